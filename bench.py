#!/usr/bin/env python
"""bench.py -- refined patches/sec of the PMVS patch-optimisation path on B200 (BASELINE.json metric).

Workload (BASELINE.json configs[4] on the configs[2] scene): DTU-shaped synthetic scene,
48 calibrated views 1600x1200, pmvs level 1 / csize 2 / wsize 7 / minImageNum 3; per GPU 1 048 576 seed
patches x 5 visible views (surface point displaced along the reference ray, normal rotated up to 20 degrees);
one step = refinePatch + final computeINCC for every patch of the batch in ONE kernel launch.

  python bench.py [--gpus N --steps K --warmup W]          the CUDA path (this repository)
  python bench.py --impl reference [...]                   the reference's own CPU code (oracle/_ref)

`value`   : whole-job patches/s with inputs resident in HBM (CUDA events around the K steps, max over ranks)
`e2e`     : the same through pmvsb_refine_batch with pinned HOST buffers (H2D + kernel + D2H inside the timing)
`roofline`: algorithmic gather bytes 588*V*(E+1) per patch / kernel time vs measured HBM peak
N > 1     : patches are sharded across ranks (weak scaling: per-GPU batch fixed), images replicated, and the
            per-wave exchange of refined patch records is an NCCL all-gather inside the step.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "refined_patches_per_sec"
UNIT = "patches/s"
VIEWS = 5
BYTES_PER_VIEW_EVAL = 588  # 49 bilinear samples x 4 texels x 3 B (include/image/image.hpp:467-475)


# --------------------------------------------------------------------------------------------------------
def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--patches", type=int, default=1 << 20, help="seed patches per GPU per step")
    ap.add_argument("--views", type=int, default=48)
    ap.add_argument("--width", type=int, default=1600)
    ap.add_argument("--height", type=int, default=1200)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU-baseline budget")
    ap.add_argument("--no-pipeline", action="store_true", help="skip the end-to-end pmvs2 wall-time measurement")
    ap.add_argument("--pipeline-stage", default="", help=argparse.SUPPRESS)   # internal: second stage of the b200 arm, see main()
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """SM clock and throttle reasons sampled WHILE the timed region runs: NVML in a thread of this process every 20 ms (a sample
    costs microseconds and launches nothing), or -- where NVML cannot be loaded -- a streaming `nvidia-smi -lms 200`."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.rows = []          # nvidia-smi rows
        self.proc = None
        self.nv = None          # (module, handle)
        self.sm, self.bits, self.mx = [], 0, None
        self.stop_flag = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            handle = None
            try:      # the CUDA ordinal is not the NVML index under CUDA_VISIBLE_DEVICES: go through the UUID
                import torch
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode() if not uuid.startswith("GPU-") else uuid.encode())
            except Exception:
                handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM)
            self.nv = (pynvml, handle)
        except Exception:
            self.nv = None

    def start(self):
        if self.nv:
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv, h = self.nv
        try:
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        except Exception:
            self.mx = None
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.bits |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
            except Exception:
                pass
            self.stop_flag.wait(0.02)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nv:
            self.stop_flag.set()
            self.t.join(timeout=2)
            nv = self.nv[0]
            names = [("hw_slowdown", nv.nvmlClocksThrottleReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                     ("sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksThrottleReasonSwPowerCap)]
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.mx, "samples": len(self.sm),
                    "reasons": sorted(n for n, bit in names if self.bits & bit), "source": "nvml, every 20 ms inside the timed region"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for k, nme in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi -lms 200"}


# --------------------------------------------------------------------------------------------------------
# workload
# --------------------------------------------------------------------------------------------------------
def build_scene(args, device):
    import __graft_entry__ as g
    pkg = g.load_package()
    synth = pkg.synth
    scene = synth.dtu_scene(views=args.views, width=args.width, height=args.height, seed=2)
    synth.render(scene, device=device, rows_per_chunk=400)
    return pkg, scene


def make_seed_patches(scene, gpu_lib, n, seed, device):
    """Config-5 seed patches: exactly VIEWS views per patch picked with sortImages' greedy rule
    (/root/reference/source/pmvs/optim.cpp:284-321), depth noise N(0,(2 dscale)^2) along the reference ray,
    normal rotated by U(0,20 deg).  Returns host numpy arrays."""
    import torch
    import __graft_entry__ as g
    synth = g.load_package().synth
    pts, nrm = synth.surface_samples(scene, n, seed, device=device)
    gen = torch.Generator(device="cpu").manual_seed(seed + 17)
    C = torch.tensor(scene.C, dtype=torch.float64, device=device)                      # (V,3)
    rays = C[None, :, :] - pts[:, None, :]                                              # (n,V,3)
    dist = rays.norm(dim=2)
    rays = rays / dist[:, :, None]
    dots = (rays * nrm[:, None, :]).sum(2)
    units = torch.where(dots > 0.2, dist / dots.clamp(min=1e-6), torch.full_like(dist, 1e30))
    t = 1.0 - math.cos(math.radians(10.0))
    ref = torch.argmin(units, dim=1)
    units.scatter_(1, ref[:, None], 0.0)
    chosen = []
    for _ in range(VIEWS):
        sel = torch.argmin(units, dim=1)
        chosen.append(sel)
        rs = torch.gather(rays, 1, sel[:, None, None].expand(-1, 1, 3))                # (n,1,3)
        f = (1.0 - (rays * rs).sum(2)).clamp(min=t / 2.0, max=t)
        units = units * (t / f)
        units.scatter_(1, sel[:, None], 1e30)
    images = torch.stack(chosen, dim=1).to(torch.int32).cpu().numpy()
    coords = np.ones((n, 4), np.float32)
    coords[:, :3] = pts.cpu().numpy()
    dsc, _ = gpu_lib.set_scales_batch(coords, images)
    # displace along the reference ray, rotate the normal
    refC = scene.C[images[:, 0]]
    ray = coords[:, :3].astype(np.float64) - refC
    ray /= np.linalg.norm(ray, axis=1, keepdims=True)
    noise = torch.randn(n, generator=gen, dtype=torch.float64).numpy() * 2.0 * dsc
    coords[:, :3] = (coords[:, :3].astype(np.float64) + ray * noise[:, None]).astype(np.float32)
    ang = torch.rand(n, generator=gen, dtype=torch.float64).numpy() * math.radians(20.0)
    axis = torch.randn(n, 3, generator=gen, dtype=torch.float64).numpy()
    nn = nrm.cpu().numpy()
    axis -= (axis * nn).sum(1, keepdims=True) * nn
    axis /= np.linalg.norm(axis, axis=1, keepdims=True)
    rot = nn * np.cos(ang)[:, None] + axis * np.sin(ang)[:, None]
    normals = np.zeros((n, 4), np.float32)
    normals[:, :3] = rot
    dsc, _ = gpu_lib.set_scales_batch(coords, images)
    return coords, normals, images, dsc.astype(np.float32)


# --------------------------------------------------------------------------------------------------------
# second half of BASELINE.json's metric: wall time of the whole pmvs2 run on the workload's scene
# --------------------------------------------------------------------------------------------------------
def pipeline_wall_time(scene, impl, ranks=1):
    """Runs the drop-in binary (impl 'b200'; ranks > 1: one pmvs2 process per GPU under torch.distributed.run, frontier shards +
    exchange of the accepted candidates per wave over peer memory) or the reference binary built from the reference's own sources (impl
    'reference', all host threads) on the scene written to disk as PPM + txt + option file; returns a dict."""
    cores = os.cpu_count() or 1
    prefix = write_scene_for_reference(scene, cores, tag="pipeline")   # its own directory: RefLib leaves feature-cache stubs in models/
    return pipeline_wall_time_at(prefix, scene_description(scene), impl, ranks)


def scene_description(scene):
    return "%s %d views %dx%d level %d csize %d" % (scene.name, scene.num, scene.width, scene.height, scene.option["level"], scene.option["csize"])


def pipeline_wall_time_at(prefix, description, impl, ranks=1, keep=False):
    """pipeline_wall_time for a scene that is already on disk under `prefix` (removed afterwards unless keep)."""
    import subprocess
    cores = os.cpu_count() or 1
    exe = os.path.join(ROOT, "cmvs-pmvs_b200", "bin", "pmvs2") if impl == "b200" else os.path.join(ROOT, "oracle", "_ref", "pmvs3_ref")
    if not os.path.exists(exe):
        return {"scene": description, "unavailable": os.path.relpath(exe, ROOT) + " not built"}
    cmd = [exe, prefix, "option.txt", "PSET"]
    env = dict(os.environ)
    if ranks > 1:
        port = int(os.environ.get("MASTER_PORT", "29500")) + 211
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(ranks), "--master-addr", "127.0.0.1", "--master-port", str(port),
               "--no-python"] + cmd
        for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT", "TORCHELASTIC_RUN_ID", "GROUP_RANK", "ROLE_RANK", "LOCAL_WORLD_SIZE", "ROLE_WORLD_SIZE"):
            env.pop(k, None)
    t0 = time.perf_counter()
    p = subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, env=env)
    secs = time.perf_counter() - t0
    out = {"scene": description,
           "binary": os.path.relpath(exe, ROOT), "ranks": ranks, "wall_seconds": secs, "host_threads": cores, "returncode": p.returncode}
    if ranks > 1:
        out["note"] = "wall_seconds includes the launcher (torch.distributed.run start-up, ~1-2 s) and N CUDA contexts coming up at once; phases_seconds.main.total is pmvs2's own clock on rank 0"
        for l in p.stderr.splitlines():
            if "exchange " in l and " ranks" in l:
                out["exchange"] = l[l.index("exchange "):].strip()
    try:
        with open(prefix + "models/option.txt.pset") as f:
            out["patches"] = sum(1 for _ in f)
    except OSError:
        out["patches"] = None
    refined = []   # the count lines are all-integer; the reference also prints a percentage line (floats, nan when a round is empty)
    for l in p.stderr.splitlines():
        if l.startswith("Total pass fail0 fail1 refinepatch:"):
            tok = l.split(":")[1].split()
            if len(tok) == 5 and all(t.isdigit() for t in tok) and not (tok[0] == "100" and int(tok[1]) + int(tok[2]) + int(tok[3]) != 100):
                refined.append(int(tok[4]))
    phases = {}   # pmvs2's own per-phase clocks ("time <name> <seconds> s" on stderr)
    for l in p.stderr.splitlines():
        tok = l.split()
        if len(tok) == 4 and tok[0] == "time" and tok[3] == "s":
            try:
                phases[tok[1]] = round(float(tok[2]), 4)
            except ValueError:
                pass
    if phases:
        out["phases_seconds"] = phases
    if refined:
        out["refined_patches"] = int(sum(refined))                      # the reference's own "refinepatch" counter
        out["refined_patches_per_sec_whole_run"] = out["refined_patches"] / secs
    if not keep:
        import shutil
        shutil.rmtree(prefix, ignore_errors=True)
    return out


def measured_traffic(patches):
    """dram__bytes_read.sum + dram__bytes_write.sum of one k_refine launch from the committed `ncu --set full` capture
    (profiles/r2_k_refine_dram.json); only valid for the launch size it was captured on."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_k_refine_dram.json")) as f:
            d = json.load(f)
        return float(d["dram_bytes_per_launch"]) if int(d["patches_per_launch"]) == int(patches) else None
    except Exception:
        return None


# --------------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU code (oracle/_ref), bounded sample per step
# --------------------------------------------------------------------------------------------------------
def write_scene_for_reference(scene, cpu_threads, tag="scene"):
    import __graft_entry__ as g
    synth = g.load_package().synth
    prefix = "/tmp/pmvs_b200_bench_%s_%d/" % (tag, os.getpid())
    scene.option = dict(scene.option)
    scene.option["CPU"] = cpu_threads
    synth.write_scene(scene, prefix)
    return prefix


def reference_rate(scene, patches, seconds, steps=1, warmup=0, full=None, ref_xtol_seconds=0.0):
    """Times COptim::refinePatch of the reference build over a bounded sample on all host cores.
    Returns dict(value, cores, kind, sample, ms_per_step, evals_per_patch).  ref_xtol_seconds > 0 adds `at_reference_xtol`:
    the same loop with the stand-in optimiser's tolerance floor removed, i.e. at the reference's own xtol_rel 1e-7."""
    from oracle import bindings as ob
    cores = os.cpu_count() or 1
    coords, normals, images, dsc = patches
    kind = "reference"
    if os.path.exists(ob.REF_SO):
        prefix = write_scene_for_reference(scene, cores)
        lib = ob.RefLib(prefix, num=scene.num, level=scene.option["level"])
    else:  # the reference build did not travel: time the C port instead
        kind = "port"
        lib = ob.OracleLib.from_scene(scene)
    probe = min(len(coords), 2048)
    r = lib.refine_batch(coords[:probe], normals[:probe], images[:probe], dsc[:probe], threads=cores)
    rate = probe / max(r["seconds"], 1e-6)
    per_step = int(max(probe, min(len(coords), rate * seconds / max(1, steps + warmup))))
    times, evals = [], []
    for it in range(warmup + steps):
        lo = (it * per_step) % max(1, len(coords) - per_step + 1)
        r = lib.refine_batch(coords[lo:lo + per_step], normals[lo:lo + per_step], images[lo:lo + per_step], dsc[lo:lo + per_step], threads=cores)
        if it >= warmup:
            times.append(r["seconds"]); evals.append(float(r["evals"].mean()))
    total = float(sum(times))
    extra = {}
    if ref_xtol_seconds > 0 and kind == "reference":
        lib.set_xtol_floor(0.0)
        m = int(max(256, min(len(coords), rate * ref_xtol_seconds / 1.6)))
        r2 = lib.refine_batch(coords[:m], normals[:m], images[:m], dsc[:m], threads=cores)
        lib.set_xtol_floor(1.0e-3)
        extra["at_reference_xtol"] = {"value": m / max(r2["seconds"], 1e-9), "unit": UNIT, "xtol": 1.0e-7, "evals_per_patch": float(r2["evals"].mean()),
                                      "ok_fraction": float(r2["ok"].mean()), "sample": "%d patches" % m,
                                      "note": "Nelder-Mead stand-in run down to the reference's xtol_rel 1e-7 (floor 0): the f32 objective is flat below ~1e-5, "
                                              "the extra evaluations walk a plateau"}
    return dict(extra, value=per_step * len(times) / total, cores=cores, kind=kind,
                sample="%d-patch sample of the %d-patch step (same generator, seed 4) per step x %d steps, COptim::refinePatch on %d host threads" % (per_step, full or len(coords), len(times), cores),
                ms_per_step=1000.0 * total / len(times), evals_per_patch=float(np.mean(evals)), unit=UNIT)


# --------------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.pipeline_stage:
        # Second stage of the b200 arm (rank 0 only).  The measuring process has replaced itself with this one (os.execv), so its
        # CUDA context and its gigabytes of device arrays are gone and pmvs2 meets the GPU the way a user's shell would hand it
        # over: with the measuring process's context still alive on the device, context creation inside pmvs2 alone took
        # 1.6 s instead of 0.6 s (profiles/r2_bench.json against the plain runs of the same box).
        with open(args.pipeline_stage) as f:
            state = json.load(f)
        os.remove(args.pipeline_stage)
        line = state["line"]
        line["pipeline"] = pipeline_wall_time_at(state["prefix"], state["description"], "b200", ranks=int(state["ranks"]), keep=True)
        line["pipeline"]["measured_from"] = ("a fresh process image after the microbench (no other CUDA context of this job on the GPU); wall_seconds is the FIRST run, "
                                             "on a GPU that has just been released; `repeat` is the same command once more right after it")
        again = pipeline_wall_time_at(state["prefix"], state["description"], "b200", ranks=int(state["ranks"]))
        line["pipeline"]["repeat"] = {"wall_seconds": again.get("wall_seconds"), "returncode": again.get("returncode"), "patches": again.get("patches"),
                                      "main.total": again.get("phases_seconds", {}).get("main.total"),
                                      "load.create_gpu_context": again.get("phases_seconds", {}).get("load.create_gpu_context")}
        print_line(json.dumps(line))
        return
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": "patch-refinement microbench (BASELINE configs[4]) on the DTU-shaped synthetic scene (configs[2]): "
                          "%d views %dx%d, level 1 csize 2 wsize 7 minImageNum 3; %d seed patches x %d views per GPU per step, "
                          "refinePatch + computeINCC" % (args.views, args.width, args.height, args.patches, VIEWS),
              "patches_per_gpu": args.patches, "views_per_patch": VIEWS, "optimizer": "in-kernel Nelder-Mead, xtol 1e-3 (scaled units: 1 = half a pixel of image motion / 3.75 degrees), maxeval 1000; against a restatement of "
                           "BOBYQA at the reference's settings on this scene: |dncc| p99 4.4e-5, depth p99 0.010 units, normal p50 0.3 deg "
                           "(profiles/r2_optimiser_bound_dtu.json)",
              "l2": "inputs larger than L2 (RGBA pyramids of 48 views = 0.49 GB, read through a texture atlas of the same size, + patch arrays); no explicit flush",
              "parallelism": "patches sharded over %d GPU(s), images replicated, refined records all-gathered every step" % world}

    if args.impl == "reference":
        if rank != 0:
            return
        import torch
        dev = "cuda:0" if torch.cuda.is_available() else "cpu"
        pkg, scene = build_scene(args, dev)
        n = min(args.patches, 1 << 16)
        # the seed patches of this arm never touch libpmvs_b200.so: setScales comes from the CPU oracle (bit-exact with the
        # kernel's, tests/test_gpu_parity.py), so the only native code this process loads is oracle/
        from oracle import bindings as ob
        orc = ob.OracleLib.from_scene(scene)

        class _S:
            def set_scales_batch(self, c, im):
                d = np.array([orc.set_scales(c[i], im[i])[0] for i in range(len(c))], np.float32)
                return d, d
        patches = make_seed_patches(scene, _S(), n, seed=4, device=dev)
        r = reference_rate(scene, patches, seconds=max(20.0, 8.0 * (args.steps + args.warmup)), steps=args.steps, warmup=args.warmup, full=args.patches,
                           ref_xtol_seconds=6.0)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
                                 "evals_per_patch": r["evals_per_patch"], "optimizer": "reference objective (COptim::my_f) + nm3 Nelder-Mead stand-in for nlopt BOBYQA, xtol 1e-3",
                                 "at_reference_xtol": r.get("at_reference_xtol")},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        if not args.no_pipeline and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "pmvs3_ref")):
            line["pipeline"] = pipeline_wall_time(scene, "reference")
        print_line(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = "cuda:%d" % local_rank
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))

    pkg, scene = build_scene(args, dev)
    lib = pkg.PmvsB200.from_scene(scene, device=local_rank)
    P = args.patches
    coords, normals, images, dsc = make_seed_patches(scene, lib, P, seed=4 + rank, device=dev)

    # ---- resident inputs --------------------------------------------------------------------------
    stream = torch.cuda.Stream(device=dev)          # a real (non-default) stream shared by torch ops, NCCL ordering and the library
    torch.cuda.set_stream(stream)
    lib.set_stream(stream.cuda_stream)
    d_coords0 = torch.from_numpy(coords).to(dev); d_normals0 = torch.from_numpy(normals).to(dev)
    d_images = torch.from_numpy(images).to(dev); d_dsc = torch.from_numpy(dsc).to(dev)
    d_coords = torch.empty_like(d_coords0); d_normals = torch.empty_like(d_normals0)
    d_ncc = torch.empty(P, dtype=torch.float32, device=dev); d_evals = torch.empty(P, dtype=torch.int32, device=dev)
    d_ok = torch.empty(P, dtype=torch.uint8, device=dev)
    # ---- the exchange of the refined records between the GPUs (north_star: accepted patches all-gathered over NVLink) ----------
    # default "peer": one kernel after the refine kernel stores every patch's 48-byte record into this rank's slot of EVERY rank's
    # mailbox (CUDA IPC mappings, peer stores over NVLink) and raises a flag there; a stream-ordered wait for all ranks' flags
    # follows; no collective is launched.  PMVSB_BENCH_GATHER=nccl (or a box whose GPUs cannot map each other):
    # pack + one NCCL all-gather after the kernel, as in round 1.
    gather = "none"
    rec = rec_all = None
    if world > 1:
        gather = os.environ.get("PMVSB_BENCH_GATHER", "peer")
        if gather == "peer":
            ok_t = torch.ones(1, dtype=torch.int32, device=dev)
            handles = torch.zeros(world * 64, dtype=torch.uint8, device=dev)
            try:
                mine = lib.peer_export(rank, world, 2 * (P * 48 + 256))
                dist.all_gather_into_tensor(handles, torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(dev))
                lib.peer_open(bytes(handles.cpu().numpy().tobytes()))
            except Exception as exc:      # every rank must learn about it: the ranks agree on the mode below
                sys.stderr.write("bench.py: peer mailboxes unavailable on rank %d (%s)\n" % (rank, exc))
                ok_t.zero_()
            dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
            if int(ok_t.item()) == 0:
                gather = "nccl"
        if gather != "peer":
            rec = torch.empty(P, 12, dtype=torch.float32, device=dev)          # refined record exchanged per wave
            rec_all = torch.empty(world * P, 12, dtype=torch.float32, device=dev)
    last_gather = [0, 0]   # device address of the gathered records of the last step, stride between ranks in floats

    def step_resident():
        d_coords.copy_(d_coords0); d_normals.copy_(d_normals0)
        if gather == "peer":
            last_gather[0], last_gather[1] = lib.refine_batch_dev_gather(P, VIEWS, d_coords.data_ptr(), d_normals.data_ptr(), d_images.data_ptr(), 0,
                                                                         d_dsc.data_ptr(), d_ncc.data_ptr(), d_evals.data_ptr(), d_ok.data_ptr())
            return
        lib.refine_batch_dev(P, VIEWS, d_coords.data_ptr(), d_normals.data_ptr(), d_images.data_ptr(), 0, d_dsc.data_ptr(),
                             d_ncc.data_ptr(), d_evals.data_ptr(), d_ok.data_ptr())
        if world > 1:
            rec[:, 0:4] = d_coords; rec[:, 4:8] = d_normals; rec[:, 8] = d_ncc; rec[:, 9] = d_ok.to(torch.float32); rec[:, 10] = d_evals.to(torch.float32)
            rec[:, 11] = 0.0
            dist.all_gather_into_tensor(rec_all, rec)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    launches0 = lib.launch_count()
    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    barrier()
    launches_before = lib.launch_count()
    e0.record(stream)
    for _ in range(args.steps):
        step_resident()
    e1.record(stream)
    barrier()
    total_ms = e0.elapsed_time(e1)
    launches = lib.launch_count() - launches_before
    clocks = sampler.stop()
    # kernel duration of the last launch (CUDA events recorded by the library on the same stream)
    kernel_ms.append(lib.last_refine_ms())
    evals_sum = int(d_evals.to(torch.int64).sum().item())
    ok_frac = float(d_ok.to(torch.float32).mean().item())
    ncc_med = float(d_ncc[d_ok.bool()].median().item()) if ok_frac > 0 else float("nan")

    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = world * P * args.steps / (total_ms / 1000.0)

    # the fused exchange checked against a plain NCCL all-gather of the same records (outside the timing): every rank must hold
    # every rank's records of the last step, bit for bit
    gather_verified = None
    if gather == "peer":
        class _DevView:      # zero-copy torch view of the mailbox slots (device memory owned by the library)
            def __init__(self, ptr, n):
                self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}
        got = torch.as_tensor(_DevView(last_gather[0], world * last_gather[1]), device=dev).view(world, last_gather[1])[:, :P * 12].reshape(world * P, 12)
        mine = torch.cat([d_coords, d_normals, d_ncc[:, None], d_ok.to(torch.float32)[:, None], d_evals.to(torch.float32)[:, None],
                          torch.zeros(P, 1, dtype=torch.float32, device=dev)], dim=1).contiguous()
        want = torch.empty(world * P, 12, dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(want, mine)
        same = torch.tensor([1 if torch.equal(got.view(torch.int32), want.view(torch.int32)) else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        gather_verified = bool(int(same.item()))
        del got, want, mine

    # ---- end to end through the host-pointer ABI call ------------------------------------------------
    h_coords = torch.from_numpy(coords).pin_memory(); h_normals = torch.from_numpy(normals).pin_memory()
    h_images = torch.from_numpy(images).pin_memory(); h_dsc = torch.from_numpy(dsc).pin_memory()
    h_ncc = torch.empty(P, dtype=torch.float32).pin_memory(); h_evals = torch.empty(P, dtype=torch.int32).pin_memory()
    h_ok = torch.empty(P, dtype=torch.uint8).pin_memory()
    import ctypes as C
    vp = lambda tns: C.c_void_p(tns.data_ptr())

    # coords / normals are in-out arguments of the call (the refined patch replaces the seed), so every step gets its OWN pre-staged
    # pinned copy of the inputs: the timed region holds the call -- host-to-device copies, kernels, device-to-host copies -- and no
    # bench-side housekeeping (a host-side reset of 33 MB per step cost 3 ms of the 127 ms).  Beyond 64 steps the sets are reused
    # and reset in the loop.
    e2e_steps = max(1, args.steps)      # the same number of timed steps as `value`
    n_sets = min(e2e_steps + 1, 64)
    w_sets = [(h_coords.clone().pin_memory(), h_normals.clone().pin_memory()) for _ in range(n_sets)]

    def step_e2e(i):
        w_coords, w_normals = w_sets[i % n_sets]
        if i >= n_sets:
            w_coords.copy_(h_coords); w_normals.copy_(h_normals)
        r = lib.lib.pmvsb_refine_batch(lib.ctx, P, VIEWS, vp(w_coords), vp(w_normals), vp(h_images), None, vp(h_dsc), vp(h_ncc),
                                       vp(h_evals), vp(h_ok))
        if r != 0:
            raise RuntimeError(lib.lib.pmvsb_last_error(lib.ctx).decode())

    step_e2e(0)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        step_e2e(1 + i)
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * P * e2e_steps / float(t.item())
    h2d = P * (16 + 16 + 4 * VIEWS + 4)
    d2h = P * (16 + 16 + 4 + 4 + 1)

    # ---- roofline of the dominant kernel (k_refine) ----------------------------------------------------
    # PRIMARY bound: the SM's texture pipe.  The kernel's gathers are TLD4 through L1TEX (hit rates 89 % L1 / 99.7 % L2, 143 MB of
    # DRAM per launch), so HBM is not what limits it; one warp-wide TLD4 occupies the pipe of its SM for clk_per_tld4 cycles
    # whatever its active lanes (tools/probe/tex_lane_probe.cu on B200: 24.9 clk per warp-row of three one-channel gathers).
    # achieved = TLD4 issued per second by this launch, peak = what 148 pipes can retire; frac = achieved / peak.
    # SECONDARY (`hbm_formula`): SURVEY.md 8d's algorithmic bytes (588 B x views x (evaluations + 1) per patch) over the measured
    # copy bandwidth, kept for continuity with round 1.
    peak, peak_kind = peaks()
    alg_bytes = BYTES_PER_VIEW_EVAL * VIEWS * (evals_sum + P)          # 588 * V * (E + 1) summed over the launch
    k_ms = float(np.mean(kernel_ms))
    achieved = alg_bytes / (k_ms / 1000.0) / 1e9
    hbm_formula = {"achieved": achieved, "peak": peak, "peak_kind": peak_kind + " copy bandwidth", "unit": "GB/s", "frac": achieved / peak,
                   "algorithmic_bytes_per_launch": alg_bytes, "traffic": measured_traffic(P)}
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    clk_hz = 1.0e6 * float(clocks.get("sm_mhz") or 1965.0)
    clk_per_tld4 = 24.9 / 3.0
    warp_evals = (evals_sum + P) / 4.0                                   # optimiser evaluations + the final computeINCC, 4 patches per warp
    tld4 = warp_evals * VIEWS * 7 * 3                                    # views x wsize rows x three one-channel gathers
    tld4_rate = tld4 / (k_ms / 1000.0) / 1e9
    tld4_peak = sms * clk_hz / clk_per_tld4 / 1e9
    roofline = {"bound": "texture_pipe", "kernel": "k_refine_g<7, atlas>", "achieved": tld4_rate, "peak": tld4_peak, "unit": "G warp-TLD4/s",
                "frac": tld4_rate / tld4_peak, "peak_kind": "measured pipe occupancy per TLD4 (%.2f clk, profiles/r1_tex_lane_probe.txt) x %d SMs x %.0f MHz" % (clk_per_tld4, sms, clk_hz / 1e6),
                "traffic": measured_traffic(P), "kernel_ms": k_ms, "evals_per_patch": evals_sum / P, "tld4_per_launch": tld4,
                "hbm_formula": hbm_formula,
                "note": "texel gathers are served by L1TEX / L2 (DRAM traffic per launch = `traffic` bytes, 0.03 % of the algorithmic bytes), so the "
                        "binding unit is the texture pipe (ncu: 75 % busy) next to instruction issue (76 %), profiles/r2_k_refine_g_full_ncu_metrics.csv; "
                        "hbm_formula is SURVEY.md 8d's algorithmic-bytes fraction of the measured HBM copy bandwidth"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps},
            "gpu_launches": int(launches), "roofline": roofline,
            "quality": {"ok_fraction": ok_frac, "median_ncc": ncc_med}}
    if world > 1:
        line["exchange"] = {"mode": gather, "bytes_per_rank_per_step": P * 48 * world,
                            "how": ("one kernel after the refine kernel packs the 48-byte records and stores them into every rank's mailbox (CUDA IPC mappings, peer stores "
                                    "over NVLink), its last block raises the flags, stream-ordered waits (cuStreamWaitValue32); no collective launch"
                                    + ("; PMVSB_GATHER_IN_KERNEL=1: the stores are issued by the refine kernel itself" if os.environ.get("PMVSB_GATHER_IN_KERNEL") == "1" else "")) if gather == "peer"
                                   else "pack (torch) + one NCCL all_gather_into_tensor after the kernel",
                            "verified_against_nccl_allgather": gather_verified}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n = min(P, 1 << 16)
        r = reference_rate(scene, (coords[:n], normals[:n], images[:n], dsc[:n]), seconds=args.cpu_seconds, full=P, ref_xtol_seconds=5.0)
        line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
                                "evals_per_patch": r["evals_per_patch"], "optimizer": "reference objective (COptim::my_f) + nm3 Nelder-Mead stand-in for nlopt BOBYQA, xtol 1e-3",
                                "at_reference_xtol": r.get("at_reference_xtol")}
    lib.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0 and not args.no_pipeline:
        if world > 1:
            time.sleep(1.0)     # the other ranks leave their GPUs
        # the scene goes to disk now; pmvs2 is timed by a fresh image of this process (see --pipeline-stage above)
        prefix = write_scene_for_reference(scene, os.cpu_count() or 1, tag="pipeline")
        state_path = "/tmp/pmvs_b200_bench_state_%d.json" % os.getpid()
        with open(state_path, "w") as f:
            json.dump({"line": line, "prefix": prefix, "description": scene_description(scene), "ranks": world}, f)
        try:
            restore_stdout()
            sys.stdout.flush(); sys.stderr.flush()
            os.execv(sys.executable, [sys.executable, os.path.abspath(__file__), "--pipeline-stage", state_path])
        except OSError as exc:      # could not re-execute: measure from here, with this process's context alive
            sys.stderr.write("bench.py: exec failed (%s), timing pmvs2 from the measuring process\n" % exc)
            os.remove(state_path)
            line["pipeline"] = pipeline_wall_time_at(prefix, scene_description(scene), "b200", ranks=world)
    if rank == 0:
        print_line(json.dumps(line))


_real_stdout_fd = None


def restore_stdout():
    """fd 1 back to the real stdout (inheritable), for a process image that replaces this one"""
    if _real_stdout_fd is not None:
        sys.stdout.flush()
        os.dup2(_real_stdout_fd, 1)


class _StdoutToStderr:
    """stdout carries exactly ONE JSON line: anything native libraries print on fd 1 meanwhile (NCCL's version banner when
    NCCL_DEBUG is set) goes to stderr; print_line() writes the JSON line to the real stdout."""
    def __enter__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)
        global print_line, _real_stdout_fd
        _real_stdout_fd = self.real

        def print_line(text, _fd=self.real):
            os.write(_fd, (text + "\n").encode())
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.real, 1)
        os.close(self.real)
        return False


def print_line(text):
    print(text, flush=True)


if __name__ == "__main__":
    with _StdoutToStderr():
        main()
