"""Run the reference binary (oracle/_ref/pmvs3_ref, CPU) and the drop-in (cmvs-pmvs_b200/bin/pmvs2, GPU) on the same
synthetic scene directory and compare the clouds: patch count, accuracy against the known surface, and the mean
nearest-neighbour distance between the two clouds.   usage: python tools/compare_pipeline.py [sphere16|small16|ring47|dtu48] [--cpu N]"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g


def nn_dist(a, b, chunk=2048):
    """mean distance from each point of a to its nearest neighbour in b"""
    import torch
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    A = torch.from_numpy(a).to(dev); B = torch.from_numpy(b).to(dev)
    out = []
    for i in range(0, len(A), chunk):
        out.append(torch.cdist(A[i:i + chunk], B).min(dim=1).values)
    return float(torch.cat(out).mean())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("scene", nargs="?", default="sphere16")
    ap.add_argument("--cpu", type=int, default=1, help="CPU option for the reference run (1 = deterministic)")
    ap.add_argument("--skip-ref", action="store_true")
    ap.add_argument("--ranks", type=int, default=1, help="pmvs2 processes (one per GPU, launched with torch.distributed.run --no-python)")
    a = ap.parse_args()
    import torch
    synth = g.load_package().synth
    if a.scene == "sphere16":
        scene = synth.sphere_scene()
    elif a.scene == "small16":
        scene = synth.sphere_scene(views=16, width=320, height=240)
    elif a.scene == "ring47":
        scene = synth.ring_scene()
    elif a.scene == "dtu48":
        scene = synth.dtu_scene()
    else:
        raise SystemExit("unknown scene")
    synth.render(scene, device="cuda" if torch.cuda.is_available() else "cpu")
    res = {"scene": scene.name, "width": scene.width, "height": scene.height, "views": scene.num}
    runs = {}
    if not a.skip_ref:
        scene.option["CPU"] = a.cpu
        pr = synth.write_scene(scene, "/tmp/cmp_ref_%s" % a.scene)
        t = time.time()
        pref = subprocess.run([os.path.join(ROOT, "oracle/_ref/pmvs3_ref"), pr, "option.txt", "PSET"], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, check=True)
        if os.path.isdir(os.path.join(ROOT, "gpurun_out")):
            open(os.path.join(ROOT, "gpurun_out", "pmvs3_ref_%s_cpu%d.log" % (a.scene, a.cpu)), "w").write(pref.stderr[-20000:])
        runs["reference_cpu%d" % a.cpu] = (time.time() - t, np.loadtxt(pr + "models/option.txt.pset", dtype=np.float32).reshape(-1, 6))
    scene.option["CPU"] = os.cpu_count() or 4
    pg = synth.write_scene(scene, "/tmp/cmp_gpu_%s" % a.scene)
    t = time.time()
    cmd = [os.path.join(ROOT, "cmvs-pmvs_b200/bin/pmvs2"), pg, "option.txt", "PATCH", "PSET"]
    if a.ranks > 1:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(a.ranks), "--master-addr", "127.0.0.1",
               "--master-port", "29741", "--no-python"] + cmd
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    if p.returncode != 0:
        print(p.stderr[-3000:])
        raise SystemExit("pmvs2 failed with %d" % p.returncode)
    runs["pmvs_b200"] = (time.time() - t, np.loadtxt(pg + "models/option.txt.pset", dtype=np.float32).reshape(-1, 6))
    open(os.path.join(ROOT, "gpurun_out", "pmvs2_%s%s.log" % (a.scene, "" if a.ranks == 1 else "_x%d" % a.ranks)), "w").write(p.stderr) if os.path.isdir(os.path.join(ROOT, "gpurun_out")) else None
    for k, (secs, pts) in runs.items():
        r = {"seconds": secs, "patches": int(len(pts))}
        if scene.kind == "sphere" and len(pts):
            rad = np.linalg.norm(pts[:, :3], axis=1)
            r["mean_abs_radius_error"] = float(np.abs(rad - 1).mean())
            al = (pts[:, 3:] * pts[:, :3] / rad[:, None]).sum(1)
            r["normal_alignment"] = float(al.mean())
            r["frac_normal_off_by_more_than_10deg"] = float((al < 0.9848).mean())
            r["frac_radius_error_above_0.005"] = float((np.abs(rad - 1) > 0.005).mean())
        res[k] = r
    if len(runs) == 2 and all(len(v[1]) for v in runs.values()):
        ra, rb = runs[list(runs)[0]][1], runs["pmvs_b200"][1]
        res["mean_nn_distance_gpu_to_ref"] = nn_dist(rb[:, :3], ra[:, :3])
        res["mean_nn_distance_ref_to_gpu"] = nn_dist(ra[:, :3], rb[:, :3])
    print(json.dumps(res))


if __name__ == "__main__":
    main()
