"""CPU suite: the C-ABI library loads and exports every symbol include/pmvs_b200.h declares (no compute
call without a GPU), the product path fails loudly without CUDA, and the multi-rank sharding plumbing
(world_size 2, gloo) partitions a frontier without loss or overlap."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.dirname(os.path.abspath(__file__))


def test_header_symbols_exported(pkg):
    hdr = open(os.path.join(ROOT, "include", "pmvs_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(pmvsb_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations found"
    import __graft_entry__ as g
    g.build()
    lib = ctypes.CDLL(pkg.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), "missing export: " + name
    assert sorted(pkg.SYMBOLS) == declared, "binding.SYMBOLS and the header disagree"
    lib.pmvsb_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.pmvsb_version()


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.PmvsError):
        pkg.PmvsB200(4)


def test_product_never_imports_oracle():
    """The product tree (cmvs-pmvs_b200/, include/) must not reference oracle/ in any way."""
    bad = []
    for base in ("cmvs-pmvs_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"(from|import)\s+oracle|pmvs_oracle|libpmvs_ref|oracle/", txt.replace("oracle/nm3.h is the written definition", "").replace("oracle/nm3.h", "")):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_shard_bounds(pkg):
    from cmvs_pmvs_b200.sharding import shard_bounds
    for n in (0, 1, 7, 8, 1000, 1 << 20):
        for world in (1, 2, 3, 8):
            cuts = [shard_bounds(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[r][1] == cuts[r + 1][0] for r in range(world - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1


WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
import __graft_entry__ as g
g.load_package()
from cmvs_pmvs_b200.sharding import shard_bounds, allgather_records
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
n = 1001
lo, hi = shard_bounds(n, rank, world)
# every rank "refines" its shard: record = (global index, 2*index), ok flag on odd indexes
idx = torch.arange(lo, hi, dtype=torch.float32)
rec = torch.stack([idx, 2 * idx, (idx %% 2)], dim=1)
allrec = allgather_records(rec, n, world)
assert allrec.shape == (n, 3)
assert torch.equal(allrec[:, 0], torch.arange(n, dtype=torch.float32))
assert torch.equal(allrec[:, 1], 2 * torch.arange(n, dtype=torch.float32))
# identical on every rank: every rank can apply the same deterministic commit
chk = torch.tensor([float(allrec.sum())]); ref = chk.clone(); dist.broadcast(ref, 0)
assert torch.equal(chk, ref)
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_allgather_world2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29617")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert "rank %d ok" % r in o


def test_genoption_matches_reference(tmp_path):
    """bin/genOption writes, byte for byte, the option-%04d files and pmvs.sh that the reference's genOption wrote for the
    same ske.dat (tests/golden/pmvs_clusters.npz, made by oracle/_ref/genOption_ref = source/genOption.cpp)."""
    import subprocess
    from scene_util import SKE_TWO_CLUSTERS
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pmvs_clusters.npz"))
    exe = os.path.join(ROOT, "cmvs-pmvs_b200", "bin", "genOption")
    assert os.path.exists(exe), "genOption not built: run __graft_entry__.build()"
    prefix = str(tmp_path) + "/"
    open(prefix + "ske.dat", "w").write(SKE_TWO_CLUSTERS)
    subprocess.run([exe, prefix] + [str(a) for a in G["args"]], check=True)
    for name in ("option-0000", "option-0001", "pmvs.sh"):
        assert open(prefix + name, "rb").read() == bytes(G["file_" + name]), name
    # defaults (level 1 csize 2 threshold 0.7 wsize 7 minImageNum 3 CPU 8) and the usage / missing-file exits
    subprocess.run([exe, prefix], check=True)
    txt = open(prefix + "option-0001").read()
    assert "threshold 0.7\n" in txt and "CPU 8\n" in txt and "timages 8 8 9 10 11 12 13 14 15 \n" in txt and "oimages 2 0 7 \n" in txt
    assert subprocess.run([exe], stderr=subprocess.PIPE).returncode == 1
    assert subprocess.run([exe, prefix + "nowhere/"], stderr=subprocess.PIPE).returncode == 1


def test_clusters_runner_schedules_and_merges(tmp_path):
    """bin/pmvs2_clusters without a GPU: a stand-in `pmvs2` next to a copy of the runner records its CUDA_VISIBLE_DEVICES and
    writes small models; the runner must start every option-%04d exactly once, never two clusters on one GPU slot at a time,
    and concatenate the models (counts add up, bodies in cluster order).  The real thing runs in tests/test_gpu_pipeline.py."""
    import shutil
    import stat
    import subprocess
    src = os.path.join(ROOT, "cmvs-pmvs_b200", "bin", "pmvs2_clusters")
    assert os.path.exists(src), "pmvs2_clusters not built: run __graft_entry__.build()"
    bindir = tmp_path / "bin"
    bindir.mkdir()
    shutil.copy2(src, bindir / "pmvs2_clusters")
    fake = bindir / "pmvs2"
    fake.write_text("""#!/bin/bash
# stand-in for pmvs2: prefix option [PATCH] [PSET]
prefix=$1; opt=$2; c=$((10#${opt#option-})); n=$((c + 2))
echo "$opt $CUDA_VISIBLE_DEVICES start $(date +%s%N)" >> ${prefix}trace.txt
sleep 0.3
{ printf 'ply\\nformat ascii 1.0\\nelement vertex %d\\nproperty float x\\nend_header\\n' $n; for i in $(seq $n); do echo "$c $i 0 0 0 1 1 2 3 0.9"; done; } > ${prefix}models/$opt.ply
{ printf 'PATCHES\\n%d\\n' $n; for i in $(seq $n); do printf 'PATCHS\\n%d %d 0 1\\n0 0 1 0\\n0.9 1 1\\n3\\n0 1 2 \\n0\\n\\n\\n' $c $i; done; } > ${prefix}models/$opt.patch
for i in $(seq $n); do echo "$c $i 0 0 0 1"; done > ${prefix}models/$opt.pset
echo "$opt $CUDA_VISIBLE_DEVICES end $(date +%s%N)" >> ${prefix}trace.txt
[ "$opt" != "option-0099" ]
""")
    fake.chmod(fake.stat().st_mode | stat.S_IEXEC)
    prefix = str(tmp_path / "scene") + "/"
    os.makedirs(prefix)
    K = 5
    for c in range(K):
        open(prefix + "option-%04d" % c, "w").write("level 1\n")
    env = dict(os.environ)
    env.pop("CUDA_VISIBLE_DEVICES", None)
    p = subprocess.run([str(bindir / "pmvs2_clusters"), prefix, "--gpus", "2", "PATCH", "PSET"], stderr=subprocess.PIPE, text=True, env=env, timeout=120)
    assert p.returncode == 0, p.stderr
    trace = [l.split() for l in open(prefix + "trace.txt")]
    starts = {t[0]: (t[1], int(t[3])) for t in trace if t[2] == "start"}
    ends = {t[0]: int(t[3]) for t in trace if t[2] == "end"}
    assert sorted(starts) == ["option-%04d" % c for c in range(K)]
    assert {d for d, _ in starts.values()} == {"0", "1"}
    for a in starts:                       # two clusters on the same device never overlap in time
        for b in starts:
            if a < b and starts[a][0] == starts[b][0]:
                assert ends[a] <= starts[b][1] or ends[b] <= starts[a][1], (a, b)
    total = sum(c + 2 for c in range(K))
    pset = open(prefix + "models/option-all.pset").read().split("\n")[:-1]
    assert len(pset) == total and pset[0] == "0 1 0 0 0 1" and pset[-1] == "%d %d 0 0 0 1" % (K - 1, K + 1)
    ply = open(prefix + "models/option-all.ply").read().split("\n")
    assert ply[2] == "element vertex %d" % total and "property float quality" in ply and len(ply) - 1 - (ply.index("end_header") + 1) == total
    patch = open(prefix + "models/option-all.patch").read()
    assert patch.startswith("PATCHES\n%d\n" % total) and patch.count("PATCHS") == total
    # error paths: unknown argument, no option files, a failing cluster
    assert subprocess.run([str(bindir / "pmvs2_clusters"), prefix, "--bogus"], stderr=subprocess.PIPE).returncode == 1
    assert subprocess.run([str(bindir / "pmvs2_clusters"), str(tmp_path) + "/empty/", "--gpus", "1"], stderr=subprocess.PIPE).returncode == 1
    shutil.rmtree(prefix + "models")
    for c in range(K, 100):
        open(prefix + "option-%04d" % c, "w").write("level 1\n")
    q = subprocess.run([str(bindir / "pmvs2_clusters"), prefix, "--gpus", "16", "--no-merge"], stderr=subprocess.PIPE, text=True, env=env, timeout=300)
    assert q.returncode == 1 and "cluster 99 FAILED" in q.stderr


def test_reference_arm_line_contract():
    """`bench.py --impl reference` (the arm the driver times beside ours): rank 0 prints ONE JSON line with the metric,
    config and unit of our arm, `impl`, a `cpu_baseline` describing the run and an `e2e` without copies; any other rank
    exits 0 without output.  Tiny scene so that it runs on the CPU suite's budget."""
    import json
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--patches", "512", "--steps", "2", "--warmup", "1",
           "--no-pipeline", "--views", "8", "--width", "320", "--height", "240"]
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "refined_patches_per_sec" and d["unit"] == "patches/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, timeout=600)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_cell_rules_match_the_reference():
    """CExpand::checkCounts / updateCounts: the code pmvs2 runs (cmvs-pmvs_b200/host/cell_rules.hpp, through lib/libpmvs_host.so)
    against the reference's own answers on its own _pgrids with pseudo-random trial counters (tests/golden/pmvs_state.npz):
    3 000 candidates, both depth regimes, then updateCounts applied in order -- return values and every counter equal."""
    import ctypes as C
    import __graft_entry__ as g
    g.build()
    lib = C.CDLL(os.path.join(ROOT, "cmvs-pmvs_b200", "lib", "libpmvs_host.so"))
    S = np.load(os.path.join(HERE, "golden", "pmvs_state.npz"))
    from scene_util import small_scene
    scene = small_scene()
    assert scene.sha256() == bytes(S["scene_sha256"]).hex()
    tn = scene.num
    lvl, cs = scene.option["level"], scene.option["csize"]
    w, h = scene.width >> lvl, scene.height >> lvl
    gw = np.full(tn, (w + cs - 1) // cs, np.int32); gh = np.full(tn, (h + cs - 1) // cs, np.int32)
    base = np.concatenate([[0], np.cumsum(gw * gh)]).astype(np.int32)
    assert base[-1] == len(S["cr_occ"])
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    occ = np.ascontiguousarray(S["cr_occ"], np.int32)
    off = np.ascontiguousarray(S["cr_off"], np.int32); im = np.ascontiguousarray(S["cr_images"], np.int32); gr = np.ascontiguousarray(S["cr_grids"], np.int32)
    voff = np.ascontiguousarray(S["cr_voff"], np.int32); vim = np.ascontiguousarray(S["cr_vimages"], np.int32); vgr = np.ascontiguousarray(S["cr_vgrids"], np.int32)
    P = len(off) - 1
    for depth, thr1 in ((1, 4), (2, 2)):
        counts = np.ascontiguousarray(S["cr_counts0"], np.uint8).copy()
        v = np.zeros(P, np.uint8)
        lib.pmvsh_check_counts_batch(tn, vp(gw), vp(gh), vp(base), vp(occ), vp(counts), P, vp(off), vp(im), vp(gr), thr1, 3, depth, vp(v))
        assert np.array_equal(v, S["cr_check_d%d" % depth]), depth
        assert 0.01 < v.mean() < 0.99      # both verdicts occur
    counts = np.ascontiguousarray(S["cr_counts0"], np.uint8).copy()
    rq = np.zeros(P, np.uint8)
    lib.pmvsh_update_counts_batch(tn, vp(gw), vp(gh), vp(base), vp(occ), vp(counts), P, vp(off), vp(im), vp(gr), vp(voff), vp(vim), vp(vgr), 2, vp(rq))
    assert np.array_equal(rq, S["cr_requeue"])
    assert np.array_equal(counts, S["cr_counts1"]) and (counts != S["cr_counts0"]).sum() > 1000
