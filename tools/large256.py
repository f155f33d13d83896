"""BASELINE.json configs[3], "Large-256": 8 disjoint object groups x 32 views of 4000x3000, laid out like a
CMVS output directory -- visualize/%08d.ppm + txt/%08d.txt with global image numbers, vis.dat (each image sees its own group),
ske.dat with one cluster per group -- then
    bin/genOption prefix 1 2 0.7 7 3 <cpu>      ->  option-0000 ... option-0007 (the reference's genOption format)
    bin/pmvs2_clusters prefix --gpus G PSET     ->  one pmvs2 per cluster per GPU ("replicas only": no collective), merged models
timed for every G in --gpus (default 1 and all visible GPUs).  Each group is its own relief object (synth.dtu_scene with a
different seed) seen only by its own cameras, which is what "disjoint" means for the reconstruction: no patch of one cluster
is visible in another's images.  Groups are rendered on the visible GPUs in turn.
--reference C runs the reference binary (oracle/_ref/pmvs3_ref, all host threads) on cluster C's option file as the CPU figure.
usage: python tools/large256.py [--clusters 8] [--views 32] [--width 4000] [--height 3000] [--gpus 1,2,4,8] [--reference 0] [--out FILE]"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g


def render_clusters(clusters, V, width, height, prefix, device):
    """renders the groups `clusters` (each its own relief object, seed 3 + c) and writes their images and cameras; returns the camera centres"""
    synth = g.load_package().synth
    centres = {}
    for c in clusters:
        scene = synth.dtu_scene(views=V, width=width, height=height, seed=3 + c)
        synth.render(scene, device=device, rows_per_chunk=200)
        for i, im in enumerate(scene.images):
            gid = c * V + i
            with open(prefix + "visualize/%08d.ppm" % gid, "wb") as f:
                f.write(b"P6\n%d %d\n255\n" % (scene.width, scene.height))
                f.write(np.ascontiguousarray(im).tobytes())
            with open(prefix + "txt/%08d.txt" % gid, "w") as f:
                f.write("CONTOUR\n")
                for r in range(3):
                    f.write(" ".join("%.9g" % float(v) for v in scene.P[i, r]) + "\n")
        centres[c] = np.asarray(scene.C)[:, :3]
        scene.images = None
    return centres


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clusters", type=int, default=8)
    ap.add_argument("--views", type=int, default=32)
    ap.add_argument("--width", type=int, default=4000)
    ap.add_argument("--height", type=int, default=3000)
    ap.add_argument("--gpus", default="")
    ap.add_argument("--only", default="", help="comma-separated clusters to generate and run (default: all)")
    ap.add_argument("--reference", type=int, default=-1, help="also time the reference binary on this cluster")
    ap.add_argument("--prefix", default="/tmp/pmvs_large256")
    ap.add_argument("--out", default="")
    ap.add_argument("--settle", type=float, default=0.0, help="seconds to wait after the scene has been rendered, so that the timed runs do not start inside the renderers' teardown")
    ap.add_argument("--logs", default="", help="directory that receives the per-cluster pmvs2 logs (phase clocks) of the last run")
    a = ap.parse_args()
    import torch
    ngpu = torch.cuda.device_count()
    synth = g.load_package().synth
    cpu = os.cpu_count() or 4
    K, V = a.clusters, a.views
    only = [int(c) for c in a.only.split(",") if c != ""] or list(range(K))
    prefix = a.prefix if a.prefix.endswith("/") else a.prefix + "/"
    shutil.rmtree(prefix, ignore_errors=True)
    for d in ("visualize", "txt", "models"):
        os.makedirs(prefix + d, exist_ok=True)
    res = {"workload": "Large-256 (BASELINE configs[3]): %d clusters x %d views %dx%d, level 1 csize 2 wsize 7 minImageNum 3" % (K, V, a.width, a.height),
           "clusters": K, "views_per_cluster": V, "size": [a.width, a.height], "host_threads": cpu, "visible_gpus": ngpu, "runs": []}
    t0 = time.time()
    centres = {}
    workers = max(1, min(ngpu, len(only)))
    if workers > 1:      # one rendering process per GPU, each takes every `workers`-th cluster
        import multiprocessing as mp
        ctx = mp.get_context("spawn")
        with ctx.Pool(workers) as pool:
            for part in pool.starmap(render_clusters, [(only[w::workers], V, a.width, a.height, prefix, "cuda:%d" % w) for w in range(workers)]):
                centres.update(part)
    else:
        centres.update(render_clusters(only, V, a.width, a.height, prefix, "cuda:0" if ngpu else "cpu"))
    res["generate_seconds"] = time.time() - t0
    n = K * V
    with open(prefix + "vis.dat", "w") as f:        # an image sees the 12 nearest cameras of its own group
        f.write("VISDATA\n%d\n" % n)
        for gid in range(n):
            c, i = divmod(gid, V)
            if c in centres:
                d = np.linalg.norm(centres[c] - centres[c][i], axis=1)
                nb = sorted(int(c * V + j) for j in np.argsort(d)[1:13])
            else:
                nb = []
            f.write("%d %d  %s\n" % (gid, len(nb), " ".join(map(str, nb))))
    with open(prefix + "ske.dat", "w") as f:        # CMVS::CBundle::writeGroups' format (source/cmvs/bundle.cpp:1462-1481)
        f.write("SKE\n%d %d\n" % (n, K))
        for c in range(K):
            t = list(range(c * V, (c + 1) * V))
            f.write("%d 0\n%s \n\n" % (len(t), " ".join(map(str, t))))
    BIN = os.path.join(ROOT, "cmvs-pmvs_b200", "bin")
    subprocess.run([os.path.join(BIN, "genOption"), prefix, "1", "2", "0.7", "7", "3", str(cpu)], check=True, stdout=subprocess.DEVNULL)
    for c in range(K):                              # clusters that were not generated are not run
        if c not in only:
            os.remove(prefix + "option-%04d" % c)
    gpus = [int(x) for x in a.gpus.split(",") if x] or sorted({1, max(1, ngpu)})
    if a.settle > 0:
        time.sleep(a.settle)
    res["settle_seconds"] = a.settle
    for G in gpus:
        if G > max(1, ngpu):
            continue
        t = time.time()
        p = subprocess.run([os.path.join(BIN, "pmvs2_clusters"), prefix, "--gpus", str(G), "PSET"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        secs = time.time() - t
        if p.returncode != 0:
            print(p.stderr[-4000:], file=sys.stderr)
            raise SystemExit("pmvs2_clusters failed")
        counts = {c: sum(1 for _ in open(prefix + "models/option-%04d.pset" % c)) for c in only}
        res["runs"].append({"gpus": G, "seconds": secs, "patches_per_cluster": counts, "merged": sum(1 for _ in open(prefix + "models/option-all.pset")),
                            "log": [l for l in p.stderr.splitlines() if l.startswith(("cluster", "merged"))]})
        print("G=%d: %.1f s, %d patches" % (G, secs, sum(counts.values())), file=sys.stderr, flush=True)
    if a.reference >= 0 and os.path.exists(os.path.join(ROOT, "oracle/_ref/pmvs3_ref")):
        opt = "option-%04d" % a.reference
        t = time.time()
        p = subprocess.run([os.path.join(ROOT, "oracle/_ref/pmvs3_ref"), prefix, opt, "PSET"], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
        secs = time.time() - t
        res["reference"] = {"cluster": a.reference, "binary": "oracle/_ref/pmvs3_ref", "host_threads": cpu, "seconds": secs, "returncode": p.returncode,
                            "patches": sum(1 for _ in open(prefix + "models/%s.pset" % opt)) if p.returncode == 0 else None}
    text = json.dumps(res)
    if a.out:
        with open(a.out, "w") as f:
            f.write(text + "\n")
    print(text)
    if a.logs:
        os.makedirs(a.logs, exist_ok=True)
        for c in only:
            src = prefix + "models/option-%04d.log" % c
            if os.path.exists(src):
                shutil.copy(src, os.path.join(a.logs, "large256_option-%04d.log" % c))
    shutil.rmtree(prefix, ignore_errors=True)


if __name__ == "__main__":
    main()
