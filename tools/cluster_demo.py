"""Config-4-shaped run at a size that fits a short GPU call: the DTU-48 scene (48 views 1600x1200) laid out like a CMVS
output directory with K clusters (ske.dat: contiguous blocks of the camera cap as target images, their two nearest cameras
of the neighbouring blocks as `oimages`; vis.dat: the 12 nearest cameras of each image), then
    bin/genOption prefix 1 2 0.7 7 3 <cpu>   ->  option-0000 ... option-(K-1), pmvs.sh
    bin/pmvs2_clusters prefix --gpus G PSET  ->  models/option-%04d.*, models/option-all.*
timed for G = 1 and G = all visible GPUs.   usage: python tools/cluster_demo.py [--clusters 4] [--views 48]"""
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clusters", type=int, default=4)
    ap.add_argument("--views", type=int, default=48)
    ap.add_argument("--width", type=int, default=1600)
    ap.add_argument("--height", type=int, default=1200)
    a = ap.parse_args()
    import torch
    synth = g.load_package().synth
    scene = synth.dtu_scene(views=a.views, width=a.width, height=a.height)
    synth.render(scene, device="cuda" if torch.cuda.is_available() else "cpu")
    cpu = os.cpu_count() or 4
    scene.option["CPU"] = cpu
    prefix = synth.write_scene(scene, "/tmp/pmvs_cluster_demo")
    n, K = scene.num, a.clusters
    C = np.asarray(scene.C)[:, :3]
    d = np.linalg.norm(C[:, None] - C[None], axis=2)
    with open(prefix + "vis.dat", "w") as f:
        f.write("VISDATA\n%d\n" % n)
        for i in range(n):
            nb = sorted(int(j) for j in np.argsort(d[i])[1:13])
            f.write("%d %d  %s\n" % (i, len(nb), " ".join(map(str, nb))))
    blocks = [list(range(c * n // K, (c + 1) * n // K)) for c in range(K)]
    with open(prefix + "ske.dat", "w") as f:
        f.write("SKE\n%d %d\n" % (n, K))
        for c, t in enumerate(blocks):
            others = [j for j in range(n) if j not in t]
            o = sorted(set(int(others[k]) for k in np.argsort(d[t][:, others].min(axis=0))[:4]))
            f.write("%d %d\n%s \n%s \n" % (len(t), len(o), " ".join(map(str, t)), " ".join(map(str, o))))
    BIN = os.path.join(ROOT, "cmvs-pmvs_b200", "bin")
    subprocess.run([os.path.join(BIN, "genOption"), prefix, "1", "2", "0.7", "7", "3", str(cpu)], check=True)
    res = {"scene": scene.name, "views": n, "size": [a.width, a.height], "clusters": K, "host_threads": cpu, "runs": []}
    for G in sorted({1, max(1, torch.cuda.device_count())}):
        t = time.time()
        p = subprocess.run([os.path.join(BIN, "pmvs2_clusters"), prefix, "--gpus", str(G), "PSET"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        secs = time.time() - t
        if p.returncode != 0:
            print(p.stderr[-3000:])
            raise SystemExit("pmvs2_clusters failed")
        counts = [sum(1 for _ in open(prefix + "models/option-%04d.pset" % c)) for c in range(K)]
        merged = sum(1 for _ in open(prefix + "models/option-all.pset"))
        res["runs"].append({"gpus": G, "seconds": secs, "patches_per_cluster": counts, "merged": merged,
                            "log": [l for l in p.stderr.splitlines() if l.startswith(("cluster", "merged", "time"))]})
    print(json.dumps(res))


if __name__ == "__main__":
    main()
