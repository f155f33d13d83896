"""Generate tests/golden/pmvs_clusters.npz: the REFERENCE's genOption (oracle/_ref/genOption_ref) run on
tests/scene_util.write_clusters(), its option files byte for byte, and the clouds the reference binary (CPU 1) writes for
the two clusters.  Run:  python tests/golden/make_golden_clusters.py   (needs /root/reference for the build)"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from scene_util import small_scene, write_clusters  # noqa: E402
from oracle.bindings import build_ref  # noqa: E402

ARGS = ["1", "2", "0.7", "7", "3", "1"]   # level csize threshold wsize minImageNum CPU


def main():
    assert build_ref()
    scene = small_scene()
    prefix = write_clusters(scene, "/tmp/pmvs_golden_clusters", cpu=1)
    subprocess.run([os.path.join(ROOT, "oracle/_ref/genOption_ref"), prefix] + ARGS, check=True)
    out = {"scene_sha256": np.frombuffer(bytes.fromhex(scene.sha256()), np.uint8), "args": np.array(ARGS)}
    for name in ("option-0000", "option-0001", "pmvs.sh"):
        out["file_" + name] = np.frombuffer(open(prefix + name, "rb").read(), np.uint8)
    for c in range(2):
        opt = "option-%04d" % c
        subprocess.run([os.path.join(ROOT, "oracle/_ref/pmvs3_ref"), prefix, opt, "PSET"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
        pts = np.loadtxt(prefix + "models/%s.pset" % opt, dtype=np.float32).reshape(-1, 6)
        out["cluster_%d_pset" % c] = pts
        print(opt, "patches", len(pts))
    path = os.path.join(HERE, "pmvs_clusters.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
