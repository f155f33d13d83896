"""Generate tests/golden/pmvs_envelope.npz: the REFERENCE binary's own run-to-run envelope on the fringe-heavy "enumerated"
option variant (tests/scene_util.option_variants: 8 target + 4 other views, csize 1).  Large parts of the sphere are seen by
barely minImageNum images there, and which of those fringes a run reaches depends on the order in which candidates are tried:
the reference itself writes between ~11 000 and ~20 000 patches for this option file depending on its CPU option (thread
interleaving) -- and even at CPU 1 its result moves by a fraction of a percent from run to run (it sorts shared_ptr values,
i.e. by allocation address, source/pmvs/seed.cpp:322, patchOrganizerS.cpp's neighbour lists).  The envelope = patch counts and
the UNION of the clouds of runs at CPU 1, 2, 4.  tests/test_gpu_pipeline.py holds the drop-in to it: every point of its cloud
lies on the union, the CPU-1 cloud is covered, the count lies inside the reference's own range.
Run:  python tests/golden/make_golden_envelope.py   (needs /root/reference for the build)"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from scene_util import small_scene, write_variant  # noqa: E402
from oracle.bindings import build_ref  # noqa: E402


def main():
    assert build_ref()
    scene = small_scene()
    counts, clouds, cpus = [], [], []
    for cpu in (1, 1, 2, 2, 4, 4):
        prefix = write_variant(scene, "enumerated", "/tmp/pmvs_golden_envelope_%d" % cpu, cpu=cpu)
        subprocess.run([os.path.join(ROOT, "oracle/_ref/pmvs3_ref"), prefix, "option.txt", "PSET"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
        pts = np.loadtxt(prefix + "models/option.txt.pset", dtype=np.float32).reshape(-1, 6)[:, :3]
        counts.append(len(pts)); clouds.append(pts); cpus.append(cpu)
        print("CPU", cpu, "patches", len(pts))
    union = np.concatenate(clouds)
    # thin the union on a grid much finer than the point spacing (the runs share most of the surface)
    key = np.round(union / 0.004).astype(np.int64)
    _, first = np.unique(key, axis=0, return_index=True)
    union = union[np.sort(first)]
    path = os.path.join(HERE, "pmvs_envelope.npz")
    np.savez_compressed(path, scene_sha256=np.frombuffer(bytes.fromhex(scene.sha256()), np.uint8), cpus=np.array(cpus, np.int32),
                        counts=np.array(counts, np.int32), union=union.astype(np.float16))
    print("wrote", path, os.path.getsize(path), "bytes; counts", counts, "union points", len(union))


if __name__ == "__main__":
    main()
