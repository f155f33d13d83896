"""Generate tests/golden/pmvs_pipeline.npz: the output of the REFERENCE binary (oracle/_ref/pmvs3_ref, the reference's own
sources, CPU 1 = deterministic) on tests/scene_util.small_scene(): patch count and the .pset cloud (x y z nx ny nz).
Run:  python tests/golden/make_golden_pipeline.py   (needs /root/reference for the build)"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from scene_util import small_scene  # noqa: E402
import __graft_entry__ as g  # noqa: E402
from oracle.bindings import build_ref  # noqa: E402


# option-file variants of the same scene (tests/scene_util.option_variants): the reference's answer for each
def run_reference(synth, scene, prefix):
    p = subprocess.run([os.path.join(ROOT, "oracle/_ref/pmvs3_ref"), prefix, "option.txt", "PATCH", "PSET"], stdout=subprocess.DEVNULL,
                       stderr=subprocess.PIPE, text=True, check=True)
    pts = np.loadtxt(prefix + "models/option.txt.pset", dtype=np.float32).reshape(-1, 6)
    counts = [int(l.split("->")[1].split()[0]) for l in p.stderr.splitlines() if "->" in l and "%" in l]
    return pts, counts


def main():
    from scene_util import option_variants, write_variant
    synth = g.load_package().synth
    assert build_ref()
    scene = small_scene()
    scene.option["CPU"] = 1
    prefix = synth.write_scene(scene, "/tmp/pmvs_golden_pipeline_scene")
    pts, counts = run_reference(synth, scene, prefix)
    extra = {}
    for name in option_variants():
        vp = write_variant(scene, name, "/tmp/pmvs_golden_pipeline_%s" % name, cpu=1)
        vpts, vcounts = run_reference(synth, scene, vp)
        extra["variant_%s_patches" % name] = np.int32(len(vpts)); extra["variant_%s_pset" % name] = vpts.astype(np.float32)
        print("variant", name, "patches", len(vpts), vcounts)
    path = os.path.join(HERE, "pmvs_pipeline.npz")
    np.savez_compressed(path, scene_sha256=np.frombuffer(bytes.fromhex(scene.sha256()), np.uint8), pset=pts.astype(np.float32),
                        stage_counts=np.array(counts, np.int32), patches=np.int32(len(pts)), **extra)
    rad = np.linalg.norm(pts[:, :3], axis=1)
    print("wrote", path, os.path.getsize(path), "bytes; patches", len(pts), "stage counts", counts, "mean |r-1|", np.abs(rad - 1).mean())


if __name__ == "__main__":
    main()
