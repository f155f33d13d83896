"""Generate tests/golden/pmvs_features.npz: the features the REFERENCE'S OWN detectors (CHarris::run,
CDifferenceOfGaussians::run through oracle/_ref/libpmvs_ref.so) find on the working-level images of
tests/scene_util.small_scene(), in the order CDetectFeatures stores them.  Run: python tests/golden/make_golden_features.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from scene_util import small_scene  # noqa: E402
import __graft_entry__ as g  # noqa: E402
from oracle.bindings import RefLib, build_ref  # noqa: E402


def main():
    synth = g.load_package().synth
    assert build_ref()
    scene = small_scene()
    prefix = synth.write_scene(scene, "/tmp/pmvs_golden_features_scene")
    ref = RefLib(prefix, num=scene.num, level=scene.option["level"])
    out = {"scene_sha256": np.frombuffer(bytes.fromhex(scene.sha256()), np.uint8)}
    off, xy, resp, typ = [0], [], [], []
    for i in range(scene.num):
        a, r, t = ref.detect_features(i, 16)
        xy.append(a); resp.append(r); typ.append(t); off.append(off[-1] + len(r))
    out["off"] = np.array(off, np.int32); out["xy"] = np.concatenate(xy).astype(np.int16)
    out["resp"] = np.concatenate(resp); out["type"] = np.concatenate(typ).astype(np.int8)
    path = os.path.join(HERE, "pmvs_features.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", off[-1], "features, Harris", int((out["type"] == 0).sum()), "DoG", int((out["type"] == 1).sum()))


if __name__ == "__main__":
    main()
