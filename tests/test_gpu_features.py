"""GPU parity of the feature detectors (pmvs_features.cuh): every plane is computed in the reference's f32 operation
order, so positions, responses, types and order are bit-exact against the reference's own detectors
(tests/golden/pmvs_features.npz) and against the C oracle on images the golden file does not cover."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_features_equal_the_reference(gpu, scene):
    F = np.load(os.path.join(HERE, "golden", "pmvs_features.npz"))
    assert scene.sha256() == bytes(F["scene_sha256"]).hex()
    for i in range(scene.num):
        xy, resp, typ = gpu.detect_features(i, 16)
        lo, hi = F["off"][i], F["off"][i + 1]
        assert len(resp) == hi - lo, i
        assert np.array_equal(xy, F["xy"][lo:hi].astype(np.float32)), i
        assert np.array_equal(resp, F["resp"][lo:hi]), i
        assert np.array_equal(typ, F["type"][lo:hi].astype(np.int32)), i


def test_features_other_block_size_and_capacity(gpu, oracle, pkg):
    """gspeedup 8 (16-pixel blocks, several per warp row) against the oracle; a short output buffer reports the full count"""
    xy, resp, typ = gpu.detect_features(3, 8)
    oxy, oresp, otyp = oracle.detect_features(3, 8)
    assert np.array_equal(xy, oxy) and np.array_equal(resp, oresp) and np.array_equal(typ, otyp)
    import ctypes as C
    n = C.c_int32()
    small = np.zeros((4, 2), np.float32); r = np.zeros(4, np.float32); t = np.zeros(4, np.int32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    assert gpu.lib.pmvsb_detect_features(gpu.ctx, 3, 8, 4, vp(small), vp(r), vp(t), C.byref(n)) == 0
    assert n.value == len(resp) and np.array_equal(r, resp[:4])
    assert gpu.lib.pmvsb_detect_features(gpu.ctx, 99, 8, 4, vp(small), vp(r), vp(t), C.byref(n)) != 0
