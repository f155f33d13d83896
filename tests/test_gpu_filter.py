"""GPU parity of the filter-stage kernels (depth maps, visibility, filterExact test, filterOutside gain) on the patch
table of the reference's own run (tests/golden/pmvs_state.npz).  Everything is bit-exact: integers, and f32 gains
evaluated in the reference's operation order."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
STORE_KEYS = ("coords", "normals", "ncc", "dscale", "img_off", "images", "grids", "vimg_off", "vimages", "vgrids", "timages")


@pytest.fixture(scope="module")
def S(scene):
    s = np.load(os.path.join(HERE, "golden", "pmvs_state.npz"))
    assert scene.sha256() == bytes(s["scene_sha256"]).hex()
    return s


@pytest.fixture(scope="module")
def state(gpu, scene, S):
    from oracle.bindings import OracleLib
    st = {k: S["st_" + k] for k in STORE_KEYS}
    o = OracleLib.from_scene(scene)
    o.set_thresholds(float(S["ncc_threshold"]), float(S["ncc_threshold_before"]))
    o.set_depth(int(S["depth_flag"]))
    o.store_set(st)
    o.build_depth_maps()
    gpu.set_thresholds(float(S["ncc_threshold"]), float(S["ncc_threshold_before"]))
    gpu.set_depth(int(S["depth_flag"]))
    gpu.store_upload(st)
    gpu.build_depth_maps()
    yield st, o
    gpu.set_depth(0)
    gpu.set_thresholds(0.7, 0.4)


def test_depth_maps_bit_exact(gpu, scene, S, state):
    st, o = state
    for i in range(scene.num):
        assert gpu.grid_dims(i) == o.grid_dims(i)
        got = gpu.depth_map(i)
        assert np.array_equal(got, o.depth_map(i)), i
    assert np.array_equal(np.concatenate([gpu.depth_map(i) for i in range(scene.num)]), S["depth_maps"])   # = the reference's own maps


def test_depth_maps_incremental(gpu, scene, S, state):
    """pmvsb_depth_maps_add (updateDepthMaps) on the second half of the table == building all at once"""
    st, o = state
    P = len(st["ncc"])
    h = P // 2
    first = {k: st[k] for k in STORE_KEYS}
    first.update(coords=st["coords"][:h], normals=st["normals"][:h], ncc=st["ncc"][:h], dscale=st["dscale"][:h], timages=st["timages"][:h],
                 img_off=st["img_off"][: h + 1], vimg_off=st["vimg_off"][: h + 1],
                 images=st["images"][: st["img_off"][h]], grids=st["grids"][: st["img_off"][h]],
                 vimages=st["vimages"][: st["vimg_off"][h]], vgrids=st["vgrids"][: st["vimg_off"][h]])
    gpu.store_upload(first)
    gpu.build_depth_maps()
    gpu.depth_maps_add(st["coords"][h:])
    got = np.concatenate([gpu.depth_map(i) for i in range(scene.num)])
    assert np.array_equal(got, S["depth_maps"])
    gpu.store_upload(st)   # restore the full table for the other tests
    gpu.build_depth_maps()


def test_set_vimages_bit_exact(gpu, scene, state):
    st, o = state
    vim, vgr, nv = gpu.set_vimages_store(scene.num)
    P = len(st["ncc"])
    rng = np.random.default_rng(1)
    total = 0
    for k in rng.integers(0, P, 1500):
        a, b = o.set_vimages(int(k), cap=scene.num)
        assert nv[k] == len(a), k
        assert np.array_equal(vim[k, : len(a)], a) and np.array_equal(vgr[k, : len(a)], b), k
        total += len(a)
    assert total > 100


def test_filter_exact_bit_exact(gpu, state):
    st, o = state
    safe = gpu.filter_exact_store()
    E = len(st["images"])
    entry_patch = np.repeat(np.arange(len(st["ncc"])), np.diff(st["img_off"]))
    rng = np.random.default_rng(2)
    for e in rng.integers(0, E, 6000):
        want = o.filter_exact_safe(int(entry_patch[e]), int(st["images"][e]), int(st["grids"][e][0]), int(st["grids"][e][1]))
        assert safe[e] == want, e
    assert 0.5 < safe.mean() <= 1.0


def test_gains_bit_exact(gpu, S, state):
    st, o = state
    g = gpu.compute_gains_store()
    assert np.array_equal(g, S["gains"])   # the reference's own computeGain on every patch


def test_store_rejects_bad_indexes(gpu, pkg, state):
    st, _ = state
    bad = dict(st)
    bad["grids"] = st["grids"].copy()
    bad["grids"][0, 0] = 10 ** 6
    with pytest.raises(pkg.PmvsError):
        gpu.store_upload(bad)
    gpu.store_upload(st)   # leave the fixture usable
    gpu.build_depth_maps()
