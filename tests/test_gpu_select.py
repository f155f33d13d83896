"""GPU parity of the visible-image-set selection kernels (preProcess / postProcess) against the oracle.
Integer results (image lists, grid cells, verdicts, _timages) must be bit-exact; dscale is f32 arithmetic in the
reference's order (bit-exact), ascale goes through a double atan (1 ulp tolerated), _tmp = score2 is f32 (exact)."""
import numpy as np
import pytest

from scene_util import make_patches

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def candidates(scene, oracle):
    # candidates as the seed / expansion rounds hand them over: 2-3 images, rough depth and normal
    pb = make_patches(scene, oracle, 500, seed=31, depth_sigma=0.01, normal_sigma=0.3)
    rng = np.random.default_rng(3)
    n0 = rng.integers(2, 4, len(pb["coords"])).astype(np.int32)
    return pb, n0


def _run_pre(gpu, scene, pb, n0):
    stride = scene.num
    P = len(pb["coords"])
    images = np.full((P, stride), 0, np.int32)
    images[:, :3] = pb["images"][:, :3]
    return gpu.pre_process_batch(pb["coords"], pb["normals"], images, n0)


def test_pre_process_bit_exact(gpu, oracle, scene, candidates):
    pb, n0 = candidates
    out = _run_pre(gpu, scene, pb, n0)
    P = len(pb["coords"])
    kept = 0
    for i in range(P):
        v, im, d, a = oracle.pre_process(pb["coords"][i], pb["normals"][i], pb["images"][i, : n0[i]], cap=scene.num)
        assert out["verdict"][i] == v, i
        assert out["nimages"][i] == len(im), i
        assert np.array_equal(out["images"][i, : len(im)], im), i
        assert out["dscale"][i] == d, i
        assert abs(float(out["ascale"][i]) - float(a)) <= 2e-7 * max(1.0, abs(float(a))), i
        kept += v == 0
    assert 0.3 * P < kept < P  # both verdicts are exercised


def test_post_process_bit_exact(gpu, oracle, scene, candidates):
    pb, n0 = candidates
    pre = _run_pre(gpu, scene, pb, n0)
    keep = np.where(pre["verdict"] == 0)[0]
    # refine on the oracle side so that both post-processes see identical patches
    P = len(keep)
    coords = np.zeros((P, 4), np.float32); normals = np.zeros((P, 4), np.float32); ncc = np.zeros(P, np.float32)
    for k, i in enumerate(keep):
        im = pre["images"][i, : pre["nimages"][i]]
        ok, c, nm, nc, _ = oracle.refine(pb["coords"][i], pb["normals"][i], im, pre["dscale"][i])
        coords[k], normals[k], ncc[k] = c, nm, nc
    out = gpu.post_process_batch(coords, normals, ncc, pre["images"][keep], pre["nimages"][keep])
    verdicts = [0, 0]
    for k, i in enumerate(keep):
        im = pre["images"][i, : pre["nimages"][i]]
        v, pim, pgr, pt, ptmp = oracle.post_process(coords[k], normals[k], ncc[k], im, cap=scene.num)
        assert out["verdict"][k] == v, k
        verdicts[v] += 1
        assert out["nimages"][k] == len(pim), k
        assert np.array_equal(out["images"][k, : len(pim)], pim), k
        if v == 0:
            assert np.array_equal(out["grids"][k, : len(pim)], pgr), k
            assert out["timages"][k] == pt and out["tmp"][k] == ptmp, k
    assert verdicts[0] > 0.5 * P


def test_pre_process_ragged_and_empty(gpu, pkg, scene, candidates):
    pb, n0 = candidates
    # zero images -> rejected, untouched
    P = 8
    images = np.zeros((P, scene.num), np.int32); images[:, :3] = pb["images"][:P, :3]
    n = np.array([0, 1, 2, 3, 3, 3, 2, 0], np.int32)
    out = gpu.pre_process_batch(pb["coords"][:P], pb["normals"][:P], images, n)
    assert out["verdict"][0] == 1 and out["nimages"][0] == 0
    # a capacity smaller than what addImages finds is an ERROR (the reference's lists are unbounded), never a silent cap
    # (on this ring of cameras a surface point has ~3 cameras inside the 60-degree cone: one listed + two found > capacity 2)
    small = np.zeros((P, 2), np.int32); small[:, :1] = pb["images"][:P, :1]
    with pytest.raises(pkg.PmvsError, match="more images than the list capacity"):
        gpu.pre_process_batch(pb["coords"][:P], pb["normals"][:P], small, np.ones(P, np.int32))
    # the context stays usable and the flag is cleared
    again = gpu.pre_process_batch(pb["coords"][:P], pb["normals"][:P], images, n)
    assert np.array_equal(again["images"], out["images"]) and np.array_equal(again["verdict"], out["verdict"])
