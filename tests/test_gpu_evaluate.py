"""pmvsb_evaluate_batch -- the reference's per-candidate contract (preProcess, refinePatch, postProcess incl. setVImagesVGrids
and check) fused into one call -- against the same stages called one by one through the ABI (each of which is pinned on the
oracle / the reference elsewhere).  Same kernels, same order: every field must be EQUAL."""
import numpy as np
import pytest

from scene_util import make_patches
from test_gpu_filter import STORE_KEYS, S, state  # noqa: F401  (fixtures)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def wave(scene, oracle):
    pb = make_patches(scene, oracle, 600, seed=41, depth_sigma=0.008, normal_sigma=0.25)
    n0 = np.random.default_rng(6).integers(2, 4, len(pb["coords"])).astype(np.int32)
    return pb, n0


def _stepwise(gpu, scene, pb, n0, depth, quad=2.5):
    P = len(pb["coords"])
    stride = scene.num
    images = np.zeros((P, stride), np.int32); images[:, :3] = pb["images"][:, :3]
    pre = gpu.pre_process_batch(pb["coords"], pb["normals"], images, n0)
    verdict = np.where(pre["verdict"] == 0, 0, 1).astype(np.int32)
    live = np.where(pre["verdict"] == 0)[0]
    ref = gpu.refine_batch(pb["coords"][live], pb["normals"][live], pre["images"][live], pre["dscale"][live], nimages=pre["nimages"][live])
    post = gpu.post_process_batch(ref["coords"], ref["normals"], ref["ncc"], pre["images"][live], pre["nimages"][live])
    verdict[live[post["verdict"] != 0]] = 2
    acc = np.where(post["verdict"] == 0)[0]
    rec = dict(index=live[acc], coords=ref["coords"][acc], normals=ref["normals"][acc], ncc=np.where(ref["ok"][acc] == 1, ref["ncc"][acc], -1.0).astype(np.float32),
               dscale=pre["dscale"][live][acc], ascale=pre["ascale"][live][acc], tmp=post["tmp"][acc], timages=post["timages"][acc],
               images=post["images"][acc], nimages=post["nimages"][acc], grids=post["grids"][acc])
    A = len(acc)
    rec["vimages"] = np.zeros((A, gpu.num_target), np.int32); rec["nv"] = np.zeros(A, np.int32); rec["vgrids"] = np.zeros((A, gpu.num_target, 2), np.int32)
    if depth >= 1 and A:
        vim, nv, vgr = gpu.set_vimages_batch(rec["coords"], rec["normals"], rec["images"], rec["nimages"], rec["vimages"], rec["nv"], rec["vgrids"])
        rec["vimages"], rec["nv"], rec["vgrids"] = vim, nv, vgr
    if depth >= 2 and A:
        gain, rej, ov = gpu.check_batch(rec["coords"], rec["normals"], ref["ncc"][acc], rec["dscale"], rec["timages"], rec["images"], rec["nimages"], rec["grids"],
                                        rec["vimages"], rec["nv"], rec["vgrids"], quad)
        rec["tmp"] = gain
        verdict[rec["index"][rej != 0]] = 2
        keep = rej == 0
        rec = {k: v[keep] for k, v in rec.items()}
    return verdict, len(live), rec


def _compare(out, verdict, refined, rec):
    assert np.array_equal(out["verdict"], verdict)
    assert out["refined"] == refined
    assert np.array_equal(out["index"], rec["index"])
    for k in ("coords", "normals", "ncc", "dscale", "ascale", "tmp", "timages"):
        assert np.array_equal(out[k], rec[k]), k
    A = len(rec["index"])
    assert np.array_equal(np.diff(out["img_off"]), rec["nimages"]) and np.array_equal(np.diff(out["vimg_off"]), rec["nv"])
    for j in range(A):
        n, nv = rec["nimages"][j], rec["nv"][j]
        assert np.array_equal(out["images"][out["img_off"][j]:out["img_off"][j + 1]], rec["images"][j, :n]), j
        assert np.array_equal(out["grids"][out["img_off"][j]:out["img_off"][j + 1]], rec["grids"][j, :n]), j
        assert np.array_equal(out["vimages"][out["vimg_off"][j]:out["vimg_off"][j + 1]], rec["vimages"][j, :nv]), j
        assert np.array_equal(out["vgrids"][out["vimg_off"][j]:out["vimg_off"][j + 1]], rec["vgrids"][j, :nv]), j


def _csr(pb, n0):
    off = np.concatenate([[0], np.cumsum(n0)]).astype(np.int32)
    images = np.concatenate([pb["images"][i, : n0[i]] for i in range(len(n0))]).astype(np.int32)
    return off, images


def test_depth0_equals_the_stages_called_one_by_one(gpu, scene, wave):
    pb, n0 = wave
    gpu.set_depth(0)
    gpu.set_thresholds(0.7, 0.4)
    verdict, refined, rec = _stepwise(gpu, scene, pb, n0, depth=0)
    off, images = _csr(pb, n0)
    out = gpu.evaluate_batch(pb["coords"], pb["normals"], off, images)
    _compare(out, verdict, refined, rec)
    assert (verdict == 0).sum() > 50 and (verdict == 1).sum() > 20 and (verdict == 2).sum() >= 1   # all three outcomes occur
    assert len(out["vimages"]) == 0
    # an empty wave is a no-op
    e = gpu.evaluate_batch(np.zeros((0, 4), np.float32), np.zeros((0, 4), np.float32), np.zeros(1, np.int32), np.zeros(0, np.int32))
    assert len(e["verdict"]) == 0 and len(e["index"]) == 0


@pytest.mark.parametrize("quad", [2.5, 0.05])
def test_depth2_equals_the_stages_called_one_by_one(gpu, scene, wave, S, state, quad):
    """against the reference's final table: setVImagesVGrids and COptim::check (gain + quadric) inside the fused call"""
    pb, n0 = wave
    st, o = state                      # table uploaded, depth maps built, thresholds and _depth of the reference's last round
    verdict, refined, rec = _stepwise(gpu, scene, pb, n0, depth=2, quad=quad)
    off, images = _csr(pb, n0)
    out = gpu.evaluate_batch(pb["coords"], pb["normals"], off, images, quad=quad)
    _compare(out, verdict, refined, rec)
    assert len(out["vimages"]) > 0
