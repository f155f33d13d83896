"""The other option shapes BASELINE.json names, stage by stage against the CPU oracle: wsize 5 and 9 (the KG<5> group kernels and
the warp-per-patch k_refine<9> / k_score<9> dispatch arms) and config 2's shape, level 0 / csize 1 (the pyramid-level pick
clamps at -level = 0, a cell is a pixel).  Same bars as tests/test_gpu_parity.py and tests/test_gpu_select.py.  The oracle itself
is pinned on the reference at these shapes and on these patches by tests/test_oracle_shapes.py (golden: pmvs_shapes.npz)."""
import copy

import numpy as np
import pytest

from scene_util import make_patches

pytestmark = pytest.mark.gpu

SHAPES = {"wsize5": dict(wsize=5), "wsize9": dict(wsize=9), "level0_csize1": dict(level=0, csize=1)}


@pytest.fixture(scope="module", params=list(SHAPES))
def shape(request, pkg, scene):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from oracle.bindings import OracleLib
    sc = copy.copy(scene)
    sc.option = dict(scene.option)
    sc.option.update(SHAPES[request.param])
    gpu = pkg.PmvsB200.from_scene(sc)
    orc = OracleLib.from_scene(sc)
    pb = make_patches(sc, orc, 300, seed=21)
    yield request.param, sc, gpu, orc, pb
    gpu.close()


def test_textures_flags_levels(shape):
    name, sc, gpu, orc, pb = shape
    tex, flag, nl = gpu.grab_tex_batch(pb["coords"], pb["normals"], pb["images"])
    n, k = pb["images"].shape
    grabbed = 0
    for p in range(n):
        for v in range(k):
            f, t, l = orc.grab_tex(pb["coords"][p], pb["normals"][p], pb["images"][p, 0], pb["images"][p, v], wsize=sc.option["wsize"])
            assert flag[p, v] == f and nl[p, v] == l, (p, v)
            if f == 0:
                grabbed += 1
                assert np.array_equal(tex[p, v], t), (p, v)
    assert grabbed > n
    if name == "level0_csize1":
        assert nl[flag == 0].min() == 0          # the finest level is really used at level 0


def test_objective_incc_scales(shape):
    name, sc, gpu, orc, pb = shape
    n = len(pb["coords"])
    rng = np.random.default_rng(7)
    x = rng.normal(size=(n, 3)) * np.array([1.5, 2.0, 2.0])
    f = gpu.eval_objective_batch(pb["coords"], pb["normals"], pb["images"], pb["dscales"], x)
    ref = np.array([orc.my_f(pb["coords"][i], pb["normals"][i], pb["images"][i], pb["dscales"][i], x[i]) for i in range(n)])
    assert np.array_equal(f == 2.0, ref == 2.0)
    assert np.abs(f - ref).max() <= 1e-4
    for robust in (1, 0):
        out = gpu.compute_incc_batch(pb["coords"], pb["normals"], pb["images"], robust=robust)
        ref = np.array([orc.compute_incc(pb["coords"][i], pb["normals"][i], pb["images"][i], robust) for i in range(n)])
        assert np.array_equal(out == 2.0, ref == 2.0) and np.abs(out - ref).max() <= 1e-4
    d, a = gpu.set_scales_batch(pb["coords"], pb["images"])
    for i in range(n):
        rd, ra = orc.set_scales(pb["coords"][i], pb["images"][i])
        assert d[i] == rd and abs(float(a[i]) - float(ra)) <= 1e-6 * abs(float(ra)), i


def test_refine(shape):
    """same stated tolerance as tests/test_gpu_parity.py::test_refine_matches_oracle"""
    name, sc, gpu, orc, pb = shape
    g = gpu.refine_batch(pb["coords"], pb["normals"], pb["images"], pb["dscales"])
    o = orc.refine_batch(pb["coords"], pb["normals"], pb["images"], pb["dscales"], threads=8)
    assert (g["ok"] == o["ok"]).mean() >= 0.99
    both = (g["ok"] == 1) & (o["ok"] == 1)
    assert both.sum() > 0.9 * len(both)
    dncc = np.abs(g["ncc"][both] - o["ncc"][both])
    depth = np.linalg.norm(g["coords"][both, :3] - o["coords"][both, :3], axis=1) / pb["dscales"][both]
    ang = np.degrees(np.arccos(np.clip((g["normals"][both, :3] * o["normals"][both, :3]).sum(1), -1, 1)))
    good = (dncc <= 2e-3) & (depth <= 0.05) & (ang <= 1.0)
    print("%s refine parity: good=%.4f median dncc=%.2e depth=%.2e ang=%.2e evals gpu=%.1f cpu=%.1f" % (
        name, good.mean(), np.median(dncc), np.median(depth), np.median(ang), g["evals"].mean(), o["evals"].mean()))
    assert good.mean() >= 0.97


def test_pre_and_post_process(shape):
    name, sc, gpu, orc, pb = shape
    P = len(pb["coords"])
    n0 = np.random.default_rng(3).integers(2, 4, P).astype(np.int32)
    images = np.zeros((P, sc.num), np.int32); images[:, :3] = pb["images"][:, :3]
    out = gpu.pre_process_batch(pb["coords"], pb["normals"], images, n0)
    kept = 0
    for i in range(P):
        v, im, d, a = orc.pre_process(pb["coords"][i], pb["normals"][i], pb["images"][i, : n0[i]], cap=sc.num)
        assert out["verdict"][i] == v and out["nimages"][i] == len(im) and np.array_equal(out["images"][i, : len(im)], im), i
        assert out["dscale"][i] == d, i
        kept += v == 0
    assert kept > 0.3 * P
    keep = np.where(out["verdict"] == 0)[0]
    ncc = np.full(len(keep), 0.9, np.float32)
    post = gpu.post_process_batch(pb["coords"][keep], pb["normals"][keep], ncc, out["images"][keep], out["nimages"][keep])
    for k, i in enumerate(keep):
        v, pim, pgr, pt, ptmp = orc.post_process(pb["coords"][i], pb["normals"][i], 0.9, out["images"][i, : out["nimages"][i]], cap=sc.num)
        assert post["verdict"][k] == v and post["nimages"][k] == len(pim) and np.array_equal(post["images"][k, : len(pim)], pim), k
        if v == 0:
            assert np.array_equal(post["grids"][k, : len(pim)], pgr) and post["timages"][k] == pt and post["tmp"][k] == ptmp, k
