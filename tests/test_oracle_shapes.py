"""CPU suite: the plain-C oracle against the reference's own answers at the OTHER option shapes BASELINE.json names -- wsize 5,
wsize 9, level 0 / csize 1 (tests/golden/make_golden_shapes.py drives oracle/_ref at each of them).  tests/test_gpu_shapes.py
compares the CUDA path with the oracle at exactly these shapes and on these patches, so this file is what anchors that
comparison on the reference.  Bit-exact, like tests/test_oracle_golden.py."""
import copy
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SHAPES = {"wsize5": dict(wsize=5), "wsize9": dict(wsize=9), "level0_csize1": dict(level=0, csize=1)}


@pytest.fixture(scope="module")
def GS():
    return np.load(os.path.join(HERE, "golden", "pmvs_shapes.npz"))


@pytest.fixture(scope="module", params=list(SHAPES))
def shaped(request, scene, GS):
    from oracle.bindings import OracleLib
    from scene_util import make_patches
    sc = copy.copy(scene)
    sc.option = dict(scene.option)
    sc.option.update(SHAPES[request.param])
    orc = OracleLib.from_scene(sc)
    G = {k.split("__", 1)[1]: GS[k] for k in GS.files if k.startswith(request.param + "__")}
    # the golden patches are the ones the GPU shape tests draw
    pb = make_patches(sc, orc, 300, seed=21)
    assert np.array_equal(pb["coords"][: len(G["coords"])], G["coords"]) and np.array_equal(pb["images"][: len(G["coords"])], G["images"])
    return request.param, sc, orc, G


def test_scales_textures(shaped):
    name, sc, orc, G = shaped
    n, w = len(G["coords"]), sc.option["wsize"]
    for i in range(n):
        d, a = orc.set_scales(G["coords"][i], G["images"][i])
        assert d == G["dscale"][i] and a == G["ascale"][i], (name, i)
    assert G["tex"].shape[2] == 3 * w * w
    for i in range(G["tex"].shape[0]):
        for v in range(G["images"].shape[1]):
            f, t, _ = orc.grab_tex(G["coords"][i], G["normals"][i], G["images"][i, 0], G["images"][i, v], wsize=w)
            assert f == G["tex_flag"][i, v], (name, i, v)
            if f == 0:
                assert np.array_equal(t, G["tex"][i, v]), (name, i, v)
    assert (G["tex_flag"] == 0).sum() > 20


def test_objective_incc(shaped):
    name, sc, orc, G = shaped
    for i in range(len(G["coords"])):
        c, nm, im, ds = G["coords"][i], G["normals"][i], G["images"][i], G["dscale"][i]
        assert orc.my_f(c, nm, im, ds, G["x"][i]) == G["my_f"][i], (name, i)
        assert orc.compute_incc(c, nm, im, 1) == G["incc_robust"][i], (name, i)
        assert orc.compute_incc(c, nm, im, 0) == G["incc_plain"][i], (name, i)
        assert np.array_equal(orc.set_inccs(c, nm, im, 0), G["set_inccs"][i]), (name, i)
    assert (G["my_f"] < 2.0).sum() > len(G["coords"]) // 2


def test_refine(shaped):
    name, sc, orc, G = shaped
    m = len(G["refine_ok"])
    for i in range(m):
        ok, c, nm, ncc, ev = orc.refine(G["coords"][i], G["normals"][i], G["images"][i], G["dscale"][i])
        assert ok == G["refine_ok"][i] and ev == G["refine_evals"][i], (name, i)
        assert np.array_equal(c, G["refine_coord"][i]) and np.array_equal(nm, G["refine_normal"][i]) and ncc == G["refine_ncc"][i], (name, i)
    assert G["refine_ok"].sum() >= m - 2


def test_pre_post_process(shaped):
    name, sc, orc, G = shaped
    for i in range(len(G["pp_coords"])):
        v, im, d, a = orc.pre_process(G["pp_coords"][i], G["pp_normals"][i], G["pp_images"][i])
        assert v == G["pre_verdict"][i], (name, i)
        assert np.array_equal(im, G["pre_images"][i, : G["pre_n"][i]]), (name, i)
        assert d == G["pre_dscale"][i] and a == G["pre_ascale"][i], (name, i)
        if v == 0:
            pv, pim, pgr, pt, ptmp = orc.post_process(G["post_in_coord"][i], G["post_in_normal"][i], G["post_in_ncc"][i], im)
            assert pv == G["post_verdict"][i], (name, i)
            assert np.array_equal(pim, G["post_images"][i, : G["post_n"][i]]), (name, i)
            assert np.array_equal(pgr, G["post_grids"][i, : G["post_n"][i]]), (name, i)
            assert pt == G["post_timages"][i] and ptmp == G["post_tmp"][i], (name, i)
    assert (G["pre_verdict"] == 0).sum() > 10
