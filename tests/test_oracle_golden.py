"""CPU suite: the plain-C oracle (oracle/pmvs_oracle.c) against golden vectors produced by the REFERENCE'S
OWN objects (tests/golden/make_golden.py).  Everything here must be bit-exact: the oracle restates the
reference's arithmetic operation by operation."""
import hashlib
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def G():
    return np.load(os.path.join(HERE, "golden", "pmvs_golden.npz"))


@pytest.fixture(scope="module")
def scene_checked(scene, G):
    got = scene.sha256()
    want = bytes(G["scene_sha256"]).hex()
    assert got == want, "synthetic scene is not bit-reproducible on this machine: %s != %s" % (got, want)
    return scene


def test_scene_reproducible(scene_checked):
    assert scene_checked.num == 16


def test_cameras(oracle, scene_checked, G):
    for i in range(scene_checked.num):
        c = oracle.camera(i, 1)
        for k in ("P", "centre", "oaxis", "xaxis", "yaxis", "zaxis", "ipscale"):
            assert np.array_equal(np.atleast_1d(c[k]).ravel(), G["cam_" + k][i].ravel()), (i, k)


def test_pyramid(oracle, scene_checked, G):
    k = 0
    for i in range(scene_checked.num):
        for l in range(scene_checked.option["level"] + 3):
            d = np.frombuffer(hashlib.sha256(oracle.image(i, l).tobytes()).digest(), np.uint8)
            assert np.array_equal(d, G["pyr_sha256"][k]), (i, l)
            k += 1
    assert np.array_equal(oracle.image(3, 3), G["pyr_img3_level3"])


def test_project_unit_scales(oracle, scene_checked, G):
    n = len(G["coords"])
    for l in (0, 1, 2):
        got = np.stack([oracle.project(G["proj_image"][i], G["coords"][i], l) for i in range(n)])
        assert np.array_equal(got, G["proj_l%d" % l])
    unit = np.array([oracle.get_unit(G["proj_image"][i], G["coords"][i]) for i in range(n)], np.float32)
    assert np.array_equal(unit, G["unit"])
    for i in range(n):
        d, a = oracle.set_scales(G["coords"][i], G["images"][i])
        assert d == G["dscale"][i] and a == G["ascale"][i], i


def test_grab_tex_normalize_dot(oracle, scene_checked, G):
    nviews = G["images"].shape[1]
    for i in range(G["tex"].shape[0]):
        for v in range(nviews):
            f, t, nl = oracle.grab_tex(G["coords"][i], G["normals"][i], G["images"][i, 0], G["images"][i, v])
            assert f == G["tex_flag"][i, v], (i, v)
            if f == 0:
                assert np.array_equal(t, G["tex"][i, v]), (i, v)
                assert 0 <= nl <= 3
    for k in range(len(G["norm_in"])):
        assert np.array_equal(oracle.normalize(G["norm_in"][k]), G["norm_out"][k])
        assert oracle.dot(G["norm_out"][k], G["norm_out"][(k + 1) % len(G["norm_in"])]) == G["dot_out"][k]
    # a constant texture has sigma 0 -> treated as 1 (optim.cpp:1057-1059)
    flat = np.full(147, 93.0, np.float32)
    assert np.array_equal(oracle.normalize(flat), np.zeros(147, np.float32))


def test_encode_decode_objective(oracle, scene_checked, G):
    n = len(G["coords"])
    for i in range(n):
        c, nm, im, ds = G["coords"][i], G["normals"][i], G["images"][i], G["dscale"][i]
        assert np.array_equal(oracle.encode(c, nm, im, ds), G["encode"][i]), i
        oc, on = oracle.decode(c, nm, im, ds, G["x"][i])
        assert np.array_equal(oc, G["decode_coord"][i]) and np.array_equal(on, G["decode_normal"][i]), i
        assert oracle.my_f(c, nm, im, ds, G["x"][i]) == G["my_f"][i], i
        assert oracle.compute_incc(c, nm, im, 1) == G["incc_robust"][i], i
        assert oracle.compute_incc(c, nm, im, 0) == G["incc_plain"][i], i
        assert np.array_equal(oracle.set_inccs(c, nm, im, 0), G["set_inccs"][i]), i
        assert np.array_equal(oracle.set_inccs_matrix(c, nm, im, 1), G["set_inccs_matrix"][i]), i
    assert (G["my_f"] < 2.0).sum() > n // 2 and (G["my_f"] == 2.0).sum() > 0   # both branches are pinned
    # fewer than two images: computeINCC returns 2.0 before touching any texture (optim.cpp:866)
    assert oracle.compute_incc(G["coords"][0], G["normals"][0], G["images"][0][:1], 1) == 2.0


def test_refine(oracle, scene_checked, G):
    m = len(G["refine_ok"])
    for i in range(m):
        ok, c, nm, ncc, ev = oracle.refine(G["coords"][i], G["normals"][i], G["images"][i], G["dscale"][i])
        assert ok == G["refine_ok"][i] and ev == G["refine_evals"][i], i
        assert np.array_equal(c, G["refine_coord"][i]) and np.array_equal(nm, G["refine_normal"][i]), i
        assert ncc == G["refine_ncc"][i], i
    # batched entry point (threads) gives the same answers as the scalar one
    r = oracle.refine_batch(G["coords"][:m], G["normals"][:m], G["images"][:m], G["dscale"][:m], threads=4)
    assert np.array_equal(r["ok"], G["refine_ok"]) and np.array_equal(r["evals"], G["refine_evals"])
    assert np.array_equal(r["coords"], G["refine_coord"]) and np.array_equal(r["ncc"], G["refine_ncc"])


def test_refine_maxeval_is_failure(oracle, scene_checked, G):
    """MAXEVAL_REACHED is not a success: the patch stays untouched (optim.cpp:644-655)."""
    oracle.set_xtol(1e-3, 1.0, 10)
    try:
        ok, c, nm, ncc, ev = oracle.refine(G["coords"][5], G["normals"][5], G["images"][5], G["dscale"][5])
        assert ok == 0 and ev == 10
        assert np.array_equal(c, G["coords"][5]) and np.array_equal(nm, G["normals"][5])
    finally:
        oracle.set_xtol(1e-3, 1.0, 1000)


def test_pre_post_process(oracle, scene_checked, G):
    n = len(G["pp_coords"])
    for i in range(n):
        v, im, d, a = oracle.pre_process(G["pp_coords"][i], G["pp_normals"][i], G["pp_images"][i])
        assert v == G["pre_verdict"][i], i
        assert np.array_equal(im, G["pre_images"][i, : G["pre_n"][i]]), i
        assert d == G["pre_dscale"][i] and a == G["pre_ascale"][i], i
        if v == 0:
            pv, pim, pgr, pt, ptmp = oracle.post_process(G["post_in_coord"][i], G["post_in_normal"][i], G["post_in_ncc"][i], im)
            assert pv == G["post_verdict"][i], i
            assert np.array_equal(pim, G["post_images"][i, : G["post_n"][i]]), i
            assert np.array_equal(pgr, G["post_grids"][i, : G["post_n"][i]]), i
            assert pt == G["post_timages"][i] and ptmp == G["post_tmp"][i], i
    assert (G["pre_verdict"] == 0).sum() > 50 and (G["pre_verdict"] == 1).sum() > 10


# ---- filter stage: the reference's own run, its depth maps, visibility tests, neighbour tests and gains ------------
@pytest.fixture(scope="module")
def S(scene):
    s = np.load(os.path.join(HERE, "golden", "pmvs_state.npz"))
    assert scene.sha256() == bytes(s["scene_sha256"]).hex()
    return s


STORE_KEYS = ("coords", "normals", "ncc", "dscale", "img_off", "images", "grids", "vimg_off", "vimages", "vgrids", "timages")


@pytest.fixture(scope="module")
def oracle_state(scene, S):
    from oracle.bindings import OracleLib
    o = OracleLib.from_scene(scene)
    o.set_thresholds(float(S["ncc_threshold"]), float(S["ncc_threshold_before"]))
    o.set_depth(int(S["depth_flag"]))
    o.store_set({k: S["st_" + k] for k in STORE_KEYS})
    o.build_depth_maps()
    return o


def test_depth_maps(oracle_state, scene, S):
    got = np.concatenate([oracle_state.depth_map(i) for i in range(scene.num)])
    assert np.array_equal(got, S["depth_maps"])
    assert (got >= 0).mean() > 0.3


def test_is_visible_and_vimages(oracle_state, S):
    q, a = S["vis_query"], S["vis_answer"]
    for i in range(len(q)):
        k, im, ix, iy, strict = int(q[i, 0]), int(q[i, 1]), int(q[i, 2]), int(q[i, 3]), float(q[i, 4])
        assert oracle_state.is_visible_k(k, im, ix, iy, strict) == a[i], i
    assert 0.05 < (a == 0).mean() < 0.95
    for j, k in enumerate(S["vis_k"]):
        vim, vgr = oracle_state.set_vimages(int(k))
        lo, hi = S["vim_off"][j], S["vim_off"][j + 1]
        assert np.array_equal(vim, S["vim"][lo:hi]) and np.array_equal(vgr, S["vgr"].reshape(-1, 2)[lo:hi]), k


def test_is_neighbor_and_gain(oracle_state, S):
    for a, b, n1, n05 in S["nb_pairs"]:
        assert oracle_state.is_neighbor(a, b, 1.0) == n1 and oracle_state.is_neighbor(a, b, 0.5) == n05, (a, b)
    gains = np.array([oracle_state.compute_gain(k) for k in range(len(S["gains"]))], np.float32)
    assert np.array_equal(gains, S["gains"])


def test_find_neighbors_empty_blocks_filter_neighbor(oracle_state, S):
    """findNeighbors (both call patterns), computeRadius, findEmptyBlocks' fill mask and filterNeighbor's verdict
    (at the option's quad and at three tighter ones) for EVERY patch of the reference's final table"""
    o = oracle_state
    P = len(S["st_ncc"])
    for k in range(P):
        assert np.array_equal(o.find_neighbors(k, 4.0, 1, 0), S["fn_m1"][S["fn_m1_off"][k]:S["fn_m1_off"][k + 1]]), k
        assert np.array_equal(o.find_neighbors(k, 4.0, 2, 1), S["fn_m2"][S["fn_m2_off"][k]:S["fn_m2_off"][k + 1]]), k
        assert o.compute_radius(k) == S["radius"][k], k
        assert o.find_empty_blocks(k)[0] == S["empty_mask"][k], k
        rej, res, cnt = o.filter_neighbor(k)
        assert rej == S["fnb_reject"][k] and cnt == S["fnb_count"][k], k
    for qi, q in enumerate(S["fnb_quads"]):
        mine = np.array([o.filter_neighbor(k, float(q))[0] for k in range(0, P, 3)], np.uint8)
        assert np.array_equal(mine, S["fnb_reject_q"][qi][::3]), q
    assert 0.2 < S["fnb_reject_q"][0].mean() < 0.8 and S["empty_mask"].min() < 63


def test_detect_features(oracle, scene_checked):
    """Harris + DoG features of every image: positions, responses (bit for bit), types and order equal the reference's"""
    F = np.load(os.path.join(HERE, "golden", "pmvs_features.npz"))
    assert scene_checked.sha256() == bytes(F["scene_sha256"]).hex()
    for i in range(scene_checked.num):
        xy, resp, typ = oracle.detect_features(i, 16)
        lo, hi = F["off"][i], F["off"][i + 1]
        assert np.array_equal(xy, F["xy"][lo:hi].astype(np.float32)), i
        assert np.array_equal(resp, F["resp"][lo:hi]) and np.array_equal(typ, F["type"][lo:hi].astype(np.int32)), i
    assert (F["type"] == 0).sum() > 100 and (F["type"] == 1).sum() > 100


def test_check(oracle_state, S):
    """COptim::check (gain + quadric test) on every table patch at the option's quad and at a tight one"""
    o = oracle_state
    P = len(S["st_ncc"])
    got = [o.check(k, 2.5) for k in range(P)]
    assert np.array_equal(np.array([r for r, _ in got], np.uint8), S["check_reject"])
    assert np.array_equal(np.array([g for _, g in got], np.float32), S["check_gain"])
    assert np.array_equal(np.array([o.check(k, 0.1)[0] for k in range(P)], np.uint8), S["check_reject_q01"])
    assert 0.01 < S["check_reject_q01"].mean() < 0.5
