"""GPU parity of the device-side cell bookkeeping and the neighbour searches built on it (pmvs_cells.cuh) on the patch
table of the reference's own run: _pgrids / _vpgrids as CSR, in-place setVImagesVGrids, table appends,
findEmptyBlocks and filterNeighbor.  Integers are bit-exact; the quadric residual is compared with a tolerance
because its 5x5 normal equations are accumulated over the lanes (double, different association)."""
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


def _slice(st, lo, hi):
    """patches [lo, hi) of a table as a self-contained table (offsets rebased)"""
    out = {k: st[k][lo:hi] for k in ("coords", "normals", "ncc", "dscale", "timages")}
    for off, keys in (("img_off", ("images", "grids")), ("vimg_off", ("vimages", "vgrids"))):
        a, b = st[off][lo], st[off][hi]
        out[off] = st[off][lo:hi + 1] - a
        for k in keys:
            out[k] = st[k][a:b]
    return out


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


def _host_cells(gpu, scene, off, images, grids, tnum):
    dims = [gpu.grid_dims(i) for i in range(tnum)]
    base = np.concatenate([[0], np.cumsum([w * h for w, h in dims])])
    lists = [[] for _ in range(base[-1])]
    for p in range(len(off) - 1):
        for e in range(off[p], off[p + 1]):
            im = images[e]
            if im < tnum:
                lists[base[im] + grids[e][1] * dims[im][0] + grids[e][0]].append(p)
    return lists


def test_cell_lists_bit_exact(gpu, scene, state):
    st, _ = state
    for visible, off, images, grids in ((0, st["img_off"], st["images"], st["grids"]), (1, st["vimg_off"], st["vimages"], st["vgrids"])):
        want = _host_cells(gpu, scene, off, images, grids.reshape(-1, 2), gpu.num_target)
        coff, clst = gpu.cell_lists(visible)
        assert len(coff) == len(want) + 1 and coff[-1] == sum(len(w) for w in want)
        assert np.array_equal(np.diff(coff), [len(w) for w in want])
        assert np.array_equal(clst, np.concatenate([np.array(w, np.int32) for w in want if w]))   # table order inside a cell


def test_store_update_vimages_bit_exact(gpu, state):
    """in-place setVImagesVGrids: from scratch == the oracle patch by patch; additive on the result adds nothing"""
    st, o = state
    off, vim, vgr = gpu.store_update_vimages(additive=0)
    P = len(st["ncc"])
    for k in range(0, P, 7):
        a, b = o.set_vimages(k)
        assert np.array_equal(vim[off[k]:off[k + 1]], a) and np.array_equal(vgr[off[k]:off[k + 1]], b), k
    off2, vim2, vgr2 = gpu.store_update_vimages(additive=1)
    assert np.array_equal(off, off2) and np.array_equal(vim, vim2) and np.array_equal(vgr, vgr2)
    # the reference's final table went through the same additive passes: its lists hold at least these images
    gpu.store_upload(st); gpu.build_depth_maps()
    off3, vim3, vgr3 = gpu.store_update_vimages(additive=1)
    assert np.array_equal(off3, st["vimg_off"]) and np.array_equal(vim3, st["vimages"])   # already complete: nothing to add
    gpu.store_upload(st); gpu.build_depth_maps()


def test_find_empty_blocks_matches_reference(gpu, S, state):
    st, o = state
    P = len(st["ncc"])
    ids = np.arange(P, dtype=np.int32)
    mask, radius = gpu.find_empty_blocks_store(ids)
    assert np.array_equal(radius, S["radius"])            # computeRadius, bit-exact against the reference's own values
    # the sector index is an integer decision: the kernel evaluates the reference's double atan2 and rounds like it
    assert np.array_equal(mask, S["empty_mask"])
    sub = np.array([5, 17, P - 1, P // 2], np.int32)       # arbitrary subsets, any order
    m2, r2 = gpu.find_empty_blocks_store(sub)
    assert np.array_equal(m2, mask[sub]) and np.array_equal(r2, radius[sub])


def test_filter_neighbor_matches_reference(gpu, S, state):
    st, o = state
    P = len(st["ncc"])
    rej, res, cnt, overflow = gpu.filter_neighbor_store(2.5)
    assert overflow == 0
    assert np.array_equal(cnt, S["fnb_count"])            # unique neighbours of findNeighbors(scale 4, margin 2, skipvis 1)
    assert np.array_equal(rej, S["fnb_reject"])
    want = np.array([o.filter_neighbor(k)[1] for k in range(P)], np.float32)
    fit = want >= 0
    assert np.array_equal(res < 0, want < 0)
    assert np.allclose(res[fit], want[fit], rtol=1e-4, atol=1e-6)
    for qi, q in enumerate(S["fnb_quads"]):
        r2, res2, _, _ = gpu.filter_neighbor_store(float(q))
        near = np.abs(res2 - q) < 1e-4 * q                 # verdicts may differ only where the residual sits on the threshold
        assert np.array_equal(r2[~near], S["fnb_reject_q"][qi][~near]), q
        assert near.mean() < 1e-3


def test_store_append_equals_upload(gpu, S, state):
    """upload the first half, append the rest in two pieces: cell lists, depth maps, gains, masks == one upload"""
    st, o = state
    P = len(st["ncc"])
    full_cells = [gpu.cell_lists(v) for v in (0, 1)]
    full_gain = gpu.compute_gains_store()
    full_safe = gpu.filter_exact_store()
    full_mask = gpu.find_empty_blocks_store(np.arange(P, dtype=np.int32))
    full_dm = np.concatenate([gpu.depth_map(i) for i in range(gpu.num_target)])
    h, h2 = P // 2, P // 2 + P // 5
    gpu.store_upload(_slice(st, 0, h)); gpu.build_depth_maps()
    gpu.store_append(_slice(st, h, h2))
    gpu.store_append(_slice(st, h2, P))
    for v in (0, 1):
        off, lst = gpu.cell_lists(v)
        assert np.array_equal(off, full_cells[v][0]) and np.array_equal(lst, full_cells[v][1])
    assert np.array_equal(np.concatenate([gpu.depth_map(i) for i in range(gpu.num_target)]), full_dm)
    assert np.array_equal(gpu.compute_gains_store(), full_gain)
    assert np.array_equal(gpu.filter_exact_store(), full_safe)
    m, r = gpu.find_empty_blocks_store(np.arange(P, dtype=np.int32))
    assert np.array_equal(m, full_mask[0]) and np.array_equal(r, full_mask[1])
    gpu.store_upload(st); gpu.build_depth_maps()


def test_append_rejects_bad_indexes(gpu, pkg, state):
    st, _ = state
    bad = _slice(st, 0, 4)
    bad["grids"] = bad["grids"].copy(); bad["grids"][0] = [100000, 0]
    with pytest.raises(pkg.PmvsError):
        gpu.store_append(bad)
    gpu.store_upload(st); gpu.build_depth_maps()


def test_check_batch_matches_reference(gpu, S, state):
    """COptim::check for candidates handed over as arrays (here: the table's own patches, so the reference's answers on
    its final table apply): gains bit-exact, verdicts equal at the option's quad and at a tight one"""
    st, o = state
    P = len(st["ncc"])
    ni = np.diff(st["img_off"]); nv = np.diff(st["vimg_off"])
    stride, vstride = int(ni.max()), max(int(nv.max()), 1)
    images = np.zeros((P, stride), np.int32); grids = np.zeros((P, stride, 2), np.int32)
    vimages = np.zeros((P, vstride), np.int32); vgrids = np.zeros((P, vstride, 2), np.int32)
    for k in range(P):
        a, b = st["img_off"][k], st["img_off"][k + 1]
        images[k, :b - a] = st["images"][a:b]; grids[k, :b - a] = st["grids"].reshape(-1, 2)[a:b]
        a, b = st["vimg_off"][k], st["vimg_off"][k + 1]
        vimages[k, :b - a] = st["vimages"][a:b]; vgrids[k, :b - a] = st["vgrids"].reshape(-1, 2)[a:b]
    gain, rej, ov = gpu.check_batch(st["coords"], st["normals"], st["ncc"], st["dscale"], st["timages"], images, ni, grids, vimages, nv, vgrids, 2.5)
    assert ov == 0
    assert np.array_equal(gain, S["check_gain"])
    assert np.array_equal(rej, S["check_reject"])
    _, rej2, _ = gpu.check_batch(st["coords"], st["normals"], st["ncc"], st["dscale"], st["timages"], images, ni, grids, vimages, nv, vgrids, 0.1)
    assert (rej2 != S["check_reject_q01"]).mean() < 1e-3     # residuals on the threshold may round either way (lane-wise double sums)
    sub = slice(100, 164)                                     # a small batch, wider strides than needed
    g3, r3, _ = gpu.check_batch(st["coords"][sub], st["normals"][sub], st["ncc"][sub], st["dscale"][sub], st["timages"][sub],
                                np.pad(images[sub], ((0, 0), (0, 5))), ni[sub], np.pad(grids[sub], ((0, 0), (0, 5), (0, 0))),
                                np.pad(vimages[sub], ((0, 0), (0, 3))), nv[sub], np.pad(vgrids[sub], ((0, 0), (0, 3), (0, 0))), 2.5)
    assert np.array_equal(g3, gain[sub]) and np.array_equal(r3, rej[sub])
