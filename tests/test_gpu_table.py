"""The resident patch table reorganised on the device (cmvs-pmvs_b200/csrc/pmvs_table.cuh) against the REFERENCE'S OWN state
transitions (tests/golden/pmvs_state.npz, the tail of make_golden_state.py): removePatch + collectPatches renumbering +
setDepthMapsVGridsVPGridsAddPatchV, filterSmallGroups and filterExact.  Integer work: everything must be equal."""
import os

import numpy as np
import pytest

from test_gpu_filter import STORE_KEYS, S, state  # noqa: F401  (fixtures)

pytestmark = pytest.mark.gpu


def _upload(gpu, st):
    gpu.store_upload(st)
    gpu.build_depth_maps()


def _csr_equal(off_a, a, off_b, b):
    return np.array_equal(off_a, off_b) and np.array_equal(a[: off_a[-1]], b[: off_b[-1]])


def test_rebuild_identity_and_removal(gpu, S, state):
    st, o = state
    P = len(st["ncc"])
    _upload(gpu, st)
    perm = gpu.store_rebuild(None, additive=1)
    assert np.array_equal(perm, np.arange(P))            # the reference's table is already in collectPatches order
    d = gpu.store_download()
    assert _csr_equal(d["img_off"], d["images"], st["img_off"], st["images"]) and np.array_equal(d["grids"], st["grids"])
    assert _csr_equal(d["vimg_off"], d["vimages"], st["vimg_off"], st["vimages"])   # complete lists: additive adds nothing
    assert np.array_equal(np.concatenate([gpu.depth_map(i) for i in range(gpu.num_target)]), S["depth_maps"])
    # a shuffled upload with the creation order handed over comes back in the reference's order
    rng = np.random.default_rng(8)
    sh = rng.permutation(P).astype(np.int32)
    cat = lambda off, a: np.concatenate([a[off[k]:off[k + 1]] for k in sh]) if off[-1] else a[:0]
    st2 = dict(coords=st["coords"][sh], normals=st["normals"][sh], ncc=st["ncc"][sh], dscale=st["dscale"][sh], timages=st["timages"][sh],
               img_off=np.concatenate([[0], np.cumsum(np.diff(st["img_off"])[sh])]).astype(np.int32), images=cat(st["img_off"], st["images"]),
               grids=cat(st["img_off"], st["grids"]),
               vimg_off=np.concatenate([[0], np.cumsum(np.diff(st["vimg_off"])[sh])]).astype(np.int32), vimages=cat(st["vimg_off"], st["vimages"]),
               vgrids=cat(st["vimg_off"], st["vgrids"]))
    _upload(gpu, st2)
    gpu.store_set_seq(sh)                                  # table patch k was created as number sh[k]
    perm = gpu.store_rebuild(None, additive=1)
    assert np.array_equal(sh[perm], np.arange(P))
    d = gpu.store_download()
    assert np.array_equal(d["seq"], np.arange(P)) and _csr_equal(d["img_off"], d["images"], st["img_off"], st["images"])
    _upload(gpu, st)


def test_removal_rebuild_small_groups_and_exact_follow_the_reference(gpu, S, state):
    st, o = state
    P = len(st["ncc"])
    _upload(gpu, st)
    # (1) removePatch + setDepthMapsVGridsVPGridsAddPatchV(1)
    perm = gpu.store_rebuild(S["frag_keep"], additive=1)
    assert np.array_equal(perm, S["frag_perm"])
    d = gpu.store_download()
    assert _csr_equal(d["vimg_off"], d["vimages"], S["frag_vimg_off"], S["frag_vimages"]) and np.array_equal(d["vgrids"], S["frag_vgrids"])
    # (2) filterSmallGroups: neighbour tests on the device, labelling walk over them
    keep, thr = gpu.filter_small_groups_store(1.0)
    assert thr == max(20, len(perm) // 10000)
    assert np.array_equal(perm[keep == 1], S["frag_groups_survivors"])
    assert 0 < (keep == 0).sum() < len(keep)
    perm2 = gpu.store_rebuild(keep, additive=1)
    _upload(gpu, st)


def test_filter_exact_follows_the_reference(gpu, S, state):
    """CFilter::filterExact on a table with occluders (a sixth of the patches moved towards their cameras by the golden generator):
    the visibility re-test prunes image lists, patches left with too few images go, setRefImage + setGrids for the rest."""
    st, o = state
    ex = {k: S["exact_st_" + k] for k in STORE_KEYS}
    gpu.store_upload(ex)
    gpu.build_depth_maps()
    keep = gpu.filter_exact_apply_store()
    assert 0 < (keep == 0).sum() < len(keep)
    perm = gpu.store_rebuild(keep, additive=2)          # renumbering only (collectPatches), as ref.state() does
    assert np.array_equal(perm, S["exact_survivors"])
    d = gpu.store_download()
    assert _csr_equal(d["img_off"], d["images"], S["exact_img_off"], S["exact_images"]) and np.array_equal(d["grids"], S["exact_grids"])
    assert np.array_equal(d["timages"], S["exact_timages"])
    assert d["img_off"][-1] < ex["img_off"][-1]          # entries were really pruned
    _upload(gpu, st)


def test_small_group_edges_equal_the_oracles_neighbour_tests(gpu, scene, S, state):
    st, o = state
    P = len(st["ncc"])
    _upload(gpu, st)
    off, adj = gpu.small_group_edges_store(1.0)
    coff, clst = gpu.cell_lists(0)
    voff, vlst = gpu.cell_lists(1)
    base = np.concatenate([[0], np.cumsum([gpu.grid_dims(i)[0] * gpu.grid_dims(i)[1] for i in range(gpu.num_target)])])
    rng = np.random.default_rng(4)
    for p in rng.integers(0, P, 150):
        e0 = st["img_off"][p]
        im = st["images"][e0]; ix, iy = st["grids"][e0]
        gw, gh = gpu.grid_dims(im)
        want = []
        for y in (-1, 0, 1):
            for x in (-1, 0, 1):
                xx, yy = ix + x, iy + y
                if xx < 0 or gw <= xx or yy < 0 or gh <= yy:
                    continue
                c = base[im] + yy * gw + xx
                for lo, ls in ((coff, clst), (voff, vlst)):
                    want += [q for q in ls[lo[c]:lo[c + 1]] if o.is_neighbor(p, q, 1.0)]
        assert list(adj[off[p]:off[p + 1]]) == want, p
