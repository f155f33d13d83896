"""Generate tests/golden/pmvs_state.npz: the filter-stage state of the REFERENCE'S OWN run (CFindMatch::run on
tests/scene_util.small_scene(), CPU 1 -- deterministic) and the reference's answers on it: depth maps
(CFilter::setDepthMaps), CPatchOrganizerS::isVisible / setVImagesVGrids, CFindMatch::isNeighbor and
CFilter::computeGain.  Run:  python tests/golden/make_golden_state.py   (needs /root/reference)"""
import ctypes
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
    prefix = synth.write_scene(scene, "/tmp/pmvs_golden_state_scene")
    ref = RefLib(prefix, num=scene.num, level=scene.option["level"], skip_features=False)
    ref.run()
    st = ref.state()
    P = len(st["ncc"])
    th = np.zeros(6, np.float32)
    ref.lib.ref_thresholds(th.ctypes.data_as(ctypes.c_void_p))
    out = {"scene_sha256": np.frombuffer(bytes.fromhex(scene.sha256()), np.uint8), "depth_flag": np.int32(ref.depth_flag()),
           "ncc_threshold": th[0], "ncc_threshold_before": th[1]}
    out.update({"st_" + k: v for k, v in st.items()})
    ref.build_depth_maps()
    out["depth_maps"] = np.concatenate([ref.depth_map(i) for i in range(scene.num)])
    rng = np.random.default_rng(99)
    ks = rng.integers(0, P, 600).astype(np.int32)
    vis_q, vis_a = [], []
    vim_off, vim, vgr = [0], [], []
    for k in ks:
        for e in range(st["img_off"][k], st["img_off"][k + 1]):
            im = st["images"][e]; gx, gy = st["grids"][e]
            for dx, dy in ((0, 0), (1, 0), (-1, 0), (0, 1), (0, -1)):
                for strict in (0.5, 1.0):
                    vis_q.append((k, im, gx + dx, gy + dy, strict)); vis_a.append(ref.is_visible_k(k, im, gx + dx, gy + dy, strict))
        a, b = ref.set_vimages(k)
        vim.append(a); vgr.append(b); vim_off.append(vim_off[-1] + len(a))
    out["vis_k"] = ks
    out["vis_query"] = np.array(vis_q, np.float32); out["vis_answer"] = np.array(vis_a, np.int8)
    out["vim_off"] = np.array(vim_off, np.int32)
    out["vim"] = np.concatenate(vim).astype(np.int32); out["vgr"] = np.concatenate(vgr).astype(np.int32)
    pairs = []
    for _ in range(4000):
        a = int(rng.integers(0, P)); b = int(rng.integers(0, P)) if rng.random() < 0.3 else min(P - 1, a + int(rng.integers(1, 30)))
        pairs.append((a, b, ref.is_neighbor(a, b, 1.0), ref.is_neighbor(a, b, 0.5)))
    out["nb_pairs"] = np.array(pairs, np.int32)
    out["gains"] = np.array([ref.compute_gain(k) for k in range(P)], np.float32)
    # neighbour searches on the same table: findNeighbors (both call patterns), computeRadius, findEmptyBlocks' fill
    # mask, filterNeighbor's verdict -- for every patch (small scene)
    nb_off1, nb1, nb_off2, nb2 = [0], [], [0], []
    for k in range(P):
        a = ref.find_neighbors(k, 4.0, 1, 0); b = ref.find_neighbors(k, 4.0, 2, 1)
        nb1.append(a); nb_off1.append(nb_off1[-1] + len(a)); nb2.append(b); nb_off2.append(nb_off2[-1] + len(b))
    out["fn_m1_off"] = np.array(nb_off1, np.int32); out["fn_m1"] = np.concatenate(nb1).astype(np.int32)
    out["fn_m2_off"] = np.array(nb_off2, np.int32); out["fn_m2"] = np.concatenate(nb2).astype(np.int32)
    out["radius"] = np.array([ref.compute_radius(k) for k in range(P)], np.float32)
    out["empty_mask"] = np.array([ref.find_empty_blocks(k) for k in range(P)], np.uint8)
    fnb = [ref.filter_neighbor(k) for k in range(P)]
    out["fnb_reject"] = np.array([r for r, _ in fnb], np.uint8); out["fnb_count"] = np.array([c for _, c in fnb], np.int32)
    # COptim::check (gain + quadric test of postProcess at depth >= 2) on every table patch, at the option's quad and a tight one
    chk = [ref.check(k, 2.5) for k in range(P)]
    out["check_reject"] = np.array([r for r, _ in chk], np.uint8); out["check_gain"] = np.array([g_ for _, g_ in chk], np.float32)
    out["check_reject_q01"] = np.array([ref.check(k, 0.1)[0] for k in range(P)], np.uint8)
    # the final table has already been through the filter at quad 2.5; tighter thresholds exercise filterQuad's fit
    out["fnb_quads"] = np.array([0.03, 0.1, 0.5], np.float32)
    out["fnb_reject_q"] = np.array([[ref.filter_neighbor(k, float(q))[0] for k in range(P)] for q in out["fnb_quads"]], np.uint8)
    # ---- expansion cell rules (CExpand::checkCounts / updateCounts) on the reference's own _pgrids, with pseudo-random trial
    # counters: candidates = real patches' lists with their cells jittered (so that empty, counted-out and occupied cells all occur)
    import ctypes as C
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    tn = scene.num
    dims = [ref.grid_dims(i) for i in range(tn)]
    occ, cnt0 = [], []
    for i in range(tn):
        o_ = np.zeros(dims[i][0] * dims[i][1], np.int32); ref.lib.ref_get_occupancy(i, vp(o_)); occ.append(o_)
        c_ = rng.integers(0, 6, dims[i][0] * dims[i][1]).astype(np.uint8); c_[rng.random(len(c_)) < 0.002] = 255; cnt0.append(c_)
        ref.lib.ref_set_counts(i, vp(c_))
    out["cr_occ"] = np.concatenate(occ); out["cr_counts0"] = np.concatenate(cnt0)
    NC = 3000
    coff, cim, cgr, cvoff, cvim, cvgr = [0], [], [], [0], [], []
    for k in rng.integers(0, P, NC):
        a, b = st["img_off"][k], st["img_off"][k + 1]
        keep = rng.random(b - a) < 0.8
        im = st["images"][a:b][keep]; gr = st["grids"][a:b][keep] + rng.integers(-2, 3, (int(keep.sum()), 2))
        cim.append(im); cgr.append(gr); coff.append(coff[-1] + len(im))
        a, b = st["vimg_off"][k], st["vimg_off"][k + 1]
        cvim.append(st["vimages"][a:b]); cvgr.append(st["vgrids"][a:b] + rng.integers(-1, 2, (b - a, 2))); cvoff.append(cvoff[-1] + b - a)
    out["cr_off"] = np.array(coff, np.int32); out["cr_images"] = np.concatenate(cim).astype(np.int32); out["cr_grids"] = np.concatenate(cgr).astype(np.int32)
    out["cr_voff"] = np.array(cvoff, np.int32); out["cr_vimages"] = np.concatenate(cvim).astype(np.int32); out["cr_vgrids"] = np.concatenate(cvgr).astype(np.int32)
    for depth, thr1 in ((1, 4), (2, 2)):
        ref.set_depth(depth); ref.lib.ref_set_count_threshold1(thr1)
        v = [ref.lib.ref_check_counts_raw(vp(np.ascontiguousarray(out["cr_images"][coff[k]:coff[k + 1]])), vp(np.ascontiguousarray(out["cr_grids"][coff[k]:coff[k + 1]])),
                                          int(coff[k + 1] - coff[k])) for k in range(NC)]
        out["cr_check_d%d" % depth] = np.array(v, np.uint8)
    rq = [ref.lib.ref_update_counts_raw(vp(np.ascontiguousarray(out["cr_images"][coff[k]:coff[k + 1]])), vp(np.ascontiguousarray(out["cr_grids"][coff[k]:coff[k + 1]])),
                                        int(coff[k + 1] - coff[k]), vp(np.ascontiguousarray(out["cr_vimages"][cvoff[k]:cvoff[k + 1]])),
                                        vp(np.ascontiguousarray(out["cr_vgrids"][cvoff[k]:cvoff[k + 1]])), int(cvoff[k + 1] - cvoff[k])) for k in range(NC)]
    out["cr_requeue"] = np.array(rq, np.uint8)
    cnt1 = []
    for i in range(tn):
        c_ = np.zeros(dims[i][0] * dims[i][1], np.uint8); ref.lib.ref_get_counts(i, vp(c_)); cnt1.append(c_)
    out["cr_counts1"] = np.concatenate(cnt1)
    print("cell rules: checkCounts rejects", out["cr_check_d1"].mean(), out["cr_check_d2"].mean(), "updateCounts requeues", out["cr_requeue"].mean(),
          "counters changed", int((out["cr_counts1"] != out["cr_counts0"]).sum()))
    ref.set_depth(int(out["depth_flag"]))
    # ---- filter-round stages that change the table (run last: they mutate the reference's state) -------------------------
    # (1) a fragmented table: 55 % of the patches removed in blobs, setDepthMapsVGridsVPGridsAddPatchV(1), then filterSmallGroups
    cell = np.floor(st["coords"][:, :3] * 4.0).astype(np.int64)
    blob = (cell[:, 0] * 73856093 ^ cell[:, 1] * 19349663 ^ cell[:, 2] * 83492791) % 100
    keep0 = ((blob >= 55) | (rng.random(P) < 0.03)).astype(np.uint8)
    out["frag_keep"] = keep0
    out["frag_perm"] = ref.remove_and_rebuild(keep0, additive=1)
    frag = ref.state()
    out["frag_vimg_off"] = frag["vimg_off"]; out["frag_vimages"] = frag["vimages"]; out["frag_vgrids"] = frag["vgrids"]
    out["frag_groups_survivors"] = ref.filter_small_groups(P)
    # (2) on what is left: a sixth of the patches moved 3 % towards their reference camera (they now occlude their neighbours in the
    # other images), everything rebuilt from empty _vimages (additive 0), then filterExact.  The GPU side starts from the state saved here.
    left = ref.state()
    nl = len(left["ncc"])
    ref.shift_patches((np.arange(nl) % 6 == 0).astype(np.uint8), 0.03)
    ref.remove_and_rebuild(np.ones(nl, np.uint8), additive=0)
    before = ref.state()
    for k_ in ("coords", "normals", "ncc", "dscale", "img_off", "images", "grids", "vimg_off", "vimages", "vgrids", "timages"):
        out["exact_st_" + k_] = before[k_]
    out["exact_survivors"] = ref.filter_exact(len(before["ncc"]))
    after = ref.state()
    out["exact_img_off"] = after["img_off"]; out["exact_images"] = after["images"]; out["exact_grids"] = after["grids"]
    out["exact_timages"] = after["timages"]
    print("fragmented table:", int(keep0.sum()), "kept ->", len(out["frag_perm"]), "-> small groups leave", len(out["frag_groups_survivors"]),
          "-> filterExact leaves", len(out["exact_survivors"]), "image entries", int(before["img_off"][-1]), "->", int(after["img_off"][-1]))
    path = os.path.join(HERE, "pmvs_state.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; patches", P, "visible", np.bincount(out["vis_answer"]), "neighbours",
          out["nb_pairs"][:, 2].mean(), "gain<0", (out["gains"] < 0).mean(), "filterNeighbor rejects", out["fnb_reject"].mean(),
          "mean neighbours", out["fnb_count"].mean(), "rejects at tighter quad", out["fnb_reject_q"].mean(axis=1))


if __name__ == "__main__":
    main()
