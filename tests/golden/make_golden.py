"""Generate tests/golden/pmvs_golden.npz from the REFERENCE'S OWN CODE (oracle/_ref/libpmvs_ref.so, built
from /root/reference by oracle/Makefile).  The reference ships no tests, fixtures or golden vectors
(SURVEY.md section 4), so these known-answer vectors are produced here, once, in the container that has
/root/reference, and committed.  Run:  python tests/golden/make_golden.py

Scene: tests/scene_util.small_scene() (sphere, 16 views 320x240, level 1, csize 2, wsize 7, minImageNum 3);
its SHA-256 is stored so a non-reproducible scene is detected before any comparison.
Optimiser behind `refine_*`: oracle/nm3.h via oracle/shim/nlopt.hpp (nlopt is absent) -- PARITY UNPINNED for
the iterates; every other vector pins reference arithmetic exactly.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from scene_util import make_patches, small_scene  # noqa: E402
import __graft_entry__ as g  # noqa: E402
from oracle.bindings import OracleLib, RefLib, build_ref  # noqa: E402


def main():
    synth = g.load_package().synth
    assert build_ref(), "oracle/_ref could not be built (needs /root/reference)"
    scene = small_scene()
    prefix = synth.write_scene(scene, "/tmp/pmvs_golden_scene")
    ref = RefLib(prefix, num=scene.num, level=scene.option["level"])
    orc = OracleLib.from_scene(scene)  # only to produce dscale inputs for make_patches; outputs come from `ref`
    out = {"scene_sha256": np.frombuffer(bytes.fromhex(scene.sha256()), np.uint8)}
    rng = np.random.default_rng(1234)

    # cameras
    for k in ("P", "centre", "oaxis", "xaxis", "yaxis", "zaxis", "ipscale"):
        out["cam_" + k] = np.stack([np.atleast_1d(ref.camera(i, 1)[k]) for i in range(scene.num)])
    # pyramids: digest per (image, level) + one full small level
    dig = []
    for i in range(scene.num):
        for l in range(scene.option["level"] + 3):
            dig.append(np.frombuffer(hashlib.sha256(ref.image(i, l).tobytes()).digest(), np.uint8))
    out["pyr_sha256"] = np.stack(dig)
    out["pyr_img3_level3"] = ref.image(3, 3)

    pb = make_patches(scene, orc, 240, seed=77, depth_sigma=0.006, normal_sigma=0.2)
    n = len(pb["coords"])
    out.update({"coords": pb["coords"], "normals": pb["normals"], "images": pb["images"]})
    # projections / units
    pimg = rng.integers(0, scene.num, n).astype(np.int32)
    out["proj_image"] = pimg
    for l in (0, 1, 2):
        out["proj_l%d" % l] = np.stack([ref.project(pimg[i], pb["coords"][i], l) for i in range(n)])
    out["unit"] = np.array([ref.get_unit(pimg[i], pb["coords"][i]) for i in range(n)], np.float32)
    # setScales
    sc = [ref.set_scales(pb["coords"][i], pb["images"][i]) for i in range(n)]
    out["dscale"] = np.array([s[0] for s in sc], np.float32)
    out["ascale"] = np.array([s[1] for s in sc], np.float32)
    # grabTex: first 40 patches, all 6 views
    flags, texs = [], []
    for i in range(40):
        for v in range(pb["images"].shape[1]):
            f, t, _ = ref.grab_tex(pb["coords"][i], pb["normals"][i], pb["images"][i, 0], pb["images"][i, v])
            flags.append(f); texs.append(t if f == 0 else np.zeros_like(t))
    out["tex_flag"] = np.array(flags, np.int32).reshape(40, -1)
    out["tex"] = np.stack(texs).reshape(40, pb["images"].shape[1], -1)
    # normalize / dot on grabbed textures
    good = [t for t, f in zip(texs, flags) if f == 0][:30]
    out["norm_in"] = np.stack(good)
    out["norm_out"] = np.stack([ref.normalize(t) for t in good])
    out["dot_out"] = np.array([ref.dot(out["norm_out"][k], out["norm_out"][(k + 1) % len(good)]) for k in range(len(good))], np.float32)
    # encode / decode / my_f / computeINCC / setINCCs
    x = rng.normal(size=(n, 3)) * np.array([1.5, 2.0, 2.0])
    x[: n // 6] = 0.0
    out["x"] = x
    out["encode"] = np.stack([ref.encode(pb["coords"][i], pb["normals"][i], pb["images"][i], out["dscale"][i]) for i in range(n)])
    dec = [ref.decode(pb["coords"][i], pb["normals"][i], pb["images"][i], out["dscale"][i], x[i]) for i in range(n)]
    out["decode_coord"] = np.stack([d[0] for d in dec]); out["decode_normal"] = np.stack([d[1] for d in dec])
    out["my_f"] = np.array([ref.my_f(pb["coords"][i], pb["normals"][i], pb["images"][i], out["dscale"][i], x[i]) for i in range(n)])
    out["incc_robust"] = np.array([ref.compute_incc(pb["coords"][i], pb["normals"][i], pb["images"][i], 1) for i in range(n)])
    out["incc_plain"] = np.array([ref.compute_incc(pb["coords"][i], pb["normals"][i], pb["images"][i], 0) for i in range(n)])
    out["set_inccs"] = np.stack([ref.set_inccs(pb["coords"][i], pb["normals"][i], pb["images"][i], 0) for i in range(n)])
    out["set_inccs_matrix"] = np.stack([ref.set_inccs_matrix(pb["coords"][i], pb["normals"][i], pb["images"][i], 1) for i in range(n)])
    # refinePatch (first 120)
    m = 120
    rr = [ref.refine(pb["coords"][i], pb["normals"][i], pb["images"][i], out["dscale"][i]) for i in range(m)]
    out["refine_ok"] = np.array([r[0] for r in rr], np.uint8)
    out["refine_coord"] = np.stack([r[1] for r in rr]); out["refine_normal"] = np.stack([r[2] for r in rr])
    out["refine_ncc"] = np.array([r[3] for r in rr], np.float32); out["refine_evals"] = np.array([r[4] for r in rr], np.int32)
    # preProcess from 3-image candidates (as expansion hands them over), then postProcess after refine
    pp = make_patches(scene, orc, 160, seed=78, depth_sigma=0.01, normal_sigma=0.3)
    out["pp_coords"] = pp["coords"]; out["pp_normals"] = pp["normals"]; out["pp_images"] = pp["images"][:, :3].copy()
    cap = scene.num
    pre_v, pre_n, pre_im, pre_d, pre_a = [], [], np.full((160, cap), -1, np.int32), [], []
    post_v, post_n, post_im, post_gr, post_t, post_tmp = [], [], np.full((160, cap), -1, np.int32), np.full((160, cap, 2), -1, np.int32), [], []
    post_in_coord, post_in_normal, post_in_ncc = np.zeros((160, 4), np.float32), np.zeros((160, 4), np.float32), np.zeros(160, np.float32)
    for i in range(160):
        v, im, d, a = ref.pre_process(pp["coords"][i], pp["normals"][i], out["pp_images"][i])
        pre_v.append(v); pre_n.append(len(im)); pre_im[i, : len(im)] = im; pre_d.append(d); pre_a.append(a)
        if v == 0:
            r = ref.refine(pp["coords"][i], pp["normals"][i], im, d)
            post_in_coord[i], post_in_normal[i], post_in_ncc[i] = r[1], r[2], r[3]
            pv, pim, pgr, pt, ptmp = ref.post_process(r[1], r[2], r[3], im)
        else:
            pv, pim, pgr, pt, ptmp = -1, np.zeros(0, np.int32), np.zeros((0, 2), np.int32), 0, np.float32(0)
        post_v.append(pv); post_n.append(len(pim)); post_im[i, : len(pim)] = pim; post_gr[i, : len(pim)] = pgr
        post_t.append(pt); post_tmp.append(ptmp)
    out.update(pre_verdict=np.array(pre_v, np.int32), pre_n=np.array(pre_n, np.int32), pre_images=pre_im,
               pre_dscale=np.array(pre_d, np.float32), pre_ascale=np.array(pre_a, np.float32),
               post_in_coord=post_in_coord, post_in_normal=post_in_normal, post_in_ncc=post_in_ncc,
               post_verdict=np.array(post_v, np.int32), post_n=np.array(post_n, np.int32), post_images=post_im, post_grids=post_gr,
               post_timages=np.array(post_t, np.int32), post_tmp=np.array(post_tmp, np.float32))
    path = os.path.join(HERE, "pmvs_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", "pre verdicts", np.bincount(out["pre_verdict"]), "post verdicts",
          np.unique(out["post_verdict"], return_counts=True))


if __name__ == "__main__":
    main()
