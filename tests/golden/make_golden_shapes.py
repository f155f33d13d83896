"""Generate tests/golden/pmvs_shapes.npz from the REFERENCE'S OWN CODE (oracle/_ref/libpmvs_ref.so) for the other option shapes
BASELINE.json names -- wsize 5, wsize 9, and config 2's level 0 / csize 1 -- so that the C oracle is pinned on the reference at
every shape the GPU parity tests use it at (tests/test_gpu_shapes.py), not only at the default one of pmvs_golden.npz.
Run (in the container that has /root/reference):  python tests/golden/make_golden_shapes.py

Per shape, on tests/scene_util.small_scene() with that option file and the patches make_patches(scene, oracle, 300, seed=21)
[:48] (the set the GPU tests draw): setScales; grabTex flags and textures (3 * wsize^2 floats); my_f at fixed x, computeINCC
robust / plain, setINCCs; refinePatch results and evaluation counts; preProcess on 3-image candidates, postProcess after
refinePatch.  The reference keeps its scene in a singleton, so each shape runs in its own process.
Optimiser behind `refine_*`: oracle/nm3.h via oracle/shim/nlopt.hpp (nlopt is absent) -- PARITY UNPINNED for the iterates."""
import copy
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SHAPES = {"wsize5": dict(wsize=5), "wsize9": dict(wsize=9), "level0_csize1": dict(level=0, csize=1)}
N, NTEX, NREF, NPP = 48, 16, 32, 40


def shaped_scene(name):
    from scene_util import small_scene
    sc = copy.copy(small_scene())
    sc.option = dict(sc.option)
    sc.option.update(SHAPES[name])
    return sc


def one(name, path):
    from scene_util import make_patches
    import __graft_entry__ as g
    from oracle.bindings import OracleLib, RefLib, build_ref
    synth = g.load_package().synth
    assert build_ref(), "oracle/_ref could not be built (needs /root/reference)"
    sc = shaped_scene(name)
    prefix = synth.write_scene(sc, "/tmp/pmvs_golden_shape_%s" % name)
    ref = RefLib(prefix, num=sc.num, level=sc.option["level"])
    assert ref.wsize == sc.option["wsize"] and ref.level == sc.option["level"] and ref.csize == sc.option["csize"]
    orc = OracleLib.from_scene(sc)      # only to draw the patches (dscale inputs); every output below comes from `ref`
    w = sc.option["wsize"]
    pb = make_patches(sc, orc, 300, seed=21)
    c, nm, im = pb["coords"][:N], pb["normals"][:N], pb["images"][:N]
    out = {"coords": c, "normals": nm, "images": im}
    s = [ref.set_scales(c[i], im[i]) for i in range(N)]
    out["dscale"] = np.array([x[0] for x in s], np.float32); out["ascale"] = np.array([x[1] for x in s], np.float32)
    flags, texs = [], []
    for i in range(NTEX):
        for v in range(im.shape[1]):
            f, t, _ = ref.grab_tex(c[i], nm[i], im[i, 0], im[i, v], wsize=w)
            flags.append(f); texs.append(t if f == 0 else np.zeros_like(t))
    out["tex_flag"] = np.array(flags, np.int32).reshape(NTEX, -1)
    out["tex"] = np.stack(texs).reshape(NTEX, im.shape[1], -1)
    rng = np.random.default_rng(4321)
    x = rng.normal(size=(N, 3)) * np.array([1.5, 2.0, 2.0])
    x[: N // 6] = 0.0
    out["x"] = x
    out["my_f"] = np.array([ref.my_f(c[i], nm[i], im[i], out["dscale"][i], x[i]) for i in range(N)])
    out["incc_robust"] = np.array([ref.compute_incc(c[i], nm[i], im[i], 1) for i in range(N)])
    out["incc_plain"] = np.array([ref.compute_incc(c[i], nm[i], im[i], 0) for i in range(N)])
    out["set_inccs"] = np.stack([ref.set_inccs(c[i], nm[i], im[i], 0) for i in range(N)])
    rr = [ref.refine(c[i], nm[i], im[i], out["dscale"][i]) for i in range(NREF)]
    out["refine_ok"] = np.array([r[0] for r in rr], np.uint8)
    out["refine_coord"] = np.stack([r[1] for r in rr]); out["refine_normal"] = np.stack([r[2] for r in rr])
    out["refine_ncc"] = np.array([r[3] for r in rr], np.float32); out["refine_evals"] = np.array([r[4] for r in rr], np.int32)
    pp = make_patches(sc, orc, NPP, seed=78, depth_sigma=0.01, normal_sigma=0.3)
    out["pp_coords"] = pp["coords"]; out["pp_normals"] = pp["normals"]; out["pp_images"] = pp["images"][:, :3].copy()
    cap = sc.num
    pre_v, pre_n, pre_im, pre_d, pre_a = [], [], np.full((NPP, cap), -1, np.int32), [], []
    post_v, post_n, post_im, post_gr, post_t, post_tmp = [], [], np.full((NPP, cap), -1, np.int32), np.full((NPP, cap, 2), -1, np.int32), [], []
    pin_c, pin_n, pin_ncc = np.zeros((NPP, 4), np.float32), np.zeros((NPP, 4), np.float32), np.zeros(NPP, np.float32)
    for i in range(NPP):
        v, pim, d, a = ref.pre_process(pp["coords"][i], pp["normals"][i], out["pp_images"][i])
        pre_v.append(v); pre_n.append(len(pim)); pre_im[i, : len(pim)] = pim; pre_d.append(d); pre_a.append(a)
        if v == 0:
            r = ref.refine(pp["coords"][i], pp["normals"][i], pim, d)
            pin_c[i], pin_n[i], pin_ncc[i] = r[1], r[2], r[3]
            pv, qim, qgr, qt, qtmp = ref.post_process(r[1], r[2], r[3], pim)
        else:
            pv, qim, qgr, qt, qtmp = -1, np.zeros(0, np.int32), np.zeros((0, 2), np.int32), 0, np.float32(0)
        post_v.append(pv); post_n.append(len(qim)); post_im[i, : len(qim)] = qim; post_gr[i, : len(qim)] = qgr
        post_t.append(qt); post_tmp.append(qtmp)
    out.update(pre_verdict=np.array(pre_v, np.int32), pre_n=np.array(pre_n, np.int32), pre_images=pre_im, pre_dscale=np.array(pre_d, np.float32),
               pre_ascale=np.array(pre_a, np.float32), post_in_coord=pin_c, post_in_normal=pin_n, post_in_ncc=pin_ncc,
               post_verdict=np.array(post_v, np.int32), post_n=np.array(post_n, np.int32), post_images=post_im, post_grids=post_gr,
               post_timages=np.array(post_t, np.int32), post_tmp=np.array(post_tmp, np.float32))
    np.savez(path, **out)
    print(name, "tex flags", np.bincount(out["tex_flag"].ravel() != 0), "refine ok", int(out["refine_ok"].sum()), "/", NREF,
          "evals", float(out["refine_evals"].mean()), "pre verdicts", np.bincount(out["pre_verdict"]), flush=True)


def main():
    if len(sys.argv) == 3:
        one(sys.argv[1], sys.argv[2])
        return
    merged = {}
    for name in SHAPES:
        tmp = "/tmp/pmvs_shape_%s.npz" % name
        subprocess.run([sys.executable, os.path.abspath(__file__), name, tmp], check=True)
        d = np.load(tmp)
        for k in d.files:
            merged["%s__%s" % (name, k)] = d[k]
        os.remove(tmp)
    path = os.path.join(HERE, "pmvs_shapes.npz")
    np.savez_compressed(path, **merged)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
