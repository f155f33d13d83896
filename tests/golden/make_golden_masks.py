"""Generate tests/golden/pmvs_masks.npz: the REFERENCE'S answers on the masks / edges / bimages.dat part of the contract
(tests/scene_util.mask_variants()).  Per variant: the working-level mask and edge map of every image as the reference built
them (readPGMImage / readPBMImage, buildMask / buildEdge, CImage::setEdge), the point gate of expandSub / collectCandidates /
postProcess, CPhoto::getEdge, COptim::removeImagesEdge, preProcess / postProcess image sets under the maps, the features
of three images, and the cloud the reference BINARY writes (CPU 1).
Run:  python tests/golden/make_golden_masks.py   (needs /root/reference)"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from scene_util import make_patches, mask_variants, small_scene, write_mask_variant  # noqa: E402
import __graft_entry__ as g  # noqa: E402
from oracle.bindings import OracleLib, RefLib, build_ref  # noqa: E402


def probe_points(scene, n, seed):
    """points on and around the sphere, some far outside every frame (bounding images, z <= 0)"""
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    r = np.where(rng.random(n) < 0.7, 1.0 + 0.02 * rng.normal(size=n), rng.uniform(0.2, 4.0, n))
    X = np.ones((n, 4), np.float32)
    X[:, :3] = (d * r[:, None]).astype(np.float32)
    return X


def main():
    assert build_ref()
    scene = small_scene()
    out = {"scene_sha256": np.frombuffer(bytes.fromhex(scene.sha256()), np.uint8)}
    orc = OracleLib.from_scene(scene)
    pb = make_patches(scene, orc, 300, seed=77, depth_sigma=0.01, normal_sigma=0.3)
    n0 = np.random.default_rng(5).integers(2, 4, 300).astype(np.int32)
    X = probe_points(scene, 3000, 11)
    out["points"] = X
    out["patch_coords"] = pb["coords"]; out["patch_normals"] = pb["normals"]; out["patch_images"] = pb["images"]; out["patch_n0"] = n0
    for name in mask_variants():
        prefix = write_mask_variant(scene, name, "/tmp/pmvs_golden_masks_%s" % name, cpu=1)
        # the reference singleton serves one scene per process: probe in a child
        ref = RefLib(prefix, num=scene.num, level=scene.option["level"], skip_features=False)
        lvl = scene.option["level"]
        for which, tag in ((0, "mask"), (1, "edge")):
            present = []
            maps = []
            for i in range(scene.num):
                m = ref.map_bytes(i, which, lvl)
                present.append(0 if m is None else 1)
                if m is not None:
                    maps.append(np.packbits(m.ravel() != 0))
            out["%s_%s_present" % (name, tag)] = np.array(present, np.uint8)
            out["%s_%s_bits" % (name, tag)] = np.concatenate(maps) if maps else np.zeros(0, np.uint8)
        out["%s_gate" % name] = np.array([ref.mask_gate(x) for x in X], np.uint8)
        out["%s_edge" % name] = np.array([[ref.get_edge(x, i) for i in range(scene.num)] for x in X[:600]], np.uint8)
        rm = [ref.remove_images_edge(pb["coords"][k], np.arange(scene.num, dtype=np.int32)[(np.arange(scene.num) + k) % 3 != 0]) for k in range(300)]
        out["%s_rm_off" % name] = np.cumsum([0] + [len(r) for r in rm]).astype(np.int32)
        out["%s_rm" % name] = np.concatenate(rm).astype(np.int32) if sum(len(r) for r in rm) else np.zeros(0, np.int32)
        pre_v, pre_off, pre_im, pre_d = [], [0], [], []
        post_v, post_off, post_im = [], [0], []
        for k in range(300):
            v, im, d, a = ref.pre_process(pb["coords"][k], pb["normals"][k], pb["images"][k, : n0[k]], cap=scene.num)
            pre_v.append(v); pre_im.append(im); pre_off.append(pre_off[-1] + len(im)); pre_d.append(d)
            # postProcess straight on the candidate (no refinement: the gate and the image sets are what is pinned here)
            v2, im2, gr2, t2, tmp2 = ref.post_process(pb["coords"][k], pb["normals"][k], 0.9, pb["images"][k, :3], cap=scene.num)
            post_v.append(v2); post_im.append(im2); post_off.append(post_off[-1] + len(im2))
        out["%s_pre_verdict" % name] = np.array(pre_v, np.int32); out["%s_pre_off" % name] = np.array(pre_off, np.int32)
        out["%s_pre_images" % name] = np.concatenate(pre_im).astype(np.int32); out["%s_pre_dscale" % name] = np.array(pre_d, np.float32)
        out["%s_post_verdict" % name] = np.array(post_v, np.int32); out["%s_post_off" % name] = np.array(post_off, np.int32)
        out["%s_post_images" % name] = np.concatenate(post_im).astype(np.int32) if sum(len(r) for r in post_im) else np.zeros(0, np.int32)
        for i in (0, 1, 6):
            xy, resp, ty = ref.detect_features(i)
            out["%s_feat%d_xy" % (name, i)] = xy; out["%s_feat%d_resp" % (name, i)] = resp; out["%s_feat%d_type" % (name, i)] = ty
        ref.lib.ref_close()
        # the reference binary on the same directory
        p = subprocess.run([os.path.join(ROOT, "oracle/_ref/pmvs3_ref"), prefix, "option.txt", "PSET"], stdout=subprocess.DEVNULL,
                           stderr=subprocess.PIPE, text=True, check=True)
        pts = np.loadtxt(prefix + "models/option.txt.pset", dtype=np.float32).reshape(-1, 6)
        out["%s_pset" % name] = pts
        print(name, "gate pass", out["%s_gate" % name].mean(), "edge pass", out["%s_edge" % name].mean(), "pre keep", 1 - np.mean(pre_v),
              "post keep", 1 - np.mean(post_v), "removeImagesEdge mean length", np.diff(out["%s_rm_off" % name]).mean(),
              "features", [len(out["%s_feat%d_resp" % (name, i)]) for i in (0, 1, 6)], "patches", len(pts))
    path = os.path.join(HERE, "pmvs_masks.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
