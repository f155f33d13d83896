"""Generate tests/golden/pmvs_seed.npz: the REFERENCE'S seed-stage answers on tests/scene_util.small_scene() before any patch
exists -- the features CSeed holds per image, COptim::collectImages' list and CSeed::collectCandidates (epipolar cell walk,
point-to-line distance, triangulation, _response) for every feature of three reference images, once with every cell open and
once with a pseudo-random third of the cells of all target images closed through _counts (CSeed::canAdd).
Run:  python tests/golden/make_golden_seed.py   (needs /root/reference)"""
import ctypes as C
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

REF_IMAGES = (0, 5, 11)


def closed_cells(scene, ref):
    """per target image: uint8 counts with _countThreshold2 (2) in a pseudo-random third of the cells"""
    out = []
    for i in range(scene.num):
        gw, gh = ref.grid_dims(i)
        rng = np.random.default_rng(100 + i)
        out.append(np.where(rng.random(gw * gh) < 0.33, 2, 0).astype(np.uint8))
    return out


def collect(ref, scene, index, feats_per_cell):
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    cap = 4096
    oi = np.zeros(cap, np.int32); oxy = np.zeros((cap, 2), np.float32); co = np.zeros((cap, 4), np.float32); rp = np.zeros(cap, np.float32)
    rows = []       # (cell, p, n)
    cand = []
    gw, gh = ref.grid_dims(index)
    for cell in range(gw * gh):
        if not ref.lib.ref_can_add(index, cell % gw, cell // gw):
            continue
        for p in range(feats_per_cell[cell]):
            n = ref.lib.ref_collect_candidates(index, cell, p, vp(oi), vp(oxy), vp(co), vp(rp), cap)
            assert n <= cap
            rows.append((cell, p, n))
            for k in range(n):
                cand.append((oi[k], oxy[k, 0], oxy[k, 1], co[k, 0], co[k, 1], co[k, 2], co[k, 3], rp[k]))
    return np.array(rows, np.int32).reshape(-1, 3), np.array(cand, np.float32).reshape(-1, 8)


def main():
    synth = g.load_package().synth
    assert build_ref()
    scene = small_scene()
    prefix = synth.write_scene(scene, "/tmp/pmvs_golden_seed_scene")
    ref = RefLib(prefix, num=scene.num, level=scene.option["level"], skip_features=False)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    out = {"scene_sha256": np.frombuffer(bytes.fromhex(scene.sha256()), np.uint8), "ref_images": np.array(REF_IMAGES, np.int32)}
    fcount = {}
    for i in range(scene.num):
        xy = np.zeros((20000, 2), np.float32); ty = np.zeros(20000, np.int32)
        n = ref.lib.ref_features(i, vp(xy), vp(ty), 20000)
        out["feat%d_xy" % i] = xy[:n].copy(); out["feat%d_type" % i] = ty[:n].copy()
        gw, gh = ref.grid_dims(i)
        cells = (np.floor(xy[:n, 1] + 0.5).astype(np.int64) // scene.option["csize"]) * gw + np.floor(xy[:n, 0] + 0.5).astype(np.int64) // scene.option["csize"]
        fcount[i] = np.bincount(cells, minlength=gw * gh)
    for index in REF_IMAGES:
        v = np.zeros(16, np.int32)
        n = ref.lib.ref_collect_images(index, vp(v), 16)
        out["views%d" % index] = v[:n].copy()
    for tag in ("open", "closed"):
        if tag == "closed":
            counts = closed_cells(scene, ref)
            for i in range(scene.num):
                ref.lib.ref_set_counts(i, vp(counts[i]))
            out["closed_counts"] = np.concatenate(counts)
        for index in REF_IMAGES:
            rows, cand = collect(ref, scene, index, fcount[index])
            out["%s_rows%d" % (tag, index)] = rows; out["%s_cand%d" % (tag, index)] = cand
            print(tag, "image", index, "features searched", len(rows), "candidates", len(cand), "max per feature", rows[:, 2].max() if len(rows) else 0)
    path = os.path.join(HERE, "pmvs_seed.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
