"""Shared test inputs: a small deterministic scene (cached on disk) and seeded patch batches."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CACHE = os.environ.get("PMVS_TEST_CACHE", "/tmp/pmvs_b200_test_cache")


def _synth():
    import __graft_entry__ as g
    return g.load_package().synth


def small_scene(views: int = 16, width: int = 320, height: int = 240):
    """Sphere scene, `views` cameras, level 1 / csize 2 / wsize 7 / minImageNum 3 (config-1 shaped, smaller)."""
    synth = _synth()
    s = synth.sphere_scene(views=views, width=width, height=height, seed=0)
    path = os.path.join(CACHE, "%s_%dx%d.npy" % (s.name, width, height))
    if os.path.exists(path):
        arr = np.load(path)
        s.images = [arr[i] for i in range(views)]
    else:
        synth.render(s)
        os.makedirs(CACHE, exist_ok=True)
        np.save(path, np.stack(s.images))
    return s


def pick_views(scene, X, N, k):
    """Reference image = most frontal camera, then the next k-1 by angle to the normal."""
    d = scene.C - X[None, :3]
    d = d / np.linalg.norm(d, axis=1, keepdims=True)
    order = np.argsort(-(d @ N[:3]), kind="stable")
    return order[:k].astype(np.int32)


def make_patches(scene, orc, n, seed, k=None, depth_sigma=0.004, normal_sigma=0.12):
    """n seed patches near the true surface with perturbed depth/normal, k views each, dscale from setScales."""
    synth = _synth()
    k = k or min(6, scene.num)
    rng = np.random.default_rng(seed)
    pts, nrm = synth.surface_samples(scene, 4 * n, seed)
    pts = pts.numpy(); nrm = nrm.numpy()
    if scene.kind == "sphere":  # the camera ring sees the equatorial band; keep 1 in 8 patches from the caps
        keep = (np.abs(pts[:, 2]) < 0.55) | (np.arange(len(pts)) % 8 == 0)
        pts, nrm = pts[keep][:n], nrm[keep][:n]
    pts, nrm = pts[:n], nrm[:n]
    coords = np.zeros((n, 4), np.float32); normals = np.zeros((n, 4), np.float32)
    images = np.zeros((n, k), np.int32); dscales = np.zeros(n, np.float32)
    for i in range(n):
        X = pts[i] * (1.0 + rng.normal() * depth_sigma)
        N = nrm[i] + rng.normal(size=3) * normal_sigma
        N = N / np.linalg.norm(N)
        coords[i, :3] = X; coords[i, 3] = 1.0
        normals[i, :3] = N
        images[i] = pick_views(scene, coords[i], normals[i], k)
        dscales[i] = orc.set_scales(coords[i], images[i])[0]
    return dict(coords=coords, normals=normals, images=images, dscales=dscales)


def option_variants():
    """Option-file variants of small_scene() that exercise the rest of the option contract (source/pmvs/option.cpp)."""
    return {
        "oimages": {"timages": (-1, 0, 12), "oimages": (-1, 12, 16)},             # four non-target images
        "visdata": {"useVisData": 1},                                              # vis.dat: +-3 neighbours on the ring
        "sequence": {"sequence": 2},                                               # only images within 2 of the reference
        "enumerated": {"timages": (8, 0, 2, 4, 6, 8, 10, 12, 14), "oimages": (4, 1, 5, 9, 13), "csize": 1, "level": 1},
    }


def write_variant(scene, name, prefix, cpu):
    """small_scene() with the variant's option file (and vis.dat) under `prefix`; returns the prefix with trailing '/'."""
    import copy
    synth = _synth()
    sc = copy.copy(scene)
    sc.option = dict(scene.option)
    sc.option.update(option_variants()[name])
    sc.option["CPU"] = cpu
    out = synth.write_scene(sc, prefix)
    n = scene.num
    with open(out + "vis.dat", "w") as f:
        f.write("VISDATA\n%d\n" % n)
        for i in range(n):
            nb = sorted({(i + d) % n for d in (-3, -2, -1, 1, 2, 3)})
            f.write("%d %d  %s\n" % (i, len(nb), " ".join(str(v) for v in nb)))
    return out


SKE_TWO_CLUSTERS = "SKE\n16 2\n8 2\n0 1 2 3 4 5 6 7 \n8 15 \n8 2\n8 9 10 11 12 13 14 15 \n0 7 \n"


def write_clusters(scene, prefix, cpu):
    """small_scene() laid out like a CMVS output directory: images, cameras, vis.dat (+-3 ring neighbours) and a ske.dat
    with two clusters (the halves of the ring as target images, two images of the other half as `oimages`), in the format
    CMVS::CBundle::writeGroups writes (source/cmvs/bundle.cpp:1462-1481).  genOption turns it into option-0000/0001."""
    out = write_variant(scene, "visdata", prefix, cpu)
    with open(out + "ske.dat", "w") as f:
        f.write(SKE_TWO_CLUSTERS)
    return out
