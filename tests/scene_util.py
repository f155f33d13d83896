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


def mask_maps(scene, which):
    """Deterministic per-image maps for the mask / edge contract (source/image/image.cpp:146-176): which = 0 masks (the top of
    every frame is outside, stored with gray values on both sides of the 127 threshold), 1 edge files (a band
    pattern with values around the `1 <` threshold).  Images 3, 7, 11, 15 have no mask file; odd images with a mask get a
    binary PBM (the reference reads it as one continuous bit stream), the others a PGM.  Returns {index: (kind, array)}."""
    h, w = scene.height, scene.width
    yy, xx = np.mgrid[0:h, 0:w]
    out = {}
    for i in range(scene.num):
        if which == 0:
            if i % 4 == 3:
                continue
            # a silhouette-style mask must keep everything any OTHER camera sees (the gate rejects a point that falls outside
            # the mask of ANY image it projects into): cut the top cap of the sphere and a small off-object corner
            inside = (yy >= h * (0.18 + 0.03 * np.cos(i))) & ~((xx < 0.1 * w) & (yy > 0.9 * h))
            if i % 2 == 1:
                out[i] = ("pbm", inside)
            else:
                out[i] = ("pgm", np.where(inside, 128 + (xx + yy + i) % 100, (xx * 3 + yy) % 128).astype(np.uint8))
        else:
            if i % 3 == 2:
                continue
            inside = ((xx + 2 * yy + 7 * i) % 23) < 19
            out[i] = ("pgm", np.where(inside, 2 + (xx + i) % 200, (xx + yy) % 2).astype(np.uint8))
    return out


def write_map_files(prefix, folder, maps):
    os.makedirs(prefix + folder, exist_ok=True)
    for i, (kind, a) in maps.items():
        h, w = a.shape
        if kind == "pgm":
            with open(prefix + "%s/%08d.pgm" % (folder, i), "wb") as f:
                f.write(b"P5\n# written by tests/scene_util.py\n%d %d\n255\n" % (w, h))
                f.write(np.ascontiguousarray(a, dtype=np.uint8).tobytes())
        else:   # P4 as CImage::writePBMImage lays it out: one bit stream without row padding, set bit = outside
            bits = np.packbits((~a.astype(bool)).ravel())
            with open(prefix + "%s/%08d.pbm" % (folder, i), "wb") as f:
                f.write(b"P4\n%d %d\n" % (w, h))
                f.write(bits.tobytes())


def mask_variants():
    """Scene variants for the masks / edges / bimages.dat part of the option contract: name -> (option updates, files)."""
    return {
        "masked": ({"useBound": 1}, ("masks", "bimages")),     # masks/ + useBound with bimages.dat = images 0 and 8
        "edgefiles": ({}, ("edges",)),                          # edges/%08d.pgm
        "setedge": ({"setEdge": 2.5}, ()),                       # edge maps computed from the images (CImage::setEdge)
        "allmaps": ({"useBound": 1, "setEdge": 0.0}, ("masks", "edges", "bimages")),
    }


def write_mask_variant(scene, name, prefix, cpu):
    import copy
    synth = _synth()
    upd, files = mask_variants()[name]
    sc = copy.copy(scene)
    sc.option = dict(scene.option)
    sc.option.update(upd)
    sc.option["CPU"] = cpu
    out = synth.write_scene(sc, prefix)
    if "masks" in files:
        write_map_files(out, "masks", mask_maps(scene, 0))
    if "edges" in files:
        write_map_files(out, "edges", mask_maps(scene, 1))
    if "bimages" in files:
        with open(out + "bimages.dat", "w") as f:
            f.write("3\n0 8 99\n")    # 99 is not a target image: ignored (option.cpp:318-321)
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
