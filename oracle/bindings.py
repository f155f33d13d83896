"""oracle/bindings.py -- TEST INFRASTRUCTURE.

ctypes wrappers with one method vocabulary over
  * RefLib   : oracle/_ref/libpmvs_ref.so   (the reference's own objects, oracle/ref_probe.cpp)
  * OracleLib: oracle/build/libpmvs_oracle.so (the plain-C restatement, oracle/pmvs_oracle.c)
so that a test can run the same call against both.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libpmvs_ref.so")
REF_BIN = os.path.join(HERE, "_ref", "pmvs3_ref")
ORACLE_SO = os.path.join(HERE, "build", "libpmvs_oracle.so")

f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def build_oracle(force: bool = False) -> str:
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(HERE, "pmvs_oracle.c")):
        subprocess.check_call(["make", "-C", HERE, "oracle"], stdout=subprocess.DEVNULL)
    return ORACLE_SO


def build_ref() -> str | None:
    """Build oracle/_ref from /root/reference when it is present; otherwise use the prebuilt files."""
    if os.path.isdir("/root/reference/source/pmvs"):
        subprocess.check_call(["make", "-C", HERE, "-j8", "ref"], stdout=subprocess.DEVNULL)
    return REF_SO if os.path.exists(REF_SO) else None


def _f4(v):
    a = np.zeros(4, dtype=np.float32)
    v = np.asarray(v, dtype=np.float32).ravel()
    a[: len(v)] = v
    return a


def _imgs(images):
    return np.ascontiguousarray(np.asarray(images, dtype=np.int32).ravel())


class _Common:
    """Methods shared by both back ends; subclasses provide self.lib, self.pfx and self._h() (handle args)."""

    def _fn(self, name):
        return getattr(self.lib, self.pfx + name)

    def image(self, index, level):
        w, h = C.c_int(), C.c_int()
        self._fn("image_dims")(*self._h(), int(index), level, C.byref(w), C.byref(h))
        out = np.zeros((h.value, w.value, 3), dtype=np.uint8)
        if out.size:
            self._fn("image_bytes")(*self._h(), int(index), level, out.ctypes.data_as(C.c_void_p))
        return out

    def camera(self, index, level=0):
        P = np.zeros(12, np.float32); ce = np.zeros(4, np.float32); oa = np.zeros(4, np.float32)
        xa = np.zeros(3, np.float32); ya = np.zeros(3, np.float32); za = np.zeros(3, np.float32)
        ips = C.c_float()
        fp = lambda a: a.ctypes.data_as(C.c_void_p)
        self._fn("camera")(*self._h(), int(index), level, fp(P), fp(ce), fp(oa), fp(xa), fp(ya), fp(za), C.byref(ips))
        return dict(P=P.reshape(3, 4), centre=ce, oaxis=oa, xaxis=xa, yaxis=ya, zaxis=za, ipscale=np.float32(ips.value))

    def project(self, index, coord, level):
        out = np.zeros(3, np.float32)
        self._fn("project")(*self._h(), int(index), _f4(coord).ctypes.data_as(C.c_void_p), level, out.ctypes.data_as(C.c_void_p))
        return out

    def get_unit(self, index, coord):
        f = self._fn("get_unit"); f.restype = C.c_float
        return np.float32(f(*self._h(), int(index), _f4(coord).ctypes.data_as(C.c_void_p)))

    def get_color(self, index, x, y, level):
        out = np.zeros(3, np.float32)
        self._fn("get_color")(*self._h(), int(index), C.c_float(x), C.c_float(y), level, out.ctypes.data_as(C.c_void_p))
        return out

    def get_paxes(self, index, coord, normal):
        px = np.zeros(4, np.float32); py = np.zeros(4, np.float32)
        self._fn("get_paxes")(*self._h(), int(index), _f4(coord).ctypes.data_as(C.c_void_p), _f4(normal).ctypes.data_as(C.c_void_p),
                              px.ctypes.data_as(C.c_void_p), py.ctypes.data_as(C.c_void_p))
        return px, py

    def normalize(self, tex):
        t = np.ascontiguousarray(tex, dtype=np.float32).copy()
        self._fn("normalize")(t.ctypes.data_as(C.c_void_p), t.size)
        return t

    def dot(self, a, b):
        f = self._fn("dot"); f.restype = C.c_float
        a = np.ascontiguousarray(a, dtype=np.float32); b = np.ascontiguousarray(b, dtype=np.float32)
        return np.float32(f(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), a.size))

    def set_scales(self, coord, images):
        im = _imgs(images); d = C.c_float(); a = C.c_float()
        self._fn("set_scales")(*self._h(), _f4(coord).ctypes.data_as(C.c_void_p), im.ctypes.data_as(C.c_void_p), len(im), C.byref(d), C.byref(a))
        return np.float32(d.value), np.float32(a.value)

    def encode(self, coord, normal, images, dscale):
        im = _imgs(images); x = np.zeros(3, np.float64)
        self._fn("encode")(*self._h(), _f4(coord).ctypes.data_as(C.c_void_p), _f4(normal).ctypes.data_as(C.c_void_p),
                           im.ctypes.data_as(C.c_void_p), len(im), C.c_float(dscale), x.ctypes.data_as(C.c_void_p))
        return x

    def decode(self, coord, normal, images, dscale, x):
        im = _imgs(images); x = np.ascontiguousarray(x, dtype=np.float64)
        oc = np.zeros(4, np.float32); on = np.zeros(4, np.float32)
        self._fn("decode")(*self._h(), _f4(coord).ctypes.data_as(C.c_void_p), _f4(normal).ctypes.data_as(C.c_void_p),
                           im.ctypes.data_as(C.c_void_p), len(im), C.c_float(dscale), x.ctypes.data_as(C.c_void_p),
                           oc.ctypes.data_as(C.c_void_p), on.ctypes.data_as(C.c_void_p))
        return oc, on

    def my_f(self, coord, normal, images, dscale, x):
        f = self._fn("my_f"); f.restype = C.c_double
        im = _imgs(images); x = np.ascontiguousarray(x, dtype=np.float64)
        return f(*self._h(), _f4(coord).ctypes.data_as(C.c_void_p), _f4(normal).ctypes.data_as(C.c_void_p),
                 im.ctypes.data_as(C.c_void_p), len(im), C.c_float(dscale), x.ctypes.data_as(C.c_void_p))

    def compute_incc(self, coord, normal, images, robust=1):
        f = self._fn("compute_incc"); f.restype = C.c_double
        im = _imgs(images)
        return f(*self._h(), _f4(coord).ctypes.data_as(C.c_void_p), _f4(normal).ctypes.data_as(C.c_void_p),
                 im.ctypes.data_as(C.c_void_p), len(im), robust)

    def set_inccs(self, coord, normal, images, robust=0):
        im = _imgs(images); out = np.zeros(len(im), np.float32)
        self._fn("set_inccs")(*self._h(), _f4(coord).ctypes.data_as(C.c_void_p), _f4(normal).ctypes.data_as(C.c_void_p),
                              im.ctypes.data_as(C.c_void_p), len(im), robust, out.ctypes.data_as(C.c_void_p))
        return out

    def set_inccs_matrix(self, coord, normal, images, robust=1):
        im = _imgs(images); out = np.zeros((len(im), len(im)), np.float32)
        self._fn("set_inccs_matrix")(*self._h(), _f4(coord).ctypes.data_as(C.c_void_p), _f4(normal).ctypes.data_as(C.c_void_p),
                                     im.ctypes.data_as(C.c_void_p), len(im), robust, out.ctypes.data_as(C.c_void_p))
        return out

    def refine(self, coord, normal, images, dscale):
        """-> (ok, coord, normal, ncc, evals)"""
        im = _imgs(images); c = _f4(coord).copy(); n = _f4(normal).copy()
        ncc = C.c_float(-1.0); ev = C.c_int(0)
        ok = self._fn("refine")(*self._h(), c.ctypes.data_as(C.c_void_p), n.ctypes.data_as(C.c_void_p),
                                im.ctypes.data_as(C.c_void_p), len(im), C.c_float(dscale), C.byref(ncc), C.byref(ev))
        return int(ok), c, n, np.float32(ncc.value), ev.value

    def refine_batch(self, coords, normals, images, dscales, threads=1):
        """coords/normals (P,4) f32, images (P,V) i32 -> dict(coords, normals, ncc, evals, ok, seconds)"""
        f = self._fn("refine_batch"); f.restype = C.c_double
        co = np.ascontiguousarray(coords, dtype=np.float32).copy(); no = np.ascontiguousarray(normals, dtype=np.float32).copy()
        im = np.ascontiguousarray(images, dtype=np.int32); ds = np.ascontiguousarray(dscales, dtype=np.float32)
        P, V = im.shape
        ncc = np.full(P, -1.0, np.float32); ev = np.zeros(P, np.int32); ok = np.zeros(P, np.uint8)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        secs = f(*self._h(), P, V, vp(co), vp(no), vp(im), vp(ds), vp(ncc), vp(ev), vp(ok), threads)
        return dict(coords=co, normals=no, ncc=ncc, evals=ev, ok=ok, seconds=secs)

    def pre_process(self, coord, normal, images, cap=256):
        """-> (verdict, images, dscale, ascale)"""
        im = np.zeros(cap, np.int32); src = _imgs(images); im[: len(src)] = src
        n = C.c_int(len(src)); d = C.c_float(); a = C.c_float()
        r = self._fn("pre_process")(*self._h(), _f4(coord).ctypes.data_as(C.c_void_p), _f4(normal).ctypes.data_as(C.c_void_p),
                                    im.ctypes.data_as(C.c_void_p), C.byref(n), cap, C.byref(d), C.byref(a))
        return int(r), im[: n.value].copy(), np.float32(d.value), np.float32(a.value)


class OracleLib(_Common):
    pfx = "pmvso_"

    def __init__(self, num, tnum=None, level=1, csize=2, wsize=7, min_image_num=3, threshold=0.7, max_angle_deg=10.0):
        self.lib = C.CDLL(build_oracle())
        self.lib.pmvso_create.restype = C.c_void_p
        self.wsize = int(wsize)
        self.ctx = C.c_void_p(self.lib.pmvso_create(num, num if tnum is None else tnum, level, csize, wsize, min_image_num,
                                                    C.c_float(threshold), C.c_float(max_angle_deg)))
        self.num = num
        self.level = level

    def _h(self):
        return (self.ctx,)

    @classmethod
    def from_scene(cls, scene):
        o = scene.option
        self = cls(scene.num, level=o["level"], csize=o["csize"], wsize=o["wsize"], min_image_num=o["minImageNum"],
                   threshold=o["threshold"])
        for i in range(scene.num):
            self.set_camera(i, scene.P[i])
            self.set_image(i, scene.images[i])
        return self

    def set_camera(self, index, P):
        P = np.ascontiguousarray(P, dtype=np.float32)
        self.lib.pmvso_set_camera(self.ctx, index, P.ctypes.data_as(C.c_void_p))

    def set_image(self, index, rgb):
        rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
        self.lib.pmvso_set_image(self.ctx, index, rgb.shape[1], rgb.shape[0], rgb.ctypes.data_as(C.c_void_p))

    def set_thresholds(self, ncc, ncc_before):
        self.lib.pmvso_set_thresholds(self.ctx, C.c_float(ncc), C.c_float(ncc_before))

    def set_xtol(self, xtol=1e-3, step=1.0, maxeval=1000):
        self.lib.pmvso_set_xtol(self.ctx, C.c_double(xtol), C.c_double(step), maxeval)

    def grab_tex(self, coord, normal, ref, index, wsize=None):
        """-> (flag, tex, newlevel); the texture has 3 * wsize^2 floats (the context's wsize: the C side writes that many)"""
        wsize = self.wsize if wsize is None else int(wsize)
        assert wsize == self.wsize, "grab_tex: the texture size is the context's wsize"
        tex = np.zeros(3 * wsize * wsize, np.float32); nl = C.c_int(-1)
        flag = self.lib.pmvso_grab_tex(self.ctx, _f4(coord).ctypes.data_as(C.c_void_p), _f4(normal).ctypes.data_as(C.c_void_p),
                                       int(ref), int(index), tex.ctypes.data_as(C.c_void_p), C.byref(nl))
        return int(flag), tex, nl.value

    def post_process(self, coord, normal, ncc, images, cap=256):
        """-> (verdict, images, grids, timages, tmp)   (_depth == 0 semantics)"""
        im = np.zeros(cap, np.int32); src = _imgs(images); im[: len(src)] = src
        n = C.c_int(len(src)); grids = np.zeros((cap, 2), np.int32); t = C.c_int(); tmp = C.c_float()
        r = self.lib.pmvso_post_process(self.ctx, _f4(coord).ctypes.data_as(C.c_void_p), _f4(normal).ctypes.data_as(C.c_void_p),
                                        C.c_float(ncc), im.ctypes.data_as(C.c_void_p), C.byref(n), cap,
                                        grids.ctypes.data_as(C.c_void_p), C.byref(t), C.byref(tmp))
        return int(r), im[: n.value].copy(), grids[: n.value].copy(), t.value, np.float32(tmp.value)

    # -- filter stage --------------------------------------------------------------------------------
    def set_depth(self, d):
        self.lib.pmvso_set_depth(self.ctx, int(d))

    def store_set(self, st):
        """st: dict from RefLib.state() / tests (coords, normals, ncc, dscale, img_off, images, grids, vimg_off, vimages, vgrids, timages)"""
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        self._store = {k: np.ascontiguousarray(v) for k, v in st.items()}
        t = self._store
        self.lib.pmvso_store_set(self.ctx, len(t["ncc"]), vp(t["coords"]), vp(t["normals"]), vp(t["ncc"]), vp(t["dscale"]), vp(t["img_off"]),
                                 vp(t["images"]), vp(t["grids"]), vp(t["vimg_off"]), vp(t["vimages"]), vp(t["vgrids"]), vp(t["timages"]))

    def grid_dims(self, image):
        gw, gh = C.c_int(), C.c_int()
        self.lib.pmvso_grid_dims(self.ctx, int(image), C.byref(gw), C.byref(gh))
        return gw.value, gh.value

    def build_depth_maps(self):
        self.lib.pmvso_build_depth_maps(self.ctx)

    def depth_map(self, image):
        gw, gh = self.grid_dims(image)
        out = np.zeros(gw * gh, np.int32)
        self.lib.pmvso_get_depth_map(self.ctx, int(image), out.ctypes.data_as(C.c_void_p))
        return out

    def is_visible_k(self, k, image, ix, iy, strict):
        t = self._store
        return self.lib.pmvso_is_visible(self.ctx, t["coords"][k].ctypes.data_as(C.c_void_p), t["normals"][k].ctypes.data_as(C.c_void_p),
                                         int(image), int(ix), int(iy), C.c_float(strict))

    def set_vimages(self, k, cap=256):
        vim = np.zeros(cap, np.int32); vgr = np.zeros((cap, 2), np.int32)
        n = self.lib.pmvso_set_vimages(self.ctx, int(k), vim.ctypes.data_as(C.c_void_p), vgr.ctypes.data_as(C.c_void_p), cap)
        return vim[:n].copy(), vgr[:n].copy()

    def filter_exact_safe(self, k, image, ix, iy):
        return self.lib.pmvso_filter_exact_safe(self.ctx, int(k), int(image), int(ix), int(iy))

    def is_neighbor(self, a, b, thr):
        return self.lib.pmvso_is_neighbor(self.ctx, int(a), int(b), C.c_float(thr))

    def compute_gain(self, k):
        f = self.lib.pmvso_compute_gain; f.restype = C.c_float
        return np.float32(f(self.ctx, int(k)))

    def detect_features(self, index, gspeedup=16, cap=65536):
        """(x, y), response, type per feature: Harris first, then DoG, strongest first"""
        xy = np.zeros((cap, 2), np.float32); resp = np.zeros(cap, np.float32); types = np.zeros(cap, np.int32)
        n = self.lib.pmvso_detect_features(self.ctx, int(index), int(gspeedup), xy.ctypes.data_as(C.c_void_p), resp.ctypes.data_as(C.c_void_p),
                                           types.ctypes.data_as(C.c_void_p), cap)
        return xy[:n].copy(), resp[:n].copy(), types[:n].copy()

    def compute_radius(self, k):
        f = self.lib.pmvso_compute_radius; f.restype = C.c_float
        return np.float32(f(self.ctx, int(k)))

    def find_neighbors(self, k, scale, margin, skipvis, cap=4096):
        out = np.zeros(cap, np.int32)
        n = self.lib.pmvso_find_neighbors(self.ctx, int(k), C.c_float(scale), int(margin), int(skipvis), out.ctypes.data_as(C.c_void_p), cap)
        return out[:min(n, cap)].copy()

    def find_empty_blocks(self, k):
        r = C.c_float()
        m = self.lib.pmvso_find_empty_blocks(self.ctx, int(k), C.byref(r))
        return int(m), np.float32(r.value)

    def check(self, k, quad=2.5):
        gain = C.c_float()
        rej = self.lib.pmvso_check(self.ctx, int(k), C.c_float(quad), C.byref(gain))
        return int(rej), np.float32(gain.value)

    def filter_neighbor(self, k, quad=2.5):
        res = C.c_float(); cnt = C.c_int()
        rej = self.lib.pmvso_filter_neighbor(self.ctx, int(k), C.c_float(quad), C.byref(res), C.byref(cnt))
        return int(rej), np.float32(res.value), cnt.value

    def close(self):
        if self.ctx:
            self.lib.pmvso_destroy(self.ctx)
            self.ctx = None


class RefLib(_Common):
    """The reference's own code.  One scene at a time per process (the reference uses a singleton)."""
    pfx = "ref_"

    def __init__(self, prefix, option="option.txt", skip_features=True, level=1, num=0):
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO)
        self.lib = C.CDLL(REF_SO)
        if not prefix.endswith("/"):
            prefix += "/"
        if skip_features:
            # detectFeatures.cpp:65-73 skips an image whose models/%08d.affin<level> exists
            os.makedirs(prefix + "models", exist_ok=True)
            for i in range(num):
                open(prefix + "models/%08d.affin%d" % (i, level), "a").close()
        self.lib.ref_open(prefix.encode(), option.encode())
        cfg = (C.c_int * 9)()
        self.lib.ref_config(cfg)
        (self.num, self.tnum, self.level, self.csize, self.wsize, self.min_image_num, self.tau, self.cpu, self.depth) = list(cfg)

    def _h(self):
        return ()

    def grab_tex(self, coord, normal, ref, index, wsize=None):
        wsize = self.wsize if wsize is None else int(wsize)
        assert wsize == self.wsize, "grab_tex: the texture size is the scene's wsize"
        tex = np.zeros(3 * wsize * wsize, np.float32)
        flag = self.lib.ref_grab_tex(_f4(coord).ctypes.data_as(C.c_void_p), _f4(normal).ctypes.data_as(C.c_void_p), int(ref), int(index),
                                     tex.ctypes.data_as(C.c_void_p))
        return int(flag), tex, None

    def set_depth(self, d):
        self.lib.ref_set_depth(d)

    def set_thresholds(self, ncc, ncc_before):
        self.lib.ref_set_thresholds(C.c_float(ncc), C.c_float(ncc_before))

    def post_process(self, coord, normal, ncc, images, cap=256):
        im = np.zeros(cap, np.int32); src = _imgs(images); im[: len(src)] = src
        n = C.c_int(len(src)); grids = np.zeros((cap, 2), np.int32)
        vim = np.zeros(cap, np.int32); nv = C.c_int(); vgr = np.zeros((cap, 2), np.int32)
        t = C.c_int(); tmp = C.c_float()
        r = self.lib.ref_post_process(_f4(coord).ctypes.data_as(C.c_void_p), _f4(normal).ctypes.data_as(C.c_void_p), C.c_float(ncc),
                                      im.ctypes.data_as(C.c_void_p), C.byref(n), cap, grids.ctypes.data_as(C.c_void_p),
                                      vim.ctypes.data_as(C.c_void_p), C.byref(nv), vgr.ctypes.data_as(C.c_void_p),
                                      C.byref(t), C.byref(tmp))
        return int(r), im[: n.value].copy(), grids[: n.value].copy(), t.value, np.float32(tmp.value)

    def run(self):
        self.lib.ref_run()

    # -- masks / edges / bounding images ---------------------------------------------------------------
    def map_bytes(self, index, which, level=None):
        """CImage::getMask(level) / getEdge(level) as (h, w) uint8, or None when the image has no such map"""
        level = self.level if level is None else level
        w, h = C.c_int(), C.c_int()
        self.lib.ref_image_dims(int(index), level, C.byref(w), C.byref(h))
        n = self.lib.ref_map_bytes(int(index), int(which), level, None)
        if n == 0:
            return None
        out = np.zeros(n, np.uint8)
        self.lib.ref_map_bytes(int(index), int(which), level, out.ctypes.data_as(C.c_void_p))
        return out[: w.value * h.value].reshape(h.value, w.value)

    def mask_gate(self, coord):
        return self.lib.ref_mask_gate(_f4(coord).ctypes.data_as(C.c_void_p))

    def get_edge(self, coord, index):
        return self.lib.ref_get_edge(_f4(coord).ctypes.data_as(C.c_void_p), int(index))

    def remove_images_edge(self, coord, images):
        im = _imgs(images).copy()
        n = self.lib.ref_remove_images_edge(_f4(coord).ctypes.data_as(C.c_void_p), im.ctypes.data_as(C.c_void_p), len(im))
        return im[:n].copy()

    def patches(self, cap=256):
        n = self.lib.ref_num_patches()
        coords = np.zeros((n, 4), np.float32); normals = np.zeros((n, 4), np.float32); nda = np.zeros((n, 3), np.float32)
        images = []
        im = np.zeros(cap, np.int32); vim = np.zeros(cap, np.int32); k = C.c_int(); kv = C.c_int()
        for i in range(n):
            self.lib.ref_get_patch(i, coords[i].ctypes.data_as(C.c_void_p), normals[i].ctypes.data_as(C.c_void_p),
                                   nda[i].ctypes.data_as(C.c_void_p), im.ctypes.data_as(C.c_void_p), C.byref(k), cap,
                                   vim.ctypes.data_as(C.c_void_p), C.byref(kv))
            images.append(im[: k.value].copy())
        return dict(coords=coords, normals=normals, ncc=nda[:, 0], dscale=nda[:, 1], ascale=nda[:, 2], images=images)

    # -- filter stage (state of the reference after run()) -------------------------------------------------
    def state(self, cap=256):
        """All patches after collectPatches(0) as arrays (CSR image lists); index == CPatch::_id."""
        P = self.lib.ref_collect_patches()
        coords = np.zeros((P, 4), np.float32); normals = np.zeros((P, 4), np.float32)
        sc = np.zeros((P, 4), np.float32); misc = np.zeros((P, 3), np.int32)
        img_off = [0]; vimg_off = [0]; images = []; grids = []; vimages = []; vgrids = []
        im = np.zeros(cap, np.int32); gr = np.zeros((cap, 2), np.int32); vim = np.zeros(cap, np.int32); vgr = np.zeros((cap, 2), np.int32)
        n, nv = C.c_int(), C.c_int()
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        for k in range(P):
            self.lib.ref_get_patch_full(k, vp(coords[k]), vp(normals[k]), vp(sc[k]), vp(im), vp(gr), C.byref(n), vp(vim), vp(vgr),
                                        C.byref(nv), vp(misc[k]), cap)
            images.append(im[: n.value].copy()); grids.append(gr[: n.value].copy())
            vimages.append(vim[: nv.value].copy()); vgrids.append(vgr[: nv.value].copy())
            img_off.append(img_off[-1] + n.value); vimg_off.append(vimg_off[-1] + nv.value)
        cat = lambda l, shape: (np.concatenate(l) if l and sum(len(x) for x in l) else np.zeros(shape, np.int32)).astype(np.int32)
        return dict(coords=coords, normals=normals, ncc=sc[:, 0].copy(), dscale=sc[:, 1].copy(), ascale=sc[:, 2].copy(), tmp=sc[:, 3].copy(),
                    img_off=np.array(img_off, np.int32), images=cat(images, (0,)), grids=cat(grids, (0, 2)),
                    vimg_off=np.array(vimg_off, np.int32), vimages=cat(vimages, (0,)), vgrids=cat(vgrids, (0, 2)),
                    timages=misc[:, 0].copy(), fix=misc[:, 1].copy())

    def grid_dims(self, image):
        gw, gh = C.c_int(), C.c_int()
        self.lib.ref_grid_dims(int(image), C.byref(gw), C.byref(gh))
        return gw.value, gh.value

    def build_depth_maps(self):
        self.lib.ref_set_depth_maps()

    def depth_map(self, image):
        gw, gh = self.grid_dims(image)
        out = np.zeros(gw * gh, np.int32)
        self.lib.ref_get_depth_map(int(image), out.ctypes.data_as(C.c_void_p))
        return out

    def is_visible_k(self, k, image, ix, iy, strict):
        return self.lib.ref_is_visible(int(k), int(image), int(ix), int(iy), C.c_float(strict))

    def set_vimages(self, k, cap=256):
        vim = np.zeros(cap, np.int32); vgr = np.zeros((cap, 2), np.int32)
        n = self.lib.ref_set_vimages(int(k), vim.ctypes.data_as(C.c_void_p), vgr.ctypes.data_as(C.c_void_p), cap)
        return vim[:n].copy(), vgr[:n].copy()

    def is_neighbor(self, a, b, thr):
        return self.lib.ref_is_neighbor(int(a), int(b), C.c_float(thr))

    def compute_gain(self, k):
        f = self.lib.ref_compute_gain; f.restype = C.c_float
        return np.float32(f(int(k)))

    def detect_features(self, index, gspeedup=16, cap=65536):
        xy = np.zeros((cap, 2), np.float32); resp = np.zeros(cap, np.float32); types = np.zeros(cap, np.int32)
        n = self.lib.ref_detect_features(int(index), int(gspeedup), xy.ctypes.data_as(C.c_void_p), resp.ctypes.data_as(C.c_void_p),
                                         types.ctypes.data_as(C.c_void_p), cap)
        return xy[:n].copy(), resp[:n].copy(), types[:n].copy()

    def compute_radius(self, k):
        f = self.lib.ref_compute_radius; f.restype = C.c_float
        return np.float32(f(int(k)))

    def find_neighbors(self, k, scale, margin, skipvis, cap=4096):
        out = np.zeros(cap, np.int32)
        n = self.lib.ref_find_neighbors(int(k), C.c_float(scale), int(margin), int(skipvis), out.ctypes.data_as(C.c_void_p), cap)
        return out[:min(n, cap)].copy()

    def find_empty_blocks(self, k):
        return int(self.lib.ref_find_empty_blocks(int(k)))

    def check(self, k, quad=2.5):
        gain = C.c_float()
        rej = self.lib.ref_check(int(k), C.c_float(quad), C.byref(gain))
        return int(rej), np.float32(gain.value)

    def filter_neighbor(self, k, quad=2.5):
        cnt = C.c_int()
        rej = self.lib.ref_filter_neighbor(int(k), C.c_float(quad), C.byref(cnt))
        return int(rej), cnt.value

    def remove_and_rebuild(self, keep, additive=1):
        """removePatch where keep == 0, then setDepthMapsVGridsVPGridsAddPatchV(additive) -> former table index of each new table patch"""
        keep = np.ascontiguousarray(keep, dtype=np.uint8)
        out = np.zeros(len(keep), np.int32)
        n = self.lib.ref_remove_and_rebuild(keep.ctypes.data_as(C.c_void_p), int(additive), out.ctypes.data_as(C.c_void_p), len(out))
        return out[:n].copy()

    def filter_small_groups(self, cap):
        out = np.zeros(cap, np.int32)
        n = self.lib.ref_filter_small_groups(out.ctypes.data_as(C.c_void_p), cap)
        return out[:n].copy()

    def shift_patches(self, flags, t):
        flags = np.ascontiguousarray(flags, dtype=np.uint8)
        self.lib.ref_shift_patches(flags.ctypes.data_as(C.c_void_p), C.c_float(t))

    def filter_exact(self, cap):
        out = np.zeros(cap, np.int32)
        n = self.lib.ref_filter_exact(out.ctypes.data_as(C.c_void_p), cap)
        return out[:n].copy()

    def set_xtol_floor(self, v):
        """x-tolerance floor of the nm3 stand-in behind the reference's nlopt call (0 = the reference's own 1e-7)"""
        self.lib.ref_set_xtol_floor(C.c_double(v))

    def depth_flag(self):
        return self.lib.ref_get_depth_flag()

    def counters(self):
        self.lib.ref_total_evals.restype = C.c_ulonglong
        self.lib.ref_total_calls.restype = C.c_ulonglong
        return self.lib.ref_total_evals(), self.lib.ref_total_calls()
