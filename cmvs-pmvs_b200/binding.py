"""ctypes binding of the C ABI in include/pmvs_b200.h (cmvs-pmvs_b200/lib/libpmvs_b200.so).

This is plumbing for tests/, bench.py and __graft_entry__: it only marshals numpy arrays (or raw device
pointers) into the library.  There is no fallback: if the CUDA library is missing, or no CUDA device can be
opened, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PMVS_B200_LIB") or os.path.join(HERE, "lib", "libpmvs_b200.so")  # override: kernel experiments

# every symbol include/pmvs_b200.h declares
SYMBOLS = [
    "pmvsb_create", "pmvsb_destroy", "pmvsb_last_error", "pmvsb_version", "pmvsb_device_count", "pmvsb_upload_camera", "pmvsb_upload_image",
    "pmvsb_upload_mask", "pmvsb_set_edge", "pmvsb_set_bimages", "pmvsb_download_mask", "pmvsb_mask_gate_batch",
    "pmvsb_remove_images_edge_batch", "pmvsb_store_set_seq", "pmvsb_store_counts", "pmvsb_store_rebuild", "pmvsb_filter_exact_apply_store",
    "pmvsb_set_features", "pmvsb_seed_candidates", "pmvsb_evaluate_batch", "pmvsb_evaluate_fetch", "pmvsb_evaluate_allgather", "pmvsb_evaluate_counts", "pmvsb_exchanged_bytes", "pmvsb_small_group_edges_store", "pmvsb_filter_small_groups_store", "pmvsb_store_download_lists", "pmvsb_set_visdata2", "pmvsb_finalize_scene", "pmvsb_set_thresholds", "pmvsb_set_optimizer", "pmvsb_image_dims",
    "pmvsb_download_image", "pmvsb_get_camera", "pmvsb_project_batch", "pmvsb_grab_tex_batch", "pmvsb_eval_objective_batch",
    "pmvsb_compute_incc_batch", "pmvsb_set_inccs_batch", "pmvsb_set_scales_batch", "pmvsb_pre_process_batch",
    "pmvsb_post_process_batch", "pmvsb_set_depth", "pmvsb_grid_dims", "pmvsb_store_upload", "pmvsb_build_depth_maps",
    "pmvsb_download_depth_map", "pmvsb_depth_maps_add", "pmvsb_store_append", "pmvsb_store_update_vimages",
    "pmvsb_store_download_vimages", "pmvsb_download_cell_lists", "pmvsb_find_empty_blocks_store", "pmvsb_filter_neighbor_store", "pmvsb_check_batch", "pmvsb_set_vimages_store", "pmvsb_filter_exact_store", "pmvsb_compute_gains_store", "pmvsb_set_vimages_batch", "pmvsb_set_ref_image_batch",
    "pmvsb_patch_colors_batch", "pmvsb_refine_batch",
    "pmvsb_refine_batch_dev", "pmvsb_refine_batch_dev_gather", "pmvsb_peer_export", "pmvsb_peer_open", "pmvsb_peer_close", "pmvsb_peer_needed", "pmvsb_detect_features", "pmvsb_comm_unique_id", "pmvsb_comm_init", "pmvsb_allgather", "pmvsb_sync", "pmvsb_stream", "pmvsb_set_stream", "pmvsb_launch_count", "pmvsb_last_refine_ms",
]


class PmvsError(RuntimeError):
    pass


def load_library() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise PmvsError("CUDA library not built: %s (run `python -c 'import __graft_entry__ as g; g.build()'`)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.pmvsb_last_error.restype = C.c_char_p
    lib.pmvsb_version.restype = C.c_char_p
    lib.pmvsb_stream.restype = C.c_void_p
    lib.pmvsb_launch_count.restype = C.c_uint64
    lib.pmvsb_last_refine_ms.restype = C.c_float
    if hasattr(lib, "pmvsb_peer_needed"):      # (older builds loaded by tools/variant_bench.py do not have it)
        lib.pmvsb_peer_needed.restype = C.c_size_t
    return lib


def _vp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a if shape is None else a.reshape(shape)


def _patch_arrays(coords, normals, images, nimages=None, dscales=None):
    coords = _f32(coords).reshape(-1, 4)
    P = coords.shape[0]
    normals = None if normals is None else _f32(normals).reshape(P, 4)
    images = np.ascontiguousarray(images, dtype=np.int32)
    if images.ndim != 2 or images.shape[0] != P:
        images = images.reshape(P, -1) if P else images.reshape(0, max(1, images.shape[-1] if images.ndim else 1))
    nimages = None if nimages is None else np.ascontiguousarray(nimages, dtype=np.int32).reshape(P)
    dscales = None if dscales is None else _f32(dscales).reshape(P)
    return P, images.shape[1], coords, normals, images, nimages, dscales


class PmvsB200:
    """One context = one GPU holding the image pyramids and camera tables of one scene."""

    def __init__(self, num_images, num_target=None, level=1, csize=2, wsize=7, min_image_num=3, threshold=0.7,
                 max_angle_deg=10.0, device=0):
        self.lib = load_library()
        self.ctx = C.c_void_p()
        r = self.lib.pmvsb_create(C.byref(self.ctx), device, num_images, num_images if num_target is None else num_target,
                                  level, csize, wsize, min_image_num, C.c_float(threshold), C.c_float(max_angle_deg))
        if r != 0:
            self.ctx = None
            raise PmvsError("pmvsb_create failed (%d): no CUDA device or bad arguments -- there is no CPU fallback" % r)
        self.num = num_images
        self.num_target = num_images if num_target is None else num_target
        self.level = level
        self.wsize = wsize
        self.tau = min(2 * min_image_num, num_images)

    # -- plumbing ---------------------------------------------------------------------------------
    def _ck(self, r):
        if r != 0:
            raise PmvsError("pmvs_b200 error %d: %s" % (r, self.lib.pmvsb_last_error(self.ctx).decode()))

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.pmvsb_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @classmethod
    def from_scene(cls, scene, device=0):
        o = scene.option
        self = cls(scene.num, level=o["level"], csize=o["csize"], wsize=o["wsize"], min_image_num=o["minImageNum"],
                   threshold=o["threshold"], device=device)
        for i in range(scene.num):
            self.upload_camera(i, scene.P[i])
            self.upload_image(i, scene.images[i])
        self.finalize_scene()
        return self

    # -- scene -------------------------------------------------------------------------------------
    def upload_camera(self, index, P):
        P = _f32(P).reshape(12)
        self._ck(self.lib.pmvsb_upload_camera(self.ctx, int(index), _vp(P)))

    def upload_image(self, index, rgb):
        rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
        self._ck(self.lib.pmvsb_upload_image(self.ctx, int(index), rgb.shape[1], rgb.shape[0], _vp(rgb)))

    def finalize_scene(self):
        self._ck(self.lib.pmvsb_finalize_scene(self.ctx))

    def upload_mask(self, index, gray, which=0):
        """which = 0: masks/%08d.pgm (inside when 127 < v); 1: edges/%08d.pgm (inside when 1 < v); before finalize_scene."""
        gray = np.ascontiguousarray(gray, dtype=np.uint8)
        self._ck(self.lib.pmvsb_upload_mask(self.ctx, int(index), int(which), gray.shape[1], gray.shape[0], _vp(gray)))

    def set_edge(self, threshold):
        self._ck(self.lib.pmvsb_set_edge(self.ctx, C.c_float(threshold)))

    def set_bimages(self, indexes):
        a = np.ascontiguousarray(indexes, dtype=np.int32).reshape(-1)
        self._ck(self.lib.pmvsb_set_bimages(self.ctx, _vp(a), len(a)))

    def set_visdata2(self, index, lst):
        a = np.ascontiguousarray(lst, dtype=np.int32).reshape(-1)
        self._ck(self.lib.pmvsb_set_visdata2(self.ctx, int(index), _vp(a), len(a)))

    def mask_map(self, index, which=0):
        """working-level mask (0) / edge map (1) of an image as (h, w) uint8, or None when it has none"""
        w, h = C.c_int(), C.c_int()
        self._ck(self.lib.pmvsb_image_dims(self.ctx, int(index), self.level, C.byref(w), C.byref(h)))
        out = np.zeros((h.value, w.value), np.uint8); present = C.c_int()
        self._ck(self.lib.pmvsb_download_mask(self.ctx, int(index), int(which), _vp(out), C.byref(present)))
        return out if present.value else None

    def mask_gate_batch(self, coords):
        coords = _f32(coords).reshape(-1, 4)
        out = np.zeros(coords.shape[0], np.uint8)
        self._ck(self.lib.pmvsb_mask_gate_batch(self.ctx, coords.shape[0], _vp(coords), _vp(out)))
        return out

    def remove_images_edge_batch(self, coords, images, nimages):
        P, stride, coords, _, images, nimages, _ = _patch_arrays(coords, None, images, nimages)
        im = images.copy(); n = nimages.copy()
        self._ck(self.lib.pmvsb_remove_images_edge_batch(self.ctx, P, stride, _vp(coords), _vp(im), _vp(n)))
        return im, n

    def set_thresholds(self, ncc, ncc_before):
        self._ck(self.lib.pmvsb_set_thresholds(self.ctx, C.c_float(ncc), C.c_float(ncc_before)))

    def set_optimizer(self, xtol=1e-3, step=1.0, maxeval=1000):
        self._ck(self.lib.pmvsb_set_optimizer(self.ctx, C.c_double(xtol), C.c_double(step), int(maxeval)))

    def image(self, index, level):
        w, h = C.c_int(), C.c_int()
        self._ck(self.lib.pmvsb_image_dims(self.ctx, int(index), level, C.byref(w), C.byref(h)))
        out = np.zeros((h.value, w.value, 3), np.uint8)
        self._ck(self.lib.pmvsb_download_image(self.ctx, int(index), level, _vp(out)))
        return out

    def camera(self, index, level=0):
        P = np.zeros(12, np.float32); ce = np.zeros(4, np.float32); oa = np.zeros(4, np.float32)
        xa = np.zeros(3, np.float32); ya = np.zeros(3, np.float32); za = np.zeros(3, np.float32); ips = C.c_float()
        self._ck(self.lib.pmvsb_get_camera(self.ctx, int(index), level, _vp(P), _vp(ce), _vp(oa), _vp(xa), _vp(ya), _vp(za), C.byref(ips)))
        return dict(P=P.reshape(3, 4), centre=ce, oaxis=oa, xaxis=xa, yaxis=ya, zaxis=za, ipscale=np.float32(ips.value))

    # -- batched calls (host arrays) -----------------------------------------------------------------
    def project_batch(self, coords, image, level):
        coords = _f32(coords).reshape(-1, 4)
        image = np.ascontiguousarray(image, dtype=np.int32).reshape(-1)
        out = np.zeros((coords.shape[0], 3), np.float32)
        self._ck(self.lib.pmvsb_project_batch(self.ctx, coords.shape[0], _vp(coords), _vp(image), level, _vp(out)))
        return out

    def grab_tex_batch(self, coords, normals, images, nimages=None):
        P, stride, coords, normals, images, nimages, _ = _patch_arrays(coords, normals, images, nimages)
        tsz = 3 * self.wsize * self.wsize
        tex = np.zeros((P, stride, tsz), np.float32); flag = np.zeros((P, stride), np.int32); nl = np.zeros((P, stride), np.int32)
        self._ck(self.lib.pmvsb_grab_tex_batch(self.ctx, P, stride, _vp(coords), _vp(normals), _vp(images), _vp(nimages),
                                               _vp(tex), _vp(flag), _vp(nl)))
        return tex, flag, nl

    def eval_objective_batch(self, coords, normals, images, dscales, x, nimages=None):
        P, stride, coords, normals, images, nimages, dscales = _patch_arrays(coords, normals, images, nimages, dscales)
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(P, 3)
        f = np.zeros(P, np.float64)
        self._ck(self.lib.pmvsb_eval_objective_batch(self.ctx, P, stride, _vp(coords), _vp(normals), _vp(images), _vp(nimages),
                                                     _vp(dscales), _vp(x), _vp(f)))
        return f

    def compute_incc_batch(self, coords, normals, images, robust=1, nimages=None):
        P, stride, coords, normals, images, nimages, _ = _patch_arrays(coords, normals, images, nimages)
        out = np.zeros(P, np.float64)
        self._ck(self.lib.pmvsb_compute_incc_batch(self.ctx, P, stride, _vp(coords), _vp(normals), _vp(images), _vp(nimages),
                                                   int(robust), _vp(out)))
        return out

    def set_inccs_batch(self, coords, normals, images, robust=0, nimages=None):
        P, stride, coords, normals, images, nimages, _ = _patch_arrays(coords, normals, images, nimages)
        out = np.zeros((P, stride), np.float32)
        self._ck(self.lib.pmvsb_set_inccs_batch(self.ctx, P, stride, _vp(coords), _vp(normals), _vp(images), _vp(nimages),
                                                int(robust), _vp(out)))
        return out

    def set_scales_batch(self, coords, images, nimages=None):
        P, stride, coords, _, images, nimages, _ = _patch_arrays(coords, None, images, nimages)
        d = np.zeros(P, np.float32); a = np.zeros(P, np.float32)
        self._ck(self.lib.pmvsb_set_scales_batch(self.ctx, P, stride, _vp(coords), _vp(images), _vp(nimages), _vp(d), _vp(a)))
        return d, a

    def pre_process_batch(self, coords, normals, images, nimages):
        """-> dict(images (P,stride), nimages, dscale, ascale, verdict); inputs are not modified."""
        P, stride, coords, normals, images, nimages, _ = _patch_arrays(coords, normals, images, nimages)
        im = images.copy(); n = nimages.copy()
        d = np.zeros(P, np.float32); a = np.zeros(P, np.float32); v = np.zeros(P, np.int32)
        self._ck(self.lib.pmvsb_pre_process_batch(self.ctx, P, stride, _vp(coords), _vp(normals), _vp(im), _vp(n), _vp(d), _vp(a), _vp(v)))
        return dict(images=im, nimages=n, dscale=d, ascale=a, verdict=v)

    def post_process_batch(self, coords, normals, ncc, images, nimages):
        """-> dict(images, nimages, grids (P,stride,2), timages, tmp, verdict)   (_depth == 0 semantics)"""
        P, stride, coords, normals, images, nimages, _ = _patch_arrays(coords, normals, images, nimages)
        ncc = _f32(ncc).reshape(P)
        im = images.copy(); n = nimages.copy()
        g = np.full((P, stride, 2), -1, np.int32); t = np.zeros(P, np.int32); tmp = np.zeros(P, np.float32); v = np.zeros(P, np.int32)
        self._ck(self.lib.pmvsb_post_process_batch(self.ctx, P, stride, _vp(coords), _vp(normals), _vp(ncc), _vp(im), _vp(n), _vp(g), _vp(t),
                                                   _vp(tmp), _vp(v)))
        return dict(images=im, nimages=n, grids=g, timages=t, tmp=tmp, verdict=v)

    # -- filter stage ------------------------------------------------------------------------------------
    def set_depth(self, depth):
        self._ck(self.lib.pmvsb_set_depth(self.ctx, int(depth)))

    def grid_dims(self, image):
        gw, gh = C.c_int(), C.c_int()
        self._ck(self.lib.pmvsb_grid_dims(self.ctx, int(image), C.byref(gw), C.byref(gh)))
        return gw.value, gh.value

    def store_upload(self, st):
        """st: dict with coords, normals, ncc, dscale, img_off, images, grids, vimg_off, vimages, vgrids, timages"""
        f = lambda k: np.ascontiguousarray(st[k], dtype=np.float32)
        i = lambda k: np.ascontiguousarray(st[k], dtype=np.int32)
        self._store_P = len(st["ncc"]); self._store_E = int(st["img_off"][-1])
        self._ck(self.lib.pmvsb_store_upload(self.ctx, self._store_P, _vp(f("coords")), _vp(f("normals")), _vp(f("ncc")), _vp(f("dscale")),
                                             _vp(i("img_off")), _vp(i("images")), _vp(i("grids")), _vp(i("vimg_off")), _vp(i("vimages")),
                                             _vp(i("vgrids")), _vp(i("timages"))))

    def store_append(self, st):
        """st: the same fields for the patches to append; offsets relative to their first entry"""
        f = lambda k: np.ascontiguousarray(st[k], dtype=np.float32)
        i = lambda k: np.ascontiguousarray(st[k], dtype=np.int32)
        n = len(st["ncc"])
        self._ck(self.lib.pmvsb_store_append(self.ctx, n, _vp(f("coords")), _vp(f("normals")), _vp(f("ncc")), _vp(f("dscale")),
                                             _vp(i("img_off")), _vp(i("images")), _vp(i("grids")), _vp(i("vimg_off")), _vp(i("vimages")),
                                             _vp(i("vgrids")), _vp(i("timages"))))
        self._store_P += n; self._store_E += int(st["img_off"][-1] - st["img_off"][0])

    def store_update_vimages(self, additive):
        """setVImagesVGrids for the whole table, in place; returns (vimg_off, vimages, vgrids[:, 2]) as CSR"""
        total = C.c_int32()
        self._ck(self.lib.pmvsb_store_update_vimages(self.ctx, int(additive), C.byref(total)))
        off = np.zeros(self._store_P + 1, np.int32); vim = np.zeros(max(total.value, 1), np.int32); vgr = np.zeros((max(total.value, 1), 2), np.int32)
        self._ck(self.lib.pmvsb_store_download_vimages(self.ctx, _vp(off), _vp(vim), _vp(vgr)))
        return off, vim[:total.value], vgr[:total.value]

    # -- the table reorganised on the device ------------------------------------------------------------
    def store_counts(self):
        P, E, VE = C.c_int32(), C.c_int32(), C.c_int32()
        self._ck(self.lib.pmvsb_store_counts(self.ctx, C.byref(P), C.byref(E), C.byref(VE)))
        return P.value, E.value, VE.value

    def store_set_seq(self, seq, first=0):
        seq = np.ascontiguousarray(seq, dtype=np.int32)
        self._ck(self.lib.pmvsb_store_set_seq(self.ctx, int(first), len(seq), _vp(seq)))

    def store_rebuild(self, keep=None, additive=1):
        """-> perm (old table index of every patch of the new table)"""
        P = self.store_counts()[0]
        k = None if keep is None else np.ascontiguousarray(keep, dtype=np.uint8)
        n = C.c_int32(); perm = np.zeros(max(P, 1), np.int32)
        self._ck(self.lib.pmvsb_store_rebuild(self.ctx, _vp(k), int(additive), C.byref(n), _vp(perm)))
        self._store_P = n.value; self._store_E = self.store_counts()[1]
        return perm[: n.value].copy()

    def store_download(self):
        """the table's lists as the device holds them: dict(seq, timages, img_off, images, grids, vimg_off, vimages, vgrids)"""
        P, E, VE = self.store_counts()
        seq = np.zeros(P, np.int32); ti = np.zeros(P, np.int32); off = np.zeros(P + 1, np.int32)
        im = np.zeros(max(E, 1), np.int32); gr = np.zeros((max(E, 1), 2), np.int32)
        self._ck(self.lib.pmvsb_store_download_lists(self.ctx, _vp(seq), _vp(ti), _vp(off), _vp(im), _vp(gr)))
        voff = np.zeros(P + 1, np.int32); vim = np.zeros(max(VE, 1), np.int32); vgr = np.zeros((max(VE, 1), 2), np.int32)
        self._ck(self.lib.pmvsb_store_download_vimages(self.ctx, _vp(voff), _vp(vim), _vp(vgr)))
        return dict(seq=seq, timages=ti, img_off=off, images=im[:E], grids=gr[:E], vimg_off=voff, vimages=vim[:VE], vgrids=vgr[:VE])

    def filter_exact_apply_store(self):
        keep = np.zeros(self.store_counts()[0], np.uint8)
        self._ck(self.lib.pmvsb_filter_exact_apply_store(self.ctx, _vp(keep)))
        return keep

    def small_group_edges_store(self, thr=1.0):
        P = self.store_counts()[0]
        off = np.zeros(P + 1, np.int32); total = C.c_int32()
        self._ck(self.lib.pmvsb_small_group_edges_store(self.ctx, C.c_float(thr), _vp(off), None, 0, C.byref(total)))
        adj = np.zeros(max(total.value, 1), np.int32)
        self._ck(self.lib.pmvsb_small_group_edges_store(self.ctx, C.c_float(thr), _vp(off), _vp(adj), total.value, C.byref(total)))
        return off, adj[: total.value]

    def filter_small_groups_store(self, thr=1.0):
        keep = np.zeros(self.store_counts()[0], np.uint8); t = C.c_int32()
        self._ck(self.lib.pmvsb_filter_small_groups_store(self.ctx, C.c_float(thr), _vp(keep), C.byref(t)))
        return keep, t.value

    def cell_lists(self, visible):
        cells = sum(self.grid_dims(i)[0] * self.grid_dims(i)[1] for i in range(self.num_target))
        off = np.zeros(cells + 1, np.int32)
        self._ck(self.lib.pmvsb_download_cell_lists(self.ctx, int(visible), _vp(off), None))
        lst = np.zeros(max(int(off[-1]), 1), np.int32)
        self._ck(self.lib.pmvsb_download_cell_lists(self.ctx, int(visible), _vp(off), _vp(lst)))
        return off, lst[:int(off[-1])]

    def find_empty_blocks_store(self, ids):
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        mask = np.zeros(len(ids), np.uint8); radius = np.zeros(len(ids), np.float32)
        self._ck(self.lib.pmvsb_find_empty_blocks_store(self.ctx, len(ids), _vp(ids), _vp(mask), _vp(radius)))
        return mask, radius

    def filter_neighbor_store(self, quad=2.5):
        P = self._store_P
        rej = np.zeros(P, np.uint8); res = np.zeros(P, np.float32); cnt = np.zeros(P, np.int32)
        ov = C.c_int32()
        self._ck(self.lib.pmvsb_filter_neighbor_store(self.ctx, C.c_float(quad), _vp(rej), _vp(res), _vp(cnt), C.byref(ov)))
        return rej, res, cnt, ov.value

    def check_batch(self, coords, normals, ncc, dscale, timages, images, nimages, grids, vimages, nv, vgrids, quad=2.5):
        """COptim::check for candidates outside the table; images (P, stride), grids (P, stride, 2), vimages (P, vstride), vgrids (P, vstride, 2)"""
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        coords = _f32(coords).reshape(-1, 4); P = coords.shape[0]
        images = i32(images).reshape(P, -1); vimages = i32(vimages).reshape(P, -1)
        gain = np.zeros(P, np.float32); rej = np.zeros(P, np.uint8); ov = C.c_int32()
        self._ck(self.lib.pmvsb_check_batch(self.ctx, P, images.shape[1], _vp(coords), _vp(_f32(normals)), _vp(_f32(ncc)), _vp(_f32(dscale)),
                                            _vp(i32(timages)), _vp(images), _vp(i32(nimages)), _vp(i32(grids)), vimages.shape[1], _vp(vimages),
                                            _vp(i32(nv)), _vp(i32(vgrids)), C.c_float(quad), _vp(gain), _vp(rej), C.byref(ov)))
        return gain, rej, ov.value

    def build_depth_maps(self):
        self._ck(self.lib.pmvsb_build_depth_maps(self.ctx))

    def depth_maps_add(self, coords):
        coords = _f32(coords).reshape(-1, 4)
        self._ck(self.lib.pmvsb_depth_maps_add(self.ctx, coords.shape[0], _vp(coords)))

    def depth_map(self, image):
        gw, gh = self.grid_dims(image)
        out = np.zeros(gw * gh, np.int32)
        self._ck(self.lib.pmvsb_download_depth_map(self.ctx, int(image), _vp(out)))
        return out

    def set_vimages_batch(self, coords, normals, images, nimages, vimages, nv, vgrids):
        """setVImagesVGrids for candidates outside the table: -> (vimages (P, vstride), nv, vgrids (P, vstride, 2)); inputs are not modified"""
        P, stride, coords, normals, images, nimages, _ = _patch_arrays(coords, normals, images, nimages)
        vim = np.ascontiguousarray(vimages, dtype=np.int32).reshape(P, -1).copy()
        n = np.ascontiguousarray(nv, dtype=np.int32).reshape(P).copy()
        vgr = np.ascontiguousarray(vgrids, dtype=np.int32).reshape(P, vim.shape[1], 2).copy()
        self._ck(self.lib.pmvsb_set_vimages_batch(self.ctx, P, stride, _vp(coords), _vp(normals), _vp(images), _vp(nimages), vim.shape[1], _vp(vim), _vp(n), _vp(vgr)))
        return vim, n, vgr

    def set_vimages_store(self, vcap):
        P = self._store_P
        vim = np.zeros((P, vcap), np.int32); vgr = np.zeros((P, vcap, 2), np.int32); nv = np.zeros(P, np.int32)
        self._ck(self.lib.pmvsb_set_vimages_store(self.ctx, int(vcap), _vp(vim), _vp(vgr), _vp(nv)))
        return vim, vgr, nv

    def filter_exact_store(self):
        safe = np.zeros(self._store_E, np.uint8)
        self._ck(self.lib.pmvsb_filter_exact_store(self.ctx, _vp(safe)))
        return safe

    def compute_gains_store(self):
        g = np.zeros(self._store_P, np.float32)
        self._ck(self.lib.pmvsb_compute_gains_store(self.ctx, _vp(g)))
        return g

    def refine_batch(self, coords, normals, images, dscales, nimages=None):
        """-> dict(coords, normals, ncc, evals, ok); inputs are not modified."""
        P, stride, coords, normals, images, nimages, dscales = _patch_arrays(coords, normals, images, nimages, dscales)
        co = coords.copy(); no = normals.copy()
        ncc = np.zeros(P, np.float32); ev = np.zeros(P, np.int32); ok = np.zeros(P, np.uint8)
        self._ck(self.lib.pmvsb_refine_batch(self.ctx, P, stride, _vp(co), _vp(no), _vp(images), _vp(nimages), _vp(dscales),
                                             _vp(ncc), _vp(ev), _vp(ok)))
        return dict(coords=co, normals=no, ncc=ncc, evals=ev, ok=ok)

    # -- device-pointer path (bench / multi-GPU plumbing) ------------------------------------------------
    def refine_batch_dev(self, P, stride, d_coords, d_normals, d_images, d_nimages, d_dscales, d_ncc, d_evals, d_ok):
        """All arguments are integer device addresses (e.g. torch.Tensor.data_ptr()); asynchronous."""
        p = lambda v: C.c_void_p(int(v)) if v else None
        self._ck(self.lib.pmvsb_refine_batch_dev(self.ctx, int(P), int(stride), p(d_coords), p(d_normals), p(d_images), p(d_nimages),
                                                 p(d_dscales), p(d_ncc), p(d_evals), p(d_ok)))

    def refine_batch_dev_gather(self, P, stride, d_coords, d_normals, d_images, d_nimages, d_dscales, d_ncc, d_evals, d_ok):
        """refine_batch_dev with the all-gather of the refined records fused into the kernel (peer stores into every rank's
        mailbox); returns (device address of rank 0's records of this call, stride between ranks in floats); asynchronous."""
        p = lambda v: C.c_void_p(int(v)) if v else None
        rec = C.c_void_p()
        rstride = C.c_size_t()
        self._ck(self.lib.pmvsb_refine_batch_dev_gather(self.ctx, int(P), int(stride), p(d_coords), p(d_normals), p(d_images), p(d_nimages),
                                                        p(d_dscales), p(d_ncc), p(d_evals), p(d_ok), C.byref(rec), C.byref(rstride)))
        return int(rec.value or 0), int(rstride.value)

    def peer_export(self, rank, world, slot_bytes) -> bytes:
        """(Re)allocates this rank's mailbox (one slot of slot_bytes per rank) and returns its 64-byte CUDA IPC handle."""
        buf = (C.c_uint8 * 64)()
        self._ck(self.lib.pmvsb_peer_export(self.ctx, int(rank), int(world), C.c_size_t(int(slot_bytes)), buf))
        self._world = int(world)
        return bytes(buf)

    def peer_open(self, handles: bytes):
        """handles = the world ranks' 64-byte handles back to back, in rank order"""
        buf = (C.c_uint8 * len(handles)).from_buffer_copy(handles)
        self._ck(self.lib.pmvsb_peer_open(self.ctx, buf))

    def peer_close(self):
        self._ck(self.lib.pmvsb_peer_close(self.ctx))

    def detect_features(self, index, gspeedup=16, cap=65536):
        xy = np.zeros((cap, 2), np.float32); resp = np.zeros(cap, np.float32); types = np.zeros(cap, np.int32)
        n = C.c_int32()
        self._ck(self.lib.pmvsb_detect_features(self.ctx, int(index), int(gspeedup), cap, _vp(xy), _vp(resp), _vp(types), C.byref(n)))
        k = min(n.value, cap)
        return xy[:k].copy(), resp[:k].copy(), types[:k].copy()

    def evaluate_batch(self, coords, normals, img_off, images, quad=2.5):
        """preProcess -> refinePatch -> postProcess for a wave of candidates in one call.
        -> dict(verdict (P,), refined, index, coords, normals, ncc, dscale, ascale, tmp, timages, img_off, images, grids, vimg_off, vimages, vgrids)"""
        coords = _f32(coords).reshape(-1, 4); normals = _f32(normals).reshape(-1, 4)
        P = coords.shape[0]
        img_off = np.ascontiguousarray(img_off, dtype=np.int32); images = np.ascontiguousarray(images, dtype=np.int32)
        A, E, VE, R = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        self._ck(self.lib.pmvsb_evaluate_batch(self.ctx, P, _vp(coords), _vp(normals), _vp(img_off), _vp(images), C.c_float(quad), C.byref(A), C.byref(E),
                                               C.byref(VE), C.byref(R)))
        a, e, ve = A.value, E.value, VE.value
        v = np.zeros(P, np.int32); idx = np.zeros(max(a, 1), np.int32); co = np.zeros((max(a, 1), 4), np.float32); no = np.zeros((max(a, 1), 4), np.float32)
        sc = np.zeros((max(a, 1), 4), np.float32); ti = np.zeros(max(a, 1), np.int32); off = np.zeros(a + 1, np.int32); im = np.zeros(max(e, 1), np.int32)
        gr = np.zeros((max(e, 1), 2), np.int32); voff = np.zeros(a + 1, np.int32); vim = np.zeros(max(ve, 1), np.int32); vgr = np.zeros((max(ve, 1), 2), np.int32)
        self._ck(self.lib.pmvsb_evaluate_fetch(self.ctx, _vp(v), _vp(idx), _vp(co), _vp(no), _vp(sc), _vp(ti), _vp(off), _vp(im), _vp(gr), _vp(voff), _vp(vim), _vp(vgr)))
        return dict(verdict=v, refined=R.value, index=idx[:a], coords=co[:a], normals=no[:a], ncc=sc[:a, 0], dscale=sc[:a, 1], ascale=sc[:a, 2], tmp=sc[:a, 3],
                    timages=ti[:a], img_off=off, images=im[:e], grids=gr[:e], vimg_off=voff, vimages=vim[:ve], vgrids=vgr[:ve])

    def set_features(self, index, xy, types):
        xy = _f32(xy).reshape(-1, 2); types = np.ascontiguousarray(types, dtype=np.int32).reshape(-1)
        self._ck(self.lib.pmvsb_set_features(self.ctx, int(index), len(types), _vp(xy), _vp(types)))

    def seed_candidates(self, index, views, blocked):
        """-> dict(ref_feature, ref_cell, ref_start, ref_count, coords (n,4), other_image, other_feature, resp)"""
        views = np.ascontiguousarray(views, dtype=np.int32).reshape(-1)
        blocked = np.ascontiguousarray(blocked, dtype=np.uint8).reshape(-1)
        nref, total = C.c_int32(), C.c_int32()
        call = lambda cr, rf, rc, rs, rn, cap, co, oi, of, rp: self._ck(self.lib.pmvsb_seed_candidates(
            self.ctx, int(index), len(views), _vp(views), _vp(blocked), cr, C.byref(nref), _vp(rf), _vp(rc), _vp(rs), _vp(rn), cap, C.byref(total),
            _vp(co), _vp(oi), _vp(of), _vp(rp)))
        call(0, None, None, None, None, 0, None, None, None, None)
        R = nref.value
        rf = np.zeros(max(R, 1), np.int32); rc = np.zeros(max(R, 1), np.int32); rs = np.zeros(max(R, 1), np.int32); rn = np.zeros(max(R, 1), np.int32)
        call(R, rf, rc, rs, rn, 0, None, None, None, None)
        n = total.value
        co = np.zeros((max(n, 1), 4), np.float32); oi = np.zeros(max(n, 1), np.int32); of = np.zeros(max(n, 1), np.int32); rp = np.zeros(max(n, 1), np.float32)
        if n:
            call(R, rf, rc, rs, rn, n, co, oi, of, rp)
        return dict(ref_feature=rf[:R], ref_cell=rc[:R], ref_start=rs[:R], ref_count=rn[:R], coords=co[:n], other_image=oi[:n], other_feature=of[:n], resp=rp[:n])

    def comm_unique_id(self):
        buf = (C.c_uint8 * 128)()
        self._ck(self.lib.pmvsb_comm_unique_id(self.ctx, buf))
        return bytes(buf)

    def comm_init(self, rank, world, unique_id: bytes):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self._ck(self.lib.pmvsb_comm_init(self.ctx, int(rank), int(world), buf))
        self._world = int(world)

    def allgather(self, send: np.ndarray):
        send = np.ascontiguousarray(send)
        recv = np.empty((getattr(self, "_world", 1),) + send.shape, send.dtype)
        self._ck(self.lib.pmvsb_allgather(self.ctx, _vp(send), C.c_size_t(send.nbytes), _vp(recv)))
        return recv

    def sync(self):
        self._ck(self.lib.pmvsb_sync(self.ctx))

    def set_stream(self, cuda_stream: int):
        self._ck(self.lib.pmvsb_set_stream(self.ctx, C.c_void_p(int(cuda_stream)) if cuda_stream else None))

    def stream(self) -> int:
        return int(self.lib.pmvsb_stream(self.ctx) or 0)

    def launch_count(self) -> int:
        return int(self.lib.pmvsb_launch_count(self.ctx))

    def last_refine_ms(self) -> float:
        return float(self.lib.pmvsb_last_refine_ms(self.ctx))
