"""Deterministic synthetic calibrated scenes (BASELINE.json `configs`).

Everything that decides a pixel value is built from IEEE-exact operations only (+ - * / sqrt floor on
float64 and integer hashing on int64), so a scene is bit-reproducible on CPU wherever it is generated;
tests/golden/ stores a SHA-256 of the image bytes next to the vectors derived from them.  Camera poses
use math.cos/sin on a handful of scalars.

A scene is written to disk in the layout the reference binary reads
(/root/reference/source/image/photoSetS.cpp:29-72, /root/reference/source/image/camera.cpp:13-54):
    <prefix>/visualize/%08d.ppm   binary P6
    <prefix>/txt/%08d.txt         "CONTOUR" + 3x4 projection matrix
    <prefix>/option.txt           pmvs option file (/root/reference/source/pmvs/option.cpp:47-109)
    <prefix>/models/              output directory
"""
from __future__ import annotations

import hashlib
import math
import os
from dataclasses import dataclass, field

import numpy as np
import torch

F64 = torch.float64
I64 = torch.int64


# --------------------------------------------------------------------------------------------------
# integer hash + value noise (exact)
# --------------------------------------------------------------------------------------------------
def _hash_u32(ix, iy, iz, seed: int):
    h = (ix * 0x8DA6B343 + iy * 0xD8163841 + iz * 0xCB1AB31F + seed * 0x9E3779B1) & 0xFFFFFFFF
    h = h ^ (h >> 15)
    h = (h * 0x2C1B3C6D) & 0xFFFFFFFF
    h = h ^ (h >> 12)
    h = (h * 0x297A2D39) & 0xFFFFFFFF
    h = h ^ (h >> 15)
    return h


def _value_noise3(x, y, z, freq: float, seed: int):
    """Trilinear (smoothstep) value noise in [0,1) at lattice frequency `freq`."""
    px, py, pz = x * freq, y * freq, z * freq
    fx, fy, fz = torch.floor(px), torch.floor(py), torch.floor(pz)
    tx, ty, tz = px - fx, py - fy, pz - fz
    tx = tx * tx * (3.0 - 2.0 * tx)
    ty = ty * ty * (3.0 - 2.0 * ty)
    tz = tz * tz * (3.0 - 2.0 * tz)
    ix, iy, iz = fx.to(I64), fy.to(I64), fz.to(I64)
    out = None
    for dz in (0, 1):
        wz = tz if dz else (1.0 - tz)
        for dy in (0, 1):
            wy = ty if dy else (1.0 - ty)
            for dx in (0, 1):
                wx = tx if dx else (1.0 - tx)
                v = _hash_u32(ix + dx, iy + dy, iz + dz, seed).to(F64) * (1.0 / 4294967296.0)
                term = v * (wx * wy * wz)
                out = term if out is None else out + term
    return out


def _texture_rgb(x, y, z, base_freq: float, seed: int):
    """Band-limited procedural colour in [0,255] (float64), 5 octaves from base_freq."""
    amps = (1.0, 0.8, 0.65, 0.5, 0.4)
    grey = None
    chans = [None, None, None]
    for o, a in enumerate(amps):
        f = base_freq * (2.0 ** o)
        g = (_value_noise3(x, y, z, f, seed * 131 + o) - 0.5) * a
        grey = g if grey is None else grey + g
        for k in range(3):
            c = (_value_noise3(x, y, z, f, seed * 131 + 17 + 7 * o + k) - 0.5) * (a * 0.5)
            chans[k] = c if chans[k] is None else chans[k] + c
    return [128.0 + 150.0 * (grey + chans[k]) for k in range(3)]


def _pixel_noise(view: int, ys, xs, k: int, seed: int):
    """~N(0,1): Irwin-Hall sum of 4 hashed uniforms (variance 4/12 -> scaled by sqrt(3))."""
    s = None
    for r in range(4):
        u = _hash_u32(xs, ys, torch.full_like(xs, view * 16 + k * 4 + r), seed + 977).to(F64) * (1.0 / 4294967296.0)
        s = u if s is None else s + u
    return (s - 2.0) * 1.7320508075688772


# --------------------------------------------------------------------------------------------------
# scene description
# --------------------------------------------------------------------------------------------------
@dataclass
class Scene:
    name: str
    width: int
    height: int
    K: np.ndarray                 # (V,3,3) float64
    R: np.ndarray                 # (V,3,3) float64 world->camera rows
    C: np.ndarray                 # (V,3) float64 optical centres
    P: np.ndarray                 # (V,3,4) float32, what the txt files carry
    images: list = field(default_factory=list)   # V x uint8 (H,W,3) numpy
    option: dict = field(default_factory=dict)
    kind: str = "sphere"
    params: dict = field(default_factory=dict)
    seed: int = 0
    base_freq: float = 4.0

    @property
    def num(self) -> int:
        return len(self.C)

    def sha256(self) -> str:
        h = hashlib.sha256()
        for im in self.images:
            h.update(np.ascontiguousarray(im).tobytes())
        h.update(np.ascontiguousarray(self.P).tobytes())
        return h.hexdigest()


def _look_at(C, T, up):
    C, T, up = (np.asarray(v, dtype=np.float64) for v in (C, T, up))
    fwd = T - C
    fwd = fwd / math.sqrt(float(fwd @ fwd))
    right = np.cross(fwd, up)
    right = right / math.sqrt(float(right @ right))
    down = np.cross(fwd, right)
    return np.stack([right, down, fwd])


def _make_cameras(centres, targets, ups, width, height, focal):
    V = len(centres)
    K = np.zeros((V, 3, 3))
    R = np.zeros((V, 3, 3))
    P = np.zeros((V, 3, 4), dtype=np.float32)
    for i in range(V):
        K[i] = [[focal, 0, (width - 1) / 2.0], [0, focal, (height - 1) / 2.0], [0, 0, 1]]
        R[i] = _look_at(centres[i], targets[i], ups[i])
        t = -R[i] @ np.asarray(centres[i], dtype=np.float64)
        P64 = K[i] @ np.concatenate([R[i], t[:, None]], axis=1)
        P[i] = P64.astype(np.float32)
    return K, R, np.asarray(centres, dtype=np.float64), P


# --------------------------------------------------------------------------------------------------
# surfaces: ray casting + ground truth
# --------------------------------------------------------------------------------------------------
def _relief_bumps(seed: int, n: int = 24, extent: float = 0.35):
    rng = np.random.default_rng(seed + 1000)
    cx = rng.uniform(-extent, extent, n)
    cy = rng.uniform(-extent, extent, n)
    rad = rng.uniform(0.08, 0.22, n)
    amp = rng.uniform(-0.18, 0.28, n) * rad
    return np.stack([cx, cy, rad, amp], axis=1)


def _relief_h(x, y, bumps):
    h = torch.zeros_like(x)
    for cx, cy, rad, amp in bumps:
        q = 1.0 - ((x - cx) * (x - cx) + (y - cy) * (y - cy)) / (rad * rad)
        q = torch.clamp(q, min=0.0)
        h = h + amp * q * q
    return h


def _relief_grad(x, y, bumps):
    gx = torch.zeros_like(x)
    gy = torch.zeros_like(x)
    for cx, cy, rad, amp in bumps:
        q = 1.0 - ((x - cx) * (x - cx) + (y - cy) * (y - cy)) / (rad * rad)
        q = torch.clamp(q, min=0.0)
        gx = gx + amp * 2.0 * q * (-2.0 * (x - cx) / (rad * rad))
        gy = gy + amp * 2.0 * q * (-2.0 * (y - cy) / (rad * rad))
    return gx, gy


def _cast(scene: Scene, ox, oy, oz, dx, dy, dz):
    """Ray origin o (scalars) + directions d (tensors) -> hit point tensors and hit mask."""
    kind = scene.kind
    if kind in ("sphere", "column"):
        ax, ay, az = scene.params["semi_axes"]
        # scale to unit sphere
        sx, sy, sz = ox / ax, oy / ay, oz / az
        ex, ey, ez = dx / ax, dy / ay, dz / az
        a = ex * ex + ey * ey + ez * ez
        b = sx * ex + sy * ey + sz * ez
        c = sx * sx + sy * sy + sz * sz - 1.0
        disc = b * b - a * c
        hit = disc > 0.0
        t = (-b - torch.sqrt(torch.clamp(disc, min=0.0))) / a
        # background: plane through the origin facing the camera keeps every pixel textured
        on = math.sqrt(ox * ox + oy * oy + oz * oz)
        nx, ny, nz = ox / on, oy / on, oz / on
        tb = -(ox * nx + oy * ny + oz * nz) / (dx * nx + dy * ny + dz * nz)
        t = torch.where(hit, t, tb)
        return ox + t * dx, oy + t * dy, oz + t * dz, hit
    if kind == "relief":
        bumps = scene.params["bumps"]
        t = -oz / dz
        for _ in range(24):
            x, y = ox + t * dx, oy + t * dy
            t = (_relief_h(x, y, bumps) - oz) / dz
        x, y = ox + t * dx, oy + t * dy
        return x, y, oz + t * dz, torch.ones_like(x, dtype=torch.bool)
    raise ValueError(kind)


def render(scene: Scene, device: str = "cpu", rows_per_chunk: int = 256) -> None:
    """Fill scene.images with uint8 (H,W,3) arrays."""
    W, H = scene.width, scene.height
    scene.images = []
    dev = torch.device(device)
    for v in range(scene.num):
        K, R, C = scene.K[v], scene.R[v], scene.C[v]
        f, cx, cy = float(K[0, 0]), float(K[0, 2]), float(K[1, 2])
        out = np.empty((H, W, 3), dtype=np.uint8)
        for y0 in range(0, H, rows_per_chunk):
            y1 = min(H, y0 + rows_per_chunk)
            ys_i, xs_i = torch.meshgrid(torch.arange(y0, y1, device=dev, dtype=I64),
                                        torch.arange(0, W, device=dev, dtype=I64), indexing="ij")
            u = (xs_i.to(F64) - cx) / f
            w = (ys_i.to(F64) - cy) / f
            # d_world = R^T (u, w, 1), written out (no matmul: keeps the arithmetic order fixed)
            dx = float(R[0, 0]) * u + float(R[1, 0]) * w + float(R[2, 0])
            dy = float(R[0, 1]) * u + float(R[1, 1]) * w + float(R[2, 1])
            dz = float(R[0, 2]) * u + float(R[1, 2]) * w + float(R[2, 2])
            X, Y, Z, hit = _cast(scene, float(C[0]), float(C[1]), float(C[2]), dx, dy, dz)
            rgb = _texture_rgb(X, Y, Z, scene.base_freq, scene.seed)
            for k in range(3):
                val = rgb[k]
                if scene.kind in ("sphere", "column"):
                    val = torch.where(hit, val, 0.35 * val + 30.0)   # dimmer, still textured background
                val = val + _pixel_noise(v, ys_i, xs_i, k, scene.seed)
                val = torch.floor(torch.clamp(val, 0.0, 255.0) + 0.5)
                out[y0:y1, :, k] = val.to(torch.uint8).cpu().numpy()
        scene.images.append(out)


# --------------------------------------------------------------------------------------------------
# the named configurations
# --------------------------------------------------------------------------------------------------
def _option(level, csize, threshold=0.7, wsize=7, min_image_num=3, cpu=1, num=16):
    return {"level": level, "csize": csize, "threshold": threshold, "wsize": wsize,
            "minImageNum": min_image_num, "CPU": cpu, "useVisData": 0, "sequence": -1,
            "timages": (-1, 0, num), "oimages": (0,)}


def sphere_scene(views: int = 16, width: int = 640, height: int = 480, seed: int = 0,
                 level: int = 1, csize: int = 2) -> Scene:
    """Config 1: unit sphere, `views` cameras on a ring of radius 3 with +-0.25 rad elevation wobble."""
    cs = []
    for i in range(views):
        th = 2.0 * math.pi * i / views
        el = 0.25 * math.sin(3.0 * th + 0.5)
        cs.append([3.0 * math.cos(el) * math.cos(th), 3.0 * math.cos(el) * math.sin(th), 3.0 * math.sin(el)])
    K, R, C, P = _make_cameras(cs, [[0, 0, 0]] * views, [[0, 0, 1]] * views, width, height, 1.2 * width)
    # finest octave ~3 px at the working level
    px_per_unit = (1.2 * width / (2 ** level)) / 2.0
    base = px_per_unit / 3.0 / 16.0
    return Scene("sphere%d" % views, width, height, K, R, C, P, kind="sphere",
                 params={"semi_axes": (1.0, 1.0, 1.0)}, seed=seed, base_freq=base,
                 option=_option(level, csize, num=views))


def ring_scene(views: int = 47, width: int = 640, height: int = 480, seed: int = 1,
               level: int = 0, csize: int = 1) -> Scene:
    """Config 2: templeRing-shaped: a tall column (ellipsoid 0.55 x 0.45 x 1.1), one camera ring."""
    cs = []
    for i in range(views):
        th = 2.0 * math.pi * i / views
        cs.append([3.4 * math.cos(th), 3.4 * math.sin(th), 0.6])
    K, R, C, P = _make_cameras(cs, [[0, 0, 0]] * views, [[0, 0, 1]] * views, width, height, 1.2 * width)
    px_per_unit = (1.2 * width / (2 ** level)) / 2.9
    base = px_per_unit / 3.0 / 16.0
    return Scene("ring%d" % views, width, height, K, R, C, P, kind="column",
                 params={"semi_axes": (0.55, 0.45, 1.1)}, seed=seed, base_freq=base,
                 option=_option(level, csize, num=views))


def dtu_scene(views: int = 48, width: int = 1600, height: int = 1200, seed: int = 2,
              level: int = 1, csize: int = 2) -> Scene:
    """Config 3: table-top relief, cameras on a 7x7-ish spherical cap at 0.7 from the table centre."""
    cs = []
    n = 7
    k = 0
    for j in range(n):
        for i in range(n):
            if k >= views:
                break
            az = (i - (n - 1) / 2.0) * 0.11 + 0.013 * ((j * 3) % 5 - 2)
            el = (j - (n - 1) / 2.0) * 0.11 + 0.011 * ((i * 2) % 5 - 2)
            cs.append([0.7 * math.sin(az) * math.cos(el), 0.7 * math.sin(el), 0.7 * math.cos(az) * math.cos(el)])
            k += 1
    K, R, C, P = _make_cameras(cs, [[0, 0, 0]] * len(cs), [[0, 1, 0]] * len(cs), width, height, 1.2 * width)
    px_per_unit = (1.2 * width / (2 ** level)) / 0.7
    base = px_per_unit / 3.0 / 16.0
    return Scene("dtu%d" % len(cs), width, height, K, R, C, P, kind="relief",
                 params={"bumps": _relief_bumps(seed)}, seed=seed, base_freq=base,
                 option=_option(level, csize, num=len(cs)))


# --------------------------------------------------------------------------------------------------
# ground truth helpers (used to make seed patches for the micro-benchmark and tests)
# --------------------------------------------------------------------------------------------------
def surface_samples(scene: Scene, n: int, seed: int, device: str = "cpu"):
    """n surface points and outward unit normals, float64 tensors (n,3)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    if scene.kind in ("sphere", "column"):
        v = torch.randn(n, 3, generator=g, dtype=F64)
        v = v / v.norm(dim=1, keepdim=True)
        ax = torch.tensor(scene.params["semi_axes"], dtype=F64)
        pts = v * ax
        nrm = v / ax
        nrm = nrm / nrm.norm(dim=1, keepdim=True)
        return pts.to(device), nrm.to(device)
    if scene.kind == "relief":
        xy = (torch.rand(n, 2, generator=g, dtype=F64) - 0.5) * 0.5
        x, y = xy[:, 0].to(device), xy[:, 1].to(device)
        bumps = scene.params["bumps"]
        z = _relief_h(x, y, bumps)
        gx, gy = _relief_grad(x, y, bumps)
        nrm = torch.stack([-gx, -gy, torch.ones_like(gx)], dim=1)
        nrm = nrm / nrm.norm(dim=1, keepdim=True)
        return torch.stack([x, y, z], dim=1), nrm
    raise ValueError(scene.kind)


# --------------------------------------------------------------------------------------------------
# disk layout for the reference binary
# --------------------------------------------------------------------------------------------------
def option_text(opt: dict) -> str:
    lines = []
    for k in ("level", "csize", "threshold", "wsize", "minImageNum", "CPU", "useVisData", "sequence"):
        lines.append("%s %s" % (k, opt[k]))
    for k in ("setEdge", "useBound", "quad", "maxAngle"):   # optional keys (source/pmvs/option.cpp:59-62, 102-106)
        if k in opt:
            lines.append("%s %s" % (k, opt[k]))
    lines.append("timages " + " ".join(str(v) for v in opt["timages"]))
    lines.append("oimages " + " ".join(str(v) for v in opt["oimages"]))
    return "\n".join(lines) + "\n"


def write_scene(scene: Scene, prefix: str, option_name: str = "option.txt") -> str:
    """Write the scene where the reference binary expects it; returns prefix with trailing '/'."""
    if not prefix.endswith("/"):
        prefix += "/"
    for d in ("visualize", "txt", "models"):
        os.makedirs(prefix + d, exist_ok=True)
    for i, im in enumerate(scene.images):
        with open(prefix + "visualize/%08d.ppm" % i, "wb") as f:
            f.write(b"P6\n%d %d\n255\n" % (scene.width, scene.height))
            f.write(np.ascontiguousarray(im).tobytes())
    for i in range(scene.num):
        with open(prefix + "txt/%08d.txt" % i, "w") as f:
            f.write("CONTOUR\n")
            for r in range(3):
                f.write(" ".join("%.9g" % float(v) for v in scene.P[i, r]) + "\n")
    with open(prefix + option_name, "w") as f:
        f.write(option_text(scene.option))
    return prefix
