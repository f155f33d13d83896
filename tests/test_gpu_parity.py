"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bars: integer / byte / index work bit-exact; my_f / computeINCC / setINCCs within 1e-4 absolute
(BASELINE.json north_star); refined depth / normal within the tolerances written in test_refine_*.
"""
import numpy as np
import pytest

from scene_util import make_patches

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def patches(scene, oracle):
    return make_patches(scene, oracle, 400, seed=11)


def test_pyramid_bit_exact(gpu, oracle, scene):
    for i in range(scene.num):
        for level in range(scene.option["level"] + 3):
            a, b = gpu.image(i, level), oracle.image(i, level)
            assert a.shape == b.shape
            assert np.array_equal(a, b), "pyramid differs: image %d level %d" % (i, level)


def test_camera_constants_bit_exact(gpu, oracle, scene):
    for i in range(scene.num):
        for level in (0, scene.option["level"]):
            a, b = gpu.camera(i, level), oracle.camera(i, level)
            for k in a:
                assert np.array_equal(a[k], b[k]), (i, level, k)


def test_project_bit_exact(gpu, oracle, scene, patches):
    rng = np.random.default_rng(5)
    n = len(patches["coords"])
    img = rng.integers(0, scene.num, n).astype(np.int32)
    for level in (0, 1, 2):
        out = gpu.project_batch(patches["coords"], img, level)
        ref = np.stack([oracle.project(img[i], patches["coords"][i], level) for i in range(n)])
        assert np.array_equal(out, ref)
    # a point behind the camera takes the (-65535, -65535, -1) branch
    behind = np.array([[scene.C[0][0] * 2, scene.C[0][1] * 2, scene.C[0][2] * 2, 1.0]], np.float32)
    out = gpu.project_batch(behind, np.zeros(1, np.int32), 1)
    assert np.array_equal(out[0], oracle.project(0, behind[0], 1))
    assert out[0, 2] == -1.0


def test_grab_tex_flags_levels_textures(gpu, oracle, patches):
    tex, flag, nl = gpu.grab_tex_batch(patches["coords"], patches["normals"], patches["images"])
    n, k = patches["images"].shape
    grabbed = 0
    for p in range(n):
        for v in range(k):
            f, t, l = oracle.grab_tex(patches["coords"][p], patches["normals"][p], patches["images"][p, 0], patches["images"][p, v])
            assert flag[p, v] == f, (p, v)
            assert nl[p, v] == l, (p, v)
            if f == 0:
                grabbed += 1
                assert np.array_equal(tex[p, v], t), "texture differs at patch %d view %d: %g" % (p, v, np.abs(tex[p, v] - t).max())
    assert grabbed > n  # the batch really sampled something


def test_my_f_within_1e4(gpu, oracle, patches):
    rng = np.random.default_rng(7)
    n = len(patches["coords"])
    x = rng.normal(size=(n, 3)) * np.array([1.5, 2.0, 2.0])
    x[: n // 8] = 0.0
    f = gpu.eval_objective_batch(patches["coords"], patches["normals"], patches["images"], patches["dscales"], x)
    ref = np.array([oracle.my_f(patches["coords"][i], patches["normals"][i], patches["images"][i], patches["dscales"][i], x[i]) for i in range(n)])
    assert np.array_equal(f == 2.0, ref == 2.0), "the 2.0 plateau (too few views) must match exactly"
    assert np.abs(f - ref).max() <= 1e-4
    assert (ref < 2.0).sum() > n // 2


def test_compute_incc_within_1e4(gpu, oracle, patches):
    n = len(patches["coords"])
    for robust in (1, 0):
        out = gpu.compute_incc_batch(patches["coords"], patches["normals"], patches["images"], robust=robust)
        ref = np.array([oracle.compute_incc(patches["coords"][i], patches["normals"][i], patches["images"][i], robust) for i in range(n)])
        assert np.array_equal(out == 2.0, ref == 2.0)
        assert np.abs(out - ref).max() <= 1e-4


def test_set_inccs_within_1e4(gpu, oracle, scene, patches):
    n = len(patches["coords"])
    # all images of the scene, ragged lengths
    rng = np.random.default_rng(9)
    stride = scene.num
    images = np.zeros((n, stride), np.int32)
    nimages = rng.integers(1, stride + 1, n).astype(np.int32)
    for i in range(n):
        rest = [j for j in range(scene.num) if j != patches["images"][i, 0]]
        images[i] = [patches["images"][i, 0]] + rest
    for robust in (0, 1):
        out = gpu.set_inccs_batch(patches["coords"], patches["normals"], images, robust=robust, nimages=nimages)
        for i in range(n):
            ref = oracle.set_inccs(patches["coords"][i], patches["normals"][i], images[i, : nimages[i]], robust)
            got = out[i, : nimages[i]]
            assert np.array_equal(got == 2.0, ref == 2.0), i
            assert np.abs(got - ref).max() <= 1e-4, i


def test_set_scales_bit_exact(gpu, oracle, patches):
    d, a = gpu.set_scales_batch(patches["coords"], patches["images"])
    n = len(patches["coords"])
    for i in range(n):
        rd, ra = oracle.set_scales(patches["coords"][i], patches["images"][i])
        assert d[i] == rd, i
        assert abs(float(a[i]) - float(ra)) <= 1e-6 * abs(float(ra)), i


def test_refine_matches_oracle(gpu, oracle, patches):
    """Same Nelder-Mead on both sides (oracle/nm3.h); sums over texels associate differently on the GPU,
    so objective values differ by ~1e-6 and an occasional simplex comparison flips.  Stated tolerance:
    for >= 97% of patches refined by both, |ncc| within 2e-3, depth within 0.05 dscale units (= 0.05 px
    of image motion) and normal within 1 degree; the optimiser verdict agrees for >= 99%."""
    g = gpu.refine_batch(patches["coords"], patches["normals"], patches["images"], patches["dscales"])
    o = oracle.refine_batch(patches["coords"], patches["normals"], patches["images"], patches["dscales"], threads=8)
    n = len(patches["coords"])
    assert (g["ok"] == o["ok"]).mean() >= 0.99
    both = (g["ok"] == 1) & (o["ok"] == 1)
    assert both.sum() > 0.9 * n
    # untouched when the optimiser failed (optim.cpp:649-655)
    failed = g["ok"] == 0
    assert np.array_equal(g["coords"][failed], patches["coords"][failed])
    dncc = np.abs(g["ncc"][both] - o["ncc"][both])
    depth = np.linalg.norm(g["coords"][both, :3] - o["coords"][both, :3], axis=1) / patches["dscales"][both]
    cosang = np.clip((g["normals"][both, :3] * o["normals"][both, :3]).sum(1), -1, 1)
    ang = np.degrees(np.arccos(cosang))
    good = (dncc <= 2e-3) & (depth <= 0.05) & (ang <= 1.0)
    print("refine parity: n=%d both=%d good=%.4f median dncc=%.2e depth=%.2e ang=%.2e identical=%.3f mean evals gpu=%.1f cpu=%.1f" % (
        n, both.sum(), good.mean(), np.median(dncc), np.median(depth), np.median(ang),
        (g["evals"][both] == o["evals"][both]).mean(), g["evals"].mean(), o["evals"].mean()))
    assert good.mean() >= 0.97
    # and the refinement did its job: photo-consistency is high on the true surface
    assert np.median(g["ncc"][both]) > 0.9


def test_refine_improves_objective(gpu, oracle, patches):
    """Size-independent property: the refined patch never scores worse than its start under my_f."""
    g = gpu.refine_batch(patches["coords"], patches["normals"], patches["images"], patches["dscales"])
    n = len(patches["coords"])
    x0 = np.zeros((n, 3))
    # encode the start normals through the oracle to evaluate f at the start
    for i in range(n):
        x0[i] = oracle.encode(patches["coords"][i], patches["normals"][i], patches["images"][i], patches["dscales"][i])
    f0 = gpu.eval_objective_batch(patches["coords"], patches["normals"], patches["images"], patches["dscales"], x0)
    x1 = np.zeros((n, 3))
    for i in range(n):
        x1[i] = oracle.encode(g["coords"][i], g["normals"][i], patches["images"][i], patches["dscales"][i])
    f1 = gpu.eval_objective_batch(g["coords"], g["normals"], patches["images"], patches["dscales"], x1)
    ok = g["ok"] == 1
    assert (f1[ok] <= f0[ok] + 1e-4).mean() >= 0.99


def test_empty_and_error_paths(gpu, pkg, patches):
    # empty batch is a no-op
    out = gpu.refine_batch(np.zeros((0, 4), np.float32), np.zeros((0, 4), np.float32), np.zeros((0, 6), np.int32), np.zeros(0, np.float32))
    assert len(out["ok"]) == 0
    # out-of-range image index is rejected on the host, with a message
    bad = patches["images"].copy()
    bad[0, 1] = 10 ** 6
    with pytest.raises(pkg.PmvsError):
        gpu.refine_batch(patches["coords"], patches["normals"], bad, patches["dscales"])
