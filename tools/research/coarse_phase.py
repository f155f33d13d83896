"""CPU study for the next kernel step: a COARSE first phase of refinePatch's optimiser on
hardware-filtered texture fetches, finished by the exact path.

The hot kernel sits at its texture-pipe floor with three one-channel gathers per bilinear sample; ONE hardware-bilinear
fetch costs a third of that (tools/probe/tex_lane_probe.cu) but interpolates with 8-bit weights (1.8 fixed point).  This
script answers, on the CPU and with the oracle's own code, what such a phase would do to the results:

  A  the oracle's refinePatch as is (Nelder-Mead on the exact objective, xtol 1e-4)                      = today's kernel
  B  Nelder-Mead on the objective sampled with weights rounded to 1/256 until the simplex is <= S, then a fresh
     Nelder-Mead on the exact objective from that point (start step T) to xtol 1e-4, exact computeINCC at the end

and reports, per (S, T): evaluations spent in each phase, and how B's patches compare with A's under the tolerances of
tests/test_gpu_parity.py::test_refine_matches_oracle (|dncc| <= 2e-3, depth <= 0.05 dscale, normal <= 1 degree).

The oracle is test infrastructure: this tool builds a PRIVATE, transformed copy of oracle/pmvs_oracle.c under /tmp (get_color
gains the rounded-weight variant and a two-phase driver is appended); nothing here is imported by the product.
usage: python tools/research/coarse_phase.py [--patches 2048] [--threads 8]"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

EXTRA = r'''
/* ---- study only: hardware-like bilinear (weights rounded to 1/256) behind a thread-local switch ---- */
static void get_color(const pmvso_ctx* c, int index, float x, float y, int level, float* rgb) {
  if (!g_coarse) { get_color_exact(c, index, x, y, level, rgb); return; }
  const int k = index * c->nlevels + level;
  const int W = c->w[k];
  const unsigned char* im = c->pix[k];
  int lx = (int)x, ly = (int)y;
  float dx1 = rintf((x - lx) * 256.0f) / 256.0f, dy1 = rintf((y - ly) * 256.0f) / 256.0f;
  const float dx0 = 1.0f - dx1, dy0 = 1.0f - dy1;
  const unsigned char* p0 = im + 3 * (ly * W + lx);
  const unsigned char* p1 = p0 + 3 * W;
  for (int ch = 0; ch < 3; ++ch)
    rgb[ch] = (p0[ch] * dx0 + p0[ch + 3] * dx1) * dy0 + (p1[ch] * dx0 + p1[ch + 3] * dx1) * dy1;
}

int pmvso_refine_two_phase(const pmvso_ctx* c, float* coord, float* normal, const int* images, int n, float dscale,
                           double xtol_coarse, double step2, float* ncc, int* evals_coarse, int* evals_exact) {
  float* texs = alloc_texs(c, c->tau);
  rctx_t r; rctx_init(&r, c, coord, normal, images, n, dscale);
  r.texs = texs;
  double p[3];
  encode(&r, coord, normal, p);
  const double lb[3] = {-HUGE_VAL, -23.99999, -23.99999};
  const double ub[3] = {HUGE_VAL, 23.99999, 23.99999};
  double x[3];
  for (int i = 0; i < 3; ++i) x[i] = fmax(fmin(p[i], ub[i]), lb[i]);
  double minf;
  int nev1 = 0, nev2 = 0;
  g_coarse = 1;
  int res = nm3_minimize(3, my_f, &r, lb, ub, x, &minf, c->step, xtol_coarse, c->maxeval, &nev1);
  g_coarse = 0;
  if (res == NM3_XTOL_REACHED) res = nm3_minimize(3, my_f, &r, lb, ub, x, &minf, step2, c->xtol, c->maxeval - nev1, &nev2);
  *evals_coarse = nev1; *evals_exact = nev2;
  int ok = 0;
  if (res == NM3_XTOL_REACHED) {
    decode(&r, x, coord, normal);
    *ncc = (float)(1.0 - unrobustincc((float)compute_incc(&r, coord, normal, 1)));
    ok = 1;
  }
  free(texs);
  return ok;
}

/* variant C: ONE Nelder-Mead run; when the simplex is <= switch_size the objective becomes the exact one, the four vertices are
   re-evaluated (4 evaluations) and re-sorted, and the same simplex carries on to xtol */
int pmvso_refine_switch(const pmvso_ctx* c, float* coord, float* normal, const int* images, int n, float dscale,
                        double switch_size, float* ncc, int* evals_coarse, int* evals_total) {
  float* texs = alloc_texs(c, c->tau);
  rctx_t r; rctx_init(&r, c, coord, normal, images, n, dscale);
  r.texs = texs;
  double p[3];
  encode(&r, coord, normal, p);
  const double lb[3] = {-HUGE_VAL, -23.99999, -23.99999};
  const double ub[3] = {HUGE_VAL, 23.99999, 23.99999};
  double x[3];
  for (int i = 0; i < 3; ++i) x[i] = fmax(fmin(p[i], ub[i]), lb[i]);
  double minf;
  int nev = 0, at = -1;
  g_coarse = 1;
  const int res = nm3_minimize_switch(3, my_f, &r, lb, ub, x, &minf, c->step, c->xtol, c->maxeval, &nev, switch_size, &at);
  g_coarse = 0;
  *evals_coarse = at < 0 ? nev : at; *evals_total = nev;
  int ok = 0;
  if (res == NM3_XTOL_REACHED && at >= 0) {
    decode(&r, x, coord, normal);
    *ncc = (float)(1.0 - unrobustincc((float)compute_incc(&r, coord, normal, 1)));
    ok = 1;
  }
  free(texs);
  return ok;
}
'''


def build_study_lib():
    src = open(os.path.join(ROOT, "oracle", "pmvs_oracle.c")).read()
    marker = "static void get_color(const pmvso_ctx* c, int index, float x, float y, int level, float* rgb) {"
    assert src.count(marker) == 1
    src = src.replace(marker, "static __thread int g_coarse = 0;\n"
                              "static void get_color(const pmvso_ctx* c, int index, float x, float y, int level, float* rgb);\n"
                              "static void get_color_exact(const pmvso_ctx* c, int index, float x, float y, int level, float* rgb) {")
    out_dir = "/tmp/pmvs_b200_study"
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "pmvs_oracle_study.c"), "w") as f:
        nm = open(os.path.join(ROOT, "oracle", "nm3.h")).read()
        body = nm[nm.index("static inline int nm3_minimize("):nm.rindex("#endif")]
        body = body.replace("nm3_minimize(", "nm3_minimize_switch(").replace("int* nevals)", "int* nevals, double switch_size, int* switched_at)")
        hook = "    if (size <= xtol) {"
        assert body.count(hook) == 1
        body = body.replace(hook, "    if (g_coarse && size <= switch_size) {\n      g_coarse = 0; *switched_at = cnt;\n"
                                  "      for (i = 0; i <= n; ++i) NM3_EVAL(fv[i], p[i]);\n      for (k = 1; k <= n; ++k) NM3_INSERT(k);\n      continue;\n    }\n" + hook)
        f.write(src + body + EXTRA)
    so = os.path.join(out_dir, "libpmvs_oracle_study.so")
    # the oracle's own flags (oracle/Makefile: baseline x86-64, no FMA contraction)
    subprocess.check_call(["gcc", "-O2", "-std=gnu11", "-ffp-contract=off", "-fPIC", "-shared", "-I", os.path.join(ROOT, "oracle"),
                           "-o", so, os.path.join(out_dir, "pmvs_oracle_study.c"), "-lm", "-lpthread"])
    return so


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--patches", type=int, default=2048)
    ap.add_argument("--threads", type=int, default=8)
    a = ap.parse_args()
    import oracle.bindings as ob
    so = build_study_lib()
    ob.build_oracle = lambda force=False: so
    from scene_util import make_patches, small_scene
    scene = small_scene()
    orc = ob.OracleLib.from_scene(scene)
    pt = make_patches(scene, orc, a.patches, seed=11)
    P = len(pt["coords"])
    A = orc.refine_batch(pt["coords"], pt["normals"], pt["images"], pt["dscales"], threads=a.threads)
    fn = orc.lib.pmvso_refine_two_phase
    fn.restype = C.c_int
    vp = lambda arr: arr.ctypes.data_as(C.c_void_p)

    def run(S, T):
        co = pt["coords"].copy(); no = pt["normals"].copy()
        ncc = np.full(P, -1.0, np.float32); e1 = np.zeros(P, np.int32); e2 = np.zeros(P, np.int32); ok = np.zeros(P, np.uint8)
        im = np.ascontiguousarray(pt["images"], np.int32)

        def one(i):
            n_ = C.c_float(-1.0); a1 = C.c_int(0); a2 = C.c_int(0)
            ok[i] = fn(orc.ctx, vp(co[i]), vp(no[i]), vp(im[i]), im.shape[1], C.c_float(float(pt["dscales"][i])), C.c_double(S), C.c_double(T),
                       C.byref(n_), C.byref(a1), C.byref(a2))
            ncc[i] = n_.value; e1[i] = a1.value; e2[i] = a2.value
        with ThreadPoolExecutor(a.threads) as ex:
            list(ex.map(one, range(P)))
        both = (ok == 1) & (A["ok"] == 1)
        dncc = np.abs(ncc[both] - A["ncc"][both])
        depth = np.linalg.norm(co[both, :3] - A["coords"][both, :3], axis=1) / pt["dscales"][both]
        ang = np.degrees(np.arccos(np.clip((no[both, :3] * A["normals"][both, :3]).sum(1), -1, 1)))
        good = (dncc <= 2e-3) & (depth <= 0.05) & (ang <= 1.0)
        return {"S": S, "T": T, "evals_coarse": float(e1.mean()), "evals_exact": float(e2.mean()), "ok_agree": float((ok == A["ok"]).mean()),
                "good": float(good.mean()), "median_dncc": float(np.median(dncc)), "p99_dncc": float(np.quantile(dncc, 0.99)),
                "median_depth": float(np.median(depth)), "median_angle_deg": float(np.median(ang)),
                "ncc_B_minus_A_mean": float((ncc[both] - A["ncc"][both]).mean()), "B_better_or_equal_within_1e-4": float((ncc[both] >= A["ncc"][both] - 1e-4).mean())}

    res = {"patches": P, "A_evals": float(A["evals"].mean()), "A_ok": float(A["ok"].mean()), "runs": []}
    print("A (exact Nelder-Mead): %.1f evaluations per patch, ok %.4f" % (res["A_evals"], res["A_ok"]), flush=True)
    for S, T in [(1e-4, 1e-3), (1e-2, 2e-2), (3e-2, 6e-2), (1e-1, 2e-1), (3e-1, 5e-1)]:
        r = run(S, T)
        res["runs"].append(r)
        print(json.dumps(r), flush=True)
    fs = orc.lib.pmvso_refine_switch
    fs.restype = C.c_int

    def run_switch(S):
        co = pt["coords"].copy(); no = pt["normals"].copy()
        ncc = np.full(P, -1.0, np.float32); e1 = np.zeros(P, np.int32); e2 = np.zeros(P, np.int32); ok = np.zeros(P, np.uint8)
        im = np.ascontiguousarray(pt["images"], np.int32)

        def one(i):
            n_ = C.c_float(-1.0); a1 = C.c_int(0); a2 = C.c_int(0)
            ok[i] = fs(orc.ctx, vp(co[i]), vp(no[i]), vp(im[i]), im.shape[1], C.c_float(float(pt["dscales"][i])), C.c_double(S),
                       C.byref(n_), C.byref(a1), C.byref(a2))
            ncc[i] = n_.value; e1[i] = a1.value; e2[i] = a2.value
        with ThreadPoolExecutor(a.threads) as ex:
            list(ex.map(one, range(P)))
        both = (ok == 1) & (A["ok"] == 1)
        dncc = np.abs(ncc[both] - A["ncc"][both])
        depth = np.linalg.norm(co[both, :3] - A["coords"][both, :3], axis=1) / pt["dscales"][both]
        ang = np.degrees(np.arccos(np.clip((no[both, :3] * A["normals"][both, :3]).sum(1), -1, 1)))
        good = (dncc <= 2e-3) & (depth <= 0.05) & (ang <= 1.0)
        return {"variant": "switch in place", "S": S, "evals_coarse": float(e1.mean()), "evals_exact": float((e2 - e1).mean()),
                "ok_agree": float((ok == A["ok"]).mean()), "good": float(good.mean()), "median_dncc": float(np.median(dncc)),
                "p99_dncc": float(np.quantile(dncc, 0.99)), "median_depth": float(np.median(depth)), "median_angle_deg": float(np.median(ang)),
                "B_better_or_equal_within_1e-4": float((ncc[both] >= A["ncc"][both] - 1e-4).mean())}
    for S in [1e-3, 3e-3, 1e-2, 3e-2, 1e-1, 3e-1]:
        r = run_switch(S)
        res["runs"].append(r)
        print(json.dumps(r), flush=True)
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", "r1_coarse_phase_study.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
