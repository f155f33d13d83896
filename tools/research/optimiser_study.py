"""CPU study: the shared Nelder-Mead definition (oracle/nm3.h) against SciPy's derivative-free minimisers on the SAME objective
(the oracle's my_f through ctypes) from the same starts: evaluations per patch and the final objective relative to nm3's optimum.
nlopt's BOBYQA (what the reference links) is not available here; this only says that the stand-in is not a wasteful choice.
usage: python tools/research/optimiser_study.py   (prints one JSON line per method)"""
import json
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from scipy.optimize import minimize
import oracle.bindings as ob
from scene_util import make_patches, small_scene
scene = small_scene(); orc = ob.OracleLib.from_scene(scene)
N = 192
pt = make_patches(scene, orc, N, seed=11)
A = orc.refine_batch(pt["coords"], pt["normals"], pt["images"], pt["dscales"], threads=8)
lb = np.array([-np.inf, -23.99999, -23.99999]); ub = -lb
def study(method, opts, label):
    ev = []; df = []; dd = []
    for i in range(N):
        if not A["ok"][i]: continue
        c, n, im, ds = pt["coords"][i], pt["normals"][i], pt["images"][i], float(pt["dscales"][i])
        x0 = orc.encode(c, n, im, ds)
        cnt = [0]
        def f(x):
            cnt[0] += 1
            return orc.my_f(c, n, im, ds, np.clip(x, lb, ub))
        r = minimize(f, x0, method=method, options=opts)
        xa = orc.encode(A["coords"][i], A["normals"][i], im, ds)   # angles of NM's optimum (they depend on the reference camera only)
        ray = c[:3].astype(np.float64) - scene.C[im[0]][:3]; ray /= np.linalg.norm(ray)
        xa[0] = float((A["coords"][i][:3].astype(np.float64) - c[:3]) @ ray) / ds   # its depth along THIS patch's ray, in dscale units
        fa = orc.my_f(c, n, im, ds, xa)
        ev.append(cnt[0]); df.append(r.fun - fa); dd.append(abs(r.x[0] - xa[0]))
    ev = np.array(ev); df = np.array(df); dd = np.array(dd)
    print(json.dumps(dict(method=label, evals=float(ev.mean()), f_minus_fNM_median=float(np.median(df)), f_minus_fNM_p90=float(np.quantile(df, .9)),
                          worse_than_NM_by_1e4=float((df > 1e-4).mean()), better_than_NM_by_1e4=float((df < -1e-4).mean()),
                          depth_diff_median=float(np.median(dd)), depth_diff_p90=float(np.quantile(dd, .9)))), flush=True)
print("NM oracle evals", float(A["evals"][A["ok"] == 1].mean()))
study("Nelder-Mead", dict(xatol=1e-4, fatol=1e-12, initial_simplex=None), "scipy NM xatol 1e-4")
study("Powell", dict(xtol=1e-2, ftol=1e-6), "Powell xtol 1e-2")
study("COBYLA", dict(rhobeg=1.0, tol=1e-4), "COBYLA rhobeg 1 tol 1e-4")
study("COBYLA", dict(rhobeg=1.0, tol=1e-3), "COBYLA rhobeg 1 tol 1e-3")
