"""CPU study: how far is the repository's optimiser (oracle/nm3.h Nelder-Mead, stopped at xtol 1e-4) from what the reference's own call
computes -- nlopt LN_BOBYQA at xtol_rel 1e-7 (source/pmvs/optim.cpp:621-644)?  nlopt is absent, so the comparison arms are
  nm3@1e-4   the shared definition (oracle, shim, kernel)
  nm3@1e-7   the same simplex run down to the reference's tolerance (floor 0)
  bobyqa     tools/research/bobyqa_dense.py, Powell's method restated for n = 3, nlopt's scaling, rho_end = 1e-7 rho_beg
all on the oracle's my_f (the reference's objective, bit-exact against the reference's own on the golden vectors) from the same
starts.  Reported per arm against bobyqa: evaluations per patch, and percentiles of |d ncc|, depth difference in dscale units
(1 unit = half a pixel of image motion) and normal angle.
usage: python tools/research/optimiser_bound_study.py [--patches 2048] [--scene small|dtu]   (writes profiles/r2_optimiser_bound_<scene>.json)"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools", "research"))
import oracle.bindings as ob
from bobyqa_dense import bobyqa
from scene_util import make_patches, small_scene


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--patches", type=int, default=2048)
    ap.add_argument("--scene", default="small")
    a = ap.parse_args()
    if a.scene == "small":
        scene = small_scene()
    else:
        import __graft_entry__ as g
        synth = g.load_package().synth
        scene = synth.dtu_scene()
        synth.render(scene, device="cpu")
    orc = ob.OracleLib.from_scene(scene)
    pt = make_patches(scene, orc, a.patches, seed=11, k=5 if a.scene == "dtu" else None)
    n = len(pt["coords"])
    arms = {}
    for name, xtol in (("nm3@1e-4", 1e-4), ("nm3@1e-3", 1e-3), ("nm3@1e-7", 1e-7)):
        orc.set_xtol(xtol, 1.0, 1000)
        r = orc.refine_batch(pt["coords"], pt["normals"], pt["images"], pt["dscales"], threads=8)
        arms[name] = dict(coords=r["coords"], normals=r["normals"], ncc=r["ncc"], evals=r["evals"].astype(np.float64), ok=r["ok"].astype(bool))
    orc.set_xtol(1e-3, 1.0, 1000)
    lb = np.array([-np.inf, -23.99999, -23.99999]); ub = -lb
    co = np.zeros((n, 4), np.float32); no = np.zeros((n, 4), np.float32); ncc = np.zeros(n, np.float32); ev = np.zeros(n); ok = np.zeros(n, bool)
    t0 = time.time()
    for i in range(n):
        c, nm, im, ds = pt["coords"][i], pt["normals"][i], pt["images"][i], pt["dscales"][i]
        x0 = np.clip(orc.encode(c, nm, im, ds), lb, ub)
        f = lambda x: orc.my_f(c, nm, im, ds, x)
        x, fx, ne, st = bobyqa(f, x0, lb, ub, xtol_rel=1e-7, maxeval=1000)
        oc, on = orc.decode(c, nm, im, ds, x)
        co[i], no[i] = oc, on
        inc = orc.compute_incc(oc, on, im, 1)
        ncc[i] = 1.0 - inc / (1.0 - 3.0 * inc)          # unrobustincc (include/pmvs/optim.hpp:90-92)
        ev[i] = ne; ok[i] = st == "xtol"
    arms["bobyqa@1e-7"] = dict(coords=co, normals=no, ncc=ncc, evals=ev, ok=ok)
    print("bobyqa arm: %.1f s" % (time.time() - t0), flush=True)
    base = arms["bobyqa@1e-7"]
    out = {"scene": scene.name, "patches": n, "views_per_patch": int(pt["images"].shape[1]), "reference_call": "nlopt LN_BOBYQA, xtol_rel 1e-7, maxeval 1000",
           "arms": {}}
    q = lambda v, p: float(np.quantile(v, p)) if len(v) else None
    for name, r in arms.items():
        both = r["ok"] & base["ok"]
        dncc = np.abs(r["ncc"][both] - base["ncc"][both])
        depth = np.linalg.norm(r["coords"][both, :3] - base["coords"][both, :3], axis=1) / pt["dscales"][both]
        ang = np.degrees(np.arccos(np.clip((r["normals"][both, :3] * base["normals"][both, :3]).sum(1), -1, 1)))
        better = (r["ncc"][both] > base["ncc"][both] + 1e-4).mean(); worse = (r["ncc"][both] < base["ncc"][both] - 1e-4).mean()
        out["arms"][name] = dict(evals_per_patch=float(r["evals"].mean()), ok=float(r["ok"].mean()), compared=int(both.sum()), mean_ncc=float(r["ncc"][both].mean()),
                                 dncc_p50=q(dncc, .5), dncc_p90=q(dncc, .9), dncc_p99=q(dncc, .99), depth_p50=q(depth, .5), depth_p90=q(depth, .9), depth_p99=q(depth, .99),
                                 angle_p50=q(ang, .5), angle_p90=q(ang, .9), angle_p99=q(ang, .99), frac_ncc_higher_by_1e4=float(better), frac_ncc_lower_by_1e4=float(worse),
                                 within_test_tolerance=float(((dncc <= 2e-3) & (depth <= 0.05) & (ang <= 1.0)).mean()))
        print(name, json.dumps(out["arms"][name]), flush=True)
    path = os.path.join(ROOT, "profiles", "r2_optimiser_bound_%s.json" % a.scene)
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
