"""CPU study: what the stopping tolerance of the shared Nelder-Mead definition (xtol = 1e-4 in scaled units,
1 unit = dscale = half a pixel of image motion) buys.  Runs the oracle's refinePatch on 2 048 patches of the test scene at several
xtol and compares each run with the xtol = 1e-4 one: evaluations per patch, |dncc|, depth (in dscale units) and normal differences.
The tolerance is part of the optimiser definition shared by oracle/shim/nlopt.hpp, oracle/pmvs_oracle.c and the kernel, so
changing it means re-pinning the golden vectors and the GPU parity tests together; this script only measures.
usage: python tools/research/xtol_study.py   (writes profiles/r1_xtol_study.json)"""
import json
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import oracle.bindings as ob
from scene_util import make_patches, small_scene
scene = small_scene()
orc = ob.OracleLib.from_scene(scene)
pt = make_patches(scene, orc, 2048, seed=11)
runs = []
base = None
for xtol in (1e-4, 3e-4, 1e-3, 3e-3, 1e-2, 3e-2):
    orc.set_xtol(xtol, 1.0, 1000)
    A = orc.refine_batch(pt["coords"], pt["normals"], pt["images"], pt["dscales"], threads=8)
    if base is None: base = A
    both = (A["ok"] == 1) & (base["ok"] == 1)
    dncc = np.abs(A["ncc"][both] - base["ncc"][both])
    depth = np.linalg.norm(A["coords"][both, :3] - base["coords"][both, :3], axis=1) / pt["dscales"][both]
    ang = np.degrees(np.arccos(np.clip((A["normals"][both, :3] * base["normals"][both, :3]).sum(1), -1, 1)))
    r = dict(xtol=xtol, evals=float(A["evals"].mean()), ok=float(A["ok"].mean()), median_dncc=float(np.median(dncc)), p99_dncc=float(np.quantile(dncc, .99)),
             median_depth_dscale=float(np.median(depth)), p99_depth_dscale=float(np.quantile(depth, .99)), median_angle_deg=float(np.median(ang)), p99_angle_deg=float(np.quantile(ang, .99)),
             mean_ncc=float(A["ncc"][both].mean()))
    print(json.dumps(r), flush=True)
    runs.append(r)
json.dump({'patches': 2048, 'runs': runs}, open(os.path.join(ROOT, 'profiles', 'r1_xtol_study.json'), 'w'), indent=1)
