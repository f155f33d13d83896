"""Kernel-variant timing on one GPU: one scene + patch set, several builds of libpmvs_b200.so.
usage: python tools/variant_bench.py [--patches N] lib1.so lib2.so ...   (paths relative to cmvs-pmvs_b200/lib)"""
import argparse
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import __graft_entry__ as g
import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--patches", type=int, default=1 << 18)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("libs", nargs="+")
    a = ap.parse_args()
    pkg = g.load_package()
    args = argparse.Namespace(views=48, width=1600, height=1200)
    _, scene = bench.build_scene(args, "cuda:0")
    base = pkg.PmvsB200.from_scene(scene)
    coords, normals, images, dsc = bench.make_seed_patches(scene, base, a.patches, seed=4, device="cuda:0")
    base.close()
    import cmvs_pmvs_b200.binding as binding
    ref = None
    for name in a.libs:
        binding.LIB_PATH = os.path.join(ROOT, "cmvs-pmvs_b200", "lib", name)
        lib = binding.PmvsB200.from_scene(scene)
        dev = "cuda:0"
        d_c0 = torch.from_numpy(coords).to(dev); d_n0 = torch.from_numpy(normals).to(dev)
        d_im = torch.from_numpy(images).to(dev); d_ds = torch.from_numpy(dsc).to(dev)
        d_c = torch.empty_like(d_c0); d_n = torch.empty_like(d_n0)
        P = a.patches
        d_ncc = torch.empty(P, dtype=torch.float32, device=dev); d_ev = torch.empty(P, dtype=torch.int32, device=dev)
        d_ok = torch.empty(P, dtype=torch.uint8, device=dev)
        ms = []
        for r in range(a.reps + 1):
            d_c.copy_(d_c0); d_n.copy_(d_n0)
            torch.cuda.synchronize()
            lib.refine_batch_dev(P, 5, d_c.data_ptr(), d_n.data_ptr(), d_im.data_ptr(), 0, d_ds.data_ptr(), d_ncc.data_ptr(), d_ev.data_ptr(), d_ok.data_ptr())
            lib.sync()
            if r > 0:
                ms.append(lib.last_refine_ms())
        ncc = d_ncc.cpu().numpy()
        cur = (ncc, d_ev.cpu().numpy(), d_c.cpu().numpy(), d_n.cpu().numpy())
        if ref is None:
            ref = cur
        same = all(np.array_equal(a, b) for a, b in zip(cur, ref))
        print("%-22s kernel ms %s  -> %.3f M patches/s  evals %.1f  ok %.4f  max|dncc vs first| %.2e  identical evals %.4f  %s" % (
            name, ["%.1f" % m for m in ms], P / (min(ms) / 1e3) / 1e6, float(d_ev.float().mean()), float(d_ok.float().mean()),
            float(np.abs(ncc - ref[0]).max()), float((cur[1] == ref[1]).mean()),
            "BIT-IDENTICAL to first (ncc, evals, coords, normals)" if same else "differs from first"), flush=True)
        if hasattr(lib.lib, "pmvsb_refine_batch_dev_gather"):
            # the same batch through the kernel variant that also posts every record to the mailboxes (world 1: its own memory)
            lib.peer_export(0, 1, 2 * (P * 48 + 256))
            for form in ("0", "1"):      # 0: one scatter kernel after the refine kernel (default), 1: stores inside the refine kernel
                os.environ["PMVSB_GATHER_IN_KERNEL"] = form
                gms, wall = [], []
                for r in range(a.reps + 1):
                    d_c.copy_(d_c0); d_n.copy_(d_n0)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    st = torch.cuda.ExternalStream(lib.stream())
                    e0.record(st)
                    lib.refine_batch_dev_gather(P, 5, d_c.data_ptr(), d_n.data_ptr(), d_im.data_ptr(), 0, d_ds.data_ptr(), d_ncc.data_ptr(), d_ev.data_ptr(), d_ok.data_ptr())
                    e1.record(st)
                    lib.sync()
                    if r > 0:
                        gms.append(lib.last_refine_ms()); wall.append(e0.elapsed_time(e1))
                got = (d_ncc.cpu().numpy(), d_ev.cpu().numpy(), d_c.cpu().numpy(), d_n.cpu().numpy())
                print("%-22s   + record gather, %s (world 1): refine kernel ms %s, whole call ms %s, %s" % (
                      name, "stores inside the refine kernel" if form == "1" else "scatter kernel after it", ["%.2f" % m for m in gms], ["%.2f" % m for m in wall],
                      "BIT-IDENTICAL to the plain kernel" if all(np.array_equal(x, y) for x, y in zip(got, cur)) else "differs from the plain kernel"), flush=True)
            os.environ.pop("PMVSB_GATHER_IN_KERNEL", None)
        lib.close()


if __name__ == "__main__":
    main()
