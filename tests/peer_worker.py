"""One rank of tests/test_gpu_peer.py: refine + fused all-gather of the records (pmvsb_refine_batch_dev_gather) with the other
ranks, which run as separate PROCESSES (on the same GPU or on others).  The 64-byte CUDA IPC handles travel through files.
usage: python tests/peer_worker.py RANK WORLD DIR [DEVICE]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import numpy as np
import torch

import __graft_entry__ as g


class DevView:      # zero-copy torch view of device memory owned by the library
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}


def wait_for(path, seconds=120.0):
    t0 = time.time()
    while not os.path.exists(path):
        if time.time() - t0 > seconds:
            raise SystemExit("peer_worker: %s did not appear" % path)
        time.sleep(0.01)


def main():
    rank, world, d = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    device = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    from scene_util import make_patches, small_scene
    from oracle.bindings import OracleLib      # only to draw the seed patches the other tests use (test infrastructure)
    pkg = g.load_package()
    scene = small_scene()
    torch.cuda.set_device(device)
    dev = "cuda:%d" % device
    lib = pkg.PmvsB200.from_scene(scene, device=device)
    orc = OracleLib.from_scene(scene)
    P = 700 + 100 * rank      # ragged: ranks bring different numbers of patches
    pb = make_patches(scene, orc, P, seed=11 + rank)
    slot = 2 * (1024 * 48 + 256)
    handle = lib.peer_export(rank, world, slot)
    with open(os.path.join(d, "handle%d.tmp" % rank), "wb") as f:
        f.write(handle)
    os.rename(os.path.join(d, "handle%d.tmp" % rank), os.path.join(d, "handle%d" % rank))
    handles = b""
    for k in range(world):
        wait_for(os.path.join(d, "handle%d" % k))
        handles += open(os.path.join(d, "handle%d" % k), "rb").read()
    if world > 1:
        lib.peer_open(handles)
    c0 = torch.from_numpy(pb["coords"]).to(dev); n0 = torch.from_numpy(pb["normals"]).to(dev)
    im = torch.from_numpy(np.ascontiguousarray(pb["images"])).to(dev); ds = torch.from_numpy(pb["dscales"].astype(np.float32)).to(dev)
    stride = pb["images"].shape[1]
    ncc = torch.empty(P, dtype=torch.float32, device=dev); ev = torch.empty(P, dtype=torch.int32, device=dev); ok = torch.empty(P, dtype=torch.uint8, device=dev)
    out = {}
    for step in range(3):      # three calls: both halves of the slots are used, and one is used again
        c = c0.clone(); n = n0.clone()
        if step == 1:      # the second call refines a perturbed start, so consecutive calls carry different records
            c[:, :3] += 1e-4
        torch.cuda.synchronize()
        ptr, rstride = lib.refine_batch_dev_gather(P, stride, c.data_ptr(), n.data_ptr(), im.data_ptr(), 0, ds.data_ptr(), ncc.data_ptr(), ev.data_ptr(), ok.data_ptr())
        lib.sync()
        got = torch.as_tensor(DevView(ptr, world * rstride), device=dev).view(world, rstride)[:, :1024 * 12].clone().cpu().numpy()
        out["gathered%d" % step] = got
        out["own%d" % step] = torch.cat([c, n, ncc[:, None], ok.float()[:, None], ev.float()[:, None], torch.zeros(P, 1, device=dev)], dim=1).cpu().numpy()
    # the plain entry point on the same start: the fused exchange must not change the results
    c = c0.clone(); n = n0.clone()
    torch.cuda.synchronize()      # the clones run on torch's stream, the library launches on its own
    lib.refine_batch_dev(P, stride, c.data_ptr(), n.data_ptr(), im.data_ptr(), 0, ds.data_ptr(), ncc.data_ptr(), ev.data_ptr(), ok.data_ptr())
    lib.sync()
    out["plain"] = torch.cat([c, n, ncc[:, None], ok.float()[:, None], ev.float()[:, None], torch.zeros(P, 1, device=dev)], dim=1).cpu().numpy()
    np.savez(os.path.join(d, "rank%d.npz" % rank), **out)
    # stay until every rank has finished reading our memory
    open(os.path.join(d, "done%d" % rank), "w").close()
    for k in range(world):
        wait_for(os.path.join(d, "done%d" % k))
    if world > 1:
        lib.peer_close()
    lib.close()


if __name__ == "__main__":
    main()
