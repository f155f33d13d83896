"""The peer-memory exchange (include/pmvs_b200.h: pmvsb_peer_*, pmvsb_refine_batch_dev_gather): the all-gather of the refined
records fused into the refine kernel.  Ranks are separate processes; on a one-GPU box they share the device (a CUDA IPC
handle of another process opens the same way, stores into the mapping land in the other process's allocation), on a multi-GPU
box the second variant puts them on different GPUs, where the stores cross NVLink.

Two forms (PMVSB_GATHER_IN_KERNEL): the records are stored into the mailboxes by one kernel that follows the refine kernel (the
default: measured faster) or by the refine kernel itself as each patch finishes.

Checked: every rank ends up with every rank's records, bit for bit what that rank's own output arrays hold, for three
consecutive calls (the two halves of a slot alternate); the fused kernel's results equal the plain pmvsb_refine_batch_dev's;
the wave exchange of pmvs2 itself over the same mailboxes is covered by tests/test_gpu_pipeline.py (peer, peer-grow, peer-mixed)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _run(world, devices, tmp_path, in_kernel=False):
    procs = []
    env = dict(os.environ, PMVSB_GATHER_IN_KERNEL="1" if in_kernel else "0")
    for r in range(world):
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "peer_worker.py"), str(r), str(world), str(tmp_path), str(devices[r])],
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env))
    outs = []
    for p in procs:
        try:
            o, e = p.communicate(timeout=600)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append((p.returncode, e))
    for rc, e in outs:
        assert rc == 0, e[-3000:]
    return [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]


def _check(res, world):
    for step in range(3):
        for me in range(world):
            for src in range(world):
                own = res[src]["own%d" % step]
                got = res[me]["gathered%d" % step][src][: own.size].reshape(own.shape)
                assert np.array_equal(got.view(np.int32), own.view(np.int32)), (step, me, src)
    for r in range(world):
        assert np.array_equal(res[r]["own0"].view(np.int32), res[r]["own2"].view(np.int32))      # same start, same result
        assert np.array_equal(res[r]["own0"].view(np.int32), res[r]["plain"].view(np.int32))       # the fused exchange changes no result
        assert not np.array_equal(res[r]["own0"], res[r]["own1"])                                    # the calls really carried different records
        assert (res[r]["own0"][:, 9] == 1).mean() > 0.9


def test_gather_single_rank(tmp_path):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    _check(_run(1, [0], tmp_path), 1)


@pytest.mark.parametrize("form", ["scatter-kernel", "in-refine-kernel"])
def test_gather_two_processes_one_gpu(form, tmp_path):
    """both forms of the exchange: the records stored by one kernel after the refine kernel (default), or by the refine kernel itself"""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    _check(_run(2, [0, 0], tmp_path, in_kernel=form == "in-refine-kernel"), 2)


def test_gather_two_gpus(tmp_path):
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    _check(_run(2, [0, 1], tmp_path), 2)
