"""Data-parallel path on real GPUs; needs at least two devices, skipped on a one-GPU box.  Each case is a
torchrun launch (one process per GPU, rendezvous on 127.0.0.1):

  * tools/peer_check.py -- the peer-memory all-reduce kernel (csrc/peer_reduce.cu) against the fp64 sum of every rank's
    input, bit-identical on all ranks, bit-identical to ncclAllReduce with two ranks, 50 CUDA-graph replays;
  * tools/dp_check.py   -- DFGANTrainer with world = N: after eager, captured and replayed iterations on different
    shards every rank holds bit-identical parameters, and they differ from a world = 1 run on the same shard.
The host-side logic (shard bookkeeping, rank-averaged losses) is covered on CPU by test_dp_gloo.py /
test_trainer_shell_gloo.py.
"""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(script: str, nproc: int, port: int, env=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", script)]
    return subprocess.run(cmd, env=dict(os.environ, **(env or {})), capture_output=True, text=True, timeout=900)


def _nproc() -> int:
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    return 2 if n < 4 else 4


def test_peer_allreduce_kernel():
    r = _torchrun("peer_check.py", _nproc(), 29611)
    assert r.returncode == 0 and "PEER OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_trainer_ranks_stay_identical():
    r = _torchrun("dp_check.py", _nproc(), 29612)
    assert r.returncode == 0 and "DP OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_trainer_ranks_stay_identical_nccl_path():
    """DFGAN_DP_PEER=0: the round-1 exchange (ncclAllReduce between CUDA-graph segments) still works."""
    r = _torchrun("dp_check.py", _nproc(), 29613, env={"DFGAN_DP_PEER": "0"})
    assert r.returncode == 0 and "DP OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
