"""Multi-rank parity of the ring / tensor-parallel paths with the real kernels: over NCCL with one rank per GPU (needs >= 2
GPUs on the box; skipped otherwise) and, on ANY GPU box, with two and four ranks sharing cuda:0 over gloo. The host logic
of the same paths is covered on CPU by tests/test_distributed_gloo.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_ring_and_tp_parity_nccl(world, built_lib):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29520 + world), os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-4000:] + res.stderr[-4000:]


@pytest.mark.parametrize("world", [2, 4])
def test_ring_and_tp_parity_ranks_sharing_one_gpu(world, built_lib):
    """tests/multi_gpu_check.py with every rank on cuda:0 and gloo as the transport: zigzag / contiguous ring attention
    (K1 accumulate steps, LSE merge in the epilogue), TP MLP (column / row shards, one all-reduce, chunked overlap path)
    and TP attention against the fp32 oracle — the multi-rank logic with the real kernels on a one-GPU box."""
    env = dict(os.environ, B200_MGC_SHARED_GPU="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29560 + world), os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert res.returncode == 0, res.stdout[-4000:] + res.stderr[-4000:]
    assert "FAIL" not in res.stdout
