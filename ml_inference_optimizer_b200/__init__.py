"""B200-native (sm_100a) attention + FusedMLP hot path behind the aslitaser/ml-inference-optimizer Python API.

Layout
  csrc/           hand-written CUDA (tcgen05 / TMEM / TMA) + the C-ABI (include/b200_attn_mlp.h)
  _lib.py, ops.py ctypes binding and torch-facing wrappers (no fallback)
  kernels/        host-side mirror of the reference's kernels/{attention,mlp,triton}
  parallelism/    ring attention (exact, overlapped) and column/row tensor parallelism over torch.distributed
  optimizer.py    the README's Optimizer
"""
__version__ = "0.1.0"

from . import ops  # noqa: F401
