"""Functional entry points under the reference's names (``kernels.triton.*``). They call hand-written sm_100a CUDA
through the C-ABI; the package name is historical."""
