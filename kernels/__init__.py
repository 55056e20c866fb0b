"""Drop-in import path: ``from kernels.attention.flash_attention import FlashAttention3`` etc. resolve to the B200
implementation in ``ml_inference_optimizer_b200.kernels``."""
import importlib
import sys

_IMPL = "ml_inference_optimizer_b200.kernels"
for _sub in ("attention", "attention.flash_attention", "attention.ring_attention", "mlp", "mlp.fused_mlp", "triton",
             "triton.flash_attention_kernels", "triton.attention_kernels", "triton.mlp_kernels", "triton.layernorm_kernels", "triton.fused_layernorm_qkv"):
    _mod = importlib.import_module(f"{_IMPL}.{_sub}")
    sys.modules[f"{__name__}.{_sub}"] = _mod
    if "." not in _sub:
        globals()[_sub] = _mod
