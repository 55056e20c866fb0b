"""Mirror of the functional API of ``kernels/triton/fused_layernorm_qkv.py`` (reference :422-611, wrappers :1073, :1118):
LayerNorm followed by the Q/K/V projections, returning head-major views ready for the attention kernels.

Here the op is the LayerNorm kernel (``b200_layernorm``) followed by ONE tcgen05 GEMM over the concatenated
[Wq; Wk; Wv] weight (``b200_linear_act``); the outputs are strided views of that single result — K1 consumes them
through its (batch, seq, head) strides without copies. (The reference fuses both into one Triton kernel that has never
run; projection fusion into the attention kernel itself is out of scope, §2.2.)"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

from ... import ops
from .. import _measure as M

HAS_TRITON = True


def _cat_bias(biases, weights):
    if all(b is None for b in biases):
        return None
    return torch.cat([b if b is not None else torch.zeros(w.shape[0], dtype=w.dtype, device=w.device)
                      for b, w in zip(biases, weights)])


def triton_fused_layernorm_qkv(hidden_states: torch.Tensor, layernorm_weight: torch.Tensor,
                               layernorm_bias: Optional[torch.Tensor], query_weight: torch.Tensor, key_weight: torch.Tensor,
                               value_weight: torch.Tensor, query_bias: Optional[torch.Tensor] = None,
                               key_bias: Optional[torch.Tensor] = None, value_bias: Optional[torch.Tensor] = None,
                               eps: float = 1e-5, num_heads: int = 0, num_kv_heads: Optional[int] = None
                               ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Returns ``(q [B,S,H,Dq], k [B,S,Hkv,Dk], v [B,S,Hkv,Dk])`` with ``Dq = Wq.rows/H`` and ``Dk = Wk.rows/Hkv``."""
    B, S, hidden = hidden_states.shape
    if num_heads == 0:  # reference :470-480: try the common head sizes
        for hs in (64, 80, 128):
            if hidden % hs == 0:
                num_heads = hidden // hs
                break
        if num_heads == 0:
            num_heads = max(1, hidden // 64)
    if num_kv_heads is None:
        num_kv_heads = num_heads
    nq, nk, nv = query_weight.shape[0], key_weight.shape[0], value_weight.shape[0]
    if nq % num_heads or nk % num_kv_heads or nv % num_kv_heads:
        raise ValueError("projection widths must be divisible by the head counts")
    normed = ops.layernorm(hidden_states, layernorm_weight, layernorm_bias, eps)
    w = torch.cat([query_weight, key_weight, value_weight], dim=0)
    b = _cat_bias((query_bias, key_bias, value_bias), (query_weight, key_weight, value_weight))
    qkv = ops.linear_act(normed, w, b)
    q, k, v = qkv.split([nq, nk, nv], dim=-1)
    return (q.view(B, S, num_heads, nq // num_heads), k.view(B, S, num_kv_heads, nk // num_kv_heads),
            v.view(B, S, num_kv_heads, nv // num_kv_heads))


pytorch_fused_layernorm_qkv = triton_fused_layernorm_qkv


def flash_compatible_wrapper(hidden_states, layernorm_weight, layernorm_bias, qkv_weight, qkv_bias=None, eps: float = 1e-5,
                             num_heads: int = 0, num_kv_heads: Optional[int] = None):
    """reference :1073-1115 — combined ``[3*hidden, hidden]`` weight; outputs ``[B,S,H,D]`` for FlashAttention3."""
    hidden = hidden_states.shape[-1]
    wq, wk, wv = qkv_weight.split(hidden, dim=0) if qkv_weight.shape[0] == 3 * hidden else qkv_weight.chunk(3, dim=0)
    bq = bk = bv = None
    if qkv_bias is not None:
        bq, bk, bv = qkv_bias.split([wq.shape[0], wk.shape[0], wv.shape[0]])
    return triton_fused_layernorm_qkv(hidden_states, layernorm_weight, layernorm_bias, wq, wk, wv, bq, bk, bv, eps,
                                      num_heads, num_kv_heads)


def ring_compatible_wrapper(hidden_states, layernorm_weight, layernorm_bias, q_weight, k_weight, v_weight, q_bias=None,
                            k_bias=None, v_bias=None, eps: float = 1e-5, num_heads: int = 0,
                            num_kv_heads: Optional[int] = None):
    """reference :1118-1161 — outputs ``[B,H,S,D]`` (the ring / TP internal layout), as views."""
    q, k, v = triton_fused_layernorm_qkv(hidden_states, layernorm_weight, layernorm_bias, q_weight, k_weight, v_weight,
                                         q_bias, k_bias, v_bias, eps, num_heads, num_kv_heads)
    return q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)


# ------------------------------------------------------------------------------------------------------------------
# The reference's own measurement / validation helpers for this file (:707-1060), same arguments and result keys.
# ------------------------------------------------------------------------------------------------------------------
def _lnqkv_problem(batch_size, seq_len, hidden_size, num_heads, num_kv_heads, device, dtype=torch.bfloat16):
    if num_heads == 0:
        num_heads = max(1, hidden_size // 64)
    num_kv_heads = num_heads if num_kv_heads is None else num_kv_heads
    head_dim = hidden_size // num_heads
    g = torch.Generator(device=device).manual_seed(0)
    r = lambda *s, sc=1.0: (torch.randn(*s, device=device, generator=g) * sc).to(dtype)
    sc = hidden_size ** -0.5
    return (r(batch_size, seq_len, hidden_size), r(hidden_size), r(hidden_size, sc=0.1),
            r(num_heads * head_dim, hidden_size, sc=sc), r(num_kv_heads * head_dim, hidden_size, sc=sc),
            r(num_kv_heads * head_dim, hidden_size, sc=sc), r(num_heads * head_dim, sc=0.1), r(num_kv_heads * head_dim, sc=0.1),
            r(num_kv_heads * head_dim, sc=0.1), num_heads, num_kv_heads)


def _unfused_lnqkv(x, lw, lb, wq, wk, wv, bq, bk, bv, fp32: bool):
    """Comparator: torch LayerNorm followed by three separate library GEMMs."""
    c = (lambda t: t.float()) if fp32 else (lambda t: t)
    n = F.layer_norm(c(x), (x.shape[-1],), c(lw), c(lb), 1e-5)
    return F.linear(n, c(wq), c(bq)), F.linear(n, c(wk), c(bk)), F.linear(n, c(wv), c(bv))


def benchmark_fused_layernorm_qkv(batch_size: int, seq_len: int, hidden_size: int, num_heads: int = 0,
                                  num_kv_heads: Optional[int] = None, device: str = "cuda", iterations: int = 100,
                                  warmup: int = 10) -> Dict[str, float]:
    """reference :707-837 — K5 + one K3 GEMM next to torch LayerNorm + three cuBLAS GEMMs."""
    res = {"batch_size": batch_size, "seq_len": seq_len, "hidden_size": hidden_size, "triton_fused_ms": 0.0,
           "separate_pytorch_ms": 0.0, "speedup": 0.0}
    if not M.cuda_ready(device):
        return res
    x, lw, lb, wq, wk, wv, bq, bk, bv, H, Hkv = _lnqkv_problem(batch_size, seq_len, hidden_size, num_heads, num_kv_heads, device)
    res.update(num_heads=H, num_kv_heads=Hkv)
    res["triton_fused_ms"] = M.time_ms(lambda: triton_fused_layernorm_qkv(x, lw, lb, wq, wk, wv, bq, bk, bv, 1e-5, H, Hkv),
                                       warmup, iterations)
    res["separate_pytorch_ms"] = M.time_ms(lambda: _unfused_lnqkv(x, lw, lb, wq, wk, wv, bq, bk, bv, False), warmup, iterations)
    res["speedup"] = res["separate_pytorch_ms"] / max(res["triton_fused_ms"], 1e-9)
    return res


def compare_with_unfused_implementation(batch_size: int, seq_len: int, hidden_size: int, num_heads: int = 0,
                                        num_kv_heads: Optional[int] = None, device: str = "cuda") -> Dict[str, float]:
    """reference :840-948 — per-projection max difference to the unfused fp32 computation."""
    if not M.cuda_ready(device):
        return {"max_difference_q": 0.0, "max_difference_k": 0.0, "max_difference_v": 0.0, "is_correct": False}
    x, lw, lb, wq, wk, wv, bq, bk, bv, H, Hkv = _lnqkv_problem(batch_size, seq_len, hidden_size, num_heads, num_kv_heads, device)
    q, k, v = triton_fused_layernorm_qkv(x, lw, lb, wq, wk, wv, bq, bk, bv, 1e-5, H, Hkv)
    rq, rk, rv = _unfused_lnqkv(x, lw, lb, wq, wk, wv, bq, bk, bv, True)
    B, S = x.shape[:2]
    dq, dk, dv = (M.max_abs_diff(a.reshape(B, S, -1), b) for a, b in ((q, rq), (k, rk), (v, rv)))
    tol = 4 * M.MAX_ABS_TOL  # the normalised activations are rounded to bf16 between the two kernels (|LN(x)| up to ~10)
    return {"max_difference_q": dq, "max_difference_k": dk, "max_difference_v": dv, "is_correct": max(dq, dk, dv) <= tol,
            "batch_size": batch_size, "seq_len": seq_len, "hidden_size": hidden_size, "num_heads": H, "num_kv_heads": Hkv}


def profile_memory_usage(batch_size: int, seq_len: int, hidden_size: int, num_heads: int = 0, num_kv_heads: Optional[int] = None,
                         device: str = "cuda") -> Dict[str, float]:
    """reference :951-1060 — peak memory of both forms (here the concatenated weight is built per call, see the module text)."""
    if not M.cuda_ready(device):
        return {"unfused_memory_mb": 0.0, "fused_memory_mb": 0.0, "memory_saving_percent": 0.0}
    x, lw, lb, wq, wk, wv, bq, bk, bv, H, Hkv = _lnqkv_problem(batch_size, seq_len, hidden_size, num_heads, num_kv_heads, device)
    mem_u, _ = M.peak_mb(lambda: _unfused_lnqkv(x, lw, lb, wq, wk, wv, bq, bk, bv, False))
    mem_f, _ = M.peak_mb(lambda: triton_fused_layernorm_qkv(x, lw, lb, wq, wk, wv, bq, bk, bv, 1e-5, H, Hkv))
    return {"unfused_memory_mb": mem_u, "fused_memory_mb": mem_f, "memory_saving_mb": mem_u - mem_f,
            "memory_saving_percent": 100.0 * (mem_u - mem_f) / max(mem_u, 1e-6), "batch_size": batch_size, "seq_len": seq_len,
            "hidden_size": hidden_size}
