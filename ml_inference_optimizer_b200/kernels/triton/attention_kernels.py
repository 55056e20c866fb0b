"""Mirror of the functional API of ``kernels/triton/attention_kernels.py``: paged decode attention, KV append and the
single-GPU "ring" (chunked online-softmax) attention."""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

from ... import ops
from .. import _measure as M

TRITON_AVAILABLE = True


def triton_paged_attention_forward(query: torch.Tensor, output: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor,
                                   block_tables: torch.Tensor, context_lengths: torch.Tensor, block_size: int,
                                   max_seq_len: int, layer_idx: int, causal: bool = True) -> None:
    """reference :1206-1311. query/output ``[B,H,q_len,D]``, cache ``[num_blocks, L, block_size, Hkv, D]``, int32
    tables/lengths. Writes ``output`` in place. q_len == 1: decode kernel K2. q_len > 1 (chunked prefill, the reference's
    BLOCK_SIZE_M = 64 case, :1251): prefill kernel K1 gathering K/V through the block table; the q_len queries are the last
    q_len tokens of each sequence and are masked causally among themselves (``causal=False`` = the reference kernel as
    written, whose causal mask is commented out, :774-777)."""
    assert query.dim() == 4 and output.shape == query.shape, "query/output must be [B,H,q_len,D]"
    assert block_tables.dtype == torch.int32 and context_lengths.dtype == torch.int32, "tables/lengths must be int32"
    assert k_cache.shape[2] == block_size, "block_size does not match the cache"
    B, H, q_len, D = query.shape
    if q_len != 1:
        ops.paged_prefill_attention(query.transpose(1, 2), k_cache, v_cache, block_tables.contiguous(),
                                    context_lengths.contiguous(), layer_idx=layer_idx, causal=causal,
                                    out=output.transpose(1, 2))
        return
    o = ops.decode_attention(query.reshape(B, H, D), k_cache, v_cache, context_lengths.contiguous(),
                             block_tables=block_tables.contiguous(), layer_idx=layer_idx, max_context_len=max_seq_len)
    output.copy_(o.view(B, H, 1, D))


def triton_reshape_and_cache(key: torch.Tensor, value: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor,
                             block_tables: torch.Tensor, context_lengths: torch.Tensor, layer_idx: int) -> None:
    """reference :1314-1407. key/value ``[B,1,Hkv,D]`` (decode only, warning at :1363-1365) are written at position
    ``context_lengths[b]-1``."""
    assert key.dim() == 4 and key.shape[1] == 1, "reshape_and_cache handles one new token per sequence"
    assert block_tables.dtype == torch.int32 and context_lengths.dtype == torch.int32, "tables/lengths must be int32"
    ops.kv_append(key[:, 0], value[:, 0], k_cache, v_cache, context_lengths.contiguous(), block_tables.contiguous(), layer_idx)


def triton_ring_attention_forward(query: torch.Tensor, key: torch.Tensor, value: torch.Tensor,
                                  attention_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """reference :909-998 — query/key/value ``[B,H,S,D]`` -> ``[B,S,H*D]``, non-causal exact attention computed
    chunk by chunk (one K1 launch covers the whole key range). Additive masks are not supported (use the module API
    with ``causal`` / key-padding lengths)."""
    if attention_mask is not None:
        raise NotImplementedError("additive attention masks are not supported; use causal / kv_lens")
    B, H, S, D = query.shape
    o = ops.flash_attn_fwd(query.transpose(1, 2), key.transpose(1, 2), value.transpose(1, 2))
    return o.reshape(B, S, H * D)


# ------------------------------------------------------------------------------------------------------------------
# The reference's own measurement helpers for this file (:1643-1800), same arguments and result keys.
# ------------------------------------------------------------------------------------------------------------------
def compare_with_flash_attention(seq_len: int, batch_size: int, hidden_size: int, num_heads: int) -> Dict[str, float]:
    """reference :1643-1732 — ``triton_ring_attention_forward`` next to the ``flash_attn`` package's kernel (fp16, non-causal).
    Returns ``{"error": ...}`` when that package cannot be imported or cannot run on this GPU, as the reference does."""
    try:
        from flash_attn import flash_attn_func
    except Exception:
        return {"error": "FlashAttention is not available for comparison"}
    head_dim = hidden_size // num_heads
    g = torch.Generator(device="cuda").manual_seed(0)
    q, k, v = (torch.randn(batch_size, seq_len, num_heads, head_dim, device="cuda", dtype=torch.float16, generator=g) for _ in range(3))
    qr, kr, vr = (t.permute(0, 2, 1, 3) for t in (q, k, v))
    try:
        flash = lambda: flash_attn_func(q, k, v, dropout_p=0.0, softmax_scale=1.0 / math.sqrt(head_dim))
        flash_out = flash()
        torch.cuda.synchronize()
    except Exception as e:  # (the package's binaries may not cover this architecture)
        return {"error": f"FlashAttention is not available for comparison: {type(e).__name__}"}
    ring_out = triton_ring_attention_forward(qr, kr, vr).reshape(batch_size, seq_len, num_heads, head_dim)
    t_flash = M.time_ms(flash, 5, 10)
    t_ring = M.time_ms(lambda: triton_ring_attention_forward(qr, kr, vr), 5, 10)
    err = (flash_out.float() - ring_out.float()).abs()
    return {"flash_attention_time_ms": t_flash, "ring_attention_time_ms": t_ring, "flash_vs_ring_speedup": t_ring / t_flash,
            "max_absolute_diff": err.max().item(), "mean_absolute_diff": err.mean().item(),
            "relative_error": err.mean().item() / flash_out.float().abs().mean().item()}


def calculate_attention_theoretical_flops(seq_len: int, batch_size: int, hidden_size: int, num_heads: int) -> Dict[str, float]:
    """reference :1735-1800 — the reference's operation-count model (one count per multiply-accumulate; softmax 5 per score
    in the materialised form, 7 in the chunked form with 128-key chunks; projections included)."""
    head_dim = hidden_size // num_heads
    proj = 4 * batch_size * seq_len * hidden_size * hidden_size            # Q, K, V and output projections
    scores = batch_size * num_heads * seq_len * seq_len                    # entries of the score matrix
    std_total = proj + 2 * scores * head_dim + 5 * scores
    chunk = min(128, seq_len)
    covered = batch_size * num_heads * seq_len * (-(-seq_len // chunk)) * chunk   # keys rounded up to whole chunks
    ring_total = proj + 2 * covered * head_dim + 7 * covered
    flash_total = proj + 2 * scores * head_dim
    return {"standard_attention_gflops": std_total / 1e9, "ring_attention_gflops": ring_total / 1e9,
            "flash_attention_gflops": flash_total / 1e9, "standard_to_ring_flops_ratio": std_total / ring_total,
            "ring_to_flash_flops_ratio": ring_total / flash_total}
