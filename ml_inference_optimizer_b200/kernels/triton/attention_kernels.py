"""Mirror of the functional API of ``kernels/triton/attention_kernels.py``: paged decode attention, KV append and the
single-GPU "ring" (chunked online-softmax) attention."""
from __future__ import annotations

from typing import Optional

import torch

from ... import ops

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
