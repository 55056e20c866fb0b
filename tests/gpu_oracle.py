"""The fp32 oracle restated on the device, chunked, so the BASELINE-size configurations can be compared IN FULL (the CPU
oracle in oracle/attn_mlp_oracle.py would need hours and ~100 GB for C3). Test infrastructure only: plain torch fp32 on
CUDA (TF32 off), same formulas as the CPU oracle, and tests/test_gpu_fullsize.py pins it to the CPU oracle on small
shapes before using it."""
import math

import torch


def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def attention_ref_gpu(q, k, v, causal=False, softmax_scale=None, causal_offset=0, kv_lens=None, q_block=1024):
    """softmax(Q K^T * scale + mask) V and logsumexp in fp32, one (batch, query block) at a time.
    q [B,Sq,Hq,D], k/v [B,Sk,Hkv,D] -> o [B,Sq,Hq,D] fp32, lse [B,Hq,Sq] fp32 (oracle.attention_ref semantics)."""
    _no_tf32()
    B, Sq, Hq, D = q.shape
    Sk, Hkv = k.shape[1], k.shape[2]
    G = Hq // Hkv
    scale = softmax_scale if softmax_scale is not None else 1.0 / math.sqrt(D)
    o = torch.empty(B, Sq, Hq, D, dtype=torch.float32, device=q.device)
    lse = torch.empty(B, Hq, Sq, dtype=torch.float32, device=q.device)
    kpos = torch.arange(Sk, device=q.device)
    for b in range(B):
        kb = k[b].float().permute(1, 0, 2)                      # [Hkv,Sk,D]
        vb = v[b].float().permute(1, 0, 2)
        for s0 in range(0, Sq, q_block):
            s1 = min(Sq, s0 + q_block)
            qb = q[b, s0:s1].float().permute(1, 0, 2).reshape(Hkv, G, s1 - s0, D)   # head h = hkv * G + g
            sc = torch.einsum("kgqd,ksd->kgqs", qb, kb) * scale                      # [Hkv,G,q,Sk]
            vis = torch.ones(s1 - s0, Sk, dtype=torch.bool, device=q.device)
            if causal:
                qpos = torch.arange(s0, s1, device=q.device) + causal_offset
                vis &= kpos.unsqueeze(0) <= qpos.unsqueeze(1)
            if kv_lens is not None:
                vis &= (kpos < int(kv_lens[b])).unsqueeze(0)
            sc = sc.masked_fill(~vis, float("-inf"))
            l = torch.logsumexp(sc, dim=-1)
            safe = torch.where(torch.isinf(l), torch.zeros_like(l), l)
            p = torch.exp(sc - safe.unsqueeze(-1))
            p = torch.where(vis, p, torch.zeros_like(p))
            ob = torch.einsum("kgqs,ksd->kgqd", p, vb).reshape(Hq, s1 - s0, D)
            o[b, s0:s1] = ob.permute(1, 0, 2)
            lse[b, :, s0:s1] = l.reshape(Hq, s1 - s0)
    return o, lse


def decode_ref_gpu(q, k_cache, v_cache, lens, softmax_scale=None, b_block=4):
    """q [B,Hq,D] against a contiguous cache [B,S,Hkv,D] with lens[b] valid keys -> o [B,Hq,D] fp32."""
    _no_tf32()
    B, Hq, D = q.shape
    S, Hkv = k_cache.shape[1], k_cache.shape[2]
    G = Hq // Hkv
    scale = softmax_scale if softmax_scale is not None else 1.0 / math.sqrt(D)
    o = torch.empty(B, Hq, D, dtype=torch.float32, device=q.device)
    pos = torch.arange(S, device=q.device)
    for b0 in range(0, B, b_block):
        b1 = min(B, b0 + b_block)
        qb = q[b0:b1].float().reshape(b1 - b0, Hkv, G, D)
        sc = torch.einsum("bkgd,bskd->bkgs", qb, k_cache[b0:b1].float()) * scale
        vis = pos.unsqueeze(0) < lens[b0:b1].unsqueeze(1)
        sc = sc.masked_fill(~vis[:, None, None, :], float("-inf"))
        p = torch.softmax(sc, dim=-1)
        o[b0:b1] = torch.einsum("bkgs,bskd->bkgd", p, v_cache[b0:b1].float()).reshape(b1 - b0, Hq, D)
    return o


def mlp_ref_gpu(x, w_up, b_up, w_down, b_down, activation, w_gate=None, b_gate=None, t_block=4096):
    """fc2(act(fc1 x)) / fc2(silu(gate x) * up x) in fp32, a block of tokens at a time (oracle.mlp_ref semantics)."""
    _no_tf32()
    F = torch.nn.functional
    f = lambda t: None if t is None else t.float()
    wu, bu, wd, bd, wg, bg = f(w_up), f(b_up), f(w_down), f(b_down), f(w_gate), f(b_gate)
    out = torch.empty(x.shape[0], w_down.shape[0], dtype=torch.float32, device=x.device)
    for t0 in range(0, x.shape[0], t_block):
        xb = x[t0:t0 + t_block].float()
        up = F.linear(xb, wu, bu)
        if activation == "swiglu":
            hmid = F.silu(F.linear(xb, wg, bg)) * up
        elif activation in ("gelu_tanh", "gelu_new"):
            hmid = 0.5 * up * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (up + 0.044715 * up.pow(3))))
        elif activation in ("gelu", "gelu_erf"):
            hmid = F.gelu(up)
        elif activation == "relu":
            hmid = torch.relu(up)
        else:
            raise ValueError(activation)
        out[t0:t0 + t_block] = F.linear(hmid, wd, bd)
    return out
