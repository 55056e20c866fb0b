"""The oracle against vectors produced by the reference itself (tests/golden/make_golden.py): this is what pins
parity. CPU only. Tolerances are fp32 round-off (the reference computes the same formulas in a different order)."""
import os

import pytest
import torch

from oracle import attn_mlp_oracle as orc

TOL = 2e-5


@pytest.fixture(scope="module")
def mlp_vecs(golden_dir):
    return torch.load(os.path.join(golden_dir, "mlp_reference_vectors.pt"))


@pytest.fixture(scope="module")
def attn_vecs(golden_dir):
    return torch.load(os.path.join(golden_dir, "attention_reference_vectors.pt"))


@pytest.mark.parametrize("act,ours", [("gelu", "gelu_tanh"), ("relu", "relu"), ("swiglu", "swiglu")])
def test_fused_transformer_mlp_module(mlp_vecs, act, ours):
    # FusedTransformerMLP("gelu") is the tanh approximation (fused_mlp.py:340-341 -> FusedMLPGeluTanh)
    d = mlp_vecs[f"FusedTransformerMLP_{act}"]
    sd = d["state_dict"]
    y = orc.mlp_ref(d["x"], sd["mlp.fc1.weight"], sd["mlp.fc1.bias"], sd["mlp.fc2.weight"], sd["mlp.fc2.bias"], ours,
                    sd.get("mlp.fc1_gate.weight"), sd.get("mlp.fc1_gate.bias"))
    assert orc.max_abs_err(y, d["y"]) < TOL


def test_bare_fused_mlp_is_exact_gelu(mlp_vecs):
    d = mlp_vecs["FusedMLP_gelu_erf"]
    sd = d["state_dict"]
    y = orc.mlp_ref(d["x"], sd["fc1.weight"], sd["fc1.bias"], sd["fc2.weight"], sd["fc2.bias"], "gelu")
    assert orc.max_abs_err(y, d["y"]) < TOL
    y_tanh = orc.mlp_ref(d["x"], sd["fc1.weight"], sd["fc1.bias"], sd["fc2.weight"], sd["fc2.bias"], "gelu_tanh")
    assert orc.max_abs_err(y_tanh, d["y"]) > 1e-5  # the two GELUs are distinguishable at this tolerance


@pytest.mark.parametrize("act", ["gelu", "relu", "swiglu"])
def test_pytorch_fused_mlp_functional(mlp_vecs, act):
    d = mlp_vecs[f"pytorch_fused_mlp_{act}"]
    y = orc.mlp_ref(d["x"], d["w1"], d["b1"], d["w2"], d["b2"], act, d["wg"] if act == "swiglu" else None,
                    d["bg"] if act == "swiglu" else None)
    assert orc.max_abs_err(y, d["y"]) < 5e-5


def test_layernorm_vs_reference(golden_dir):
    vecs = torch.load(os.path.join(golden_dir, "layernorm_reference_vectors.pt"))
    d = vecs["pytorch_layernorm"]
    assert orc.max_abs_err(orc.layernorm_ref(d["x"], d["w"], d["b"], d["eps"]), d["y"]) < TOL
    d = vecs["pytorch_layernorm_residual"]
    assert orc.max_abs_err(orc.layernorm_ref(d["x"], d["w"], d["b"], d["eps"], d["r"], d["alpha"]), d["y"]) < TOL
    d = vecs["pytorch_layernorm_nobias"]
    assert orc.max_abs_err(orc.layernorm_ref(d["x"], d["w"], None, d["eps"]), d["y"]) < TOL


def _bhsd_to_bshd(t):
    return t.permute(0, 2, 1, 3).contiguous()


@pytest.mark.parametrize("name,causal", [("ring_fallback_noncausal", False), ("ring_fallback_causal_finite_mask", True),
                                         ("ring_fallback_cross", False)])
def test_attention_vs_reference_online_softmax(attn_vecs, name, causal):
    d = attn_vecs[name]
    q, k, v = (_bhsd_to_bshd(d[n]) for n in ("q", "k", "v"))
    o, lse = orc.attention_ref(q, k, v, causal=causal)
    B, S, H, D = o.shape
    assert orc.max_abs_err(o.reshape(B, S, H * D), d["y"]) < TOL
    assert torch.isfinite(lse).all()


def test_attention_vs_reference_eager_softmax(attn_vecs):
    d = attn_vecs["sp_local_attention"]
    q, k, v = (_bhsd_to_bshd(d[n]) for n in ("q", "k", "v"))
    o, _ = orc.attention_ref(q, k, v)
    assert orc.max_abs_err(o, _bhsd_to_bshd(d["y"])) < TOL
    d = attn_vecs["sp_local_attention_padmask"]
    o, _ = orc.attention_ref(q, k, v, kv_lens=d["kv_lens"].to(torch.int32))
    assert orc.max_abs_err(o, _bhsd_to_bshd(d["y"])) < TOL


def test_lse_merge_reproduces_full_attention():
    torch.manual_seed(0)
    q, k, v = torch.randn(2, 40, 4, 32), torch.randn(2, 96, 2, 32), torch.randn(2, 96, 2, 32)
    full, lse_full = orc.attention_ref(q, k, v)
    o1, l1 = orc.attention_ref(q, k[:, :50], v[:, :50])
    o2, l2 = orc.attention_ref(q, k[:, 50:], v[:, 50:])
    o, lse = orc.lse_merge_ref(o1, l1, o2, l2)
    assert orc.max_abs_err(o, full) < 1e-5 and orc.max_abs_err(lse, lse_full) < 1e-5


def test_fully_masked_rows_are_zero_with_minus_inf_lse():
    torch.manual_seed(0)
    q, k, v = torch.randn(1, 8, 2, 16), torch.randn(1, 8, 2, 16), torch.randn(1, 8, 2, 16)
    o, lse = orc.attention_ref(q, k, v, causal=True, causal_offset=-3)
    assert torch.equal(o[:, :3], torch.zeros_like(o[:, :3]))
    assert torch.isinf(lse[:, :, :3]).all() and torch.isfinite(lse[:, :, 3:]).all()


def test_decode_matches_prefill_last_row_and_paged_equals_contiguous():
    torch.manual_seed(0)
    B, S, Hq, Hkv, D, bs = 2, 48, 4, 2, 16, 16
    q, k, v = torch.randn(B, S, Hq, D), torch.randn(B, S, Hkv, D), torch.randn(B, S, Hkv, D)
    o_full, lse_full = orc.attention_ref(q, k, v, causal=True)
    lens = torch.tensor([S, S], dtype=torch.int32)
    o_dec, lse_dec = orc.decode_attention_ref(q[:, -1], k, v, lens)
    assert orc.max_abs_err(o_dec, o_full[:, -1]) < 1e-6
    assert orc.max_abs_err(lse_dec, lse_full[:, :, -1]) < 1e-5
    # paged layout [num_blocks, L, block_size, Hkv, D]
    nblk = B * S // bs
    perm = torch.randperm(nblk).view(B, S // bs).to(torch.int32)
    kc, vc = torch.zeros(nblk, 2, bs, Hkv, D), torch.zeros(nblk, 2, bs, Hkv, D)
    for b in range(B):
        for t in range(S):
            kc[perm[b, t // bs], 1, t % bs] = k[b, t]
            vc[perm[b, t // bs], 1, t % bs] = v[b, t]
    o_p, _ = orc.decode_attention_ref(q[:, -1], kc, vc, lens, block_tables=perm, layer_idx=1)
    assert orc.max_abs_err(o_p, o_dec) < 1e-6


def test_ring_emulation_contiguous_and_zigzag():
    torch.manual_seed(0)
    B, S, H, D, n = 1, 64, 2, 16, 4
    q, k, v = torch.randn(B, S, H, D), torch.randn(B, S, H, D), torch.randn(B, S, H, D)
    full, _ = orc.attention_ref(q, k, v, causal=True)
    # contiguous partition (communication.py:651-659)
    pos = [torch.arange(r * S // n, (r + 1) * S // n) for r in range(n)]
    outs = orc.ring_attention_ref([q[:, p] for p in pos], [k[:, p] for p in pos], [v[:, p] for p in pos], True, pos)
    for p, o in zip(pos, outs):
        assert orc.max_abs_err(o, full[:, p]) < 1e-5
    # zigzag partition: rank r owns chunks r and 2n-1-r
    c = S // (2 * n)
    pos = [torch.cat([torch.arange(r * c, (r + 1) * c), torch.arange((2 * n - 1 - r) * c, (2 * n - r) * c)]) for r in range(n)]
    outs = orc.ring_attention_ref([q[:, p] for p in pos], [k[:, p] for p in pos], [v[:, p] for p in pos], True, pos)
    for p, o in zip(pos, outs):
        assert orc.max_abs_err(o, full[:, p]) < 1e-5


def test_tp_mlp_emulation_equals_dense():
    torch.manual_seed(0)
    x = torch.randn(10, 32)
    wu, bu, wg, bg = torch.randn(64, 32), torch.randn(64), torch.randn(64, 32), torch.randn(64)
    wd, bd = torch.randn(32, 64), torch.randn(32)
    for act in ("gelu_tanh", "swiglu"):
        dense = orc.mlp_ref(x, wu, bu, wd, bd, act, wg, bg)
        for tp in (2, 4):
            assert orc.max_abs_err(orc.tp_mlp_ref(x, wu, bu, wd, bd, act, tp, wg, bg), dense) < 1e-3
