"""CPU checks of the host-side mirror of the reference API: signatures, validation, model surgery (weights copied),
mask lowering, partition helpers. No kernel runs here."""
import inspect
import math

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from kernels.attention.flash_attention import (FlashAttention3, FlashAttentionConfig, FlashAttentionLayer, FlashSelfAttention,
                                               ModelConverter, key_padding_mask_to_lengths)
from kernels.attention.ring_attention import RingAttentionConfig, RingCrossAttention, RingSelfAttention, calculate_theoretical_flops
from kernels.mlp.fused_mlp import (FusedMLP, FusedMLPConfig, FusedMLPGeluTanh, FusedMLPReLU, FusedMLPSwiGLU, FusedTransformerMLP,
                                   MLPConverter)
from ml_inference_optimizer import Optimizer
from parallelism import communication as comm
from parallelism.sequence_parallel import SequenceParallelConfig
from parallelism.tensor_parallel import TensorParallelConfig, TensorParallelMLP, activation_name


def test_signatures_match_reference_appendix_a():
    assert list(inspect.signature(FlashAttentionConfig).parameters) == [
        "block_size", "causal", "softmax_scale", "dropout_p", "return_softmax", "use_triton", "memory_efficient", "precision",
        "normalize_query", "fp8_ortho_matrix"]
    assert list(inspect.signature(FlashAttention3.forward).parameters) == ["self", "q", "k", "v", "mask"]
    assert list(inspect.signature(FlashAttentionLayer.__init__).parameters) == [
        "self", "hidden_size", "num_attention_heads", "config", "num_kv_heads"]
    assert list(inspect.signature(FusedMLPConfig).parameters) == [
        "activation_fn", "dropout_prob", "use_triton", "precision", "fuse_bias_gelu", "recompute_activation",
        "sequence_parallel", "tensor_parallel", "checkpoint_activation"]
    assert list(inspect.signature(FusedTransformerMLP.__init__).parameters) == [
        "self", "hidden_size", "intermediate_size", "activation_fn", "config"]
    assert list(inspect.signature(Optimizer.optimize).parameters) == [
        "self", "use_flash_attention", "use_fused_mlp", "tensor_parallel_size"]
    assert list(inspect.signature(RingAttentionConfig).parameters)[:4] == ["world_size", "chunk_size", "fuse_qkv", "use_flash_attention"]
    from kernels.triton.attention_kernels import triton_paged_attention_forward, triton_reshape_and_cache
    paged = inspect.signature(triton_paged_attention_forward).parameters
    assert list(paged)[:9] == [
        "query", "output", "k_cache", "v_cache", "block_tables", "context_lengths", "block_size", "max_seq_len", "layer_idx"]
    # anything past the reference's arguments (the short-q `causal` switch) must be optional
    assert all(p.default is not inspect.Parameter.empty for p in list(paged.values())[9:])
    assert list(inspect.signature(triton_reshape_and_cache).parameters) == [
        "key", "value", "k_cache", "v_cache", "block_tables", "context_lengths", "layer_idx"]
    from kernels.triton.mlp_kernels import triton_fused_mlp
    assert list(inspect.signature(triton_fused_mlp).parameters) == [
        "hidden_states", "fc1_weight", "fc1_bias", "fc2_weight", "fc2_bias", "activation", "fc1_gate_weight", "fc1_gate_bias"]
    assert list(inspect.signature(comm.all_reduce).parameters) == [
        "tensor", "op", "async_op", "group", "use_fp16", "use_bf16", "use_unbalanced", "stream"]


def test_config_validation():
    with pytest.raises(ValueError):
        FlashAttentionConfig(precision="int8")
    with pytest.raises(ValueError):
        RingAttentionConfig(world_size=0)
    with pytest.raises(ValueError):
        RingAttentionConfig(attention_dropout=1.0)
    with pytest.raises(ValueError):
        SequenceParallelConfig(world_size=4, sp_size=3)
    with pytest.raises(ValueError):
        SequenceParallelConfig(attention_handling="ulysses")
    with pytest.raises(AssertionError):
        TensorParallelConfig(world_size=4, tp_size=3)
    with pytest.raises(ValueError):
        FlashAttentionLayer(100, 3)


def test_module_attribute_names_and_shapes():
    layer = FlashAttentionLayer(256, 8, num_kv_heads=2)
    assert layer.k_proj.out_features == 2 * 32 and layer.q_proj.bias is not None
    fused = FlashSelfAttention(256, 8, num_kv_heads=2)
    assert fused.qkv_proj.out_features == 256 + 2 * 2 * 32
    assert set(dict(FusedMLPSwiGLU(64, 128).named_parameters())) == {
        "fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias", "fc1_gate.weight", "fc1_gate.bias"}
    assert isinstance(FusedTransformerMLP(64, 128, "gelu").mlp, FusedMLPGeluTanh)
    assert isinstance(FusedTransformerMLP(64, 128, "relu").mlp, FusedMLPReLU)
    assert FusedMLP(64, 128, FusedMLPConfig(activation_fn="gelu"))._activation() == "gelu_erf"
    assert FusedTransformerMLP(64, 128, "gelu").mlp._activation() == "gelu_tanh"
    ring = RingSelfAttention(128, 4, RingAttentionConfig())
    assert ring.qkv_proj.out_features == 384 and hasattr(ring, "out_proj")
    assert hasattr(RingCrossAttention(128, 4, RingAttentionConfig()), "k_proj")
    tp = TensorParallelMLP(64, 256)
    assert tp.dense_h_to_4h.weight.shape == (256, 64) and tp.dense_4h_to_h.weight.shape == (64, 256)


def test_no_cpu_fallback_in_modules():
    with pytest.raises(ValueError, match="CUDA"):
        FlashAttention3()(torch.randn(1, 8, 2, 64), torch.randn(1, 8, 2, 64), torch.randn(1, 8, 2, 64))
    with pytest.raises(ValueError, match="CUDA"):
        FusedTransformerMLP(64, 128)(torch.randn(1, 8, 64))


def test_key_padding_mask_lowering():
    m = torch.tensor([[1, 1, 1, 0, 0], [1, 1, 1, 1, 1]])
    assert key_padding_mask_to_lengths(m, 5).tolist() == [3, 5]
    assert key_padding_mask_to_lengths(m.unsqueeze(1).bool(), 5).tolist() == [3, 5]
    with pytest.raises(NotImplementedError):
        key_padding_mask_to_lengths(torch.tensor([[0, 1, 1, 1, 1]]), 5)  # left padding
    with pytest.raises(NotImplementedError):
        key_padding_mask_to_lengths(torch.ones(1, 5, 5), 5)  # dense mask


def test_activation_mapping():
    assert activation_name(F.gelu) == "gelu_erf"
    assert activation_name(nn.GELU(approximate="tanh")) == "gelu_tanh"
    assert activation_name(F.relu) == "relu"
    assert activation_name(F.silu) == "swiglu"
    with pytest.raises(ValueError):
        activation_name(torch.tanh)


def test_activation_resolution_is_strict():
    """converters resolve the activation from the module and refuse what has no fused epilogue (the reference
    maps by substring, fused_mlp.py:440-470): QuickGELU / Tanh / GeGLU must raise, exact vs tanh GELU stay distinct."""
    from transformers.activations import ACT2FN

    from ml_inference_optimizer_b200.kernels.mlp.fused_mlp import resolve_activation
    from ml_inference_optimizer_b200.parallelism.tensor_parallel import ModelParallelConverter, TensorParallelConfig

    assert resolve_activation(ACT2FN["gelu_new"]) == "gelu_tanh" and resolve_activation(ACT2FN["gelu"]) == "gelu_erf"
    assert resolve_activation(ACT2FN["gelu_pytorch_tanh"]) == "gelu_tanh" and resolve_activation(ACT2FN["silu"]) == "silu"
    assert resolve_activation("gelu_new") == "gelu_tanh" and resolve_activation(nn.GELU()) == "gelu_erf"
    for bad in (ACT2FN["quick_gelu"], ACT2FN["tanh"], nn.Mish(), torch.tanh, "mish", None):
        with pytest.raises(ValueError):
            resolve_activation(bad)

    class FcMLP(nn.Module):
        def __init__(self, act):
            super().__init__()
            self.fc1, self.fc2, self.activation_fn = nn.Linear(16, 32), nn.Linear(32, 16), act

    class Gated(nn.Module):
        def __init__(self, act):
            super().__init__()
            self.gate_proj, self.up_proj, self.down_proj = nn.Linear(16, 32), nn.Linear(16, 32), nn.Linear(32, 16)
            self.act_fn = act

    class Holder(nn.Module):
        def __init__(self, m):
            super().__init__()
            self.mlp = m

    h = Holder(FcMLP(nn.GELU()))                       # exact GELU stays exact
    MLPConverter().convert_model(h)
    assert h.mlp.inner.mlp._activation() == "gelu_erf"
    h = Holder(FcMLP(ACT2FN["gelu_new"]))
    MLPConverter().convert_model(h)
    assert h.mlp.inner.mlp._activation() == "gelu_tanh"
    for bad in (FcMLP(ACT2FN["quick_gelu"]), FcMLP(nn.SiLU()), FcMLP(nn.Tanh()), Gated(nn.GELU(approximate="tanh"))):
        with pytest.raises(ValueError):                 # SiLU without a gate, QuickGELU, Tanh, GeGLU: no fused epilogue
            MLPConverter().convert_model(Holder(bad))
    with pytest.raises(ValueError):
        ModelParallelConverter(TensorParallelConfig()).convert_model(Holder(Gated(nn.GELU())))


def test_hf_mask_lowering_to_kv_lens():
    """HF's 2-D keep mask / 4-D additive causal+padding mask -> per-sequence key counts; left padding, windows and
    non-causal structure raise (never a silent unmasked run)."""
    from ml_inference_optimizer_b200.kernels.attention.flash_attention import _HFAttentionAdapter as A

    B, S = 2, 6
    keep = torch.tensor([[1, 1, 1, 1, 0, 0], [1, 1, 1, 1, 1, 1]])
    assert A._mask_to_kv_lens(keep, B, S, S, True).tolist() == [4, 6]
    assert A._mask_to_kv_lens(torch.ones(B, S), B, S, S, True) is None      # no padding: nothing to mask
    causal = torch.tril(torch.ones(S, S, dtype=torch.bool))
    vis = causal[None, None] & keep.bool()[:, None, None, :]
    additive = torch.zeros(B, 1, S, S).masked_fill(~vis, torch.finfo(torch.float32).min)
    assert A._mask_to_kv_lens(additive, B, S, S, True).tolist() == [4, 6]
    assert A._mask_to_kv_lens(vis, B, S, S, True).tolist() == [4, 6]
    # cached decode: one query row against Sk keys
    assert A._mask_to_kv_lens(additive[:, :, -1:, :], B, 1, S, True).tolist() == [4, 6]
    with pytest.raises(NotImplementedError):                                   # left padding
        A._mask_to_kv_lens(torch.flip(keep, dims=[1]), B, S, S, True)
    window = vis & ~torch.tril(torch.ones(S, S, dtype=torch.bool), -3)[None, None]
    with pytest.raises(NotImplementedError):                                   # sliding window hides early keys
        A._mask_to_kv_lens(window, B, S, S, True)
    with pytest.raises(NotImplementedError):                                   # bidirectional mask on a causal module
        A._mask_to_kv_lens(torch.ones(B, 1, S, S, dtype=torch.bool), B, S, S, True)
    with pytest.raises(NotImplementedError):
        A._mask_to_kv_lens(torch.ones(B, 3, S, S), B, S, S, True)


def test_sequence_parallel_without_sp_group_stays_local():
    """sp_size == 1 has no SP group; ring / full handling must run the local kernel, not a ring over WORLD."""
    from ml_inference_optimizer_b200.parallelism.sequence_parallel import SequenceParallelAttention, SequenceParallelConfig

    calls = []

    class Backend:
        def attn(self, q, k, v, causal, scale):
            calls.append("attn")
            return torch.zeros_like(q), torch.zeros(q.shape[0], q.shape[2], q.shape[1])

    for handling in ("ring", "full", "local"):
        m = SequenceParallelAttention(32, 4, SequenceParallelConfig(world_size=1, sp_size=1, attention_handling=handling),
                                      attention_dropout=0.0)
        m.backend = Backend()
        m(torch.randn(1, 8, 32))
    assert calls == ["attn"] * 3


def test_converters_copy_weights_gpt2():
    from transformers import GPT2Config, GPT2LMHeadModel

    torch.manual_seed(0)
    cfg = GPT2Config(n_layer=2, n_head=4, n_embd=128, vocab_size=256, n_positions=64)
    model = GPT2LMHeadModel(cfg).eval()
    ref_attn = model.transformer.h[0].attn
    ref_mlp = model.transformer.h[0].mlp
    c_attn_w, c_fc_w, c_proj_b = ref_attn.c_attn.weight.clone(), ref_mlp.c_fc.weight.clone(), ref_mlp.c_proj.bias.clone()
    ModelConverter(FlashAttentionConfig(causal=True, precision="bf16")).convert_model(model)
    MLPConverter().convert_model(model)
    new_attn = model.transformer.h[0].attn.inner
    new_mlp = model.transformer.h[0].mlp.inner.mlp
    assert isinstance(new_attn, FlashSelfAttention) and new_attn.config.causal
    assert torch.equal(new_attn.qkv_proj.weight, c_attn_w.t())  # Conv1D stores [in, out]
    assert isinstance(new_mlp, FusedMLPGeluTanh)  # gelu_new -> tanh GELU
    assert torch.equal(new_mlp.fc1.weight, c_fc_w.t()) and torch.equal(new_mlp.fc2.bias, c_proj_b)
    assert model.transformer.h[0].attn.layer_idx == 0 and model.transformer.h[1].attn.layer_idx == 1


def test_mlp_converter_llama_style_and_load_from_standard():
    class LlamaMLP(nn.Module):
        def __init__(self):
            super().__init__()
            self.gate_proj, self.up_proj = nn.Linear(32, 96, bias=False), nn.Linear(32, 96, bias=False)
            self.down_proj = nn.Linear(96, 32, bias=False)
            self.act_fn = nn.SiLU()

    class Block(nn.Module):
        def __init__(self):
            super().__init__()
            self.mlp = LlamaMLP()

    blk = Block()
    gate_w = blk.mlp.gate_proj.weight.clone()
    MLPConverter().convert_model(blk)
    fused = blk.mlp.inner.mlp
    assert isinstance(fused, FusedMLPSwiGLU) and torch.equal(fused.fc1_gate.weight, gate_w)
    assert torch.count_nonzero(fused.fc1.bias) == 0  # missing biases become zeros (reference fused_mlp.py:549-555)
    t = FusedTransformerMLP(32, 96, "swiglu")
    sd = {"x.gate_proj.weight": torch.randn(96, 32), "x.up_proj.weight": torch.randn(96, 32), "x.down_proj.weight": torch.randn(32, 96)}
    t.load_from_standard_mlp(sd, prefix="x.")
    assert torch.equal(t.mlp.fc1_gate.weight, sd["x.gate_proj.weight"]) and torch.equal(t.mlp.fc2.weight, sd["x.down_proj.weight"])


def test_sequence_partition_roundtrip_single_process():
    x = torch.arange(2 * 16 * 3, dtype=torch.float32).view(2, 16, 3)
    for part in ("contiguous", "zigzag"):
        shards = [comm.scatter_along_sequence_dim(x, 4, partition=part, rank=r) for r in range(4)]
        assert all(s.shape == (2, 4, 3) for s in shards)
        if part == "zigzag":  # rank 0 owns chunks 0 and 7
            assert torch.equal(shards[0], torch.cat([x[:, 0:2], x[:, 14:16]], dim=1))
        covered = torch.cat(shards, dim=1)
        assert sorted(covered[0, :, 0].tolist()) == x[0, :, 0].tolist()
    with pytest.raises(ValueError):
        comm.scatter_along_sequence_dim(x[:, :15], 4, rank=0)


def test_flops_formula_and_memory_model():
    assert calculate_theoretical_flops(128, 2, 64, 4) == 3 * 2 * 128 * 64 * 64 + 2 * (2 * 4 * 128 * 128 * 16) + 2 * 128 * 64 * 64
    mem = FlashAttention3().get_theoretical_memory_usage(4096, 8, 12, 64)
    assert mem["memory_reduction_factor"] > 10


def test_optimizer_profile_and_cpu_refusal():
    from transformers import GPT2Config, GPT2LMHeadModel

    model = GPT2LMHeadModel(GPT2Config(n_layer=1, n_head=2, n_embd=64, vocab_size=128, n_positions=32))
    opt = Optimizer(model)
    prof = opt.profile()
    assert prof["attention_modules"] == ["transformer.h.0.attn"] and prof["mlp_modules"] == ["transformer.h.0.mlp"]
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        opt.optimize()


def test_paged_kv_cache_bookkeeping():
    from baseline.inference import BlockManager, PagedKVCache

    c = PagedKVCache(num_blocks=6, block_size=4, num_layers=2, num_heads=2, head_dim=8, dtype=torch.float32, device="cpu")
    k, v = c.get_physical_caches()
    # [num_blocks, L, block_size, Hkv, D] (reference baseline/inference.py:1077-1084); D is the PHYSICAL width: 64 or 128
    # columns (what the decode kernels are built for), a narrower head leaves zero columns behind it
    assert k.shape == (6, 2, 4, 2, 64) and c.head_dim == 8 and c.block_manager.physical_head_dim == 64
    assert PagedKVCache(2, 4, 1, 2, 128, dtype=torch.float32, device="cpu").get_physical_caches()[0].shape == (2, 1, 4, 2, 128)
    assert PagedKVCache(2, 4, 1, 2, 96, dtype=torch.float32, device="cpu").get_physical_caches()[0].shape == (2, 1, 4, 2, 128)
    with pytest.raises(ValueError, match="head_dim"):
        PagedKVCache(2, 4, 1, 2, 136, dtype=torch.float32, device="cpu")
    c.allocate_blocks_for_sequence(0, 6)
    assert len(c.get_block_table(0)) == 2 and c.get_sequence_length(0) == 6
    for _ in range(3):
        c.append_token(0)  # 9 tokens -> third block
    assert len(c.get_block_table(0)) == 3 and c.get_sequence_length(0) == 9
    c.allocate_blocks_for_sequence(1, 4)
    bt, lens = c.device_tables([0, 1])
    assert bt.dtype == torch.int32 and bt.shape == (2, 3) and lens.tolist() == [9, 4]
    c.allocate_blocks_for_sequence(2, 8)
    with pytest.raises(MemoryError):
        c.allocate_blocks_for_sequence(3, 4)  # 6 blocks exhausted
    c.free_sequence(0)
    assert c.block_manager.get_num_free_blocks() == 3
    # prefill scatter lands where the block table says
    kk = torch.arange(1 * 4 * 2 * 8, dtype=torch.float32).view(1, 4, 2, 8)
    c.write_prefill(1, [1], kk, kk + 1)
    blk = c.get_block_table(1)[0]
    assert torch.equal(k[blk, 1, ..., :8], kk[0]) and torch.equal(v[blk, 1, ..., :8], kk[0] + 1) and k[blk, 0].abs().sum() == 0
    assert k[blk, 1, ..., 8:].abs().sum() == 0 and v[blk, 1, ..., 8:].abs().sum() == 0      # the padding columns stay zero
    bm = BlockManager(2, 4, 1, 1, 8, torch.float32, "cpu")
    b0 = bm.allocate_block()
    bm.increase_ref_count(b0)
    bm.free_block(b0)
    assert bm.get_num_free_blocks() == 1  # still referenced once
    bm.free_block(b0)
    assert bm.get_num_free_blocks() == 2


def test_inference_runner_metrics_cpu():
    from baseline.inference import BasicInferenceRunner

    lin = torch.nn.Linear(4, 4)
    out, metrics = BasicInferenceRunner(lin, device="cpu").run_inference(torch.randn(2, 4))
    assert out.shape == (2, 4) and "total_time_ms" in metrics


def test_benchmark_metric_definitions():
    from benchmarks import metrics as m

    assert m.calculate_throughput(8, 128, 2.0) == 4.0
    st = m.calculate_latency_statistics([0.1 * i for i in range(1, 101)])
    assert abs(st["p90"] - 9.1) < 1e-9 and abs(st["p99"] - 10.0) < 1e-9 and st["min"] == 0.1
    assert m.calculate_latency_statistics([])["mean"] == 0.0
    assert abs(m.calculate_scaling_efficiency(118.44, 16.25, 8) - 91.1) < 0.1
    a, b = torch.tensor([1.0, 2.0]), torch.tensor([1.1, 2.0])
    assert abs(m.calculate_relative_error(a, b) - 5.0) < 1e-4 and abs(m.calculate_max_absolute_error(a, b) - 0.1) < 1e-6
    assert m.calculate_relative_error(a, torch.zeros(3)) == float("inf")


def test_benchmark_runner_host_logic(tmp_path):
    """Row f4: the runner keeps the reference's config defaults, result tree, validation rules and JSON export
    (reference benchmarks/runners.py:28-50, :250-330); timing itself needs a GPU and fails loudly on CPU inputs."""
    from benchmarks.runners import BenchmarkConfig, BenchmarkRunner, ModelBenchmarkRunner

    cfg = BenchmarkConfig("toy", [1], [8], ["baseline"], save_results=False)
    assert cfg.devices == ["cuda:0"] and cfg.num_iterations == 100 and cfg.warmup_iterations == 10
    assert cfg.to_dict()["precision"] == "fp16"
    r = BenchmarkRunner(cfg, results_dir=str(tmp_path))
    with pytest.raises(ValueError):
        BenchmarkRunner(BenchmarkConfig("toy", [1], [8], [], precision="int3", save_results=False))
    with pytest.raises(NotImplementedError):
        r.setup_model_variants()
    a = torch.arange(6.0).view(2, 3)
    assert r.validate_model_outputs(a, a + 1e-5)
    assert not r.validate_model_outputs(a, a + 1.0)
    assert r.validate_model_outputs((a, a), (a, a)) and not r.validate_model_outputs((a,), (a, a))
    assert r.validate_model_outputs({"logits": a}, {"logits": a}) and not r.validate_model_outputs("x", "x")
    path = r.save_benchmark_results({"t": a, "nested": {"v": (a, 1)}}, "out.json")
    import json
    assert json.load(open(path))["nested"]["v"][1] == 1
    with pytest.raises(RuntimeError):
        r.measure_performance(torch.nn.Identity(), {"input": torch.zeros(1, 4)})
    m = ModelBenchmarkRunner(BenchmarkConfig("toy", [2], [8], ["nonsense"], save_results=False), lambda: torch.nn.Linear(2, 2))
    ids = m.generate_test_inputs(2, 8)["input_ids"]
    assert ids.shape == (2, 8) and torch.equal(ids, m.generate_test_inputs(2, 8)["input_ids"])


def test_add_paged_attention_requires_attention_layers():
    from baseline.model_utils import add_paged_attention_to_model

    with pytest.raises(ValueError):
        add_paged_attention_to_model(torch.nn.Sequential(torch.nn.Linear(4, 4)))


def test_model_loader_registry_and_random_init():
    """``baseline.model_loader.load_model`` returns ``(model, loader)`` (reference baseline/model_loader.py:466-489); without
    network the HF loader builds the named architecture from its config with seeded random weights."""
    from baseline.model_loader import (BaseModelLoader, TorchModelLoader, load_model, model_registry, register_custom_loader,
                                       register_custom_pattern)

    model, loader = load_model("gpt2", device="cpu", dtype=torch.float32, random_init=True,
                               config_overrides={"n_layer": 1, "n_embd": 64, "n_head": 4, "vocab_size": 97})
    assert type(model).__name__ == "GPT2LMHeadModel" and loader.get_model_config()["n_layer"] == 1
    ids = loader.get_sample_input(2, 5)
    assert ids.shape == (2, 5) and int(ids.max()) < 97
    with torch.no_grad():
        assert model(ids).logits.shape == (2, 5, 97)
    m2, _ = load_model("gpt2", device="cpu", random_init=True, config_overrides={"n_layer": 1, "n_embd": 64, "n_head": 4, "vocab_size": 97})
    assert torch.equal(m2.transformer.wte.weight, model.transformer.wte.weight)  # seeded

    class Toy(BaseModelLoader):
        def __init__(self, device="cpu", dtype=None):
            self.device = device
        def load_model(self, model_name, **kw):
            return nn.Linear(3, 3)
        def get_sample_input(self, b, s):
            return torch.zeros(b, 3)
        def get_model_config(self):
            return {"toy": True}

    register_custom_loader("toy", Toy)
    register_custom_pattern(r"^toy-.*", Toy)
    assert "toy" in model_registry.list_registered_loaders()
    m, l = load_model("toy-1", device="cpu")
    assert isinstance(m, nn.Linear) and l.get_model_config() == {"toy": True}
    m, l = load_model("whatever", loader_name="torch", device="cpu", factory=lambda: nn.Linear(2, 2))
    assert isinstance(l, TorchModelLoader) and l.get_model_config()["parameters"] == 6
    with pytest.raises(ValueError):
        load_model("x", loader_name="nope", device="cpu")


def test_measurement_helpers_mirror_the_reference_signatures_and_result_keys():
    """The reference ships benchmark_* / compare_with_* / validate_* / profile_memory_usage helpers next to each kernel
    (flash_attention_kernels.py:1786-2060, attention_kernels.py:1643-1800, mlp_kernels.py:810-1090,
    layernorm_kernels.py:318-600, fused_layernorm_qkv.py:707-1060, ring_attention.py:838-1040). The mirrors take the same
    positional arguments and report the same keys; without a GPU they report zeros exactly as the reference's do (a report,
    not a fallback: nothing is computed)."""
    from kernels.attention import ring_attention as ra
    from kernels.triton import attention_kernels as ak
    from kernels.triton import flash_attention_kernels as fk
    from kernels.triton import fused_layernorm_qkv as lq
    from kernels.triton import layernorm_kernels as ln
    from kernels.triton import mlp_kernels as mk

    want = {
        fk.benchmark_flash_attention: ["seq_len", "batch_size", "num_heads", "head_dim", "device", "causal", "iterations", "warmup"],
        fk.compare_with_standard_attention: ["seq_len", "batch_size", "num_heads", "head_dim", "device"],
        fk.compare_with_xformers: ["seq_len", "batch_size", "num_heads", "head_dim", "device"],
        ak.compare_with_flash_attention: ["seq_len", "batch_size", "hidden_size", "num_heads"],
        ak.calculate_attention_theoretical_flops: ["seq_len", "batch_size", "hidden_size", "num_heads"],
        mk.benchmark_fused_mlp: ["batch_size", "seq_len", "hidden_size", "intermediate_size", "activation", "device", "dtype",
                                 "num_warmup", "num_iter"],
        mk.validate_fused_mlp: ["batch_size", "seq_len", "hidden_size", "intermediate_size", "activation", "device", "dtype"],
        mk.profile_memory_usage: ["batch_size", "seq_len", "hidden_size", "intermediate_size", "activation", "device"],
        ln.benchmark_layernorm: ["batch_size", "seq_len", "hidden_size", "device", "iterations", "warmup"],
        ln.compare_with_torch_layernorm: ["batch_size", "seq_len", "hidden_size", "device"],
        ln.profile_memory_usage: ["batch_size", "seq_len", "hidden_size", "device"],
        lq.benchmark_fused_layernorm_qkv: ["batch_size", "seq_len", "hidden_size", "num_heads", "num_kv_heads", "device",
                                           "iterations", "warmup"],
        lq.compare_with_unfused_implementation: ["batch_size", "seq_len", "hidden_size", "num_heads", "num_kv_heads", "device"],
        lq.profile_memory_usage: ["batch_size", "seq_len", "hidden_size", "num_heads", "num_kv_heads", "device"],
        ra.benchmark_ring_attention: ["seq_len", "batch_size", "hidden_size", "num_heads"],
        ra.compare_with_standard_attention: ["seq_len", "batch_size", "hidden_size", "num_heads"],
        ra.calculate_theoretical_flops: ["seq_len", "batch_size", "hidden_size", "num_heads"],
    }
    for fn, names in want.items():
        got = list(inspect.signature(fn).parameters)
        assert got[:len(names)] == names, (fn.__module__, fn.__name__, got)
    # the reference's operation-count model, values taken from the reference function itself (tests/golden is not needed:
    # three integers per case)
    f = ak.calculate_attention_theoretical_flops(4096, 2, 1024, 16)
    assert f["standard_attention_gflops"] == pytest.approx(105.763569664) and f["ring_attention_gflops"] == pytest.approx(106.837311488)
    assert f["flash_attention_gflops"] == pytest.approx(103.079215104)
    f = ak.calculate_attention_theoretical_flops(100, 1, 256, 4)     # ragged last chunk
    assert f["standard_to_ring_flops_ratio"] == pytest.approx(0.9974695075661724)
    assert f["ring_to_flash_flops_ratio"] == pytest.approx(1.0089358660130718)
    if not torch.cuda.is_available():
        assert fk.benchmark_flash_attention(128, 1, 2, 64) == {
            "sequence_length": 128, "batch_size": 1, "num_heads": 2, "head_dim": 64, "causal": False, "flash_attention_ms": 0.0,
            "pytorch_attention_ms": 0.0, "speedup": 0.0}
        assert fk.compare_with_standard_attention(128, 1, 2, 64)["is_correct"] is False
        assert mk.validate_fused_mlp(1, 8, 64, 128) == {"is_correct": False, "max_diff": 0.0}
        assert mk.benchmark_fused_mlp(1, 8, 64, 128)["triton_time_ms"] == 0.0
        assert ln.compare_with_torch_layernorm(1, 8, 64) == {"max_difference": 0.0, "is_correct": False}
        assert lq.compare_with_unfused_implementation(1, 8, 64)["is_correct"] is False
        assert set(ln.profile_memory_usage(1, 8, 64)) == {"torch_memory_mb", "triton_memory_mb", "memory_saving_percent"}
    with pytest.raises(ValueError, match="Unsupported activation"):
        mk._mlp_problem(1, 8, 64, 128, "tanh", "cpu", torch.bfloat16)


def test_tensor_parallel_linear_and_converter_surface():
    """reference tensor_parallel.py:152-159/:189-204/:271-327 (reset_parameters, get_master_weight), :497-509
    (transpose_for_scores), :601-614, :729-816 (convert_to_column/row_parallel, distribute_model) — single process (tp = 1
    shards are the whole matrix) plus a tp = 4 shard computed for a chosen rank."""
    from parallelism.tensor_parallel import ColumnParallelLinear, ModelParallelConverter, RowParallelLinear, TensorParallelAttention

    torch.manual_seed(0)
    lin = nn.Linear(32, 48)
    conv = ModelParallelConverter()                       # config defaults to a single rank, as in the reference (:622)
    col, row = conv.convert_to_column_parallel(lin), conv.convert_to_row_parallel(lin)
    assert isinstance(col, ColumnParallelLinear) and torch.equal(col.weight, lin.weight) and torch.equal(col.bias, lin.bias)
    assert isinstance(row, RowParallelLinear) and torch.equal(row.weight, lin.weight) and torch.equal(row.bias, lin.bias)
    assert col.get_master_weight() is col.weight and row.get_master_weight() is row.weight
    w0 = col.weight.clone()
    col.reset_parameters()
    bound = 1 / math.sqrt(32)
    assert not torch.equal(col.weight, w0) and col.weight.abs().max() <= bound + 1e-6 and col.bias.abs().max() <= bound
    cfg4 = TensorParallelConfig(world_size=4, tp_size=4)
    cfg4.tp_rank = lambda: 2                                # pick a rank without a process group
    conv4 = ModelParallelConverter(cfg4)
    c4, r4 = conv4.convert_to_column_parallel(lin), conv4.convert_to_row_parallel(lin)
    assert torch.equal(c4.weight, lin.weight[24:36]) and torch.equal(c4.bias, lin.bias[24:36])
    assert torch.equal(r4.weight, lin.weight[:, 16:24]) and torch.equal(r4.bias, lin.bias)   # row-parallel bias stays whole
    attn = TensorParallelAttention(64, 4, TensorParallelConfig())
    x = torch.randn(2, 5, 64)
    t = attn.transpose_for_scores(x)
    assert t.shape == (2, 4, 5, 16) and torch.equal(t[1, 3, 2], x[1, 2, 48:64])
    assert attn.num_heads_per_partition == 4 and attn.attention_head_size == 16
    assert attn.communication_schedule_optimized is False
    attn.optimize_communication_schedule()
    assert attn.communication_schedule_optimized is True
    tiny = nn.Sequential(nn.Linear(8, 8))
    assert set(ModelParallelConverter(TensorParallelConfig(world_size=2, tp_size=1)).distribute_model(tiny)) == {0, 1}


def test_parallel_utils_tensor_helpers():
    """reference parallel_utils.py:217-286, :426-556."""
    from parallelism import parallel_utils as pu

    t = torch.arange(24.0).view(2, 3, 4)
    assert torch.equal(pu.split_tensor_into_1d_equal_chunks(t, 1), t.view(-1))
    assert torch.equal(pu.gather_1d_tensor_chunks(t.view(-1), t.shape, 1), t)
    chunk = pu.split_tensor_into_1d_equal_chunks(t, 1)
    chunk[0] = -1
    assert t.view(-1)[0] == 0                                   # a copy, as in the reference (.clone())
    p = nn.Parameter(torch.zeros(4, 4))
    assert pu.get_parallel_tensor_info(p)["is_parallel"] is False
    pu.set_tensor_model_parallel_attributes(p, True, 0, 1)
    q = torch.zeros(2)
    pu.copy_tensor_model_parallel_attributes(q, p)
    info = pu.get_parallel_tensor_info(q)
    assert (info["is_parallel"], info["parallel_dim"], info["parallel_stride"], info["shape"]) == (True, 0, 1, q.shape)
    mask = torch.ones(2, 1, 8, 8)
    assert pu.create_attention_mask_for_tp(mask, 1) is mask and pu.create_attention_mask_for_tp(None, 4) is None
    assert pu.create_attention_mask_for_tp(mask, 4).shape == (2, 1, 8, 2)   # (rank 0 without a group: first key block)


def test_sequence_parallel_converter_surface():
    """reference sequence_parallel.py:326-342, :734-920: convert_model = deep copy -> attention layers -> MLP layers -> wrapped in
    SequenceShardedModule; weights are copied (the reference builds fresh random modules, Appendix B);
    partition_input_data / gather_output_data are inverses for both partitions."""
    from parallelism.sequence_parallel import (SequenceParallelAttention, SequenceParallelConverter, SequenceParallelMLP,
                                               SequenceShardedModule)

    class Attn(nn.Module):
        def __init__(self):
            super().__init__()
            self.num_attention_heads = 4
            self.q_proj, self.k_proj, self.v_proj, self.o_proj = (nn.Linear(32, 32) for _ in range(4))

    class MLP(nn.Module):
        def __init__(self):
            super().__init__()
            self.fc1, self.fc2, self.act = nn.Linear(32, 96), nn.Linear(96, 32), nn.ReLU()

    class Gated(nn.Module):
        def __init__(self):
            super().__init__()
            self.gate_proj, self.up_proj, self.down_proj = nn.Linear(32, 96), nn.Linear(32, 96), nn.Linear(96, 32)

    class Block(nn.Module):
        def __init__(self):
            super().__init__()
            self.attn, self.mlp, self.gated = Attn(), MLP(), Gated()

    torch.manual_seed(0)
    blk = Block()
    cfg = SequenceParallelConfig(world_size=1, sp_size=1)
    out = SequenceParallelConverter(cfg, causal=True).convert_model(blk)
    assert isinstance(out, SequenceShardedModule) and isinstance(blk.attn, Attn)          # the original is untouched
    new = out.module
    assert isinstance(new.attn, SequenceParallelAttention) and torch.equal(new.attn.query.weight, blk.attn.q_proj.weight)
    assert isinstance(new.mlp, SequenceParallelMLP) and torch.equal(new.mlp.dense_h_to_4h.weight, blk.mlp.fc1.weight)
    assert torch.equal(new.mlp.dense_4h_to_h.bias, blk.mlp.fc2.bias) and isinstance(new.mlp.activation, nn.ReLU)
    assert isinstance(new.gated, Gated)                                                   # gated blocks run unchanged on the shard
    out.config.communication_dtype = torch.bfloat16
    out.optimize_for_inference()
    assert all(p.dtype == torch.bfloat16 for p in out.parameters())
    class NoAct(nn.Module):                      # a two-Linear block that does not say what sits between them: refused, not guessed
        def __init__(self):
            super().__init__()
            self.fc1, self.fc2 = nn.Linear(32, 96), nn.Linear(96, 32)

    with pytest.raises(ValueError, match="no activation attribute"):
        SequenceParallelConverter(cfg).convert_mlp_layers(nn.Sequential(NoAct()))
    with pytest.raises(ValueError, match="no activation attribute"):
        from parallelism.tensor_parallel import ModelParallelConverter
        ModelParallelConverter().convert_model(nn.Sequential(NoAct()))
    ids = torch.arange(2 * 16).view(2, 16)
    for part in ("contiguous", "zigzag"):
        conv = SequenceParallelConverter(SequenceParallelConfig(world_size=4, sp_size=4), partition=part)
        parts = conv.partition_input_data({"input_ids": ids, "labels": torch.tensor([1, 2])})
        assert len(parts) == 4 and all(p["input_ids"].shape == (2, 4) and torch.equal(p["labels"], torch.tensor([1, 2])) for p in parts)
        assert torch.equal(conv.gather_output_data([p["input_ids"] for p in parts]), ids)


def test_kv_cache_runner_helpers_and_fusion_registry_cpu():
    """Row f1, host logic: the plain KVCache (reference baseline/inference.py:791-1043), the runner helpers (:616-784) and the
    module-pattern fusion registry (:26-281: matching, slot replacement, Sequential renumbering, weights copied)."""
    from baseline.inference import (BasicInferenceRunner, FusionPattern, FusionRegistry, KVCache, TransformerInferenceRunner,
                                    fusion_registry)
    from kernels.mlp.fused_mlp import FusedMLP as _FusedMLP

    c = KVCache(max_batch_size=2, max_seq_len=8, use_block_storage=True, block_size=4)
    with pytest.raises(RuntimeError, match="not initialized"):
        c.get_kv_cache(0)
    assert c.get_memory_usage() == {"total_memory_mb": 0}
    c.initialize(num_layers=2, num_heads=3, head_dim=8, dtype=torch.float32, device="cpu")
    assert c.get_kv_cache(0, 1) == (None, None)
    k, v = torch.randn(5, 3, 8), torch.randn(5, 3, 8)
    for layer in range(2):                      # the same 5 tokens into both layers: length 5, not 10
        c.append(layer, 1, k, v)
    assert c.current_seq_lengths == [0, 5]
    c.append(0, 1, k[:2], v[:2])
    gk, gv = c.get_kv_cache(0, 1)
    assert torch.equal(gk, torch.cat([k, k[:2]])) and torch.equal(gv, torch.cat([v, v[:2]])) and c.current_seq_lengths == [0, 7]
    with pytest.raises(ValueError, match="exceeds maximum"):
        c.append(0, 1, k[:2], v[:2])
    kc, vc, lens = c.decode_views(0)
    assert kc.shape == (2, 8, 3, 64) and lens.tolist() == [0, 7] and lens.dtype == torch.int32   # physical width 64: zero columns past head_dim
    assert kc[..., 8:].abs().sum() == 0
    use = c.get_memory_usage()
    assert use["total_memory_mb"] == pytest.approx(2 * 2 * 2 * 8 * 3 * 64 * 4 / 2 ** 20)
    assert use["memory_efficiency"] == pytest.approx((2 + 2) / (2 * 2 * 2))   # blocks touched: layer 0 -> 2, layer 1 -> 2, of 8
    c.reset()
    assert c.current_seq_lengths == [0, 0] and c.is_initialized
    c.clear()
    assert not c.is_initialized and c.k_caches == {}

    runner = BasicInferenceRunner(nn.Linear(4, 4), device="cpu")
    runner.warmup(torch.randn(2, 4), iterations=2)
    res = runner.run_batch_inference([torch.randn(2, 4), torch.randn(3, 4)])
    assert [r[0].shape[0] for r in res] == [2, 3] and runner.last_batch_metrics["avg_inference_time_ms"] > 0
    assert "model_inference" in runner.profile_model(torch.randn(2, 4), use_cuda=False)["table"]
    t = TransformerInferenceRunner(nn.Linear(4, 4), device="cpu")      # no CUDA: no cache, plain forward
    assert t.get_kv_cache_stats() == {"kv_cache_enabled": False} and t.run_inference(torch.randn(2, 4))[0].shape == (2, 4)

    assert [p.name for p in fusion_registry.patterns] == ["linear_gelu_linear", "linear_relu_linear"]
    seq = nn.Sequential(nn.LayerNorm(16), nn.Linear(16, 32), nn.ReLU(), nn.Linear(32, 16), nn.Dropout(0.0))
    fused = fusion_registry.fuse_modules(seq)
    assert isinstance(seq[1], nn.Linear) and list(fused._modules) == ["0", "1", "2"]       # copy by default, renumbered
    assert isinstance(fused[1], _FusedMLP) and fused[1]._activation() == "relu"
    assert torch.equal(fused[1].fc1.weight, seq[1].weight) and torch.equal(fused[1].fc2.bias, seq[3].bias)

    class Named(nn.Module):
        def __init__(self):
            super().__init__()
            self.up, self.act, self.down, self.norm = nn.Linear(16, 32, bias=False), nn.GELU(), nn.Linear(32, 16), nn.LayerNorm(16)

    m = fusion_registry.fuse_modules(Named(), inplace=True)
    assert isinstance(m.up, _FusedMLP) and m.up._activation() == "gelu_erf" and not hasattr(m, "act") and not hasattr(m, "down")
    assert torch.count_nonzero(m.up.fc1.bias) == 0                                       # a missing bias becomes zeros
    assert fusion_registry.fuse_modules(nn.Sequential(nn.Linear(16, 32), nn.GELU(approximate="tanh"), nn.Linear(32, 16)))[0]._activation() == "gelu_tanh"
    with pytest.raises(ValueError, match="hidden -> intermediate -> hidden"):
        fusion_registry.fuse_modules(nn.Sequential(nn.Linear(16, 32), nn.ReLU(), nn.Linear(32, 8)))
    reg = FusionRegistry()
    assert reg.fuse_modules(seq) is not seq and reg.find_matching_pattern([seq[0]]) is None
    reg.register_pattern(FusionPattern("ln_linear", [nn.LayerNorm, nn.Linear], lambda mods: nn.Sequential(*mods)))
    assert reg.patterns[0].description == "Fuses LayerNorm + Linear" and reg.patterns[0].match([seq[0], seq[1]])
    assert isinstance(reg.fuse_modules(seq)[0], nn.Sequential)


def test_model_utils_inspection_helpers():
    """reference baseline/model_utils.py:18-260, :455-598."""
    from baseline import model_utils as mu

    class Attn(nn.Module):
        def __init__(self):
            super().__init__()
            self.q_proj, self.k_proj, self.v_proj = nn.Linear(8, 8), nn.Linear(8, 8), nn.Linear(8, 8)

    class Block(nn.Module):
        def __init__(self):
            super().__init__()
            self.self_attn = Attn()
            self.mlp = nn.Sequential(nn.Linear(8, 16), nn.GELU(), nn.Linear(16, 8))
            self.drop = nn.Dropout(0.1)

    m = Block()
    size = mu.get_model_size(m)
    n = 3 * (8 * 8 + 8) + (8 * 16 + 16) + (16 * 8 + 8)
    assert size["total_params"] == size["trainable_params"] == n and size["param_bytes_per_element"] == 4
    assert size["param_memory_mb"] == pytest.approx(n * 4 / 2 ** 20) and size["param_dtype"] == "torch.float32"
    assert mu.get_attention_modules(m) == [m.self_attn]
    assert mu.get_mlp_modules(m) == [m.mlp, m.mlp[0], m.mlp[1], m.mlp[2]]   # "mlp" in the module path: the block and its children
    layers = mu.get_model_layers(m)
    assert m.self_attn.q_proj in layers and m.mlp not in layers and m.drop not in layers and m.self_attn not in layers
    assert [name for name, _ in mu.find_modules_by_type(m, nn.Linear)] == ["self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj",
                                                                          "mlp.0", "mlp.2"]
    assert next(mu.convert_precision(m, "bf16").parameters()).dtype == torch.bfloat16
    with pytest.raises(ValueError, match="Unsupported precision"):
        mu.convert_precision(m, "int3")
    assert mu.create_random_input(2, 3, 4, device="cpu").shape == (2, 3, 4)
    mu.freeze_layers(m, ["self_attn.*", "mlp.2.bias"])
    frozen = sorted(name for name, p in m.named_parameters() if not p.requires_grad)
    assert len(frozen) == 7 and "mlp.2.bias" in frozen and "mlp.0.weight" not in frozen
    m = m.float()
    sd = {"mlp.0.weight": torch.ones(16, 8), "mlp.2.weight": torch.ones(3, 3), "nope": torch.zeros(1)}
    res = mu.load_partial_weights(m, sd)
    assert torch.equal(m.mlp[0].weight, torch.ones(16, 8)) and "mlp.2.weight" not in res.unexpected_keys
    with pytest.raises(RuntimeError, match="Shape-mismatched keys"):
        mu.load_partial_weights(m, sd, strict=True)
