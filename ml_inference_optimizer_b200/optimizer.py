"""``Optimizer`` — the user-facing entry point the reference documents in its README (README.md:56-82) but never
implements:

    from ml_inference_optimizer import Optimizer
    optimizer = Optimizer(model)
    profile   = optimizer.profile()
    optimized = optimizer.optimize(use_flash_attention=True, use_fused_mlp=True, tensor_parallel_size=1)
    text/ids  = optimized.generate(input_text="...", max_new_tokens=100)

``optimize`` swaps attention blocks for the K1/K2-backed modules (``kernels.attention.flash_attention.ModelConverter``)
and MLP blocks for the K3-backed ``FusedTransformerMLP`` (``kernels.mlp.fused_mlp.MLPConverter``), copying weights;
``tensor_parallel_size > 1`` shards MLPs with ``parallelism.tensor_parallel.ModelParallelConverter`` (needs an
initialised process group of that size).
"""
from __future__ import annotations

import time
from typing import Any, Dict, Optional

import torch
import torch.nn as nn

from .kernels.attention.flash_attention import FlashAttentionConfig, ModelConverter
from .kernels.mlp.fused_mlp import FusedMLPConfig, MLPConverter

__all__ = ["Optimizer"]


def _precision_name(dtype: torch.dtype) -> str:
    return {torch.float16: "fp16", torch.bfloat16: "bf16"}.get(dtype, "bf16")


class Optimizer:
    def __init__(self, model: nn.Module, tokenizer: Any = None):
        self.model = model
        self.tokenizer = tokenizer
        self.applied: Dict[str, Any] = {}

    # -------------------------------------------------------------------------------------------------------
    def profile(self, sample_inputs: Optional[Dict[str, torch.Tensor]] = None, iters: int = 3) -> Dict[str, Any]:
        """Module census plus (if ``sample_inputs`` is given) a timed forward — the data the README's ``profile()``
        promises for deciding what to optimise."""
        model = self.model
        n_params = sum(p.numel() for p in model.parameters())
        conv_attn, conv_mlp = ModelConverter(), MLPConverter()
        attn = [n for n, m in model.named_modules() if conv_attn._is_attention_module(m)]
        mlps = [n for n, m in model.named_modules() if conv_mlp._detect_mlp_type(m) is not None]
        out: Dict[str, Any] = {"num_parameters": n_params, "attention_modules": attn, "mlp_modules": mlps,
                               "dtype": str(next(model.parameters()).dtype), "device": str(next(model.parameters()).device)}
        if sample_inputs is not None:
            cuda = next(model.parameters()).is_cuda
            with torch.no_grad():
                model(**sample_inputs)
                if cuda:
                    torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(iters):
                    model(**sample_inputs)
                if cuda:
                    torch.cuda.synchronize()
            out["forward_ms"] = (time.perf_counter() - t0) / iters * 1e3
        return out

    # -------------------------------------------------------------------------------------------------------
    def optimize(self, use_flash_attention: bool = True, use_fused_mlp: bool = True, tensor_parallel_size: int = 1) -> nn.Module:
        model = self.model
        param = next(model.parameters())
        if not param.is_cuda:
            raise RuntimeError("Optimizer.optimize targets a model on a CUDA (sm_100) device: there is no CPU fallback")
        prec = _precision_name(param.dtype)
        if tensor_parallel_size > 1:
            import torch.distributed as dist
            from .parallelism.tensor_parallel import ModelParallelConverter, TensorParallelConfig
            if not dist.is_initialized() or dist.get_world_size() % tensor_parallel_size != 0:
                raise RuntimeError("tensor_parallel_size > 1 needs an initialised process group whose size it divides")
            cfg = TensorParallelConfig(world_size=dist.get_world_size(), tp_size=tensor_parallel_size)
            ModelParallelConverter(cfg).convert_model(model)
            self.applied["tensor_parallel_size"] = tensor_parallel_size
        if use_flash_attention:
            ModelConverter(FlashAttentionConfig(causal=True, precision=prec)).convert_model(model)
            self.applied["flash_attention"] = True
        if use_fused_mlp and tensor_parallel_size == 1:
            MLPConverter(FusedMLPConfig(precision=prec)).convert_model(model)
            self.applied["fused_mlp"] = True
        model.eval()
        self._attach_generate(model)
        return model

    # -------------------------------------------------------------------------------------------------------
    def _attach_generate(self, model: nn.Module) -> None:
        hf_generate = getattr(model, "generate", None)
        tokenizer = self.tokenizer

        def generate(input_text: Optional[str] = None, max_new_tokens: int = 20, input_ids: Optional[torch.Tensor] = None,
                     **kwargs):
            """Greedy generation. ``input_text`` needs a tokenizer (``Optimizer(model, tokenizer)``); ``input_ids`` works
            without one and returns ids."""
            dev = next(model.parameters()).device
            if input_ids is None:
                if input_text is None:
                    raise ValueError("pass input_text= or input_ids=")
                if tokenizer is None:
                    raise ValueError("input_text needs a tokenizer: construct Optimizer(model, tokenizer)")
                input_ids = tokenizer(input_text, return_tensors="pt")["input_ids"]
            input_ids = input_ids.to(dev)
            with torch.no_grad():
                if hf_generate is not None:
                    kwargs.setdefault("do_sample", False)
                    kwargs.setdefault("pad_token_id", 0)
                    out = hf_generate(input_ids=input_ids, max_new_tokens=max_new_tokens, **kwargs)
                else:
                    out = input_ids
                    for _ in range(max_new_tokens):
                        logits = model(out)
                        logits = logits[0] if isinstance(logits, tuple) else getattr(logits, "logits", logits)
                        out = torch.cat([out, logits[:, -1].argmax(-1, keepdim=True)], dim=1)
            if input_text is not None and tokenizer is not None:
                return tokenizer.decode(out[0], skip_special_tokens=True)
            return out

        model.generate = generate
