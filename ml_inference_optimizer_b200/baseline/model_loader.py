"""``baseline/model_loader.py`` surface (reference :14-520): the loader registry that hands models to the runners and
to ``Optimizer``. Not on the hot path — kept so ``from baseline.model_loader import load_model`` stays a drop-in
(SURVEY.md Appendix A); what it returns is what the converters of this package (K1/K3 shells) take.

Differences from the reference, on purpose:
  * ``HuggingFaceModelLoader.load_model(name, random_init=True)`` builds the architecture from its config with seeded
    random weights instead of downloading a checkpoint — the only way to get "GPT-2 small" / Llama shapes on a box
    without network (synthetic data, random-init weights); without ``random_init`` it loads with
    ``local_files_only`` semantics left to ``transformers``;
  * no ``device_map="auto"``: the model goes to ``device`` in one piece (one process per GPU; sharding is
    ``parallelism/``'s job).
"""
from __future__ import annotations

import re
from abc import ABC, abstractmethod
from typing import Any, Dict, List, Optional, Tuple, Type

import torch
import torch.nn as nn

__all__ = ["BaseModelLoader", "HuggingFaceModelLoader", "TorchModelLoader", "ModelRegistry", "model_registry", "load_model",
           "register_custom_loader", "register_custom_pattern"]


class BaseModelLoader(ABC):
    """Abstract loader (reference :14-53)."""

    @abstractmethod
    def load_model(self, model_name: str, **kwargs) -> nn.Module: ...

    @abstractmethod
    def get_sample_input(self, batch_size: int, seq_len: int) -> torch.Tensor: ...

    @abstractmethod
    def get_model_config(self) -> Dict[str, Any]: ...


#: architectures that can be built without a checkpoint: name -> (config class name, config kwargs)
_KNOWN_CONFIGS = {
    "gpt2": ("GPT2Config", {}),  # defaults = GPT-2 small: 12 layers, 12 heads, 768 hidden, gelu_new (124.4 M parameters)
    "llama-2-7b": ("LlamaConfig", dict(hidden_size=4096, intermediate_size=11008, num_hidden_layers=32, num_attention_heads=32,
                                       num_key_value_heads=32, vocab_size=32000, max_position_embeddings=8192)),
    "llama-3-8b": ("LlamaConfig", dict(hidden_size=4096, intermediate_size=14336, num_hidden_layers=32, num_attention_heads=32,
                                       num_key_value_heads=8, vocab_size=128256, max_position_embeddings=8192,
                                       rope_theta=500000.0)),
}


class HuggingFaceModelLoader(BaseModelLoader):
    """Causal-LM loader (reference :56-153)."""

    def __init__(self, device: str = "cuda", dtype: Optional[torch.dtype] = None):
        self.device = device
        self.dtype = dtype
        self.model: Optional[nn.Module] = None
        self.tokenizer = None
        self.config = None

    def load_model(self, model_name: str, **kwargs) -> nn.Module:
        import transformers

        random_init = kwargs.pop("random_init", False)
        seed = kwargs.pop("seed", 0)
        torch_dtype = kwargs.pop("torch_dtype", self.dtype)
        if random_init:
            key = model_name.lower()
            if key in _KNOWN_CONFIGS:
                cls_name, cfg_kwargs = _KNOWN_CONFIGS[key]
                cfg_kwargs = {**cfg_kwargs, **kwargs.pop("config_overrides", {})}
                self.config = getattr(transformers, cls_name)(**cfg_kwargs)
            else:
                self.config = transformers.AutoConfig.from_pretrained(model_name, **kwargs.pop("config_overrides", {}))
            self.config._attn_implementation = kwargs.pop("attn_implementation", "eager")
            torch.manual_seed(seed)
            self.model = transformers.AutoModelForCausalLM.from_config(self.config)
        else:
            self.config = transformers.AutoConfig.from_pretrained(model_name)
            try:
                self.tokenizer = transformers.AutoTokenizer.from_pretrained(model_name)
            except Exception:  # a tokenizer is optional for the runners (they take token ids)
                self.tokenizer = None
            self.model = transformers.AutoModelForCausalLM.from_pretrained(model_name, torch_dtype=torch_dtype, **kwargs)
        if torch_dtype is not None:
            self.model = self.model.to(torch_dtype)
        self.model = self.model.to(self.device).eval()
        return self.model

    def get_sample_input(self, batch_size: int, seq_len: int) -> torch.Tensor:
        vocab = getattr(self.config, "vocab_size", 50257) if self.config is not None else 50257
        g = torch.Generator().manual_seed(0)
        return torch.randint(0, vocab, (batch_size, seq_len), generator=g).to(self.device)

    def get_model_config(self) -> Dict[str, Any]:
        if self.config is None:
            raise ValueError("Model not loaded. Call load_model() first.")  # reference :147
        return self.config.to_dict()


class TorchModelLoader(BaseModelLoader):
    """Plain ``torch.load`` / factory loader (reference :255-365): ``model_name`` is a path to a pickled module, or
    ``factory=callable`` builds it."""

    def __init__(self, device: str = "cuda", dtype: Optional[torch.dtype] = None):
        self.device, self.dtype, self.model, self.input_shape = device, dtype, None, None

    def load_model(self, model_name: str, **kwargs) -> nn.Module:
        factory = kwargs.pop("factory", None)
        self.input_shape = kwargs.pop("input_shape", None)
        self.model = factory(**kwargs) if factory is not None else torch.load(model_name, map_location="cpu", weights_only=False)
        if self.dtype is not None:
            self.model = self.model.to(self.dtype)
        self.model = self.model.to(self.device).eval()
        return self.model

    def get_sample_input(self, batch_size: int, seq_len: int) -> torch.Tensor:
        shape = (batch_size, seq_len) if self.input_shape is None else (batch_size, *self.input_shape)
        return torch.randn(*shape, device=self.device, dtype=self.dtype or torch.float32)

    def get_model_config(self) -> Dict[str, Any]:
        if self.model is None:
            raise ValueError("Model not loaded. Call load_model() first.")
        return {"class": type(self.model).__name__, "parameters": sum(p.numel() for p in self.model.parameters())}


class ModelRegistry:
    """name / regex-pattern -> loader class (reference :368-459)."""

    def __init__(self):
        self.loaders: Dict[str, Type[BaseModelLoader]] = {}
        self.patterns: List[Tuple[str, Type[BaseModelLoader]]] = []
        self.default_loader: Optional[Type[BaseModelLoader]] = None
        self.register_loader("huggingface", HuggingFaceModelLoader)
        self.register_loader("torch", TorchModelLoader)
        self.register_pattern(r".*\.(pt|pth)$", TorchModelLoader)
        self.set_default_loader(HuggingFaceModelLoader)

    def register_loader(self, name: str, loader_cls: Type[BaseModelLoader]) -> None:
        self.loaders[name] = loader_cls

    def register_pattern(self, pattern: str, loader_cls: Type[BaseModelLoader]) -> None:
        self.patterns.append((pattern, loader_cls))

    def set_default_loader(self, loader_cls: Type[BaseModelLoader]) -> None:
        self.default_loader = loader_cls

    def get_loader_for_model(self, model_name: str, loader_name: Optional[str] = None, **kwargs) -> BaseModelLoader:
        if loader_name is not None:
            if loader_name not in self.loaders:
                raise ValueError(f"Unknown loader: {loader_name}")
            return self.loaders[loader_name](**kwargs)
        for pattern, loader_cls in self.patterns:
            if re.match(pattern, model_name, re.IGNORECASE):
                return loader_cls(**kwargs)
        if self.default_loader:
            return self.default_loader(**kwargs)
        raise ValueError(f"Could not determine appropriate loader for model: {model_name}")

    def list_registered_loaders(self) -> Dict[str, Type[BaseModelLoader]]:
        return self.loaders.copy()


model_registry = ModelRegistry()


def load_model(model_name: str, loader_name: Optional[str] = None, device: str = "cuda", dtype: Optional[torch.dtype] = None,
               **kwargs) -> Tuple[nn.Module, BaseModelLoader]:
    """``(model, loader)`` through the registry (reference :466-489)."""
    loader = model_registry.get_loader_for_model(model_name, loader_name=loader_name, device=device, dtype=dtype)
    return loader.load_model(model_name, **kwargs), loader


def register_custom_loader(name: str, loader_cls: Type[BaseModelLoader]) -> None:
    model_registry.register_loader(name, loader_cls)


def register_custom_pattern(pattern: str, loader_cls: Type[BaseModelLoader]) -> None:
    model_registry.register_pattern(pattern, loader_cls)
