"""Host-side mirror of the reference's ``kernels/mlp/fused_mlp.py``: same classes, constructor arguments, attribute
names (``fc1``, ``fc1_gate``, ``fc2``, ``mlp``) and state-dict keys; ``forward`` runs K3 (``b200_fused_mlp``):
tcgen05 GEMM1 with bias + activation (or the SwiGLU gate/up pair) fused into the epilogue, then tcgen05 GEMM2.

Activation per class, as in the reference: ``FusedMLPGeluTanh`` and
``FusedTransformerMLP("gelu")`` use the tanh approximation (fused_mlp.py:223-237, :340-341), the bare
``FusedMLP(activation_fn="gelu")`` the exact erf GELU (:162-163).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, List, Optional

import torch
import torch.nn as nn

from ... import ops

__all__ = ["FusedMLPConfig", "FusedMLP", "FusedMLPGeluTanh", "FusedMLPSwiGLU", "FusedMLPReLU", "FusedTransformerMLP",
           "MLPConverter", "resolve_activation"]


def resolve_activation(fn) -> str:
    """Strict map from an activation (string, function, ``nn.Module`` or HF activation object) to a fused epilogue:
    ``"gelu_tanh"``, ``"gelu_erf"``, ``"relu"`` or ``"silu"`` (the gate activation of SwiGLU). Anything without a fused
    epilogue (QuickGELU, plain Tanh, Mish, GeGLU gates ...) raises instead of being approximated — a converted model must
    compute what the original computed. (The reference maps by substring, fused_mlp.py:440-470.)"""
    import functools
    import torch.nn.functional as F

    if fn is None:
        raise ValueError("the MLP block has no activation attribute (act_fn / activation_fn / act / activation)")
    if isinstance(fn, functools.partial):
        if fn.func is F.gelu:
            return "gelu_tanh" if fn.keywords.get("approximate") == "tanh" else "gelu_erf"
        fn = fn.func
    if isinstance(fn, str):
        key = fn.lower()
        table = {"gelu_new": "gelu_tanh", "gelu_tanh": "gelu_tanh", "gelu_pytorch_tanh": "gelu_tanh", "gelu_fast": "gelu_tanh",
                 "gelu": "gelu_erf", "gelu_erf": "gelu_erf", "gelu_python": "gelu_erf", "relu": "relu", "silu": "silu",
                 "swish": "silu", "swiglu": "silu"}
        if key in table:
            return table[key]
        raise ValueError(f"activation {fn!r} has no fused epilogue (supported: gelu, gelu_new/gelu_tanh, relu, silu)")
    if isinstance(fn, nn.GELU):
        return "gelu_tanh" if fn.approximate == "tanh" else "gelu_erf"
    if isinstance(fn, nn.ReLU) or fn is F.relu or fn is torch.relu:
        return "relu"
    if isinstance(fn, nn.SiLU) or fn is F.silu:
        return "silu"
    if fn is F.gelu:
        return "gelu_erf"
    name = (getattr(fn, "__name__", None) or type(fn).__name__).lower()
    by_class = {"newgeluactivation": "gelu_tanh", "pytorchgelutanh": "gelu_tanh", "gelutanh": "gelu_tanh",
                "fastgeluactivation": "gelu_tanh",  # 0.5x(1+tanh(0.79788456x(1+0.044715x^2))): the same function
                "geluactivation": "gelu_erf", "reluactivation": "relu", "siluactivation": "silu", "gelu_new": "gelu_tanh",
                "gelu": "gelu_erf", "relu": "relu", "silu": "silu", "swish": "silu"}
    if name in by_class:
        if name == "geluactivation" and getattr(fn, "act", None) is not None and getattr(fn.act, "__name__", "") == "_gelu_python":
            return "gelu_erf"
        return by_class[name]
    raise ValueError(f"activation {fn!r} has no fused epilogue (supported: exact / tanh GELU, ReLU, SiLU gates); "
                     "refusing to convert rather than computing something else")


@dataclass
class FusedMLPConfig:
    """reference: kernels/mlp/fused_mlp.py:15-25 (field names kept; ``use_triton`` etc. are accepted and ignored)."""
    activation_fn: str = "gelu"
    dropout_prob: float = 0.0
    use_triton: bool = True
    precision: str = "fp16"
    fuse_bias_gelu: bool = True
    recompute_activation: bool = False
    sequence_parallel: bool = False
    tensor_parallel: bool = False
    checkpoint_activation: bool = False


_PRECISION = {"fp16": torch.float16, "bf16": torch.bfloat16}


class FusedMLP(nn.Module):
    """reference: kernels/mlp/fused_mlp.py:28-202. ``fc2(act(fc1(x)))`` with bias=True Linears."""

    _kernel_activation: Optional[str] = None  # subclasses pin the activation

    def __init__(self, hidden_size: int, intermediate_size: int, config: Optional[FusedMLPConfig] = None):
        super().__init__()
        self.config = config or FusedMLPConfig()
        self.hidden_size = hidden_size
        self.intermediate_size = intermediate_size
        self.fc1 = nn.Linear(hidden_size, intermediate_size, bias=True)
        self.fc2 = nn.Linear(intermediate_size, hidden_size, bias=True)
        self.dropout = nn.Dropout(self.config.dropout_prob) if self.config.dropout_prob > 0 else None

    def _activation(self) -> str:
        if self._kernel_activation is not None:
            return self._kernel_activation
        act = self.config.activation_fn
        if act == "gelu":
            return "gelu_erf"  # exact GELU for the bare module, fused_mlp.py:162-163
        if act in ("gelu_erf", "gelu_exact"):
            return "gelu_erf"
        if act in ("gelu_tanh", "gelu_new", "gelu_pytorch_tanh"):
            return "gelu_tanh"
        if act == "relu":
            return "relu"
        raise ValueError(f"Unsupported activation function: {act}")

    def _compute_dtype(self, x: torch.Tensor) -> torch.dtype:
        if x.dtype in (torch.float16, torch.bfloat16):
            return x.dtype
        return _PRECISION.get(self.config.precision, torch.bfloat16)

    def _cast(self, t: Optional[torch.Tensor], dt: torch.dtype) -> Optional[torch.Tensor]:
        return None if t is None else (t if t.dtype == dt else t.to(dt))

    def forward(self, hidden_states: torch.Tensor) -> torch.Tensor:
        if not hidden_states.is_cuda:
            raise ValueError("Input tensor must be on CUDA device: the B200 FusedMLP has no CPU fallback")  # :85-90
        if self.dropout is not None and self.training:
            raise NotImplementedError("dropout inside the fused MLP is not implemented (inference path)")
        orig = hidden_states.dtype
        dt = self._compute_dtype(hidden_states)
        x = self._cast(hidden_states, dt)
        gate_w = gate_b = None
        if hasattr(self, "fc1_gate"):
            gate_w, gate_b = self._cast(self.fc1_gate.weight, dt), self._cast(self.fc1_gate.bias, dt)
        y = ops.fused_mlp(x, self._cast(self.fc1.weight, dt), self._cast(self.fc1.bias, dt), self._cast(self.fc2.weight, dt),
                          self._cast(self.fc2.bias, dt), self._activation(), gate_w, gate_b)
        return y if y.dtype == orig else y.to(orig)  # output dtype = input dtype (:143-145)


class FusedMLPGeluTanh(FusedMLP):
    """reference :205-237 — tanh-approximate GELU (HF ``gelu_new``)."""
    _kernel_activation = "gelu_tanh"


class FusedMLPReLU(FusedMLP):
    """reference :299-315."""
    _kernel_activation = "relu"


class FusedMLPSwiGLU(FusedMLP):
    """reference :240-296 — ``fc2(silu(fc1_gate(x)) * fc1(x))``; ``fc1`` is the value/up projection."""
    _kernel_activation = "swiglu"

    def __init__(self, hidden_size: int, intermediate_size: int, config: Optional[FusedMLPConfig] = None):
        super().__init__(hidden_size, intermediate_size, config)
        self.fc1_gate = nn.Linear(hidden_size, intermediate_size, bias=True)


class FusedTransformerMLP(nn.Module):
    """reference :318-396 — picks the fused implementation from ``activation_fn`` and exposes it as ``self.mlp``."""

    def __init__(self, hidden_size: int, intermediate_size: int, activation_fn: str = "gelu",
                 config: Optional[FusedMLPConfig] = None):
        super().__init__()
        self.config = config or FusedMLPConfig()
        self.config.activation_fn = activation_fn
        if activation_fn == "gelu":
            self.mlp = FusedMLPGeluTanh(hidden_size, intermediate_size, self.config)
        elif activation_fn == "swiglu":
            self.mlp = FusedMLPSwiGLU(hidden_size, intermediate_size, self.config)
        elif activation_fn == "relu":
            self.mlp = FusedMLPReLU(hidden_size, intermediate_size, self.config)
        else:
            self.mlp = FusedMLP(hidden_size, intermediate_size, self.config)

    def forward(self, hidden_states: torch.Tensor) -> torch.Tensor:
        return self.mlp(hidden_states)

    def load_from_standard_mlp(self, state_dict: Dict[str, torch.Tensor], prefix: str = "") -> None:
        """reference :362-396 (BERT-style ``dense`` / ``output.dense`` and Llama-style ``gate/up/down_proj`` keys)."""
        key_mapping = {
            f"{prefix}dense.weight": "mlp.fc1.weight", f"{prefix}dense.bias": "mlp.fc1.bias",
            f"{prefix}output.dense.weight": "mlp.fc2.weight", f"{prefix}output.dense.bias": "mlp.fc2.bias",
        }
        if isinstance(self.mlp, FusedMLPSwiGLU) and f"{prefix}gate_proj.weight" in state_dict:
            key_mapping.update({
                f"{prefix}gate_proj.weight": "mlp.fc1_gate.weight", f"{prefix}gate_proj.bias": "mlp.fc1_gate.bias",
                f"{prefix}up_proj.weight": "mlp.fc1.weight", f"{prefix}up_proj.bias": "mlp.fc1.bias",
                f"{prefix}down_proj.weight": "mlp.fc2.weight", f"{prefix}down_proj.bias": "mlp.fc2.bias",
            })
        own = self.state_dict()
        with torch.no_grad():
            for src, dst in key_mapping.items():
                if src in state_dict and dst in own:
                    own[dst].copy_(state_dict[src])
        self.load_state_dict(own)


class _MLPAdapter(nn.Module):
    """Stand-in for a HuggingFace MLP block: same call signature (a single tensor), fused execution inside."""

    def __init__(self, inner: FusedTransformerMLP):
        super().__init__()
        self.inner = inner

    def forward(self, hidden_states, *args, **kwargs):
        return self.inner(hidden_states)


class MLPConverter:
    """reference: kernels/mlp/fused_mlp.py:399-614 — scan a model and replace MLP blocks by fused ones, COPYING the
    weights (GPT-2 ``c_fc``/``c_proj`` Conv1D stored [in,out]; Llama ``gate/up/down_proj`` without biases — the
    missing biases become zeros, as the reference does at :549-555)."""

    def __init__(self, config: Optional[FusedMLPConfig] = None):
        self.config = config or FusedMLPConfig()
        self.activation_map = {"gelu": "gelu", "relu": "relu", "silu": "silu", "swish": "silu", "swiglu": "swiglu",
                               "gelu_new": "gelu"}

    @staticmethod
    def _module_activation(module: nn.Module):
        for attr in ("act_fn", "activation_fn", "act", "activation"):
            if getattr(module, attr, None) is not None:
                return getattr(module, attr)
        cfg = getattr(module, "config", None)
        for attr in ("hidden_act", "activation_function"):
            if cfg is not None and getattr(cfg, attr, None) is not None:
                return getattr(cfg, attr)
        return None

    def _detect_mlp_type(self, module: nn.Module) -> Optional[Dict[str, Any]]:
        """Structure decides the kind; the activation is RESOLVED from the module (``resolve_activation``), never guessed:
        a block whose activation has no fused epilogue raises."""
        cls = type(module).__name__
        if isinstance(module, (FusedMLP, FusedTransformerMLP, _MLPAdapter)):
            return None
        # FusedTransformerMLP spelling: "gelu" = tanh GELU (reference :340-341), "gelu_erf" = exact
        spell = {"gelu_tanh": "gelu", "gelu_erf": "gelu_erf", "relu": "relu"}

        def plain(kind, hidden, inter, default=None):
            act = self._module_activation(module)
            name = resolve_activation(act if act is not None else default)
            if name not in spell:
                raise ValueError(f"{cls}: activation {name!r} has no fused epilogue for an un-gated MLP")
            return {"kind": kind, "hidden": hidden, "intermediate": inter, "activation": spell[name]}

        if hasattr(module, "c_fc") and hasattr(module, "c_proj") and "mlp" in cls.lower():
            w = module.c_fc.weight  # Conv1D: [in, out]
            return plain("gpt2", w.shape[0], w.shape[1])
        if all(hasattr(module, a) for a in ("gate_proj", "up_proj", "down_proj")):
            gate_act = resolve_activation(self._module_activation(module))
            if gate_act != "silu":
                raise ValueError(f"{cls}: gated MLP with a {gate_act} gate (GeGLU-style) has no fused epilogue; only SiLU gates "
                                 "(SwiGLU) are fused")
            return {"kind": "llama", "hidden": module.up_proj.in_features, "intermediate": module.up_proj.out_features,
                    "activation": "swiglu"}
        if hasattr(module, "fc1") and hasattr(module, "fc2") and isinstance(module.fc1, nn.Linear) and "mlp" in cls.lower():
            return plain("fc", module.fc1.in_features, module.fc1.out_features)
        if hasattr(module, "dense_h_to_4h") and hasattr(module, "dense_4h_to_h"):
            return plain("megatron", module.dense_h_to_4h.in_features, module.dense_h_to_4h.out_features, default="gelu")
        return None

    @staticmethod
    def _copy(dst: nn.Linear, weight: torch.Tensor, bias: Optional[torch.Tensor]):
        with torch.no_grad():
            dst.weight.copy_(weight.to(dst.weight.dtype))
            if bias is None:
                dst.bias.zero_()
            else:
                dst.bias.copy_(bias.to(dst.bias.dtype))

    def _create_fused_mlp(self, module: nn.Module, info: Dict[str, Any]) -> nn.Module:
        cfg = FusedMLPConfig(**{**self.config.__dict__})
        fused = FusedTransformerMLP(info["hidden"], info["intermediate"], info["activation"], cfg)
        ref = next(module.parameters())
        kind = info["kind"]
        if kind == "gpt2":
            self._copy(fused.mlp.fc1, module.c_fc.weight.t(), module.c_fc.bias)
            self._copy(fused.mlp.fc2, module.c_proj.weight.t(), module.c_proj.bias)
        elif kind == "llama":
            self._copy(fused.mlp.fc1_gate, module.gate_proj.weight, module.gate_proj.bias)
            self._copy(fused.mlp.fc1, module.up_proj.weight, module.up_proj.bias)
            self._copy(fused.mlp.fc2, module.down_proj.weight, module.down_proj.bias)
        elif kind == "fc":
            self._copy(fused.mlp.fc1, module.fc1.weight, module.fc1.bias)
            self._copy(fused.mlp.fc2, module.fc2.weight, module.fc2.bias)
        else:
            self._copy(fused.mlp.fc1, module.dense_h_to_4h.weight, module.dense_h_to_4h.bias)
            self._copy(fused.mlp.fc2, module.dense_4h_to_h.weight, module.dense_4h_to_h.bias)
        fused.to(device=ref.device, dtype=ref.dtype)
        return _MLPAdapter(fused)

    def convert_model(self, model: nn.Module, target_class_names: Optional[List[str]] = None) -> nn.Module:
        """reference :560-614."""
        for name, sub in list(model.named_children()):
            info = None
            if target_class_names is None or type(sub).__name__ in target_class_names:
                info = self._detect_mlp_type(sub)
            if info is not None:
                setattr(model, name, self._create_fused_mlp(sub, info))
            else:
                self.convert_model(sub, target_class_names)
        return model
