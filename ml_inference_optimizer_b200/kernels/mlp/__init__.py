from .fused_mlp import (FusedMLP, FusedMLPConfig, FusedMLPGeluTanh, FusedMLPReLU, FusedMLPSwiGLU,  # noqa: F401
                        FusedTransformerMLP, MLPConverter)
