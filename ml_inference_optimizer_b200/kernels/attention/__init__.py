from .flash_attention import (FlashAttention3, FlashAttentionConfig, FlashAttentionLayer, FlashSelfAttention,  # noqa: F401
                              ModelConverter)
