"""Mirror of the slice of the reference's ``baseline`` package that drives the hot path end to end: the paged KV cache
bookkeeping (``BlockManager`` / ``SequenceMetadata`` / ``PagedKVCache``), a working inference runner and a greedy
generation loop over the paged cache."""
from .inference import (BasicInferenceRunner, BlockManager, InferenceRunner, PagedKVCache, SequenceMetadata,  # noqa: F401
                        create_inference_runner, generate_paged)
