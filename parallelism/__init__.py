"""Drop-in import path: ``from parallelism.sequence_parallel import SequenceParallelAttention`` etc. resolve to
``ml_inference_optimizer_b200.parallelism``."""
import importlib
import sys

_IMPL = "ml_inference_optimizer_b200.parallelism"
for _sub in ("communication", "parallel_utils", "ring", "sequence_parallel", "tensor_parallel"):
    _mod = importlib.import_module(f"{_IMPL}.{_sub}")
    sys.modules[f"{__name__}.{_sub}"] = _mod
    globals()[_sub] = _mod
from ml_inference_optimizer_b200.parallelism import *  # noqa: F401,F403,E402
