"""Drop-in import path: ``from baseline.inference import PagedKVCache, create_inference_runner`` resolves to
``ml_inference_optimizer_b200.baseline``. (``baseline/_ref`` is a git-ignored scratch location.)"""
import importlib
import sys

_mod = importlib.import_module("ml_inference_optimizer_b200.baseline.inference")
sys.modules[f"{__name__}.inference"] = _mod
inference = _mod
_mu = importlib.import_module("ml_inference_optimizer_b200.baseline.model_utils")
sys.modules[f"{__name__}.model_utils"] = _mu
model_utils = _mu
_ml = importlib.import_module("ml_inference_optimizer_b200.baseline.model_loader")
sys.modules[f"{__name__}.model_loader"] = _ml
model_loader = _ml
from ml_inference_optimizer_b200.baseline import *  # noqa: F401,F403,E402
