"""``from ml_inference_optimizer import Optimizer`` — the package name the reference's README uses (README.md:57-58)."""
from ml_inference_optimizer_b200.optimizer import Optimizer  # noqa: F401
