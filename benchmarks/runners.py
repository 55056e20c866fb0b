"""Benchmark runners with the reference's protocol and result keys (reference ``benchmarks/runners.py``: ``BenchmarkConfig``
:28-50, ``BenchmarkRunner`` :53-330, ``ThroughputBenchmarkRunner`` :333-360, ``LatencyBenchmarkRunner`` :363-405,
``ScalingBenchmarkRunner`` :450-526), so results of this repo drop into the reference's report / dashboard JSON
(SURVEY.md §8 f4).

What differs, on purpose:
  * latencies are device times (CUDA events on the current stream, one pair per iteration, read after a single
    synchronize) instead of ``time.time()`` around a synchronize — the reference's number includes a host round trip
    per iteration;
  * ``ScalingBenchmarkRunner`` really measures the multi-rank leg when ``torch.distributed`` is initialised (one process
    per GPU; max over ranks) — the reference leaves ``results["multi_gpu"]`` empty (:499-506);
  * ``ModelBenchmarkRunner`` supplies the ``setup_model_variants`` / ``generate_test_inputs`` the reference leaves
    abstract: "baseline" = the model as given, "optimized" = ``Optimizer(model).optimize(...)`` (K1 + K3 swapped in).
"""
from __future__ import annotations

import copy
import json
import os
import time
from dataclasses import asdict, dataclass
from typing import Any, Callable, Dict, List, Optional

import torch
import torch.nn as nn

from benchmarks.metrics import calculate_latency_statistics, calculate_scaling_efficiency, calculate_throughput


@dataclass
class BenchmarkConfig:
    """Same fields and defaults as the reference dataclass (:28-50)."""
    model_name: str
    batch_sizes: List[int]
    sequence_lengths: List[int]
    optimization_types: List[str]
    num_iterations: int = 100
    warmup_iterations: int = 10
    devices: Optional[List[str]] = None
    precision: str = "fp16"
    save_results: bool = True
    profiling: bool = False
    validate_outputs: bool = True

    def __post_init__(self):
        if self.devices is None:
            self.devices = ["cuda:0"]

    def to_dict(self) -> Dict[str, Any]:
        return asdict(self)


def _first_tensor(inputs: Dict[str, torch.Tensor]) -> torch.Tensor:
    return inputs.get("input_ids", next(iter(inputs.values())))


def _tensors_of(out: Any) -> List[torch.Tensor]:
    if isinstance(out, torch.Tensor):
        return [out]
    if isinstance(out, dict) or hasattr(out, "keys"):
        return [v for v in (out[k] for k in out.keys()) if isinstance(v, torch.Tensor)]
    if isinstance(out, (tuple, list)):
        return [v for v in out if isinstance(v, torch.Tensor)]
    return []


class BenchmarkRunner:
    """Base runner (reference :53-330)."""

    def __init__(self, config: BenchmarkConfig, results_dir: str = "benchmark_results"):
        self.config = config
        self.results_dir = results_dir
        if config.save_results:
            os.makedirs(results_dir, exist_ok=True)
        self.primary_device = torch.device(config.devices[0])
        self.dtype = self._get_dtype_from_precision(config.precision)

    def _get_dtype_from_precision(self, precision: str) -> torch.dtype:
        table = {"fp32": torch.float32, "fp16": torch.float16, "bf16": torch.bfloat16}
        if precision not in table:
            raise ValueError(f"Unsupported precision: {precision}")  # reference :96
        return table[precision]

    # -- to be provided by subclasses (reference :160-183) --
    def setup_model_variants(self) -> Dict[str, nn.Module]:
        raise NotImplementedError("Subclasses must implement setup_model_variants()")

    def generate_test_inputs(self, batch_size: int, seq_len: int) -> Dict[str, torch.Tensor]:
        raise NotImplementedError("Subclasses must implement generate_test_inputs()")

    def _warm(self, model: nn.Module, inputs: Dict[str, torch.Tensor]) -> None:
        with torch.no_grad():
            for _ in range(self.config.warmup_iterations):
                model(**inputs)

    def measure_performance(self, model: nn.Module, inputs: Dict[str, torch.Tensor]) -> Dict[str, float]:
        """Result keys of the reference (:185-248); latencies are device times."""
        model.eval()
        first = _first_tensor(inputs)
        batch_size, seq_len = first.shape[0], first.shape[1]
        dev = first.device
        if dev.type != "cuda":
            raise RuntimeError("measure_performance times on the device with CUDA events; inputs must be CUDA tensors")
        n = self.config.num_iterations
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
        torch.cuda.synchronize(dev)
        torch.cuda.reset_peak_memory_stats(dev)
        mem_before = torch.cuda.memory_allocated(dev)
        with torch.no_grad():
            for i in range(n):
                starts[i].record()
                model(**inputs)
                ends[i].record()
        torch.cuda.synchronize(dev)
        latencies = [s.elapsed_time(e) / 1e3 for s, e in zip(starts, ends)]
        peak_extra_mb = (torch.cuda.max_memory_allocated(dev) - mem_before) / 2**20
        avg = sum(latencies) / len(latencies)
        stats = calculate_latency_statistics(latencies)
        return {
            "avg_latency_ms": avg * 1e3,
            "throughput_samples_per_sec": calculate_throughput(batch_size, seq_len, avg),
            "latency_p50_ms": stats["p50"] * 1e3,
            "latency_p90_ms": stats["p90"] * 1e3,
            "latency_p95_ms": stats["p95"] * 1e3,
            "latency_p99_ms": stats["p99"] * 1e3,
            "memory_usage_mb": peak_extra_mb,
            "batch_size": batch_size,
            "sequence_length": seq_len,
        }

    def validate_model_outputs(self, baseline_outputs: Any, optimized_outputs: Any, rtol: float = 1e-3,
                               atol: float = 1e-3) -> bool:
        """``allclose`` over tensors / tuples / dicts, False for anything else (reference :250-297)."""
        a, b = _tensors_of(baseline_outputs), _tensors_of(optimized_outputs)
        if not a or len(a) != len(b):
            return False
        return all(x.shape == y.shape and torch.allclose(x.float(), y.float(), rtol=rtol, atol=atol) for x, y in zip(a, b))

    def save_benchmark_results(self, results: Dict[str, Any], filename: str) -> str:
        def plain(obj):
            if isinstance(obj, torch.Tensor):
                return obj.tolist()
            if isinstance(obj, dict):
                return {k: plain(v) for k, v in obj.items()}
            if isinstance(obj, (list, tuple)):
                return [plain(v) for v in obj]
            return obj

        path = os.path.join(self.results_dir, filename)
        os.makedirs(self.results_dir, exist_ok=True)
        with open(path, "w") as f:
            json.dump(plain(results), f, indent=2)
        return path

    def _bench_variants(self, variants: Dict[str, nn.Module], key_fmt: Callable[[str, int, int], str],
                        flat: bool) -> Dict[str, Any]:
        out: Dict[str, Any] = {}
        for bs in self.config.batch_sizes:
            for sl in self.config.sequence_lengths:
                inputs = {k: v.to(self.primary_device) for k, v in self.generate_test_inputs(bs, sl).items()}
                base_out = None
                if self.config.validate_outputs and "baseline" in variants:
                    with torch.no_grad():
                        base_out = variants["baseline"].to(self.primary_device)(**inputs)
                per_cfg: Dict[str, Any] = {}
                for name, model in variants.items():
                    model = model.to(self.primary_device)
                    self._warm(model, inputs)
                    perf = self.measure_performance(model, inputs)
                    if base_out is not None and name != "baseline":
                        with torch.no_grad():
                            perf["output_validation"] = self.validate_model_outputs(base_out, model(**inputs))
                    if flat:
                        out[key_fmt(name, bs, sl)] = perf
                    else:
                        per_cfg[name] = perf
                if not flat:
                    out[f"bs{bs}_seq{sl}"] = per_cfg
        return out

    def run_benchmarks(self) -> Dict[str, Any]:
        """Every (batch, seq) x variant; same result tree as the reference (:98-158)."""
        results = {"config": self.config.to_dict(), "timestamp": time.time(), "benchmarks": {}}
        results["benchmarks"] = self._bench_variants(self.setup_model_variants(), lambda n, b, s: n, flat=False)
        if self.config.save_results:
            self.save_benchmark_results(results, f"{int(time.time())}_{self.config.model_name}.json")
        return results


class ThroughputBenchmarkRunner(BenchmarkRunner):
    """Adds ``tokens_per_second`` (reference :333-360)."""

    def measure_performance(self, model, inputs):
        r = super().measure_performance(model, inputs)
        r["tokens_per_second"] = r["batch_size"] * r["sequence_length"] / (r["avg_latency_ms"] / 1e3)
        return r


class LatencyBenchmarkRunner(BenchmarkRunner):
    """Adds first-call latency and jitter (reference :363-405)."""

    def measure_performance(self, model, inputs):
        dev = _first_tensor(inputs).device
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        with torch.no_grad():
            s.record()
            model(**inputs)
            e.record()
        torch.cuda.synchronize(dev)
        r = super().measure_performance(model, inputs)
        r["first_token_latency_ms"] = s.elapsed_time(e)
        r["latency_jitter_ms"] = r["latency_p99_ms"] - r["latency_p50_ms"]
        return r


class ScalingBenchmarkRunner(BenchmarkRunner):
    """Single-GPU numbers plus, under ``torch.distributed`` (one process per GPU), the multi-rank leg and
    ``scaling_efficiency`` = (multi throughput / single throughput) / num_gpus (reference :450-526).

    ``setup_parallel_variants(world_size)`` returns the models to run on every rank (tensor-/sequence-parallel shells
    built from ``parallelism/``); the default is the single-GPU variants, i.e. data-parallel replicas over a
    ``world_size``-times larger global batch."""

    def setup_parallel_variants(self, world_size: int) -> Dict[str, nn.Module]:
        return self.setup_model_variants()

    def run_benchmarks(self) -> Dict[str, Any]:
        import torch.distributed as dist

        results = {"config": self.config.to_dict(), "timestamp": time.time(), "benchmarks": {}}
        key = lambda n, b, s: f"{n}_bs{b}_seq{s}"
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        single = self._bench_variants(self.setup_model_variants(), key, flat=True)
        results["single_gpu"] = single
        if world > 1:
            multi = self._bench_variants(self.setup_parallel_variants(world), key, flat=True)
            for k, m in multi.items():  # job time = slowest rank; replicas process world x the batch
                t = torch.tensor([m["avg_latency_ms"]], device=self.primary_device)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                m["avg_latency_ms"] = t.item()
                m["throughput_samples_per_sec"] = calculate_throughput(m["batch_size"] * world, m["sequence_length"],
                                                                       t.item() / 1e3)
            results["multi_gpu"] = multi
            results["scaling_efficiency"] = {
                k: {"speedup": multi[k]["throughput_samples_per_sec"] / s["throughput_samples_per_sec"],
                    "scaling_efficiency": multi[k]["throughput_samples_per_sec"] / s["throughput_samples_per_sec"] / world,
                    "num_gpus": world}
                for k, s in single.items() if k in multi}
        if self.config.save_results and (world == 1 or dist.get_rank() == 0):
            self.save_benchmark_results(results, f"{int(time.time())}_{self.config.model_name}_scaling.json")
        return results


class ModelBenchmarkRunner(ThroughputBenchmarkRunner):
    """Concrete runner: ``model_factory()`` builds the (HF-style) model once; variants are taken from
    ``config.optimization_types`` — "baseline", "flash_attention", "fused_mlp", "optimized" (= both)."""

    def __init__(self, config: BenchmarkConfig, model_factory: Callable[[], nn.Module], vocab_size: int = 50257,
                 results_dir: str = "benchmark_results"):
        super().__init__(config, results_dir)
        self.model_factory = model_factory
        self.vocab_size = vocab_size

    def setup_model_variants(self) -> Dict[str, nn.Module]:
        from ml_inference_optimizer_b200.optimizer import Optimizer

        base = self.model_factory().eval().to(self.primary_device, self.dtype)
        flags = {"flash_attention": (True, False), "fused_mlp": (False, True), "optimized": (True, True)}
        variants: Dict[str, nn.Module] = {}
        for name in self.config.optimization_types:
            if name == "baseline":
                variants[name] = base
            elif name in flags:
                fa, mlp = flags[name]
                variants[name] = Optimizer(copy.deepcopy(base)).optimize(use_flash_attention=fa, use_fused_mlp=mlp)
            else:
                raise ValueError(f"unknown optimization type {name!r}; expected baseline / {' / '.join(flags)}")
        return variants

    def generate_test_inputs(self, batch_size: int, seq_len: int) -> Dict[str, torch.Tensor]:
        g = torch.Generator().manual_seed(1000 * batch_size + seq_len)
        return {"input_ids": torch.randint(0, self.vocab_size, (batch_size, seq_len), generator=g)}
