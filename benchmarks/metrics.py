"""Metric definitions of the reference's ``benchmarks/metrics.py`` (same names, formulas and units), so results of this
repo drop into the reference's report / dashboard JSON. Pure Python / torch-on-CPU helpers."""
from __future__ import annotations

import statistics
from typing import Dict, List

import torch


def calculate_throughput(batch_size: int, seq_len: int, time_seconds: float) -> float:
    """samples per second (reference :15-27)."""
    return batch_size / time_seconds if time_seconds > 0 else 0


def calculate_latency_statistics(latencies: List[float]) -> Dict[str, float]:
    """mean / median / min / max / p50 / p90 / p95 / p99 / stddev (reference :30-79)."""
    keys = ("mean", "median", "min", "max", "p50", "p90", "p95", "p99", "stddev")
    if not latencies:
        return {k: 0.0 for k in keys}
    s = sorted(latencies)
    n = len(s)
    median = statistics.median(s)
    return {"mean": statistics.mean(s), "median": median, "min": s[0], "max": s[-1], "p50": median,
            "p90": s[int(n * 0.9)], "p95": s[int(n * 0.95)], "p99": s[int(n * 0.99)],
            "stddev": statistics.stdev(s) if n > 1 else 0.0}


def calculate_memory_reduction(baseline_memory: float, optimized_memory: float) -> float:
    """percent reduction, clamped at 0 (reference :150-169)."""
    if baseline_memory <= 0:
        return 0.0
    return max(0.0, (baseline_memory - optimized_memory) / baseline_memory * 100)


def calculate_scaling_efficiency(single_gpu_time: float, multi_gpu_time: float, num_gpus: int) -> float:
    """(T1 / TN) / N in percent (reference :172-190)."""
    if multi_gpu_time <= 0 or num_gpus <= 0:
        return 0.0
    return (single_gpu_time / multi_gpu_time) / num_gpus * 100


def calculate_communication_overhead(total_time: float, computation_time: float) -> float:
    """percent of the total time not spent computing (reference :193-208)."""
    if total_time <= 0:
        return 0.0
    return max(0.0, total_time - computation_time) / total_time * 100


def calculate_relative_error(baseline: torch.Tensor, optimized: torch.Tensor) -> float:
    """mean(|a-b| / (|a| + 1e-8)) * 100 (reference :211-238)."""
    if baseline.shape != optimized.shape:
        return float("inf")
    a, b = baseline.detach().double().cpu(), optimized.detach().double().cpu()
    return float(((a - b).abs() / (a.abs() + 1e-8)).mean() * 100)


def calculate_max_absolute_error(baseline: torch.Tensor, optimized: torch.Tensor) -> float:
    """max |a-b| (reference :241-262)."""
    if baseline.shape != optimized.shape:
        return float("inf")
    return float((baseline.detach().double().cpu() - optimized.detach().double().cpu()).abs().max())
