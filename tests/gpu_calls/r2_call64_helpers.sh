#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_modules.py -q -m gpu -x -k "measurement_helpers or ring_module" > gpurun_out/c64_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/c64_tests.log
timeout 300 python - > gpurun_out/c64_helpers.jsonl 2> gpurun_out/c64_helpers.err <<'PY'
import json, sys
sys.path.insert(0, ".")
from kernels.triton import flash_attention_kernels as fk, mlp_kernels as mk, layernorm_kernels as ln, fused_layernorm_qkv as lq, attention_kernels as ak
from kernels.attention import ring_attention as ra
print(json.dumps({"benchmark_flash_attention": fk.benchmark_flash_attention(4096, 8, 12, 64, causal=True, iterations=20, warmup=3)}))
print(json.dumps({"benchmark_fused_mlp": mk.benchmark_fused_mlp(8, 4096, 768, 3072, "gelu", num_warmup=3, num_iter=20)}))
print(json.dumps({"benchmark_layernorm": ln.benchmark_layernorm(8, 4096, 4096, iterations=20, warmup=3)}))
print(json.dumps({"benchmark_fused_layernorm_qkv": lq.benchmark_fused_layernorm_qkv(8, 4096, 768, 12, iterations=20, warmup=3)}))
print(json.dumps({"compare_with_flash_attention": ak.compare_with_flash_attention(4096, 2, 4096, 32)}))
print(json.dumps({"ring.compare_with_standard_attention": ra.compare_with_standard_attention(4096, 1, 1024, 16)}))
PY
tail -15 gpurun_out/c64_tests.log; cat gpurun_out/c64_helpers.jsonl; tail -5 gpurun_out/c64_helpers.err
