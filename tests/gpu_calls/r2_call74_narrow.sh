#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_modules.py -q -m gpu -x -k "narrow_heads or head_dim_96 or paged or decode or kv_cache or generation" > gpurun_out/c74_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/c74_tests.log
tail -30 gpurun_out/c74_tests.log
