#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_modules.py -q -m gpu -x -k "kv_cache_transformer_runner or paged_generation or add_paged" > gpurun_out/c66_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/c66_tests.log
tail -40 gpurun_out/c66_tests.log
