#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_modules.py -q -m gpu -x > gpurun_out/c78_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/c78_tests.log
tail -12 gpurun_out/c78_tests.log
