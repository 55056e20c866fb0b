#!/bin/bash
# arbitrary head_dim in K1: parity tests + every attention test + a probe next to cuDNN
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_modules.py -q -m gpu -x -k "fa_ or attention or ring or kernels_write" > gpurun_out/c61_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/c61_tests.log
timeout 300 python tests/head_dim_probe.py > gpurun_out/c61_head_dim.jsonl 2> gpurun_out/c61_head_dim.err
tail -5 gpurun_out/c61_tests.log; cat gpurun_out/c61_head_dim.jsonl
