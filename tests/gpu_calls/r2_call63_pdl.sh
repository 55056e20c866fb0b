#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_modules.py -q -m gpu -x -k "mlp or linear or graph or paged_generation or optimizer" > gpurun_out/c63_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/c63_tests.log
timeout 400 python tests/pdl_probe.py > gpurun_out/c63_pdl.jsonl 2> gpurun_out/c63_pdl.err
tail -4 gpurun_out/c63_tests.log; cat gpurun_out/c63_pdl.jsonl; tail -3 gpurun_out/c63_pdl.err
