#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "stream_k" > gpurun_out/c67_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/c67_tests.log
tail -30 gpurun_out/c67_tests.log
if grep -q "tests rc=0" gpurun_out/c67_tests.log; then
  timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_modules.py -q -m gpu -x -k "mlp or linear or graph or paged_generation or optimizer" > gpurun_out/c67_tests2.log 2>&1
  echo "tests2 rc=$?" >> gpurun_out/c67_tests2.log
  tail -5 gpurun_out/c67_tests2.log
  timeout 300 python tests/streamk_probe.py > gpurun_out/c67_streamk.jsonl 2> gpurun_out/c67_streamk.err
  cat gpurun_out/c67_streamk.jsonl; tail -3 gpurun_out/c67_streamk.err
fi
