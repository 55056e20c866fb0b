#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "stream_k" > gpurun_out/c68_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/c68_tests.log
tail -30 gpurun_out/c68_tests.log
if grep -q "tests rc=0" gpurun_out/c68_tests.log; then
  timeout 500 python tests/streamk_probe.py > gpurun_out/c68_streamk.jsonl 2> gpurun_out/c68_streamk.err
  cat gpurun_out/c68_streamk.jsonl; tail -3 gpurun_out/c68_streamk.err
fi
