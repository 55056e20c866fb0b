#!/bin/bash
mkdir -p gpurun_out
timeout 500 python tests/streamk_probe.py > gpurun_out/c69_streamk.jsonl 2> gpurun_out/c69_streamk.err
cat gpurun_out/c69_streamk.jsonl; tail -3 gpurun_out/c69_streamk.err
