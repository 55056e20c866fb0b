#!/bin/bash
mkdir -p gpurun_out
timeout 500 python tests/packed_weight_probe.py > gpurun_out/c73_packed.jsonl 2> gpurun_out/c73_packed.err
cat gpurun_out/c73_packed.jsonl; tail -5 gpurun_out/c73_packed.err
