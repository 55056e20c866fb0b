#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tests/skinny_gemm_target.py > gpurun_out/c62_plain.log 2>&1 || { tail -5 gpurun_out/c62_plain.log; exit 1; }
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,sm__cycles_elapsed.max --clock-control none --csv --log-file gpurun_out/c62_skinny_launches.csv python tests/skinny_gemm_target.py > gpurun_out/c62_ncu.log 2>&1
tail -3 gpurun_out/c62_ncu.log
