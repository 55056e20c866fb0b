#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tests/skinny_gemm_target.py > gpurun_out/c70_plain.log 2>&1 || { tail -5 gpurun_out/c70_plain.log; exit 1; }
timeout 500 ncu --set full --clock-control none --import-source on -k regex:gemm_act_kernel -s 2 -c 2 -o gpurun_out/r2g_skinny_up_gate python tests/skinny_gemm_target.py > gpurun_out/c70_ncu.log 2>&1
tail -3 gpurun_out/c70_ncu.log; ls -la gpurun_out/r2g_skinny_up_gate.ncu-rep
