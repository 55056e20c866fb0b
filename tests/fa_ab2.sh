#!/bin/bash
L=ml_inference_optimizer_b200/libb200_attn_mlp.so
cp gpurun_in/old.so $L; echo "== old"; timeout 100 python tests/fa_overhead_probe.py 128 2>&1 | tail -7
cp gpurun_in/new.so $L; echo "== new persistent"; timeout 100 python tests/fa_overhead_probe.py 128 2>&1 | tail -7
echo "== new non-persistent"; B200_FA_PERSISTENT=0 timeout 100 python tests/fa_overhead_probe.py 128 2>&1 | tail -7
