#!/bin/bash
# dev tool: same-box A/B of two builds of the library (gpurun_in/old.so vs gpurun_in/new.so), interleaved
L=ml_inference_optimizer_b200/libb200_attn_mlp.so
for rep in 1 2; do
  for v in old new; do
    cp gpurun_in/$v.so $L
    echo "== $v (rep $rep)"; timeout 100 python tests/attn_quick.py std 2>&1 | tail -3
  done
done
cp gpurun_in/new.so $L
echo "== new, B200_FA_PERSISTENT=0"; B200_FA_PERSISTENT=0 timeout 100 python tests/attn_quick.py std 2>&1 | tail -3
