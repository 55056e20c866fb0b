#!/bin/bash
# dev tool: cycle-count A/B (sm__cycles_elapsed.max under ncu is independent of the power-capped clock)
set -e
VARIANTS="${VARIANTS:-old new}"
rm -f gpurun_out/fa_ab_plain.txt
for v in $VARIANTS; do timeout 200 python tests/fa_ab_ncu.py $v >> gpurun_out/fa_ab_plain.txt 2>&1 || { tail -5 gpurun_out/fa_ab_plain.txt; exit 1; }; done
timeout 600 ncu --target-processes all --metrics gpu__time_duration.sum,sm__cycles_elapsed.max --clock-control none --csv --log-file gpurun_out/fa_ab_ncu.csv bash -c "for v in $VARIANTS; do python tests/fa_ab_ncu.py \$v; done" > gpurun_out/fa_ab_ncu.log 2>&1
python tests/fa_ab_parse.py $VARIANTS
