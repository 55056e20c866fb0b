#!/bin/bash
mkdir -p gpurun_out
timeout 400 python bench.py --steps 10 --warmup 3 --no-secondary --no-ring --no-cpu-baseline > gpurun_out/c77_bench.json 2> gpurun_out/c77_bench.err; echo rc=$?
tail -c 300 gpurun_out/c77_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c77_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','prefill_tokens_per_s','pct_of_bf16_peak')})
PY
