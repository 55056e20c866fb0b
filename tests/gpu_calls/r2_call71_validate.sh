#!/bin/bash
# full validation of the final tree: build from source on the box, every GPU test, smoke, the default bench line and the reference arm
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_final_pytest.log 2>&1
tail -6 gpurun_out/r2_final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" > gpurun_out/r2_final_smoke.log 2>&1; tail -2 gpurun_out/r2_final_smoke.log
timeout 900 python bench.py > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err; echo bench rc=$?
tail -c 400 gpurun_out/r2_final_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_final_bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','clocks')})
print({k:(v.get('tflops') or v.get('gbs'), v.get('speedup_vs_unfused'), v.get('cudnn_sdpa_tflops')) for k,v in d['secondary'].items()})
PY
