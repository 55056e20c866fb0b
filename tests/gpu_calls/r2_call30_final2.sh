set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_pytest6.log 2>&1
tail -5 gpurun_out/r2_pytest6.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" 2>&1 | tail -2
timeout 900 python bench.py --steps 50 > gpurun_out/r2_bench_n1b.json 2> gpurun_out/r2_bench_n1b.err; echo rc=$?
tail -c 300 gpurun_out/r2_bench_n1b.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n1b.json').read().strip().splitlines()[-1])
print('value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'e2e ms',round(d['e2e']['ms_per_step'],2),'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],3), d['roofline']['traffic'], d['roofline']['traffic_source'], 'fa',round(d['roofline_attention']['achieved'],1))
print({k:(round(v.get('gbs',0),0) or round(v.get('tflops',0),1), v.get('speedup_vs_unfused')) for k,v in d['secondary'].items()})
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -c 700
