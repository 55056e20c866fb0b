cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "decode or paged or generate or abi" 2>&1 | tail -3
timeout 300 python tests/decode_split_probe.py 2>&1 | tee gpurun_out/r2_decode_splits.jsonl | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); r=d['us_by_splits(0=auto)']; print(d['case'],d['ctas_per_split'],'auto',r['0'],'best',d['best'])"
timeout 120 python tests/decode_probe.py | tee gpurun_out/r2_decode_probe_final2.jsonl
