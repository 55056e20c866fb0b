cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out

timeout 300 python tests/decode_split_probe.py 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); r=d['us_by_splits(0=auto)']; print(d['case'],d['ctas_per_split'],r)" | tee gpurun_out/r2_decode_splits_after.txt

