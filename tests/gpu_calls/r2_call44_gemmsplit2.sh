cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "mlp or linear or abi" 2>&1 | tail -3
timeout 600 python tests/gemm_split_probe.py 2>&1 | tee gpurun_out/r2_gemm_splits_after.jsonl | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); r=d['us_by_splits(0=model)']; print(d['case'],'model',r['0'],'best',d['best'],'cublas',d['cublas_us'])"
