cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "mlp or linear or golden or fullsize" 2>&1 | tail -3
timeout 300 python tests/gemm_epilogue_probe.py 2>&1 | tee gpurun_out/r2_gemm_epilogue_probe_after.jsonl
