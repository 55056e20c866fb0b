cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tests/gemm_split_probe.py 2>&1 | tee gpurun_out/r2_gemm_splits.jsonl
