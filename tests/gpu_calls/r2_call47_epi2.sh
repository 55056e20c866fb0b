cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out

timeout 300 python tests/gemm_epilogue_probe.py 2>&1 | tee gpurun_out/r2_gemm_epilogue_probe_after.jsonl
