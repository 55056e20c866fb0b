cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --launch-timeout 0 --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "layernorm_vs_oracle or ring_variant or (pair_kernel_vs_oracle and 640)" > gpurun_out/r2_sanitizer.log 2>&1; echo rc=$?
grep -E "ERROR SUMMARY|passed|failed|Invalid|error" gpurun_out/r2_sanitizer.log | head -12
