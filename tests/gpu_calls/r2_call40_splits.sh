cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tests/decode_split_probe.py 2>&1 | tee gpurun_out/r2_decode_splits.jsonl
