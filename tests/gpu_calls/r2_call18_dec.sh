cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "decode or layernorm" > gpurun_out/r2_pytest_dec.log 2>&1; tail -4 gpurun_out/r2_pytest_dec.log
for l in old dec2 old dec2; do B200_LIB_PATH=gpurun_in/$l.so timeout 120 python tests/decode_probe.py; done 2>&1 | tee gpurun_out/r2_decode_probe.jsonl
