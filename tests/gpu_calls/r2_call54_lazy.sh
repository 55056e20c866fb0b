cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_modules.py tests/test_gpu_fullsize.py -m gpu -x -q -k "decode or paged or generate or write_only" 2>&1 | tail -3
for l in base lazy base lazy; do B200_LIB_PATH=gpurun_in/$l.so timeout 120 python tests/decode_probe.py | grep -v mha; done 2>&1 | tee gpurun_out/r2_decode_lazy_probe.jsonl
