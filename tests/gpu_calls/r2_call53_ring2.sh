cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B200_GQA_RING=1 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_modules.py -m gpu -x -q -k "decode or paged or generate or write_only" 2>&1 | tail -3
for r in 0 1 0 1; do B200_GQA_RING=$r TAG=ring$r timeout 120 python tests/decode_probe.py | grep -v mha; done 2>&1 | tee gpurun_out/r2_decode_ring_probe.jsonl
