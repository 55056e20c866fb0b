set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_pytest5.log 2>&1
tail -6 gpurun_out/r2_pytest5.log
timeout 900 python bench.py --steps 50 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo rc=$?
tail -c 400 gpurun_out/r2_bench_n1.err
timeout 200 python tests/decode_probe.py > gpurun_out/r2_decode_probe_final.jsonl 2>&1; cat gpurun_out/r2_decode_probe_final.jsonl
timeout 300 python tests/ln_probe.py > gpurun_out/r2_ln_probe3.log 2>&1; grep "narrow_max=dflt wide_ctas=occ" gpurun_out/r2_ln_probe3.log
# launch list of the same command (shares, not absolutes)
python bench.py --steps 5 --warmup 3 --no-secondary --no-ring --no-cpu-baseline > gpurun_out/ncu_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_bench_launches_ncu.csv python bench.py --steps 5 --warmup 3 --no-secondary --no-ring --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
# full captures of the dominant kernels
python tests/ncu_target.py mlp 3 > gpurun_out/ncu_plain_mlp.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_act_pair -s 2 -c 4 -o gpurun_out/r2f_gemm_pair python tests/ncu_target.py mlp 3 > gpurun_out/ncu_gemm_pair.log 2>&1
python tests/ncu_target.py fa 2 > gpurun_out/ncu_plain_fa.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fa_fwd -s 1 -c 1 -o gpurun_out/r2f_fa python tests/ncu_target.py fa 2 > gpurun_out/ncu_fa.log 2>&1
python tests/ncu_target.py decode 2 > gpurun_out/ncu_plain_dec.log 2>&1 && ncu --set full --clock-control none -k regex:decode_kernel -s 1 -c 1 -o gpurun_out/r2f_decode python tests/ncu_target.py decode 2 > gpurun_out/ncu_dec.log 2>&1
python tests/ncu_target.py decode_gqa 2 > gpurun_out/ncu_plain_decg.log 2>&1 && ncu --set full --clock-control none -k regex:decode_gqa -s 1 -c 1 -o gpurun_out/r2f_decode_gqa python tests/ncu_target.py decode_gqa 2 > gpurun_out/ncu_decg.log 2>&1
python tests/ncu_target.py ln 2 > gpurun_out/ncu_plain_ln.log 2>&1 && ncu --set full --clock-control none -k regex:layernorm -s 1 -c 1 -o gpurun_out/r2f_layernorm python tests/ncu_target.py ln 2 > gpurun_out/ncu_ln.log 2>&1
ls -la gpurun_out/*.ncu-rep
