cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --no-secondary --no-ring --no-cpu-baseline > gpurun_out/ncu_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_bench_launches_ncu.csv python bench.py --steps 5 --warmup 3 --no-secondary --no-ring --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python tests/ncu_target.py mlp 3 > gpurun_out/ncu_plain_mlp.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_act_pair -s 2 -c 4 -o gpurun_out/r2f_gemm_pair python tests/ncu_target.py mlp 3 > gpurun_out/ncu_gemm_pair.log 2>&1
python tests/ncu_target.py mlp_c2 2 > gpurun_out/ncu_plain_c2.log 2>&1 && ncu --set full --clock-control none -k regex:gemm_act_pair -s 2 -c 2 -o gpurun_out/r2f_gemm_pair_c2 python tests/ncu_target.py mlp_c2 2 > gpurun_out/ncu_c2.log 2>&1
python tests/ncu_target.py decode_gqa 2 > gpurun_out/ncu_plain_decg.log 2>&1 && ncu --set full --clock-control none -k regex:decode_gqa -s 1 -c 1 -o gpurun_out/r2f_decode_gqa python tests/ncu_target.py decode_gqa 2 > gpurun_out/ncu_decg.log 2>&1
ls -la gpurun_out/r2f_*.ncu-rep gpurun_out/r2f_bench_launches_ncu.csv
