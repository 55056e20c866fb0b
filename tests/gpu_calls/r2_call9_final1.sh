set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_pytest3.log 2>&1
tail -6 gpurun_out/r2_pytest3.log
python bench.py --steps 50 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo rc=$?
tail -c 600 gpurun_out/r2_bench_n1.err
# launch list of the same command (shares, not absolutes)
python bench.py --steps 5 --warmup 3 --no-secondary --no-ring --no-cpu-baseline > gpurun_out/ncu_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches_ncu.csv python bench.py --steps 5 --warmup 3 --no-secondary --no-ring --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
# full captures of the dominant kernels
python tests/ncu_target.py mlp 3 > gpurun_out/ncu_plain_mlp.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_act_pair -s 2 -c 4 -o gpurun_out/prof_gemm_pair_r2 python tests/ncu_target.py mlp 3 > gpurun_out/ncu_gemm_pair.log 2>&1
python tests/ncu_target.py fa 2 > gpurun_out/ncu_plain_fa.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fa_fwd -s 1 -c 1 -o gpurun_out/prof_fa_r2 python tests/ncu_target.py fa 2 > gpurun_out/ncu_fa.log 2>&1
B200_MLP_FUSED=1 python tests/ncu_target.py mlp_c2 2 > gpurun_out/ncu_plain_c2f.log 2>&1 && B200_MLP_FUSED=1 ncu --set full --clock-control none --import-source on -k regex:fused_mlp -s 1 -c 1 -o gpurun_out/prof_fusedmlp_c2_r2 python tests/ncu_target.py mlp_c2 2 > gpurun_out/ncu_c2f.log 2>&1
B200_MLP_FUSED=0 python tests/ncu_target.py mlp_c2 2 > gpurun_out/ncu_plain_c2t.log 2>&1 && B200_MLP_FUSED=0 ncu --set full --clock-control none --import-source on -k regex:gemm_act_pair -s 2 -c 2 -o gpurun_out/prof_mlp_c2_two_r2 python tests/ncu_target.py mlp_c2 2 > gpurun_out/ncu_c2t.log 2>&1
python tests/ncu_target.py decode 2 > gpurun_out/ncu_plain_dec.log 2>&1 && ncu --set full --clock-control none -k regex:decode_kernel -s 1 -c 1 -o gpurun_out/prof_decode_r2 python tests/ncu_target.py decode 2 > gpurun_out/ncu_dec.log 2>&1
ls -la gpurun_out/*.ncu-rep
