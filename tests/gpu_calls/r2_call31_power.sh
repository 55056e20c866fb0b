cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for pr in 0 1; do echo "PAIR=$pr"; B200_FA_PAIR=$pr SECS=3 timeout 200 python tests/fa_power_probe.py 2>&1 | grep -v again | grep "c3_causal\|c3_full"; done | tee gpurun_out/r2_fa_power_pair.log
