cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tests/ln_probe.py > gpurun_out/r2_ln_probe5.log 2>&1; grep "COPY\|narrow_max=dflt wide_ctas=occ" gpurun_out/r2_ln_probe5.log | awk '!seen[$0]++'
