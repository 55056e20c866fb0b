cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "layernorm" 2>&1 | tail -2
timeout 300 python tests/ln_probe.py > gpurun_out/r2_ln_probe4.log 2>&1; grep "narrow_max=dflt wide_ctas=occ" gpurun_out/r2_ln_probe4.log
