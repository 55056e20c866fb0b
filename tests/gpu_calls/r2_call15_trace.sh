cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B200_TRACE_LIB=gpurun_in/trace.so timeout 120 python tests/fa_trace.py 128 0 > gpurun_out/fa_trace_new.txt 2>&1; cat gpurun_out/fa_trace_new.txt
