#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29561 tests/ring_timeline.py > gpurun_out/r2_ring_timeline_n4.jsonl 2> gpurun_out/r2_ring_timeline_n4.err; echo rc=$?
grep '^{' gpurun_out/r2_ring_timeline_n4.jsonl | head -4 | cut -c1-1500
tail -3 gpurun_out/r2_ring_timeline_n4.err
