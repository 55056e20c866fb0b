"""Developer probe: bench.py's e2e leg in isolation and after the other legs (which one slows it down?)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
args = argparse.Namespace(gpus=1, steps=10, warmup=3, impl="ours", workload="c3", no_cpu_baseline=True, no_secondary=True, no_ring=True, group_rows=0)
w = bench.WORKLOADS["c3"]
cx = bench.Ctx()
step, st = bench.build_step(cx, w, args)
def e2e(tag):
    r = bench.e2e_leg(cx, w, st, args)
    print(json.dumps({"probe": "e2e_bisect", "when": tag, "ms_per_step": round(r["ms"], 3)}), flush=True)
e2e("fresh")
for _ in range(20): step()
cx.torch.cuda.synchronize()
e2e("after 20 steps")
s = bench.ClockSampler(0); s.start(); time.sleep(0.5)
for _ in range(20): step()
cx.torch.cuda.synchronize()
s.stop(time.time() - 0.3, time.time())
e2e("after sampler start/stop")
p = bench.parity_check(cx)
e2e("after parity_check")
import torch
print("threads", torch.get_num_threads(), "affinity", len(os.sched_getaffinity(0)))
