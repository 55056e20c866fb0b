import sys, os
sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo")
import perf_probe as pp
pp.attn(4, 8192, 32, 32, 128, True, ("ours",))
pp.attn(4, 8192, 32, 32, 128, False, ("ours",))
pp.attn(8, 4096, 12, 12, 64, True, ("ours",))
