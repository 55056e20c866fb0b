import sys, os
sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo")
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import perf_probe as pp
which = sys.argv[1] if len(sys.argv) > 1 else "std"
if which == "std":
    pp.attn(4, 8192, 32, 32, 128, True, ("ours",))
    pp.attn(4, 8192, 32, 32, 128, False, ("ours",))
    pp.attn(8, 4096, 12, 12, 64, True, ("ours",))
else:
    pp.attn(8, 4096, 12, 12, 64, True, ("ours",))
    pp.attn(8, 4096, 12, 12, 64, False, ("ours",))
    pp.attn(2, 16384, 12, 12, 64, False, ("ours", "sdpa"))
    pp.attn(2, 16384, 12, 12, 64, True, ("ours", "sdpa"))
    pp.attn(1, 32768, 16, 16, 128, True, ("ours", "sdpa"))
