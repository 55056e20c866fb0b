"""Dev tool: cycle-count A/B of attention-kernel builds inside ONE process (for ncu --metrics sm__cycles_elapsed.max):
loads gpurun_in/<name>.so side by side through ctypes and launches the same problems on each."""
import ctypes, os, shutil, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ml_inference_optimizer_b200._lib import SIGNATURES, c_int64_p

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

def load(path):
    lib = ctypes.CDLL(path)
    res, args = SIGNATURES["b200_fa_fwd"]
    lib.b200_fa_fwd.restype, lib.b200_fa_fwd.argtypes = res, args
    lib.b200_last_error.restype = ctypes.c_char_p
    return lib

def run(lib, q, k, v, o, causal):
    B, Sq, Hq, D = q.shape
    Sk, Hkv = k.shape[1], k.shape[2]
    st = lambda t: (ctypes.c_int64 * 3)(*t.stride()[:3])
    rc = lib.b200_fa_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), None, B, Sq, Sk, Hq, Hkv, D, st(q), st(k),
                         st(v), st(o), D ** -0.5, int(causal), 0, None, 0, torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.b200_last_error()

names = sys.argv[1:2] or ["new"]
if names[0].startswith("pair"):
    os.environ["B200_FA_PAIR"] = "1"  # the CTA-pair kernel of the same library (read per call)  # one build per process: template-static state is shared between copies of the library
libs = {}
for n in names:
    src = n.split(":")[0]
    path = os.path.join(ROOT, "gpurun_in", src + ".so")
    if ":" in n:  # "new:np" = the same build with B200_FA_PERSISTENT=0 (needs its own copy: the flag is read once per library)
        cp = os.path.join(ROOT, "gpurun_in", n.replace(":", "_") + ".so")
        shutil.copy(path, cp)
        path = cp
    libs[n] = load(path)

torch.manual_seed(0)
B, H, D, S = 4, 32, 128, 8192
if os.environ.get("FA_AB_D") == "64":
    B, H, D, S = 8, 12, 64, 8192   # GPT-2 head shape, longer sequence so the fit has the same tile counts
q = torch.randn(B, S, H, D, device="cuda", dtype=torch.bfloat16)
k = torch.randn_like(q); v = torch.randn_like(q); o = torch.empty_like(q)
ev = lambda: torch.cuda.Event(enable_timing=True)
for n, lib in libs.items():
    if n.endswith(":np"): os.environ["B200_FA_PERSISTENT"] = "0"
    else: os.environ.pop("B200_FA_PERSISTENT", None)
    run(lib, q, k, v, o, True)  # first call reads the flag
torch.cuda.synchronize()
os.environ.pop("B200_FA_PERSISTENT", None)
cases = [("causal", S, True), ("full", S, False)] + [(f"sk{t*128}", t * 128, False) for t in (1, 2, 4, 8, 16, 32)]
for rep in range(2):
    for name, Sk, causal in cases:
        for n, lib in libs.items():
            s, e = ev(), ev()
            s.record(); run(lib, q, k[:, :Sk], v[:, :Sk], o, causal); e.record(); torch.cuda.synchronize()
            print(f"rep{rep} {name:8s} {n:8s} {s.elapsed_time(e)*1e3:9.1f} us", flush=True)
