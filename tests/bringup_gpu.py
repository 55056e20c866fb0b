"""GPU bring-up checks with verbose diagnostics (run on a B200 through gpurun; not collected by pytest).

    python tests/bringup_gpu.py            # every case, each in its own process under a timeout
    python tests/bringup_gpu.py gemm_small # one case in this process

Each case compares the C-ABI path against the fp32 oracle and prints error statistics plus, on failure, a map of
where the errors are — enough to tell a descriptor/swizzle mistake from a protocol mistake in one round trip.
"""
from __future__ import annotations

import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _stats(name, got, ref, atol, rtol_mean=None):
    import torch
    got, ref = got.float().cpu(), ref.float().cpu()
    finite = torch.isfinite(got).all().item()
    diff = (got - ref).abs()
    diff = torch.where(torch.isfinite(diff), diff, torch.full_like(diff, 1e30))
    mx = diff.max().item()
    mean_rel = (diff / (ref.abs() + 1e-8)).mean().item()
    ok = finite and mx <= atol and (rtol_mean is None or mean_rel <= rtol_mean)
    print(f"[{name}] max_abs={mx:.4e} mean_rel={mean_rel:.4e} finite={finite} ref_absmax={ref.abs().max().item():.3f} "
          f"-> {'OK' if ok else 'FAIL'}", flush=True)
    return ok, diff


def _blockmap(diff2d, br, bc, thresh):
    """print a map of blocks whose max error exceeds thresh"""
    R, C = diff2d.shape
    lines = []
    for r0 in range(0, R, br):
        row = ""
        for c0 in range(0, C, bc):
            row += "X" if diff2d[r0:r0 + br, c0:c0 + bc].max().item() > thresh else "."
        lines.append(row)
    print("\n".join(lines[:64]), flush=True)


# ------------------------------------------------------------------------------------------------ cases
def case_decode():
    import torch
    from ml_inference_optimizer_b200 import ops
    from oracle import attn_mlp_oracle as orc
    torch.manual_seed(0)
    ok_all = True
    for (B, Hq, Hkv, D, S, splits) in [(2, 4, 4, 128, 300, 1), (3, 8, 2, 128, 1000, 0), (2, 4, 4, 64, 517, 3),
                                      (4, 8, 8, 128, 2048, 0), (2, 8, 1, 64, 700, 2)]:
        q = torch.randn(B, Hq, D, device="cuda", dtype=torch.bfloat16)
        kc = torch.randn(B, S, Hkv, D, device="cuda", dtype=torch.bfloat16)
        vc = torch.randn(B, S, Hkv, D, device="cuda", dtype=torch.bfloat16)
        lens = torch.randint(1, S + 1, (B,), device="cuda", dtype=torch.int32)
        lens[0] = S
        o, lse = ops.decode_attention(q, kc, vc, lens, num_splits=splits, return_lse=True)
        torch.cuda.synchronize()
        ro, rl = orc.decode_attention_ref(q.cpu(), kc.cpu(), vc.cpu(), lens.cpu())
        a, _ = _stats(f"decode B{B} Hq{Hq} Hkv{Hkv} D{D} S{S} splits{splits} O", o, ro, 2e-2)
        b, _ = _stats(f"decode B{B} Hq{Hq} Hkv{Hkv} D{D} S{S} splits{splits} LSE", lse, rl, 1e-2)
        ok_all &= a and b
    # paged
    B, Hq, Hkv, D, bs, L, nblk = 3, 8, 4, 128, 16, 2, 64
    kc = torch.randn(nblk, L, bs, Hkv, D, device="cuda", dtype=torch.bfloat16)
    vc = torch.randn(nblk, L, bs, Hkv, D, device="cuda", dtype=torch.bfloat16)
    lens = torch.tensor([37, 256, 129], device="cuda", dtype=torch.int32)
    perm = torch.randperm(nblk)[:B * 16].view(B, 16).to(device="cuda", dtype=torch.int32)
    q = torch.randn(B, Hq, D, device="cuda", dtype=torch.bfloat16)
    o, lse = ops.decode_attention(q, kc, vc, lens, block_tables=perm, layer_idx=1, return_lse=True)
    torch.cuda.synchronize()
    ro, rl = orc.decode_attention_ref(q.cpu(), kc.cpu(), vc.cpu(), lens.cpu(), block_tables=perm.cpu(), layer_idx=1)
    a, _ = _stats("decode paged O", o, ro, 2e-2)
    b, _ = _stats("decode paged LSE", lse, rl, 1e-2)
    ok_all &= a and b
    # kv_append
    key = torch.randn(B, Hkv, D, device="cuda", dtype=torch.bfloat16)
    val = torch.randn(B, Hkv, D, device="cuda", dtype=torch.bfloat16)
    kc2, vc2 = kc.clone(), vc.clone()
    ops.kv_append(key, val, kc2, vc2, lens, block_tables=perm, layer_idx=1)
    torch.cuda.synchronize()
    rk, rv = kc.cpu().clone(), vc.cpu().clone()
    orc.kv_append_ref(key.cpu(), val.cpu(), rk, rv, lens.cpu(), perm.cpu(), 1)
    a = torch.equal(kc2.cpu(), rk) and torch.equal(vc2.cpu(), rv)
    print(f"[kv_append paged] bit-exact={a}", flush=True)
    ok_all &= a
    return ok_all


def _gemm_case(T, K, N, act, seed=0, bias=True):
    import torch
    from ml_inference_optimizer_b200 import ops
    from oracle import attn_mlp_oracle as orc
    torch.manual_seed(seed)
    x = torch.randn(T, K, device="cuda", dtype=torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
    b = (torch.randn(N, device="cuda") * 0.5).to(torch.bfloat16) if bias else None
    wg = bg = None
    if act == "swiglu":
        wg = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
        bg = (torch.randn(N, device="cuda") * 0.5).to(torch.bfloat16) if bias else None
    y = ops.linear_act(x, w, b, act, wg, bg)
    torch.cuda.synchronize()
    ref = orc.linear_act_ref(x.cpu(), w.cpu(), None if b is None else b.cpu(), act, None if wg is None else wg.cpu(),
                             None if bg is None else bg.cpu())
    tol = 2e-2 * max(1.0, ref.abs().max().item() / 4)
    ok, diff = _stats(f"linear_act T{T} K{K} N{N} act={act}", y, ref, tol, 2e-2)
    if not ok:
        _blockmap(diff, 8, 8, tol)
        print("got[0,:8]", y[0, :8].float().cpu().tolist())
        print("ref[0,:8]", ref[0, :8].tolist())
        print("got[1,:8]", y[1, :8].float().cpu().tolist())
        print("ref[1,:8]", ref[1, :8].tolist())
    return ok


def case_gemm_small():
    ok = _gemm_case(128, 64, 256, None, bias=False)
    ok &= _gemm_case(128, 128, 256, None, bias=False)
    ok &= _gemm_case(128, 256, 256, None)
    return ok


def case_gemm_shapes():
    ok = True
    for (T, K, N, act) in [(256, 512, 512, None), (300, 768, 3072, "gelu_tanh"), (1000, 3072, 768, None),
                           (512, 1024, 1024, "gelu"), (512, 1024, 1024, "relu"), (384, 512, 384, "swiglu"),
                           (2048, 4096, 1408, "swiglu"), (100, 264, 200, "gelu_tanh")]:
        ok &= _gemm_case(T, K, N, act)
    return ok


def case_mlp():
    import torch
    from ml_inference_optimizer_b200 import ops
    from oracle import attn_mlp_oracle as orc
    ok = True
    for (T, h, i, act) in [(512, 768, 3072, "gelu_tanh"), (1024, 1024, 2816, "swiglu"), (77, 256, 512, "relu")]:
        torch.manual_seed(1)
        x = torch.randn(T, h, device="cuda", dtype=torch.bfloat16)
        wu = (torch.randn(i, h, device="cuda") * 0.02).to(torch.bfloat16)
        bu = (torch.randn(i, device="cuda") * 0.02).to(torch.bfloat16)
        wd = (torch.randn(h, i, device="cuda") * 0.02).to(torch.bfloat16)
        bd = (torch.randn(h, device="cuda") * 0.02).to(torch.bfloat16)
        wg = bg = None
        if act == "swiglu":
            wg = (torch.randn(i, h, device="cuda") * 0.02).to(torch.bfloat16)
            bg = (torch.randn(i, device="cuda") * 0.02).to(torch.bfloat16)
        y = ops.fused_mlp(x, wu, bu, wd, bd, act, wg, bg)
        torch.cuda.synchronize()
        ref = orc.mlp_ref(x.cpu(), wu.cpu(), bu.cpu(), wd.cpu(), bd.cpu(), act, None if wg is None else wg.cpu(),
                          None if bg is None else bg.cpu())
        a, _ = _stats(f"fused_mlp T{T} h{h} i{i} {act}", y, ref, 2e-2, 2e-2)
        ok &= a
    return ok


def _fa_case(B, Sq, Sk, Hq, Hkv, D, causal, offset=0, kv_lens=None, seed=0, dtype=None):
    import torch
    from ml_inference_optimizer_b200 import ops
    from oracle import attn_mlp_oracle as orc
    dtype = dtype or torch.bfloat16
    torch.manual_seed(seed)
    q = torch.randn(B, Sq, Hq, D, device="cuda", dtype=dtype)
    k = torch.randn(B, Sk, Hkv, D, device="cuda", dtype=dtype)
    v = torch.randn(B, Sk, Hkv, D, device="cuda", dtype=dtype)
    lens = None if kv_lens is None else torch.tensor(kv_lens, device="cuda", dtype=torch.int32)
    o, lse = ops.flash_attn_fwd(q, k, v, causal=causal, causal_offset=offset, kv_lens=lens, return_lse=True)
    torch.cuda.synchronize()
    ro, rl = orc.attention_ref(q.cpu(), k.cpu(), v.cpu(), causal=causal, causal_offset=offset,
                               kv_lens=None if lens is None else lens.cpu())
    name = f"fa B{B} Sq{Sq} Sk{Sk} Hq{Hq} Hkv{Hkv} D{D} causal={causal} off={offset} lens={kv_lens} {dtype}"
    a, diff = _stats(name + " O", o, ro, 2e-2, 3e-2)
    lse_c, rl_c = lse.cpu(), rl
    both_inf = torch.isinf(lse_c) & torch.isinf(rl_c) & (lse_c < 0) & (rl_c < 0)
    b, dl = _stats(name + " LSE", torch.where(both_inf, torch.zeros_like(lse_c), lse_c),
                   torch.where(both_inf, torch.zeros_like(rl_c), rl_c), 1e-2)
    if not a:
        d2 = diff[0, :, 0, :]  # [Sq, D] of batch 0 head 0
        print("error map (rows = 16 queries, cols = 8 dims), batch0 head0:")
        _blockmap(d2, 16, 8, 2e-2)
        print("got[0,0,0,:8]", o[0, 0, 0, :8].float().cpu().tolist())
        print("ref[0,0,0,:8]", ro[0, 0, 0, :8].tolist())
        print("got[0,1,0,:8]", o[0, 1, 0, :8].float().cpu().tolist())
        print("ref[0,1,0,:8]", ro[0, 1, 0, :8].tolist())
    if not b:
        print("lse got", lse[0, 0, :8].cpu().tolist())
        print("lse ref", rl[0, 0, :8].tolist())
    return a and b


def case_fa_small():
    ok = _fa_case(1, 128, 128, 1, 1, 128, False)
    ok &= _fa_case(1, 256, 256, 1, 1, 128, False)
    ok &= _fa_case(1, 256, 256, 2, 2, 128, True)
    return ok


def case_fa_d64():
    ok = _fa_case(1, 128, 128, 1, 1, 64, False)
    ok &= _fa_case(2, 512, 512, 3, 3, 64, True)
    return ok


def case_fa_shapes():
    import torch
    ok = True
    ok &= _fa_case(2, 1024, 1024, 4, 4, 128, True)
    ok &= _fa_case(2, 1024, 1024, 8, 2, 128, True)          # GQA
    ok &= _fa_case(1, 300, 300, 2, 2, 128, True)            # ragged
    ok &= _fa_case(1, 77, 333, 2, 1, 64, False)             # cross lengths
    ok &= _fa_case(2, 512, 512, 2, 2, 128, False, kv_lens=[512, 100])
    ok &= _fa_case(1, 256, 512, 2, 2, 128, True, offset=256)   # bottom-right / ring step (later chunk)
    ok &= _fa_case(1, 256, 256, 2, 2, 128, True, offset=-128)  # rows with no visible key
    ok &= _fa_case(1, 512, 512, 2, 2, 128, True, dtype=torch.float16)
    ok &= _fa_case(1, 2048, 2048, 2, 2, 128, True, seed=3)
    return ok


def case_merge():
    import torch
    from ml_inference_optimizer_b200 import ops
    from oracle import attn_mlp_oracle as orc
    torch.manual_seed(0)
    B, S, H, D = 2, 200, 4, 128
    oa = torch.randn(B, S, H, D, device="cuda")
    la = torch.randn(B, H, S, device="cuda")
    ob = torch.randn(B, S, H, D, device="cuda", dtype=torch.bfloat16)
    lb = torch.randn(B, H, S, device="cuda")
    la[0, 0, :5] = float("-inf")
    lb[0, 1, :5] = float("-inf")
    la[1, 2, :3] = float("-inf")
    lb[1, 2, :3] = float("-inf")
    ro, rl = orc.lse_merge_ref(oa.cpu(), la.cpu(), ob.cpu(), lb.cpu())
    ops.lse_merge(oa, la, ob, lb)
    torch.cuda.synchronize()
    a, _ = _stats("lse_merge O", oa, ro, 1e-4)
    both = torch.isinf(la.cpu()) & torch.isinf(rl)
    b, _ = _stats("lse_merge LSE", torch.where(both, torch.zeros_like(rl), la.cpu()), torch.where(both, torch.zeros_like(rl), rl), 1e-4)
    out = ops.cast_out(oa, torch.bfloat16)
    torch.cuda.synchronize()
    c = torch.equal(out.cpu(), oa.cpu().to(torch.bfloat16))
    print(f"[cast_out] bit-exact={c}", flush=True)
    return a and b and c


CASES = {
    "decode": case_decode,
    "merge": case_merge,
    "gemm_small": case_gemm_small,
    "gemm_shapes": case_gemm_shapes,
    "mlp": case_mlp,
    "fa_small": case_fa_small,
    "fa_d64": case_fa_d64,
    "fa_shapes": case_fa_shapes,
}


def main():
    if len(sys.argv) > 1 and sys.argv[1] != "all":
        ok = True
        for name in sys.argv[1:]:
            t0 = time.time()
            try:
                r = CASES[name]()
            except Exception as e:  # noqa: BLE001
                import traceback
                traceback.print_exc()
                r = False
            print(f"== case {name}: {'PASS' if r else 'FAIL'} ({time.time() - t0:.1f}s)", flush=True)
            ok &= bool(r)
        sys.exit(0 if ok else 1)
    results = {}
    for name in CASES:
        t0 = time.time()
        try:
            res = subprocess.run(["timeout", "-s", "KILL", "240", sys.executable, __file__, name], capture_output=True,
                                 text=True)
            out = res.stdout + res.stderr
            code = res.returncode
        except Exception as e:  # noqa: BLE001
            out, code = str(e), -1
        print(out[-6000:], flush=True)
        results[name] = code
        print(f"#### {name}: exit {code} in {time.time() - t0:.1f}s", flush=True)
    print("SUMMARY", results, flush=True)
    sys.exit(0 if all(v == 0 for v in results.values()) else 1)


if __name__ == "__main__":
    main()
