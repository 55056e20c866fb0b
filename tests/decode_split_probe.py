"""Developer probe: decode attention time vs number of KV splits for small problems (few (batch, kv head) pairs)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ml_inference_optimizer_b200 import ops
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda"); bf = torch.bfloat16
for name, B, S, Hq, Hkv, D in (("mqa_b32", 32, 8192, 16, 1, 128), ("gqa_b8", 8, 8192, 32, 8, 128), ("gqa_b1_32k", 1, 32768, 32, 8, 128),
                               ("gqa_b64", 64, 8192, 32, 8, 128), ("gqa_b16_s2k", 16, 2048, 32, 8, 128), ("mha_b4", 4, 8192, 32, 32, 128),
                               ("gqa4_b64_d64", 64, 8192, 16, 4, 64), ("gqa4_b8_d64", 8, 8192, 16, 4, 64), ("mha_b8_d64", 8, 4096, 12, 12, 64),
                               ("mha_b1_d64_1k", 1, 1024, 12, 12, 64), ("gqa_b4_s512", 4, 512, 32, 8, 128)):
    q = torch.randn(B, Hq, D, device="cuda", dtype=bf)
    kc, vc = (torch.randn(B, S, Hkv, D, device="cuda", dtype=bf) for _ in range(2))
    lens = torch.full((B,), S, device="cuda", dtype=torch.int32)
    res = {}
    for splits in (0, 1, 2, 3, 4, 6, 8, 16, 32, 64, 128):
        if splits > (S + 255) // 256: continue
        try:
            for _ in range(3): ops.decode_attention(q, kc, vc, lens, num_splits=splits)
        except Exception as e:
            res[splits] = str(e)[:40]; continue
        ts = []
        for _ in range(11):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); ops.decode_attention(q, kc, vc, lens, num_splits=splits); e.record(); torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        res[splits] = round(sorted(ts)[5] * 1e3, 1)
    nbytes = 2.0 * B * S * Hkv * D * 2
    best = min((v, k) for k, v in res.items() if isinstance(v, float))
    print(json.dumps({"case": name, "ctas_per_split": B * Hkv, "us_by_splits(0=auto)": res, "best": best, "best_gbs": round(nbytes / best[0] / 1e3, 0),
                      "auto_gbs": round(nbytes / res[0] / 1e3, 0)}), flush=True)
