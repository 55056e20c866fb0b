"""Dev tool: for one kernel of the built library, list every tcgen05.mma (UTC*MMA) issue group in SASS order with the
number of instructions that separate consecutive MMAs and the straight-line instruction count between the last
mbarrier try-wait and the first MMA of the group (work on the critical issue path).
usage: python tests/sass_mma_gaps.py [substring of the mangled kernel name] [library]"""
import re, subprocess, sys
pat = sys.argv[1] if len(sys.argv) > 1 else "fa_fwd_kernelILi128E13__nv_bfloat16"
lib = sys.argv[2] if len(sys.argv) > 2 else "ml_inference_optimizer_b200/libb200_attn_mlp.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
ops, on = [], False
for line in txt.splitlines():
    if "Function :" in line:
        if on: break
        on = pat in line
        continue
    if on:
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(.*?);", line)
        if m: ops.append(m.group(1).strip())
mma = [i for i, o in enumerate(ops) if re.search(r"UTC\w*MMA", o)]
groups, cur = [], [mma[0]]
for a, b in zip(mma, mma[1:]):
    if b - a > 40: groups.append(cur); cur = [b]
    else: cur.append(b)
groups.append(cur)
for g in groups:
    back = 0
    i = g[0] - 1
    while i >= 0 and "TRYWAIT" not in ops[i] and "BSYNC" not in ops[i] and not ops[i].startswith("BRA"):
        if "NOP" not in ops[i]: back += 1
        i -= 1
    kind = "QK" if "gdesc" in ops[g[0]].split(",")[0] else "PV"
    r2ur = sum("R2UR" in o for o in ops[i:g[-1]])
    print(f"{kind} group @{g[0]:5d}: {len(g)} MMAs, lead-in {back:3d} instrs, gaps {[b - a - 1 for a, b in zip(g, g[1:])]}, R2UR {r2ur}")
print("total instrs", len(ops), "LDL", sum(o.startswith("LDL") for o in ops), "STL", sum(o.startswith("STL") for o in ops))
