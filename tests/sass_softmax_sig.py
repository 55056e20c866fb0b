"""Dev tool: signature of the softmax hot loop of the attention kernel in SASS (instruction count between the S load and
the two P stores, MUFU placement), to compare builds offline. usage: python tests/sass_softmax_sig.py lib.so [...]"""
import re, subprocess, sys
pat = "fa_fwd_kernelILi128E13__nv_bfloat16"
for lib in sys.argv[1:]:
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
    ld = [i for i, o in enumerate(ops) if o.startswith("LDTM")]
    st = [i for i, o in enumerate(ops) if o.startswith("STTM")]
    first = ld[3]                      # 4th LDTM = end of the S load
    st_after = [i for i in st if i > first]
    resc_end = st_after[3]             # 4 STTMs of the (rare) rescale path
    st1, st2 = st_after[4], st_after[5]
    cnt = lambda a, b, key: sum(key in o for o in ops[a:b])
    nonresc = (ld[4] - first) + (st1 - resc_end)
    print(f"{lib}: S-load->STTM1 {nonresc} instrs (MUFU {cnt(first, st1, 'MUFU')}, FFMA2 {cnt(first, st1, 'FFMA2')}, F2FP {cnt(first, st1, 'F2FP')}, "
          f"SEL {cnt(first, st1, 'SEL')}, MOV {cnt(first, st1, 'MOV')}), STTM1->STTM2 {st2 - st1} instrs (MUFU {cnt(st1, st2, 'MUFU')}), "
          f"total {len(ops)}")
