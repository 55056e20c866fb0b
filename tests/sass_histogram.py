"""Per-kernel SASS opcode histogram of the built library (no GPU needed): the mnemonics that prove tcgen05 / TMEM / TMA /
cluster-launch-control / multimem use, per B200_PROFILING.md.   python tests/sass_histogram.py > profiles/rN_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "ml_inference_optimizer_b200", "libb200_attn_mlp.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCOMMA", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAPF", "UTMACMDFLUSH", "UCGABAR",
        "UTCATOMSWS", "SYNCS", "HMMA", "MUFU", "LDGMC", "STGMC", "REDMC", "MULTIMEM", "LDG", "STG", "RED", "ATOM", "BAR", "LDS", "STS", "LDSM", "LDL", "STL"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = per.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["_total"] += 1
            for k in KEYS:
                if op.startswith(k):
                    cur[k] += 1
                    break
    names = demangle(list(per))
    tot = collections.Counter()
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(per)} kernels; counts of SASS instructions by mnemonic prefix")
    for fn, c in per.items():
        nm = re.sub(r"\(.*", "", names.get(fn, fn))[:110]
        body = " ".join(f"{k}={c[k]}" for k in KEYS if c[k])
        print(f"{nm}: total={c['_total']} {body}")
        tot.update(c)
    print("# library totals: " + " ".join(f"{k}={tot[k]}" for k in KEYS if tot[k]))


if __name__ == "__main__":
    main()
