"""Dev tool: tabulate gpurun_out/fa_ab_ncu.csv (see fa_ab3.sh)."""
import collections, csv, io, sys
names = sys.argv[1:]
lines = [l for l in open('gpurun_out/fa_ab_ncu.csv').read().splitlines() if l.startswith('"')]
data = collections.defaultdict(dict)
for row in csv.DictReader(io.StringIO("\n".join(lines))):
    if 'fa_fwd' not in row['Kernel Name']:
        continue
    data[(row['Process ID'], int(row['ID']))][row['Metric Name']] = float(row['Metric Value'].replace(',', ''))
pids = []
for pid, _ in data:
    if pid not in pids:
        pids.append(pid)
cases = ['warm'] + ['causal', 'full', 'sk128', 'sk256', 'sk512', 'sk1024', 'sk2048', 'sk4096'] * 2
tab = collections.defaultdict(dict)
for pid, name in zip(pids, names):
    ks = sorted(k for k in data if k[0] == pid)
    for k, c in zip(ks, cases):
        tab[c].setdefault(name, []).append(data[k]['sm__cycles_elapsed.max'])
for c in cases[1:9]:
    print(f"{c:8s}", ' | '.join(f"{n}: " + ','.join(f"{cy/1e3:.0f}" for cy in tab[c][n]) + " kcyc" for n in names))
for n in names:
    import os
    waves = (8 * 12 * 32 if os.environ.get('FA_AB_D') == '64' else 4096) / 148
    slope = (tab['full'][n][0] - tab['sk4096'][n][0]) / 32 / waves
    print(f"{n}: {slope:.0f} cycles per KV iteration (2 query tiles), fixed {tab['sk128'][n][0]/waves - slope:.0f} cycles per item")
