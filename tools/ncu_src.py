"""Summarise an ncu report's source page for one kernel: stall breakdown + hottest SASS lines.
usage: python tools/ncu_src.py report.ncu-rep kernel_regex [top_n] [--list lo hi]"""
import csv, subprocess, sys, io
rep, kre = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3].isdigit() else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[h]
ia, ii, it, isamp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Avg. Threads Executed'), hdr.index('# Samples')
stall_cols = [(j, c) for j, c in enumerate(hdr) if c.startswith('stall_') and 'Not Issued' not in c]
data = []
for r in rows[h + 1:]:
    if len(r) > ii and r[ii].isdigit():
        data.append(r)
    if r and r[0] == 'Address' : break   # first kernel instance only
tot = sum(int(r[ii]) for r in data); ts = sum(int(r[isamp]) for r in data)
print('SASS instr', len(data), 'warp-instr executed', tot, 'samples', ts)
agg = {}
for r in data:
    for j, c in stall_cols:
        if r[j].isdigit(): agg[c] = agg.get(c, 0) + int(r[j])
for c, v in sorted(agg.items(), key=lambda x: -x[1])[:10]: print(f'  {c:28s} {100*v/ts:5.1f}%')
if '--list' in sys.argv:
    k = sys.argv.index('--list'); lo, hi = int(sys.argv[k+1]), int(sys.argv[k+2])
    sel = list(enumerate(data))[lo:hi]
else:
    sel = sorted(sorted(enumerate(data), key=lambda x: -int(x[1][isamp]))[:topn])
for n, r in sel:
    st = sorted(((int(r[j]) if r[j].isdigit() else 0, c.replace('stall_','')) for j, c in stall_cols), reverse=True)[:2]
    print(f"{n:5d} {r[ia][:58]:58s} {r[ii]:>10s} {r[it]:>5s} {100*int(r[isamp])/ts:5.1f}% {st[0][1]}:{st[0][0]} {st[1][1]}:{st[1][0]}")
