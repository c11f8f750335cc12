"""Aggregate an `ncu --page source --csv` (SASS view) export per barrier-delimited segment of a kernel: samples,
instructions, top stall reasons and hottest instructions.  Usage: python tools/ncu_segments.py src.csv [min_share]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
hdr = rows[1]
data = [r for r in rows[2:] if len(r) >= len(hdr)]
ci = {n: i for i, n in enumerate(hdr)}
stalls = [n for n in hdr if n.startswith('stall_') and 'Not Issued' not in n]
new = lambda: {'n': 0, 'samples': 0, 'inst': 0, 'st': {s: 0 for s in stalls}, 'top': [], 'wf': 0, 'wfi': 0}
seg, cur, tot = [], new(), 0
for r in data:
    src = r[ci['Source']].strip()
    smp, ins = int(r[ci['# Samples']] or 0), int(r[ci['Instructions Executed']] or 0)
    cur['n'] += 1; cur['samples'] += smp; cur['inst'] += ins; tot += smp
    cur['wf'] += int(r[ci['L1 Wavefronts Shared']] or 0); cur['wfi'] += int(r[ci['L1 Wavefronts Shared Ideal']] or 0)
    for s in stalls:
        cur['st'][s] += int(r[ci[s]] or 0)
    cur['top'].append((smp, src, ins, r[ci['L1 Wavefronts Shared']], r[ci['L1 Wavefronts Shared Ideal']]))
    if 'BAR.SYNC' in src:
        seg.append(cur); cur = new()
seg.append(cur)
print('total samples', tot)
for i, s in enumerate(seg):
    if s['samples'] < tot * thr:
        continue
    st = sorted(s['st'].items(), key=lambda x: -x[1])[:5]
    print(f"seg{i:2d} sass {s['n']:4d} samples {s['samples']:6d} ({100 * s['samples'] / tot:4.1f}%) inst {s['inst']:9d} smem wavefronts {s['wf']}/{s['wfi']} | "
          + ' '.join(f"{k[6:]}={v}" for k, v in st))
    for t in sorted(s['top'], key=lambda x: -x[0])[:4]:
        print(f"        {t[0]:5d} {t[1][:72]:72s} exec {t[2]} wf {t[3]}/{t[4]}")
