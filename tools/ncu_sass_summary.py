#!/usr/bin/env python
"""Summarise an `ncu --page source --csv --print-source sass` export: stall samples and shared-memory wavefronts per
instruction class and per code region (regions are cut at warp-level barriers: WARPSYNC / SYNCS / BAR)."""
import csv, sys, re, collections
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows[:10]):
    if r and r[0] == 'Address':
        hdr, start = r, i + 1
        break
col = {h: i for i, h in enumerate(hdr)}
def f(r, name):
    try: return float(r[col[name]])
    except Exception: return 0.0
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = collections.Counter(); byop = collections.defaultdict(collections.Counter)
regions = []; cur = collections.Counter(); cur_ops = collections.Counter(); nreg = 0
exec_max = max(f(r, 'Instructions Executed') for r in rows[start:] if len(r) == len(hdr))
for r in rows[start:]:
    if len(r) != len(hdr): continue
    sass = r[col['Source']].strip()
    op = re.sub(r'^@!?U?P\d+\s+', '', sass).split()[0].split('.')[0] if sass else '?'
    smp = f(r, '# Samples'); ex = f(r, 'Instructions Executed')
    wf = f(r, 'L1 Wavefronts Shared'); wfi = f(r, 'L1 Wavefronts Shared Ideal')
    byop[op]['samples'] += smp; byop[op]['exec'] += ex; byop[op]['wf'] += wf; byop[op]['wf_ideal'] += wfi; byop[op]['n'] += 1
    for s_ in stalls: byop[op][s_] += f(r, s_); tot[s_] += f(r, s_)
    tot['samples'] += smp; tot['exec'] += ex; tot['wf'] += wf; tot['wf_ideal'] += wfi
    cur['samples'] += smp; cur['exec'] += ex; cur['wf'] += wf; cur['wf_ideal'] += wfi; cur['n'] += 1; cur_ops[op] += 1
    for s_ in stalls: cur[s_] += f(r, s_)
    if op in ('WARPSYNC', 'SYNCS', 'BAR', 'EXIT'):
        regions.append((nreg, cur, cur_ops)); nreg += 1; cur = collections.Counter(); cur_ops = collections.Counter()
regions.append((nreg, cur, cur_ops))
print(f"total samples {tot['samples']:.0f}  exec {tot['exec']:.0f}  shared wavefronts {tot['wf']:.0f} (ideal {tot['wf_ideal']:.0f})")
print("stall totals:", {k[6:]: int(v) for k, v in tot.items() if k.startswith('stall_') and v > 0})
print("\nby opcode (top by samples):")
for op, c in sorted(byop.items(), key=lambda kv: -kv[1]['samples'])[:22]:
    top = sorted(((k[6:], v) for k, v in c.items() if k.startswith('stall_') and v > 0), key=lambda kv: -kv[1])[:4]
    print(f"  {op:10s} n={c['n']:4.0f} exec={c['exec']:10.0f} samples={c['samples']:6.0f} wf={c['wf']:9.0f} ideal={c['wf_ideal']:9.0f}  {top}")
print("\nregions (between warp barriers), samples / exec / wavefronts:")
for i, c, ops in regions:
    if c['samples'] < 0.005 * tot['samples'] and c['wf'] == 0: continue
    top = sorted(((k[6:], v) for k, v in c.items() if k.startswith('stall_') and v > 0), key=lambda kv: -kv[1])[:4]
    mix = ', '.join(f"{k}:{v}" for k, v in ops.most_common(6))
    print(f"  R{i:02d} n={c['n']:4.0f} exec={c['exec']:10.0f} samples={c['samples']:6.0f} ({100*c['samples']/tot['samples']:4.1f}%) wf={c['wf']:9.0f}/{c['wf_ideal']:9.0f}  {top}  [{mix}]")
