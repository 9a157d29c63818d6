"""Hot SASS regions of one kernel in an .ncu-rep: contiguous instruction runs with equal execution counts, with their share of
executed instructions, of warp-state samples and the three dominant stall reasons.
usage: python tools/ncu_regions.py REPORT KERNEL_REGEX [N_REGIONS]"""
import csv, io, re, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
nreg = int(sys.argv[3]) if len(sys.argv) > 3 else 16
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass', '--kernel-name', f'regex:{kre}'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
hdr = rows[hi[0]]; ix = {h: i for i, h in enumerate(hdr)}
end = hi[1] - 1 if len(hi) > 1 else len(rows)
body = [r for r in rows[hi[0] + 1:end] if len(r) > 8]
cnt = [int(r[ix['Instructions Executed']] or 0) for r in body]
smp = [int(r[ix['# Samples']] or 0) for r in body]
stallcols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
runs, s = [], 0
for i in range(1, len(body) + 1):
    if i == len(body) or abs(cnt[i] - cnt[s]) > 0.02 * max(cnt[s], 1):
        runs.append((s, i, cnt[s])); s = i
tot, ts = sum(cnt), sum(smp)
print('kernel', kre, 'sass', len(body), 'warp-inst', tot, 'samples', ts)
tot_st = {h: sum(int(r[ix[h]] or 0) for r in body) for h in stallcols}
print('stalls overall:', ', '.join(f'{k[6:]} {v / ts * 100:.1f}%' for k, v in sorted(tot_st.items(), key=lambda x: -x[1])[:9]))
for s, e, c in sorted(sorted(runs, key=lambda x: -sum(smp[x[0]:x[1]]))[:nreg]):
    ops = {}
    for r in body[s:e]:
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)', r[ix['Source']]); op = m.group(2) if m else '?'
        ops[op] = ops.get(op, 0) + 1
    st = {h: sum(int(r[ix[h]] or 0) for r in body[s:e]) for h in stallcols}
    top = ', '.join(f'{k[6:]} {v}' for k, v in sorted(st.items(), key=lambda x: -x[1])[:3])
    opstr = ' '.join(f'{k}:{v}' for k, v in sorted(ops.items(), key=lambda x: -x[1])[:5])
    print(f'[{s:5d},{e:5d}) n={e - s:4d} exec={c:9d} inst%={(e - s) * c / tot * 100:5.1f} smp%={sum(smp[s:e]) / ts * 100:5.1f} | {opstr} | {top}')
