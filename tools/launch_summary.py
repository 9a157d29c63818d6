"""Summary of an ncu launch list (`--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`) of
`bench.py --steps 1 --warmup 1`: per kernel, launches / total time / share of the step / DRAM bytes per launch, over the
launches of ONE 400-pose job (the timed step = the second `k_step_consts`-to-`k_update` run of 20 reverse steps).
Writes profiles/conv_lv3_traffic.json (read by bench.py for roofline.traffic) when --traffic-json is given."""
import argparse, collections, csv, json, re, sys

ap = argparse.ArgumentParser()
ap.add_argument('csv')
ap.add_argument('--job', type=int, default=1, help='which 20-step job of the capture (0 = warm-up, 1 = timed step)')
ap.add_argument('--rev-steps', type=int, default=20)
ap.add_argument('--traffic-json')
args = ap.parse_args()
rows = collections.OrderedDict()
with open(args.csv) as f:
    lines = [l for l in f if not l.startswith('==')]
for r in csv.DictReader(lines):
    d = rows.setdefault(int(r['ID']), {'name': r['Kernel Name']})
    d[r['Metric Name']] = float(r['Metric Value'].replace(',', ''))
launches = [rows[k] for k in sorted(rows)]
short = lambda n: re.sub(r'\(.*', '', n.replace('ddk::', '').replace('void ', ''))
# jobs are delimited by k_update launches: rev_steps of them per job
upd = [i for i, l in enumerate(launches) if 'k_update' in l['name']]
first = [i for i, l in enumerate(launches) if 'k_step_consts' in l['name']]
lo = first[args.job * args.rev_steps]
hi = upd[(args.job + 1) * args.rev_steps - 1] + 1
sel = launches[lo:hi]
tot = sum(l['gpu__time_duration.sum'] for l in sel)
agg = collections.OrderedDict()
for l in sel:
    a = agg.setdefault(short(l['name']), [0, 0.0, 0.0])
    a[0] += 1; a[1] += l['gpu__time_duration.sum']
    a[2] += l.get('dram__bytes_read.sum', 0.0) + l.get('dram__bytes_write.sum', 0.0)
print(f'launches {lo}..{hi} ({len(sel)}), device time {tot / 1e6:.1f} ms (cold-cache, serialised: compare shares)')
for k, (n, t, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{k:34s} n={n:4d}  {t / 1e6:8.2f} ms  {100 * t / tot:5.1f} %   dram {b / n / 1e6:9.2f} MB/launch  {b / max(t, 1) :7.1f} GB/s')
if args.traffic_json:
    is3 = lambda l, k: (k + '<3') in l['name'] or (k + '<(int)3') in l['name']
    k3 = [l for l in sel if is3(l, 'k_conv_tcr')] or [l for l in sel if is3(l, 'k_conv_fused')]
    both = [l for l in sel if is3(l, 'k_conv_tcr') or is3(l, 'k_conv_fused') or is3(l, 'k_acc_tc')]   # the kernels of one 84-wide conv layer
    b = sum(l.get('dram__bytes_read.sum', 0.0) + l.get('dram__bytes_write.sum', 0.0) for l in both)
    t = sum(l['gpu__time_duration.sum'] for l in both)
    json.dump({'kernel': ' + '.join(sorted(set(short(l['name']) for l in both))), 'workload': 'dense', 'launches': len(k3), 'dram_bytes_per_launch': b / max(len(k3), 1),
               'ms_per_launch_under_ncu': t / max(len(k3), 1) / 1e6, 'share_of_step_under_ncu': t / tot,
               'source': 'ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none '
                         'python bench.py --steps 1 --warmup 1 --no-cpu-baseline (launches of one 400-pose job, job index %d of the capture)' % args.job},
              open(args.traffic_json, 'w'), indent=1)
