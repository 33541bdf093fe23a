#!/usr/bin/env python
"""Join gpurun_out/ops.csv (per-op CUDA-event timings from bench.py --dump-ops) with the planner's debug lines."""
import csv, sys
ops = list(csv.DictReader(open(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/ops.csv')))
plans = [l.strip() for l in open(sys.argv[2] if len(sys.argv) > 2 else 'gpurun_out/plan.txt') if l.startswith('conv_plan')]
pi = 0; tot = 0; agg = {}
for o in ops:
    k = int(o['kind']); us = float(o['us_avg']); tot += us
    pl = ''
    if k == 0:
        pl = plans[pi] if pi < len(plans) else ''
        pi += 1
    key = pl.split(' grid')[0] if pl else f'kind{k}'
    a = agg.setdefault(key, [0, 0.0, 0.0, pl]); a[0] += 1; a[1] += us; a[2] += float(o['gflop'])
    if '-v' in sys.argv:
        print(f"{o['op']:>3} k{k} L{o['layer']:>3} {float(o['gflop']):7.3f}GF {us:7.2f}us {float(o['tflops']):6.1f}TF  {pl[10:]}")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{a[1]:8.1f}us  x{a[0]:<3} avg {a[1]/a[0]:6.1f}us {a[2]/max(a[1],1e-9)*1e3:7.1f}TF  {a[3][10:] if a[3] else key}")
print('total us', round(tot, 1))
