#!/usr/bin/env python
"""Copy one measurement run from gpurun_out/ into profiles/ (bench line, secondary configs, reference arm, launch list)
and print its summary.  usage: tools/archive_run.py <tag>   (reads gpurun_out/<tag>_{bench,configs,ref}.log, <tag>_launches.csv)"""
import collections, csv, json, shutil, sys

tag = sys.argv[1]
d = json.loads(open(f'gpurun_out/{tag}_bench.log').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'launches', d['gpu_launches'])
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['pcie'])
print('roof', {k: d['roofline'][k] for k in ('kernel', 'achieved', 'peak', 'frac', 'traffic', 'algorithmic_bytes')})
print('stages', {k: round(v, 2) for k, v in d['roofline']['stage_ms'].items()})
print('cpu', d['cpu_baseline']['value'])
c = d['create']
print('create', c['value'], c['ms_per_step'], c['e2e']['value'], c['e2e']['ms_per_step'], {k: round(v, 2) for k, v in c['stage_ms'].items()},
      c['c_gpu_over_c_ref'], c['cpu_baseline']['value'])
for l in open(f'gpurun_out/{tag}_configs.log'):
    j = json.loads(l)
    print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in j.items() if k not in ('path', 'note', 'checked', 'codec')})
print(open(f'gpurun_out/{tag}_ref.log').read()[:160])
for a, b in ((f'{tag}_bench.log', f'r1_{tag}_bench_1024.log'), (f'{tag}_configs.log', f'r1_{tag}_configs.log'), (f'{tag}_ref.log', f'r1_{tag}_reference_arm.log')):
    shutil.copy('gpurun_out/' + a, 'profiles/' + b)
rows = [r for r in csv.reader(open(f'gpurun_out/{tag}_launches.csv')) if len(r) > 10]
h = rows[0]; ki = h.index('Kernel Name'); vi = h.index('Metric Value'); ui = h.index('Metric Unit')
raw = []
for r in rows[1:]:
    name = r[ki].split('(')[0].replace('void ', '')
    v = float(r[vi].replace(',', '')); u = r[ui]
    ms = v / 1e6 if u in ('nsecond', 'ns') else v / 1e3 if u in ('usecond', 'us') else v
    raw.append((r[h.index('ID')], name, ms))
with open(f'profiles/r1_{tag}_launches_raw.csv', 'w', newline='') as f:
    w = csv.writer(f); w.writerow(['id', 'kernel', 'ms'])
    for r in raw: w.writerow([r[0], r[1], f'{r[2]:.4f}'])
idx = [i for i, r in enumerate(raw) if 'LzCfg<4' in r[1]]
step = raw[idx[1] + 1:idx[2] + 1]
agg = collections.OrderedDict()
for id_, name, ms in step:
    agg.setdefault(name, [0, 0.0]); agg[name][0] += 1; agg[name][1] += ms
tot = sum(v[1] for v in agg.values())
with open(f'profiles/r1_{tag}_launches_summary.csv', 'w', newline='') as f:
    f.write('# ncu launch list (gpu__time_duration.sum, --clock-control none, -c 60) of: python bench.py --entries 1024 --steps 2 --warmup 1 --e2e-steps 0 --create 0\n')
    f.write(f'# ONE timed kernel-only step (the launches between the 2nd and 3rd 4-warp zstd_lz launch; all 60: r1_{tag}_launches_raw.csv).\n')
    f.write('# cold-cache + serialised: compare SHARES with bench.py stage_share, not absolutes.\n')
    w = csv.writer(f); w.writerow(['kernel', 'launches', 'total_ms', 'share'])
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]): w.writerow([k, n, f'{ms:.3f}', f'{ms / tot:.4f}'])
print(open(f'profiles/r1_{tag}_launches_summary.csv').read())
