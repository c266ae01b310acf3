"""Summarise an `ncu --csv` log (launch list with dram bytes / duration) per kernel family -> JSON + table.

    python scripts/ncu_summarise.py gpurun_out/step_traffic.csv profiles/r2_traffic.json "source text"
"""
import csv, json, re, sys, collections

path, out_json = sys.argv[1], sys.argv[2]
source = sys.argv[3] if len(sys.argv) > 3 else path
rows = []
with open(path, newline='') as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    rows.append(r)
acc = collections.defaultdict(lambda: dict(n=0, rd=0.0, wr=0.0, us=0.0))
ids = set()


def to_bytes(v, unit):
    v = float(v.replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)


def to_us(v, unit):
    v = float(v.replace(',', ''))
    return v * {'ns': 1e-3, 'us': 1, 'ms': 1e3, 's': 1e6}.get(unit, 1)


for r in rows:
    name = re.sub(r'\(.*', '', r['Kernel Name'])
    name = re.sub(r'^void ', '', name)
    a = acc[name]
    m, v, u = r['Metric Name'], r['Metric Value'], r['Metric Unit']
    if m == 'dram__bytes_read.sum':
        a['rd'] += to_bytes(v, u)
    elif m == 'dram__bytes_write.sum':
        a['wr'] += to_bytes(v, u)
    elif m == 'gpu__time_duration.sum':
        a['us'] += to_us(v, u)
        a['n'] += 1
tot_rd, tot_wr = sum(a['rd'] for a in acc.values()), sum(a['wr'] for a in acc.values())
tot_us = sum(a['us'] for a in acc.values())
ours = {k: v for k, v in acc.items() if not k.startswith(('at::', 'cudnn', 'cutlass', 'nhwc', 'implicit_', 'wgrad_alg', 'sm100', 'sm90', 'void cu'))}
res = {'source': source, 'dram_bytes_per_step': tot_rd + tot_wr, 'dram_read_bytes': tot_rd, 'dram_write_bytes': tot_wr,
       'kernels': sum(a['n'] for a in acc.values()), 'serialised_device_ms': tot_us / 1e3,
       'by_kernel': {k: {'launches': v['n'], 'ms': round(v['us'] / 1e3, 3), 'dram_gb': round((v['rd'] + v['wr']) / 1e9, 3)}
                     for k, v in sorted(acc.items(), key=lambda kv: -(kv[1]['rd'] + kv[1]['wr']))[:60]}}
json.dump(res, open(out_json, 'w'), indent=1)
print(f'{res["kernels"]} kernels, {tot_us / 1e3:.1f} ms serialised, DRAM read {tot_rd / 1e9:.2f} GB + write {tot_wr / 1e9:.2f} GB')
for k, v in list(res['by_kernel'].items())[:25]:
    print(f'  {v["dram_gb"]:8.2f} GB {v["ms"]:8.2f} ms  n={v["launches"]:5d}  {k[:90]}')
