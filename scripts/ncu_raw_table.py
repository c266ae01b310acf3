"""Per-kernel table from an `ncu --page raw --csv` dump (optionally gzipped):  python scripts/ncu_raw_table.py file.csv[.gz] [N]"""
import csv, gzip, re, collections, sys
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
op = gzip.open if path.endswith('.gz') else open
rows = list(csv.reader(op(path, 'rt')))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
M = dict(t='gpu__time_duration.sum', sm='sm__throughput.avg.pct_of_peak_sustained_elapsed',
         dram='FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed',
         warps='sm__warps_active.avg.pct_of_peak_sustained_active', issue='smsp__issue_active.avg.pct_of_peak_sustained_active',
         regs='launch__registers_per_thread', grid='launch__grid_size', l1='l1tex__throughput.avg.pct_of_peak_sustained_active',
         lts='lts__throughput.avg.pct_of_peak_sustained_elapsed', fma='sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
         lsu='sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', rd='dram__bytes_read.sum', wr='dram__bytes_write.sum',
         long_sb='smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
         short_sb='smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
         mio='smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
         lg='smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
         bar='smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
         wait='smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
         tensor='sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active' if 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active' in ix else 'sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active')


def f(r, k):
    try:
        return float(r[ix[M[k]]].replace(',', ''))
    except Exception:
        return float('nan')


def mb(r, k):
    if M[k] not in ix:
        return float('nan')
    v = f(r, k)
    return v * {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1, 'Gbyte': 1e3}.get(units[ix[M[k]]], 1)


agg = collections.OrderedDict()
for r in data:
    name = re.sub(r'\(.*', '', r[ix['Kernel Name']]).replace('void ', '')
    a = agg.setdefault((name, r[ix[M['grid']]]), [])
    a.append(r)
tu = {'ns': 1e-3, 'us': 1, 'ms': 1e3}.get(units[ix[M['t']]], 1)
print(f'{len(data)} kernels, {sum(f(r, "t") for r in data) * tu / 1e3:.2f} ms')
print(f"{'kernel':38s} {'grid':>6s}  n    us/ea   sm%  dram% warps issue regs   l1%  lts%  fma%  lsu% tens%  MB/ea  TB/s | stalls/issue long short mio lg bar wait")
for (name, grid), rs in sorted(agg.items(), key=lambda kv: -sum(f(r, 't') for r in kv[1]))[:top]:
    n = len(rs)
    av = lambda k: sum(f(r, k) for r in rs) / n
    t = av('t') * tu
    m = sum(mb(r, 'rd') + mb(r, 'wr') for r in rs) / n
    print(f"{name[:38]:38s} {grid:>6s} {n:2d} {t:8.1f} {av('sm'):5.1f} {av('dram'):5.1f} {av('warps'):5.1f} {av('issue'):5.1f} {av('regs'):4.0f} "
          f"{av('l1'):5.1f} {av('lts'):5.1f} {av('fma'):5.1f} {av('lsu'):5.1f} {av('tensor'):5.1f} {m:7.1f} {m / t:6.2f} | "
          f"{av('long_sb'):5.2f} {av('short_sb'):5.2f} {av('mio'):5.2f} {av('lg'):5.2f} {av('bar'):5.2f} {av('wait'):5.2f}")
