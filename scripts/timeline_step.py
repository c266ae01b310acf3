"""Timeline statistics of ONE replay of the captured search step (torch.profiler / kineto kernel records): how much of
the step has 0 / 1 / 2-3 / 4+ kernels in flight, the largest idle gaps, and the kernels that run alone the longest.

    python scripts/timeline_step.py [bf16|fp32] [B] [serial|concurrent]
"""
import os, sys, json, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
import senas_b200
from senas_b200.loss import SegmentationLosses

mode = sys.argv[1] if len(sys.argv) > 1 else 'bf16'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
conc = (sys.argv[3] if len(sys.argv) > 3 else 'concurrent') == 'concurrent'
senas_b200.exact_fp32(); senas_b200.set_conv_mode(mode); torch.backends.cudnn.benchmark = True
if mode == 'bf16':
    torch.backends.cudnn.allow_tf32 = True
dev = 'cuda:0'
torch.manual_seed(0)
m = senas_b200.NAS(1, 32, 2, depth=5, meta_node_num=3, use_sharing=False, double_down_channel=False, supervision=False).to(dev).train()
w = torch.optim.SGD(m.parameters(), lr=5e-3, momentum=0.9, weight_decay=3e-4)
a = torch.optim.Adam(m.arch_parameters(), lr=1e-4, betas=(0.5, 0.999), weight_decay=1e-3)
crit = SegmentationLosses('dice_ce')
g = torch.Generator().manual_seed(1)
xs = [torch.randn(B, 1, 256, 256, generator=g).to(dev) for _ in range(2)]
ys = [(torch.rand(B, 256, 256, generator=g) > 0.8).long().to(dev) for _ in range(2)]
step = senas_b200.GraphedSearchStep(m, crit, w, a, (xs[0], ys[0], xs[1], ys[1]), concurrent_cells=conc, fused_optim=True)
for _ in range(2):
    step(xs[0], ys[0], xs[1], ys[1])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(xs[0], ys[0], xs[1], ys[1])
    torch.cuda.synchronize()
path = os.path.join(ROOT, 'gpurun_out', 'step_trace.json')
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))['traceEvents'] if e.get('cat') in ('kernel', 'gpu_memset', 'gpu_memcpy') and 'dur' in e]
os.remove(path)
ev.sort(key=lambda e: e['ts'])
t0, t1 = ev[0]['ts'], max(e['ts'] + e['dur'] for e in ev)
print(f'{len(ev)} device activities on {len(set(e["args"].get("stream") for e in ev))} streams, span {(t1 - t0) / 1e3:.2f} ms, '
      f'sum of durations {sum(e["dur"] for e in ev) / 1e3:.2f} ms')
pts = []
for e in ev:
    pts.append((e['ts'], 1)); pts.append((e['ts'] + e['dur'], -1))
pts.sort()
hist, cur, last = collections.Counter(), 0, t0
gaps = []
for t, d in pts:
    if t > last:
        hist[min(cur, 8)] += t - last
        if cur == 0:
            gaps.append((t - last, last - t0))
    cur += d; last = t
tot = sum(hist.values())
print('time with k kernels in flight: ' + '  '.join(f'{k}{"+" if k == 8 else ""}: {100 * v / tot:.1f}%' for k, v in sorted(hist.items())))
print(f'idle gaps: {len(gaps)} totalling {sum(g for g, _ in gaps) / 1e3:.2f} ms; largest: ' +
      ', '.join(f'{g:.0f}us@{at / 1e3:.1f}ms' for g, at in sorted(gaps, reverse=True)[:8]))
# kernels running alone: attribute single-occupancy time to the kernel name
alone = collections.Counter()
active = []
idx = 0
evs = sorted(ev, key=lambda e: e['ts'])
bounds = sorted(set([e['ts'] for e in ev] + [e['ts'] + e['dur'] for e in ev]))
import bisect
starts = [e['ts'] for e in evs]
live = []
j = 0
for bi in range(len(bounds) - 1):
    lo, hi = bounds[bi], bounds[bi + 1]
    while j < len(evs) and evs[j]['ts'] <= lo:
        live.append(evs[j]); j += 1
    live = [e for e in live if e['ts'] + e['dur'] > lo]
    if len(live) == 1:
        e = live[0]
        grid, block = e['args'].get('grid', [0, 0, 0]), e['args'].get('block', [0, 0, 0])
        thr = grid[0] * grid[1] * grid[2] * block[0] * block[1] * block[2]
        small = thr < 148 * 1024  # less than half of what the 148 SMs can hold (2048 threads each)
        alone[(e['name'][:100], small)] += hi - lo
print(f'time as the ONLY kernel in flight: {sum(alone.values()) / 1e3:.2f} ms, of which kernels with < 148k threads '
      f'{sum(v for (k, sm), v in alone.items() if sm) / 1e3:.2f} ms')
print('by kernel (S = small grid):')
for (k, sm), v in alone.most_common(40):
    print(f'  {v / 1e3:7.2f} ms  {"S" if sm else " "}  {k}')
