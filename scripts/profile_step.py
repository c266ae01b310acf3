"""Kernel table of one eager search step (torch.profiler, CUDA activities): every kernel of the step -- libsenas_b200's
and the stock PyTorch blocks around the cells (stems, Shrink/Rectify blocks, loss, optimizers) -- by total device time.
Lanes are off so that durations are those of kernels running alone.

    python scripts/profile_step.py [bf16|fp32] [B] [rows]
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
import senas_b200
from senas_b200.loss import SegmentationLosses

mode = sys.argv[1] if len(sys.argv) > 1 else 'bf16'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
rows = int(sys.argv[3]) if len(sys.argv) > 3 else 60
senas_b200.exact_fp32(); senas_b200.set_conv_mode(mode); torch.backends.cudnn.benchmark = True
if mode == 'bf16':
    torch.backends.cudnn.allow_tf32 = True
senas_b200._lib.get().senas_set_lanes(0)
dev = 'cuda:0'
torch.manual_seed(0)
m = senas_b200.NAS(1, 32, 2, depth=5, meta_node_num=3, use_sharing=False, double_down_channel=False, supervision=False).to(dev).train()
w = torch.optim.SGD(m.parameters(), lr=5e-3, momentum=0.9, weight_decay=3e-4)
a = torch.optim.Adam(m.arch_parameters(), lr=1e-4, betas=(0.5, 0.999), weight_decay=1e-3)
crit = SegmentationLosses('dice_ce')
g = torch.Generator().manual_seed(1)
xs = [torch.randn(B, 1, 256, 256, generator=g).to(dev) for _ in range(2)]
ys = [(torch.rand(B, 256, 256, generator=g) > 0.8).long().to(dev) for _ in range(2)]


def step():
    a.zero_grad(); crit(m(xs[1]), ys[1]).backward(); a.step()
    w.zero_grad(); l = crit(m(xs[0]), ys[0]); l.backward(); torch.nn.utils.clip_grad_norm_(m.parameters(), 5); w.step()


for _ in range(2): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
ev = prof.key_averages()
ours = ('gather_mac', 'conv_wgrad', 'conv_tc', 'dw_', 'pw_', 'adapter_', 'node_', 'bn_reduce', 'bn_finalize', 'rows_reduce',
        'wgrad_reduce', 'pack_dy', 'cast_bf16', 'tc_wgrad_reduce')
tot = sum(e.device_time_total for e in ev) / 1e3
mine = sum(e.device_time_total for e in ev if any(k in e.key for k in ours)) / 1e3
print(f'total device time {tot:.1f} ms; libsenas_b200 kernels {mine:.1f} ms; everything else {tot - mine:.1f} ms')
print('--- everything else, by device time')
for e in sorted([e for e in ev if not any(k in e.key for k in ours)], key=lambda e: -e.device_time_total)[:rows]:
    print(f'{e.device_time_total/1e3:9.3f} ms  n={e.count:5d}  {e.key[:150]}')
print('--- libsenas_b200, by device time')
for e in sorted([e for e in ev if any(k in e.key for k in ours)], key=lambda e: -e.device_time_total)[:rows]:
    print(f'{e.device_time_total/1e3:9.3f} ms  n={e.count:5d}  {e.key[:150]}')
