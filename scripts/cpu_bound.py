"""How long does the host need to ENQUEUE one search step vs how long the GPU needs to run it (dev aid)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, senas_b200
from senas_b200.loss import SegmentationLosses
tf32 = len(sys.argv) > 1 and sys.argv[1] == 'tf32'
senas_b200.exact_fp32(); senas_b200.set_conv_mode('bf16')
if tf32:
    torch.backends.cudnn.allow_tf32 = True
torch.backends.cudnn.benchmark = True
dev = 'cuda:0'
torch.manual_seed(0)
m = senas_b200.NAS(1, 32, 2, depth=5, meta_node_num=3, use_sharing=False, double_down_channel=False, supervision=False).to(dev).train()
crit = SegmentationLosses('dice_ce')
w_opt = torch.optim.SGD(m.parameters(), lr=5e-3, momentum=0.9, weight_decay=3e-4)
a_opt = torch.optim.Adam(m.arch_parameters(), lr=1e-4, betas=(0.5, 0.999), weight_decay=1e-3)
x = torch.randn(16, 1, 256, 256, device=dev); y = (torch.rand(16, 256, 256, device=dev) > 0.8).long()
def step():
    a_opt.zero_grad(); crit(m(x), y).backward(); a_opt.step()
    w_opt.zero_grad(); l = crit(m(x), y); l.backward(); torch.nn.utils.clip_grad_norm_(m.parameters(), 5); w_opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
for _ in range(2):
    t0 = time.perf_counter(); step(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f'tf32={tf32}: enqueue {1e3*(t1-t0):.1f} ms, total {1e3*(t2-t0):.1f} ms')
# split: forward only / backward only enqueue
t0 = time.perf_counter(); out = m(x); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f'forward: enqueue {1e3*(t1-t0):.1f} ms, total {1e3*(t2-t0):.1f} ms')
l = crit(out, y); torch.cuda.synchronize()
t0 = time.perf_counter(); l.backward(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f'backward: enqueue {1e3*(t1-t0):.1f} ms, total {1e3*(t2-t0):.1f} ms')
t0 = time.perf_counter(); torch.nn.utils.clip_grad_norm_(m.parameters(), 5); w_opt.step(); w_opt.zero_grad(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f'clip+sgd+zero: enqueue {1e3*(t1-t0):.1f} ms, total {1e3*(t2-t0):.1f} ms')
