import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, senas_b200
from senas_b200.loss import SegmentationLosses
senas_b200.exact_fp32(); senas_b200.set_conv_mode('bf16'); torch.backends.cudnn.benchmark = True
dev = 'cuda:0'
def build():
    torch.manual_seed(0)
    m = senas_b200.NAS(1, 32, 2, depth=5, meta_node_num=3, use_sharing=False, double_down_channel=False, supervision=False).to(dev).train()
    w = torch.optim.SGD(m.parameters(), lr=5e-3, momentum=0.9, weight_decay=3e-4)
    a = torch.optim.Adam(m.arch_parameters(), lr=1e-4, betas=(0.5, 0.999), weight_decay=1e-3)
    return m, w, a
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
g = torch.Generator().manual_seed(1)
xs = [torch.randn(B, 1, 256, 256, generator=g).to(dev) for _ in range(2)]
ys = [(torch.rand(B, 256, 256, generator=g) > 0.8).long().to(dev) for _ in range(2)]
crit = SegmentationLosses('dice_ce')
# eager reference trajectory
m, w, a = build()
def eager():
    a.zero_grad(); crit(m(xs[1]), ys[1]).backward(); a.step()
    w.zero_grad(); l = crit(m(xs[0]), ys[0]); l.backward(); torch.nn.utils.clip_grad_norm_(m.parameters(), 5); w.step(); return l.item()
le = [eager() for _ in range(6)]
# graphed: 3 warm-up steps inside the constructor, then 3 replays = steps 4..6 of the same trajectory
m2, w2, a2 = build()
step = senas_b200.GraphedSearchStep(m2, crit, w2, a2, (xs[0], ys[0], xs[1], ys[1]), force_segments=len(sys.argv) > 2 and sys.argv[2] != 'one', capture_error_mode=sys.argv[3] if len(sys.argv) > 3 else 'global')
lg = [step(xs[0], ys[0], xs[1], ys[1]).item() for _ in range(2)]
print('eager losses ', le)
print('graph losses ', lg, '(capture = step 4, replays = steps 5, 6)')
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): step(xs[0], ys[0], xs[1], ys[1])
e1.record(); torch.cuda.synchronize()
print(f'graphed step: {e0.elapsed_time(e1)/5:.1f} ms  -> {B/(e0.elapsed_time(e1)/5e3):.1f} img/s')
