"""Where does the replayed search step leave the eager one? (GPU)  usage: diag_graphed.py fp32|bf16 flags(e.g. 101)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import torch
import senas_b200
from senas_b200.loss import SegmentationLosses
from test_gpu_parity_r2 import _new_nas, _optimizers, _batches

mode = sys.argv[1] if len(sys.argv) > 1 else 'fp32'
flags = [c == '1' for c in (sys.argv[2] if len(sys.argv) > 2 else '101')]
senas_b200.exact_fp32()
senas_b200.set_conv_mode(mode)
batches = _batches(len(flags), 2, 64)
ref = _new_nas().cuda().train()
w_opt, a_opt = _optimizers(ref)
crit = SegmentationLosses('dice_ce')
arch = senas_b200.Architecture(ref, a_opt, crit)
want = []
for (xt, yt, xv, yv), fa in zip(batches, flags):
    if fa:
        arch.step(xv, yv)
    w_opt.zero_grad()
    loss = crit(ref(xt), yt)
    loss.backward()
    torch.nn.utils.clip_grad_norm_(ref.parameters(), 5)
    w_opt.step()
    want.append((loss.item(), {k: v.detach().clone() for k, v in ref.state_dict().items()}))
m = _new_nas().cuda().train()
init = {k: v.detach().clone() for k, v in m.state_dict().items()}
w2, a2 = _optimizers(m)
step = senas_b200.GraphedSearchStep(m, SegmentationLosses('dice_ce'), w2, a2, batches[0], grad_clip=5.0, warmup=3)
print('restored:', all(torch.equal(v, init[k]) for k, v in m.state_dict().items()))
prev_w, prev_g = init, init
for i, ((xt, yt, xv, yv), fa) in enumerate(zip(batches, flags)):
    loss = step(xt, yt, xv, yv, arch=fa)
    torch.cuda.synchronize()
    got = {k: v.detach().clone() for k, v in m.state_dict().items()}
    print(f'step {i} arch={fa} loss {loss.item():.7f} want {want[i][0]:.7f}')
    rows = []
    for k, v in want[i][1].items():
        if not v.is_floating_point():
            if not torch.equal(got[k], v):
                rows.append((1e9, k))
            continue
        uw, ug = (v - prev_w[k]).double(), (got[k] - prev_g[k]).double()
        scale = max(uw.abs().max().item(), 1e-3 * v.abs().max().item(), 1e-7)
        rows.append(((ug - uw).abs().max().item() / scale, k))
    rows.sort(reverse=True)
    for e, k in rows[:6]:
        print(f'    {e:.2e} {k}')
    prev_w, prev_g = want[i][1], got
