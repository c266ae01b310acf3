"""Full-supernet fwd+bwd: per-cell output / output-gradient / arch-gradient diff, GPU path vs CPU oracle (dev aid)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'oracle')]
import torch
import senas_b200, senas_oracle as oracle
from senas_b200.cell import Cell
from helpers import max_err
senas_b200.exact_fp32()
DEV = 'cuda:0'
B, H = 2, 64
torch.manual_seed(0)
m = senas_b200.NAS(1, 32, 2, depth=5, meta_node_num=3, use_sharing=False, double_down_channel=False, supervision=False)
store = oracle.clone_store(m.state_dict())
m = m.to(DEV); m.train()
gen = torch.Generator().manual_seed(1234)
x = torch.randn(B, 1, H, H, generator=gen); y = (torch.rand(B, H, H, generator=gen) > 0.8).long()
# oracle with recording
rec_o = []
ocell = oracle.cell
def cell_rec(p, *a, **k):
    o = ocell(p, *a, **k); o.retain_grad(); rec_o.append((p.prefix, o)); return o
oracle.cell = cell_rec
lo = oracle.dice_ce_loss(oracle.nas_forward(store, x)[-1], y); lo.backward()
oracle.cell = ocell
# ours with recording
rec_g = []
fwd = Cell.forward
def fwd_rec(self, *a):
    o = fwd(self, *a); o.retain_grad(); rec_g.append(o); return o
Cell.forward = fwd_rec
lg = oracle.dice_ce_loss(m(x.to(DEV))[-1], y.to(DEV)); lg.backward()
Cell.forward = fwd
print('loss', lo.item(), lg.item())
for (pref, o), g in zip(rec_o, rec_g):
    print(f'{pref:28s} out {max_err(g, o.detach()):.1e}  grad_out {max_err(g.grad, o.grad):.1e}')
for n in ('alphas_dn', 'alphas_up', 'alphas_dn_nm', 'alphas_up_nm', 'betas_dn', 'betas_up', 'gamma'):
    print(n, f'{max_err(getattr(m, n).grad, store[n].grad):.1e}', getattr(m, n).grad.abs().max().item())
worst = sorted(((max_err(p.grad, store[n].grad), n) for n, p in m.named_parameters() if store[n].grad is not None), reverse=True)[:12]
for e, n in worst: print(f'{e:.1e} {n}')
