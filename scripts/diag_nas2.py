"""Full-supernet backward: gradient at the inputs / output of every Cell.nodes(), GPU path vs fp64 oracle (dev aid)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'oracle')]
import torch
import senas_b200, senas_oracle as oracle
from senas_b200.cell import Cell
senas_b200.exact_fp32()
DEV = 'cuda:0'
B, H = 2, 64
torch.manual_seed(0)
m = senas_b200.NAS(1, 32, 2, depth=5, meta_node_num=3, use_sharing=False, double_down_channel=False, supervision=False)
gen = torch.Generator().manual_seed(1234)
x = torch.randn(B, 1, H, H, generator=gen); y = (torch.rand(B, H, H, generator=gen) > 0.8).long()
store = {k: (v.detach().clone().to(DEV).double() if v.is_floating_point() else v.clone().to(DEV)) for k, v in m.state_dict().items()}
for n, _ in m.named_parameters(): store[n].requires_grad_(True)
rec_o = []
ocn = oracle.cell_nodes
def cn_rec(p, ct, in0, in1, *a, **k):
    in0.retain_grad(); in1.retain_grad()
    o = ocn(p, ct, in0, in1, *a, **k); o.retain_grad(); rec_o.append((p.prefix, in0, in1, o)); return o
oracle.cell_nodes = cn_rec
oracle.dice_ce_loss(oracle.nas_forward(store, x.to(DEV).double())[-1], y.to(DEV)).backward()
oracle.cell_nodes = ocn
m = m.to(DEV); m.train()
rec_g = []
nd = Cell.nodes
def nd_rec(self, in0, in1, *a):
    in0.retain_grad(); in1.retain_grad()
    o = nd(self, in0, in1, *a); o.retain_grad(); rec_g.append((in0, in1, o)); return o
Cell.nodes = nd_rec
oracle.dice_ce_loss(m(x.to(DEV))[-1], y.to(DEV)).backward()
Cell.nodes = nd
def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm()).item()
for (pref, i0, i1, o), (g0, g1, go) in zip(rec_o, rec_g):
    print(f'{pref:28s} fwd: in0 {rel(g0, i0):.1e} in1 {rel(g1, i1):.1e} cat {rel(go, o):.1e} | bwd: gcat {rel(go.grad, o.grad):.1e} gin0 {rel(g0.grad, i0.grad):.1e} gin1 {rel(g1.grad, i1.grad):.1e}'
          f'  strides in0 {tuple(g0.stride())} in1 {tuple(g1.stride())} gcat {tuple(go.grad.stride())}')
