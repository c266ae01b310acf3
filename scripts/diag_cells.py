"""Per-cell parity inside a real NAS forward: capture each Cell.nodes() input on the GPU, replay it through the
oracle on the CPU with the same parameters and a random cotangent (development aid)."""
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
m = senas_b200.NAS(1, 32, 2, depth=5, meta_node_num=3, use_sharing=False, double_down_channel=False, supervision=False).to(DEV)
m.train()
captured = []
orig = Cell.nodes
def rec(self, in0, in1, wn, wc, b):
    captured.append((self, in0.detach().clone(), in1.detach().clone(), wn.detach().clone(), wc.detach().clone(), b.detach().clone()))
    return orig(self, in0, in1, wn, wc, b)
Cell.nodes = rec
gen = torch.Generator().manual_seed(1234)
x = torch.randn(B, 1, H, H, generator=gen).to(DEV)
with torch.no_grad():
    m(x)
Cell.nodes = orig
names = {id(mod): n for n, mod in m.named_modules()}
for (c, in0, in1, wn, wc, b) in captured:
    ctype = 'down' if c._ops[0]._op_type.name == 'DOWN' else 'up'
    store = oracle.clone_store(c.state_dict())
    t = [v.cpu().clone().requires_grad_(True) for v in (in0, in1, wn, wc, b)]
    ref = oracle.cell_nodes(oracle.Params(store), ctype, *t)
    gout = torch.randn(ref.shape)
    ref.backward(gout)
    c.zero_grad()
    g = [v.clone().requires_grad_(True) for v in (in0, in1, wn, wc, b)]
    out = orig(c, *g)
    out.backward(gout.to(DEV))
    norm = c._norm_rows
    ga = torch.where(norm, g[2].grad, g[3].grad).cpu(); ga_ref = torch.where(norm.cpu(), t[2].grad, t[3].grad)
    worst, wn_ = 0, ''
    for n, p in c._ops.named_parameters():
        e = max_err(p.grad, store['_ops.' + n].grad)
        if e > worst: worst, wn_ = e, n
    print(f'{names[id(c)]:28s} {ctype:4s} in1 {tuple(in1.shape[2:])} out {max_err(out, ref.detach()):.1e} gin0 {max_err(g[0].grad, t[0].grad):.1e} '
          f'gin1 {max_err(g[1].grad, t[1].grad):.1e} galpha {max_err(ga, ga_ref):.1e} gbeta {max_err(g[4].grad, t[4].grad):.1e} worst-param {worst:.1e} {wn_}')
