import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'oracle')]
import torch, senas_b200, senas_oracle as oracle
from helpers import max_err
senas_b200.exact_fp32(); senas_b200.set_conv_mode(sys.argv[1] if len(sys.argv) > 1 else 'bf16')
DEV='cuda:0'
torch.manual_seed(11)
c = senas_b200.Cell(3, 1, 32, 32, 32, 'up'); c.apply(senas_b200.weights_init)
store = oracle.clone_store(c.state_dict())
in0, in1 = torch.randn(2, 32, 16, 128), torch.randn(2, 32, 8, 64)
b = torch.ones(9)
cg = c.to(DEV)
for edge in (0, 2, 5):
    for cand in (2, 3):
        wn = torch.zeros(9, 6); wn[:, 1] = 1.0   # everything on 'none' ...
        wn[edge] = 0; wn[edge, cand] = 1.0        # ... except one conv candidate of one in0 edge
        wc = torch.zeros(9, 6); wc[:, 0] = 1.0
        t = [v.clone().requires_grad_(True) for v in (in0, in1, wn, wc, b)]
        st = oracle.clone_store(c.state_dict())
        ref = oracle.cell_nodes(oracle.Params(st), 'up', *t, training=True)
        gout = torch.randn(ref.shape, generator=torch.Generator().manual_seed(1)); ref.backward(gout)
        g = [v.to(DEV).requires_grad_(True) for v in (in0, in1, wn, wc, b)]
        cg.zero_grad()
        out = cg.nodes(*g); out.backward(gout.to(DEV))
        name = f'_ops.{edge}._ops.{cand}.0.weight'
        pg = dict(cg.named_parameters())[name].grad
        print(f'edge {edge} cand {cand}: cat {max_err(out, ref.detach()):.1e} gin0 {max_err(g[0].grad, t[0].grad):.1e} dW {max_err(pg, st[name].grad):.1e}')
