"""Diagnose a config-2-size parity failure: one-hot alpha per candidate, per-tensor errors (GPU)."""
import copy
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import torch
import senas_b200
import senas_oracle as oracle
from helpers import OP_BY_ID, OP_NAME, max_err

op_id, c_in, H, B = [int(v) for v in (sys.argv[1:5] + ['3', '32', '64', '16'][len(sys.argv) - 1:])]
senas_b200.exact_fp32()
torch.manual_seed(300 + op_id + c_in + H)
m = senas_b200.MixedOp(c_in, 8, OP_BY_ID[op_id])
m.apply(senas_b200.weights_init)
for mod in m.modules():
    if isinstance(mod, torch.nn.BatchNorm2d):
        mod.weight.data.uniform_(0.5, 1.5)
        mod.bias.data.normal_(0, 0.3)
gen = torch.Generator().manual_seed(H + op_id)
x = torch.randn(B, c_in, H, H, generator=gen)
gout = None
for k in list(range(6)) + [-1]:
    alpha = torch.zeros(6)
    if k >= 0:
        alpha[k] = 1.0
    else:
        alpha = torch.softmax(torch.randn(6, generator=gen), -1)
    store = oracle.clone_store(m.state_dict())
    xo, ao = x.clone().requires_grad_(True), alpha.clone().requires_grad_(True)
    ref = oracle.mixed_op(oracle.Params(store), OP_NAME[op_id], xo, ao, True)
    if gout is None:
        gout = torch.randn(ref.shape, generator=gen)
    ref.backward(gout)
    for lanes in (0, -1):
        senas_b200._lib.get().senas_set_lanes(lanes)
        mg = copy.deepcopy(m).to('cuda')
        xg, ag = x.cuda().requires_grad_(True), alpha.cuda().requires_grad_(True)
        out = mg(xg, ag, ag)
        out.backward(gout.cuda())
        torch.cuda.synchronize()
        errs = {'out': max_err(out, ref.detach()), 'gx': max_err(xg.grad, xo.grad), 'ga': max_err(ag.grad, ao.grad)}
        for n, p in mg.named_parameters():
            e = max_err(p.grad, store[n].grad)
            if e > 1e-4:
                errs[n] = e
        d = (xg.grad.cpu() - xo.grad).abs()
        bad = (d > 1e-4 * xo.grad.abs().max()).nonzero()
        print(f'cand {k} lanes {lanes}:', {a: f'{b:.1e}' for a, b in errs.items()}, 'bad gx elems', len(bad),
              bad[:3].tolist(), bad[-2:].tolist())
