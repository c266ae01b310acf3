import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'oracle')]
import torch, senas_b200, senas_oracle as oracle
from senas_b200.ops import OpType
from helpers import max_err
senas_b200.exact_fp32(); senas_b200.set_conv_mode('bf16')
DEV='cuda:0'
for H in (16, 24, 40):
    torch.manual_seed(11)
    m = senas_b200.MixedOp(32, 8, OpType.NORM); m.apply(senas_b200.weights_init)
    x = torch.randn(2, 32, H, 128)
    for cand in (2, 3):
        alpha = torch.zeros(6); alpha[1] = 1.0; alpha[cand] = 1.0
        st = oracle.clone_store(m.state_dict())
        xo, ao = x.clone().requires_grad_(True), alpha.clone().requires_grad_(True)
        ref = oracle.mixed_op(oracle.Params(st), 'NORM', xo, ao, True)
        gout = torch.randn(ref.shape, generator=torch.Generator().manual_seed(1)); ref.backward(gout)
        mg = m.to(DEV); mg.zero_grad()
        xg, ag = x.to(DEV).requires_grad_(True), alpha.to(DEV).requires_grad_(True)
        out = mg(xg, ag, ag); out.backward(gout.to(DEV))
        name = f'_ops.{cand}.0.weight'
        # what bf16 rounding of the operands alone does to this gradient (oracle with rounded dy and W is hard to get; use x/W rounding on forward as a scale)
        print(f'H={H} cand {cand}: out {max_err(out, ref.detach()):.1e} gx {max_err(xg.grad, xo.grad):.1e} dW {max_err(dict(mg.named_parameters())[name].grad, st[name].grad):.1e} |gx|max {xo.grad.abs().max():.2e} rms {xo.grad.pow(2).mean().sqrt():.2e}')
        m = m.cpu()
