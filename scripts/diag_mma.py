"""gather_mma vs gather_mac on the same MixedOp (GPU): where do they differ?  usage: diag_mma.py op c_in H B"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import copy, torch
import senas_b200
from helpers import OP_BY_ID, max_err
op_id, c_in, H, B = [int(v) for v in sys.argv[1:5]]
lib = senas_b200._lib.get()
senas_b200.exact_fp32(); senas_b200.set_conv_mode('bf16')
torch.manual_seed(1)
m0 = senas_b200.MixedOp(c_in, 8, OP_BY_ID[op_id]); m0.apply(senas_b200.weights_init)
x = torch.randn(B, c_in, H, H)
res = {}
for k in (2, 3, 1):
    alpha = torch.zeros(6); alpha[k] = 1.0
    for mma in (0, 1):
        lib.senas_set_gather_mma(mma)
        m = copy.deepcopy(m0).cuda()
        xg, ag = x.cuda().requires_grad_(True), alpha.cuda().requires_grad_(True)
        out = m(xg, ag, ag)
        torch.manual_seed(2)
        out.backward(torch.randn(out.shape, device='cuda'))
        torch.cuda.synchronize()
        res[mma] = (out.detach().cpu(), xg.grad.cpu())
    o_err, g_err = max_err(res[1][0], res[0][0]), max_err(res[1][1], res[0][1])
    d = (res[1][1] - res[0][1]).abs() > 0.02 * res[0][1].abs().max()
    print(f'cand {k}: out {o_err:.2e} gx {g_err:.2e} bad {d.sum().item()} of {d.numel()}')
    if d.any():
        for py in (0, 1):
            for px in (0, 1):
                print('   phase', py, px, d[:, :, py::2, px::2].float().mean().item())
        print('   by channel block', [round(d[:, c:c + 8].float().mean().item(), 3) for c in range(0, c_in, 8)])
        print('   by column 16-block', [round(d[:, :, :, c:c + 16].float().mean().item(), 3) for c in range(0, H, 16)])
        print('   by row', [round(d[:, :, r].float().mean().item(), 2) for r in range(0, min(H, 16))])
