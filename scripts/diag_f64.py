"""Deviation from an fp64 ground truth (oracle in double) of: the fp32 oracle on CPU, the fp32 oracle on the GPU
(cuDNN/ATen, TF32 off) and the senas_b200 CUDA path, for one full-supernet fwd+bwd (dev aid / basis of the
full-network parity test)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'oracle')]
import torch
import senas_b200, senas_oracle as oracle
senas_b200.exact_fp32()
B, H = int(sys.argv[1]) if len(sys.argv) > 1 else 2, int(sys.argv[2]) if len(sys.argv) > 2 else 64
torch.manual_seed(0)
m = senas_b200.NAS(1, 32, 2, depth=5, meta_node_num=3, use_sharing=False, double_down_channel=False, supervision=False)
gen = torch.Generator().manual_seed(1234)
x = torch.randn(B, 1, H, H, generator=gen); y = (torch.rand(B, H, H, generator=gen) > 0.8).long()
names = [n for n, _ in m.named_parameters()]
def run_oracle(dev, dtype):
    store = {}
    for k, v in m.state_dict().items():
        t = v.detach().clone().to(dev)
        if t.is_floating_point(): t = t.to(dtype)
        store[k] = t
    for n in names: store[n].requires_grad_(True)
    loss = oracle.dice_ce_loss(oracle.nas_forward(store, x.to(dev).to(dtype))[-1], y.to(dev)); loss.backward()
    return loss.item(), {n: store[n].grad.detach().double().cpu() for n in names}
truth = run_oracle('cuda:0', torch.float64)
a = run_oracle('cpu', torch.float32)
b = run_oracle('cuda:0', torch.float32)
mg = m.to('cuda:0'); mg.train()
loss = oracle.dice_ce_loss(mg(x.to('cuda:0'))[-1], y.to('cuda:0')); loss.backward()
c = (loss.item(), {n: p.grad.detach().double().cpu() for n, p in mg.named_parameters()})
def rel(u, v):  # global L2 relative error over a set of tensors
    num = sum(((u[n] - v[n]) ** 2).sum() for n in v); den = sum((v[n] ** 2).sum() for n in v)
    return (num / den).sqrt().item()
arch = ('alphas_dn', 'alphas_up', 'alphas_dn_nm', 'alphas_up_nm', 'betas_dn', 'betas_up', 'gamma')
mixed = [n for n in names if '._ops.' in n]
other = [n for n in names if n not in mixed and n not in arch]
print('loss  truth %.9f  cpu32 %.9f  gpu32 %.9f  ours %.9f' % (truth[0], a[0], b[0], c[0]))
for label, r in (('oracle fp32 cpu', a), ('oracle fp32 gpu', b), ('senas_b200    ', c)):
    print(label, ' arch %.2e  mixedop-params %.2e  other-params %.2e' % (
        rel({n: r[1][n] for n in arch}, {n: truth[1][n] for n in arch}),
        rel({n: r[1][n] for n in mixed}, {n: truth[1][n] for n in mixed}),
        rel({n: r[1][n] for n in other}, {n: truth[1][n] for n in other})),
        ' '.join('%s %.1e' % (n, rel({n: r[1][n]}, {n: truth[1][n]})) for n in arch))
