"""Noise floor: the oracle (pure torch) evaluated on CPU vs on the GPU (cuDNN, TF32 off) for the full supernet."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'oracle')]
import torch
import senas_b200, senas_oracle as oracle
from helpers import max_err
senas_b200.exact_fp32()
B, H = int(sys.argv[1]) if len(sys.argv) > 1 else 2, int(sys.argv[2]) if len(sys.argv) > 2 else 64
torch.manual_seed(0)
m = senas_b200.NAS(1, 32, 2, depth=5, meta_node_num=3, use_sharing=False, double_down_channel=False, supervision=False)
gen = torch.Generator().manual_seed(1234)
x = torch.randn(B, 1, H, H, generator=gen); y = (torch.rand(B, H, H, generator=gen) > 0.8).long()
res = {}
for dev in ('cpu', 'cuda:0'):
    store = {k: v.to(dev) for k, v in oracle.clone_store(m.state_dict()).items()}
    for k, v in store.items():
        if v.is_floating_point() and not k.endswith(('running_mean', 'running_var')): v.requires_grad_(True)
    rec = []
    ocell = oracle.cell
    def cell_rec(p, *a, **k):
        o = ocell(p, *a, **k); o.retain_grad(); rec.append((p.prefix, o)); return o
    oracle.cell = cell_rec
    loss = oracle.dice_ce_loss(oracle.nas_forward(store, x.to(dev))[-1], y.to(dev)); loss.backward()
    oracle.cell = ocell
    res[dev] = (loss.item(), rec, store)
print('loss', res['cpu'][0], res['cuda:0'][0])
for (pref, a), (_, b) in zip(res['cpu'][1], res['cuda:0'][1]):
    print(f'{pref:28s} out {max_err(b, a.detach()):.1e}  grad_out {max_err(b.grad, a.grad):.1e}')
for n in ('alphas_dn', 'alphas_up', 'alphas_dn_nm', 'alphas_up_nm', 'betas_dn', 'betas_up', 'gamma'):
    print(n, f"{max_err(res['cuda:0'][2][n].grad, res['cpu'][2][n].grad):.1e}")
