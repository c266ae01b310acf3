"""Ad-hoc GPU diagnostics: per-tensor error tables for the edge cases (development aid)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'oracle')]
import numpy as np, torch
import senas_b200, senas_oracle as oracle
from helpers import OP_BY_ID, OP_NAME, golden, max_err
senas_b200.exact_fp32()
DEV = 'cuda:0'

def mixed_case(op_id, c_in, B, H, W):
    torch.manual_seed(100 + op_id + c_in + H)
    m = senas_b200.MixedOp(c_in, 8, OP_BY_ID[op_id]); m.apply(senas_b200.weights_init)
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.weight.data.uniform_(0.5, 1.5); mod.bias.data.normal_(0, 0.3)
    store = oracle.clone_store(m.state_dict())
    x = torch.randn(B, c_in, H, W); alpha = torch.softmax(torch.randn(6), -1)
    xo, ao = x.clone().requires_grad_(True), alpha.clone().requires_grad_(True)
    ref = oracle.mixed_op(oracle.Params(store), OP_NAME[op_id], xo, ao, True)
    gout = torch.randn(ref.shape); ref.backward(gout)
    m = m.to(DEV)
    xg, ag = x.to(DEV).requires_grad_(True), alpha.to(DEV).requires_grad_(True)
    out = m(xg, ag, ag); out.backward(gout.to(DEV))
    print(f'--- mixed op={op_id} c_in={c_in} B={B} {H}x{W} -> {tuple(ref.shape)}')
    print('out', max_err(out, ref.detach()), 'gx', max_err(xg.grad, xo.grad), 'galpha', max_err(ag.grad, ao.grad))
    for n, p in m.named_parameters():
        e = max_err(p.grad, store[n].grad)
        if e > 1e-5: print('   ', n, f'{e:.2e}', p.grad.abs().max().item(), store[n].grad.abs().max().item())

def search_case():
    g = golden('nas_search_2steps')
    B, H, seed, steps = [int(v) for v in g['meta']]
    torch.manual_seed(seed)
    m = senas_b200.NAS(1, 32, 2, depth=5, meta_node_num=3, use_sharing=False, double_down_channel=False, supervision=False).to(DEV)
    w_opt = torch.optim.SGD(m.parameters(), lr=5e-3, momentum=0.9, weight_decay=3e-4)
    a_opt = torch.optim.Adam(m.arch_parameters(), lr=1e-4, betas=(0.5, 0.999), weight_decay=1e-3)
    crit = lambda outs, y: oracle.dice_ce_loss(outs[-1], y)
    arch = senas_b200.Architecture(m, a_opt, crit)
    gen = torch.Generator().manual_seed(1234)
    names = ('alphas_dn', 'alphas_up', 'alphas_dn_nm', 'alphas_up_nm', 'betas_dn', 'betas_up', 'gamma')
    losses = []
    for s in range(steps):
        xt = torch.randn(B, 1, H, H, generator=gen).to(DEV); yt = (torch.rand(B, H, H, generator=gen) > 0.8).long().to(DEV)
        xv = torch.randn(B, 1, H, H, generator=gen).to(DEV); yv = (torch.rand(B, H, H, generator=gen) > 0.8).long().to(DEV)
        arch.step(xv, yv)
        if s == 0:
            for n in names: print('archgrad', n, f'{max_err(getattr(m, n).grad, g["archgrad." + n]):.2e}')
        w_opt.zero_grad(); loss = crit(m(xt), yt); losses.append(loss.item()); loss.backward()
        torch.nn.utils.clip_grad_norm_(m.parameters(), 5); w_opt.step()
    print('losses', losses, list(g['losses']))
    for n in names: print('arch', n, (getattr(m, n).detach().cpu() - torch.from_numpy(g['arch.' + n])).abs().max().item())
    print(repr(m.genotype())); print(str(g['genotype'])); print('genotype equal', repr(m.genotype()) == str(g['genotype']))

def repro_case():
    torch.manual_seed(3)
    c = senas_b200.Cell(3, 1, 32, 32, 32, 'up').to(DEV)
    in0, in1 = torch.randn(2, 32, 32, 32, device=DEV), torch.randn(2, 32, 16, 16, device=DEV)
    wn, wc = torch.softmax(torch.randn(9, 6, device=DEV), -1), torch.softmax(torch.randn(9, 6, device=DEV), -1)
    b = torch.softmax(torch.randn(9, device=DEV), -1)
    outs = []
    for _ in range(3):
        c.zero_grad()
        a, bb = in0.clone().requires_grad_(True), in1.clone().requires_grad_(True)
        wn_, b_ = wn.clone().requires_grad_(True), b.clone().requires_grad_(True)
        o = c.nodes(a, bb, wn_, wc, b_)
        o.backward(torch.sin(torch.arange(o.numel(), device=DEV, dtype=torch.float32)).view_as(o))
        outs.append([o.detach().clone(), a.grad.clone(), bb.grad.clone(), wn_.grad.clone(), b_.grad.clone()] + [p.grad.clone() for p in c._ops.parameters()])
    names = ['out', 'gin0', 'gin1', 'gwn', 'gbeta'] + [n for n, _ in c._ops.named_parameters()]
    for i, n in enumerate(names):
        if not (torch.equal(outs[0][i], outs[1][i]) and torch.equal(outs[0][i], outs[2][i])):
            print('NOT reproducible:', n, (outs[0][i] - outs[1][i]).abs().max().item())
    print('repro check done')

if __name__ == '__main__':
    mixed_case(1, 32, 1, 1, 1); mixed_case(2, 32, 1, 1, 3); mixed_case(3, 8, 2, 8, 8)
    repro_case(); search_case()
