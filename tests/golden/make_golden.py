"""Generate the golden fixtures from the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py        # needs /root/reference (or $SENAS_REF)

For every MixedOp flavour on the search path (OpType x c_in), for a down and
an up Cell, and for the whole NAS supernet on a small input, the reference modules are run forward
and backward on seeded inputs and everything needed to replay the case is stored as float32 .npz:
inputs, the full state dict before the step, outputs, input/alpha/beta/parameter gradients and the
BatchNorm buffers after the step.  The oracle (oracle/senas_oracle.py) and the CUDA path are both
checked against these files; the files are small and committed, the reference is not.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import ref_shim  # noqa: E402

cell_mod, ss_mod, ops_mod = ref_shim.load()
from utils.loss import SegmentationLosses  # noqa: E402  (reference package)
from utils.utils import weights_init  # noqa: E402


def npd(d):
    return {k: v.detach().cpu().numpy() for k, v in d.items()}


def save(name, **arrays):
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **arrays)
    print(name, f'{os.path.getsize(path) / 1024:.0f} KiB')


def randomise_bn(mod, gen):
    """Non-trivial gamma/beta/running stats so that every BN term of the math is exercised."""
    for m in mod.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.data = 0.5 + torch.rand(m.weight.shape, generator=gen)
            m.bias.data = 0.3 * torch.randn(m.bias.shape, generator=gen)
            m.running_mean.data = 0.2 * torch.randn(m.running_mean.shape, generator=gen)
            m.running_var.data = 0.5 + torch.rand(m.running_var.shape, generator=gen)


def mixed_case(tag, op_type, c_in, B, H, W, seed, training=True):
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    m = cell_mod.MixedOp(c_in, 8, op_type)
    m.apply(weights_init)
    randomise_bn(m, gen)
    m.train(training)
    x = torch.randn(B, c_in, H, W, generator=gen, requires_grad=True)
    alpha = F.softmax(torch.randn(6, generator=gen), -1).requires_grad_(True)
    state = {k: v.clone() for k, v in m.state_dict().items()}
    out = m(x, alpha, alpha)
    arrays = {'x': x.detach().numpy(), 'alpha': alpha.detach().numpy(), 'out': out.detach().numpy(),
              'meta': np.array([c_in, B, H, W, int(training), op_type.value['id']])}
    if training:
        gout = torch.randn(out.shape, generator=gen)
        out.backward(gout)
        arrays.update(gout=gout.numpy(), gx=x.grad.numpy(), galpha=alpha.grad.numpy())
        arrays.update({'grad.' + k: v.grad.numpy() for k, v in m.named_parameters()})
        arrays.update({'after.' + k: v.numpy() for k, v in m.state_dict().items() if 'running' in k or 'num_batches' in k})
    arrays.update({'state.' + k: v.numpy() for k, v in state.items()})
    save(tag, **arrays)


def cell_case(tag, cell_type, B, H, seed, training=True):
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    c = cell_mod.Cell(3, 1, 32, 32, 32, cell_type)
    c.apply(weights_init)
    randomise_bn(c, gen)
    c.train(training)
    # in0 enters at twice in1's size for both cell types (down: rectified by preprocess0; up: skip input)
    in0 = torch.randn(B, 32, 2 * H, 2 * H, generator=gen, requires_grad=True)
    in1 = torch.randn(B, 32, H, H, generator=gen, requires_grad=True)
    wn = F.softmax(torch.randn(9, 6, generator=gen), -1).requires_grad_(True)
    wc = F.softmax(torch.randn(9, 6, generator=gen), -1).requires_grad_(True)
    betas = F.softmax(torch.randn(9, generator=gen), -1).requires_grad_(True)
    state = {k: v.clone() for k, v in c.state_dict().items()}
    out = c(in0, in1, wn, wc, betas)
    if not training:  # infer() path (experiments/search_arc.py:301-330): running statistics, forward only
        arrays = {'in0': in0.detach().numpy(), 'in1': in1.detach().numpy(), 'wn': wn.detach().numpy(),
                  'wc': wc.detach().numpy(), 'betas': betas.detach().numpy(), 'out': out.detach().numpy()}
        arrays.update({'state.' + k: v.numpy() for k, v in state.items()})
        save(tag, **arrays)
        return
    gout = torch.randn(out.shape, generator=gen)
    out.backward(gout)
    arrays = {'in0': in0.detach().numpy(), 'in1': in1.detach().numpy(), 'wn': wn.detach().numpy(),
              'wc': wc.detach().numpy(), 'betas': betas.detach().numpy(), 'out': out.detach().numpy(),
              'gout': gout.numpy(), 'gin0': in0.grad.numpy(), 'gin1': in1.grad.numpy(), 'gwn': wn.grad.numpy(),
              'gwc': wc.grad.numpy(), 'gbetas': betas.grad.numpy()}
    arrays.update({'grad.' + k: v.grad.numpy() for k, v in c.named_parameters()})
    arrays.update({'state.' + k: v.numpy() for k, v in state.items()})
    arrays.update({'after.' + k: v.numpy() for k, v in c.state_dict().items() if 'running' in k})
    save(tag, **arrays)


def nas_case(tag, B, H, seed, steps):
    """Whole supernet, fixed seed: loss trajectory, arch gradients, genotype after `steps` search steps
    (arch step + weight step with the PROMISE12 optimisers, experiments/search_arc.py:252-293)."""
    dev = torch.device('cpu')
    torch.manual_seed(seed)
    m = ss_mod.NAS(1, 32, 2, depth=5, meta_node_num=3, use_sharing=False, double_down_channel=False,
                   supervision=False, device=dev)
    crit = SegmentationLosses('dice_ce')
    w_opt = torch.optim.SGD(m.parameters(), lr=5e-3, momentum=0.9, weight_decay=3e-4)
    a_opt = torch.optim.Adam(m.arch_parameters(), lr=1e-4, betas=(0.5, 0.999), weight_decay=1e-3)
    arch = ss_mod.Architecture(m, a_opt, crit)
    gen = torch.Generator().manual_seed(1234)
    losses, arrays = [], {}
    m.train()
    for s in range(steps):
        xt = torch.randn(B, 1, H, H, generator=gen)
        yt = (torch.rand(B, H, H, generator=gen) > 0.8).long()
        xv = torch.randn(B, 1, H, H, generator=gen)
        yv = (torch.rand(B, H, H, generator=gen) > 0.8).long()
        arch.step(xv, yv)
        if s == 0:
            arrays.update({'archgrad.' + k: getattr(m, k).grad.numpy().copy()
                           for k in ('alphas_dn', 'alphas_up', 'alphas_dn_nm', 'alphas_up_nm', 'betas_dn', 'betas_up', 'gamma')})
        w_opt.zero_grad()
        loss = crit(m(xt), yt)
        losses.append(loss.item())
        loss.backward()
        torch.nn.utils.clip_grad_norm_(m.parameters(), 5)
        w_opt.step()
    g = m.genotype()
    arrays.update(losses=np.array(losses), meta=np.array([B, H, seed, steps]),
                  genotype=np.array(repr(g)))
    arrays.update({'arch.' + k: getattr(m, k).detach().numpy() for k in
                   ('alphas_dn', 'alphas_up', 'alphas_dn_nm', 'alphas_up_nm', 'betas_dn', 'betas_up', 'gamma')})
    save(tag, **arrays)


def nas_eval_case(tag, B, H, seed):
    """``NAS.eval()`` forward (the infer() pass of experiments/search_arc.py:301-330) of the fixed-seed supernet with
    non-trivial BatchNorm parameters / running statistics drawn from a seeded generator in ``modules()`` order (the
    mirror model applies the same function: identical module order, tested through the identical state-dict keys), so
    only the input and the logits need to be stored."""
    torch.manual_seed(seed)
    m = ss_mod.NAS(1, 32, 2, depth=5, meta_node_num=3, use_sharing=False, double_down_channel=False,
                   supervision=False, device=torch.device('cpu'))
    gen = torch.Generator().manual_seed(seed + 1)
    randomise_bn(m, gen)
    with torch.no_grad():
        for n in ('alphas_dn', 'alphas_up', 'alphas_dn_nm', 'alphas_up_nm', 'betas_dn', 'betas_up', 'gamma'):
            getattr(m, n).copy_(0.5 * torch.randn(getattr(m, n).shape, generator=gen))
    m.eval()
    x = torch.randn(B, 1, H, H, generator=gen)
    with torch.no_grad():
        out = m(x)[-1]
    save(tag, x=x.numpy(), out=out.numpy(), meta=np.array([B, H, seed]))


if __name__ == '__main__':
    OT = ops_mod.OpType
    if len(sys.argv) > 1 and sys.argv[1] == 'r2':  # round-2 additions only (eval-mode cases)
        mixed_case('mixed_down32_eval', OT.DOWN, 32, 2, 10, 12, seed=31, training=False)
        mixed_case('mixed_norm8_eval', OT.NORM, 8, 2, 7, 9, seed=32, training=False)
        cell_case('cell_up_eval', 'up', 2, 6, seed=33, training=False)
        cell_case('cell_down_eval', 'down', 2, 6, seed=34, training=False)
        nas_eval_case('nas_eval', 2, 64, seed=0)
        sys.exit(0)
    mixed_case('mixed_norm32', OT.NORM, 32, 2, 12, 20, seed=11)
    mixed_case('mixed_norm8', OT.NORM, 8, 3, 9, 18, seed=12)
    mixed_case('mixed_down32', OT.DOWN, 32, 2, 14, 18, seed=13)
    mixed_case('mixed_down32_odd', OT.DOWN, 32, 2, 11, 9, seed=14)
    mixed_case('mixed_up32', OT.UP, 32, 2, 7, 10, seed=15)
    mixed_case('mixed_norm32_eval', OT.NORM, 32, 2, 8, 8, seed=16, training=False)
    mixed_case('mixed_up32_eval', OT.UP, 32, 1, 5, 6, seed=17, training=False)
    cell_case('cell_down', 'down', 2, 8, seed=21)
    cell_case('cell_up', 'up', 2, 8, seed=22)
    nas_case('nas_search_2steps', 2, 64, seed=0, steps=2)
    mixed_case('mixed_down32_eval', OT.DOWN, 32, 2, 10, 12, seed=31, training=False)
    mixed_case('mixed_norm8_eval', OT.NORM, 8, 2, 7, 9, seed=32, training=False)
    cell_case('cell_up_eval', 'up', 2, 6, seed=33, training=False)
    cell_case('cell_down_eval', 'down', 2, 6, seed=34, training=False)
    nas_eval_case('nas_eval', 2, 64, seed=0)
