"""The oracle (oracle/senas_oracle.py) against every golden fixture produced by the unmodified
reference (tests/golden/make_golden.py).  CPU only.  This is what pins the oracle."""
import numpy as np
import pytest
import torch

import senas_oracle as oracle
from helpers import OP_NAME, golden, golden_names, sub

TOL = 1e-6


def close(a, b, tol=TOL):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return (a - b).abs().max().item() <= tol * max(1.0, b.abs().max().item())


@pytest.mark.parametrize('name', golden_names('mixed_'))
def test_mixed_op_matches_reference(name):
    g = golden(name)
    c_in, B, H, W, training, op_id = [int(v) for v in g['meta']]
    store = oracle.clone_store(sub(g, 'state.'))
    x = torch.from_numpy(g['x']).requires_grad_(True)
    alpha = torch.from_numpy(g['alpha']).requires_grad_(True)
    out = oracle.mixed_op(oracle.Params(store), OP_NAME[op_id], x, alpha, bool(training))
    assert close(out, g['out'])
    if training:
        out.backward(torch.from_numpy(g['gout']))
        assert close(x.grad, g['gx']) and close(alpha.grad, g['galpha'])
        for k, v in sub(g, 'grad.').items():
            assert close(store[k].grad, v), k
        for k, v in sub(g, 'after.').items():
            assert close(store[k], v), k


@pytest.mark.parametrize('name,cell_type', [('cell_down', 'down'), ('cell_up', 'up')])
def test_cell_matches_reference(name, cell_type):
    g = golden(name)
    store = oracle.clone_store(sub(g, 'state.'))
    t = {k: torch.from_numpy(g[k]).requires_grad_(True) for k in ('in0', 'in1', 'wn', 'wc', 'betas')}
    out = oracle.cell(oracle.Params(store), cell_type, t['in0'], t['in1'], t['wn'], t['wc'], t['betas'])
    assert close(out, g['out'])
    out.backward(torch.from_numpy(g['gout']))
    for a, b in (('in0', 'gin0'), ('in1', 'gin1'), ('wn', 'gwn'), ('wc', 'gwc'), ('betas', 'gbetas')):
        assert close(t[a].grad, g[b]), a
    for k, v in sub(g, 'grad.').items():
        assert close(store[k].grad, v), k
    for k, v in sub(g, 'after.').items():
        assert close(store[k], v), k


@pytest.mark.parametrize('name,cell_type', [('cell_down_eval', 'down'), ('cell_up_eval', 'up')])
def test_cell_eval_matches_reference(name, cell_type):
    """infer() path (experiments/search_arc.py:301-330): running statistics, forward only."""
    g = golden(name)
    store = oracle.clone_store(sub(g, 'state.'), requires_grad=False)
    t = {k: torch.from_numpy(g[k]) for k in ('in0', 'in1', 'wn', 'wc', 'betas')}
    with torch.no_grad():
        out = oracle.cell(oracle.Params(store), cell_type, t['in0'], t['in1'], t['wn'], t['wc'], t['betas'],
                          training=False)
    assert close(out, g['out'])


def randomise_like_golden(m, seed):
    """Same draws, in the same ``modules()`` order, as tests/golden/make_golden.py:nas_eval_case."""
    gen = torch.Generator().manual_seed(seed + 1)
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.weight.data = 0.5 + torch.rand(mod.weight.shape, generator=gen)
            mod.bias.data = 0.3 * torch.randn(mod.bias.shape, generator=gen)
            mod.running_mean.data = 0.2 * torch.randn(mod.running_mean.shape, generator=gen)
            mod.running_var.data = 0.5 + torch.rand(mod.running_var.shape, generator=gen)
    with torch.no_grad():
        for n in ('alphas_dn', 'alphas_up', 'alphas_dn_nm', 'alphas_up_nm', 'betas_dn', 'betas_up', 'gamma'):
            getattr(m, n).copy_(0.5 * torch.randn(getattr(m, n).shape, generator=gen))
    return gen


def test_nas_eval_matches_reference():
    """``NAS.eval()`` forward of the fixed-seed supernet with non-trivial running statistics (nas_eval.npz)."""
    import senas_b200
    g = golden('nas_eval')
    B, H, seed = [int(v) for v in g['meta']]
    torch.manual_seed(seed)
    m = senas_b200.NAS(1, 32, 2, depth=5, meta_node_num=3, use_sharing=False, double_down_channel=False,
                       supervision=False)
    gen = randomise_like_golden(m, seed)
    x = torch.randn(B, 1, H, H, generator=gen)
    assert torch.equal(x, torch.from_numpy(g['x']))  # same generator stream as the reference run => same parameters
    with torch.no_grad():
        out = oracle.nas_forward({k: v for k, v in m.state_dict().items()}, x, training=False)[-1]
    assert close(out, g['out'], 1e-5)


def test_fixed_seed_search_matches_reference():
    """Two search steps (arch step + weight step) of the whole supernet from seed 0: same losses, same
    arch gradients, same alpha tables, same genotype as the reference (nas_search_2steps.npz)."""
    import senas_b200
    g = golden('nas_search_2steps')
    B, H, seed, steps = [int(v) for v in g['meta']]
    torch.manual_seed(seed)
    m = senas_b200.NAS(1, 32, 2, depth=5, meta_node_num=3, use_sharing=False, double_down_channel=False,
                       supervision=False)
    store = {k: v for k, v in m.state_dict().items()}
    names = [n for n, _ in m.named_parameters()]
    params = [store[n].requires_grad_(True) for n in names]
    arch_names = ['alphas_dn', 'alphas_up', 'alphas_dn_nm', 'alphas_up_nm', 'betas_dn', 'betas_up', 'gamma']
    w_opt = torch.optim.SGD(params, lr=5e-3, momentum=0.9, weight_decay=3e-4)
    a_opt = torch.optim.Adam([store[n] for n in arch_names], lr=1e-4, betas=(0.5, 0.999), weight_decay=1e-3)
    gen = torch.Generator().manual_seed(1234)
    losses = []
    for s in range(steps):
        xt = torch.randn(B, 1, H, H, generator=gen)
        yt = (torch.rand(B, H, H, generator=gen) > 0.8).long()
        xv = torch.randn(B, 1, H, H, generator=gen)
        yv = (torch.rand(B, H, H, generator=gen) > 0.8).long()
        a_opt.zero_grad()
        oracle.dice_ce_loss(oracle.nas_forward(store, xv)[-1], yv).backward()
        a_opt.step()
        if s == 0:
            for n in arch_names:
                assert close(store[n].grad, g['archgrad.' + n], 1e-5), n
        w_opt.zero_grad()
        loss = oracle.dice_ce_loss(oracle.nas_forward(store, xt)[-1], yt)
        losses.append(loss.item())
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 5)
        w_opt.step()
    assert np.allclose(losses, g['losses'], rtol=1e-6)
    for n in arch_names:
        assert close(store[n], g['arch.' + n], 1e-6), n
    assert repr(oracle.genotype({k: v.detach() for k, v in store.items()})) == str(g['genotype'])


def test_tap_geometry_matches_torch():
    """Appendix A of SURVEY.md: the gather form used by the kernels, written out in numpy, against
    F.conv2d / F.conv_transpose2d for every (k, dilation, OpType) on the path."""
    rng = np.random.default_rng(0)
    for k, dil in ((3, 1), (5, 2), (5, 3), (5, 1)):
        pad = (k // 2) * dil
        x = rng.standard_normal((1, 1, 7, 6)).astype(np.float32)
        w = rng.standard_normal((1, 1, k, k)).astype(np.float32)
        xt, wt = torch.from_numpy(x), torch.from_numpy(w)
        for op in ('NORM', 'DOWN', 'UP'):
            ref = oracle.conv(xt, wt, k, dil, op).numpy()[0, 0]
            H, W = x.shape[2:]
            out = np.zeros_like(ref)
            for oy in range(ref.shape[0]):
                for ox in range(ref.shape[1]):
                    acc = 0.0
                    for ky in range(k):
                        for kx in range(k):
                            if op == 'UP':
                                ny, nx = oy + pad - dil * ky, ox + pad - dil * kx
                                if ny % 2 or nx % 2:
                                    continue
                                iy, ix = ny // 2, nx // 2
                            else:
                                s = 1 if op == 'NORM' else 2
                                iy, ix = s * oy + dil * (ky - k // 2), s * ox + dil * (kx - k // 2)
                            if 0 <= iy < H and 0 <= ix < W:
                                acc += x[0, 0, iy, ix] * w[0, 0, ky, kx]
                    out[oy, ox] = acc
            assert np.allclose(out, ref, atol=1e-5), (k, dil, op)
