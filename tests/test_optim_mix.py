"""the flat-buffer optimizer kernels and the gamma-mix concat of libsenas_b200 against the PyTorch ops
they replace (torch.nn.utils.clip_grad_norm_ + torch.optim.SGD, torch.optim.Adam, lerp + cat and their autograd).
CPU part: the same kernel sources on the test-only emulator (tests/emu); GPU part: the product library."""
import copy
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'emu'))


def _ptr(t):
    return t.data_ptr()


def _sgd_case(lib, dev, n, max_norm, steps=3, stream=0):
    torch.manual_seed(n)
    p0 = torch.randn(n, device=dev)
    grads = [torch.randn(n, device=dev) * (3.0 if max_norm > 0 else 1.0) for _ in range(steps)]
    # reference: the PyTorch ops on a parameter split into ragged tensors
    sizes, left = [], n
    while left:
        s = min(left, 1 + (len(sizes) * 37) % 501)
        sizes.append(s)
        left -= s
    ps = [torch.nn.Parameter(t.clone()) for t in p0.split(sizes)]
    opt = torch.optim.SGD(ps, lr=5e-3, momentum=0.9, weight_decay=3e-4)
    p, m = p0.clone(), torch.zeros(n, device=dev)
    lr = torch.tensor(5e-3, device=dev)
    scratch, norm = torch.zeros(2048, device=dev), torch.zeros((), device=dev)
    for it, g in enumerate(grads):
        for q, gq in zip(ps, g.split(sizes)):
            q.grad = gq.clone()
        want_norm = torch.nn.utils.clip_grad_norm_(ps, max_norm) if max_norm > 0 else None
        opt.step()
        gg = g.clone()
        assert lib.senas_sgd_clip_step(_ptr(p), _ptr(gg), _ptr(m), n, _ptr(lr), 0.9, 3e-4, float(max_norm), _ptr(scratch),
                                       _ptr(norm), stream) == 0, lib.senas_last_error()
        if dev != 'cpu':
            torch.cuda.synchronize()
        want = torch.cat([q.detach() for q in ps])
        assert torch.allclose(p, want, rtol=2e-6, atol=1e-7), (it, (p - want).abs().max().item())
        if max_norm > 0:
            assert abs(norm.item() - want_norm.item()) <= 1e-5 * want_norm.item()
            want_g = torch.cat([q.grad for q in ps])
            assert torch.allclose(gg, want_g, rtol=2e-6, atol=1e-8)  # clipped in place, like clip_grad_norm_
        if it == 1:  # a scheduler changes the learning rate: a device scalar, nothing else
            lr.fill_(2.5e-3)
            opt.param_groups[0]['lr'] = 2.5e-3


def _adam_case(lib, dev, n, stream=0):
    torch.manual_seed(5)
    p0 = torch.randn(n, device=dev) * 1e-3
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=1e-4, betas=(0.5, 0.999), weight_decay=1e-3)
    p, ea, es = p0.clone(), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    step, lr = torch.zeros((), device=dev), torch.tensor(1e-4, device=dev)
    for it in range(4):
        g = torch.randn(n, device=dev) * 10 ** (-it)
        ref.grad = g.clone()
        opt.step()
        assert lib.senas_adam_step(_ptr(p), _ptr(g), _ptr(ea), _ptr(es), _ptr(step), n, _ptr(lr), 0.5, 0.999, 1e-8, 1e-3,
                                   stream) == 0, lib.senas_last_error()
        if dev != 'cpu':
            torch.cuda.synchronize()
        assert step.item() == it + 1
        # Adam's first steps are +-lr whatever the gradient: compare the UPDATE, relative to lr
        assert (p - ref.detach()).abs().max().item() <= 2e-5 * 1e-4 * (it + 1) + 1e-10, (it, (p - ref.detach()).abs().max().item())


def _mix_ref(gamma, gidx, ts):
    parts = [ts[0]]
    for k, g in enumerate(gidx):
        parts.append(ts[k] * gamma[g][0] + ts[k + 1] * gamma[g][1])
    return torch.cat(parts, 1)


def _mix_case_raw(lib, dev, n_t, B, C, H, W, stream=0):
    """The three mix kernels through the raw C ABI (NHWC buffers) against autograd of the PyTorch expression."""
    torch.manual_seed(n_t * 100 + H)
    gidx = [(2 * k + 1) % 6 for k in range(n_t - 1)]  # distinct rows, as in the supernet
    gamma = torch.softmax(torch.randn(6, 2, device=dev), -1).requires_grad_(True)
    ts = [torch.randn(B, C, H, W, device=dev).requires_grad_(True) for _ in range(n_t)]
    ref = _mix_ref(gamma, gidx, ts)
    gout = torch.randn_like(ref)
    ref.backward(gout)
    nh = [t.detach().permute(0, 2, 3, 1).contiguous() for t in ts]  # NHWC
    npix = B * H * W
    out = torch.zeros(B, H, W, C * n_t, device=dev)
    gam = gamma.detach().contiguous()
    for s in range(n_t):
        if s == 0:
            rc = lib.senas_mix_forward(_ptr(nh[0]), C, None, 0, None, _ptr(out), C * n_t, 0, C, npix, stream)
        else:
            rc = lib.senas_mix_forward(_ptr(nh[s - 1]), C, _ptr(nh[s]), C, _ptr(gam) + 8 * gidx[s - 1], _ptr(out), C * n_t, C * s,
                                       C, npix, stream)
        assert rc == 0, lib.senas_last_error()
    assert torch.allclose(out.permute(0, 3, 1, 2), ref.detach(), rtol=1e-6, atol=1e-6)
    g = gout.permute(0, 2, 3, 1).contiguous()
    dgamma, scratch = torch.zeros(6, 2, device=dev), torch.zeros(1184, device=dev)
    for s in range(1, n_t):
        assert lib.senas_mix_backward(_ptr(nh[s - 1]), C, _ptr(nh[s]), C, _ptr(gam) + 8 * gidx[s - 1], _ptr(g), C * n_t, C * s, C,
                                      npix, None, None, _ptr(dgamma) + 8 * gidx[s - 1], _ptr(scratch), stream) == 0
    for k in range(n_t):
        d = torch.zeros(B, H, W, C, device=dev)
        w0, o0 = (None, 0) if k == 0 else (None, -1)
        w1, o1 = (_ptr(gam) + 8 * gidx[k], C * (k + 1)) if k + 1 < n_t else (None, -1)
        w2, o2 = (_ptr(gam) + 8 * gidx[k - 1] + 4, C * k) if k >= 1 else (None, -1)
        assert lib.senas_mix_dx(_ptr(g), C * n_t, w0, o0, w1, o1, w2, o2, _ptr(d), C, npix, stream) == 0
        assert torch.allclose(d.permute(0, 3, 1, 2), ts[k].grad, rtol=1e-5, atol=1e-6), k
    if dev != 'cpu':
        torch.cuda.synchronize()
    assert torch.allclose(dgamma, gamma.grad, rtol=1e-4, atol=1e-4 * gamma.grad.abs().max().item())


# ------------------------------------------------------------------------------------------------- CPU (emulator)
@pytest.fixture(scope='module')
def emu():
    import build_emu
    from senas_b200 import _lib
    return _lib.bind(build_emu.build())


@pytest.mark.parametrize('n,max_norm', [(4096, 5.0), (10007, 5.0), (777, 0.0), (3, 5.0)])
def test_sgd_clip_step_emulated(emu, n, max_norm):
    _sgd_case(emu, 'cpu', n, max_norm)


def test_adam_step_emulated(emu):
    _adam_case(emu, 'cpu', 246)


@pytest.mark.parametrize('n_t,B,C,H,W', [(2, 1, 32, 3, 5), (4, 2, 32, 4, 4), (3, 1, 8, 2, 7)])
def test_mix_concat_emulated(emu, n_t, B, C, H, W):
    _mix_case_raw(emu, 'cpu', n_t, B, C, H, W)


def test_fused_search_optim_arena_emulated(emu):
    """FusedSearchOptim on a (CPU) supernet with the emulated kernels: parameters re-homed into the arena keep their
    values / names / shapes, every fused cell's gradient slice has the layout of its graph's flat gradient buffer, and
    Adam on the arch slice followed by clip + SGD on the whole arena equals torch.optim step for step."""
    import senas_b200
    torch.manual_seed(0)

    def make():
        torch.manual_seed(0)
        m = senas_b200.NAS(1, 32, 2, depth=3, meta_node_num=3, use_sharing=False, double_down_channel=False, supervision=False)
        return (m, torch.optim.SGD(m.parameters(), lr=5e-3, momentum=0.9, weight_decay=3e-4),
                torch.optim.Adam(m.arch_parameters(), lr=1e-4, betas=(0.5, 0.999), weight_decay=1e-3))
    ref, w_ref, a_ref = make()
    m, w, a = make()
    before = {k: v.clone() for k, v in m.state_dict().items()}
    fopt = senas_b200.FusedSearchOptim(m, w, a, grad_clip=5.0, lib=emu)
    assert list(m.state_dict().keys()) == list(before.keys())
    for k, v in m.state_dict().items():
        assert torch.equal(v, before[k]), k
    lo, hi = fopt.flat_p.data_ptr(), fopt.flat_p.data_ptr() + 4 * fopt.n
    assert all(lo <= p.data_ptr() < hi for p in m.parameters())
    assert len(fopt.runners) == 2 + 3 + 1  # depth 3: two down cells, the triangle of 2 + 1 up cells, the head cell
    for r, (o, n) in zip(fopt.runners, fopt.runner_slices):
        assert r.grad_buffer.data_ptr() == fopt.flat_g.data_ptr() + 4 * o and n == r.grad_floats
        off = o
        for p, sz in zip(r.params, r.sizes):  # the arena order of a cell IS its graph's gradient order
            assert fopt.layout[id(p)] == (off, sz) and p.data_ptr() == fopt.flat_p.data_ptr() + 4 * off
            off += sz
    gen = torch.Generator().manual_seed(1)
    for it in range(3):
        # architecture step: gradients of the arch parameters only
        for p_ref, p in zip(ref.arch_parameters(), m.arch_parameters()):
            g = torch.randn(p.shape, generator=gen) * 1e-2
            p_ref.grad = g.clone()
            fopt.grad_views[id(p)].copy_(g)
        a_ref.step()
        fopt.adam_step()
        # weight step: gradients of everything (the arch parameters are in model.parameters() too, search_arc.py)
        for p_ref, p in zip(ref.parameters(), m.parameters()):
            g = torch.randn(p.shape, generator=gen) * 0.5
            p_ref.grad = g.clone()
            fopt.grad_views[id(p)].copy_(g)
        torch.nn.utils.clip_grad_norm_(ref.parameters(), 5.0)
        w_ref.step()
        fopt.sgd_step()
        if it == 0:
            w_ref.param_groups[0]['lr'] = w.param_groups[0]['lr'] = 2.5e-3
            fopt.sync_lr()
    for (k, p_ref), p in zip(ref.named_parameters(), m.parameters()):
        assert torch.allclose(p, p_ref, rtol=1e-5, atol=2e-7), (k, (p - p_ref).abs().max().item())
    fopt.publish_state()
    sd = w.state_dict()
    assert len(sd['state']) == len(fopt.params)
    k0 = next(iter(w_ref.state_dict()['state']))
    assert torch.allclose(sd['state'][k0]['momentum_buffer'], w_ref.state_dict()['state'][k0]['momentum_buffer'], rtol=1e-5, atol=1e-7)
    assert float(a.state_dict()['state'][0]['step']) == 3.0


# ------------------------------------------------------------------------------------------------- GPU (product)
@pytest.mark.gpu
@pytest.mark.parametrize('n,max_norm', [(1 << 21, 5.0), (1970001, 5.0), (1003, 0.0)])
def test_sgd_clip_step_gpu(n, max_norm):
    from senas_b200 import _lib
    _sgd_case(_lib.get(), 'cuda', n, max_norm, stream=torch.cuda.current_stream().cuda_stream)


@pytest.mark.gpu
def test_adam_step_gpu():
    from senas_b200 import _lib
    _adam_case(_lib.get(), 'cuda', 246, stream=torch.cuda.current_stream().cuda_stream)


@pytest.mark.gpu
@pytest.mark.parametrize('n_t,B,H', [(2, 16, 128), (4, 4, 64), (3, 2, 16)])
def test_mix_concat_autograd_gpu(n_t, B, H):
    """senas_b200.supernet.mix_concat (the autograd wrapper the supernet calls) == lerp/cat, forward and backward."""
    from senas_b200.supernet import mix_concat
    torch.manual_seed(n_t + H)
    dev = 'cuda'
    gidx = [(2 * k + 1) % 6 for k in range(n_t - 1)]
    gl = torch.randn(6, 2, device=dev, requires_grad=True)
    ts = [torch.randn(B, 32, H, H, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
          for _ in range(n_t)]
    ref = _mix_ref(torch.softmax(gl, -1), gidx, ts)
    gout = torch.randn_like(ref)
    ref.backward(gout)
    want = [gl.grad.clone()] + [t.grad.clone() for t in ts]
    gl.grad = None
    for t in ts:
        t.grad = None
    out = mix_concat(torch.softmax(gl, -1), gidx, ts)
    assert out.shape == ref.shape and torch.allclose(out, ref.detach(), rtol=1e-6, atol=1e-6)
    out.backward(gout)
    for w, t in zip(want[1:], ts):
        assert torch.allclose(t.grad, w, rtol=1e-5, atol=1e-6)
    assert torch.allclose(gl.grad, want[0], rtol=1e-3, atol=1e-4 * want[0].abs().max().item())


@pytest.mark.gpu
def test_fused_search_optim_step_equals_stock_step():
    """GraphedSearchStep(fused_optim=True): arenas + senas_sgd_clip_step / senas_adam_step (+ the cells' gradients
    produced in the arena) against the same captured step with clip_grad_norm_ / SGD / Adam of PyTorch: same loss, same
    weights, same arch parameters after two steps; the optimizers' state_dict keeps working; lr changes need no re-capture."""
    import senas_b200
    from senas_b200.loss import SegmentationLosses
    senas_b200.exact_fp32()
    senas_b200.set_conv_mode('fp32')
    dev = 'cuda'
    B, H = 2, 64
    gen = torch.Generator().manual_seed(3)
    batch = [torch.randn(B, 1, H, H, generator=gen).to(dev), (torch.rand(B, H, H, generator=gen) > 0.8).long().to(dev),
             torch.randn(B, 1, H, H, generator=gen).to(dev), (torch.rand(B, H, H, generator=gen) > 0.8).long().to(dev)]
    crit = SegmentationLosses('dice_ce')

    def make():
        torch.manual_seed(0)
        m = senas_b200.NAS(1, 32, 2, depth=5, meta_node_num=3, use_sharing=False, double_down_channel=False,
                           supervision=False).to(dev).train()
        return (m, torch.optim.SGD(m.parameters(), lr=5e-3, momentum=0.9, weight_decay=3e-4),
                torch.optim.Adam(m.arch_parameters(), lr=1e-4, betas=(0.5, 0.999), weight_decay=1e-3))

    results = []
    for fused_optim in (False, True):
        m, w, a = make()
        init = {k: v.detach().clone() for k, v in m.state_dict().items()}
        step = senas_b200.GraphedSearchStep(m, crit, w, a, batch, grad_clip=5.0, warmup=2, fused_optim=fused_optim)
        for k, v in m.state_dict().items():
            assert torch.equal(v, init[k]), f'warm-up left a trace in {k}'
        losses = [step(*batch).item()]
        torch.cuda.synchronize()
        after1 = {k: v.detach().clone() for k, v in m.state_dict().items()}
        w.param_groups[0]['lr'] = 2.5e-3  # CosineAnnealingLR would
        ncap = len(step._variants)
        losses.append(step(*batch).item())
        torch.cuda.synchronize()
        if fused_optim:
            assert len(step._variants) == ncap and step.fopt is not None  # a device scalar: no re-capture
            sd = w.state_dict()
            assert len(sd['state']) == len(step.params)
            assert float(a.state_dict()['state'][0]['step']) == 2.0
        else:
            assert len(step._variants) == ncap  # (same key, re-captured in place)
        results.append((losses, after1, {k: v.detach().clone() for k, v in m.state_dict().items()}, init))
        step.release()
    (l0, a0, s0, init), (l1, a1, s1, _) = results
    assert abs(l0[0] - l1[0]) <= 2e-5 * abs(l0[0]) and abs(l0[1] - l1[1]) <= 1e-3 * abs(l0[1]), (l0, l1)
    # step 1: same gradients (same graph up to the optimizer), so the updates agree tensor by tensor
    worst = 0.0
    for k, v in a0.items():
        if not v.is_floating_point():
            assert torch.equal(v, a1[k]), k
            continue
        upd0, upd1 = (v - init[k]).double(), (a1[k] - init[k]).double()
        s = upd0.abs().max().item()
        if s > 1e-12 and not k.startswith(('alphas', 'betas', 'gamma')):
            ulp = 2 * 1.1920929e-07 * v.abs().max().item()  # (an update of ~150 ulps of a BatchNorm weight: allow 2 ulps)
            worst = max(worst, max((upd1 - upd0).abs().max().item() - ulp, 0.0) / s)
    assert worst <= 5e-3, worst  # (fp32 noise floor: cuDNN atomics, reduction order)
    for k in a0:
        if k.startswith(('alphas', 'betas', 'gamma')):  # Adam: +-lr per step whatever the gradient
            assert (a0[k] - a1[k]).abs().max().item() <= 2.5e-4, k
    # step 2 (momentum, changed lr): its gradients are those of a network that already differs at the noise level, and
    # ReLU-boundary flips amplify that tensor by tensor (tests/test_gpu_parity_r2.py) -- gate the update as a whole
    num = sum(((s1[k] - a1[k]).double() - (s0[k] - a0[k]).double()).pow(2).sum().item() for k in s0
              if s0[k].is_floating_point() and not k.startswith(('alphas', 'betas', 'gamma')) and 'running' not in k)
    den = sum((s0[k] - a0[k]).double().pow(2).sum().item() for k in s0
              if s0[k].is_floating_point() and not k.startswith(('alphas', 'betas', 'gamma')) and 'running' not in k)
    assert (num / den) ** 0.5 <= 5e-2, (num / den) ** 0.5  # (chaotic growth of a 1e-7 difference: see the floor in test_gpu_parity_r2)


# ------------------------------------------------------------------------------------------------- row f1: ConvBn blocks
@pytest.mark.gpu
@pytest.mark.parametrize('kind,c_in,B,H,W,training', [('shrink', 32, 4, 64, 64, True), ('shrink', 128, 2, 128, 128, True),
                                                      ('shrink', 96, 2, 64, 128, True), ('rectify', 24, 4, 64, 64, True),
                                                      ('rectify', 24, 2, 256, 256, True), ('shrink', 64, 2, 64, 64, False)])
def test_convbn_block_vs_pytorch(kind, c_in, B, H, W, training):
    """ShrinkBlock / RectifyBlock (utils/operations.py:206-232) through senas_convbn_forward/backward (tcgen05, bf16
    operands) against the same nn.Modules evaluated by PyTorch in fp32 (cuDNN, TF32 off).  Inputs and conv weights are
    bf16-representable, so the forward differs by accumulation order only (1e-4); in backward the only rounded tensor is
    dy, every gradient meets the 2e-2 gate of the bf16 mode (BASELINE.json north_star)."""
    import senas_b200
    from senas_b200 import ops
    senas_b200.exact_fp32()
    dev = 'cuda'
    torch.manual_seed(c_in + H)
    blk = (ops.ShrinkBlock(c_in, 32) if kind == 'shrink' else ops.RectifyBlock(c_in, 32)).to(dev)
    blk.apply(senas_b200.weights_init)
    with torch.no_grad():
        blk.conv.weight.copy_(blk.conv.weight.bfloat16().float())
        blk.norm.weight.uniform_(0.5, 1.5)
        blk.norm.bias.normal_(0, 0.3)
        blk.norm.running_mean.normal_(0, 0.1)
        blk.norm.running_var.uniform_(0.5, 2.0)
    blk.train(training)
    ref = copy.deepcopy(blk)
    x = torch.randn(B, c_in, H, W, device=dev).bfloat16().float().contiguous(memory_format=torch.channels_last)
    gout = torch.randn(B, 32, H, W, device=dev).contiguous(memory_format=torch.channels_last)
    xr = x.clone().requires_grad_(True)
    senas_b200.set_conv_mode('fp32')
    try:
        want = ref(xr)
        want.backward(gout)
        senas_b200.set_conv_mode('bf16')
        xg = x.clone().requires_grad_(True)
        lib = senas_b200._lib.get()
        n0 = lib.senas_launch_count()
        got = blk(xg)
        got.backward(gout)
        torch.cuda.synchronize()
        assert lib.senas_launch_count() > n0, 'the fused path did not run'
    finally:
        senas_b200.set_conv_mode('fp32')

    def rel(a, b):
        return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()
    assert rel(got, want.detach()) <= 1e-4, rel(got, want.detach())
    assert rel(xg.grad, xr.grad) <= 2e-2, rel(xg.grad, xr.grad)
    assert rel(blk.conv.weight.grad, ref.conv.weight.grad) <= 2e-2
    assert rel(blk.norm.weight.grad, ref.norm.weight.grad) <= 2e-2
    assert rel(blk.norm.bias.grad, ref.norm.bias.grad) <= 2e-2
    assert rel(blk.norm.running_mean, ref.norm.running_mean) <= 1e-4
    assert rel(blk.norm.running_var, ref.norm.running_var) <= 1e-4
    assert torch.equal(blk.norm.num_batches_tracked, ref.norm.num_batches_tracked)


@pytest.mark.gpu
def test_arch_grads_only_step_equals_full_backward_step():
    """GraphedSearchStep(arch_grads_only=True): the architecture step asks for the alpha / beta / gamma gradients only
    (torch.autograd.grad + senas_bwd_args_t.skip_wgrad: no weight-gradient kernel is launched in that pass).  The reference
    computes those weight gradients and zeroes them before any use (experiments/search_arc.py:268-271), so the model after
    the step is the same: arch parameters, weights, BatchNorm buffers."""
    import senas_b200
    from senas_b200.loss import SegmentationLosses
    senas_b200.exact_fp32()
    senas_b200.set_conv_mode('fp32')
    dev = 'cuda'
    B, H = 2, 64
    gen = torch.Generator().manual_seed(4)
    batch = [torch.randn(B, 1, H, H, generator=gen).to(dev), (torch.rand(B, H, H, generator=gen) > 0.8).long().to(dev),
             torch.randn(B, 1, H, H, generator=gen).to(dev), (torch.rand(B, H, H, generator=gen) > 0.8).long().to(dev)]
    crit = SegmentationLosses('dice_ce')
    lib = senas_b200._lib.get()
    out = []
    for only in (False, True):
        torch.manual_seed(0)
        m = senas_b200.NAS(1, 32, 2, depth=5, meta_node_num=3, use_sharing=False, double_down_channel=False,
                           supervision=False).to(dev).train()
        w = torch.optim.SGD(m.parameters(), lr=5e-3, momentum=0.9, weight_decay=3e-4)
        a = torch.optim.Adam(m.arch_parameters(), lr=1e-4, betas=(0.5, 0.999), weight_decay=1e-3)
        init = {k: v.detach().clone() for k, v in m.state_dict().items()}
        n0 = lib.senas_launch_count()
        step = senas_b200.GraphedSearchStep(m, crit, w, a, batch, grad_clip=5.0, warmup=1, fused_optim=True, arch_grads_only=only)
        launches = lib.senas_launch_count() - n0  # warm-up step + capture pass
        loss = step(*batch).item()
        torch.cuda.synchronize()
        out.append((loss, {k: v.detach().clone() for k, v in m.state_dict().items()}, init, launches))
        step.release()
    (l0, s0, init, n_full), (l1, s1, _, n_only) = out
    assert n_only < n_full - 200, (n_full, n_only)  # the weight-gradient launches of the arch pass are gone
    assert abs(l0 - l1) <= 2e-5 * abs(l0), (l0, l1)
    worst = 0.0
    for k, v in s0.items():
        if not v.is_floating_point():
            assert torch.equal(v, s1[k]), k
        elif k.startswith(('alphas', 'betas', 'gamma')):
            assert (v - s1[k]).abs().max().item() <= 2.5e-4, k
        else:
            upd0, upd1 = (v - init[k]).double(), (s1[k] - init[k]).double()
            s = upd0.abs().max().item()
            if s > 1e-12:
                ulp = 2 * 1.1920929e-07 * v.abs().max().item()
                worst = max(worst, max((upd1 - upd0).abs().max().item() - ulp, 0.0) / s)
    assert worst <= 5e-3, worst


# ------------------------------------------------------------------------------------------------- row f4: the loss
def _loss_case(lib, dev, B, C, H, W, channels_last):
    """senas_dice_ce_forward/backward (through the autograd wrapper) against the PyTorch expression of senas_b200.loss
    (itself checked against the reference's SegmentationLosses('dice_ce') in tests/test_oracle_golden.py)."""
    from senas_b200 import loss as L
    torch.manual_seed(B * 100 + C)
    logits = (torch.randn(B, C, H, W) * 3).to(dev)
    if channels_last:
        logits = logits.contiguous(memory_format=torch.channels_last)
    target = torch.randint(0, C, (B, H, W)).to(dev)
    a = logits.clone().requires_grad_(True)
    L.fused_loss[0] = False
    try:
        want = L.DiceCrossEntropyLoss()(a, target)
    finally:
        L.fused_loss[0] = True
    (want * 1.7).backward()
    b = logits.clone().requires_grad_(True)
    got = L._DiceCEFn.apply(b, target, 1e-5, lib)
    (got * 1.7).backward()
    assert abs(got.item() - want.item()) <= 2e-6 * abs(want.item()), (got.item(), want.item())
    assert torch.allclose(b.grad, a.grad, rtol=1e-4, atol=1e-6 * a.grad.abs().max().item()), (b.grad - a.grad).abs().max().item()


@pytest.mark.parametrize('B,C,H,W,cl', [(2, 2, 16, 16, True), (1, 2, 5, 7, False), (3, 5, 9, 4, True), (2, 8, 3, 3, False)])
def test_dice_ce_loss_emulated(emu, B, C, H, W, cl):
    _loss_case(emu, 'cpu', B, C, H, W, cl)


@pytest.mark.gpu
@pytest.mark.parametrize('B,C,H,W,cl', [(16, 2, 256, 256, True), (4, 2, 64, 64, False), (2, 5, 33, 17, True)])
def test_dice_ce_loss_gpu(B, C, H, W, cl):
    from senas_b200 import _lib
    _loss_case(_lib.get(), 'cuda', B, C, H, W, cl)


@pytest.mark.gpu
def test_stem_basic_block_through_convbn_vs_pytorch():
    """stem1's BasicBlock (utils/operations.py:235-268: conv3x3 + BN + ReLU + conv3x3 + BN + residual) through two
    senas_convbn_* calls (ops.stem_convbn; bf16 mode) against the same module evaluated by PyTorch in fp32."""
    import senas_b200
    from senas_b200 import ops
    senas_b200.exact_fp32()
    dev = 'cuda'
    torch.manual_seed(7)
    blk = ops.BasicBlock(32, 32).to(dev)
    blk.apply(senas_b200.weights_init)
    with torch.no_grad():
        for conv in (blk.conv1, blk.conv2):
            conv.weight.copy_(conv.weight.bfloat16().float())
    ref = copy.deepcopy(blk)
    x = torch.randn(4, 32, 128, 128, device=dev).bfloat16().float().contiguous(memory_format=torch.channels_last)
    gout = torch.randn(4, 32, 128, 128, device=dev).contiguous(memory_format=torch.channels_last)
    xr = x.clone().requires_grad_(True)
    senas_b200.set_conv_mode('fp32')
    want = ref(xr)
    want.backward(gout)
    old = ops.stem_convbn[0]
    ops.stem_convbn[0] = True
    senas_b200.set_conv_mode('bf16')
    try:
        xg = x.clone().requires_grad_(True)
        n0 = senas_b200._lib.get().senas_launch_count()
        got = blk(xg)
        got.backward(gout)
        torch.cuda.synchronize()
        assert senas_b200._lib.get().senas_launch_count() - n0 >= 20
    finally:
        ops.stem_convbn[0] = old
        senas_b200.set_conv_mode('fp32')

    def rel(a, b):
        return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()
    # the second conv reads relu(bn1(.)), which is NOT bf16-representable: forward at the 2e-2 gate of the mode
    assert rel(got, want.detach()) <= 2e-2
    assert rel(xg.grad, xr.grad) <= 2e-2
    for n, p in blk.named_parameters():
        assert rel(p.grad, dict(ref.named_parameters())[n].grad) <= 2e-2, n
