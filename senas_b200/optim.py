"""gradient clipping + SGD on the weights and Adam on the architecture parameters as fused
kernels over flat buffers (experiments/search_arc.py:282-293, Architecture.step).

The supernet has 3 367 parameter tensors.  ``FusedSearchOptim`` re-homes every ``p.data`` into ONE fp32 arena (modules,
``state_dict`` and checkpoints are unaffected: the tensors keep their shapes and names, only their storage moves), keeps
one gradient arena and one momentum arena of the same layout, and replaces

    torch.nn.utils.clip_grad_norm_(model.parameters(), grad_clip); w_optimizer.step()      # ~300 foreach launches
    a_optimizer.step()                                                                      # ~20 launches

by ``senas_sgd_clip_step`` (2 launches) and ``senas_adam_step`` (1 launch) of libsenas_b200.  Layout of the arenas:

    [ architecture parameters | parameters outside the fused cells | cell 0 | cell 1 | ... ]

where every cell's slice follows the order of its fused graph's flat gradient buffer, so ``senas_graph_backward`` writes
the cell's weight gradients **directly into the arena** (``GraphRunner.grad_buffer``): no per-parameter gradient tensor,
no packing copy, and in data-parallel runs the slice is all-reduced in place while backward continues.  The learning
rates live in device scalars: a scheduler (CosineAnnealingLR, search_arc.py:296) takes effect on the next replay of the
captured step without re-capturing it.  The torch optimizers stay the owners of the hyper-parameters and their
``state_dict()`` keeps working: their per-parameter state entries are views of the arenas (``publish_state``).

Lifetime: the arenas belong to this object and the parameters alias them -- build it after the model is on its device
and do not move the model (``.to()`` / ``.cuda()``) afterwards (a moved Cell drops its fused graph and with it the
``grad_buffer`` binding; ``check_direct`` reports that, GraphedSearchStep raises during its warm-up).
"""
import ctypes as C

import torch

from . import _lib


def _pad4(n):
    return (n + 3) & ~3


class _NullCtx:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


class FusedSearchOptim:
    def __init__(self, model, w_opt, a_opt, grad_clip=5.0, lib=None):
        from .fused import GraphRunner  # noqa: F401  (runner discovery below)
        self.lib = lib if lib is not None else _lib.get()
        self.model, self.w_opt, self.a_opt, self.grad_clip = model, w_opt, a_opt, float(grad_clip or 0.0)
        if len(w_opt.param_groups) != 1 or len(a_opt.param_groups) != 1:
            raise NotImplementedError('senas_b200.FusedSearchOptim: one parameter group per optimizer')
        gw, ga = w_opt.param_groups[0], a_opt.param_groups[0]
        if not isinstance(w_opt, torch.optim.SGD) or gw.get('nesterov') or gw.get('dampening', 0) != 0 or gw.get('maximize'):
            raise NotImplementedError('senas_b200.FusedSearchOptim: weights need torch.optim.SGD(momentum, weight_decay)')
        if not isinstance(a_opt, torch.optim.Adam) or ga.get('amsgrad') or ga.get('maximize'):
            raise NotImplementedError('senas_b200.FusedSearchOptim: architecture parameters need torch.optim.Adam')
        seen, params = set(), []
        for p in gw['params']:
            if id(p) not in seen and p.requires_grad:
                seen.add(id(p))
                params.append(p)
        self.params = params
        self.arch = [p for p in ga['params']]
        arch_ids = {id(p) for p in self.arch}
        if any(id(p) not in seen for p in self.arch):
            raise NotImplementedError('senas_b200.FusedSearchOptim: the SGD must also own the architecture parameters '
                                      '(search_arc.py builds it on model.parameters())')
        dev = params[0].device
        if any(p.dtype != torch.float32 or p.device != dev for p in params):
            raise NotImplementedError('senas_b200.FusedSearchOptim: fp32 parameters on one device')
        # fused graphs (one per Cell / free-standing MixedOp): their parameter order IS the arena order of their slice
        from .cell import Cell, MixedOp
        self.runners = []
        cell_ids, inside = set(), set()
        for m in model.modules():
            if isinstance(m, Cell):
                inside.update(id(op) for op in m._ops)  # these MixedOps travel through the Cell's graph
            elif not isinstance(m, MixedOp) or id(m) in inside:
                continue
            r = m._ensure_runner()
            if any(id(p) not in seen for p in r.params):
                raise NotImplementedError('senas_b200.FusedSearchOptim: a fused cell has parameters the SGD does not own')
            if any(p.numel() % 4 for p in r.params):
                raise NotImplementedError('senas_b200.FusedSearchOptim: cell parameter sizes must be multiples of 4')
            self.runners.append(r)
            cell_ids.update(id(p) for p in r.params)
        others = [p for p in params if id(p) not in arch_ids and id(p) not in cell_ids]
        self.layout = {}  # id(p) -> (offset, numel)
        off = 0
        for p in self.arch:
            self.layout[id(p)] = (off, p.numel())
            off += _pad4(p.numel())
        self.n_arch = off
        for p in others:
            self.layout[id(p)] = (off, p.numel())
            off += _pad4(p.numel())
        self.n_rest = off  # [0, n_rest): everything that does not come back through a fused graph's flat buffer
        self.runner_slices = []
        for r in self.runners:
            self.runner_slices.append((off, r.grad_floats))
            for p, n in zip(r.params, r.sizes):
                self.layout[id(p)] = (off, n)
                off += n
        self.n = off
        self.rest_params = self.arch + others
        self.flat_p = torch.zeros(self.n, device=dev)
        self.flat_g = torch.zeros(self.n, device=dev)
        self.flat_m = torch.zeros(self.n, device=dev)
        self.adam_avg = torch.zeros(self.n_arch, device=dev)
        self.adam_sq = torch.zeros(self.n_arch, device=dev)
        self.adam_t = torch.zeros((), device=dev)
        self.scratch = torch.zeros(2048, device=dev)
        self.norm = torch.zeros((), device=dev)
        self.lr_w = torch.tensor(float(gw['lr']), device=dev)
        self.lr_a = torch.tensor(float(ga['lr']), device=dev)
        self._lr_host = (float(gw['lr']), float(ga['lr']))
        with torch.no_grad():
            for p in params:
                o, n = self.layout[id(p)]
                v = self.flat_p[o:o + n].view_as(p)
                v.copy_(p.data)
                p.data = v
        self.grad_views = {id(p): self.flat_g[o:o + n].view_as(p) for p in params for o, n in [self.layout[id(p)]]}
        self.rest_grad_views = [self.grad_views[id(p)] for p in self.rest_params]
        self.arch_grad_views = [self.grad_views[id(p)] for p in self.arch]
        for r, (o, n) in zip(self.runners, self.runner_slices):
            r.grad_buffer = self.flat_g[o:o + n]
            r.refresh()  # parameter storage moved
        self.load_from_torch_state()

    # -- state exchange with the torch optimizers -------------------------------------------------------------------
    def load_from_torch_state(self):
        """Import momentum / Adam moments that the torch optimizers already hold (resume); zero otherwise."""
        with torch.no_grad():
            self.flat_m.zero_(), self.adam_avg.zero_(), self.adam_sq.zero_(), self.adam_t.zero_()
            for p in self.params:
                st = self.w_opt.state.get(p, {})
                buf = st.get('momentum_buffer')
                o, n = self.layout[id(p)]
                if torch.is_tensor(buf) and buf.data_ptr() != self.flat_m[o:o + n].data_ptr():
                    self.flat_m[o:o + n].copy_(buf.reshape(-1))
            for p in self.arch:
                st = self.a_opt.state.get(p, {})
                o, n = self.layout[id(p)]
                if torch.is_tensor(st.get('exp_avg')) and st['exp_avg'].data_ptr() != self.adam_avg[o:o + n].data_ptr():
                    self.adam_avg[o:o + n].copy_(st['exp_avg'].reshape(-1))
                    self.adam_sq[o:o + n].copy_(st['exp_avg_sq'].reshape(-1))
                    self.adam_t.fill_(float(st['step']))

    def publish_state(self):
        """Make the torch optimizers' per-parameter state entries views of the arenas (``state_dict()`` / checkpoints)."""
        for p in self.params:
            o, n = self.layout[id(p)]
            self.w_opt.state[p] = {'momentum_buffer': self.flat_m[o:o + n].view_as(p)}
        for p in self.arch:
            o, n = self.layout[id(p)]
            self.a_opt.state[p] = {'step': self.adam_t, 'exp_avg': self.adam_avg[o:o + n].view_as(p),
                                   'exp_avg_sq': self.adam_sq[o:o + n].view_as(p)}

    def sync_lr(self):
        """Copy the optimizers' current learning rates into the device scalars (no-op when unchanged)."""
        cur = (float(self.w_opt.param_groups[0]['lr']), float(self.a_opt.param_groups[0]['lr']))
        if cur != self._lr_host:
            self.lr_w.fill_(cur[0]), self.lr_a.fill_(cur[1])
            self._lr_host = cur

    # -- the updates ------------------------------------------------------------------------------------------------
    def _stream(self):
        return torch.cuda.current_stream(self.flat_p.device).cuda_stream if self.flat_p.is_cuda else 0

    def pack_rest(self, params=None, views=None):
        """Gradients that autograd produced as separate tensors (everything outside the fused cells) -> arena."""
        params = self.rest_params if params is None else params
        views = self.rest_grad_views if views is None else views
        grads = []
        for p in params:
            if p.grad is None:
                raise RuntimeError('senas_b200.FusedSearchOptim: a parameter received no gradient (torch.optim.SGD would '
                                   'skip it, the flat update cannot)')
            grads.append(p.grad)
        torch._foreach_copy_(views, grads)

    def check_direct(self):
        """True iff every fused cell's parameter gradients of the last backward already live in the arena."""
        for r in self.runners:
            for p in r.params:
                if p.grad is None or p.grad.data_ptr() != self.grad_views[id(p)].data_ptr():
                    return False
        return True

    def _guard(self):
        """Kernels launch on the CURRENT device: make it the arenas' (a model on cuda:1 while cuda:0 is current)."""
        return torch.cuda.device(self.flat_p.device) if self.flat_p.is_cuda else _NullCtx()

    def adam_step(self):
        g = self.a_opt.param_groups[0]
        with self._guard():
            self._adam_step(g)

    def _adam_step(self, g):
        _lib.check(self.lib, self.lib.senas_adam_step(
            self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.adam_avg.data_ptr(), self.adam_sq.data_ptr(),
            self.adam_t.data_ptr(), self.n_arch, self.lr_a.data_ptr(), float(g['betas'][0]), float(g['betas'][1]),
            float(g['eps']), float(g['weight_decay']), self._stream()))

    def sgd_step(self):
        g = self.w_opt.param_groups[0]
        with self._guard():
            self._sgd_step(g)

    def _sgd_step(self, g):
        _lib.check(self.lib, self.lib.senas_sgd_clip_step(
            self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.flat_m.data_ptr(), self.n, self.lr_w.data_ptr(),
            float(g['momentum']), float(g['weight_decay']), self.grad_clip, self.scratch.data_ptr(), self.norm.data_ptr(),
            self._stream()))
