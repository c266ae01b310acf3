"""Data-parallel search: one process per GPU, batch-sharded, gradients of the weights and of the
architecture parameters (alpha/beta/gamma) all-reduced over NCCL (NVLink 5 / NVSwitch) in flat
buckets that are launched from autograd hooks while backward is still running.

Replaces the reference's in-process replica path (search/senas_search.py:262-279, broken as shipped,
SURVEY.md section 2.2).  Semantics kept from it: BatchNorm statistics stay local to each replica; the
loss is the global-batch loss (senas_b200.loss with ``group``), hence gradients are SUMMED.
"""
import torch
import torch.distributed as dist


class GradBuckets:
    """Buckets follow reverse parameter order (head -> up cells -> down cells -> stem), i.e. roughly the
    order in which backward produces gradients; arch parameters (which accumulate over every cell) go
    into the last bucket."""

    def __init__(self, params, arch_params=(), bucket_floats=1 << 19, group=None, exclude=()):
        self.group = group
        arch_ids = {id(p) for p in arch_params} | set(exclude)
        seen, ordered = set(), []
        for p in reversed([p for p in params if p.requires_grad]):
            if id(p) not in seen and id(p) not in arch_ids:
                seen.add(id(p))
                ordered.append(p)
        buckets, cur, n = [], [], 0
        for p in ordered:
            cur.append(p)
            n += p.numel()
            if n >= bucket_floats:
                buckets.append(cur)
                cur, n = [], 0
        if cur:
            buckets.append(cur)
        arch = [p for p in arch_params if p.requires_grad and id(p) not in set(exclude)]
        if arch:
            buckets.append(list({id(p): p for p in arch}.values()))
        self.buckets = buckets
        self.flat = [torch.zeros(sum(p.numel() for p in b), dtype=torch.float32, device=b[0].device) for b in buckets]
        self.where = {}
        for bi, b in enumerate(buckets):
            for p in b:
                self.where[id(p)] = bi
        self.pending = [0] * len(buckets)
        self.work = [None] * len(buckets)
        self.expect = [len(b) for b in buckets]
        self.hooks = [p.register_post_accumulate_grad_hook(self._hook) for b in buckets for p in b]
        self.enabled = True

    def _hook(self, p):
        if not self.enabled:
            return
        bi = self.where[id(p)]
        self.pending[bi] += 1
        if self.pending[bi] == self.expect[bi]:
            self._launch(bi)

    def _launch(self, bi):
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.buckets[bi]]
        torch._foreach_copy_(list(self.flat[bi].split([g.numel() for g in grads])), [g.reshape(-1) for g in grads])
        self.work[bi] = dist.all_reduce(self.flat[bi], group=self.group, async_op=True)

    def finish(self):
        """Wait for every bucket (launching those whose hooks did not all fire, e.g. unused parameters)
        and write the reduced gradients back.  Call after ``loss.backward()``, before clip / step."""
        for bi, b in enumerate(self.buckets):
            if self.work[bi] is None:
                self._launch(bi)
        for bi, b in enumerate(self.buckets):
            self.work[bi].wait()
            outs = self.flat[bi].split([p.numel() for p in b])
            for p, o in zip(b, outs):
                if p.grad is None:
                    p.grad = o.view_as(p).clone()
                else:
                    p.grad.copy_(o.view_as(p))
            self.work[bi], self.pending[bi] = None, 0


class FusedGradReducer:
    """All-reduce of the fused cells' parameter gradients without per-parameter hooks: every fused backward hands
    over its flat gradient buffer (one per MixedOp / Cell graph, ~0.1 M floats for a cell) and the NCCL all-reduce is
    launched on it at once, overlapping the rest of backward.  Autograd then adopts *views* of that buffer as
    ``p.grad`` (it steals a gradient when ``p.grad is None``), so the reduced values appear in place.  ``finish``
    verifies the aliasing and falls back to a copy if autograd had to accumulate instead."""

    def __init__(self, group=None):
        from . import fused
        self.group, self.pending = group, []
        fused.set_grad_sink(self._sink)

    def _sink(self, runner, flat):
        self.pending.append((runner, flat, dist.all_reduce(flat, group=self.group, async_op=True)))

    def finish(self):
        for runner, flat, work in self.pending:
            work.wait()
            p0 = runner.params[0]
            if p0.grad is None or p0.grad.data_ptr() != flat.data_ptr():  # not adopted as a view: write back
                for p, g in zip(runner.params, torch.split(flat, runner.sizes)):
                    p.grad = g.view_as(p).clone()
        self.pending.clear()

    def owned(self, module):
        """ids of the parameters that travel through the fused graphs of ``module`` (every MixedOp candidate)."""
        from .cell import MixedOp
        return {id(p) for m in module.modules() if isinstance(m, MixedOp) for p in m.parameters()}


def broadcast_parameters(module, src=0, group=None):
    """Make every replica start from rank ``src``'s weights, arch parameters and BN buffers."""
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src, group=group)
