"""CUDA-graph capture of a whole search step (arch step + weight step, experiments/search_arc.py:252-293).

The supernet has static shapes and the step issues ~9 000 small kernels plus ~3 400 parameter tensors worth of
autograd / optimizer bookkeeping, so an eagerly launched step is bound by the host (measured on B200: enqueue time ==
step time).  Capturing forward, loss, backward, gradient clipping and both optimizer steps makes the step cost what the
kernels cost.  libsenas_b200 is capture-safe: it allocates nothing, never synchronises, and all its plans / scratch
buffers are created during the warm-up iterations that precede the capture.

Data parallel (one process per GPU).  The step consists of three segments that exchange gradients only through two
flat buckets they pack / unpack themselves:

    segment 1: arch pass (fwd, loss, bwd)            -> pack arch grads (174 floats, pre-scaled 1/world)
    all-reduce(bucket_arch)
    segment 2: unpack, Adam on alpha/beta/gamma, weight pass (fwd, loss, bwd) -> pack all grads (1.97 M floats)
    all-reduce(bucket_all)
    segment 3: unpack, clip_grad_norm_, SGD

With ``comm`` (senas_b200.comm.Comm, the NCCL communicator owned by libsenas_b200) the all-reduces are captured
INTO the graph: the whole data-parallel step is ONE replayed graph, no host round trip between the segments, and the
weight gradients are **bucketed per cell and overlapped with backward**: every fused cell hands its flat gradient
buffer (0.1 M floats) to ``_sink`` the moment its backward kernels are enqueued; a side stream scales it by 1/world and
all-reduces it in place while autograd continues with the cells below (the graph records this as a branch that joins
before gradient clipping).  Only the parameters outside the cells (stems, Shrink/Rectify blocks, head: 0.3 M of the
1.97 M floats) are reduced after backward, as one bucket.  ``overlap=False`` keeps the single post-backward bucket.  Without
it (``group`` only) the collectives of ``torch.distributed`` stay outside -- capturing those deadlocked on this stack
(watchdog thread) -- and the step is three graphs with two eager all-reduces between them (round-1 path, kept as the
fallback).

BatchNorm statistics and the soft-dice sums stay local to each rank (replica semantics of the reference's DataParallel
path for BN; per-shard dice is the stated choice of SURVEY.md H8 for the graphed path), gradients are averaged.
"""
import os

import torch
import torch.distributed as dist


def _flat_views(flat, tensors):
    return [v.view_as(t) for v, t in zip(flat.split([t.numel() for t in tensors]), tensors)]


class GraphedSearchStep:
    """``step = GraphedSearchStep(model, criterion, w_opt, a_opt, (xt, yt, xv, yv), grad_clip=5, group=None)`` then
    ``loss = step(x_train, y_train, x_valid, y_valid)`` (device or pinned-host tensors of the captured shapes)."""

    def __init__(self, model, criterion, w_opt, a_opt, example, grad_clip=5.0, warmup=3, group=None,
                 force_segments=False, capture_error_mode='global', concurrent_cells=True, defer_wgrad=False,
                 restore_state=True, comm=None, overlap=True, fused_optim=False, arch_grads_only=False):
        self.static = [t.clone() for t in example]
        # independent cells of one level of the UNet++ triangle on separate streams: the captured graph overlaps the
        # small latency-bound cells with the large one of the level (senas_b200/supernet.py).  Only while warming up
        # and capturing: eager forwards outside the graphs (validation) keep the serial single-stream walk.
        self.defer_wgrad = bool(defer_wgrad)
        self._net = getattr(model, 'net', None)
        self._concurrent = bool(concurrent_cells) and self._net is not None and hasattr(self._net, 'concurrent_cells')
        self.model, self.criterion, self.w_opt, self.a_opt = model, criterion, w_opt, a_opt
        self.grad_clip, self.group, self.comm = grad_clip, group, comm
        self.world = comm.world if comm is not None else (dist.get_world_size(group) if group is not None else 1)
        self.segmented = self.world > 1 or force_segments
        self.capture_error_mode = capture_error_mode
        for g in a_opt.param_groups:  # Adam keeps `step` on the device when capturable
            g['capturable'] = True
        seen, self.params = set(), []
        for p in model.parameters():
            if id(p) not in seen and p.requires_grad:
                seen.add(id(p))
                self.params.append(p)
        self.arch = [p for g in a_opt.param_groups for p in g['params']]
        dev = self.params[0].device
        self.overlap = bool(overlap) and comm is not None and self.world > 1
        if self.overlap and self.defer_wgrad:
            # the per-cell all-reduce is ordered after the cell's backward through the CALLER's stream; deferred weight-gradient
            # lanes are not joined to it yet -- the two options exclude each other
            self.defer_wgrad = False
        # opt-in: the architecture step asks autograd for the alpha / beta / gamma gradients only.  The reference's arch
        # step (search/senas_search.py Architecture.step: loss.backward()) also produces every weight gradient and
        # search_arc.py:271 zeroes them before any use; with this flag they are never computed (torch.autograd.grad on
        # the arch parameters + senas_bwd_args_t.skip_wgrad).  Same parameters after the step; OFF by default because it
        # is work the reference performs inside the step.
        self.arch_grads_only = bool(arch_grads_only)
        self.n_rest = 0
        # clip + SGD / Adam as flat-buffer kernels of libsenas_b200 (senas_b200/optim.py).  Parameters,
        # gradients and momentum move into arenas; the cells' weight gradients are produced in the gradient arena.
        self.fopt = None
        if fused_optim:
            from .optim import FusedSearchOptim
            self.fopt = FusedSearchOptim(model, w_opt, a_opt, grad_clip)
            self.params = self.fopt.params
            self._comm_stream = torch.cuda.Stream() if self.overlap else None
            self.all_views = [self.fopt.grad_views[id(p)] for p in self.params]
            if self.segmented:
                self.bucket_arch = self.fopt.flat_g[:self.fopt.n_arch]
                self.bucket_all = self.fopt.flat_g
                self.n_rest = self.fopt.n_rest
        elif self.segmented:
            self.bucket_arch = torch.zeros(sum(p.numel() for p in self.arch), device=dev)
            self.bucket_all = torch.zeros(sum(p.numel() for p in self.params), device=dev)
            self.arch_views = _flat_views(self.bucket_arch, self.arch)
            # flat layout [parameters outside the fused cells | parameters of the fused cells]: with ``overlap`` the
            # second part arrives already averaged (per-cell all-reduces issued during backward), only the first part
            # is all-reduced after backward
            owned = self._owned_ids() if self.overlap else set()
            order = [p for p in self.params if id(p) not in owned] + [p for p in self.params if id(p) in owned]
            self.n_rest = sum(p.numel() for p in self.params if id(p) not in owned)
            where = dict(zip((id(p) for p in order), _flat_views(self.bucket_all, order)))
            self.all_views = [where[id(p)] for p in self.params]
            self._comm_stream = torch.cuda.Stream() if self.overlap else None
        # The warm-up iterations are real optimizer steps (they create the optimizer state tensors whose addresses the
        # graphs bake in, let cudnn.benchmark pick algorithms and size the library's scratch).  With ``restore_state``
        # the model (weights, BatchNorm buffers, arch parameters) and both optimizers are put back afterwards, in
        # place, so that the first replay is the first step of the search.
        snap = self._snapshot() if restore_state else None
        self._set_concurrent(True)
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(warmup):
                    self._run(capture=False, arch=True)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            # the library's scratch buffers are baked into the graphs: pin them (fused.scratch_for never frees a buffer
            # it handed out, and refuses to regrow a pinned slot silently)
            from . import fused
            self._scratch_ptrs = fused.pin_scratch(self.params[0].device)
            self._variants = {}          # arch flag -> (graphs, lr the SGD update was captured with)
            self._capture(arch=True)
        finally:
            self._set_concurrent(False)
        if snap is not None:
            self._restore(snap)
        if self.fopt is not None:
            if snap is not None:
                self.fopt.load_from_torch_state()  # (warm-up steps ran on the arenas: back to the optimizers' own state)
            self.fopt.publish_state()

    def release(self):
        """Destroy the captured graphs (required before the NCCL communicator they captured is destroyed)."""
        self._variants = {}
        self.graphs = []
        torch.cuda.synchronize()

    def _owned_ids(self):
        """ids of the parameters whose gradients come back in the fused graphs' flat buffers (every MixedOp candidate)."""
        from .cell import MixedOp
        return {id(p) for m in self.model.modules() if isinstance(m, MixedOp) for p in m.parameters()}

    def _sink(self, runner, flat):
        """Called from a fused cell's backward right after its kernels were enqueued (fused.set_grad_sink): average this
        cell's weight gradients over the ranks on the side stream while backward goes on.  Autograd adopts views of
        ``flat`` as p.grad (the weight pass starts from p.grad = None), so the reduced values are what clip / SGD see."""
        cs = self._comm_stream
        cs.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(cs):
            flat.mul_(1.0 / self.world)
            self.comm.all_reduce_(flat, stream=cs)

    def flat_grads(self):
        """The (averaged, after the step: clipped) gradients of the last weight pass in ``model.parameters()`` order."""
        return torch.cat([v.reshape(-1) for v in self.all_views])

    # -- state ------------------------------------------------------------------------------------------------
    def _set_concurrent(self, on):
        if self._concurrent:
            self._net.concurrent_cells = bool(on)

    def _snapshot(self):
        model = {k: v.detach().clone() for k, v in self.model.state_dict().items()}
        return model, [o.state_dict() for o in (self.w_opt, self.a_opt)]

    def _restore(self, snap):
        """In place (the graphs hold the addresses): model tensors back to their values, optimizer state tensors that did
        not exist before the warm-up back to zero (SGD's first step sets buf = grad, i.e. 0.9 * 0 + grad; Adam starts from
        exp_avg = exp_avg_sq = step = 0), those that did exist back to their values."""
        model, opts = snap
        with torch.no_grad():
            for k, v in self.model.state_dict().items():
                v.copy_(model[k])
            for opt, old in zip((self.w_opt, self.a_opt), opts):
                old_state = old['state']
                index = {id(p): i for i, p in enumerate(p for g in opt.param_groups for p in g['params'])}
                for p, st in opt.state.items():
                    prev = old_state.get(index[id(p)], {})
                    for name, t in st.items():
                        if not torch.is_tensor(t):
                            continue
                        if name in prev and torch.is_tensor(prev[name]):
                            t.copy_(prev[name])
                        else:
                            t.zero_()
        torch.cuda.synchronize()

    def _lr(self):
        return tuple(float(g['lr']) for g in self.w_opt.param_groups) + tuple(float(g['lr']) for g in self.a_opt.param_groups)

    def _capture(self, arch):
        self.graphs = []
        self._set_concurrent(True)
        try:
            self._run(capture=True, arch=arch)
        finally:
            self._set_concurrent(False)
        self._variants[arch] = (self.graphs, self._lr(), self.loss)  # (each capture has its own static loss tensor)

    # -- the three segments -----------------------------------------------------------------------------------
    # Segments exchange gradients only through the two flat buckets (ordinary allocations): a segment packs its
    # gradients and drops them before it ends, the next one adopts persistent *views* of the bucket as p.grad, so no
    # tensor that lives in one graph's private memory pool is touched by another graph.
    def _seg1(self):
        if not self._arch:  # epochs before alpha_begin (experiments/search_arc.py:268): no architecture step
            for p in self.params:
                p.grad = None
            return
        xt, yt, xv, yv = self.static
        # the arch pass also produces (unused) weight gradients; dropping the stale ones first makes autograd adopt
        # the new buffers instead of launching ~3 400 in-place adds (same values reach both optimizers either way)
        for p in self.params:
            p.grad = None
        self.a_opt.zero_grad(set_to_none=True)
        if self.arch_grads_only:
            from . import fused
            loss = self.criterion(self.model(xv), yv)
            fused.set_skip_wgrad(True)
            try:
                grads = torch.autograd.grad(loss, self.arch, allow_unused=True)
            finally:
                fused.set_skip_wgrad(False)
            for p, g in zip(self.arch, grads):
                p.grad = g if g is not None else torch.zeros_like(p)
        else:
            self._backward(self.criterion(self.model(xv), yv))
        if self.fopt is not None:
            self.fopt.pack_rest(self.fopt.arch, self.fopt.arch_grad_views)
            if self.world > 1:
                self.bucket_arch.mul_(1.0 / self.world)
            for p in self.params:
                p.grad = None
        elif self.segmented:
            torch._foreach_copy_(self.arch_views, [p.grad for p in self.arch])
            self.bucket_arch.mul_(1.0 / self.world)
            for p in self.params:
                p.grad = None

    def _capture_stream(self):
        """High-priority capture stream: the kernels the step enqueues on the caller's stream (node statistics, the stock
        blocks between the cells) are its serial spine and must not queue behind the low-priority weight-gradient lanes
        of the library; stream priorities are recorded in the graph's kernel nodes."""
        if os.environ.get('SENAS_CAPTURE_PRIORITY', '1') == '0':
            return torch.cuda.Stream()
        if getattr(self, '_cap_stream', None) is None:
            self._cap_stream = torch.cuda.Stream(priority=-1)
        return self._cap_stream

    def _backward(self, loss):
        """backward, optionally (``defer_wgrad``) with the weight-gradient lanes of every fused call left running (joined
        by the next call of the same slot, or here at the end).  Measured: 102.6 ms vs 102.5 ms per step -- the next
        fused call of the slot follows too closely for the tails to hide, so it is off by default."""
        from . import fused
        if self.defer_wgrad:
            fused.set_defer(True)
        try:
            loss.backward()
        finally:
            if self.defer_wgrad:
                fused.set_defer(False)  # flushes: the current stream waits for everything still pending

    def _seg2(self):
        xt, yt, xv, yv = self.static
        if self._arch:
            if self.fopt is not None:
                self.fopt.adam_step()
            else:
                if self.segmented:
                    for p, v in zip(self.arch, self.arch_views):
                        p.grad = v
                self.a_opt.step()
        self.w_opt.zero_grad(set_to_none=True)
        loss = self.criterion(self.model(xt), yt)
        if self.overlap:
            from . import fused
            fused.set_grad_sink(self._sink)
            try:
                self._backward(loss)
            finally:
                fused.set_grad_sink(None)
            torch.cuda.current_stream().wait_stream(self._comm_stream)  # join the all-reduce branch
        else:
            self._backward(loss)
        if self.fopt is not None:
            if not self._capturing and not self.fopt.check_direct():
                raise RuntimeError('senas_b200: autograd did not adopt the arena views of the fused cells\' gradients')
            self.fopt.pack_rest()  # the cells' gradients are already in the arena
            if self.world > 1:
                (self.bucket_all[:self.n_rest] if self.overlap else self.bucket_all).mul_(1.0 / self.world)
            for p in self.params:
                p.grad = None
        elif self.segmented:
            torch._foreach_copy_(self.all_views, [p.grad for p in self.params])
            (self.bucket_all[:self.n_rest] if self.overlap else self.bucket_all).mul_(1.0 / self.world)
            for p in self.params:
                p.grad = None
        self.loss = loss.detach()

    def _seg3(self):
        if self.fopt is not None:
            self.fopt.sgd_step()
            return
        if self.segmented:
            for p, v in zip(self.params, self.all_views):
                p.grad = v
        torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.grad_clip)
        self.w_opt.step()

    def _run(self, capture, arch=True):
        self._arch = bool(arch)
        self._capturing = bool(capture)
        segs = (self._seg1, self._seg2, self._seg3)
        if self.segmented and self.comm is not None and self.world > 1:  # one graph with the NCCL all-reduces inside
            def whole():
                self._seg1()
                if self._arch:
                    self.comm.all_reduce_(self.bucket_arch)
                self._seg2()
                self.comm.all_reduce_(self.bucket_all[:self.n_rest] if self.overlap else self.bucket_all)
                self._seg3()
            if capture:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=self._capture_stream(), capture_error_mode=self.capture_error_mode):
                    whole()
                self.graphs = [g]
            else:
                whole()
            return
        if not self.segmented:  # nothing to exchange: one graph
            if capture:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=self._capture_stream(), capture_error_mode=self.capture_error_mode):
                    for s in segs:
                        s()
                self.graphs = [g]
            else:
                for s in segs:
                    s()
            return
        for i, s in enumerate(segs):
            if capture:  # each graph keeps its own memory pool; nothing executes while capturing
                g = torch.cuda.CUDAGraph()
                # thread_local: the NCCL watchdog thread keeps polling CUDA events while we capture
                with torch.cuda.graph(g, stream=self._capture_stream(), capture_error_mode=self.capture_error_mode):
                    s()
                self.graphs.append(g)
            else:
                s()
                if i < 2 and self.world > 1 and (i == 1 or self._arch):
                    dist.all_reduce(self.bucket_arch if i == 0 else self.bucket_all, group=self.group)

    def __call__(self, xt, yt, xv, yv, arch=True):
        """One search step.  ``arch=False`` skips the architecture step (the driver does for ``epoch < alpha_begin``,
        experiments/search_arc.py:268).  The optimizers' learning rates are kernel arguments of the captured updates:
        when a scheduler has changed them since the capture (CosineAnnealingLR, once per epoch, search_arc.py:296) the
        step is captured again -- no warm-up is needed, nothing executes during a capture."""
        from . import fused
        fused.check_scratch(self._scratch_ptrs)
        arch = bool(arch)
        var = self._variants.get(arch)
        if self.fopt is not None:
            self.fopt.sync_lr()  # device scalars: no re-capture
        if var is None or (self.fopt is None and var[1] != self._lr()):
            self._capture(arch)
            var = self._variants[arch]
        graphs = var[0]
        for dst, src in zip(self.static, (xt, yt, xv, yv)):
            if src is not dst:
                dst.copy_(src, non_blocking=True)
        self.loss = var[2]
        if len(graphs) == 1:  # single GPU, or data parallel with the all-reduces captured (comm)
            graphs[0].replay()
        else:
            graphs[0].replay()
            if self.world > 1 and arch:
                dist.all_reduce(self.bucket_arch, group=self.group)
            graphs[1].replay()
            if self.world > 1:
                dist.all_reduce(self.bucket_all, group=self.group)
            graphs[2].replay()
        return self.loss
