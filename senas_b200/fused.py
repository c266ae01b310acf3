"""Host driver of one edge graph (a MixedOp, or the node loop + concat of a Cell).

Builds the C-ABI descriptor from the parameter modules, owns the opaque graph handle, and wraps
``senas_graph_forward`` / ``senas_graph_backward`` in a ``torch.autograd.Function``.  PyTorch
is used for device memory, streams and autograd plumbing only.
"""
import ctypes as C

import torch

from . import _lib
from .ops import KIND_AVG_POOL, KIND_CONV, KIND_DEPSEP, KIND_IDENTITY, KIND_NONE, KIND_SE_CONV, KIND_UP_SAMPLE


def _bn_slots(bn):
    return [bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.num_batches_tracked]


def candidate_slots(op):
    """(kind, ksize, dilation, [12 slot tensors or None]) for one candidate block, in the slot order documented in
    include/senas_b200.h.  Structural (duck-typed) so that it reads senas_b200's parameter containers and the
    reference's own modules (utils/operations.py:89-115,167-203) alike."""
    slots = [None] * _lib.SLOTS
    if hasattr(op, 'norm') and hasattr(op, 'module'):  # AdapterBlock
        mod = op.module
        if isinstance(mod, torch.nn.Identity):
            kind = KIND_IDENTITY
        elif isinstance(mod, torch.nn.AvgPool2d):
            kind = KIND_AVG_POOL
        elif isinstance(mod, torch.nn.Upsample):
            kind = KIND_UP_SAMPLE
        elif type(mod).__name__ == 'ZeroOp':
            kind = KIND_NONE
        else:
            raise TypeError(f'unsupported adapter module {type(mod).__name__}')
        slots[0] = op.conv.weight if hasattr(op, 'conv') else None
        slots[1:6] = _bn_slots(op.norm)
        return kind, 0, 1, slots
    if isinstance(op, torch.nn.Sequential) and len(op) in (2, 3, 5):
        conv = op[0]
        k, dil = conv.kernel_size[0], conv.dilation[0]
        slots[0] = conv.weight
        slots[1:6] = _bn_slots(op[1])
        if len(op) == 2:
            return KIND_CONV, k, dil, slots
        if len(op) == 3:
            slots[6], slots[7] = op[2].excitation[0].weight, op[2].excitation[2].weight
            return KIND_SE_CONV, k, dil, slots
        slots[6] = op[3].weight
        slots[7:12] = _bn_slots(op[4])
        return KIND_DEPSEP, k, 1, slots
    raise TypeError(f'not a MixedOp candidate: {type(op).__name__}')


_scratch = {}
_flags = {'tc_bf16': False, 'override': None}  # override: graph flags forced by the emulator tests
_grad_sink = [None]
_skip_wgrad = [False]


def set_skip_wgrad(on):
    """While set, every fused backward computes data / alpha / beta gradients only and returns no parameter gradients
    (``senas_bwd_args_t.skip_wgrad``): for the architecture step, whose weight gradients the reference discards
    (experiments/search_arc.py:268-271: ``architect.step`` is followed by ``model_optimizer.zero_grad()``)."""
    _skip_wgrad[0] = bool(on)



def set_grad_sink(fn):
    """``fn(runner, flat_grad)`` is called from every fused backward right after the kernels are enqueued, with the
    graph's flat parameter-gradient buffer (senas_b200.dp all-reduces it in place while backward continues)."""
    _grad_sink[0] = fn


def set_conv_mode(mode):
    """'fp32' : every convolution in exact fp32 FMA (gate 1e-4 against the reference);
    'bf16' : the dense / dilated / transposed c_in = 32 convolutions of a cell run as tcgen05 implicit GEMMs with
    TMA-staged NHWC bf16 operands and fp32 TMEM accumulation (gate 2e-2)."""
    if mode not in ('fp32', 'bf16'):
        raise ValueError(mode)
    _flags['tc_bf16'] = mode == 'bf16'


def get_conv_mode():
    return 'bf16' if _flags['tc_bf16'] else 'fp32'


_slot = [0]
_defer = [False]
_held = {}  # slot -> buffers that deferred weight-gradient kernels of the last backward call of the slot may still read


def set_defer(on, lib=None):
    """Deferred join of the weight-gradient lanes (``senas_set_defer``): a fused backward then returns as soon as the
    data / alpha / beta gradients are ordered on the stream, the weight-gradient kernels keep running beside whatever
    comes next, and ``flush()`` must be called before anything reads the parameter gradients.  Used by
    ``GraphedSearchStep``; off by default."""
    lib = lib if lib is not None else _lib.get()
    if not on:
        flush(lib)
    _defer[0] = bool(on)
    lib.senas_set_defer(int(bool(on)))


def flush(lib=None):
    """Make the current stream wait for all deferred weight-gradient work and release the buffers held for it."""
    lib = lib if lib is not None else _lib.get()
    if torch.cuda.is_available():
        lib.senas_flush(torch.cuda.current_stream().cuda_stream)
    _held.clear()



def set_slot(slot):
    """Select the lane set / scratch buffer of the following fused calls (see ``senas_set_slot``): a host that runs
    independent cells concurrently on different CUDA streams gives every stream its own slot."""
    _slot[0] = int(slot)


def get_slot():
    return _slot[0]


_retired = []   # superseded scratch buffers: never freed (a captured CUDA graph may have their address baked in)
_pinned = {}    # (device, slot) -> data_ptr recorded by a live capture (GraphedSearchStep)


def scratch_for(device, nbytes, slot=0):
    """One grow-only scratch buffer per (device, slot), shared by every graph that runs in that slot.  A buffer that has
    been handed out is never returned to the allocator: when a later call needs more bytes the old block is retired
    (kept alive), because captured CUDA graphs keep using its address.  Growing a slot that a live capture has pinned is
    allowed for eager calls (they get the new block; replays keep the old one), see ``check_scratch``."""
    key = (torch.device(device), slot)
    buf = _scratch.get(key)
    if buf is None or buf.numel() < nbytes:
        if buf is not None:
            _retired.append(buf)
        buf = torch.empty(int(nbytes * 1.25) + 1024, dtype=torch.uint8, device=device)
        _scratch[key] = buf
    return buf


def pin_scratch(device):
    """Called by a CUDA-graph capture after its warm-up: the scratch blocks of ``device`` as {slot: data_ptr}.  The
    blocks stay allocated for the life of the process (``scratch_for`` retires, never frees)."""
    device = torch.device(device)
    ptrs = {slot: buf.data_ptr() for (dev, slot), buf in _scratch.items() if dev == device}
    for slot, ptr in ptrs.items():
        _pinned[(device, slot)] = ptr
    return {'device': device, 'ptrs': ptrs, 'keep': [buf for (dev, _), buf in _scratch.items() if dev == device]}


def check_scratch(pinned):
    """Raise if a scratch block that a capture baked in is no longer alive (cannot happen through ``scratch_for``;
    guards against someone clearing the cache)."""
    if pinned is None:
        return
    alive = {b.data_ptr() for b in pinned['keep']}
    for slot, ptr in pinned['ptrs'].items():
        if ptr not in alive:
            raise RuntimeError(f'senas_b200: scratch block of slot {slot} captured by a CUDA graph was released')


def _nhwc(t):
    return t.contiguous(memory_format=torch.channels_last)


class GraphRunner:
    """``edges``: list of (candidate module list, src state, dst node, op_type id, c_in)."""

    def __init__(self, edges, n_inputs, n_nodes, node_relu, lib=None):
        self.lib = lib if lib is not None else _lib.get()
        self.edges, self.n_inputs, self.n_nodes, self.node_relu = edges, n_inputs, n_nodes, node_relu
        self.handle = None
        self._plans = {}
        # optional caller-owned destination of the flat parameter-gradient buffer (senas_b200.optim: a slice of the
        # gradient arena, so the cell's weight gradients are produced in place -- no per-call allocation, no packing copy)
        self.grad_buffer = None
        self._build()

    def _build(self):
        lib = self.lib
        desc = _lib.GraphDesc()
        desc.n_inputs, desc.n_nodes, desc.n_edges = self.n_inputs, self.n_nodes, len(self.edges)
        desc.c_out, desc.node_relu = 8, int(self.node_relu)
        self.tc_bf16 = _flags['tc_bf16']
        desc.reserved = 1 if self.tc_bf16 else 0  # bit 0 = bf16 mode (tcgen05 convs with bf16 operands)
        if _flags['override'] is not None:
            desc.reserved = int(_flags['override'])
        self.params, self.sizes, self.shapes = [], [], []
        off = 0
        for e, (cands, src, dst, op_type, c_in) in enumerate(self.edges):
            ed = desc.edge[e]
            ed.src, ed.dst, ed.op_type, ed.c_in = src, dst, op_type, c_in
            for k, op in enumerate(cands):
                kind, ks, dil, slots = candidate_slots(op)
                ed.kind[k], ed.ksize[k], ed.dilation[k] = kind, ks, dil
                for s, t in enumerate(slots):
                    ed.grad_off[k][s] = -1
                    if t is None:
                        ed.param[k][s] = None
                        continue
                    if not t.is_contiguous():
                        raise RuntimeError('senas_b200: parameters must be contiguous')
                    ed.param[k][s] = t.data_ptr()
                    if t.is_floating_point() and isinstance(t, torch.nn.Parameter):
                        ed.grad_off[k][s] = off
                        self.params.append(t)
                        self.sizes.append(t.numel())
                        self.shapes.append(t.shape)
                        off += t.numel()
        desc.grad_floats = off
        self.grad_floats = off
        self.device = self.params[0].device
        self._fingerprint = (self.params[0].data_ptr(), self.params[-1].data_ptr(), self.device)
        h = C.c_void_p()
        _lib.check(lib, lib.senas_graph_create(C.byref(desc), C.byref(h)))
        if self.handle is not None:
            lib.senas_graph_destroy(self.handle)
        self.handle, self._plans, self._desc = h, {}, desc

    def refresh(self):
        """Re-read parameter pointers if the module was moved (``.to()``) since the graph was built."""
        fp = (self.params[0].data_ptr(), self.params[-1].data_ptr(), self.params[0].device)
        if fp != self._fingerprint or self.tc_bf16 != _flags['tc_bf16']:
            self._build()

    def __del__(self):
        try:
            if self.handle is not None:
                self.lib.senas_graph_destroy(self.handle)
        except Exception:
            pass

    def plan(self, batch, hs, ws):
        key = (batch, tuple(hs), tuple(ws))
        info = self._plans.get(key)
        if info is None:
            info = _lib.PlanInfo()
            ih, iw = (C.c_int32 * 2)(*(list(hs) + [0])[:2]), (C.c_int32 * 2)(*(list(ws) + [0])[:2])
            _lib.check(self.lib, self.lib.senas_graph_plan(self.handle, batch, ih, iw, C.byref(info)))
            self._plans[key] = info
        return info

    # -- raw calls (tensors already NHWC fp32 on self.device) ---------------------------------
    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream if self.device.type == 'cuda' else 0

    def forward(self, ins, alpha, beta, training, slot=0):
        if self.device.type == 'cuda' and torch.cuda.current_device() != self.device.index:
            with torch.cuda.device(self.device):  # the library launches on, and allocates tables on, the current device
                return self.forward(ins, alpha, beta, training, slot)
        B = ins[0].shape[0]
        hs, ws = [t.shape[2] for t in ins], [t.shape[3] for t in ins]
        info = self.plan(B, hs, ws)
        out = torch.empty((B, 8 * self.n_nodes, info.out_h, info.out_w), dtype=torch.float32, device=self.device,
                          memory_format=torch.channels_last)
        saved = torch.empty(info.saved_bytes, dtype=torch.uint8, device=self.device)
        scratch = scratch_for(self.device, info.scratch_bytes, slot)
        a = _lib.FwdArgs()
        a.batch, a.training = B, int(training)
        for i, t in enumerate(ins):
            a.in_h[i], a.in_w[i], a.in_[i], a.in_ld[i] = t.shape[2], t.shape[3], t.data_ptr(), t.shape[1]
        a.alpha, a.beta = alpha.data_ptr(), (beta.data_ptr() if beta is not None else None)
        a.out, a.out_ld = out.data_ptr(), out.shape[1]
        a.saved, a.scratch, a.stream = saved.data_ptr(), scratch.data_ptr(), self._stream()
        self.lib.senas_set_slot(slot)
        held = _held.pop(slot, None)  # the call below first orders the slot's pending lanes before itself
        _lib.check(self.lib, self.lib.senas_graph_forward(self.handle, C.byref(a)))
        del held
        return out, saved

    def backward(self, ins, alpha, beta, out, grad_out, saved, training, need_in, slot=0):
        if self.device.type == 'cuda' and torch.cuda.current_device() != self.device.index:
            with torch.cuda.device(self.device):
                return self.backward(ins, alpha, beta, out, grad_out, saved, training, need_in, slot)
        B = ins[0].shape[0]
        hs, ws = [t.shape[2] for t in ins], [t.shape[3] for t in ins]
        info = self.plan(B, hs, ws)
        scratch = scratch_for(self.device, info.scratch_bytes, slot)
        n_edges = len(self.edges)
        g_alpha = torch.empty((n_edges, 6), dtype=torch.float32, device=self.device)
        g_beta = torch.empty((n_edges,), dtype=torch.float32, device=self.device) if beta is not None else None
        g_params = self.grad_buffer
        if g_params is None:
            g_params = torch.empty((self.grad_floats,), dtype=torch.float32, device=self.device)
        elif g_params.numel() != self.grad_floats or g_params.device != self.device or not g_params.is_contiguous():
            raise RuntimeError('senas_b200: GraphRunner.grad_buffer does not match the graph\'s gradient layout')
        g_ins = [torch.empty_like(t, memory_format=torch.channels_last) if need else None
                 for t, need in zip(ins, need_in)]
        a = _lib.BwdArgs()
        a.batch, a.training = B, int(training)
        for i, t in enumerate(ins):
            a.in_h[i], a.in_w[i], a.in_[i], a.in_ld[i] = t.shape[2], t.shape[3], t.data_ptr(), t.shape[1]
            if g_ins[i] is not None:
                a.grad_in[i], a.grad_in_ld[i] = g_ins[i].data_ptr(), t.shape[1]
        a.alpha, a.beta = alpha.data_ptr(), (beta.data_ptr() if beta is not None else None)
        a.out, a.out_ld = out.data_ptr(), out.shape[1]
        a.grad_out, a.grad_out_ld = grad_out.data_ptr(), grad_out.shape[1]
        a.saved, a.scratch = saved.data_ptr(), scratch.data_ptr()
        a.grad_alpha = g_alpha.data_ptr()
        a.grad_beta = g_beta.data_ptr() if g_beta is not None else None
        a.grad_params, a.stream = g_params.data_ptr(), self._stream()
        a.skip_wgrad = int(_skip_wgrad[0])
        self.lib.senas_set_slot(slot)
        held = _held.pop(slot, None)
        _lib.check(self.lib, self.lib.senas_graph_backward(self.handle, C.byref(a)))
        del held
        if _defer[0]:  # weight-gradient kernels of this call may still be reading these when we return
            _held[slot] = (ins, out, grad_out, saved, g_params)
        return g_ins, g_alpha, g_beta, g_params

    # -- autograd entry ------------------------------------------------------------------------
    def apply(self, ins, alpha, beta, training):
        """ins: list of [B,C,H,W] fp32 tensors; alpha [E,6]; beta [E] or None -> [B,8*nodes,Ho,Wo]."""
        self.refresh()
        ins = [_nhwc(t.float()) for t in ins]
        alpha = alpha.float().contiguous()
        beta = beta.float().contiguous() if beta is not None else None
        if torch.is_grad_enabled() and (any(t.requires_grad for t in ins) or alpha.requires_grad or
                                        any(p.requires_grad for p in self.params)):
            in1 = ins[1] if len(ins) > 1 else None
            return _GraphFn.apply(self, training, alpha, beta, ins[0], in1, *self.params)
        out, _ = self.forward(ins, alpha, beta, training, _slot[0])
        return out


class _GraphFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, runner, training, alpha, beta, in0, in1, *params):
        ins = [in0] if in1 is None else [in0, in1]
        ctx.slot = _slot[0]  # backward runs on the same stream (autograd) and must use the same lane set / scratch
        out, saved = runner.forward(ins, alpha, beta, training, ctx.slot)
        ctx.runner, ctx.training, ctx.saved_buf, ctx.has_in1 = runner, training, saved, in1 is not None
        ctx.has_beta = beta is not None
        ctx.save_for_backward(alpha, beta, in0, in1, out)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        alpha, beta, in0, in1, out = ctx.saved_tensors
        runner = ctx.runner
        if ctx.saved_buf is None:
            raise RuntimeError('senas_b200: backward through the same fused graph twice is not supported')
        ins = [in0] if in1 is None else [in0, in1]
        need_in = [ctx.needs_input_grad[4]] + ([ctx.needs_input_grad[5]] if in1 is not None else [])
        g_ins, g_alpha, g_beta, g_params = runner.backward(ins, alpha, beta, out, _nhwc(grad_out.float()),
                                                           ctx.saved_buf, ctx.training, need_in, ctx.slot)
        ctx.saved_buf = None  # the library overwrote parts of it (dz in place of z)
        if _skip_wgrad[0]:  # (grad_params is undefined: the caller asked for no parameter gradients)
            return (None, None, g_alpha, g_beta, g_ins[0], g_ins[1] if in1 is not None else None, *([None] * len(runner.params)))
        if _grad_sink[0] is not None:
            _grad_sink[0](runner, g_params)
        grads = [g.view(s) for g, s in zip(torch.split(g_params, runner.sizes), runner.shapes)]
        return (None, None, g_alpha, g_beta, g_ins[0], g_ins[1] if in1 is not None else None, *grads)
