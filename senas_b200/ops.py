"""Candidate-operation zoo of the SENAS supernet: parameter containers only.

Mirrors the public surface of the reference's ``utils/operations.py`` (``OPS`` registry
``:8-21``, candidate lists ``:23-48``, ``OpType`` ``:51-54``, ``build_ops`` ``:57-78``) so that
``state_dict`` keys, parameter order and fixed-seed initialisation are identical, but the
candidate blocks here own **no arithmetic**: the hot path (every candidate inside a
``MixedOp``) is evaluated by the sm_100a kernels behind ``libsenas_b200.so``.  Calling a
candidate block directly raises -- there is deliberately no PyTorch fallback.

The blocks that are *outside* the hot path (``ShrinkBlock``, ``RectifyBlock``, ``ReLUConv``,
the stem ``BasicBlock``) are ordinary ``torch.nn`` modules.
"""
from enum import Enum

import torch
import torch.nn as nn

DownOps = ['avg_pool', 'se_conv_3', 'dil_3_conv_5', 'dil_2_conv_5', 'dep_sep_conv_3', 'dep_sep_conv_5']
UpOps = ['up_sample', 'se_conv_3', 'dil_3_conv_5', 'dil_2_conv_5', 'dep_sep_conv_3', 'dep_sep_conv_5']
NormOps = ['identity', 'none', 'dil_3_conv_5', 'dil_2_conv_5', 'dep_sep_conv_3', 'dep_sep_conv_5']


class OpType(Enum):
    """Same member names and ``value`` payloads as the reference enum (operations.py:51-54)."""
    UP = {'id': 1, 'ops': UpOps}
    DOWN = {'id': 2, 'ops': DownOps}
    NORM = {'id': 3, 'ops': NormOps}


# candidate "kind" ids shared with the C-ABI (include/senas_b200.h, SENAS_KIND_*)
KIND_NONE, KIND_IDENTITY, KIND_AVG_POOL, KIND_UP_SAMPLE = 0, 1, 2, 3
KIND_CONV, KIND_SE_CONV, KIND_DEPSEP = 4, 5, 6


def same_padding(kernel_size, dilation=1):
    """``get_same_padding`` (utils/utils.py:21-29) times dilation (operations.py:119-120)."""
    if kernel_size % 2 == 0:
        raise AssertionError('kernel size should be odd number')
    return (kernel_size // 2) * dilation


def _geometry(op_type):
    stride = 1 if op_type == OpType.NORM else 2
    transposed = op_type == OpType.UP
    return stride, transposed, (1 if transposed else 0)


def _weight(c_in, c_ot, k, stride, dilation, transposed, out_pad, groups=1):
    pad = same_padding(k, dilation)
    if transposed:
        return nn.ConvTranspose2d(c_in, c_ot, k, stride=stride, padding=pad, dilation=dilation, bias=False,
                                  output_padding=out_pad, groups=groups)
    return nn.Conv2d(c_in, c_ot, k, stride=stride, padding=pad, dilation=dilation, bias=False, groups=groups)


class _KernelOnly:
    """Mixin: the block is a parameter store for the CUDA path, not a callable."""

    def forward(self, *a, **k):  # noqa: D401
        raise RuntimeError(
            f'{type(self).__name__} is evaluated inside the fused senas_b200 MixedOp kernels; '
            'it has no standalone (PyTorch) forward')


class ZeroOp(_KernelOnly, nn.Module):
    def __init__(self, stride=1):
        super().__init__()
        self.stride = stride


class SEBlock(_KernelOnly, nn.Module):
    """Squeeze-excite parameter store; ``mid = 1`` for c = 8 (operations.py:186-203)."""

    def __init__(self, c, r=16):
        super().__init__()
        self.mid = c // r if c > r else 1
        self.squeeze = nn.AdaptiveAvgPool2d(1)
        self.excitation = nn.Sequential(nn.Linear(c, self.mid, bias=False), nn.ReLU(inplace=True),
                                        nn.Linear(self.mid, c, bias=False), nn.Sigmoid())


class AdapterBlock(_KernelOnly, nn.Module):
    """pool / upsample / identity / zero -> optional 1x1 conv -> BN (operations.py:167-183)."""

    def __init__(self, c_in, c_ot, module, kind):
        super().__init__()
        self.c_in, self.c_ot, self.kind = c_in, c_ot, kind
        self.module = module
        if c_in != c_ot:
            self.conv = nn.Conv2d(c_in, c_ot, kernel_size=1, bias=False)
        self.norm = nn.BatchNorm2d(c_ot, affine=True)


class ConvBn(_KernelOnly, nn.Sequential):
    kind = KIND_CONV

    def __init__(self, c_in, c_ot, k, op_type, dilation=1):
        stride, tr, op = _geometry(op_type)
        super().__init__(_weight(c_in, c_ot, k, stride, dilation, tr, op), nn.BatchNorm2d(c_ot))
        self.k, self.dilation = k, dilation


class ConvBnSe(_KernelOnly, nn.Sequential):
    kind = KIND_SE_CONV

    def __init__(self, c_in, c_ot, k, op_type, dilation=1):
        stride, tr, op = _geometry(op_type)
        super().__init__(_weight(c_in, c_ot, k, stride, dilation, tr, op), nn.BatchNorm2d(c_ot), SEBlock(c_ot))
        self.k, self.dilation = k, dilation


class DepSepConv(_KernelOnly, nn.Sequential):
    kind = KIND_DEPSEP

    def __init__(self, c_in, c_ot, k, op_type):
        stride, tr, op = _geometry(op_type)
        super().__init__(_weight(c_in, c_in, k, stride, 1, tr, op, groups=c_in), nn.BatchNorm2d(c_in),
                         nn.ReLU(inplace=True), _weight(c_in, c_ot, 1, 1, 1, False, 0), nn.BatchNorm2d(c_ot))
        self.k, self.dilation = k, 1


def build_ops(op_name, op_type, c_in=None, c_ot=None, dp=0):
    """Candidate factory; same names / argument meaning / error as operations.py:57-78."""
    if dp:
        raise NotImplementedError('dropout inside candidates is not part of the search path (dp=0 at cell.py:29)')
    stride = 1 if op_type == OpType.NORM else 2
    if op_name == 'avg_pool':
        return AdapterBlock(c_in, c_ot, nn.AvgPool2d(3, stride=stride, padding=1, count_include_pad=False),
                            KIND_AVG_POOL)
    if op_name == 'se_conv_3':
        return ConvBnSe(c_in, c_ot, 3, op_type)
    if op_name == 'dil_3_conv_5':
        return ConvBn(c_in, c_ot, 5, op_type, dilation=3)
    if op_name == 'dil_2_conv_5':
        return ConvBn(c_in, c_ot, 5, op_type, dilation=2)
    if op_name == 'dep_sep_conv_3':
        return DepSepConv(c_in, c_ot, 3, op_type)
    if op_name == 'dep_sep_conv_5':
        return DepSepConv(c_in, c_ot, 5, op_type)
    raise NotImplementedError()


def _not_on_search_path(name):
    def make(c_in, c_ot, op_type, dp):
        raise NotImplementedError(f"'{name}' is registered by the reference (utils/operations.py:12,15) but is in none of its "
                                  'candidate lists (DownOps / UpOps / NormOps, :23-48); senas_b200 has no kernel for it')
    return make


OPS = {
    'none': lambda c_in, c_ot, op_type, dp: AdapterBlock(c_in, c_ot, ZeroOp(stride=1), KIND_NONE),
    'max_pool': _not_on_search_path('max_pool'),
    'conv_3': _not_on_search_path('conv_3'),
    'identity': lambda c_in, c_ot, op_type, dp: AdapterBlock(c_in, c_ot, nn.Identity(), KIND_IDENTITY),
    'avg_pool': lambda c_in, c_ot, op_type, dp: build_ops('avg_pool', op_type, c_in, c_ot),
    'up_sample': lambda c_in, c_ot, op_type, dp: AdapterBlock(
        c_in, c_ot, nn.Upsample(scale_factor=2, mode='bilinear', align_corners=False), KIND_UP_SAMPLE),
    'se_conv_3': lambda c_in, c_ot, op_type, dp: build_ops('se_conv_3', op_type, c_in, c_ot, dp=dp),
    'dil_3_conv_5': lambda c_in, c_ot, op_type, dp: build_ops('dil_3_conv_5', op_type, c_in, c_ot, dp=dp),
    'dil_2_conv_5': lambda c_in, c_ot, op_type, dp: build_ops('dil_2_conv_5', op_type, c_in, c_ot, dp=dp),
    'dep_sep_conv_3': lambda c_in, c_ot, op_type, dp: build_ops('dep_sep_conv_3', op_type, c_in, c_ot, dp=dp),
    'dep_sep_conv_5': lambda c_in, c_ot, op_type, dp: build_ops('dep_sep_conv_5', op_type, c_in, c_ot, dp=dp),
}


# ------------------------------------------------------------------------------------------
# Blocks around the hot path (rows f1/f3 of SURVEY.md section 8: stock PyTorch for now)
# ------------------------------------------------------------------------------------------
class ReLUConv(nn.Sequential):
    def __init__(self, c_in, c_ot, kernel_size=3):
        super().__init__(nn.ReLU(inplace=False), _weight(c_in, c_ot, kernel_size, 1, 1, False, 0))


class StemConvBn(nn.Sequential):
    """The 7x7 stem (``ConvBn(in_channels, c, kernel_size=7)``, senas_search.py:30)."""

    def __init__(self, c_in, c_ot, kernel_size):
        super().__init__(_weight(c_in, c_ot, kernel_size, 1, 1, False, 0), nn.BatchNorm2d(c_ot))

    def forward(self, x):
        # a 1-channel input is NCHW and NHWC at once and cuDNN then answers in NCHW, which sends the whole head of the
        # network (this BatchNorm at 256 x 256, the max-pool, stem1) down the slow NCHW kernels (measured: 16 ms of a
        # 230 ms step in cudnn::bn_*_1C11 alone); pin the activation layout the cells use.
        return self[1](self[0](x).contiguous(memory_format=torch.channels_last))


class _ConvBnFn(torch.autograd.Function):
    """[ReLU ->] Conv2d 3x3 (c_in -> 32) -> BatchNorm2d through ``senas_convbn_forward/backward`` (SURVEY row f1,
    utils/operations.py:206-232): tcgen05 implicit GEMM with the BatchNorm statistics out of the conv epilogue."""

    @staticmethod
    def forward(ctx, x, weight, gamma, beta, norm, relu_in):
        import ctypes as C
        from . import _lib, fused
        lib = _lib.get()
        x = x.contiguous(memory_format=torch.channels_last)
        B, c_in, H, W = x.shape
        sb, scb = C.c_int64(), C.c_int64()
        _lib.check(lib, lib.senas_convbn_workspace(B, H, W, c_in, C.byref(sb), C.byref(scb)))
        dev = x.device
        out = torch.empty((B, 32, H, W), dtype=torch.float32, device=dev, memory_format=torch.channels_last)
        saved = torch.empty(sb.value, dtype=torch.uint8, device=dev)
        slot = fused.get_slot()
        scratch = fused.scratch_for(dev, scb.value, ('convbn', slot))
        a = _lib.ConvBnArgs()
        a.batch, a.h, a.w, a.c_in, a.relu_in, a.training = B, H, W, c_in, int(relu_in), int(norm.training)
        a.x, a.x_ld = x.data_ptr(), c_in
        a.weight, a.gamma, a.beta = weight.data_ptr(), gamma.data_ptr(), beta.data_ptr()
        a.running_mean, a.running_var = norm.running_mean.data_ptr(), norm.running_var.data_ptr()
        a.num_batches_tracked = norm.num_batches_tracked.data_ptr()
        a.momentum, a.eps = float(norm.momentum), float(norm.eps)
        a.out, a.saved, a.scratch = out.data_ptr(), saved.data_ptr(), scratch.data_ptr()
        a.stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            _lib.check(lib, lib.senas_convbn_forward(C.byref(a)))
        ctx.save_for_backward(x, weight, gamma, beta)
        ctx.saved_buf, ctx.slot, ctx.relu_in, ctx.training = saved, slot, int(relu_in), int(norm.training)
        ctx.bn = (float(norm.momentum), float(norm.eps))
        return out

    @staticmethod
    def backward(ctx, g):
        import ctypes as C
        from . import _lib, fused
        lib = _lib.get()
        x, weight, gamma, beta = ctx.saved_tensors
        g = g.contiguous(memory_format=torch.channels_last)
        B, c_in, H, W = x.shape
        dev = x.device
        sb, scb = C.c_int64(), C.c_int64()
        _lib.check(lib, lib.senas_convbn_workspace(B, H, W, c_in, C.byref(sb), C.byref(scb)))
        scratch = fused.scratch_for(dev, scb.value, ('convbn', ctx.slot))
        gx = torch.empty_like(x, memory_format=torch.channels_last) if ctx.needs_input_grad[0] else None
        skip_w = fused._skip_wgrad[0]  # architecture step: data gradient only
        gw = None if skip_w else torch.empty_like(weight)
        gg, gb = torch.empty_like(gamma), torch.empty_like(beta)
        a = _lib.ConvBnArgs()
        a.batch, a.h, a.w, a.c_in, a.relu_in, a.training = B, H, W, c_in, ctx.relu_in, ctx.training
        a.x, a.x_ld = x.data_ptr(), c_in
        a.weight, a.gamma, a.beta = weight.data_ptr(), gamma.data_ptr(), beta.data_ptr()
        a.momentum, a.eps = ctx.bn
        a.saved, a.scratch = ctx.saved_buf.data_ptr(), scratch.data_ptr()
        a.grad_out, a.grad_out_ld = g.data_ptr(), 32
        a.grad_x = gx.data_ptr() if gx is not None else None
        a.grad_weight = gw.data_ptr() if gw is not None else None
        a.grad_gamma, a.grad_beta = gg.data_ptr(), gb.data_ptr()
        a.stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            _lib.check(lib, lib.senas_convbn_backward(C.byref(a)))
        if skip_w:
            return gx, None, None, None, None, None
        return gx, gw, gg, gb, None, None


def _convbn_on_tensor_cores(x, conv, norm):
    """The fused tcgen05 path covers the bf16 conv mode (2e-2 gate), maps that are a multiple of 64 pixels wide and the
    channel counts of the search supernet; everything else (exact fp32 mode, narrow maps) keeps PyTorch's modules."""
    from . import fused
    return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and fused.get_conv_mode() == 'bf16' and fused_convbn[0]
            and x.shape[3] % 64 == 0 and conv.out_channels == 32 and conv.in_channels in (24, 32, 64, 96, 128)
            and norm.affine and norm.track_running_stats and norm.momentum is not None)


import os as _os
stem_convbn = [_os.environ.get('SENAS_STEM_CONVBN', '0') == '1']  # stem1's BasicBlock through the same op (bf16 mode)
fused_convbn = [_os.environ.get('SENAS_NO_CONVBN', '0') != '1']  # switch (A/B runs, tests): False keeps cuDNN / ATen for the Shrink / Rectify blocks in every mode


class ShrinkBlock(nn.Module):
    def __init__(self, c_in, c_ot):
        super().__init__()
        self.act = nn.ReLU(inplace=False)
        self.conv = nn.Conv2d(c_in, c_ot, kernel_size=3, padding=1, bias=False)
        self.norm = nn.BatchNorm2d(c_ot)

    def forward(self, x):
        if _convbn_on_tensor_cores(x, self.conv, self.norm):
            return _ConvBnFn.apply(x, self.conv.weight, self.norm.weight, self.norm.bias, self.norm, True)
        return self.norm(self.conv(self.act(x)))


class RectifyBlock(nn.Module):
    def __init__(self, c_in, c_ot, cell_type='down'):
        super().__init__()
        self.cell_type = cell_type
        self.conv = nn.Conv2d(c_in, c_ot, kernel_size=3, padding=1, bias=False)
        self.norm = nn.BatchNorm2d(c_ot)

    def forward(self, x):
        if _convbn_on_tensor_cores(x, self.conv, self.norm):
            return _ConvBnFn.apply(x, self.conv.weight, self.norm.weight, self.norm.bias, self.norm, False)
        return self.norm(self.conv(x))


class _AvgPool2dNCHW(nn.AvgPool2d):
    """``nn.AvgPool2d`` evaluated on an NCHW-contiguous input.  PyTorch 2.11's CUDA avg_pool2d *backward*
    returns wrong gradients for channels_last inputs (measured on B200: max abs error 0.25-0.5 against the
    CPU result for k=3, s=2, p=1, either ``count_include_pad``; the NCHW path is exact -- see
    scripts/diag_pool.py), and the rest of the network runs channels_last."""

    def forward(self, x):
        if x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] % 4 == 0:
            return _AvgPoolNHWCFn.apply(x)  # libsenas_b200 kernels on the channels_last tensor, no layout round trip
        return super().forward(x.contiguous()).contiguous(memory_format=torch.channels_last)


class _AvgPoolNHWCFn(torch.autograd.Function):
    """AvgPool2d(3, stride 2, padding 1, count_include_pad=False) on channels_last fp32 CUDA tensors through
    ``senas_avgpool_forward / backward`` (row f1 of the scope table)."""

    @staticmethod
    def forward(ctx, x):
        from . import _lib
        lib = _lib.get()
        x = x.contiguous(memory_format=torch.channels_last)
        B, C, H, W = x.shape
        y = torch.empty((B, C, (H + 1) // 2, (W + 1) // 2), dtype=x.dtype, device=x.device,
                        memory_format=torch.channels_last)
        _lib.check(lib, lib.senas_avgpool_forward(x.data_ptr(), C, y.data_ptr(), B, H, W, C,
                                                  torch.cuda.current_stream(x.device).cuda_stream))
        ctx.shape = (B, C, H, W)
        return y

    @staticmethod
    def backward(ctx, gy):
        from . import _lib
        lib = _lib.get()
        B, C, H, W = ctx.shape
        gy = gy.contiguous(memory_format=torch.channels_last)
        gx = torch.empty((B, C, H, W), dtype=gy.dtype, device=gy.device, memory_format=torch.channels_last)
        _lib.check(lib, lib.senas_avgpool_backward(gy.data_ptr(), gx.data_ptr(), B, H, W, C,
                                                   torch.cuda.current_stream(gy.device).cuda_stream))
        return gx


def build_rectify(c_in, c_ot, cell_type):
    act = nn.ReLU(inplace=False)
    if cell_type == 'up':
        if c_in == c_ot:
            return nn.Sequential(act, nn.Upsample(scale_factor=2, mode='bilinear', align_corners=False),
                                 nn.BatchNorm2d(c_ot))
        return nn.Sequential(act, nn.ConvTranspose2d(c_in, c_ot, kernel_size=1, stride=2, output_padding=1,
                                                     bias=False), nn.BatchNorm2d(c_ot))
    if c_in == c_ot:
        return nn.Sequential(act, _AvgPool2dNCHW(3, stride=2, padding=1, count_include_pad=False),
                             nn.BatchNorm2d(c_ot))
    return nn.Sequential(act, nn.Conv2d(c_in, c_ot, kernel_size=1, stride=2, bias=False), nn.BatchNorm2d(c_ot))


class BasicBlock(nn.Module):
    """ResNet basic block of ``stem1`` (operations.py:235-268); out of the hot path."""
    expansion = 1

    def __init__(self, inplanes, planes):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, kernel_size=3, stride=1, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, stride=1, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)

    def forward(self, x):
        if stem_convbn[0] and _convbn_on_tensor_cores(x, self.conv1, self.bn1) and _convbn_on_tensor_cores(x, self.conv2, self.bn2):
            # the same library op as the Shrink / Rectify blocks: conv1 + bn1, then ReLU (folded into the bf16 cast and the
            # data-gradient epilogue) + conv2 + bn2; the residual add stays a PyTorch kernel
            h = _ConvBnFn.apply(x, self.conv1.weight, self.bn1.weight, self.bn1.bias, self.bn1, False)
            out = _ConvBnFn.apply(h, self.conv2.weight, self.bn2.weight, self.bn2.bias, self.bn2, True)
            return out + x
        out = self.bn2(self.conv2(self.relu(self.bn1(self.conv1(x)))))
        out += x
        return out


def weights_init(m):
    """Same initialisation rule as utils/utils.py:240-250."""
    if isinstance(m, nn.Linear):
        nn.init.xavier_normal_(m.weight)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
    elif isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
        nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
    elif isinstance(m, nn.BatchNorm2d):
        nn.init.constant_(m.weight, 1)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
