"""``dice_ce`` segmentation loss of the search configuration (``SegmentationLosses('dice_ce')``,
reference utils/loss/loss.py:9-27, 45-70, 124-159, 173-228).  Outside the hot path: plain PyTorch.

With ``group`` set (data-parallel search, senas_b200.dp) the soft-dice statistics tp/fp/fn are
summed over all ranks (3 x C floats, exact global-batch dice as the reference computes on its
gathered output) and the cross entropy is divided by the world size, so that a SUM all-reduce of
the gradients equals the gradient of the global-batch loss.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class _NullCtx:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


fused_loss = [True]  # CUDA, <= 8 classes, no process group: libsenas_b200's dice_ce kernels


class _DiceCEFn(torch.autograd.Function):
    """mean CE + soft dice (classes >= 1) through ``senas_dice_ce_forward/backward``: 3 launches instead of ~15."""

    @staticmethod
    def forward(ctx, logits, target, smooth, lib):
        from . import _lib
        B, C = logits.shape[0], logits.shape[1]
        HW = logits[0, 0].numel()
        st = logits.stride()
        # pixels must be enumerable with ONE stride (NCHW-contiguous or channels_last both are)
        if logits.dim() != 4 or st[2] != logits.shape[3] * st[3]:
            logits = logits.contiguous()
            st = logits.stride()
        target = target.long().contiguous()
        dev = logits.device
        loss = torch.empty((), dtype=torch.float32, device=dev)
        coef = torch.empty(3 * C + 1, dtype=torch.float32, device=dev)
        scratch = torch.empty(296 * 25, dtype=torch.float32, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream if logits.is_cuda else 0
        with (torch.cuda.device(dev) if logits.is_cuda else _NullCtx()):  # the library launches on the current device
            _lib.check(lib, lib.senas_dice_ce_forward(logits.data_ptr(), target.data_ptr(), B, C, HW, st[0], st[1], st[3], 1.0,
                                                      float(smooth), loss.data_ptr(), coef.data_ptr(), scratch.data_ptr(),
                                                      stream))
        ctx.save_for_backward(logits, target, coef)
        ctx.lib, ctx.geo = lib, (B, C, HW, st[0], st[1], st[3])
        return loss

    @staticmethod
    def backward(ctx, g):
        from . import _lib
        logits, target, coef = ctx.saved_tensors
        B, C, HW, sn, sc, sp = ctx.geo
        d = torch.empty_like(logits)  # (preserves the strides of a dense tensor)
        if d.stride() != logits.stride():
            raise RuntimeError('senas_b200: dice_ce backward needs dense logits')
        g = g.float().contiguous()
        stream = torch.cuda.current_stream(logits.device).cuda_stream if logits.is_cuda else 0
        with (torch.cuda.device(logits.device) if logits.is_cuda else _NullCtx()):
            _lib.check(ctx.lib, ctx.lib.senas_dice_ce_backward(logits.data_ptr(), target.data_ptr(), B, C, HW, sn, sc, sp,
                                                               coef.data_ptr(), g.data_ptr(), d.data_ptr(), stream))
        return d, None, None, None


class DiceCrossEntropyLoss(nn.Module):
    def __init__(self, smooth=1e-5, group=None):
        super().__init__()
        self.smooth, self.group = smooth, group

    def forward(self, logits, target):
        if (fused_loss[0] and self.group is None and logits.is_cuda and logits.dtype == torch.float32 and logits.dim() == 4
                and 2 <= logits.shape[1] <= 8):
            from . import _lib
            return _DiceCEFn.apply(logits, target, self.smooth, _lib.get())
        prob = F.softmax(logits, 1)
        onehot = torch.zeros_like(prob).scatter_(1, target.long().unsqueeze(1), 1)
        axes = (0, 2, 3)
        stats = torch.stack([(prob * onehot).sum(axes), (prob * (1 - onehot)).sum(axes),
                             ((1 - prob) * onehot).sum(axes)])
        world = 1
        if self.group is not None:
            import torch.distributed as dist
            world = dist.get_world_size(self.group)
            total = stats.detach().clone()
            dist.all_reduce(total, group=self.group)
            stats = stats + (total - stats.detach())
        tp, fp, fn = stats[0], stats[1], stats[2]
        dc = (2 * tp + self.smooth) / (2 * tp + fp + fn + self.smooth + 1e-8)
        return F.cross_entropy(logits, target.long()) / world + (1 - dc[1:].mean())


class SegmentationLosses(nn.Module):
    """Same call convention as the reference: ``criterion(outputs_list, target)`` uses ``outputs[-1]``."""

    def __init__(self, name='dice_ce', group=None):
        super().__init__()
        if name == 'cross_entropy':
            self.loss = nn.CrossEntropyLoss()
        elif name == 'dice_ce':
            self.loss = DiceCrossEntropyLoss(group=group)
        else:
            raise NotImplementedError(name)

    def forward(self, outputs, target):
        return self.loss(outputs[-1], target)
