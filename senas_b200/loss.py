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


class DiceCrossEntropyLoss(nn.Module):
    def __init__(self, smooth=1e-5, group=None):
        super().__init__()
        self.smooth, self.group = smooth, group

    def forward(self, logits, target):
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
