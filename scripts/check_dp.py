"""torchrun --nproc-per-node 2 scripts/check_dp.py : data-parallel gradient check on real GPUs.
Gradients produced by the overlapped reducers (FusedGradReducer + GradBuckets) must equal a plain all-reduce (SUM)
of the per-rank local gradients of the same shards."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import senas_b200
from senas_b200 import fused
from senas_b200.dp import FusedGradReducer, GradBuckets, broadcast_parameters
from senas_b200.loss import SegmentationLosses
rank, world, lr = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
dev = torch.device('cuda', lr); torch.cuda.set_device(dev)
dist.init_process_group('nccl', device_id=dev)
senas_b200.exact_fp32()
torch.manual_seed(0)
m = senas_b200.NAS(1, 32, 2, depth=5, meta_node_num=3, use_sharing=False, double_down_channel=False, supervision=False).to(dev)
broadcast_parameters(m)
crit = SegmentationLosses('dice_ce', group=dist.group.WORLD)
g = torch.Generator().manual_seed(100 + rank)
x = torch.randn(2, 1, 64, 64, generator=g).to(dev); y = (torch.rand(2, 64, 64, generator=g) > 0.8).long().to(dev)
# reference: local gradients, then one plain all-reduce per tensor
m.zero_grad(); crit(m(x), y).backward()
ref = []
for p in m.parameters():
    t = p.grad.detach().clone(); dist.all_reduce(t); ref.append(t)
# overlapped reducers
fr = FusedGradReducer()
hb = GradBuckets(list(m.parameters()), m.arch_parameters(), exclude=fr.owned(m))
for it in range(2):
    m.zero_grad(); crit(m(x), y).backward(); fr.finish(); hb.finish()
err = max(((p.grad - r).abs().max() / r.abs().max().clamp_min(1e-12)).item() for p, r in zip(m.parameters(), ref))
aliased = sum(1 for p in m.parameters() if p.grad is not None)
print(f'rank {rank}: max relative difference overlapped-vs-plain all-reduce = {err:.2e} over {aliased} tensors')
assert err < 1e-5
dist.destroy_process_group()
