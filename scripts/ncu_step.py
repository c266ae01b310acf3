"""ONE eagerly launched search step inside a cudaProfilerStart/Stop range, for `ncu --profile-from-start off`:
DRAM bytes and device time of every kernel of the step (libsenas_b200's and the stock blocks').

    python scripts/ncu_step.py [bf16|fp32] [B]          # plain run: prints the step's launch count
    ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
        --clock-control none --csv --log-file gpurun_out/step_traffic.csv python scripts/ncu_step.py bf16 16
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import senas_b200
from senas_b200.loss import SegmentationLosses

mode = sys.argv[1] if len(sys.argv) > 1 else 'bf16'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
senas_b200.exact_fp32(); senas_b200.set_conv_mode(mode); torch.backends.cudnn.benchmark = True
if mode == 'bf16':
    torch.backends.cudnn.allow_tf32 = True
lib = senas_b200._lib.get()
lib.senas_set_lanes(0)  # serial launches: ncu serialises anyway
dev = 'cuda:0'
torch.manual_seed(0)
m = senas_b200.NAS(1, 32, 2, depth=5, meta_node_num=3, use_sharing=False, double_down_channel=False, supervision=False).to(dev).train()
w = torch.optim.SGD(m.parameters(), lr=5e-3, momentum=0.9, weight_decay=3e-4)
a = torch.optim.Adam(m.arch_parameters(), lr=1e-4, betas=(0.5, 0.999), weight_decay=1e-3)
crit = SegmentationLosses('dice_ce')
g = torch.Generator().manual_seed(1)
xs = [torch.randn(B, 1, 256, 256, generator=g).to(dev) for _ in range(2)]
ys = [(torch.rand(B, 256, 256, generator=g) > 0.8).long().to(dev) for _ in range(2)]


# the product sequence (arena optimizer, fused mix / ConvBn blocks): the segments of GraphedSearchStep, launched eagerly
gs = senas_b200.GraphedSearchStep(m, crit, w, a, (xs[0], ys[0], xs[1], ys[1]), warmup=1, fused_optim=True)
gs.static = [xs[0], ys[0], xs[1], ys[1]]


def step():
    gs._run(capture=False, arch=True)


step()  # warm-up: plans, scratch, cudnn.benchmark
torch.cuda.synchronize()
n0 = lib.senas_launch_count()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(f'one search step, mode={mode} B={B}: {lib.senas_launch_count() - n0} libsenas_b200 launches')
