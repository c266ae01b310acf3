"""One fwd+bwd of the largest cell of the supernet (head up-cell: in0 32x256x256, in1 32x128x128, batch 16) --
the workload used for the ncu captures under profiles/ (the full bench command issues ~165k launches per run and
ncu's interception cost of ~13 ms per launch makes that impractical)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import senas_b200
mode = sys.argv[1] if len(sys.argv) > 1 else 'bf16'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
res = int(sys.argv[4]) if len(sys.argv) > 4 else 256   # node resolution of the up cell (in0 res x res, in1 res/2)
senas_b200.exact_fp32(); senas_b200.set_conv_mode(mode)
torch.manual_seed(0)
dev = 'cuda:0'
c = senas_b200.Cell(3, 1, 32, 32, 32, 'up'); c.apply(senas_b200.weights_init); c = c.to(dev)
in0 = torch.randn(B, 32, res, res, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
in1 = torch.randn(B, 32, res // 2, res // 2, device=dev).relu().contiguous(memory_format=torch.channels_last).requires_grad_(True)
wn, wc = torch.softmax(torch.randn(9, 6, device=dev), -1), torch.softmax(torch.randn(9, 6, device=dev), -1)
b = torch.softmax(torch.randn(9, device=dev), -1)
lib = senas_b200._lib.get()
for i in range(iters):
    if i == iters - 1: lib.senas_profile(1)
    out = c.nodes(in0, in1, wn, wc, b)
    out.backward(torch.ones_like(out))
torch.cuda.synchronize(); lib.senas_profile(0)
prof = senas_b200._lib.profile_dump(lib)
tot = sum(v['ms'] for v in prof.values())
print(f'mode={mode} B={B} total {tot:.2f} ms in {sum(v["launches"] for v in prof.values())} launches')
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]['ms']):
    print(f'{k:14s} {v["ms"]:8.3f} ms {100*v["ms"]/tot:5.1f}%  n={v["launches"]:4d}  {v["flops"]/(v["ms"]*1e-3)/1e12 if v["ms"] else 0:8.2f} TFLOP/s  {v["bytes"]/(v["ms"]*1e-3)/1e9 if v["ms"] else 0:8.1f} GB/s')
