"""Where does the search step go, cell by cell?  Times one fwd+bwd of every distinct (cell type, resolution) of the
supernet as a replayed CUDA graph (so launch gaps are what they are inside the real captured step), prints the
per-family device profile of a few of them, and the implied total over the 15 cells x 2 passes of a search step.

    python scripts/profile_cells.py [bf16|fp32] [B] [families: comma list of cell names or 'none']
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import senas_b200

mode = sys.argv[1] if len(sys.argv) > 1 else 'bf16'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
fam = sys.argv[3].split(',') if len(sys.argv) > 3 else ['up128', 'up32']
senas_b200.exact_fp32(); senas_b200.set_conv_mode(mode)
dev = 'cuda:0'
lib = senas_b200._lib.get()
# (name, cell type, node resolution, multiplicity in NAS(depth=5))
CELLS = [('up256', 'up', 256, 1), ('up128', 'up', 128, 4), ('up64', 'up', 64, 3), ('up32', 'up', 32, 2),
         ('up16', 'up', 16, 1), ('dn64', 'down', 64, 1), ('dn32', 'down', 32, 1), ('dn16', 'down', 16, 1),
         ('dn8', 'down', 8, 1)]
total = 0.0
for name, kind, res, mult in CELLS:
    torch.manual_seed(0)
    c = senas_b200.Cell(3, 1, 32, 32, 32, kind); c.apply(senas_b200.weights_init); c = c.to(dev)
    r0 = res if kind == 'up' else 2 * res   # in0 is already pre-processed to in1's geometry rules:
    r1 = res // 2 if kind == 'up' else 2 * res  # up: in0 at node res, in1 at res/2; down: both at 2*res
    in0 = torch.randn(B, 32, r0, r0, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    in1 = torch.randn(B, 32, r1, r1, device=dev).relu().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    wn, wc = torch.softmax(torch.randn(9, 6, device=dev), -1), torch.softmax(torch.randn(9, 6, device=dev), -1)
    b = torch.softmax(torch.randn(9, device=dev), -1)
    go = None

    def fb():
        out = c.nodes(in0, in1, wn, wc, b)
        g = go if go is not None else torch.ones_like(out)
        out.backward(g)
        return out

    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            in0.grad = in1.grad = None
            out = fb()
        go = torch.ones_like(out)
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    n0 = lib.senas_launch_count()
    graph = torch.cuda.CUDAGraph()
    in0.grad = in1.grad = None
    with torch.cuda.graph(graph):
        fb()
    nl = lib.senas_launch_count() - n0
    for _ in range(2): graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    it = 5 if res >= 128 else 20
    e0.record()
    for _ in range(it): graph.replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / it
    total += ms * mult
    print(f'{name:6s} x{mult}  graph fwd+bwd {ms:8.3f} ms   {nl} launches   {1e3*ms/max(nl,1):6.2f} us/launch', flush=True)
    if name in fam:
        lib.senas_profile(1)
        in0.grad = in1.grad = None
        fb()
        torch.cuda.synchronize(); lib.senas_profile(0)
        prof = senas_b200._lib.profile_dump(lib)
        tot = sum(v['ms'] for v in prof.values())
        print(f'   eager per-family profile of {name}: {tot:.2f} ms in {sum(v["launches"] for v in prof.values())} launches')
        for k, v in sorted(prof.items(), key=lambda kv: -kv[1]['ms']):
            print(f'   {k:22s} {v["ms"]:8.3f} ms {100*v["ms"]/tot:5.1f}%  n={v["launches"]:4d}  '
                  f'{v["flops"]/(v["ms"]*1e-3)/1e12 if v["ms"] else 0:8.2f} TFLOP/s  '
                  f'{v["bytes"]/(v["ms"]*1e-3)/1e9 if v["ms"] else 0:8.1f} GB/s')
    del c, in0, in1, graph
print(f'cells of one pass: {total:.1f} ms; search step (2 passes): {2*total:.1f} ms')
