#!/usr/bin/env python
"""bench.py -- SENAS supernet search step on B200: images/sec, roofline of the dominant kernel, CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

One "step" = one search step of experiments/search_arc.py:252-293 (epoch >= alpha_begin) on synthetic
PROMISE12-shaped 1x256x256 slices: Architecture.step on a validation batch (fwd, dice_ce, bwd, Adam on
alpha/beta/gamma) followed by the weight step on a training batch (fwd, dice_ce, bwd, clip_grad_norm 5,
SGD) of the supernet NAS(1, 32, 2, depth=5, meta_node_num=3) -- configs/senas/senas_promise12.yml.
images/sec = training images per step / step time.  16 images per GPU at every N
(N=1: BASELINE config 2; N=8: global batch 128 = config 3), i.e. weak scaling.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# the captured step is a DAG over ~60 streams (lanes x concurrent cells): give the device as many hardware queues as it has
os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')

# algorithmic work per image (SURVEY.md section 8d / BASELINE.md section 3), search step = 2 x (fwd + bwd)
STEP_GFLOP_PER_IMG = 176.0          # whole supernet
STEP_GFLOP_PER_IMG_MIXED = 129.7    # MixedOp path only


def peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            p = json.load(f)
        return dict(hbm=p['hbm_gbs'], tf=p.get('bf16_tflops_sustained', p['bf16_tflops']), src='measured')
    except Exception:
        return dict(hbm=6650.0, tf=1590.0, src='fallback')


def synth(B, size, seed, pin):
    import torch
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 1, size, size, generator=g)
    y = (torch.rand(B, size, size, generator=g) > 0.8).long()
    return (x.pin_memory(), y.pin_memory()) if pin else (x, y)


# ------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference's CPU implementation
# ------------------------------------------------------------------------------------------------------
def cpu_search_step_factory(size, batch, seed=0):
    import torch
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import senas_oracle as oracle
    import senas_b200
    torch.manual_seed(seed)
    m = senas_b200.NAS(1, 32, 2, depth=5, meta_node_num=3, use_sharing=False, double_down_channel=False,
                       supervision=False)  # parameter container only: the arithmetic below is the oracle's
    store = dict(m.state_dict())
    names = [n for n, _ in m.named_parameters()]
    params = [store[n].requires_grad_(True) for n in names]
    arch = [store[n] for n in ('alphas_dn', 'alphas_up', 'alphas_dn_nm', 'alphas_up_nm', 'betas_dn', 'betas_up', 'gamma')]
    w_opt = torch.optim.SGD(params, lr=5e-3, momentum=0.9, weight_decay=3e-4)
    a_opt = torch.optim.Adam(arch, lr=1e-4, betas=(0.5, 0.999), weight_decay=1e-3)
    xt, yt = synth(batch, size, 1234, False)
    xv, yv = synth(batch, size, 4321, False)

    def step():
        a_opt.zero_grad()
        oracle.dice_ce_loss(oracle.nas_forward(store, xv)[-1], yv).backward()
        a_opt.step()
        w_opt.zero_grad()
        loss = oracle.dice_ce_loss(oracle.nas_forward(store, xt)[-1], yt)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 5)
        w_opt.step()
        return loss.item()

    return step


def run_reference(args, rank):
    if rank != 0:
        return
    import torch
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    batch = args.ref_batch
    step = cpu_search_step_factory(args.size, batch)
    for _ in range(min(args.warmup, 1)):
        step()
    steps = max(1, min(args.steps, args.ref_max_steps))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    v = batch / dt
    sample = f'{steps} search step(s) of the full supernet at batch {batch}, 1x{args.size}x{args.size}, fp32, oracle port'
    print(json.dumps({
        'impl': 'reference', 'metric': 'search_step_images_per_sec', 'value': v, 'unit': 'images/s', 'n_gpus': args.gpus,
        'steps': steps, 'warmup': min(args.warmup, 1), 'ms_per_step': dt * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'SENAS supernet search step (arch step + weight step), NAS(1,32,2,depth=5,nodes=3), '
                               f'{args.batch} x 1x{args.size}x{args.size} per GPU, global batch {args.batch * max(1, args.gpus)}',
                   'parallelism': 'host cpu', 'launch': f'PyTorch CPU, {cores} threads, bounded sample (see cpu_baseline.sample)'},
        'cpu_baseline': {'value': v, 'unit': 'images/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': v, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))


# ------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        try:
            self.p = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '100',
                                       '-i', str(index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return None
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], 0.0, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.f.read().splitlines():
            c = [t.strip() for t in line.split(',')]
            if len(c) < 7:
                continue
            try:
                sm.append(float(c[0]))
                mx = max(mx, float(c[1]))
            except ValueError:
                continue
            for n, v in zip(names, c[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        os.unlink(self.f.name)
        if not sm:
            return None
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': mx, 'reasons': sorted(reasons), 'samples': len(sm)}


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    # libraries print banners on stdout at fd level ("NCCL version ..."): keep stdout for the ONE JSON line
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    import senas_b200
    from senas_b200 import _lib
    from senas_b200.dp import FusedGradReducer, GradBuckets, broadcast_parameters
    from senas_b200.loss import SegmentationLosses

    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    lib = _lib.get()
    _lib.check(lib, lib.senas_device_check(local_rank))
    torch.backends.cudnn.benchmark = True  # as experiments/search_arc.py:72 (stems / pre / post convs)
    senas_b200.exact_fp32()                # no TF32 in the stock-PyTorch blocks around the cells
    senas_b200.set_conv_mode(args.conv_mode)
    if args.conv_mode == 'bf16':           # reduced-precision mode (gate 2e-2): the stock cuDNN convs may use TF32
        torch.backends.cudnn.allow_tf32 = True

    B, size = args.batch, args.size
    torch.manual_seed(0)
    model = senas_b200.NAS(1, 32, 2, depth=5, meta_node_num=3, use_sharing=False, double_down_channel=False,
                           supervision=False).to(dev)
    model.train()
    group = dist.group.WORLD if world > 1 else None
    crit = SegmentationLosses('dice_ce', group=group)
    w_opt = torch.optim.SGD(model.parameters(), lr=5e-3, momentum=0.9, weight_decay=3e-4)
    a_opt = torch.optim.Adam(model.arch_parameters(), lr=1e-4, betas=(0.5, 0.999), weight_decay=1e-3)
    buckets = None
    if world > 1:
        broadcast_parameters(model)

    def make_eager_reducers():                    # eager DP path: NCCL all-reduces overlapped with backward
        fused_red = FusedGradReducer()            # cells: flat gradient buffers straight from the kernels
        hooks = GradBuckets(list(model.parameters()), model.arch_parameters(), exclude=fused_red.owned(model))

        class _Both:                              # the rest (stems, pre/post blocks, arch parameters): hook buckets
            @staticmethod
            def finish():
                fused_red.finish()
                hooks.finish()
        return _Both

    if world > 1 and args.no_graph:
        buckets = make_eager_reducers()

    host = [synth(B, size, 1234 + 17 * rank + i, True) for i in range(4)]  # train0, valid0, train1, valid1
    devb = [(x.to(dev), y.to(dev)) for x, y in host]

    def search_step(xt, yt, xv, yv):
        a_opt.zero_grad()
        crit(model(xv), yv).backward()
        if buckets:
            buckets.finish()
        a_opt.step()
        w_opt.zero_grad()
        loss = crit(model(xt), yt)
        loss.backward()
        if buckets:
            buckets.finish()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 5)
        w_opt.step()
        return loss

    # The step is static-shaped and host-bound when launched eagerly (~9k kernels + 3.4k parameter tensors of
    # autograd/optimizer bookkeeping): capture arch step + weight step once and replay (senas_b200.GraphedSearchStep).
    graphed, graph_note, eager_step, launches_per_step = None, 'eager', search_step, None
    if not args.no_graph:
        try:
            n_before = lib.senas_launch_count()
            if world > 1:  # graphed DP path: local dice per rank, gradients averaged (senas_b200/graphs.py)
                crit = SegmentationLosses('dice_ce')
            graphed = senas_b200.GraphedSearchStep(model, crit, w_opt, a_opt, (*devb[0], *devb[1]), grad_clip=5.0,
                                                   warmup=3, group=group,
                                                   capture_error_mode='thread_local' if world > 1 else 'global',
                                                   concurrent_cells=not args.serial_cells, defer_wgrad=args.defer_wgrad)
            launches_per_step = (lib.senas_launch_count() - n_before) // 4   # 3 warm-up steps + 1 capture pass
            graph_note = 'cuda-graph (whole search step captured once, replayed per step)'
            search_step = lambda xt, yt, xv, yv: graphed(xt, yt, xv, yv)  # noqa: E731
        except Exception as e:  # keep measuring, but say so
            graph_note = f'eager (graph capture failed: {type(e).__name__}: {str(e)[:120]})'
            torch.cuda.synchronize()
            if world > 1:
                crit = SegmentationLosses('dice_ce', group=group)
                buckets = make_eager_reducers()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    resident = lambda i: search_step(*devb[(2 * i) % 4], *devb[(2 * i + 1) % 4])

    def e2e(i):
        (xt, yt), (xv, yv) = host[(2 * i) % 4], host[(2 * i + 1) % 4]
        loss = search_step(xt.to(dev, non_blocking=True), yt.to(dev, non_blocking=True),
                           xv.to(dev, non_blocking=True), yv.to(dev, non_blocking=True))
        return loss.item()

    for i in range(args.warmup):
        resident(i)
    clocks = ClockSampler(local_rank) if rank == 0 else None
    n0 = lib.senas_launch_count()
    ms = timed(resident, args.steps)
    launches = lib.senas_launch_count() - n0
    if graphed is not None:  # replayed graph nodes do not pass through the library's launch counter
        launches = launches_per_step * args.steps
    ms_e2e = timed(e2e, args.steps)
    clk = clocks.stop() if clocks else None

    # per-kernel-family device time (CUDA events on the launch stream), one more step
    lib.senas_profile(1)
    barrier()
    eager_step(*devb[0], *devb[1])   # eagerly launched so that every launch is bracketed by its own events
    barrier()
    lib.senas_profile(0)
    prof = _lib.profile_dump(lib)
    total_ms = sum(v['ms'] for v in prof.values()) or 1.0

    if rank != 0:
        return
    pk = peaks()
    gB = B * world
    ms_step, ms_step_e2e = ms / args.steps, ms_e2e / args.steps
    value, value_e2e = gB / (ms_step * 1e-3), gB / (ms_step_e2e * 1e-3)
    # dominant kernel = the family with the largest device time among those that carry algorithmic work (the 'reduce' /
    # finalize families are fixed-order partial-sum folds: overhead of the bit-reproducible reductions, no flops/bytes
    # of the reference graph; their share is reported under kernel_families like everything else)
    work = {k: v for k, v in prof.items() if v['flops'] > 0 or v['bytes'] > 0} or prof
    name, t = max(work.items(), key=lambda kv: kv[1]['ms'])
    tensor_bound = name.startswith('conv_')  # conv_fwd / conv_dgrad / conv_wgrad / conv_tc_*
    if tensor_bound:
        ach = t['flops'] / (t['ms'] * 1e-3) / 1e12
        roof = {'kernel': name, 'bound': 'tensor', 'achieved': ach, 'peak': pk['tf'], 'unit': 'TFLOP/s',
                'frac': ach / pk['tf'], 'traffic': None}
    else:
        ach = t['bytes'] / (t['ms'] * 1e-3) / 1e9
        roof = {'kernel': name, 'bound': 'hbm', 'achieved': ach, 'peak': pk['hbm'], 'unit': 'GB/s',
                'frac': ach / pk['hbm'], 'traffic': None}
    roof.update(peak_source=pk['src'], launches=t['launches'], avg_launch_ms=t['ms'] / max(1, t['launches']),
                share_of_senas_kernels=t['ms'] / total_ms,
                step_tensor_frac=STEP_GFLOP_PER_IMG * 1e9 * value / world / (pk['tf'] * 1e12))
    families = {k: {'ms': round(v['ms'], 3), 'launches': v['launches'],
                    'tflops': round(v['flops'] / (v['ms'] * 1e-3) / 1e12, 3) if v['ms'] > 0 else 0,
                    'gbs': round(v['bytes'] / (v['ms'] * 1e-3) / 1e9, 1) if v['ms'] > 0 else 0}
                for k, v in sorted(prof.items(), key=lambda kv: -kv[1]['ms'])}
    out = {
        'metric': 'search_step_images_per_sec', 'value': value, 'unit': 'images/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'bf16' if args.conv_mode == 'bf16' else 'f32', 'data': 'synthetic',
        'config': {'workload': f'SENAS supernet search step (arch step + weight step), NAS(1,32,2,depth=5,nodes=3), '
                               f'{B} x 1x{size}x{size} per GPU, global batch {gB}',
                   'parallelism': f'dp{world}', 'launch': graph_note, 'conv_mode': args.conv_mode + (' (tcgen05 bf16 operands, fp32 accumulate/storage)' if args.conv_mode == 'bf16' else ' (exact FMA)'), 'l2': 'no flush: each step streams several GB of activations (>> 126 MB L2)'},
        'e2e': {'value': value_e2e, 'unit': 'images/s', 'ms_per_step': ms_step_e2e,
                'h2d_bytes_per_step': 2 * B * size * size * (4 + 8), 'd2h_bytes_per_step': 4},
        'gpu_launches': int(launches), 'roofline': roof, 'kernel_families': families, 'clocks': clk,
    }
    if world == 1 and not args.no_cpu:
        import torch as _t
        cores = os.cpu_count()
        _t.set_num_threads(cores)
        cpu_search_step_factory(64, 1)()                      # thread-pool / allocator warm-up on a tiny input
        step = cpu_search_step_factory(size, args.ref_batch)
        t0 = time.perf_counter()
        step()
        step()
        dt = (time.perf_counter() - t0) / 2
        out['cpu_baseline'] = {'value': args.ref_batch / dt, 'unit': 'images/s', 'cores': cores, 'kind': 'port',
                               'sample': f'2 search steps of the full supernet at batch {args.ref_batch}, '
                                         f'1x{size}x{size}, fp32, all host threads, oracle port of the reference'}
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    print(json.dumps(out), flush=True)
    os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=16, help='images per GPU')
    ap.add_argument('--size', type=int, default=256)
    ap.add_argument('--ref-batch', type=int, default=4, help='CPU sample: images per (bounded) reference step')
    ap.add_argument('--ref-max-steps', type=int, default=4)
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='launch the step eagerly instead of replaying a CUDA graph')
    ap.add_argument('--serial-cells', action='store_true', help='do not run independent cells of a level on separate streams')
    ap.add_argument('--defer-wgrad', action='store_true', help='leave the weight-gradient lanes of a fused backward running (joined by the next call of the slot)')
    ap.add_argument('--conv-mode', default='bf16', choices=['fp32', 'bf16'],
                    help='bf16: tcgen05 implicit-GEMM convs with bf16 operands / fp32 accumulation; fp32: exact FMA path')
    args = ap.parse_args()
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        run_reference(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3
    run_ours(args, rank, world, local_rank)


if __name__ == '__main__':
    main()
