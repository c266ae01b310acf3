#!/usr/bin/env python
"""bench.py -- SENAS supernet search step on B200: images/sec, roofline of the dominant kernel, CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|torch_eager]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Arms: ``ours`` (default) = senas_b200 on the GPU(s); ``reference`` = the UNMODIFIED reference (oracle/_ref, staged by
oracle/make_ref.py) on the box's host cores -- the reported CPU baseline; ``torch_eager`` = the same unmodified
reference module tree on ONE B200 through stock PyTorch (fp32, cudnn.benchmark=True as experiments/search_arc.py:72),
eagerly launched and as a replayed CUDA graph -- "the kernel to beat" of SURVEY.md section 2.3 / BASELINE.md 4.2.  The default
run also embeds a short torch_eager measurement (``reference_b200``) and an fp32-mode line (``fp32_mode``).

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
STEP_MB_PER_IMG = 390.4             # MixedOp-boundary bytes, bf16 activations


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
# reference arms: the UNMODIFIED reference (oracle/ref_env.py finds it: /root/reference here, oracle/_ref on the GPU box)
# ------------------------------------------------------------------------------------------------------
WORKLOAD = ('SENAS supernet search step (arch step + weight step), NAS(1,32,2,depth=5,nodes=3), '
            '{B} x 1x{S}x{S} per GPU, global batch {G}')


def ref_env():
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import ref_env as re_
    return re_


def reference_step_factory(size, batch, device='cpu', seed=0):
    """search step of experiments/search_arc.py:252-293 on the reference's own NAS / Architecture / SegmentationLosses
    classes with the optimizers of configs/senas/senas_promise12.yml; synthetic batches as for our arm."""
    import contextlib
    import io
    re_ = ref_env()
    with contextlib.redirect_stdout(io.StringIO()):   # the reference prints "Using loss: ..." on stdout
        model = re_.make_nas(seed=seed).to(device)
        model.train()
        step, w_opt, a_opt = re_.make_search_step(model)
    xt, yt = synth(batch, size, 1234, False)
    xv, yv = synth(batch, size, 4321, False)
    data = [t.to(device) for t in (xt, yt, xv, yv)]
    return step, data, model, (w_opt, a_opt)


def cpu_reference_sample(size, batch, steps, warmup):
    """(seconds per step, what ran) of the reference on all host threads."""
    import torch
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    step, data, _, _ = reference_step_factory(size, batch)
    for _ in range(warmup):
        step(*data)
    t0 = time.perf_counter()
    for _ in range(steps):
        step(*data)
    return (time.perf_counter() - t0) / steps, cores


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation (PyTorch CPU kernels) of the same workload, all host
    threads.  A step is a whole search step at the full batch unless K + W such steps would exceed --ref-budget-s, in
    which case the per-step sample shrinks to the batch that fits -- and config.workload says which batch ran."""
    if rank != 0:
        return
    import torch
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    batch = args.ref_batch if args.ref_batch > 0 else args.batch
    step, data, _, _ = reference_step_factory(args.size, batch)
    t0 = time.perf_counter()
    step(*data)                                   # warm-up step 1 (also the probe that sizes the sample)
    t_probe = time.perf_counter() - t0
    total = args.steps + args.warmup
    if args.ref_batch <= 0 and t_probe * total > args.ref_budget_s:
        batch = max(1, int(batch * args.ref_budget_s / (t_probe * total)))
        step, data, _, _ = reference_step_factory(args.size, batch)
        step(*data)
    for _ in range(max(0, args.warmup - 1)):
        step(*data)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(*data)
    dt = (time.perf_counter() - t0) / args.steps
    v = batch / dt
    re_ = ref_env()
    sample = (f'{args.steps} search steps (+{args.warmup} warm-up) of the full supernet at batch {batch}, '
              f'1x{args.size}x{args.size}, fp32, {cores} threads; unmodified reference from {re_.kind()}')
    print(json.dumps({
        'impl': 'reference', 'metric': 'search_step_images_per_sec', 'value': v, 'unit': 'images/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD.format(B=batch, S=args.size, G=batch),
                   'parallelism': 'host cpu', 'launch': f'PyTorch CPU (oneDNN), {cores} threads',
                   'sample_batch': batch, 'full_batch': args.batch},
        'cpu_baseline': {'value': v, 'unit': 'images/s', 'cores': cores, 'kind': 'reference', 'sample': sample},
        'e2e': {'value': v, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))


def torch_eager_measure(args, steps, warmup, graph=True):
    """The unmodified reference on ONE B200 through stock PyTorch: fp32 parameters and activations, cuDNN/ATen kernels,
    cudnn.benchmark=True and PyTorch's default TF32 policy for cuDNN convolutions -- exactly what
    experiments/search_arc.py:61-76 sets up.  Times the eagerly launched search step and the same step captured into a
    CUDA graph (the same capture our arm uses), CUDA events around the timed steps."""
    import torch
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    torch.backends.cudnn.enabled = True
    torch.backends.cudnn.benchmark = True
    B, size = args.batch, args.size
    step, data, model, (w_opt, a_opt) = reference_step_factory(size, B, device=dev)
    out = {'batch': B, 'dtype': 'f32', 'cudnn_benchmark': True,
           'cudnn_allow_tf32': bool(torch.backends.cudnn.allow_tf32), 'source': ref_env().kind()}

    def timed(fn, n):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    eager = lambda: step(*data)  # noqa: E731
    for _ in range(warmup):
        eager()
    ms = timed(eager, steps)
    out['eager'] = {'ms_per_step': ms, 'images_per_sec': B / (ms * 1e-3), 'steps': steps, 'warmup': warmup}
    out['peak_mem_gb'] = torch.cuda.max_memory_allocated() / 2 ** 30
    if graph:
        try:
            del step, model, w_opt, a_opt, eager
            torch.cuda.empty_cache()
            # fresh model / optimizers: Adam must be capturable from its first step (its `step` counters live on the
            # device then); nothing else differs from the eager run
            # and a device-resident dice_ce: the reference's own loss builds its one-hot target on the host every step
            # (utils/loss/loss.py:203-206), a host->device copy that cannot be captured; senas_b200.loss is the same
            # arithmetic in plain torch ops.  The model, Architecture and optimizers are the reference's.
            _step, data, model, (w_opt, a_opt) = reference_step_factory(size, B, device=dev)
            for g in a_opt.param_groups:
                g['capturable'] = True
            from senas_b200.loss import SegmentationLosses as _DevLoss
            crit = _DevLoss('dice_ce')

            def eager():
                xt, yt, xv, yv = data
                a_opt.zero_grad()
                crit(model(xv), yv).backward()
                a_opt.step()
                w_opt.zero_grad()
                loss = crit(model(xt), yt)
                loss.backward()
                torch.nn.utils.clip_grad_norm_(model.parameters(), 5.0)
                w_opt.step()
                return loss
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    eager()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            cg = torch.cuda.CUDAGraph()
            with torch.cuda.graph(cg):
                eager()
            for _ in range(max(3, warmup)):
                cg.replay()
            ms = timed(cg.replay, steps)
            out['cuda_graph'] = {'ms_per_step': ms, 'images_per_sec': B / (ms * 1e-3), 'steps': steps}
        except Exception as e:
            out['cuda_graph'] = {'failed': f'{type(e).__name__}: {str(e)[:200]}'}
    return out


def run_torch_eager(args, rank):
    if rank != 0:
        return
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    r = torch_eager_measure(args, args.steps, max(3, args.warmup))
    best = r['eager']
    if 'images_per_sec' in r.get('cuda_graph', {}) and r['cuda_graph']['images_per_sec'] > best['images_per_sec']:
        best = r['cuda_graph']
    line = {'impl': 'torch_eager', 'metric': 'search_step_images_per_sec', 'value': best['images_per_sec'],
            'unit': 'images/s', 'n_gpus': 1, 'steps': args.steps, 'warmup': max(3, args.warmup),
            'ms_per_step': best['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': WORKLOAD.format(B=args.batch, S=args.size, G=args.batch), 'parallelism': 'dp1',
                       'launch': 'unmodified reference, stock PyTorch (cuDNN/ATen) on one B200; value = the faster of '
                                 'eager and CUDA-graph replay'},
            'detail': r}
    os.dup2(real_stdout, 1)
    print(json.dumps(line), flush=True)


def subprocess_json(argv, timeout):
    """Run bench.py with other arguments in a fresh process (own CUDA context / allocator / cuDNN settings) and return
    the JSON line it prints, or {'failed': why}."""
    try:
        res = subprocess.run([sys.executable, os.path.abspath(__file__)] + argv, stdout=subprocess.PIPE,
                             stderr=subprocess.PIPE, text=True, timeout=timeout)
        for line in reversed(res.stdout.strip().splitlines()):
            if line.startswith('{'):
                return json.loads(line)
        return {'failed': f'rc={res.returncode}: {res.stderr.strip()[-300:]}'}
    except Exception as e:
        return {'failed': f'{type(e).__name__}: {str(e)[:200]}'}


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
    buckets, comm = None, None
    if world > 1:
        broadcast_parameters(model)
        if not args.no_comm:  # NCCL communicator owned by libsenas_b200: its all-reduces are captured into the step's graph
            try:
                from senas_b200.comm import Comm
                comm = Comm(group=group, device=dev)
            except Exception as e:
                print(f'senas_b200.Comm unavailable ({type(e).__name__}: {e}); torch.distributed all-reduces between graphs',
                      file=sys.stderr)

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
                                                   warmup=3, group=group, comm=comm,
                                                   capture_error_mode='thread_local' if world > 1 else 'global',
                                                   concurrent_cells=not args.serial_cells, defer_wgrad=args.defer_wgrad,
                                                   overlap=not args.no_overlap, fused_optim=not args.no_fused_optim,
                                                   arch_grads_only=args.arch_grads_only)
            launches_per_step = (lib.senas_launch_count() - n_before) // 4   # 3 warm-up steps + 1 capture pass
            graph_note = 'cuda-graph (whole search step captured once, replayed per step)'
            if world > 1:
                graph_note = ('cuda-graph (one graph per step; NCCL all-reduces of the libsenas_b200 communicator captured '
                              'inside: per-cell weight-gradient buckets on a side stream overlapped with backward, one '
                              'bucket for the rest, one for the arch gradients)' if comm is not None else
                              'cuda-graph (three graphs per step, two torch.distributed all-reduces between them)')
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

    mem_main = torch.cuda.max_memory_allocated() / 2 ** 30
    # per-kernel-family device time (CUDA events on the launch stream), one more step
    lib.senas_profile(1)
    barrier()
    eager_step(*devb[0], *devb[1])   # eagerly launched so that every launch is bracketed by its own events
    barrier()
    lib.senas_profile(0)
    prof = _lib.profile_dump(lib)
    total_ms = sum(v['ms'] for v in prof.values()) or 1.0

    # BASELINE config 3: global batch 128 at N = 2 / 4 / 8 (64 / 32 / 16 images per GPU).  The weak-scaling line above
    # keeps 16 images per GPU at every N (N = 8 IS config 3); at N = 2, 4 the config-3 point is measured here as well, on
    # a second captured step of the larger per-GPU batch.
    config3 = None
    if world == 8 and B * world == 128:
        config3 = {'global_batch': 128, 'per_gpu_batch': B, 'same_as_value': True}
    elif world in (2, 4) and graphed is not None and not args.no_config3:
        try:
            graphed.release()
            del search_step
            torch.cuda.empty_cache()
            B3 = 128 // world
            host3 = [synth(B3, size, 4321 + 17 * rank + i, False) for i in range(2)]
            dev3 = [(x.to(dev), y.to(dev)) for x, y in host3]
            step3 = senas_b200.GraphedSearchStep(model, crit, w_opt, a_opt, (*dev3[0], *dev3[1]), grad_clip=5.0, warmup=2,
                                                 group=group, comm=comm, capture_error_mode='thread_local',
                                                 concurrent_cells=not args.serial_cells, overlap=not args.no_overlap,
                                                 fused_optim=not args.no_fused_optim, arch_grads_only=args.arch_grads_only)
            for _ in range(2):
                step3(*dev3[0], *dev3[1])
            n3 = min(args.steps, 10)
            ms3 = timed(lambda i: step3(*dev3[0], *dev3[1]), n3) / n3
            config3 = {'global_batch': 128, 'per_gpu_batch': B3, 'ms_per_step': ms3, 'value': 128 / (ms3 * 1e-3),
                       'unit': 'images/s', 'steps': n3, 'peak_mem_gb': torch.cuda.max_memory_allocated() / 2 ** 30}
            step3.release()
            del step3
        except Exception as e:
            config3 = {'failed': f'{type(e).__name__}: {str(e)[:200]}'}
    def teardown():
        """Every rank: graphs first, then the communicator, then the process group; a watchdog ends the process if a
        collective teardown blocks (the JSON line is already out by then)."""
        if world <= 1:
            return
        import threading
        threading.Timer(30.0, lambda: os._exit(0)).start()
        try:
            if graphed is not None:
                graphed.release()
            torch.cuda.synchronize()
            if comm is not None:
                comm.destroy()
            dist.barrier()
            dist.destroy_process_group()
        finally:
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)

    if rank != 0:
        teardown()
        return
    pk = peaks()
    gB = B * world
    ms_step, ms_step_e2e = ms / args.steps, ms_e2e / args.steps
    value, value_e2e = gB / (ms_step * 1e-3), gB / (ms_step_e2e * 1e-3)
    # roofline.  `frac` is the SURVEY 8(d) quantity for the path: algorithmic FLOPs of the reference graph per image
    # (176 GFLOP, no credit for recompute / padding) x images/s per GPU / the measured sustained bf16 peak; the HBM
    # counterpart (390.4 MB/img MixedOp-boundary bytes) sits beside it.  `dominant_family` is the kernel family with the
    # largest device time of the step (reduction folds included), with its own achieved rate; `traffic` = DRAM bytes of
    # one step from the committed ncu pass (profiles/r2_traffic.json), null when that file is absent.
    per_gpu = value / world
    ach_tf = STEP_GFLOP_PER_IMG * 1e9 * per_gpu / 1e12
    ach_gb = STEP_MB_PER_IMG * 1e6 * per_gpu / 1e9
    name, t = max(prof.items(), key=lambda kv: kv[1]['ms'])
    dom = {'kernel': name, 'ms_per_step': t['ms'], 'launches': t['launches'],
           'avg_launch_ms': t['ms'] / max(1, t['launches']), 'share_of_senas_kernels': t['ms'] / total_ms,
           'tflops': t['flops'] / (t['ms'] * 1e-3) / 1e12, 'gbs': t['bytes'] / (t['ms'] * 1e-3) / 1e9}
    dom['frac_tensor'], dom['frac_hbm'] = dom['tflops'] / pk['tf'], dom['gbs'] / pk['hbm']
    traffic = None
    try:
        with open(os.path.join(ROOT, 'profiles', 'r2_traffic.json')) as f:
            traffic = json.load(f)
    except Exception:
        pass
    roof = {'kernel': 'search step, all kernels (algorithmic 176 GFLOP/img, SURVEY 8d)', 'bound': 'tensor',
            'achieved': ach_tf, 'peak': pk['tf'], 'unit': 'TFLOP/s', 'frac': ach_tf / pk['tf'],
            'traffic': traffic.get('dram_bytes_per_step') if traffic else None,
            'traffic_over_algorithmic': (traffic['dram_bytes_per_step'] / (STEP_MB_PER_IMG * 1e6 * B)
                                         if traffic and traffic.get('dram_bytes_per_step') else None),
            'traffic_source': traffic.get('source') if traffic else None,
            'peak_source': pk['src'] + ' (bf16_tflops_sustained: the kernels are timed inside a long step)',
            'hbm': {'achieved': ach_gb, 'peak': pk['hbm'], 'unit': 'GB/s', 'frac': ach_gb / pk['hbm'],
                    'algorithmic_mb_per_img': STEP_MB_PER_IMG},
            'dominant_family': dom}
    families = {k: {'ms': round(v['ms'], 3), 'launches': v['launches'],
                    'tflops': round(v['flops'] / (v['ms'] * 1e-3) / 1e12, 3) if v['ms'] > 0 else 0,
                    'gbs': round(v['bytes'] / (v['ms'] * 1e-3) / 1e9, 1) if v['ms'] > 0 else 0}
                for k, v in sorted(prof.items(), key=lambda kv: -kv[1]['ms'])}
    out = {
        'metric': 'search_step_images_per_sec', 'value': value, 'unit': 'images/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'bf16' if args.conv_mode == 'bf16' else 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD.format(B=B, S=size, G=gB),
                   'parallelism': f'dp{world}', 'launch': graph_note, 'conv_mode': args.conv_mode + (' (tcgen05 bf16 operands, fp32 accumulate)' if args.conv_mode == 'bf16' else ' (exact FMA)'), 'l2': 'no flush: each step streams several GB of activations (>> 126 MB L2)',
                   'arch_step': ('alpha/beta/gamma gradients only (opt-in --arch-grads-only: the weight gradients the reference computes and discards in the architecture step are not computed)' if args.arch_grads_only else 'full backward, as the reference (weight gradients computed and discarded)'),
                   'between_cells': ('fused: ' + ', '.join(n for n, on in (('flat-arena clip+SGD/Adam (f4)', not args.no_fused_optim and graphed is not None), ('gamma-mix concat (f3)', os.environ.get('SENAS_NO_MIX', '0') != '1'), ('Shrink/Rectify ConvBn on tcgen05 (f1)', os.environ.get('SENAS_NO_CONVBN', '0') != '1' and args.conv_mode == 'bf16')) if on))},
        'e2e': {'value': value_e2e, 'unit': 'images/s', 'ms_per_step': ms_step_e2e,
                'h2d_bytes_per_step': 2 * B * size * size * (4 + 8), 'd2h_bytes_per_step': 4},
        'gpu_launches': int(launches), 'roofline': roof, 'kernel_families': families, 'clocks': clk,
    }
    if config3 is not None:
        out['config3'] = config3
    out['peak_mem_gb'] = mem_main
    if world == 1 and not args.no_cpu:
        # the unmodified reference on the host cores: ONE search step at the full batch after a tiny warm-up
        cpu_reference_sample(64, 1, 1, 0)                      # thread pool / allocator warm-up on a tiny input
        dt, cores = cpu_reference_sample(size, B, 1, 0)
        out['cpu_baseline'] = {'value': B / dt, 'unit': 'images/s', 'cores': cores, 'kind': 'reference',
                               'sample': f'1 search step of the full supernet at batch {B}, 1x{size}x{size}, fp32, '
                                         f'{cores} host threads, unmodified reference from {ref_env().kind()}'}
    if world == 1 and not args.no_ref_gpu:
        # "the kernel to beat": the unmodified reference through stock PyTorch on this same B200 (own process)
        r = subprocess_json(['--impl', 'torch_eager', '--steps', str(min(args.steps, 10)), '--warmup', '3',
                             '--batch', str(B), '--size', str(size)], 900)
        out['reference_b200'] = r.get('detail', r)
    if world == 1 and args.conv_mode == 'bf16' and not args.no_fp32_line:
        r = subprocess_json(['--conv-mode', 'fp32', '--steps', str(min(args.steps, 10)), '--warmup', '3', '--no-cpu',
                             '--no-ref-gpu', '--batch', str(B), '--size', str(size)], 900)
        out['fp32_mode'] = ({k: r[k] for k in ('value', 'ms_per_step', 'dtype') if k in r} if 'value' in r else r)
        if 'value' in r:
            out['fp32_mode']['e2e'] = r['e2e']['value']
            out['fp32_mode']['roofline_frac'] = r['roofline']['frac']
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    print(json.dumps(out), flush=True)
    os.dup2(2, 1)
    teardown()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference', 'torch_eager'])
    ap.add_argument('--batch', type=int, default=16, help='images per GPU')
    ap.add_argument('--size', type=int, default=256)
    ap.add_argument('--ref-batch', type=int, default=0,
                    help='--impl reference: images per step of the CPU sample (0 = the full batch, shrunk only if K + W steps '
                         'would exceed --ref-budget-s)')
    ap.add_argument('--ref-budget-s', type=float, default=540.0)
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--no-ref-gpu', action='store_true', help='skip the reference-on-B200 (stock PyTorch) leg')
    ap.add_argument('--no-fp32-line', action='store_true', help='skip the fp32-mode measurement beside the bf16 one')
    ap.add_argument('--no-graph', action='store_true', help='launch the step eagerly instead of replaying a CUDA graph')
    ap.add_argument('--no-comm', action='store_true', help='data parallel: keep the all-reduces outside the graphs (torch.distributed)')
    ap.add_argument('--no-overlap', action='store_true', help='data parallel: one post-backward weight-gradient bucket instead of per-cell buckets overlapped with backward')
    ap.add_argument('--no-fused-optim', action='store_true', help='clip_grad_norm_ / SGD / Adam of PyTorch inside the graph instead of the flat-buffer kernels of libsenas_b200 (row f4)')
    ap.add_argument('--arch-grads-only', action='store_true', help='opt-in: the architecture step computes alpha / beta / gamma gradients only (the reference also computes, then discards, every weight gradient there)')
    ap.add_argument('--no-config3', action='store_true', help='N = 2, 4: skip the BASELINE config-3 leg (global batch 128)')
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
    if args.impl == 'torch_eager':
        run_torch_eager(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3
    run_ours(args, rank, world, local_rank)


if __name__ == '__main__':
    main()
