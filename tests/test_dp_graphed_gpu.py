"""The graphed data-parallel search step (what ``bench.py --gpus N`` times) on 2 GPUs: gradients and the resulting
update == the mean over ranks of the per-shard local gradients, computed eagerly with explicit all-reduces.
`-m gpu`; skipped on a box with fewer than 2 GPUs (run it with ``gpurun --gpus 2``)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, result, use_comm, overlap=True):
    import torch.distributed as dist
    import senas_b200
    from senas_b200.dp import broadcast_parameters
    from senas_b200.loss import SegmentationLosses
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dev = torch.device('cuda', rank)
    torch.cuda.set_device(dev)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    senas_b200.exact_fp32()
    senas_b200.set_conv_mode('fp32')
    B, H = 2, 64
    gen = torch.Generator().manual_seed(1000 + rank)  # a different shard per rank
    xt = torch.randn(B, 1, H, H, generator=gen).to(dev)
    yt = (torch.rand(B, H, H, generator=gen) > 0.8).long().to(dev)
    xv = torch.randn(B, 1, H, H, generator=gen).to(dev)
    yv = (torch.rand(B, H, H, generator=gen) > 0.8).long().to(dev)

    def new_model():
        torch.manual_seed(0)
        m = senas_b200.NAS(1, 32, 2, depth=5, meta_node_num=3, use_sharing=False, double_down_channel=False,
                           supervision=False).to(dev).train()
        broadcast_parameters(m)
        return (m, torch.optim.SGD(m.parameters(), lr=5e-3, momentum=0.9, weight_decay=3e-4),
                torch.optim.Adam(m.arch_parameters(), lr=1e-4, betas=(0.5, 0.999), weight_decay=1e-3))

    crit = SegmentationLosses('dice_ce')  # local dice per rank: the stated semantics of the graphed DP path
    # --- eager restatement: local gradients, explicit mean over ranks -----------------------------------------
    m, w_opt, a_opt = new_model()
    init = {k: v.detach().clone() for k, v in m.state_dict().items()}
    a_opt.zero_grad()
    crit(m(xv), yv).backward()
    for p in m.arch_parameters():
        dist.all_reduce(p.grad)
        p.grad.div_(world)
    a_opt.step()
    w_opt.zero_grad()
    loss = crit(m(xt), yt)
    loss.backward()
    params = [p for p in m.parameters() if p.requires_grad]
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
        dist.all_reduce(p.grad)
        p.grad.div_(world)
    want_grads = torch.cat([p.grad.reshape(-1) for p in params]).clone()
    torch.nn.utils.clip_grad_norm_(m.parameters(), 5)
    want_clipped = torch.cat([p.grad.reshape(-1) for p in params]).clone()
    w_opt.step()
    want = {k: v.detach().clone() for k, v in m.state_dict().items()}
    want_loss = loss.item()
    # --- the graphed DP step -----------------------------------------------------------------------------------
    m2, w2, a2 = new_model()
    comm = None
    if use_comm:  # NCCL communicator of libsenas_b200: both all-reduces captured inside ONE graph
        from senas_b200.comm import Comm
        comm = Comm(group=dist.group.WORLD, device=dev)
    step = senas_b200.GraphedSearchStep(m2, crit, w2, a2, (xt, yt, xv, yv), grad_clip=5.0, warmup=2,
                                        group=dist.group.WORLD, comm=comm, capture_error_mode='thread_local', overlap=overlap)
    assert step.overlap == (bool(use_comm) and overlap)
    assert len(step.graphs) == (1 if use_comm else 3), len(step.graphs)
    for k, v in m2.state_dict().items():
        assert torch.equal(v, init[k]), f'warm-up left a trace in {k}'
    got_loss = step(xt, yt, xv, yv).item()
    torch.cuda.synchronize()
    got_clipped = step.flat_grads()  # averaged over ranks by the all-reduces, then clipped in place by graph 3
    got = m2.state_dict()
    errs = {}
    scale = want_clipped.abs().max().item()
    errs['grads'] = (got_clipped - want_clipped).abs().max().item() / scale
    errs['loss'] = abs(got_loss - want_loss) / abs(want_loss)
    worst_upd = 0.0
    for k, v in want.items():
        if not v.is_floating_point():
            assert torch.equal(got[k], v), k
            continue
        upd_w, upd_g = (v - init[k]).double(), (got[k] - init[k]).double()
        s = upd_w.abs().max().item()
        if s > 1e-12 and not k.startswith(('alphas', 'betas', 'gamma')):
            # an update of a BatchNorm weight (1.0) by lr * grad ~ 1e-5 is ~150 ulps of the parameter: one ulp of rounding
            # in p - lr * buf is 1/147 of the update (seen on the box), so allow 2 ulps of the parameter on top
            ulp = 2 * 1.1920929e-07 * v.abs().max().item()
            worst_upd = max(worst_upd, max((upd_g - upd_w).abs().max().item() - ulp, 0.0) / s)
    errs['update'] = worst_upd
    errs['unclipped_norm'] = want_grads.norm().item()
    # every rank must hold the same weights after the step
    flat = torch.cat([p.detach().reshape(-1) for p in m2.parameters()])
    ref = flat.clone()
    dist.broadcast(ref, 0)
    errs['rank_divergence'] = (flat - ref).abs().max().item()
    result[rank] = errs
    step.release()  # graphs before the communicator they captured
    if comm is not None:
        comm.destroy()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs (gpurun --gpus 2)')
@pytest.mark.parametrize('use_comm,overlap', [(True, True), (True, False), (False, False)])
def test_graphed_dp_step_equals_mean_of_local_gradients(use_comm, overlap):
    """(True, True): per-cell buckets all-reduced on a side stream while backward runs, inside the one captured graph
    (the benched path); (True, False): one post-backward bucket inside the graph; (False, False): three graphs with
    torch.distributed all-reduces between them."""
    import torch.multiprocessing as mp
    world = 2
    mgr = mp.Manager()
    result = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), result, use_comm, overlap), nprocs=world, join=True)
    assert len(result) == world
    for rank, e in result.items():
        assert e['loss'] <= 2e-5, (rank, e)
        assert e['grads'] <= 1e-3, (rank, e)      # fp32 noise floor of the composed network (cuDNN atomics, reduction order)
        assert e['update'] <= 5e-3, (rank, e)
        assert e['rank_divergence'] == 0.0, (rank, e)
