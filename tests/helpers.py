"""Shared test plumbing: golden fixtures -> senas_b200 modules / oracle stores, comparison helpers."""
import glob
import os

import numpy as np
import torch

import senas_b200
from senas_b200.ops import OpType

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
OP_BY_ID = {1: OpType.UP, 2: OpType.DOWN, 3: OpType.NORM}
OP_NAME = {1: 'UP', 2: 'DOWN', 3: 'NORM'}


def golden(name):
    z = np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)
    return {k: z[k] for k in z.files}


def golden_names(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, prefix + '*.npz')))


def sub(g, prefix):
    return {k[len(prefix):]: torch.from_numpy(v.copy()) for k, v in g.items() if k.startswith(prefix)}


def max_err(a, b):
    """max |a-b| / max|b|: the 'fp32 within 1e-4 relative' gate of BASELINE.json, taken per tensor."""
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.numel() == 0:
        return 0.0
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-6)).item()


def mixed_module(g):
    c_in, B, H, W, training, op_id = [int(v) for v in g['meta']]
    m = senas_b200.MixedOp(c_in, 8, OP_BY_ID[op_id])
    m.load_state_dict(sub(g, 'state.'))
    m.train(bool(training))
    return m


def cell_module(g, cell_type):
    c = senas_b200.Cell(3, 1, 32, 32, 32, cell_type)
    c.load_state_dict(sub(g, 'state.'))
    c.train()
    return c


def run_graph_raw(runner, ins, alpha, beta, gout, training=True):
    """Drive a GraphRunner directly (no autograd): out, input/alpha/beta grads, {param: grad}."""
    nhwc = lambda t: t.contiguous(memory_format=torch.channels_last)
    ins = [nhwc(t.detach().float()) for t in ins]
    alpha = alpha.detach().float().contiguous()
    beta = beta.detach().float().contiguous() if beta is not None else None
    out, saved = runner.forward(ins, alpha, beta, training)
    res = {'out': out.clone()}
    if gout is not None:
        g_ins, g_alpha, g_beta, g_params = runner.backward(ins, alpha, beta, out, nhwc(gout.float()), saved, training,
                                                           [True] * len(ins))
        grads = [g.view(s) for g, s in zip(torch.split(g_params, runner.sizes), runner.shapes)]
        res.update(g_ins=g_ins, g_alpha=g_alpha, g_beta=g_beta, g_params=grads)
    return res
