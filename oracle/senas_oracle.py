"""CPU oracle for the SENAS supernet-search hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import this file; the product package ``senas_b200`` never does.

What it is: a functional (state-dict driven) restatement, in plain ``torch.nn.functional`` fp32
on the CPU, of the reference's algorithm for the path named by BASELINE.json:

* candidate operations           -- /root/reference/utils/operations.py:8-21,57-130,155-203
* ``MixedOp.forward``            -- /root/reference/search/cell.py:32-36
* ``Cell.forward`` (nodes, cat)  -- /root/reference/search/cell.py:92-110
* ``SenasSearch.forward``        -- /root/reference/search/senas_search.py:76-112
* ``NAS.forward`` / ``genotype`` -- /root/reference/search/senas_search.py:203-260
* ``dice_ce`` loss               -- /root/reference/utils/loss/loss.py:45-70,124-159,173-228

The arithmetic itself is PyTorch's (the reference pins ``torch==1.8.1`` in requirements.txt:5 and
calls ``nn.Conv2d`` / ``ConvTranspose2d`` / ``BatchNorm2d`` / ``AvgPool2d`` / ``Upsample`` /
``Linear``); it is installed here, so those primitives are *called*, not restated.  Gradients are
autograd's over this forward, exactly as in the reference.

Pinning: the reference ships no tests or golden vectors for this path, so
the oracle is pinned against the reference itself: ``tests/golden/make_golden.py`` imports
``/root/reference`` (build container only), runs its ``MixedOp`` / ``Cell`` / ``NAS`` on seeded
inputs and stores inputs, state dicts, outputs and gradients under ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks this file against every one of them.
"""
import math
from collections import namedtuple

import numpy as np
import torch
import torch.nn.functional as F

DOWN_OPS = ['avg_pool', 'se_conv_3', 'dil_3_conv_5', 'dil_2_conv_5', 'dep_sep_conv_3', 'dep_sep_conv_5']
UP_OPS = ['up_sample', 'se_conv_3', 'dil_3_conv_5', 'dil_2_conv_5', 'dep_sep_conv_3', 'dep_sep_conv_5']
NORM_OPS = ['identity', 'none', 'dil_3_conv_5', 'dil_2_conv_5', 'dep_sep_conv_3', 'dep_sep_conv_5']
CANDIDATES = {'UP': UP_OPS, 'DOWN': DOWN_OPS, 'NORM': NORM_OPS}

BN_EPS, BN_MOMENTUM = 1e-5, 0.1


class Params:
    """A view on a flat ``{name: tensor}`` dict under a key prefix (state_dict naming)."""

    def __init__(self, store, prefix=''):
        self.store, self.prefix = store, prefix

    def sub(self, name):
        return Params(self.store, f'{self.prefix}{name}.')

    def __getitem__(self, name):
        return self.store[self.prefix + name]

    def has(self, name):
        return (self.prefix + name) in self.store


def batch_norm(x, p, training):
    """``nn.BatchNorm2d`` (operations.py:133-134): biased var to normalise, unbiased var and
    momentum 0.1 into the running buffers, ``num_batches_tracked += 1``."""
    if training:
        p['num_batches_tracked'].add_(1)
    return F.batch_norm(x, p['running_mean'], p['running_var'], p['weight'], p['bias'], training,
                        BN_MOMENTUM, BN_EPS)


def conv(x, w, k, dilation, op_type, groups=1):
    """``build_weight`` (operations.py:118-130) with ``build_ops`` geometry (:57-60)."""
    pad = (k // 2) * dilation
    if op_type == 'UP':
        return F.conv_transpose2d(x, w, None, stride=2, padding=pad, output_padding=1, groups=groups,
                                  dilation=dilation)
    return F.conv2d(x, w, None, stride=1 if op_type == 'NORM' else 2, padding=pad, dilation=dilation,
                    groups=groups)


def se_block(x, p):
    """``SEBlock.forward`` (operations.py:199-203)."""
    b, c = x.shape[:2]
    q = F.adaptive_avg_pool2d(x, 1).view(b, c)
    h = F.relu(F.linear(q, p['excitation.0.weight']))
    s = torch.sigmoid(F.linear(h, p['excitation.2.weight']))
    return x * s.view(b, c, 1, 1).expand_as(x)


def candidate(name, op_type, p, x, training):
    """One entry of ``OPS`` applied to ``x`` (operations.py:8-21)."""
    if name in ('none', 'identity', 'avg_pool', 'up_sample'):  # AdapterBlock, :167-183
        if name == 'none':
            t = x.mul(0.)
        elif name == 'identity':
            t = x
        elif name == 'avg_pool':
            t = F.avg_pool2d(x, 3, stride=1 if op_type == 'NORM' else 2, padding=1, count_include_pad=False)
        else:
            t = F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=False)
        if p.has('conv.weight'):
            t = F.conv2d(t, p['conv.weight'])
        return batch_norm(t, p.sub('norm'), training)
    if name in ('dil_3_conv_5', 'dil_2_conv_5'):  # ConvBn, :89-95
        d = 3 if name == 'dil_3_conv_5' else 2
        return batch_norm(conv(x, p['0.weight'], 5, d, op_type), p.sub('1'), training)
    if name == 'se_conv_3':  # ConvBnSe, :98-104
        t = batch_norm(conv(x, p['0.weight'], 3, 1, op_type), p.sub('1'), training)
        return se_block(t, p.sub('2'))
    if name in ('dep_sep_conv_3', 'dep_sep_conv_5'):  # DepSepConv, :107-115
        k = 3 if name.endswith('3') else 5
        t = conv(x, p['0.weight'], k, 1, op_type, groups=x.shape[1])
        t = F.relu(batch_norm(t, p.sub('1'), training))
        t = F.conv2d(t, p['3.weight'])
        return batch_norm(t, p.sub('4'), training)
    raise NotImplementedError(name)


def mixed_op(p, op_type, x, weights, training=True):
    """``MixedOp.forward`` (cell.py:32-36): python ``sum`` starts at int 0, left to right."""
    out = 0
    for k, name in enumerate(CANDIDATES[op_type]):
        out = out + weights[k] * candidate(name, op_type, p.sub(f'_ops.{k}'), x, training)
    return out


def edge_types(cell_type, n_nodes=3):
    """(source state, OpType) of every edge, in ``Cell._ops`` order (cell.py:76-90)."""
    out = []
    for i in range(n_nodes):
        for j in range(2 + i):
            if j < 2:
                t = 'DOWN' if cell_type == 'down' else ('UP' if j > 0 else 'NORM')
            else:
                t = 'NORM'
            out.append((j, t))
    return out


def cell_nodes(p, cell_type, in0, in1, w_norm, w_chg, betas, training=True, n_nodes=3):
    """Node loop + concat of ``Cell.forward`` (cell.py:95-110) on already pre-processed inputs."""
    states = [in0, in1]
    types = edge_types(cell_type, n_nodes)
    offset = 0
    for i in range(n_nodes):
        node = None
        for j, h in enumerate(states):
            e = offset + j
            t = types[e][1]
            w = w_norm[e] if t == 'NORM' else w_chg[e]
            feat = betas[e] * mixed_op(p.sub(f'_ops.{e}'), t, h, w, training)
            node = feat if j == 0 else node + feat
        offset += len(states)
        states.append(F.relu(node))
    return torch.cat(states[-n_nodes:], dim=1)


def shrink_block(p, x, training):  # operations.py:206-218
    return batch_norm(F.conv2d(F.relu(x), p['conv.weight'], padding=1), p.sub('norm'), training)


def rectify_down(p, x, c_in, c_ot, training):  # build_rectify 'down', operations.py:141-152
    x = F.relu(x)
    if c_in == c_ot:
        x = F.avg_pool2d(x, 3, stride=2, padding=1, count_include_pad=False)
    else:
        x = F.conv2d(x, p['1.weight'], stride=2)
    return batch_norm(x, p.sub('2'), training)


def cell(p, cell_type, in0, in1, w_norm, w_chg, betas, training=True, n_nodes=3):
    """Whole ``Cell.forward`` (cell.py:92-110)."""
    if cell_type == 'down':
        in0 = rectify_down(p.sub('preprocess0'), in0, in0.shape[1], in1.shape[1], training)
    else:
        in0 = shrink_block(p.sub('preprocess0'), in0, training)
    in1 = F.relu(in1)
    cat = cell_nodes(p, cell_type, in0, in1, w_norm, w_chg, betas, training, n_nodes)
    pp = p.sub('post_process')
    return batch_norm(F.conv2d(cat, pp['conv.weight'], padding=1), pp.sub('norm'), training)


def basic_block(p, x, training):  # operations.py:235-268
    out = F.relu(batch_norm(F.conv2d(x, p['conv1.weight'], padding=1), p.sub('bn1'), training))
    out = batch_norm(F.conv2d(out, p['conv2.weight'], padding=1), p.sub('bn2'), training)
    return out + x


def supernet(p, x, a_dn_nm, a_up_nm, a_dn, a_up, b_dn, b_up, gamma, depth=5, n_nodes=3, training=True):
    """``SenasSearch.forward`` without deep supervision (senas_search.py:76-112)."""
    s0 = batch_norm(F.conv2d(x, p['stem0.0.weight'], padding=3), p.sub('stem0.1'), training)
    t = F.max_pool2d(F.relu(s0), 3, stride=2, padding=1)
    cell_out = [basic_block(p.sub('blocks.0.0.2'), t, training)]  # == stem1.2 (same modules, first registered name)
    for j in range(1, depth):
        a = s0 if j == 1 else cell_out[-2]
        cell_out.append(cell(p.sub(f'blocks.0.{j}'), 'down', a, cell_out[-1], a_dn_nm, a_dn, b_dn, training,
                             n_nodes))
    for j in reversed(range(depth - 1)):
        for i in range(1, depth - j):
            ides = list(range(j, i + j))
            gidx = [sum(range(k + j)) + j for k in range(1, i)]
            parts = [cell_out[ides[0]]] + [cell_out[ides[k]] * gamma[g][0] + cell_out[ides[k + 1]] * gamma[g][1]
                                           for k, g in enumerate(gidx)]
            cell_out[i + j] = cell(p.sub(f'blocks.{i}.{j}'), 'up', torch.cat(parts, dim=1), cell_out[i + j],
                                   a_up_nm, a_up, b_up, training, n_nodes)
    hp = p.sub('head_block.0')
    h = cell(hp.sub('up_cell'), 'up', s0, cell_out[-1], a_up_nm, a_up, b_up, training, n_nodes)
    return [F.conv2d(F.relu(h), hp['segmentation_head.1.weight'], padding=1)]


def arch_softmax(store, n_nodes=3):
    """Softmaxes of ``NAS.forward`` (senas_search.py:248-260).  The beta segments follow the
    reference literally: ``offset = len(betas_dn)`` is the number of segments appended so far
    (:254), so node i uses ``betas[i : 2*i + 2]``."""
    sm = lambda t: F.softmax(t, dim=-1)
    bd, bu = [], []
    for i in range(n_nodes):
        off = len(bd)
        bd.append(sm(store['betas_dn'][off:off + 2 + i]))
        bu.append(sm(store['betas_up'][off:off + 2 + i]))
    return dict(a_dn_nm=sm(store['alphas_dn_nm']), a_up_nm=sm(store['alphas_up_nm']), a_dn=sm(store['alphas_dn']),
                a_up=sm(store['alphas_up']), b_dn=torch.cat(bd), b_up=torch.cat(bu), gamma=sm(store['gamma']))


def nas_forward(store, x, depth=5, n_nodes=3, training=True):
    """``NAS.forward`` on a ``NAS.state_dict()``-shaped store."""
    a = arch_softmax(store, n_nodes)
    return supernet(Params(store, 'net.'), x, depth=depth, n_nodes=n_nodes, training=training, **a)


# ------------------------------------------------------------------------------------------
# loss (utils/loss/loss.py) -- dice_ce = CE + soft dice over batch+spatial axes, foreground only
# ------------------------------------------------------------------------------------------
def dice_ce_loss(logits, target, smooth=1e-5):
    prob = F.softmax(logits, 1)
    onehot = torch.zeros_like(prob).scatter_(1, target.long().unsqueeze(1), 1)
    axes = (0, 2, 3)
    tp = (prob * onehot).sum(axes)
    fp = (prob * (1 - onehot)).sum(axes)
    fn = ((1 - prob) * onehot).sum(axes)
    dc = (2 * tp + smooth) / (2 * tp + fp + fn + smooth + 1e-8)
    return F.cross_entropy(logits, target.long()) + (1 - dc[1:].mean())


# ------------------------------------------------------------------------------------------
# genotype derivation (senas_search.py:203-244, utils/genotype.py:13-90) -- integer/index work
# ------------------------------------------------------------------------------------------
Genotype = namedtuple('Genotype', ['down', 'down_concat', 'up', 'up_concat', 'gamma'])


def _best_op(row, names):
    best = None
    for k in range(len(row)):
        if names[k] != 'none' and (best is None or row[k] > row[best]):
            best = k
    return best


def parse_cell(w_norm, w_chg, cell_type, n_nodes=3):
    """``GenoParser.parse`` (utils/genotype.py:13-90)."""
    gene, start, n = [], 0, 2
    n_chg = 2 if cell_type == 'down' else 1
    chg_names = UP_OPS if cell_type == 'up' else DOWN_OPS
    nc = w_norm.shape[0]
    for _ in range(n_nodes):
        end, chg_end = start + n, start + n_chg
        m_norm, m_chg = np.zeros(nc, dtype=bool), np.zeros(nc, dtype=bool)
        if cell_type == 'down':
            m_norm[chg_end:end] = True
            m_chg[start:chg_end] = True
        else:
            m_norm[chg_end + 1:end] = True
            m_norm[start:chg_end] = True
            m_chg[chg_end] = True
        W1, W2 = w_norm[m_norm].copy(), w_chg[m_chg].copy()
        item1, item2 = [], []
        if len(W2) >= 1:
            strength = lambda x: -max(W2[x][k] for k in range(len(W2[x])) if chg_names[k] != 'none')
            for j in sorted(range(n_chg), key=strength)[:min(len(W2), 2)]:
                kb = _best_op(W2[j], chg_names)
                item2.append((W2[j][kb], chg_names[kb], j if cell_type == 'down' else j + 1))
        if len(W1) > 0:
            strength = lambda x: -max(W1[x][k] for k in range(len(W1[x])) if NORM_OPS[k] != 'none')
            for j in sorted(range(len(W1)), key=strength)[:min(len(W1), 2)]:
                kb = _best_op(W1[j], NORM_OPS)
                item1.append((W1[j][kb], NORM_OPS[kb], 0 if j == 0 and cell_type == 'up' else j + n_chg))
        if len(W1) > 0 and len(W2) > 0 and len(W1[0]) != len(W2[0]):
            scale = min(len(W1[0]), len(W2[0])) / max(len(W1[0]), len(W2[0]))
            if len(W1[0]) > len(W2[0]):
                item2 = [(w * scale, o, f) for (w, o, f) in item2]
            else:
                item1 = [(w * scale, o, f) for (w, o, f) in item1]
        item1 += item2
        gene += [(o, f) for (_, o, f) in sorted(item1)[-2:]]
        start = end
        n += 1
    return gene


def genotype(store, depth=5, n_nodes=3):
    """``NAS.genotype`` (senas_search.py:203-244)."""
    a = {k: v.detach().cpu().clone() for k, v in arch_softmax(store, n_nodes).items()}
    k = a['a_dn'].shape[0]
    for j in range(k):
        a['a_dn_nm'][j, :] = a['a_dn_nm'][j, :] * a['b_dn'][j].item()
        a['a_dn'][j, :] = a['a_dn'][j, :] * a['b_dn'][j].item()
        a['a_up_nm'][j, :] = a['a_up_nm'][j, :] * a['b_up'][j].item()
        a['a_up'][j, :] = a['a_up'][j, :] * a['b_up'][j].item()
    down = parse_cell(a['a_dn_nm'].numpy(), a['a_dn'].numpy(), 'down', n_nodes)
    up = parse_cell(a['a_up_nm'].numpy(), a['a_up'].numpy(), 'up', n_nodes)
    concat = range(2, n_nodes + 2)
    g = a['gamma']
    idx = torch.topk(g[:, 1], len(g) // 2, largest=False).indices
    gl = g.argmax(1).tolist()
    gl = [v if i not in idx else 0 for i, v in enumerate(gl)]
    path = [gl[sum(range(i)): sum(range(i)) + i] for i in range(1, depth - 1)]
    path = sum([(v[:v.index(1)] + [1] * len(v[v.index(1):])) if (1 in v) else v for v in path], [])
    return Genotype(down=down, down_concat=concat, up=up, up_concat=concat, gamma=path)


def clone_store(sd, requires_grad=True):
    """Detached fp32 CPU copy of a state dict; float tensors become leaves that require grad."""
    out = {}
    for k, v in sd.items():
        t = v.detach().cpu().clone()
        if requires_grad and t.is_floating_point() and not k.endswith(('running_mean', 'running_var')):
            t.requires_grad_(True)
        out[k] = t
    return out
