"""Search network and controller around the fused cells: API mirror of the reference's
``search/senas_search.py`` (``Head`` :5-13, ``SenasSearch`` :16-112, ``NAS`` :115-279,
``Architecture`` :282-303).  Constructor signatures, attribute names, ``state_dict`` keys,
``parameters()`` order and fixed-seed initial values are identical to the reference; the
``Cell`` objects it builds run the sm_100a kernels (``senas_b200.cell``).

Everything outside the cells (stems, gamma-mixed skip concat, head conv) is stock PyTorch --
rows f1/f3 of SURVEY.md section 8 ("next").  Activations travel in ``channels_last`` (NHWC)
memory so the cells read and write them without layout conversion.
"""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from .cell import Cell
from .genotype import GenoParser, Genotype
from .ops import (BasicBlock, DownOps, NormOps, ReLUConv, StemConvBn, UpOps, weights_init)


class Head(nn.Module):
    def __init__(self, meta_node_num, double_down, c_in0, c_in1, nclass):
        super().__init__()
        self.up_cell = Cell(meta_node_num, double_down, c_in0, c_in1, c_in1, cell_type='up')
        self.segmentation_head = ReLUConv(c_in1, nclass, kernel_size=3)

    def forward(self, s0, ot, weights_up_norm, weights_up, betas_up):
        return self.segmentation_head(self.up_cell(s0, ot, weights_up_norm, weights_up, betas_up))


class SenasSearch(nn.Module):
    """UNet++-shaped supernet: ``depth-1`` down cells, a nested triangle of up cells, one head."""

    def __init__(self, in_channels, c, nclass, depth, meta_node_num=3, double_down_channel=False,
                 supervision=False):
        super().__init__()
        assert depth >= 2, 'depth must >= 2'
        self._depth, self._double_down_channel = depth, double_down_channel
        self._supervision, self._meta_node_num = supervision, meta_node_num
        dd = 2 if double_down_channel else 1
        c_in0 = c_in1 = c_curr = c

        self.blocks = nn.ModuleList()
        self.stem0 = StemConvBn(in_channels, c_in0, kernel_size=7)
        self.stem1 = nn.Sequential(nn.ReLU(inplace=False), nn.MaxPool2d(3, stride=2, padding=1),
                                   BasicBlock(c_in0, c_in1))
        widths = [[c_in1]]  # widths[i][j] = channels produced by blocks[i][j]
        down = nn.ModuleList([self.stem1])
        for _ in range(1, depth):
            c_curr = int(dd * c_curr)
            down.append(Cell(meta_node_num, dd, c_in0, c_in1, c_curr, cell_type='down'))
            widths[0].append(c_curr)
            c_in0, c_in1 = c_in1, c_curr
        self.blocks.append(down)
        for i in range(1, depth):
            row, row_w = nn.ModuleList(), []
            for j in range(depth - i):
                head_in0 = sum(widths[k][j] for k in range(i))
                row.append(Cell(meta_node_num, dd, head_in0, widths[i - 1][j + 1], widths[0][j], cell_type='up'))
                row_w.append(widths[0][j])
            widths.append(row_w)
            self.blocks.append(row)
        self.head_block = nn.ModuleList([Head(meta_node_num, dd, c, widths[-1][0], nclass)])

    # ``concurrent_cells``: the up cells (i, j) of one level i of the nested triangle depend only on level i - 1, so
    # they are independent of each other; when set (GraphedSearchStep does), every level runs its cells on separate
    # CUDA streams (cell j = 0, the largest, on the caller's stream), each with its own lane set / scratch slot, so that
    # the small latency-bound cells hide under the large one in the captured graph.  Same operations on the same data:
    # results do not depend on the setting.
    concurrent_cells = False
    down_side_streams = os.environ.get('SENAS_DOWN_STREAMS', '1') != '0'  # (only with concurrent_cells)
    fused_mix = os.environ.get('SENAS_NO_MIX', '0') != '1'  # gamma mix + concat through libsenas_b200 (row f3); False: torch.lerp / torch.cat
    gamma_rows_sum_to_one = False  # set by NAS (which softmaxes gamma); a direct caller may pass any gamma

    def _cell_stream(self, j, device):
        pool = self.__dict__.setdefault('_cell_streams', {})
        key = (str(device), j)
        if key not in pool:
            pool[key] = torch.cuda.Stream(device=device, priority=-1)
        return pool[key]

    def forward(self, x, alpha_dn_nm, alpha_up_nm, alpha_dn, alpha_up, beta_dn, beta_up, gamma):
        from . import fused
        if x.dim() == 4:
            x = x.contiguous(memory_format=torch.channels_last)
        depth = self._depth
        s0 = self.stem0(x)
        # out[i][j]: output of cell (i, j); row 0 is the down path.  The reference (senas_search.py:87-107) walks
        # j = depth-2 .. 0 outside and i inside while overwriting cell_out[i + j]; cell (i, j) reads cell_out[j .. i+j-1]
        # = out[0..i-1][j] and cell_out[i + j] = out[i-1][j+1], which the level-by-level walk below provides unchanged.
        side = self.concurrent_cells and x.is_cuda
        out = [[self.stem1(s0)]]
        for j in range(1, depth):
            prev = s0 if j == 1 else out[0][-2]
            if side and j >= 2 and self.down_side_streams:
                # The down path is a strict chain in forward, but in BACKWARD the small down cells (32^2 and below: ~300
                # latency-bound launches each) only wait for the small up cells of their own column; on the caller's stream
                # they would queue behind the 128^2 cell of level 1.  A stream of their own (joined at once in forward)
                # lets autograd run their backward beside it: 86.9 -> 84.8 ms per step.  (A fully dependency-driven
                # schedule -- every cell waiting only for the events of its own inputs, so that the forward down path
                # also overlaps with column 0 -- measured the same, 84.9 vs 84.8 ms, and was not kept.)
                cur = torch.cuda.current_stream(x.device)
                st = self._cell_stream(100 + j, x.device)
                st.wait_stream(cur)
                fused.set_slot(100 + j)
                try:
                    with torch.cuda.stream(st):
                        o = self.blocks[0][j](prev, out[0][-1], alpha_dn_nm, alpha_dn, beta_dn)
                finally:
                    fused.set_slot(0)
                cur.wait_stream(st)
                out[0].append(o)
                continue
            out[0].append(self.blocks[0][j](prev, out[0][-1], alpha_dn_nm, alpha_dn, beta_dn))
        for i in range(1, depth):
            cur = torch.cuda.current_stream(x.device) if side else None
            row, joins = [], []
            if side:  # fork every side stream BEFORE cell j = 0 is enqueued on the caller's stream
                for j in range(1, depth - i):
                    self._cell_stream(j, x.device).wait_stream(cur)
            for j in range(depth - i):
                gidx = [sum(range(k + j)) + j for k in range(1, i)]
                st = self._cell_stream(j, x.device) if (side and j > 0) else None
                if st is not None:
                    fused.set_slot(j)
                try:
                    with (torch.cuda.stream(st) if st is not None else _nullctx()):
                        parts = [out[0][j]]
                        if self.fused_mix and x.is_cuda and gidx:
                            in0 = mix_concat(gamma, gidx, [out[k][j] for k in range(i)])
                            row.append(self.blocks[i][j](in0, out[i - 1][j + 1], alpha_up_nm, alpha_up, beta_up))
                            continue
                        for k, g in enumerate(gidx):
                            if self.gamma_rows_sum_to_one:
                                # softmax pairs (NAS.forward, senas_search.py:260): a*g0 + b*g1 == lerp(a, b, g1), one
                                # kernel instead of three; the logits receive the same gradient g0*g1*(dL/dg1 - dL/dg0)
                                parts.append(torch.lerp(out[k][j], out[k + 1][j], gamma[g][1]))
                            else:
                                parts.append(out[k][j] * gamma[g][0] + out[k + 1][j] * gamma[g][1])
                        in0 = torch.cat(parts, dim=1)
                        row.append(self.blocks[i][j](in0, out[i - 1][j + 1], alpha_up_nm, alpha_up, beta_up))
                finally:
                    if st is not None:
                        fused.set_slot(0)
                        joins.append(st)
            for st in joins:
                cur.wait_stream(st)
            out.append(row)
        head = self.head_block[-1]
        if self._supervision:
            return [head(s0, out[i][0] if i else out[0][0], alpha_up_nm, alpha_up, beta_up) for i in range(depth)]
        return [head(s0, out[depth - 1][0], alpha_up_nm, alpha_up, beta_up)]


class _MixConcat(torch.autograd.Function):
    """``cat([T0, g[0]*T0 + g[1]*T1, ...], dim=1)`` (search/senas_search.py:96-107) written by
    libsenas_b200 straight into one NHWC concat buffer: one launch per 32-channel slot, no intermediate mix tensors, no
    ``torch.cat``; backward = one launch per skip tensor (its up to three slices of the concat gradient, weighted) plus the
    two dot products per gamma pair (fixed-order reduction)."""

    @staticmethod
    def forward(ctx, gamma, gidx, *ts):
        from . import _lib
        lib = _lib.get()
        ts = [t.contiguous(memory_format=torch.channels_last) for t in ts]
        B, C, H, W = ts[0].shape
        n = len(ts)
        gamma = gamma.contiguous()
        out = torch.empty((B, C * n, H, W), dtype=torch.float32, device=ts[0].device, memory_format=torch.channels_last)
        st = torch.cuda.current_stream(out.device).cuda_stream
        npix = B * H * W
        with torch.cuda.device(out.device):
            for s in range(n):
                if s == 0:
                    rc = lib.senas_mix_forward(ts[0].data_ptr(), C, None, 0, None, out.data_ptr(), C * n, 0, C, npix, st)
                else:
                    rc = lib.senas_mix_forward(ts[s - 1].data_ptr(), C, ts[s].data_ptr(), C, gamma.data_ptr() + 8 * gidx[s - 1],
                                               out.data_ptr(), C * n, C * s, C, npix, st)
                _lib.check(lib, rc)
        ctx.gidx, ctx.n, ctx.shape = list(gidx), n, (B, C, H, W)
        ctx.save_for_backward(gamma, *ts)
        return out

    @staticmethod
    def backward(ctx, g):
        from . import _lib
        lib = _lib.get()
        gamma, *ts = ctx.saved_tensors
        B, C, H, W = ctx.shape
        n, gidx = ctx.n, ctx.gidx
        g = g.contiguous(memory_format=torch.channels_last)
        dev = g.device
        st = torch.cuda.current_stream(dev).cuda_stream
        npix = B * H * W
        dgamma = torch.zeros_like(gamma)
        scratch = torch.empty(1184 * max(n - 1, 1), dtype=torch.float32, device=dev)
        gp = gamma.data_ptr()
        outs = []
        with torch.cuda.device(dev):
            for s in range(1, n):  # d gamma[g][0] = <g_s, T_{s-1}>, d gamma[g][1] = <g_s, T_s>
                _lib.check(lib, lib.senas_mix_backward(ts[s - 1].data_ptr(), C, ts[s].data_ptr(), C, gp + 8 * gidx[s - 1],
                                                       g.data_ptr(), C * n, C * s, C, npix, None, None,
                                                       dgamma.data_ptr() + 8 * gidx[s - 1], scratch.data_ptr() + 4 * 1184 * (s - 1), st))
            for k in range(n):
                if not ctx.needs_input_grad[2 + k]:
                    outs.append(None)
                    continue
                d = torch.empty((B, C, H, W), dtype=torch.float32, device=dev, memory_format=torch.channels_last)
                # slices T_k reached: slot 0 (k == 0, coefficient 1), `a` of slot k + 1, `b` of slot k
                w0, o0 = (None, 0) if k == 0 else (None, -1)
                w1, o1 = (gp + 8 * gidx[k], C * (k + 1)) if k + 1 < n else (None, -1)
                w2, o2 = (gp + 8 * gidx[k - 1] + 4, C * k) if k >= 1 else (None, -1)
                _lib.check(lib, lib.senas_mix_dx(g.data_ptr(), C * n, w0, o0, w1, o1, w2, o2, d.data_ptr(), C, npix, st))
                outs.append(d)
        return (dgamma, None, *outs)


def mix_concat(gamma, gidx, tensors):
    """``cat([T0, gamma[g0,0]*T0 + gamma[g0,1]*T1, gamma[g1,0]*T1 + gamma[g1,1]*T2, ...], 1)`` for CUDA fp32 tensors."""
    if len(tensors) == 1:
        return tensors[0]
    return _MixConcat.apply(gamma.float(), tuple(gidx), *tensors)


class _nullctx:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


class NAS(nn.Module):
    """Owner of the architecture parameters; same constructor as senas_search.py:117-120.

    ``multi_gpus`` is accepted for signature compatibility only: the reference's in-process
    replica path (:262-279) is broken as shipped; data parallelism here is one process per GPU
    (``senas_b200.dp``), so ``device_ids`` is always ``[0]``.
    """

    def __init__(self, input_c, c, num_classes, depth, meta_node_num=4, use_sharing=True,
                 double_down_channel=True, use_softmax_head=False, supervision=False, multi_gpus=False,
                 device='cuda'):
        super().__init__()
        self._use_sharing, self._meta_node_num, self._depth = use_sharing, meta_node_num, depth
        self.net = SenasSearch(input_c, c, num_classes, depth, meta_node_num, double_down_channel, supervision)
        self.net.apply(weights_init)
        self.net.gamma_rows_sum_to_one = True  # forward() below always passes softmax(gamma)
        self.device_ids = [0]
        self._init_alphas()

    def _init_alphas(self):
        k = sum(2 + i for i in range(self._meta_node_num))
        self.alphas_dn = nn.Parameter(1e-3 * torch.randn(k, len(DownOps)))
        self.alphas_up = nn.Parameter(1e-3 * torch.randn(k, len(UpOps)))
        self.alphas_dn_nm = nn.Parameter(1e-3 * torch.randn(k, len(NormOps)))
        self.alphas_up_nm = self.alphas_dn_nm if self._use_sharing else nn.Parameter(
            1e-3 * torch.randn(k, len(NormOps)))
        self.betas_dn = nn.Parameter(1e-3 * torch.randn(k))
        self.betas_up = nn.Parameter(1e-3 * torch.randn(k))
        self.gamma = nn.Parameter(1e-3 * torch.randn(sum(range(self._depth - 1)), 2))
        self._arch_parameters = [self.alphas_dn, self.alphas_up, self.alphas_dn_nm, self.alphas_up_nm,
                                 self.betas_dn, self.betas_up, self.gamma]

    def alphas_dict(self):
        return {'alphas_dn': self.alphas_dn, 'alphas_dn_nm': self.alphas_dn_nm, 'alphas_up': self.alphas_up,
                'alphas_up_nm': self.alphas_up_nm}

    def betas_dict(self):
        return {'betas_dn': self.betas_dn, 'betas_up': self.betas_up}

    def load_params(self, alphas_dict, betas_dict):
        """Key names as read by the reference (senas_search.py:170-176), including its quirk of
        dropping ``gamma`` from the arch-parameter list."""
        self.alphas_dn = alphas_dict['alphas_down']
        self.alphas_up = alphas_dict['alphas_up']
        self.alphas_dn_nm = alphas_dict['alphas_normal_down']
        self.alphas_up_nm = alphas_dict['alphas_normal_up']
        self.betas_dn = betas_dict['betas_down']
        self.betas_up = betas_dict['betas_up']
        self._arch_parameters = [self.alphas_dn, self.alphas_up, self.alphas_dn_nm, self.alphas_up_nm,
                                 self.betas_dn, self.betas_up]

    def arch_parameters(self):
        return self._arch_parameters

    def _beta_softmax(self, betas, detach=False):
        """Per-node softmax segments exactly as the reference takes them (senas_search.py:212-216,
        253-257): ``offset = len(list_of_segments)`` there, i.e. segment ``i`` is
        ``betas[i : 2*i + 2]`` (overlapping windows), *not* the cumulative edge offset."""
        segs = []
        for i in range(self._meta_node_num):
            off = len(segs)
            s = F.softmax(betas[off:off + 2 + i], dim=-1)
            segs.append(s.detach().cpu() if detach else s)
        return torch.cat(segs, dim=0)

    def genotype(self):
        """senas_search.py:203-244 -- CPU index work on the softmaxed tables."""
        sm = lambda t: F.softmax(t, dim=-1).detach().cpu()
        a_dn_nm, a_dn, a_up_nm, a_up = sm(self.alphas_dn_nm), sm(self.alphas_dn), sm(self.alphas_up_nm), sm(
            self.alphas_up)
        b_dn, b_up = self._beta_softmax(self.betas_dn, True), self._beta_softmax(self.betas_up, True)
        for j in range(a_dn.shape[0]):
            a_dn_nm[j, :] = a_dn_nm[j, :] * b_dn[j].item()
            a_dn[j, :] = a_dn[j, :] * b_dn[j].item()
            a_up_nm[j, :] = a_up_nm[j, :] * b_up[j].item()
            a_up[j, :] = a_up[j, :] * b_up[j].item()
        parser = GenoParser(self._meta_node_num)
        gene_down = parser.parse(a_dn_nm.numpy(), a_dn.numpy(), cell_type='down')
        gene_up = parser.parse(a_up_nm.numpy(), a_up.numpy(), cell_type='up')
        concat = range(2, self._meta_node_num + 2)
        gamma = sm(self.gamma)
        idx = torch.topk(gamma[:, 1], len(gamma) // 2, largest=False).indices
        g = gamma.argmax(1).tolist()
        g = [v if i not in idx else 0 for i, v in enumerate(g)]
        path = [g[sum(range(i)): sum(range(i)) + i] for i in range(1, self._depth - 1)]
        path = sum([(v[:v.index(1)] + [1] * len(v[v.index(1):])) if (1 in v) else v for v in path], [])
        return Genotype(down=gene_down, down_concat=concat, up=gene_up, up_concat=concat, gamma=path)

    def forward(self, x):
        sm = lambda t: F.softmax(t, dim=-1)
        return self.net(x, sm(self.alphas_dn_nm), sm(self.alphas_up_nm), sm(self.alphas_dn), sm(self.alphas_up),
                        self._beta_softmax(self.betas_dn), self._beta_softmax(self.betas_up), sm(self.gamma))


class Architecture(object):
    """First-order DARTS step on the architecture parameters (senas_search.py:282-303)."""

    def __init__(self, model, arch_optimizer, criterion):
        self.model, self.optimizer, self.criterion = model, arch_optimizer, criterion

    def step(self, input_valid, target_valid):
        self.optimizer.zero_grad()
        loss = self.criterion(self.model(input_valid), target_valid)
        loss.backward()
        self.optimizer.step()
