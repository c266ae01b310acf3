"""``MixedOp`` and ``Cell``: API mirror of the reference's ``search/cell.py`` (:5-43, :46-110) with
the arithmetic of the hot path moved into ``libsenas_b200.so``.

* same constructor signatures, attribute names (``_ops``, ``_op_type``, ``k``, ``c_out``,
  ``c_part``, ``preprocess0/1``, ``node_activation``, ``post_process``), ``state_dict`` keys and
  parameter order, so reference checkpoints load both ways and fixed-seed init is identical;
* ``MixedOp.forward(x, alpha_normal, alpha_up_dn)`` and
  ``Cell.forward(in0, in1, weights_norm, weights_chg, betas)`` keep their meaning;
* the six candidates, their BatchNorms, the softmax(alpha)-weighted sum, the per-edge beta, the
  node sum + ReLU and the concat run as sm_100a kernels (one edge graph per MixedOp / per Cell).
  ``preprocess0`` and ``post_process`` stay stock PyTorch for now.

There is no PyTorch fallback for the fused part: tensors must live on a B200.
"""
import torch
import torch.nn as nn

from .fused import GraphRunner
from .ops import OPS, OpType, RectifyBlock, ShrinkBlock, build_rectify


def _require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f'senas_b200.{what}: the fused MixedOp/Cell path runs only on a B200 (sm_100a) GPU; '
                           f'got a {t.device.type} tensor and there is no CPU fallback')


class MixedOp(nn.Module):
    def __init__(self, c_in, c_out, op_type):
        super().__init__()
        self._ops = nn.ModuleList()
        self._op_type = op_type
        self.k = 1  # PC-DARTS style partial-channel factor; 1 in the reference (cell.py:14) => no skip branch
        self.c_out = c_out
        self.c_part = int(c_out // self.k)
        self._c_in = c_in
        for pri in self._op_type.value['ops']:
            self._ops.append(OPS[pri](c_in, self.c_part, self._op_type, dp=0))
        self._runner = None

    def _edge(self, src, dst):
        return (list(self._ops), src, dst, self._op_type.value['id'], self._c_in)

    def _ensure_runner(self):
        if self._runner is None:
            self._runner = GraphRunner([self._edge(0, 0)], n_inputs=1, n_nodes=1, node_relu=False)
        return self._runner

    def forward(self, x, alpha_normal, alpha_up_dn):
        _require_cuda(x, 'MixedOp')
        self._ensure_runner()
        w = alpha_normal if self._op_type == OpType.NORM else alpha_up_dn
        return self._runner.apply([x], w.reshape(1, -1), None, self.training)

    def _apply(self, fn, *a, **k):  # parameter storage may move: drop the graph (rebuilt lazily)
        self._runner = None
        return super()._apply(fn, *a, **k)


class Cell(nn.Module):
    def __init__(self, meta_node_num, double_down, c_in0, c_in1, c_out, cell_type):
        super().__init__()
        self.k = 4  # "shrink": every MixedOp works on c_out / 4 channels (cell.py:53)
        # what libsenas_b200 has kernels for (every senas config: init_channels 32, meta_node_num 3, no channel doubling);
        # say so here instead of failing at the first forward
        if meta_node_num < 1 or meta_node_num > 3:
            raise NotImplementedError(f'senas_b200.Cell: meta_node_num = {meta_node_num} is not supported (1..3; the '
                                      'reference configs use 3): a node with 5 incoming MixedOps has 30 candidate terms')
        cp = int((c_out // double_down) // self.k) if cell_type == 'down' else int(c_out // self.k)
        if cp != 8 or c_in1 != 32:
            raise NotImplementedError(f'senas_b200.Cell: c_in1 = {c_in1}, c_out / 4 = {cp}: the kernels cover the 32-channel '
                                      'search supernet (MixedOps 32 -> 8 and 8 -> 8; double_down_channel=False)')
        self._meta_node_num = meta_node_num
        self._input_num = 2
        if cell_type == 'down':
            self.preprocess0 = build_rectify(c_in0, c_in1, cell_type)
            c_part = int((c_out // double_down) // self.k)
        else:
            self.preprocess0 = ShrinkBlock(c_in0, c_in1)
            c_part = int(c_out // self.k)
        self.preprocess1 = nn.ReLU(inplace=False)
        self.node_activation = nn.ReLU(inplace=True)
        self.post_process = RectifyBlock(c_part * meta_node_num, c_out, cell_type=cell_type)
        self._ops = nn.ModuleList()
        srcs, dsts, norm_rows = [], [], []
        for i in range(meta_node_num):
            for j in range(self._input_num + i):
                if j < self._input_num:
                    if cell_type == 'down':
                        op = MixedOp(c_in1, c_part, OpType.DOWN)
                    elif j > 0:
                        op = MixedOp(c_in1, c_part, OpType.UP)
                    else:
                        op = MixedOp(c_in1, c_part, OpType.NORM)
                else:
                    op = MixedOp(c_part, c_part, OpType.NORM)
                self._ops.append(op)
                srcs.append(j)
                dsts.append(i)
                norm_rows.append(op._op_type == OpType.NORM)
        self._srcs, self._dsts = srcs, dsts
        self._norm_rows = torch.tensor(norm_rows, dtype=torch.bool).view(-1, 1)
        self._runner = None

    def _apply(self, fn, *a, **k):
        self._runner = None
        return super()._apply(fn, *a, **k)

    def _ensure_runner(self):
        if self._runner is None:
            edges = [op._edge(s, d) for op, s, d in zip(self._ops, self._srcs, self._dsts)]
            self._runner = GraphRunner(edges, n_inputs=2, n_nodes=self._meta_node_num, node_relu=True)
        return self._runner

    def nodes(self, in0, in1, weights_norm, weights_chg, betas):
        """Node loop + concat (cell.py:95-110) on pre-processed inputs -> [B, 8*nodes, H, W]."""
        _require_cuda(in1, 'Cell')
        self._ensure_runner()
        if self._norm_rows.device != weights_norm.device:
            self._norm_rows = self._norm_rows.to(weights_norm.device)
        alpha = torch.where(self._norm_rows, weights_norm, weights_chg)  # row used by each edge (cell.py:33-36)
        return self._runner.apply([in0, in1], alpha, betas, self.training)

    def forward(self, in0, in1, weights_norm, weights_chg, betas):
        in0 = self.preprocess0(in0)
        in1 = self.preprocess1(in1)
        return self.post_process(self.nodes(in0, in1, weights_norm, weights_chg, betas))
