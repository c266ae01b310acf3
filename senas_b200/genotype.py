"""Genotype derivation: API mirror of the reference's ``utils/genotype.py`` (``Genotype`` :5,
``GenoParser.parse`` :13-90).  Pure CPU index work on the softmaxed alpha/beta tables -- it is
not on the GPU path and must stay bit-exact, so the selection rules (strict ``>`` for the best
op, stable ``sorted`` on negated strengths, tuple ordering of the final top-2) are kept as is.
"""
from collections import namedtuple

import numpy as np

from .ops import DownOps, NormOps, UpOps

Genotype = namedtuple('Genotype', ['down', 'down_concat', 'up', 'up_concat', 'gamma'])


def _strongest(row, names):
    """Index of the largest non-'none' weight; first one wins ties (strict '>')."""
    best = None
    for k, w in enumerate(row):
        if names[k] == 'none':
            continue
        if best is None or w > row[best]:
            best = k
    return best


def _edge_strength(row, names):
    return max(w for k, w in enumerate(row) if names[k] != 'none')


class GenoParser:
    def __init__(self, meta_node_num=4):
        self._meta_node_num = meta_node_num

    def parse(self, weights1, weights2, cell_type):
        """``weights1``: [edges, ops] normal-op table; ``weights2``: up/down-op table."""
        down = cell_type == 'down'
        resize_inputs = 2 if down else 1          # how many of a node's first edges change resolution
        resize_names = DownOps if down else UpOps
        n_edges = weights1.shape[0]
        gene, first, fan_in = [], 0, 2
        for _ in range(self._meta_node_num):
            last, resize_end = first + fan_in, first + resize_inputs
            normal_rows = np.zeros(n_edges, dtype=bool)
            resize_rows = np.zeros(n_edges, dtype=bool)
            if down:
                normal_rows[resize_end:last] = True
                resize_rows[first:resize_end] = True
            else:                                 # up cell: |norm|up|norm|...|
                normal_rows[resize_end + 1:last] = True
                normal_rows[first:resize_end] = True
                resize_rows[resize_end] = True
            wn, wr = weights1[normal_rows].copy(), weights2[resize_rows].copy()
            picks_n, picks_r = [], []
            if len(wr) >= 1:
                order = sorted(range(resize_inputs), key=lambda e: -_edge_strength(wr[e], resize_names))
                for e in order[:min(len(wr), 2)]:
                    k = _strongest(wr[e], resize_names)
                    picks_r.append((wr[e][k], resize_names[k], e if down else e + 1))
            if len(wn) > 0:
                order = sorted(range(len(wn)), key=lambda e: -_edge_strength(wn[e], NormOps))
                for e in order[:min(len(wn), 2)]:
                    k = _strongest(wn[e], NormOps)
                    picks_n.append((wn[e][k], NormOps[k], 0 if (e == 0 and not down) else e + resize_inputs))
            if len(wn) > 0 and len(wr) > 0 and len(wn[0]) != len(wr[0]):
                a, b = len(wn[0]), len(wr[0])
                scale = min(a, b) / max(a, b)
                if a > b:
                    picks_r = [(w * scale, o, f) for (w, o, f) in picks_r]
                else:
                    picks_n = [(w * scale, o, f) for (w, o, f) in picks_n]
            picks_n += picks_r
            gene += [(o, f) for (_, o, f) in sorted(picks_n)[-2:]]
            first = last
            fan_in += 1
        return gene
