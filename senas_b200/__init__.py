"""senas_b200 -- B200-native (sm_100a) supernet-search hot path of SENAS.

Public surface mirrors the reference modules for this path:
``MixedOp``, ``Cell`` (search/cell.py), ``OPS``, ``OpType`` (utils/operations.py),
``SenasSearch``, ``NAS``, ``Architecture`` (search/senas_search.py), ``Genotype``/``GenoParser``
(utils/genotype.py).  ``patch_reference()`` reroutes the reference's own classes instead.
"""
from .ops import OPS, OpType, DownOps, UpOps, NormOps, weights_init  # noqa: F401
from .cell import MixedOp, Cell  # noqa: F401
from .genotype import Genotype, GenoParser  # noqa: F401
from .supernet import Head, SenasSearch, NAS, Architecture  # noqa: F401
from .build import build  # noqa: F401

__version__ = '0.1.0'
