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
from .patch import patch_reference  # noqa: F401
from .fused import set_conv_mode, get_conv_mode  # noqa: F401
from .graphs import GraphedSearchStep  # noqa: F401
from .optim import FusedSearchOptim  # noqa: F401

__version__ = '0.1.0'


def exact_fp32():
    """fp32 mode gate (1e-4 relative vs the reference): the fused MixedOp/Cell kernels are plain fp32
    FMA already; this turns off cuDNN/cuBLAS TF32 for the stock-PyTorch blocks around them (stems,
    ShrinkBlock / RectifyBlock convs), which PyTorch enables by default for convolutions."""
    import torch
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
