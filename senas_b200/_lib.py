"""ctypes binding of ``libsenas_b200.so`` (C ABI in ``include/senas_b200.h``).

The library is built in-tree by ``senas_b200.build.build()`` (nvcc, sm_100a only) and loaded
from ``senas_b200/lib/``.  There is no fallback: if the shared object is missing, or the device
is not sm_100, importing the hot path raises.
"""
import ctypes as C
import os

MAX_CAND, SLOTS, MAX_EDGES, MAX_NODES = 6, 12, 16, 3
OP_UP, OP_DOWN, OP_NORM = 1, 2, 3

LIB_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'lib')
LIB_PATH = os.path.join(LIB_DIR, 'libsenas_b200.so')


class EdgeDesc(C.Structure):
    _fields_ = [('src', C.c_int32), ('dst', C.c_int32), ('op_type', C.c_int32), ('c_in', C.c_int32),
                ('kind', C.c_int32 * MAX_CAND), ('ksize', C.c_int32 * MAX_CAND), ('dilation', C.c_int32 * MAX_CAND),
                ('param', (C.c_void_p * SLOTS) * MAX_CAND), ('grad_off', (C.c_int64 * SLOTS) * MAX_CAND)]


class GraphDesc(C.Structure):
    _fields_ = [('n_inputs', C.c_int32), ('n_nodes', C.c_int32), ('n_edges', C.c_int32), ('c_out', C.c_int32),
                ('node_relu', C.c_int32), ('reserved', C.c_int32), ('grad_floats', C.c_int64),
                ('edge', EdgeDesc * MAX_EDGES)]


class PlanInfo(C.Structure):
    _fields_ = [('out_h', C.c_int32), ('out_w', C.c_int32), ('saved_bytes', C.c_int64), ('scratch_bytes', C.c_int64)]


class FwdArgs(C.Structure):
    _fields_ = [('batch', C.c_int32), ('training', C.c_int32), ('in_h', C.c_int32 * 2), ('in_w', C.c_int32 * 2),
                ('in_', C.c_void_p * 2), ('in_ld', C.c_int64 * 2), ('alpha', C.c_void_p), ('beta', C.c_void_p),
                ('out', C.c_void_p), ('out_ld', C.c_int64), ('saved', C.c_void_p), ('scratch', C.c_void_p),
                ('stream', C.c_void_p)]


class BwdArgs(C.Structure):
    _fields_ = [('batch', C.c_int32), ('training', C.c_int32), ('in_h', C.c_int32 * 2), ('in_w', C.c_int32 * 2),
                ('in_', C.c_void_p * 2), ('in_ld', C.c_int64 * 2), ('alpha', C.c_void_p), ('beta', C.c_void_p),
                ('out', C.c_void_p), ('out_ld', C.c_int64), ('grad_out', C.c_void_p), ('grad_out_ld', C.c_int64),
                ('saved', C.c_void_p), ('scratch', C.c_void_p), ('grad_in', C.c_void_p * 2),
                ('grad_in_ld', C.c_int64 * 2), ('grad_alpha', C.c_void_p), ('grad_beta', C.c_void_p),
                ('grad_params', C.c_void_p), ('stream', C.c_void_p), ('skip_wgrad', C.c_int32), ('reserved_', C.c_int32)]


class ConvBnArgs(C.Structure):
    _fields_ = [('batch', C.c_int32), ('h', C.c_int32), ('w', C.c_int32), ('c_in', C.c_int32), ('relu_in', C.c_int32),
                ('training', C.c_int32), ('x', C.c_void_p), ('x_ld', C.c_int64), ('weight', C.c_void_p), ('gamma', C.c_void_p),
                ('beta', C.c_void_p), ('running_mean', C.c_void_p), ('running_var', C.c_void_p),
                ('num_batches_tracked', C.c_void_p), ('momentum', C.c_float), ('eps', C.c_float), ('out', C.c_void_p),
                ('saved', C.c_void_p), ('scratch', C.c_void_p), ('grad_out', C.c_void_p), ('grad_out_ld', C.c_int64),
                ('grad_x', C.c_void_p), ('grad_weight', C.c_void_p), ('grad_gamma', C.c_void_p), ('grad_beta', C.c_void_p),
                ('stream', C.c_void_p)]


EXPORTS = ['senas_version', 'senas_last_error', 'senas_device_check', 'senas_graph_create', 'senas_graph_destroy',
           'senas_graph_plan', 'senas_graph_forward', 'senas_graph_backward', 'senas_avgpool_forward',
           'senas_avgpool_backward', 'senas_launch_count', 'senas_set_lanes', 'senas_set_slot', 'senas_set_defer', 'senas_flush', 'senas_set_ds_fused', 'senas_set_z_bfloat', 'senas_set_gather_mma', 'senas_comm_unique_id', 'senas_comm_init',
           'senas_comm_allreduce', 'senas_comm_destroy', 'senas_sgd_clip_step', 'senas_adam_step', 'senas_dice_ce_forward', 'senas_dice_ce_backward', 'senas_mix_forward',
           'senas_mix_backward', 'senas_mix_dx', 'senas_convbn_workspace', 'senas_convbn_forward', 'senas_convbn_backward',
           'senas_profile',
           'senas_profile_dump']


def bind(path):
    """Load a build of the C ABI and declare its prototypes."""
    lib = C.CDLL(path)
    lib.senas_version.restype = C.c_char_p
    lib.senas_last_error.restype = C.c_char_p
    lib.senas_device_check.argtypes = [C.c_int]
    lib.senas_graph_create.argtypes = [C.POINTER(GraphDesc), C.POINTER(C.c_void_p)]
    lib.senas_graph_destroy.argtypes = [C.c_void_p]
    lib.senas_graph_destroy.restype = None
    lib.senas_graph_plan.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                     C.POINTER(PlanInfo)]
    lib.senas_graph_forward.argtypes = [C.c_void_p, C.POINTER(FwdArgs)]
    lib.senas_graph_backward.argtypes = [C.c_void_p, C.POINTER(BwdArgs)]
    lib.senas_launch_count.restype = C.c_int64
    lib.senas_avgpool_forward.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                          C.c_void_p]
    lib.senas_avgpool_backward.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    lib.senas_set_lanes.argtypes = [C.c_int]
    lib.senas_set_slot.argtypes = [C.c_int]
    lib.senas_set_defer.argtypes = [C.c_int]
    lib.senas_set_ds_fused.argtypes = [C.c_int]
    lib.senas_set_z_bfloat.argtypes = [C.c_int]
    lib.senas_set_gather_mma.argtypes = [C.c_int]
    lib.senas_comm_unique_id.argtypes = [C.c_void_p]
    lib.senas_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    lib.senas_comm_allreduce.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    lib.senas_comm_destroy.argtypes = [C.c_void_p]
    lib.senas_flush.argtypes = [C.c_void_p]
    lib.senas_dice_ce_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                          C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.senas_dice_ce_backward.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.senas_convbn_workspace.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.senas_convbn_forward.argtypes = [C.POINTER(ConvBnArgs)]
    lib.senas_convbn_backward.argtypes = [C.POINTER(ConvBnArgs)]
    lib.senas_mix_dx.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
                                 C.c_void_p, C.c_int32, C.c_int64, C.c_void_p]
    lib.senas_sgd_clip_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_float, C.c_float,
                                        C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.senas_adam_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                    C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p]
    lib.senas_mix_forward.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
                                      C.c_int32, C.c_int32, C.c_int64, C.c_void_p]
    lib.senas_mix_backward.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
                                       C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p]
    lib.senas_profile.argtypes = [C.c_int]
    lib.senas_profile_dump.argtypes = [C.c_char_p, C.c_int64]
    lib.senas_profile_dump.restype = C.c_int64
    return lib


_LIB = None


def get():
    """The product library.  Raises if it has not been built -- there is no other code path."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                '(nvcc, sm_100a). senas_b200 has no PyTorch/CPU fallback for the MixedOp/Cell path.')
        _LIB = bind(LIB_PATH)
    return _LIB


def profile_dump(lib=None):
    """{family: dict(launches, ms, flops, bytes)} of the last senas_profile recording."""
    lib = lib or get()
    n = lib.senas_profile_dump(None, 0)
    buf = C.create_string_buffer(n + 1)
    lib.senas_profile_dump(buf, n + 1)
    out = {}
    for line in buf.value.decode().splitlines():
        name, cnt, ms, fl, by = line.split()
        out[name] = dict(launches=int(cnt), ms=float(ms), flops=float(fl), bytes=float(by))
    return out


def check(lib, rc):
    if rc != 0:
        raise RuntimeError('senas_b200: ' + lib.senas_last_error().decode())
