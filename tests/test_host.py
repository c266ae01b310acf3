"""Host-side logic on the CPU: ABI surface, module tree, genotype parser, no-fallback behaviour, DP buckets."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import senas_b200
import senas_oracle as oracle
from helpers import golden
from senas_b200 import _lib
from senas_b200.ops import OpType

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_library_exports_every_declared_symbol():
    """The nvcc build loads without a GPU and exports exactly what include/senas_b200.h declares."""
    path = senas_b200.build()
    lib = ctypes.CDLL(path)
    header = open(os.path.join(ROOT, 'include', 'senas_b200.h')).read()
    declared = set(re.findall(r'\b(senas_[a-z_]+)\s*\(', header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for sym in declared:
        assert hasattr(lib, sym), sym
    bound = _lib.bind(path)
    assert b'sm_100a' in bound.senas_version()
    assert bound.senas_launch_count() == 0


def test_cuda_binary_is_sm100a_only():
    out = subprocess.run(['cuobjdump', '--list-elf', senas_b200.build()], capture_output=True, text=True).stdout
    archs = set(re.findall(r'sm_\d+a?', out))
    assert archs == {'sm_100a'}, archs


def test_no_cpu_fallback():
    m = senas_b200.MixedOp(32, 8, OpType.NORM)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        m(torch.randn(1, 32, 8, 8), torch.ones(6) / 6, torch.ones(6) / 6)
    c = senas_b200.Cell(3, 1, 32, 32, 32, 'up')
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        c(torch.randn(1, 32, 8, 8), torch.randn(1, 32, 4, 4), torch.ones(9, 6) / 6, torch.ones(9, 6) / 6, torch.ones(9))
    with pytest.raises(RuntimeError, match='no standalone'):
        m._ops[2](torch.randn(1, 32, 8, 8))


def test_product_package_never_imports_the_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, 'senas_b200')):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(root, f)).read()
                assert 'senas_oracle' not in src and 'import oracle' not in src and 'libsenas_emu' not in src, f


def test_module_tree_shape():
    """6718 state-dict entries / 1 967 798 parameters in 3367 tensors, candidate order."""
    torch.manual_seed(0)
    m = senas_b200.NAS(1, 32, 2, depth=5, meta_node_num=3, use_sharing=False, double_down_channel=False, supervision=False)
    assert len(m.state_dict()) == 6718
    ps = list(m.parameters())
    assert len(ps) == 3367 and sum(p.numel() for p in ps) == 1967798
    cell = m.net.blocks[1][0]
    types = [op._op_type.name for op in cell._ops]
    assert types == ['NORM', 'UP', 'NORM', 'UP', 'NORM', 'NORM', 'UP', 'NORM', 'NORM']
    assert [op._c_in for op in cell._ops] == [32, 32, 32, 32, 8, 32, 32, 8, 8]
    up = cell._ops[1]
    assert dict(up.named_parameters())['_ops.1.0.weight'].shape == (32, 8, 3, 3)      # ConvTranspose2d layout
    assert dict(up.named_parameters())['_ops.1.2.excitation.0.weight'].shape == (1, 8)
    assert 'conv.weight' not in dict(m.net.blocks[0][1]._ops[4]._ops[0].named_parameters())  # identity 8->8: bare BN


@pytest.mark.skipif(not os.path.isdir('/root/reference/search'), reason='reference tree only in the build container')
def test_module_tree_matches_reference_bit_for_bit():
    import ref_shim
    _, ss, _ = ref_shim.load()
    torch.manual_seed(0)
    ref = ss.NAS(1, 32, 2, depth=5, meta_node_num=3, use_sharing=False, double_down_channel=False, supervision=False,
                 device=torch.device('cpu'))
    torch.manual_seed(0)
    mine = senas_b200.NAS(1, 32, 2, depth=5, meta_node_num=3, use_sharing=False, double_down_channel=False, supervision=False)
    a, b = ref.state_dict(), mine.state_dict()
    assert list(a.keys()) == list(b.keys()) and all(torch.equal(a[k], b[k]) for k in a)
    assert [n for n, _ in ref.named_parameters()] == [n for n, _ in mine.named_parameters()]
    assert ref.genotype() == mine.genotype()


def test_genotype_parser_matches_oracle_on_random_tables():
    rng = np.random.default_rng(0)
    parser = senas_b200.GenoParser(3)
    for _ in range(50):
        w1 = rng.random((9, 6)).astype(np.float32)
        w2 = rng.random((9, 6)).astype(np.float32)
        for ct in ('down', 'up'):
            assert parser.parse(w1, w2, ct) == oracle.parse_cell(w1, w2, ct, 3)


def test_genotype_of_golden_arch_tables():
    g = golden('nas_search_2steps')
    torch.manual_seed(0)
    m = senas_b200.NAS(1, 32, 2, depth=5, meta_node_num=3, use_sharing=False, double_down_channel=False, supervision=False)
    with torch.no_grad():
        for n in ('alphas_dn', 'alphas_up', 'alphas_dn_nm', 'alphas_up_nm', 'betas_dn', 'betas_up', 'gamma'):
            getattr(m, n).copy_(torch.from_numpy(g['arch.' + n]))
    assert repr(m.genotype()) == str(g['genotype'])


def test_graph_descriptor_validation():
    """The C ABI rejects malformed graphs with a message instead of crashing (no device needed)."""
    lib = _lib.bind(senas_b200.build())
    d = _lib.GraphDesc()
    d.n_inputs, d.n_nodes, d.n_edges, d.c_out = 1, 1, 1, 16
    h = ctypes.c_void_p()
    assert lib.senas_graph_create(ctypes.byref(d), ctypes.byref(h)) != 0
    assert b'c_out' in lib.senas_last_error()
    d.c_out, d.edge[0].c_in, d.edge[0].op_type = 8, 24, 3
    assert lib.senas_graph_create(ctypes.byref(d), ctypes.byref(h)) != 0
    assert b'c_in' in lib.senas_last_error()


def test_unsupported_configurations_fail_at_construction():
    """the reference's NAS defaults (meta_node_num=4, double_down_channel=True) are outside what the kernels
    cover; say so in the constructor, not at the first forward.  'max_pool' / 'conv_3' are registry keys of the reference
    that no candidate list uses."""
    import pytest
    import senas_b200
    from senas_b200.ops import OPS, OpType
    with pytest.raises(NotImplementedError, match='meta_node_num'):
        senas_b200.Cell(4, 1, 32, 32, 32, 'up')
    with pytest.raises(NotImplementedError, match='32-channel'):
        senas_b200.Cell(3, 2, 32, 32, 64, 'up')
    with pytest.raises(NotImplementedError):
        senas_b200.NAS(1, 32, 2, depth=5)  # the reference's defaults: 4 nodes, doubled channels
    for name in ('max_pool', 'conv_3'):
        with pytest.raises(NotImplementedError, match='candidate lists'):
            OPS[name](32, 8, OpType.NORM, 0)
    senas_b200.NAS(1, 32, 2, depth=3, meta_node_num=3, use_sharing=False, double_down_channel=False)  # supported


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the contract's reference arm: the unmodified reference on the host cores) prints ONE JSON
    line with the agreed keys; run here on a tiny sample (2 x 1x64x64, one step)."""
    import json
    res = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--size', '64', '--batch', '2',
                          '--steps', '1', '--warmup', '1'], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1, res.stdout
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'search_step_images_per_sec' and d['unit'] == 'images/s'
    assert d['higher_is_better'] is True and d['value'] > 0 and d['steps'] == 1 and d['warmup'] == 1
    assert d['cpu_baseline']['kind'] == 'reference' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'unmodified reference' in d['cpu_baseline']['sample']
