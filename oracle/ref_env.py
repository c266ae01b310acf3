"""Import and run the UNMODIFIED reference (RayburnChen/senas).  TEST / BENCH INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()``, ``bench.py``'s reference / baseline arms and the acceptance scripts
may import this; the product (``senas_b200/``) never does (tests/test_host.py checks it).

The reference tree is looked up as ``$SENAS_REF``, ``/root/reference`` (build container), ``oracle/_ref`` (staged
byte-for-byte by ``oracle/make_ref.py``; this is what exists on the GPU box).  Nothing of it is modified: the
shims below only supply third-party modules that are not installed, two torchvision names
removed since torchvision 0.13, and a synthetic PROMISE12-shaped dataset (there is no data and no network).
"""
import importlib
import importlib.abc
import importlib.machinery
import importlib.util
import os
import runpy
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
CANDIDATES = [os.environ.get('SENAS_REF'), '/root/reference', os.path.join(HERE, '_ref')]

# packages the reference imports at module top that are not installed here (and are not needed by the search path)
STUB_PACKAGES = ('graphviz', 'ptflops', 'torchstat', 'tensorboardX', 'adabound', 'SimpleITK', 'skimage', 'pydicom',
                 'nibabel', 'imageio', 'matplotlib', 'segmentation_models_pytorch', 'pretrainedmodels', 'timm',
                 'efficientnet_pytorch', 'visdom', 'pynvml')


def root():
    for c in CANDIDATES:
        if c and os.path.isdir(os.path.join(c, 'search')):
            return c
    return None


def available():
    return root() is not None


def kind():
    """'reference' -- it is the reference's own code either way; says where it was found."""
    r = root()
    return None if r is None else ('oracle/_ref (staged copy of the unmodified reference)'
                                   if os.path.abspath(r) == os.path.join(HERE, '_ref') else r)


class _Anything:
    """Permissive stand-in: any attribute, any call, usable as a base class / context manager / iterator."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith('__') and name.endswith('__'):
            raise AttributeError(name)
        return _Anything()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def __iter__(self):
        return iter(())


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith('__') and name.endswith('__'):
            raise AttributeError(name)
        return _Anything


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def __init__(self, names):
        self.names = tuple(names)

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split('.')[0] in self.names:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


_installed = []


def install_shims():
    """Stub the third-party packages that really are missing (returns their names) and restore the two torchvision
    names the reference's augmentations still use (utils/augmentations/__init__.py:16,18)."""
    if _installed:
        return _installed[0]
    missing = []
    for name in STUB_PACKAGES:
        if name in sys.modules:
            continue
        try:
            if importlib.util.find_spec(name) is None:
                missing.append(name)
        except Exception:
            missing.append(name)
    sys.meta_path.append(_StubFinder(missing))
    _installed.append(tuple(missing))
    try:
        import torchvision.transforms as T
        import torchvision.transforms.transforms as TT
        for mod in (T, TT):
            if not hasattr(mod, 'Scale'):
                mod.Scale = T.Resize
            if not hasattr(mod, 'RandomSizedCrop'):
                mod.RandomSizedCrop = T.RandomResizedCrop
    except Exception:
        pass
    return _installed[0]


def load():
    """The reference's own modules: (search.cell, search.senas_search, utils.operations)."""
    r = root()
    if r is None:
        raise RuntimeError('reference tree not found (SENAS_REF, /root/reference, oracle/_ref): run oracle/make_ref.py '
                           'in the build container')
    install_shims()
    if r not in sys.path:
        sys.path.insert(0, r)
    cell = importlib.import_module('search.cell')
    ss = importlib.import_module('search.senas_search')
    ops = importlib.import_module('utils.operations')
    return cell, ss, ops


def load_loss():
    load()
    return importlib.import_module('utils.loss.loss')


def make_nas(depth=5, nodes=3, c=32, in_ch=1, n_classes=2, seed=0):
    """The reference's NAS exactly as experiments/search_arc.py:113-118 builds it from senas_promise12.yml."""
    import torch
    _, ss, _ = load()
    torch.manual_seed(seed)
    return ss.NAS(in_ch, c, n_classes, depth, meta_node_num=nodes, use_sharing=False, double_down_channel=False,
                  supervision=False, multi_gpus=False, device=torch.device('cpu'))


def make_search_step(model, lr_w=5e-3, lr_a=1e-4, grad_clip=5.0):
    """One search step of experiments/search_arc.py:252-293 (epoch >= alpha_begin) on the reference's own classes:
    Architecture.step(valid) then zero_grad / forward / dice_ce / backward / clip_grad_norm_ / SGD.step, with the
    optimizers of configs/senas/senas_promise12.yml.  Returns (step_fn, model_optimizer, arch_optimizer)."""
    import torch
    _, ss, _ = load()
    loss_mod = load_loss()
    crit = loss_mod.SegmentationLosses('dice_ce')
    w_opt = torch.optim.SGD(model.parameters(), lr=lr_w, momentum=0.9, weight_decay=3e-4)
    a_opt = torch.optim.Adam(model.arch_parameters(), lr=lr_a, betas=(0.5, 0.999), weight_decay=1e-3)
    arch = ss.Architecture(model, arch_optimizer=a_opt, criterion=crit)

    def step(xt, yt, xv, yv):
        arch.step(xv, yv)
        w_opt.zero_grad()
        loss = crit(model(xt), yt)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), grad_clip)
        w_opt.step()
        return loss

    return step, w_opt, a_opt


# ---------------------------------------------------------------------------------------------------------------
# the untouched driver: experiments/search_arc.py under runpy
# ---------------------------------------------------------------------------------------------------------------
def _synthetic_dataset_class(n_samples, size, seed):
    import torch
    from torch.utils import data

    class SyntheticPromise12(data.Dataset):
        """PROMISE12-shaped samples (utils/datasets/promise12.py:288-299,345-418): z-scored float32 1xSxS slice and a
        {0,1} long mask; attributes the driver reads (search_arc.py:85-86) as the real class provides them."""
        num_class, in_channels, NUM_CLASS, IN_CHANNELS = 2, 1, 2, 1
        class_weight = None

        def __init__(self, root=None, split='train', mode=None, **kw):
            g = torch.Generator().manual_seed(seed)
            self.x = torch.randn(n_samples, 1, size, size, generator=g)
            self.y = (torch.rand(n_samples, size, size, generator=g) > 0.8).long()

        def __len__(self):
            return self.x.shape[0]

        def __getitem__(self, i):
            return self.x[i], self.y[i]

    return SyntheticPromise12


def run_search_arc(workdir, epochs=1, n_samples=8, size=64, batch_size=2, gpu=False, alpha_begin=0, seed=1234,
                   before_run=None, sequential_sampler=True):
    """``runpy`` the reference's experiments/search_arc.py UNTOUCHED, in ``workdir/experiments`` (the driver uses
    cwd-relative paths), on a synthetic promise12 dataset and a copy of configs/senas/senas_promise12.yml in which
    only run-length knobs differ (epoch, batch_size, n_workers, alpha_begin, gpu).  ``before_run()`` is called after
    the reference modules are importable and before the driver starts (the place for senas_b200.patch_reference()).
    Returns the driver's globals (``search_network`` holds the finished SearchNetwork)."""
    import torch
    import yaml
    r = root()
    if r is None:
        raise RuntimeError('reference tree not found')
    load()
    import utils.datasets as ds
    import utils.utils as uu
    ds.datasets['promise12'] = _synthetic_dataset_class(n_samples, size, seed)
    uu.get_gpus_memory_info = lambda: (0, [0])           # utils/utils.py:146 shells nvidia-smi and parses its text
    if not (gpu and torch.cuda.is_available()):
        uu.gpu_memory = lambda n=0: ['cpu', '0 GB', '0 GB', '0 GB', '0 GB']
    import torch.utils.data.sampler as smp
    orig_sampler = smp.SubsetRandomSampler
    if sequential_sampler:
        # SubsetRandomSampler draws from the global RNG; a fixed order keeps two runs on different devices comparable

        class _InOrder(smp.Sampler):
            def __init__(self, indices, generator=None):
                self.indices = list(indices)

            def __iter__(self):
                return iter(self.indices)

            def __len__(self):
                return len(self.indices)
        smp.SubsetRandomSampler = _InOrder
    with open(os.path.join(r, 'configs', 'senas', 'senas_promise12.yml')) as fh:
        cfg = yaml.load(fh, Loader=yaml.FullLoader)
    s = cfg['searching']
    s['epoch'], s['batch_size'], s['n_workers'], s['alpha_begin'], s['gpu'] = epochs, batch_size, 0, alpha_begin, bool(gpu)
    s['report_freq'] = 1
    exp = os.path.join(workdir, 'experiments')
    os.makedirs(exp, exist_ok=True)
    cfg_path = os.path.join(workdir, 'senas_promise12_short.yml')
    with open(cfg_path, 'w') as fh:
        yaml.dump(cfg, fh)
    if before_run is not None:
        before_run()
    argv, cwd = sys.argv, os.getcwd()
    sys.argv = ['search_arc.py', '--config', cfg_path]
    os.chdir(exp)
    try:
        return runpy.run_path(os.path.join(r, 'experiments', 'search_arc.py'), run_name='__main__')
    finally:
        sys.argv = argv
        os.chdir(cwd)
        smp.SubsetRandomSampler = orig_sampler
