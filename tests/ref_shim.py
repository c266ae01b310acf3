"""Import the unmodified reference (/root/reference) in the build container.  TEST HELPER ONLY.

The reference needs three third-party modules that are not installed (graphviz, ptflops, torchstat,
pulled in by utils/__init__.py and utils/utils.py:17-18); permissive stubs are enough for the
search path.  Nothing here runs on the GPU box: `/root/reference` does not
exist there and every caller is skipped when `available()` is False.
"""
import importlib
import os
import sys
import types

REF = os.environ.get('SENAS_REF', '/root/reference')


def available():
    return os.path.isdir(os.path.join(REF, 'search'))


def load():
    """Returns the reference's (search.cell, search.senas_search, utils.operations) modules."""
    if not available():
        raise RuntimeError('reference tree not present')
    for name, attrs in (('graphviz', ['Digraph']), ('ptflops', ['get_model_complexity_info']), ('torchstat', ['stat'])):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                m = types.ModuleType(name)
                for a in attrs:
                    setattr(m, a, lambda *x, **k: None)
                sys.modules[name] = m
    # our own top-level test package is called "tests"; the reference's packages are `search`, `utils`
    if REF not in sys.path:
        sys.path.insert(0, REF)
    cell = importlib.import_module('search.cell')
    ss = importlib.import_module('search.senas_search')
    ops = importlib.import_module('utils.operations')
    return cell, ss, ops
