#!/usr/bin/env python
"""Recipe for ``oracle/_ref/``: the UNMODIFIED reference, staged so that it travels to the GPU box.

TEST / BENCH INFRASTRUCTURE ONLY.  The reference (RayburnChen/senas) is pure Python: there is nothing to
compile, so "building" it means staging the files of the search path -- byte for byte -- where the GPU box can
import them (``/root/reference`` does not exist there).  ``oracle/_ref/`` is git-ignored (no reference source
ever enters the history) but not gpurun-ignored.  Run by ``__graft_entry__.build()`` whenever
``/root/reference`` is present; on the GPU box the staged copy is used as is.

What is staged: ``search/``, ``utils/``, ``models/``, ``experiments/``, ``configs/`` (``*.py`` / ``*.yml`` only,
~0.4 MB).  Not staged: vendored ``segmentation_models_pytorch`` (baseline models, stubbed at import), ``kohonen``,
binaries, images, slides.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, '_ref')
SRC_DEFAULT = os.environ.get('SENAS_REF_SRC', '/root/reference')
TREES = ('search', 'utils', 'models', 'experiments', 'configs')
EXTS = ('.py', '.yml', '.yaml')


def make(src=SRC_DEFAULT, dst=DST, quiet=False):
    """Copy the search-path files of ``src`` into ``dst``; returns the manifest {relative path: sha256}."""
    if not os.path.isdir(os.path.join(src, 'search')):
        raise FileNotFoundError(f'no reference tree at {src}')
    manifest = {}
    for tree in TREES:
        for root, _dirs, files in os.walk(os.path.join(src, tree)):
            for f in sorted(files):
                if not f.endswith(EXTS):
                    continue
                sp = os.path.join(root, f)
                rel = os.path.relpath(sp, src)
                dp = os.path.join(dst, rel)
                os.makedirs(os.path.dirname(dp), exist_ok=True)
                shutil.copyfile(sp, dp)
                with open(dp, 'rb') as fh:
                    manifest[rel] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(dst, 'MANIFEST.json'), 'w') as fh:
        json.dump({'source': src, 'files': manifest}, fh, indent=1, sort_keys=True)
    if not quiet:
        print(f'oracle/_ref: staged {len(manifest)} files from {src}')
    return manifest


if __name__ == '__main__':
    make(*(sys.argv[1:2] or [SRC_DEFAULT]))
