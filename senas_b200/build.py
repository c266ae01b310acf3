"""In-tree build of libsenas_b200.so: hand-written CUDA for sm_100a only (no other arch, no JIT cache)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'csrc', 'graph.cu')
DEPS = [SRC, os.path.join(HERE, 'csrc', 'kernels.cuh'), os.path.join(HERE, 'csrc', 'platform.h'), os.path.join(HERE, 'csrc', 'conv_tc.cuh'), os.path.join(HERE, 'csrc', 'comm.cuh'), os.path.join(HERE, 'csrc', 'optim.cuh'), os.path.join(HERE, 'csrc', 'convbn.cuh'),
        os.path.join(os.path.dirname(HERE), 'include', 'senas_b200.h')]
OUT = os.path.join(HERE, 'lib', 'libsenas_b200.so')
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-shared',
              '-Xcompiler', '-fPIC']


def build(force=False, verbose=False):
    """Compile the CUDA extension (cross-compiles without a GPU). Returns the .so path."""
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS):
        return OUT
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + [SRC, '-o', OUT]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + res.stdout)
    if verbose:
        print(res.stdout)
    return OUT
