"""Build tests/emu/libsenas_emu.so: the kernel sources compiled by g++ against the fiber emulator
(cpu_emu.h).  TEST INFRASTRUCTURE ONLY -- used by tests/test_emu_kernels.py to check kernel logic
against the oracle in a container without a GPU; never loaded by the senas_b200 package."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, 'senas_b200', 'csrc')
OUT = os.path.join(HERE, 'libsenas_emu.so')


def build(force=False):
    deps = [os.path.join(CSRC, f) for f in ('graph.cu', 'kernels.cuh', 'platform.h')] + [
        os.path.join(HERE, 'cpu_emu.h'), os.path.join(ROOT, 'include', 'senas_b200.h')]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    cmd = ['g++', '-O2', '-g', '-std=c++17', '-x', 'c++', '-DSENAS_EMU', '-I' + HERE, '-I' + CSRC, '-shared', '-fPIC',
           os.path.join(CSRC, 'graph.cu'), '-o', OUT]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError('emulator build failed:\n' + res.stdout)
    return OUT
