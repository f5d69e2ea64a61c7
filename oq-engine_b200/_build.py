"""Compile liboqg.so (the CUDA kernels + C ABI) in-tree for sm_100a."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'csrc', 'oqg.cu')
CSRC = os.path.join(HERE, 'csrc')
DEPS = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)
              if f.endswith(('.cu', '.cuh', '.h'))) + [
    os.path.join(HERE, '..', 'include', 'oqg.h')]
LIB = os.path.join(HERE, 'liboqg.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo',
         '-std=c++17', '-Xcompiler', '-fPIC', '-shared']


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS if os.path.exists(d))


def build(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> liboqg.so"""
    if not force and not stale():
        return LIB
    cmd = [NVCC] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + [
        '-o', LIB, SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed:\n%s\n%s' % (res.stdout, res.stderr))
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == '__main__':
    print(build(force=True, verbose=True))
