"""Compile liboqg.so (the CUDA kernels + C ABI) in-tree for sm_100a.

Staleness is decided by CONTENT: the SHA-256 of every source under csrc/,
of include/oqg.h and of the compiler flags is stored beside the library
(liboqg.stamp) when it is built; the library is rebuilt whenever the stamp
differs (an edited source, other flags, another nvcc path).  Modification
times are only the fallback for a library that has no stamp."""
import os
import hashlib
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'csrc', 'oqg.cu')
CSRC = os.path.join(HERE, 'csrc')
DEPS = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)
              if f.endswith(('.cu', '.cuh', '.h'))) + [
    os.path.join(HERE, '..', 'include', 'oqg.h')]
LIB = os.path.join(HERE, 'liboqg.so')
STAMP = os.path.join(HERE, 'liboqg.stamp')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo',
         '-std=c++17', '-Xcompiler', '-fPIC', '-shared']


def fingerprint():
    """SHA-256 over the flags and the contents of the sources"""
    h = hashlib.sha256(' '.join(FLAGS).encode())
    for d in DEPS:
        h.update(os.path.basename(d).encode())
        with open(d, 'rb') as f:
            h.update(f.read())
    return h.hexdigest()


kernel_fingerprint = fingerprint   # identifies the build a profile belongs to


def stale():
    if not os.path.exists(LIB):
        return True
    if os.path.exists(STAMP):
        with open(STAMP) as f:
            return f.read().strip() != fingerprint()
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
    with open(STAMP, 'w') as f:
        f.write(fingerprint() + '\n')
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == '__main__':
    print(build(force=True, verbose=True))
