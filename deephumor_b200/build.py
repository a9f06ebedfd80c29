"""Builds deephumor_b200/libdeephumor_sm100.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m deephumor_b200.build [--force] [--verbose]
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(HERE, 'csrc', 'build')
LIB = os.path.join(HERE, 'libdeephumor_sm100.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '-Xcompiler', '-fPIC',
         '-Xcompiler', '-Wall', '--expt-relaxed-constexpr', '-Xptxas', '-v']


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def _digest():
    h = hashlib.sha256(' '.join(FLAGS).encode())
    for f in sorted(os.listdir(CSRC)) + ['../../include/deephumor_b200.h']:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p):
            h.update(open(p, 'rb').read())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, 'stamp')
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    if not os.path.exists(NVCC):
        raise RuntimeError(f'nvcc not found at {NVCC}; cannot build {LIB}')

    def compile_one(src):
        obj = os.path.join(OBJ, src[:-3] + '.o')
        cmd = [NVCC] + FLAGS + ['-c', os.path.join(CSRC, src), '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'nvcc failed on {src}:\n{r.stdout}\n{r.stderr}')
        with open(os.path.join(OBJ, src[:-3] + '.ptxas.log'), 'w') as f:
            f.write(r.stderr)
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(compile_one, _sources()))
    r = subprocess.run([NVCC, '-shared', '-o', LIB] + objs, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
    open(stamp, 'w').write(dig)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
