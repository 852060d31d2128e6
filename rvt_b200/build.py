"""Build the C-ABI CUDA library in-tree: rvt_b200/lib/librvt_b200.so (sm_100a only).

nvcc cross-compiles without a GPU; the .so is git-ignored but travels with gpurun snapshots.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'csrc', 'capi.cu')
LIB_DIR = os.path.join(HERE, 'lib')
LIB = os.path.join(LIB_DIR, 'librvt_b200.so')

NVCC_FLAGS = ['-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
              '-Xcompiler', '-fPIC', '-shared']


def sources():
    d = os.path.join(HERE, 'csrc')
    inc = os.path.join(os.path.dirname(HERE), 'include')
    return [os.path.join(d, f) for f in sorted(os.listdir(d))] + \
           [os.path.join(inc, f) for f in sorted(os.listdir(inc))]


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = os.environ.get('NVCC', 'nvcc')
    tmp = LIB + '.building'
    cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-o', tmp, SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + ' '.join(cmd) + '\n' + res.stdout + res.stderr)
    os.replace(tmp, LIB)          # atomic: a concurrent snapshot never sees a half-written library
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
