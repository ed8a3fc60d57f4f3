"""Builds pasero_b200/libpasero_b200.so from csrc/*.cu with nvcc for sm_100a (in-tree, no JIT cache).

The .so carries a static CUDA runtime and resolves cuTensorMapEncodeTiled through
cudaGetDriverEntryPoint, so it has no link-time dependency on libcuda and cross-compiles on a box
without a GPU.
"""
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
BUILD = os.path.join(ROOT, 'build', 'pasero_b200')
LIB = os.path.join(HERE, 'libpasero_b200.so')

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
    '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr', '-Xptxas', '-v',
] + os.environ.get('PB200_NVCC_EXTRA', '').split()


def _nvcc() -> str:
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        raise RuntimeError('nvcc not found: cannot build libpasero_b200.so')
    return nvcc


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def _digest(path: str) -> str:
    h = hashlib.sha1()
    deps = [path] + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cuh'))
    deps.append(os.path.join(ROOT, 'include', 'pasero_b200.h'))
    for dep in deps:
        with open(dep, 'rb') as f:
            h.update(f.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile(src: str, verbose: bool) -> str:
    obj = os.path.join(BUILD, os.path.basename(src)[:-3] + '.o')
    stamp = obj + '.sha1'
    digest = _digest(src)
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == digest:
        return obj
    cmd = [_nvcc()] + NVCC_FLAGS + ['-c', src, '-o', obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = obj + '.log'
    with open(log, 'w') as f:
        f.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError(f'nvcc failed for {src}:\n{res.stdout}\n{res.stderr}')
    if verbose:
        sys.stderr.write(res.stderr)
    with open(stamp, 'w') as f:
        f.write(digest)
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    if force:
        for f in os.listdir(BUILD):
            os.remove(os.path.join(BUILD, f))
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), srcs))
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        cmd = [_nvcc(), '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a']
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f'link failed:\n{res.stdout}\n{res.stderr}')
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
