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


SHARED_SYMBOLS = ('pb200_abi_version', 'pb200_last_error_string', 'pb200_num_sms')    # runtime.cu: exist once


def _write_f16_symbols() -> None:
    """csrc/f16_symbols.h: the renames that let the float16 build (-DPB200_HALF) of every kernel file link next to the
    bfloat16 build: namespace pb200 -> pb200_h and the suffix _f16 on every entry point declared in include/pasero_b200.h."""
    import re
    with open(os.path.join(ROOT, 'include', 'pasero_b200.h')) as f:
        names = sorted(set(re.findall(r'\b(pb200_[a-z0-9_]+)\s*\(', f.read())))
    lines = ['// pasero_b200 — symbol names of the float16 build of the kernels. Every .cu of this directory is compiled twice: as it',
             '// stands (16-bit element type = bfloat16) and with -DPB200_HALF (IEEE half), where this header renames the namespace and',
             '// every exported entry point (suffix _f16) so that both builds link into one library. The three library-wide entry points',
             '// of runtime.cu exist once. Generated from include/pasero_b200.h by pasero_b200/build.py: do not edit.',
             '#pragma once', '#define pb200 pb200_h']
    lines += [f'#define {n} {n}_f16' for n in names if n not in SHARED_SYMBOLS]
    text = '\n'.join(lines) + '\n'
    path = os.path.join(CSRC, 'f16_symbols.h')
    if not os.path.exists(path) or open(path).read() != text:
        with open(path, 'w') as f:
            f.write(text)


def _digest(path: str, half: bool) -> str:
    h = hashlib.sha1()
    deps = [path] + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cuh') or f.endswith('.h'))
    deps.append(os.path.join(ROOT, 'include', 'pasero_b200.h'))
    for dep in deps:
        with open(dep, 'rb') as f:
            h.update(f.read())
    h.update(' '.join(NVCC_FLAGS + (['-DPB200_HALF'] if half else [])).encode())
    return h.hexdigest()


def _compile(src: str, verbose: bool, half: bool = False) -> str:
    obj = os.path.join(BUILD, os.path.basename(src)[:-3] + ('_f16.o' if half else '.o'))
    stamp = obj + '.sha1'
    digest = _digest(src, half)
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == digest:
        return obj
    cmd = [_nvcc()] + NVCC_FLAGS + (['-DPB200_HALF'] if half else []) + ['-c', src, '-o', obj]
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
    _write_f16_symbols()
    # every kernel file twice: bfloat16 (as it stands) and float16 (-DPB200_HALF); runtime.cu (library-wide plumbing) once
    jobs = [(s, False) for s in _sources()] + [(s, True) for s in _sources() if os.path.basename(s) != 'runtime.cu']
    with ThreadPoolExecutor(max_workers=min(os.cpu_count() or 8, len(jobs))) as ex:
        objs = list(ex.map(lambda j: _compile(j[0], verbose, j[1]), jobs))
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        cmd = [_nvcc(), '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a']
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f'link failed:\n{res.stdout}\n{res.stderr}')
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
