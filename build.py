"""Build the in-tree native artefacts (also called by __graft_entry__.build()).

  * dynast-release_b200/libdynast_b200.so  -- the product: CUDA kernels + C-ABI, sm_100a only
  * oracle/_build/liboracle.so             -- TEST INFRASTRUCTURE: plain-C restatement of the reference

Both are git-ignored and travel to the GPU box with the snapshot.
"""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, 'dynast-release_b200')
LIB = os.path.join(PKG, 'libdynast_b200.so')
ORACLE_LIB = os.path.join(ROOT, 'oracle', '_build', 'liboracle.so')

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '-Xptxas=-v', '-Xcompiler',
    '-fPIC', '-Xcompiler', '-O3'
]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_product(force=False, verbose=False):
    srcs = sorted(glob.glob(os.path.join(PKG, 'csrc', '*.cu')) + glob.glob(os.path.join(PKG, 'csrc', '*.cpp')))
    deps = srcs + glob.glob(os.path.join(PKG, 'csrc', '*.cuh')) + glob.glob(os.path.join(ROOT, 'include', '*.h'))
    if not force and not _newer(LIB, deps):
        return LIB
    from concurrent.futures import ThreadPoolExecutor

    def compile_one(src):
        obj = os.path.join(PKG, 'csrc', os.path.splitext(os.path.basename(src))[0] + '.o')
        if force or _newer(obj, [src] + [d for d in deps if d.endswith(('.cuh', '.h'))]):
            cmd = ['nvcc'] + NVCC_FLAGS + ['-c', src, '-o', obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if verbose or r.returncode:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode:
                raise RuntimeError(f'nvcc failed on {src}')
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        objs = list(pool.map(compile_one, srcs))
    cmd = ['nvcc', '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-o', LIB] + objs + ['-lz', '-lpthread']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError('link failed')
    return LIB


def build_oracle(force=False):
    srcs = sorted(glob.glob(os.path.join(ROOT, 'oracle', '*.c')))
    if not force and not _newer(ORACLE_LIB, srcs):
        return ORACLE_LIB
    os.makedirs(os.path.dirname(ORACLE_LIB), exist_ok=True)
    # -ffp-contract=off: the reference (numba/LLVM without fastmath) never fuses multiply-add
    cmd = ['gcc', '-O2', '-ffp-contract=off', '-fno-fast-math', '-shared', '-fPIC', '-o', ORACLE_LIB] + srcs + ['-lm']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError('oracle build failed')
    return ORACLE_LIB


if __name__ == '__main__':
    force = '--force' in sys.argv
    print(build_oracle(force))
    print(build_product(force, verbose='-v' in sys.argv))
