"""Build libdbbcuda.so in-tree with nvcc for sm_100a (no other architecture is supported).

``python -m dbb_b200.build`` or ``__graft_entry__.build()``.  Objects are rebuilt only when the source
(or a header) is newer; translation units compile in parallel.
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(CSRC, '_obj')
LIB = os.path.join(HERE, 'libdbbcuda.so')
SOURCES = ['table.cu', 'pack.cu', 'count.cu', 'pairs.cu', 'refine.cu', 'prune.cu', 'knn.cu', 'greedy.cu', 'pca.cu', 'split.cu', 'synth.cu', 'host_api.cu',
           'hostpack.cpp']          # (.cpp: host-only code with x86 target attributes, compiled by g++ directly)
HEADERS = [os.path.join(CSRC, 'common.cuh'), os.path.join(HERE, '..', 'include', 'dbb_cuda.h')]
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
CXX = os.environ.get('CXX', 'g++')
CXXFLAGS = ['-O3', '-std=c++17', '-fPIC', '-fvisibility=hidden', '-Wall']
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
         '-Xcompiler', '-fPIC,-fvisibility=hidden,-ffp-contract=off', '--expt-relaxed-constexpr']


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src, verbose):
    obj = os.path.join(OBJ, os.path.splitext(src)[0] + '.o')
    path = os.path.join(CSRC, src)
    if not _stale(obj, [path] + HEADERS):
        return obj, ''
    if src.endswith('.cpp'):
        cmd = [CXX] + CXXFLAGS + ['-c', path, '-o', obj]
    else:
        cmd = [NVCC] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', path, '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('compiler failed for %s:\n%s\n%s' % (src, r.stdout, r.stderr))
    return obj, r.stderr


def build(verbose=False, force=False):
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            if os.path.isfile(os.path.join(OBJ, f)):
                os.remove(os.path.join(OBJ, f))
    with concurrent.futures.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        results = list(ex.map(lambda s: _compile(s, verbose), SOURCES))
    objs = [o for o, _ in results]
    log = ''.join(l for _, l in results)
    if force or _stale(LIB, objs):
        cmd = [NVCC, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-cudart', 'static', '-o', LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('link failed:\n%s\n%s' % (r.stdout, r.stderr))
    if verbose:
        print(log)
    return LIB


if __name__ == '__main__':
    print(build(verbose='-v' in sys.argv, force='-f' in sys.argv))
