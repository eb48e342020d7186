"""In-tree build of libtheanet_b200.so for sm_100a (nvcc cross-compiles without a GPU).

    python -m theanet_b200.build [--force]

The shared library lands in theanet_b200/lib/ (git-ignored, shipped to the GPU box by gpurun).
Objects are rebuilt only when their source (or a header) is newer.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, 'lib')
LIB = os.path.join(LIBDIR, 'libtheanet_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
         '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr', '-Xptxas', '-v'] + \
    os.environ.get('TN_NVCC_EXTRA', '').split()


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    hs.append(os.path.join(ROOT, 'include', 'theanet_b200.h'))
    return hs


def _newer(a, b):
    return (not os.path.exists(b)) or os.path.getmtime(a) > os.path.getmtime(b)


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(LIBDIR, 'obj')
    os.makedirs(objdir, exist_ok=True)
    hdr_time = max(os.path.getmtime(h) for h in headers())
    jobs = []
    objs = []
    for src in sources():
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src[:-3] + '.o')
        objs.append(o)
        if force or _newer(s, o) or hdr_time > os.path.getmtime(o):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        cmd = [NVCC] + FLAGS + ['-c', s, '-o', o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = os.path.join(objdir, os.path.basename(s) + '.ptxas.log')
        with open(log, 'w') as f:
            f.write(r.stderr)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed for {}:\n{}'.format(s, r.stderr))
        if verbose:
            print(r.stderr)
        return o

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(compile_one, jobs))
    if jobs or not os.path.exists(LIB):
        cmd = [NVCC, '-shared', '-o', LIB] + objs + ['-lcudart_static', '-ldl', '-lrt', '-lpthread']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('link failed:\n' + r.stderr)
    return LIB


if __name__ == '__main__':
    lib = build(force='--force' in sys.argv, verbose='-v' in sys.argv)
    print(lib)
