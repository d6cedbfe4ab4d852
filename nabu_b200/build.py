"""Builds nabu_b200/libnabu_b200.so (the C-ABI in include/nabu_b200.h) with nvcc for sm_100a.

In-tree on purpose: the .so travels with the repo snapshot to the GPU box; nothing is JIT-compiled
at run time.  `python -m nabu_b200.build` rebuilds what is stale; `--force` rebuilds everything.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(HERE, 'csrc', 'build')
LIB = os.path.join(HERE, 'libnabu_b200.so')

NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
FLAGS = ['-O3', '-lineinfo', '-std=c++17', '-Xcompiler', '-fPIC',
         '-I', os.path.join(ROOT, 'include'), '-I', CSRC]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, log=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.h', '.cuh'))]
    headers.append(os.path.join(ROOT, 'include', 'nabu_b200.h'))
    jobs = []
    objs = []
    for src in _sources():
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src[:-3] + '.o')
        objs.append(o)
        if force or _stale(o, [s] + headers):
            jobs.append([NVCC] + ARCH + FLAGS + ['-c', s, '-o', o])

    def run(cmd):
        if verbose:
            print(' '.join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed:\n%s\n%s' % (' '.join(cmd), r.stderr))
        return r

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(run, jobs))
    link = None
    if jobs or force or _stale(LIB, objs):
        link = [NVCC] + ARCH + ['-shared', '-o', LIB] + objs + ['-lcudart_static', '-lpthread', '-ldl', '-lrt']
        run(link)
    if log:
        import hashlib
        import time
        with open(os.path.join(OBJ, 'build.log'), 'w') as fid:
            fid.write('# %s: %d translation units compiled, library %s\n' % (time.strftime('%Y-%m-%d %H:%M:%S'), len(jobs),
                                                                              'linked' if link else 'up to date'))
            for cmd in jobs + ([link] if link else []):
                fid.write(' '.join(cmd) + '\n')
            fid.write('sha256 %s  %s\n' % (hashlib.sha256(open(LIB, 'rb').read()).hexdigest(), os.path.relpath(LIB, ROOT)))
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
