"""
In-tree build of libh3d.so (hand-written sm_100a CUDA behind a C ABI).

    python -m hic3defdr_b200.build [--force]

nvcc cross-compiles without a GPU; the resulting .so sits next to this file so
that it travels to the GPU box with the repository snapshot.
"""
import concurrent.futures
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(HERE, 'build')
LIB = os.path.join(HERE, 'libh3d.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3',
         '-std=c++17', '-Xcompiler', '-fPIC']


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, 'rb') as f:
            h.update(os.path.basename(p).encode())   # path independent
            h.update(f.read())
    return h.hexdigest()


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)
          if f.endswith(('.cuh', '.h'))]
    hs.append(os.path.join(os.path.dirname(HERE), 'include', 'h3d.h'))
    return hs


def _compile(src):
    obj = os.path.join(OBJ, src[:-3] + '.o')
    stamp = obj + '.sha'
    want = _digest([os.path.join(CSRC, src)] + _headers() + [__file__])
    if os.path.exists(obj) and os.path.exists(stamp) and \
            open(stamp).read() == want:
        return obj, False
    cmd = [NVCC] + FLAGS + ['-c', os.path.join(CSRC, src), '-o', obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed for %s:\n%s\n%s'
                           % (src, res.stdout, res.stderr))
    with open(stamp, 'w') as f:
        f.write(want)
    return obj, True


def build(force=False, verbose=True):
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    srcs = _sources()
    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
        results = list(ex.map(_compile, srcs))
    objs = [o for o, _ in results]
    rebuilt = any(r for _, r in results)
    if rebuilt or not os.path.exists(LIB):
        cmd = [NVCC, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-Xcompiler', '-pthread',
               '-o', LIB] + objs
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError('link failed:\n%s\n%s' % (res.stdout,
                                                          res.stderr))
        if verbose:
            print('built %s (%d sources recompiled)'
                  % (LIB, sum(r for _, r in results)), file=sys.stderr)
    return LIB


if __name__ == '__main__':
    build(force='--force' in sys.argv)
