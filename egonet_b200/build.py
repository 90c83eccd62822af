"""Build the native library in-tree: egonet_b200/lib/libegonet_b200.so.

    python -m egonet_b200.build            # incremental (per-file objects under egonet_b200/_build)
    python -m egonet_b200.build --force

nvcc cross-compiles for sm_100a without a GPU.  cudart is linked statically and
the driver API (cuTensorMapEncodeTiled) is resolved at run time through
cudaGetDriverEntryPoint, so the .so loads on a machine with no NVIDIA driver
(the CPU-side tests check its exported symbols there).
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(HERE, '_build')
LIB = os.path.join(HERE, 'lib', 'libegonet_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
SOURCES = ['common.cu', 'decode.cu', 'pose.cu', 'lifter.cu', 'conv_simt.cu', 'conv_tc.cu', 'hrnet_engine.cu', 'hrnet_train.cu', 'loss.cu', 'crop.cu', 'pnp.cu', 'eval.cu']
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
         '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden', '--expt-relaxed-constexpr',
         '-I', os.path.join(ROOT, 'include'), '-I', SRC]


def _deps_mtime():
    m = 0.0
    for d in (SRC, os.path.join(ROOT, 'include')):
        for f in os.listdir(d):
            if f.endswith(('.h', '.cuh')):
                m = max(m, os.path.getmtime(os.path.join(d, f)))
    return m


def _compile(src, force, hdr_mtime):
    obj = os.path.join(OBJ, src.replace('.cu', '.o'))
    path = os.path.join(SRC, src)
    if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(path), hdr_mtime):
        return obj, None
    cmd = [NVCC] + FLAGS + ['-Xptxas', '-v', '-c', path, '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (src, r.stdout[-4000:], r.stderr[-8000:]))
    return obj, r.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    hdr = _deps_mtime()
    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        results = list(ex.map(lambda s: _compile(s, force, hdr), SOURCES))
    objs = [o for o, _ in results]
    if verbose:
        for (_, log), s in zip(results, SOURCES):
            if log:
                print('==== %s\n%s' % (s, log))
    rebuilt = any(log is not None for _, log in results)
    if rebuilt or not os.path.exists(LIB):
        cmd = [NVCC, '-shared', '-o', LIB] + objs + ['-cudart', 'static', '-Xlinker', '--no-undefined',
                                                     '-lpthread', '-ldl', '-lrt']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('link failed:\n%s\n%s' % (r.stdout, r.stderr))
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
