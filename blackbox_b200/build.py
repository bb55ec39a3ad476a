"""Build libbbx.so in-tree (blackbox_b200/_build/libbbx.so) with nvcc for sm_100a.

    python -m blackbox_b200.build [--force] [--verbose]

-fmad=false: the parity contract is one IEEE operation per arithmetic step (bit-exact against
the CPU oracle, which is compiled with -ffp-contract=off); default -prec-div / -prec-sqrt.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT_DIR = os.path.join(HERE, '_build')
LIB = os.path.join(OUT_DIR, 'libbbx.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-fmad=false',
         '-std=c++17', '--compiler-options', '-fPIC', '-Xptxas', '-v'] + os.environ.get('BBX_NVCC_EXTRA', '').split()


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def _deps_mtime():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    hdrs.append(os.path.join(HERE, '..', 'include', 'bbx.h'))
    return max(os.path.getmtime(h) for h in hdrs)


def _compile(src, force, verbose):
    obj = os.path.join(OUT_DIR, src[:-3] + '.o')
    path = os.path.join(CSRC, src)
    if (not force and os.path.exists(obj)
            and os.path.getmtime(obj) >= max(os.path.getmtime(path), _deps_mtime())):
        return obj, False, ''
    cmd = [NVCC] + FLAGS + ['-c', path, '-o', obj]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed for {}:\n{}'.format(src, res.stdout))
    if verbose:
        print(' '.join(cmd))
    return obj, True, res.stdout


def build(force=False, verbose=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(lambda s: _compile(s, force, verbose), sources()))
    objs = [r[0] for r in results]
    log = ''.join(r[2] for r in results)
    if log:
        with open(os.path.join(OUT_DIR, 'ptxas.log'), 'w') as fh:
            fh.write(log)
    if any(r[1] for r in results) or not os.path.exists(LIB):
        cmd = [NVCC, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a']
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if res.returncode != 0:
            raise RuntimeError('link failed:\n' + res.stdout)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
