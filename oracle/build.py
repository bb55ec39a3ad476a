"""Build the oracle's C library (oracle/_build/libbbo.so) with gcc.

-ffp-contract=off keeps every float32 operation a separate IEEE operation (no FMA), which is
what the CUDA kernels are compiled to match (-fmad=false).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'csrc', 'bbo.c')
OUT_DIR = os.path.join(HERE, '_build')
OUT = os.path.join(OUT_DIR, 'libbbo.so')


def build(force=False, verbose=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if (not force and os.path.exists(OUT)
            and os.path.getmtime(OUT) >= os.path.getmtime(SRC)):
        return OUT
    cmd = ['gcc', '-O2', '-fopenmp', '-ffp-contract=off', '-fno-fast-math', '-fPIC',
           '-shared', '-fvisibility=hidden', '-std=c99', '-o', OUT, SRC, '-lm']
    if verbose:
        print(' '.join(cmd))
    subprocess.check_call(cmd)
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
