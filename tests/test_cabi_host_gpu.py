"""The C ABI used the way a non-Python host would: tests/cabi/host_c_abi.c (plain C, cudaMalloc'd
buffers, include/bbx.h) is compiled with gcc against libbbx.so and run; it checks
bbx_stack_median, bbx_xtalk and the FITS codec against scalar loops of its own."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA = os.environ.get('CUDA_HOME', '/usr/local/cuda')


def _compile(tmp_path):
    from blackbox_b200 import build
    lib = build.build()
    exe = str(tmp_path / 'host_c_abi')
    cmd = ['gcc', '-O1', '-Wall', '-Werror', '-o', exe, os.path.join(ROOT, 'tests', 'cabi', 'host_c_abi.c'),
           '-I' + os.path.join(ROOT, 'include'), '-I' + os.path.join(CUDA, 'include'),
           '-L' + os.path.dirname(lib), '-lbbx', '-L' + os.path.join(CUDA, 'lib64'), '-lcudart', '-lm',
           '-Wl,-rpath,' + os.path.dirname(lib), '-Wl,-rpath,' + os.path.join(CUDA, 'lib64')]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


@pytest.mark.skipif(shutil.which('gcc') is None, reason='no gcc')
def test_c_host_compiles_against_the_header(tmp_path):
    """include/bbx.h is valid C (not only C++) and the library links from a C program."""
    exe = _compile(tmp_path)
    assert os.access(exe, os.X_OK)


@pytest.mark.gpu
def test_c_host_runs(tmp_path):
    exe = _compile(tmp_path)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    assert 'C ABI host: all ok' in res.stdout
