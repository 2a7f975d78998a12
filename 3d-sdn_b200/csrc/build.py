#!/usr/bin/env python
"""nvcc build of libsdn_b200.so for sm_100a, in-tree (3d-sdn_b200/sdn_b200/libsdn_b200.so)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(os.path.dirname(HERE), 'sdn_b200', 'libsdn_b200.so')
SOURCES = [os.path.join(HERE, 'sdn_raster.cu'), os.path.join(HERE, 'sdn_linear_tc.cu'),
           os.path.join(HERE, 'sdn_objects.cu'), os.path.join(HERE, 'sdn_composite.cu'),
           os.path.join(HERE, 'sdn_conv_tc.cu')]
# -fmad=false: visibility arithmetic must not be contracted (DESIGN.md "arithmetic contract")
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-fmad=false', '-std=c++17',
         '-Xcompiler', '-fPIC', '-shared', '-I' + os.path.join(ROOT, 'include')]


def build(verbose=True, force=False):
    deps = SOURCES + [os.path.join(ROOT, 'include', 'sdn_b200.h'), os.path.abspath(__file__)]
    if not force and os.path.isfile(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    cmd = ['nvcc'] + FLAGS + ['-o', OUT] + SOURCES
    if verbose:
        print('[build]', ' '.join(cmd))
    subprocess.check_call(cmd)
    return OUT


if __name__ == '__main__':
    build(force='--force' in sys.argv)
