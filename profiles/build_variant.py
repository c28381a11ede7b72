#!/usr/bin/env python
"""Instrumented / experimental build of ONE source next to the product library:
   python profiles/build_variant.py <name> <source.cu> <nvcc flags...>
-> vognet_pytorch_b200/libvog_b200_<name>.so (the other objects are the product build's); run a script against it with
VOG_B200_SO=<path>.  Built here (nvcc cross-compiles), travels to the GPU box like the product .so."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vognet_pytorch_b200 import _lib  # noqa: E402

name, src, extra = sys.argv[1], sys.argv[2], sys.argv[3:]
_lib.build()
objdir = os.path.join(_lib.CSRC, '_obj')
cflags = [f for f in _lib.NVCC_FLAGS if f != '-shared']
obj = os.path.join(objdir, f'{src}.{name}.o')
subprocess.run(['nvcc'] + cflags + extra + ['-c', '-o', obj, src], cwd=_lib.CSRC, check=True)
objs = [os.path.join(objdir, s + '.o') for s in _lib.SOURCES if s != src] + [obj]
out = os.path.join(os.path.dirname(_lib.SO_PATH), f'libvog_b200_{name}.so')
subprocess.run(['nvcc', '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', out] + objs, check=True)
os.remove(obj)
print(out)
