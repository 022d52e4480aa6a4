#!/usr/bin/env python
"""Builds an experimental variant of the library from a copy of csrc/ (A/B runs on the GPU box):

    python tools/build_variant.py NAME CSRC_DIR   ->  variants/lib_NAME.so   (git-ignored, travels with gpurun)
    DGN_LIB_PATH=variants/lib_NAME.so python bench.py ...
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from decagon_b200.build import FLAGS, NVCC  # noqa: E402

name, src = sys.argv[1], sys.argv[2]
out_dir = os.path.join(ROOT, 'variants')
obj_dir = os.path.join('/tmp', 'variant_' + name)
os.makedirs(out_dir, exist_ok=True)
os.makedirs(obj_dir, exist_ok=True)
procs, objs = [], []
for f in sorted(os.listdir(src)):
    if f.endswith('.cu'):
        obj = os.path.join(obj_dir, f[:-3] + '.o')
        procs.append(subprocess.Popen([NVCC] + FLAGS + ['-I', os.path.join(ROOT, 'decagon_b200', 'csrc'), '-c', os.path.join(src, f), '-o', obj]))
        objs.append(obj)
if any(p.wait() != 0 for p in procs):
    sys.exit('nvcc failed')
lib = os.path.join(out_dir, 'lib_%s.so' % name)
subprocess.check_call([NVCC, '-shared', '-o', lib] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a'])
print(lib)
