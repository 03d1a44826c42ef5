#!/usr/bin/env python
"""tuning builds: tools/build_variant.py <name> [-DMACRO=VALUE ...] compiles the whole library with extra defines into
featuredetection_b200/csrc/variants/libfdb200_<name>.so (selected at run time with FDB_LIB=<path>; *.so files are git-ignored
but travel to the GPU box)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g

name, defs = sys.argv[1], sys.argv[2:]
out = os.path.join(g.CSRC, "variants")
obj = os.path.join(out, "_" + name)
os.makedirs(obj, exist_ok=True)
flags = [f for f in g.NVCC_FLAGS if f != "-shared"] + defs
nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
jobs = [[nvcc] + flags + ["-c", "-o", os.path.join(obj, s + ".o"), os.path.join(g.CSRC, s)] for s in g.SOURCES]
with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
    list(ex.map(lambda c: subprocess.check_call(c, cwd=g.CSRC, stderr=subprocess.DEVNULL), jobs))
lib = os.path.join(out, "libfdb200_%s.so" % name)
subprocess.check_call([nvcc, "-shared", "-o", lib] + [os.path.join(obj, s + ".o") for s in g.SOURCES] + ["-lz"], cwd=g.CSRC)
print(lib)
