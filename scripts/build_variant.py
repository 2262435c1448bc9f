"""Builds a variant of libls1b200.so with extra -D flags for ONE source file (kernel sweeps on the GPU box: LS1B200_LIB=<path>).

    python scripts/build_variant.py <name> <source.cu> -DLS1_PF_LCAP=54 -DLS1_PF_SOFT=36
 -> ls1_mardyn_b200/variants/libls1b200_<name>.so
"""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ls1_mardyn_b200 import build as B

name, src, defs = sys.argv[1], sys.argv[2], sys.argv[3:]
B.build()
vdir = os.path.join(B.HERE, "variants"); os.makedirs(vdir, exist_ok=True)
obj = os.path.join(vdir, "%s_%s.o" % (src[:-3], name))
p = subprocess.run([B.NVCC, *B.FLAGS, *defs, "-c", os.path.join(B.CSRC, src), "-o", obj], capture_output=True, text=True)
open(obj[:-2] + ".ptxas.log", "w").write(p.stdout + p.stderr)
if p.returncode: sys.exit(p.stdout + p.stderr)
objs = [obj if s == src else os.path.join(B.OBJ, s[:-3] + ".o") for s in B.SOURCES]
lib = os.path.join(vdir, "libls1b200_%s.so" % name)
p = subprocess.run([B.NVCC, "-shared", "-o", lib, *objs, "-lcudart"], capture_output=True, text=True)
if p.returncode: sys.exit(p.stdout + p.stderr)
print(lib)
