"""Builds a kernel-variant copy of the library: python scripts/build_variant.py <name> [-DMACRO=VALUE ...]
-> variants/libb200world_<name>.so (git-ignored, travels with gpurun); select it with B2W_LIB=variants/libb200world_<name>.so."""
import os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from idiaptts_b200 import build as B
name, extra = sys.argv[1], sys.argv[2:]
obj = "/tmp/b2w_variant_%s" % name
os.makedirs(obj, exist_ok=True)
os.makedirs(os.path.join(ROOT, "variants"), exist_ok=True)
def one(src):
    o = os.path.join(obj, src[:-3] + ".o")
    r = subprocess.run([B._nvcc()] + B.NVCC_FLAGS + extra + ["-c", os.path.join(B.CSRC, src), "-o", o], capture_output=True, text=True)
    if r.returncode: raise RuntimeError(r.stderr)
    return o
with ThreadPoolExecutor(8) as ex: objs = list(ex.map(one, B.SOURCES))
out = os.path.join(ROOT, "variants", "libb200world_%s.so" % name)
subprocess.run([B._nvcc(), "-shared", "-o", out] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"], check=True)
print(out)
