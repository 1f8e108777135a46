"""Builds idiaptts_b200/libb200world.so (hand-written CUDA, sm_100a only) in-tree with nvcc.

    python -m idiaptts_b200.build            # incremental: one object per .cu, relink when anything changed
"""
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
LIB = os.path.join(HERE, "libb200world.so")
SOURCES = ["api.cu", "cheaptrick.cu", "cheaptrick_fast.cu", "d4c.cu", "d4c_fast.cu", "mcep.cu", "mgcep.cu", "mcep_tc.cu", "mc2sp_tc.cu", "labels.cu", "synth.cu", "synth_fast.cu", "vtln.cu", "vtln_tc.cu", "mlpg.cu", "metrics.cu", "dio.cu", "corpus_io.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--extended-lambda",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libb200world.so cannot be built")
    return exe


def _digest(paths):
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def build(force=False, verbose=True):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "b200world.h"))
    nvcc = _nvcc()
    jobs = []
    for src in SOURCES:
        sp = os.path.join(CSRC, src)
        obj = os.path.join(OBJ, src[:-3] + ".o")
        stamp = obj + ".sha"
        dig = _digest([sp] + headers)
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
            continue
        jobs.append((sp, obj, stamp, dig))

    def compile_one(job):
        sp, obj, stamp, dig = job
        cmd = [nvcc] + NVCC_FLAGS + ["-c", sp, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (sp, r.stdout, r.stderr))
        with open(stamp, "w") as f:
            f.write(dig)
        return sp

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for done in ex.map(compile_one, jobs):
                if verbose:
                    print("[b200world] compiled", os.path.basename(done), file=sys.stderr)
    if jobs or not os.path.exists(LIB):
        objs = [os.path.join(OBJ, s[:-3] + ".o") for s in SOURCES]
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
        if verbose:
            print("[b200world] linked", LIB, file=sys.stderr)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
