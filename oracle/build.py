"""Compiles oracle/c/world_oracle.c (gcc) into oracle/_build/<cpu-key>/liboracle.so.

The library is built with -march=native, so it is keyed by the CPU's feature flags: a copy that travelled from another
machine is never loaded on a CPU it was not built for.  Test infrastructure / CPU baseline only."""
import hashlib
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "c", "world_oracle.c")


def _cpu_key():
    flags = ""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    flags = line
                    break
    except OSError:
        pass
    with open(SRC, "rb") as f:
        src = f.read()
    return hashlib.sha256(flags.encode() + src).hexdigest()[:16]


def lib_path():
    return os.path.join(HERE, "_build", _cpu_key(), "liboracle.so")


def build(verbose=False):
    out = lib_path()
    if os.path.exists(out):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cmd = ["gcc", "-O3", "-march=native", "-ffp-contract=off", "-fno-math-errno", "-shared", "-fPIC", "-o", out, SRC, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("gcc failed:\n" + r.stderr)
    if verbose:
        print("[oracle] built", out)
    return out


if __name__ == "__main__":
    print(build(verbose=True))
