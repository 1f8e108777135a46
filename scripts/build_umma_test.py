"""Builds the tcgen05 building-block self-test (scripts/umma_test.cu) as its own development library,
variants/libumma_test.so -- it is deliberately NOT linked into the product libb200world.so."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build_umma_test():
    from idiaptts_b200 import build as B
    out = os.path.join(ROOT, "variants", "libumma_test.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    src = os.path.join(ROOT, "scripts", "umma_test.cu")
    if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        subprocess.run([B._nvcc()] + B.NVCC_FLAGS + ["-I", B.CSRC, "-shared", src, os.path.join(B.CSRC, "api.cu"), "-o", out], check=True)
    return out


if __name__ == "__main__":
    print(build_umma_test())
