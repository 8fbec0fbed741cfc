"""Build libchimera_b200.so in-tree with nvcc (-gencode arch=compute_100a,code=sm_100a)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libchimera_b200.so")


def is_stale():
    if not os.path.exists(LIB) or not os.path.exists(os.path.join(HERE, "_npalloc.so")):
        return True
    t = os.path.getmtime(LIB)
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            if f.endswith((".cu", ".cuh", ".h", ".c", "Makefile")) and os.path.getmtime(os.path.join(root, f)) > t:
                return True
    return False


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a (cross-compiles without a GPU)."""
    if not force and not is_stale():
        return LIB
    cmd = ["make", "-C", CSRC, "-j4"] + (["-B"] if force else [])
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout)
    if out.returncode != 0:
        raise RuntimeError("building libchimera_b200.so failed")
    return LIB


if __name__ == "__main__":
    build(force=True, verbose=True)
