"""In-tree build of the CUDA library (nvcc cross-compiles for sm_100a without a GPU)."""
import os
import subprocess

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB = os.path.join(CSRC, "libsynthpy_b200.so")


def build(force=False, verbose=False):
    if force and os.path.exists(LIB):
        os.remove(LIB)
    out = subprocess.run(["make", "-C", CSRC, "libsynthpy_b200.so"], capture_output=True, text=True)
    if verbose or out.returncode:
        print(out.stdout, out.stderr)
    if out.returncode:
        raise RuntimeError("nvcc build of libsynthpy_b200.so failed")
    return LIB
