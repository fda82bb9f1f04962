"""Build tests/host_harness/libhost_kernels.so: the product's kernel arithmetic compiled for the host (test only)."""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libhost_kernels.so")
SRC = os.path.join(HERE, "host_kernels.cu")
CSRC = os.path.join(os.path.dirname(os.path.dirname(HERE)), "mahakala_b200", "csrc")


def build(force=False):
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-O2", "-std=c++17", "--expt-relaxed-constexpr",
                               "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler",
                               "-fPIC,-mfma,-fopenmp,-ffp-contract=off", "-shared", "-o", SO, SRC, "-ccbin",
                               "/usr/bin/g++", "-lgomp"])
    return SO


def lib():
    L = ctypes.CDLL(build())
    return L
