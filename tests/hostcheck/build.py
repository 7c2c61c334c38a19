"""Builds the test-only host instantiation of the device routines (see hostcheck.cu)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libccx_hostcheck.so")
ROOT = os.path.dirname(os.path.dirname(HERE))


def build():
    srcs = [os.path.join(HERE, "hostcheck.cu")] + [
        os.path.join(ROOT, "chinesecheckersagent_b200", "csrc", f) for f in ("ccx_device.cuh",)]
    if os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in srcs):
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    subprocess.check_call([nvcc, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
                           "-shared", "-cudart", "none", "-o", LIB, srcs[0]], env=env)
    return LIB
