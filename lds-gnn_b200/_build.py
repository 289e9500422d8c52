"""Builds liblds_b200.so in-tree with one nvcc invocation (sm_100a only, -lineinfo for ncu source pages)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "liblds_b200.so")
SOURCES = ["lds_common.cu", "lds_k1_sample.cu", "lds_k2_propagate.cu", "lds_k3_theta_update.cu", "lds_fused_small.cu", "lds_outer_step.cu", "lds_spmm.cu", "lds_skinny.cu", "lds_theta0.cu", "lds_k1_packed.cu", "lds_k2_packed.cu", "lds_adam.cu", "lds_rowops.cu"]


def nvcc_path():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "lds_b200.h")]
    return any(os.path.getmtime(d) > built for d in deps if os.path.isfile(d))


def build(force=False, verbose=False):
    """Compile every CUDA source of the package into lib/liblds_b200.so. Returns the library path."""
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--threads", "0",
           "-Xcompiler", "-fPIC", "-shared", "-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout + proc.stderr)
    return LIB_PATH
