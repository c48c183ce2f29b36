"""Build libevosops_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

Safe under torchrun: every rank may call build(); the compile happens under an exclusive file lock, into a temporary
file that is renamed over the library only when nvcc has finished, so no process ever dlopens a half-written file."""
import fcntl
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libevosops_b200.so")
SOURCES = [os.path.join(CSRC, "sops_host.cu")]
DEPS = SOURCES + [os.path.join(CSRC, "sops_kernels.cuh"), os.path.join(CSRC, "sops_rng.cuh"),
                  os.path.join(ROOT, "include", "evosops.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC",
              "-I", os.path.join(ROOT, "include")]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    return "nvcc"


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False):
    """Compile the CUDA kernels + C ABI into evosops_b200/libevosops_b200.so; returns its path."""
    if not (force or stale()):
        return LIB
    with open(os.path.join(HERE, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if force or stale():                              # another process may have built it while we waited
                tmp = LIB + ".tmp.%d" % os.getpid()
                cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + SOURCES
                try:
                    subprocess.check_call(cmd, cwd=ROOT)
                    os.replace(tmp, LIB)
                finally:
                    if os.path.exists(tmp):
                        os.remove(tmp)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB
