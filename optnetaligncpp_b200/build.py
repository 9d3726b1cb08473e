"""Builds the CUDA engine in-tree: optnetaligncpp_b200/libona_b200.so (sm_100a only)."""
import glob
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
# ONA_LIB_PATH: load a prebuilt variant instead (kernel A/B experiments, profiles/kernel_ab.py)
LIB = os.environ.get("ONA_LIB_PATH") or os.path.join(HERE, "libona_b200.so")
SOURCES = [os.path.join(HERE, "csrc", "ona_engine.cu")]
# every file under csrc/ is a dependency (ona_engine.cu includes all the .cuh headers)
DEPS = sorted(glob.glob(os.path.join(HERE, "csrc", "*"))) + [os.path.join(ROOT, "include", "ona_b200.h")]


def nvcc_path():
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def host_compiler():
    # the image's default g++ wrapper links libstdc++ statically, which clashes with the
    # libstdc++ other Python extensions load; the system compiler links it dynamically
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a: cross-compiles without a GPU."""
    if os.environ.get("ONA_LIB_PATH") or (not force and not stale()):
        return LIB
    cmd = [nvcc_path(), "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a",
           "-lineinfo", "-fmad=false", "-Xcompiler", "-fPIC", "-shared", "-ccbin", host_compiler(),
           "-o", LIB] + SOURCES
    cmd += os.environ.get("ONA_NVCC_FLAGS", "").split()     # tuning experiments only
    if verbose:
        cmd += ["-Xptxas", "-v"]
    out = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(" ".join(cmd))
        print(out.stdout + out.stderr)
    if out.returncode != 0:
        raise RuntimeError("nvcc failed")
    return LIB


if __name__ == "__main__":
    build(force=True, verbose=True)
