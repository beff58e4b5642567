"""Build of the sm_100a shared library libxnb_hotpath.so (C-ABI in include/xnb_hotpath.h), in-tree.

nvcc cross-compiles without a GPU.  The product has no CPU fallback: if the library is missing, `exanbody_b200.capi`
raises instead of computing anything."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libxnb_hotpath.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-ccbin", "/usr/bin/g++", "-Xcompiler", "-fPIC", "-shared", "-ldl"]


def sources():
    return [os.path.join(SRC, f) for f in sorted(os.listdir(SRC))] + [os.path.join(os.path.dirname(HERE), "include", "xnb_hotpath.h")]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(f) > t for f in sources())


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = [NVCC] + FLAGS + ["-o", LIB, os.path.join(SRC, "xnb_hotpath.cu"), os.path.join(SRC, "xnb_host_inputs.cpp"),
                                           os.path.join(SRC, "xnb_host_decomp.cpp")]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return LIB


HOST = os.path.join(HERE, "host")
DECK = os.path.join(OUT_DIR, "lj_deck")
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def build_host(force=False):
    """C++17 operator mirror (exanbody_b200/host) + the LJ deck program, linked against libxnb_hotpath.so"""
    lib = build()
    src = [os.path.join(HOST, f) for f in ("xnb_operators.cpp", "lj_deck.cpp")]
    dep = src + [os.path.join(HOST, "xnb_operators.hpp"), lib]
    if not force and os.path.exists(DECK) and all(os.path.getmtime(f) <= os.path.getmtime(DECK) for f in dep):
        return DECK
    cmd = [CXX, "-std=c++17", "-O2", "-Wall", "-o", DECK] + src + ["-L" + OUT_DIR, "-lxnb_hotpath", "-Wl,-rpath,$ORIGIN"]
    subprocess.check_call(cmd)
    return DECK


if __name__ == "__main__":
    print(build(force=True, verbose=True))
