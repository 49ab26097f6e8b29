"""Build recipe for the CUDA shared library (in-tree, sm_100a only).

The output is ``trio_binning_b200/kmers_c<EXT_SUFFIX>`` -- the very file name
the reference's loader scans for (reference src/trio_binning/kmers.py:30-38) --
a plain shared object without ``PyInit_`` that is only ever dlopen'ed by ctypes.
"""
import os
import shutil
import subprocess
import sys
from importlib.machinery import EXTENSION_SUFFIXES

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
SOURCES = ["capi.cu", "set_build.cu", "classify.cu", "gather_bench.cu", "host_pack.cpp",
           "fastx_pipeline.cpp", "gz_stream.cpp", "fastq_deflate.cpp"]
HEADERS = ["kmer_device.cuh", "launchers.h", "gz_stream.h", "fastq_deflate.h", os.path.join("..", "..", "include", "trio_kmers_b200.h")]
LIB_NAME = "kmers_c" + EXTENSION_SUFFIXES[0]
LIB_PATH = os.path.join(PKG, LIB_NAME)

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC,-O3,-Wall,-pthread",
    "-shared", "-cudart", "static", "-lz",
]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the CUDA library cannot be built (no CPU fallback exists)")


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, extra_flags=()):
    if not force and not needs_build():
        return LIB_PATH
    extra_env = os.environ.get("TB_NVCC_EXTRA", "").split()  # e.g. -DTB_MINI_MIN_BLOCKS=6 (tuning experiments)
    cmd = [find_nvcc()] + NVCC_FLAGS + list(extra_flags) + extra_env + [os.path.join(CSRC, f) for f in SOURCES]
    cmd += ["-o", LIB_PATH]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return LIB_PATH


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True,
          extra_flags=["-Xptxas", "-v"] if "--ptxas-v" in sys.argv else [])
    print(LIB_PATH)
