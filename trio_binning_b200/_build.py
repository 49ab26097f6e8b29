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
SOURCES = ["capi.cu", "capi_count.cu", "set_build.cu", "classify.cu", "kmer_count.cu", "gather_bench.cu", "host_pack.cpp",
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


def have_nvcc():
    try:
        find_nvcc()
        return True
    except RuntimeError:
        return False


def source_digest():
    """Hash of everything the library is built from.  Kept beside the library so that staleness
    does not depend on file times (a snapshot copied to another machine scrambles them)."""
    import hashlib

    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    with open(os.path.abspath(__file__), "rb") as fh:
        h.update(fh.read())
    h.update(os.environ.get("TB_NVCC_EXTRA", "").encode())
    return h.hexdigest()


DIGEST_PATH = os.path.join(PKG, "kmers_c.build_digest")


def needs_build():
    if not os.path.exists(LIB_PATH) or not os.path.exists(DIGEST_PATH):
        return True
    with open(DIGEST_PATH) as f:
        return f.read().strip() != source_digest()


def _compile_one(nvcc, src, obj, flags, verbose):
    cmd = [nvcc] + flags + ["-c", src, "-o", obj]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)


def build(force=False, verbose=False, extra_flags=()):
    """Compile every source to build/<name>.o (only those older than their inputs, in parallel)
    and link the shared library.  One object per translation unit keeps the edit-compile loop
    of a single kernel file at seconds."""
    if not force and not needs_build():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor

    nvcc = find_nvcc()
    extra_env = os.environ.get("TB_NVCC_EXTRA", "").split()  # e.g. -DTB_MINI_MIN_BLOCKS=6 (tuning experiments)
    cflags = [f for f in NVCC_FLAGS if f not in ("-shared", "-lz")]
    for drop in ("-cudart", "static"):
        cflags.remove(drop)
    cflags += list(extra_flags) + extra_env
    obj_dir = os.path.join(PKG, "build")
    os.makedirs(obj_dir, exist_ok=True)
    stamp = os.path.join(obj_dir, "flags.txt")
    flag_text = " ".join(cflags)
    if not os.path.exists(stamp) or open(stamp).read() != flag_text:
        force = True
    header_time = max(os.path.getmtime(os.path.join(CSRC, h)) for h in HEADERS)
    header_time = max(header_time, os.path.getmtime(os.path.abspath(__file__)))
    jobs, objs = [], []
    for name in SOURCES:
        src = os.path.join(CSRC, name)
        obj = os.path.join(obj_dir, os.path.splitext(name)[0] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), header_time):
            jobs.append((src, obj))
    with ThreadPoolExecutor(max(1, min(len(jobs), os.cpu_count() or 1))) as pool:
        list(pool.map(lambda j: _compile_one(nvcc, j[0], j[1], cflags, verbose), jobs))
    with open(stamp, "w") as f:
        f.write(flag_text)
    link = [nvcc, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a"] + objs + ["-lz", "-lpthread", "-o", LIB_PATH]
    if verbose:
        print(" ".join(link), file=sys.stderr)
    subprocess.run(link, check=True)
    with open(DIGEST_PATH, "w") as f:
        f.write(source_digest() + "\n")
    return LIB_PATH


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True,
          extra_flags=["-Xptxas", "-v"] if "--ptxas-v" in sys.argv else [])
    print(LIB_PATH)
