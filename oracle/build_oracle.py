"""Build the CPU oracle (TEST INFRASTRUCTURE, not product code).

  oracle/liboracle.so            <- oracle/kmers_oracle.c (our restatement)
  oracle/_ref/kmers_ref.so       <- /root/reference/c/kmers.c, unmodified, -O2
  oracle/_ref/kmers_ref_zinit.so <- same source, `-include oracle/zero_malloc.h`

The two _ref libraries are only (re)built when /root/reference exists (the
build container); on the GPU box the prebuilt files travel with the snapshot.
Reference sources are compiled where they lie and never copied into the repo.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/c/kmers.c"
REF_DIR = os.path.join(HERE, "_ref")


def _stale(out, srcs):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(s) > t for s in srcs if os.path.exists(s))


def _run(cmd):
    subprocess.run(cmd, check=True)


def build(verbose=False):
    out = os.path.join(HERE, "liboracle.so")
    src = os.path.join(HERE, "kmers_oracle.c")
    if _stale(out, [src]):
        _run(["gcc", "-O2", "-std=c99", "-Wall", "-shared", "-fPIC", src, "-o", out])
        if verbose:
            print("built", out)
    if os.path.exists(REF_SRC):
        os.makedirs(REF_DIR, exist_ok=True)
        ref = os.path.join(REF_DIR, "kmers_ref.so")
        if _stale(ref, [REF_SRC]):
            # same flags setuptools used for the reference's own Extension
            _run(["gcc", "-O2", "-shared", "-fPIC", REF_SRC, "-o", ref])
            if verbose:
                print("built", ref)
        zref = os.path.join(REF_DIR, "kmers_ref_zinit.so")
        zh = os.path.join(HERE, "zero_malloc.h")
        if _stale(zref, [REF_SRC, zh]):
            _run(["gcc", "-O2", "-shared", "-fPIC", "-include", zh, REF_SRC, "-o", zref])
            if verbose:
                print("built", zref)
    return out


if __name__ == "__main__":
    build(verbose=True)
    sys.exit(0)
