"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes access to
  * ``liboracle.so``            our C restatement (oracle/kmers_oracle.c)
  * ``_ref/kmers_ref.so``       the UNMODIFIED reference c/kmers.c, compiled here
  * ``_ref/kmers_ref_zinit.so`` the same source with zero-initialised malloc
and a numpy "two-stream" restatement (`streams_count`) of the same arithmetic
in the shape the CUDA kernel uses (flat 2-bit F / E streams + never-written
masks), so that formulation is validated on the CPU before it is trusted on
the GPU.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  Parity status: pinned (see the header
of kmers_oracle.c).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class _RefSet(C.Structure):
    # reference struct, c/kmers.c:12-38 (mirrored by kmers.py:53-59)
    _fields_ = [
        ("kmers", C.POINTER(C.c_uint64)),
        ("full", C.POINTER(C.c_ubyte)),
        ("hash_size", C.c_int),
        ("k", C.c_ubyte),
        ("num_kmers", C.c_int),
    ]


def _ensure_built():
    from . import build_oracle  # noqa: WPS433  (local import: build on demand)

    build_oracle.build()


class Oracle:
    """Our C restatement."""

    def __init__(self):
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            _ensure_built()
        lib = C.CDLL(path)
        P = C.POINTER(_RefSet)
        lib.orc_kmer_to_int.argtypes = [C.c_char_p, C.c_ubyte]
        lib.orc_kmer_to_int.restype = C.c_uint64
        lib.orc_reverse_complement.argtypes = [C.c_char_p, C.c_char_p, C.c_ubyte]
        lib.orc_hash_function.argtypes = [C.c_uint64]
        lib.orc_hash_function.restype = C.c_uint
        lib.orc_set_from_file.argtypes = [C.c_char_p]
        lib.orc_set_from_file.restype = P
        lib.orc_set_from_ints.argtypes = [C.c_void_p, C.c_uint64, C.c_int]
        lib.orc_set_from_ints.restype = P
        lib.orc_set_from_sorted.argtypes = [C.c_void_p, C.c_uint64, C.c_int]
        lib.orc_set_from_sorted.restype = P
        lib.orc_set_free.argtypes = [P]
        lib.orc_set_contains_int.argtypes = [P, C.c_uint64]
        lib.orc_set_contains_int.restype = C.c_int
        lib.orc_count_kmers_in_read.argtypes = [C.c_char_p, P, P, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.orc_count_kmers_in_read_n.argtypes = [C.c_char_p, C.c_long, P, P,
                                                  C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.orc_count_kmers_in_reads.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, P, P,
                                                 C.c_void_p, C.c_void_p]
        lib.orc_scaling_factors.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        lib.orc_decide.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_double,
                                   C.c_void_p, C.c_void_p, C.c_void_p]
        self.lib = lib

    # -- scalar helpers
    def kmer_to_int(self, kmer):
        b = kmer if isinstance(kmer, bytes) else kmer.encode()
        return int(self.lib.orc_kmer_to_int(b, len(b)))

    def reverse_complement(self, kmer):
        b = kmer if isinstance(kmer, bytes) else kmer.encode()
        out = C.create_string_buffer(b"x" * len(b), len(b))
        self.lib.orc_reverse_complement(b, out, len(b))
        return out.raw.decode("latin-1")

    def hash_function(self, x):
        return int(self.lib.orc_hash_function(C.c_uint64(x)))

    # -- sets
    def set_from_file(self, path):
        s = self.lib.orc_set_from_file(os.fsencode(path))
        if not s:
            raise IOError(path)
        return s

    def set_from_ints(self, keys, k):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        return self.lib.orc_set_from_ints(keys.ctypes.data, len(keys), k)

    def set_from_sorted(self, keys, k):
        """Set over an ASCENDING uint64 array (membership by binary search, no table build):
        for the full-size configs, where the reference's table would take minutes to fill.
        The array is kept alive by this Oracle object."""
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        s = self.lib.orc_set_from_sorted(keys.ctypes.data, len(keys), k)
        self._keep = getattr(self, "_keep", {})
        self._keep[C.addressof(s.contents)] = keys
        return s

    def set_free(self, s):
        getattr(self, "_keep", {}).pop(C.addressof(s.contents), None)
        self.lib.orc_set_free(s)

    def contains(self, s, key):
        return bool(self.lib.orc_set_contains_int(s, C.c_uint64(key)))

    # -- counting
    def count(self, read, A, B):
        b = read if isinstance(read, bytes) else read.encode("utf-8")
        ca, cb = C.c_int(), C.c_int()
        self.lib.orc_count_kmers_in_read_n(b, len(b), A, B, C.byref(ca), C.byref(cb))
        return ca.value, cb.value

    def count_batch(self, bases, offsets, A, B):
        """bases: uint8 array of concatenated reads, offsets: int64[R+1]."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(offsets) - 1
        ca = np.zeros(n, dtype=np.int32)
        cb = np.zeros(n, dtype=np.int32)
        self.lib.orc_count_kmers_in_reads(bases.ctypes.data, offsets.ctypes.data, n, A, B,
                                          ca.ctypes.data, cb.ctypes.data)
        return ca, cb

    def count_batch_threads(self, bases, offsets, A, B, threads=None):
        """count_batch over `threads` host threads (ctypes drops the GIL; sets are read-only)."""
        from concurrent.futures import ThreadPoolExecutor

        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(offsets) - 1
        threads = max(1, min(threads or (os.cpu_count() or 1), n))
        ca = np.zeros(n, dtype=np.int32)
        cb = np.zeros(n, dtype=np.int32)
        # slices of about equal numbers of bases
        cuts = np.searchsorted(offsets, np.linspace(0, offsets[-1], threads * 4 + 1)[1:-1])
        bounds = np.unique(np.concatenate([[0], cuts, [n]]))

        def work(i):
            lo, hi = int(bounds[i]), int(bounds[i + 1])
            self.lib.orc_count_kmers_in_reads(bases.ctypes.data, offsets[lo:].ctypes.data, hi - lo, A, B,
                                              ca[lo:].ctypes.data, cb[lo:].ctypes.data)

        with ThreadPoolExecutor(threads) as pool:
            list(pool.map(work, range(len(bounds) - 1)))
        return ca, cb

    def scaling_factors(self, na, nb):
        sa, sb = C.c_double(), C.c_double()
        self.lib.orc_scaling_factors(na, nb, C.byref(sa), C.byref(sb))
        return sa.value, sb.value

    def decide(self, ca, cb, sa, sb):
        ca = np.ascontiguousarray(ca, dtype=np.int32)
        cb = np.ascontiguousarray(cb, dtype=np.int32)
        n = len(ca)
        a = np.zeros(n, dtype=np.float64)
        b = np.zeros(n, dtype=np.float64)
        bins = np.zeros(n, dtype=np.uint8)
        self.lib.orc_decide(ca.ctypes.data, cb.ctypes.data, n, sa, sb,
                            a.ctypes.data, b.ctypes.data, bins.ctypes.data)
        return a, b, bins


class CompiledReference:
    """The reference's own c/kmers.c, compiled by oracle/build_oracle.py.

    ``zinit=True`` loads the zero-initialised-malloc build (Q3-UB contract)."""

    def __init__(self, zinit=False):
        name = "kmers_ref_zinit.so" if zinit else "kmers_ref.so"
        path = os.path.join(HERE, "_ref", name)
        if not os.path.exists(path):
            _ensure_built()
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        lib = C.CDLL(path)
        P = C.POINTER(_RefSet)
        # the four symbols the reference's kmers.py binds (kmers.py:62-93)
        lib.create_kmer_hash_set.argtypes = [C.c_char_p]
        lib.create_kmer_hash_set.restype = P
        lib.count_kmers_in_read.argtypes = [C.c_char_p, P, P, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.kmer_to_int.argtypes = [C.c_char_p, C.c_ubyte]
        lib.kmer_to_int.restype = C.c_uint64
        lib.reverse_complement.argtypes = [C.c_char_p, C.c_char_p, C.c_ubyte]
        lib.hash_function.argtypes = [C.c_uint64]
        lib.hash_function.restype = C.c_uint
        self.lib = lib

    @staticmethod
    def available(zinit=False):
        name = "kmers_ref_zinit.so" if zinit else "kmers_ref.so"
        return os.path.exists(os.path.join(HERE, "_ref", name)) or os.path.exists("/root/reference/c/kmers.c")

    def kmer_to_int(self, kmer):
        b = kmer if isinstance(kmer, bytes) else kmer.encode()
        return int(self.lib.kmer_to_int(b, len(b)))

    def reverse_complement(self, kmer):
        b = kmer if isinstance(kmer, bytes) else kmer.encode()
        out = C.create_string_buffer(b"x" * len(b), len(b) + 1)
        self.lib.reverse_complement(b, out, len(b))
        return out.raw[: len(b)].decode("latin-1")

    def hash_function(self, x):
        return int(self.lib.hash_function(C.c_uint64(x)))

    def set_from_file(self, path):
        return self.lib.create_kmer_hash_set(os.fsencode(path))

    def count(self, read, A, B):
        b = read if isinstance(read, bytes) else read.encode("utf-8")
        ca, cb = C.c_int(), C.c_int()
        self.lib.count_kmers_in_read(b, A, B, C.byref(ca), C.byref(cb))
        return ca.value, cb.value


# --------------------------------------------------------------------------
# numpy two-stream restatement (the formulation the CUDA kernels implement)
# --------------------------------------------------------------------------

_CODE = np.zeros(256, dtype=np.uint8)
_VALID = np.zeros(256, dtype=bool)
for _i, _c in enumerate(b"ACGT"):
    _CODE[_c] = _i
    _VALID[_c] = True


def streams(read_bytes):
    """F, E, valid for one read.

    F[x] = code(read[x]) with non-ACGT -> 0.
    E[x] = code of the nearest valid base at or before x (0 if none).
    """
    r = np.frombuffer(read_bytes, dtype=np.uint8)
    F = _CODE[r]
    valid = _VALID[r]
    idx = np.where(valid, np.arange(len(r)), -1)
    last = np.maximum.accumulate(idx) if len(r) else idx
    E = np.where(last >= 0, F[np.maximum(last, 0)], 0).astype(np.uint8)
    return F, E, valid


def never_written(valid, k):
    """U[j] for j<k: number of leading windows whose reverse-complement digit for
    forward slot j has never been written (Q3-UB): first valid position >= j,
    minus j (len-j if there is none)."""
    n = len(valid)
    U = np.zeros(k, dtype=np.int64)
    for j in range(min(k, n)):
        p = j
        while p < n and not valid[p]:
            p += 1
        U[j] = p - j
    return U


def window_keys(read_bytes, k):
    """Canonical lookup key of every window of one read, via the two streams.

    fwd(i) = sum_j F[i+j] << 2j
    rc(i)  = sum_j (3 - E[i+j]) << 2(k-1-j), digit forced to 0 while i < U[j]
    """
    F, E, valid = streams(read_bytes)
    n = len(F)
    W = max(0, n - k + 1)
    if W == 0:
        return np.zeros(0, dtype=np.uint64)
    U = never_written(valid, k)
    fwd = np.zeros(W, dtype=np.uint64)
    rc = np.zeros(W, dtype=np.uint64)
    win = np.arange(W)
    for j in range(k):
        fwd |= F[j:j + W].astype(np.uint64) << np.uint64(2 * j)
        digit = (3 - E[j:j + W]).astype(np.uint64)
        digit[win < U[j]] = 0
        rc |= digit << np.uint64(2 * (k - 1 - j))
    return np.minimum(fwd, rc)


def streams_count(read_bytes, keys_a, keys_b, k):
    """(count_A, count_B) with Python sets of RAW stored keys (Q1) and the
    A-then-B rule (Q4)."""
    ca = cb = 0
    for key in window_keys(read_bytes, k).tolist():
        if key in keys_a:
            ca += 1
        elif key in keys_b:
            cb += 1
    return ca, cb
