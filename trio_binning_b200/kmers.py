"""k-mer sets and per-read k-mer counting on a B200.

Host-side mirror of the reference's ctypes boundary
(reference src/trio_binning/kmers.py): same public names, argument meaning and
error behaviour --

    create_kmer_hash_set(path) -> HashSet          (kmers.py:104-122)
    count_kmers_in_read(read, A, B) -> (int, int)  (kmers.py:125-154)
    get_number_kmers_in_set(set) -> int            (kmers.py:157-159)
    kmer_to_int(kmer) -> int                       (kmers.py:80-82)
    reverse_complement(kmer) -> str                (kmers.py:89-93)

-- plus what a GPU needs and the reference lacks: many reads per call
(`count_kmers_in_reads`, `classify_reads`), a double-buffered asynchronous
`Classifier`, device selection and explicit destruction.

Everything is computed by the CUDA library ``kmers_c<EXT_SUFFIX>`` that sits
next to this file (built from csrc/ for sm_100a).  There is no CPU fallback:
if the library is missing and cannot be built, importing this module raises
ImportError, exactly like the reference does when its C library is missing
(kmers.py:36-37); if no GPU is usable, the calls raise RuntimeError.
"""
import ctypes as C
import os
import sys
from importlib.machinery import EXTENSION_SUFFIXES
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))


def _locate_library() -> str:
    """Same search as the reference (kmers.py:30-34): ``kmers_c`` + any of the
    interpreter's extension suffixes, next to this module; last match wins."""
    found = ""
    for suffix in EXTENSION_SUFFIXES:
        candidate = os.path.join(_PKG_DIR, "kmers_c" + suffix)
        if os.path.isfile(candidate):
            found = candidate
    return found


_lib_path = _locate_library()
try:
    # Source checkout: (re)build in-tree when the library is missing or older than csrc/
    # (_build.build() is a no-op when it is current).  A stale library with a missing
    # compiler is still loaded -- the GPU box has no reason to rebuild what travelled with it.
    from . import _build

    if not _lib_path or (_build.needs_build() and _build.have_nvcc()):
        _lib_path = _build.build()
except Exception as exc:  # noqa: BLE001
    if not _lib_path:
        raise ImportError(
            "Cannot load kmers_c library (the CUDA build for sm_100a is missing and "
            f"could not be produced: {exc}). There is no CPU fallback."
        ) from exc
    print(f"trio_binning_b200: rebuilding the stale CUDA library failed ({exc}); loading it as it is",
          file=sys.stderr)

lib = C.cdll.LoadLibrary(_lib_path)


class _HashSet(C.Structure):
    """`hash_set` of include/trio_kmers_b200.h.  The first five fields are the
    reference's struct (c/kmers.c:12-38, kmers.py:53-59); the rest is opaque."""

    _fields_ = [
        ("kmers", C.POINTER(C.c_uint64)),
        ("full", C.POINTER(C.c_ubyte)),
        ("hash_size", C.c_int),
        ("k", C.c_ubyte),
        ("num_kmers", C.c_int),
        ("magic", C.c_uint32),
        ("device", C.c_int),
        ("d_table", C.c_void_p),
        ("n_buckets", C.c_uint32),
        ("has_all_ones", C.c_uint32),
        ("n_distinct", C.c_uint64),
        ("build_ms", C.c_double),
    ]


HashSet = C.POINTER(_HashSet)
_I32P = C.POINTER(C.c_int32)

# ---- legacy symbols (the four the reference binds, kmers.py:62-93) --------------
lib.create_kmer_hash_set.argtypes = [C.c_char_p]
lib.create_kmer_hash_set.restype = HashSet
lib.count_kmers_in_read.argtypes = [C.c_char_p, HashSet, HashSet, C.POINTER(C.c_int), C.POINTER(C.c_int)]
lib.count_kmers_in_read.restype = None
lib.kmer_to_int.argtypes = [C.c_char_p, C.c_ubyte]
lib.kmer_to_int.restype = C.c_uint64
lib.reverse_complement.argtypes = [C.c_char_p, C.c_char_p, C.c_ubyte]
lib.reverse_complement.restype = None
lib.hash_function.argtypes = [C.c_uint64]
lib.hash_function.restype = C.c_uint
lib.kmer_in_hash_set.argtypes = [C.c_char_p, C.c_char_p, HashSet]
lib.kmer_in_hash_set.restype = C.c_char

# ---- B200 API ------------------------------------------------------------------
lib.tb_last_error.restype = C.c_char_p
lib.tb_version.restype = C.c_char_p
lib.tb_device_count.restype = C.c_int
lib.tb_set_device.argtypes = [C.c_int]
lib.tb_get_device.restype = C.c_int
lib.tb_device_info.argtypes = [C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int),
                               C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
lib.tb_host_alloc.argtypes = [C.c_uint64]
lib.tb_host_alloc.restype = C.c_void_p
lib.tb_host_free.argtypes = [C.c_void_p]
lib.tb_host_free.restype = None
lib.tb_set_create_from_file.argtypes = [C.c_char_p, C.c_int]
lib.tb_set_create_from_file.restype = HashSet
lib.tb_set_create_from_keys.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int]
lib.tb_set_create_from_keys.restype = HashSet
lib.tb_set_destroy.argtypes = [HashSet]
lib.tb_set_destroy.restype = None
lib.tb_set_contains_keys.argtypes = [HashSet, C.c_void_p, C.c_uint64, C.c_void_p]
lib.tb_set_load_factor.argtypes = [C.c_double]
lib.tb_scaling_factors.argtypes = [HashSet, HashSet, C.POINTER(C.c_double), C.POINTER(C.c_double)]
lib.tb_scaling_factors.restype = None
lib.tb_classify_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, HashSet, HashSet,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
lib.tb_classifier_create.argtypes = [HashSet, HashSet, C.c_int64, C.c_int64, C.c_int]
lib.tb_classifier_create.restype = C.c_void_p
lib.tb_classifier_destroy.argtypes = [C.c_void_p]
lib.tb_classifier_destroy.restype = None
lib.tb_classifier_submit.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
lib.tb_classifier_wait.argtypes = [C.c_void_p, C.c_int]
lib.tb_classifier_mark.argtypes = [C.c_void_p, C.c_int, C.c_int]
lib.tb_classifier_elapsed.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]
lib.tb_classifier_rerun_resident.argtypes = [C.c_void_p, C.c_int, C.c_int,
                                             C.POINTER(C.c_float), C.POINTER(C.c_float)]
lib.tb_classifier_rerun_resident_flush.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint64,
                                                   C.POINTER(C.c_float), C.POINTER(C.c_float)]
lib.tb_classifier_filter_passes.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint64)]
lib.tb_classifier_sector_runs.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint64)]
lib.tb_classifier_minimizer_mode.argtypes = [C.c_void_p]
lib.tb_classifier_pair_info.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_int)]
lib.tb_classifier_filter_info.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float)]
lib.tb_classifier_stats.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                    C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
lib.tb_gather_bench.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.POINTER(C.c_double), C.POINTER(C.c_float)]
lib.tb_h2d_bench.argtypes = [C.c_uint64, C.c_int, C.POINTER(C.c_double)]


class _PackedBatch(C.Structure):
    """`tb_packed_batch` of include/trio_kmers_b200.h."""

    _fields_ = [
        ("n_reads", C.c_int64), ("total_bases", C.c_int64), ("n_tiles", C.c_int64),
        ("F", C.c_void_p), ("cap_words", C.c_int64),
        ("dirty", C.c_void_p), ("cap_tiles", C.c_int64),
        ("dirty_idx", C.c_void_p), ("e_words", C.c_void_p),
        ("cap_dirty", C.c_int64), ("n_dirty", C.c_int64),
        ("ub_read", C.c_void_p), ("ub_skip", C.c_void_p), ("ub_off", C.c_void_p),
        ("ub_bytes", C.c_void_p),
        ("cap_ub", C.c_int64), ("n_ub", C.c_int64), ("cap_ub_bytes", C.c_int64),
        ("ub_bytes_used", C.c_int64),
    ]


lib.tb_pack_reads.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.POINTER(_PackedBatch)]
lib.tb_classifier_capacity.argtypes = [C.c_void_p] + [C.POINTER(C.c_int64)] * 7
lib.tb_classifier_submit_packed.argtypes = [C.c_void_p, C.c_int, C.POINTER(_PackedBatch), C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]


class StreamStats(C.Structure):
    """`tb_stream_stats` of include/trio_kmers_b200.h."""

    _fields_ = [(n, C.c_uint64) for n in ("n_reads", "n_bases", "n_batches", "input_bytes", "reads_a",
                                          "reads_b", "reads_u", "output_raw_bytes", "output_gz_bytes")] + \
               [(n, C.c_double) for n in ("seconds_total", "seconds_parse", "seconds_gpu_wait", "seconds_write",
                                          "seconds_inflate", "seconds_pack", "seconds_deflate_cpu")] + \
               [("devices_used", C.c_uint64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


lib.tb_classify_fastx.argtypes = [C.c_char_p, HashSet, HashSet, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int,
                                  C.c_int, C.c_int64, C.c_int, C.POINTER(StreamStats)]


class StreamOptions(C.Structure):
    """`tb_stream_options` of include/trio_kmers_b200.h."""

    _fields_ = [("n_devices", C.c_int), ("sets_a", C.POINTER(HashSet)), ("sets_b", C.POINTER(HashSet)),
                ("gzip_output", C.c_int), ("gzip_level", C.c_int), ("fastq_encoder", C.c_int),
                ("round_robin", C.c_int), ("tsv_fd", C.c_int), ("batch_bases", C.c_int64), ("n_threads", C.c_int),
                ("py_text_mode", C.c_int), ("py_buffer_size", C.c_int64)]


lib.tb_classify_fastx_multi.argtypes = [C.c_char_p, C.POINTER(StreamOptions), C.c_char_p, C.c_char_p, C.c_char_p,
                                        C.POINTER(StreamStats)]
lib.tb_set_clone.argtypes = [HashSet, C.c_int]
lib.tb_set_clone.restype = HashSet
lib.tb_route_fastx.argtypes = [C.c_char_p, C.c_char_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                               C.c_int, C.c_int64, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int64, C.c_int,
                               C.POINTER(StreamStats)]
lib.tb_counter_create.argtypes = [C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int]
lib.tb_counter_create.restype = C.c_void_p
lib.tb_fastx_cat.argtypes = [C.c_char_p, C.c_int, C.c_int]
lib.tb_fastx_cat.restype = C.c_int64
lib.tb_counter_slot_bytes.argtypes = [C.c_int]
lib.tb_counter_slot_bytes.restype = C.c_int
lib.tb_counter_destroy.argtypes = [C.c_void_p]
lib.tb_counter_destroy.restype = None
lib.tb_counter_add_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
lib.tb_counter_add_fastx.argtypes = [C.c_void_p, C.c_char_p, C.c_int64, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
lib.tb_counter_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
lib.tb_counter_histogram.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
lib.tb_counter_select_unique.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                         C.c_void_p, C.c_uint64]
lib.tb_counter_select_unique.restype = C.c_int64
lib.tb_float_repr.argtypes = [C.c_double, C.c_char_p, C.c_int]
lib.tb_fastx_dump.argtypes = [C.c_char_p, C.c_int]
lib.tb_fastx_dump.restype = C.c_int64
lib.tb_gunzip_fd.argtypes = [C.c_char_p, C.c_int, C.c_int64]
lib.tb_gunzip_fd.restype = C.c_int64
lib.tb_crc32.argtypes = [C.c_uint32, C.c_char_p, C.c_int64]
lib.tb_crc32.restype = C.c_uint32
lib.tb_crc32_accelerated.restype = C.c_int
lib.tb_fastq_gzip.argtypes = [C.c_char_p, C.c_int64, C.c_char_p, C.c_int64, C.c_int]
lib.tb_fastq_gzip.restype = C.c_int64
lib.tb_gzip_copy.argtypes = [C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int64]
lib.tb_gzip_copy.restype = C.c_int64


def _error() -> str:
    return (lib.tb_last_error() or b"unknown error").decode("utf-8", "replace")


def _check(rc: int) -> None:
    if rc != 0:
        raise RuntimeError(_error())


# =============================================================================
# the reference's public functions
# =============================================================================

def kmer_to_int(kmer: str) -> int:
    """Convert a k-mer to its integer form: first base in the two least
    significant bits, A=0 C=1 G=2 T=3, anything else 0 (c/kmers.c:50-72)."""
    raw = kmer.encode("utf-8")
    return int(lib.kmer_to_int(raw, C.c_ubyte(len(kmer))))


def reverse_complement(kmer: str) -> str:
    """Reverse-complement a k-mer; non-ACGT positions come back as 'x' because the
    native routine does not write them (c/kmers.c:78-91, kmers.py:91)."""
    raw = kmer.encode("utf-8")
    out = C.create_string_buffer(b"x" * len(kmer), len(kmer) + 1)
    lib.reverse_complement(raw, out, len(kmer))
    return out.raw[: len(kmer)].decode("utf-8")


def create_kmer_hash_set(kmer_file_path: str) -> "HashSet":
    """Read a list of k-mers (one per line) into a GPU-resident hash set.

    Mirrors kmers.py:104-122: IOError if the path is not a file, a progress line
    on stderr, then the native call (here: mmap + sm_100a parse/insert kernel).
    """
    if not os.path.isfile(kmer_file_path):
        raise IOError(f"Specified file {kmer_file_path} does not exist or is not file.")
    print(f"Reading k-mers in {kmer_file_path}...", file=sys.stderr)
    handle = lib.create_kmer_hash_set(os.fsencode(kmer_file_path))
    if not handle:
        raise RuntimeError(f"create_kmer_hash_set({kmer_file_path}): {_error()}")
    return handle


def count_kmers_in_read(read: str, kmers_hap_a: "HashSet", kmers_hap_b: "HashSet") -> Tuple[int, int]:
    """Count the k-mers of one read found in set A and, failing that, in set B
    (kmers.py:125-154).  One read = one batch of one through the CUDA kernels;
    use `count_kmers_in_reads` for throughput."""
    count_a, count_b = C.c_int(), C.c_int()
    lib.count_kmers_in_read(read.encode("utf-8"), kmers_hap_a, kmers_hap_b,
                            C.byref(count_a), C.byref(count_b))
    if count_a.value < 0:
        raise RuntimeError(f"count_kmers_in_read: {_error()}")
    return count_a.value, count_b.value


def get_number_kmers_in_set(kmer_hash_set: "HashSet") -> int:
    """Number of k-mers (= lines of the file) in a set (kmers.py:157-159)."""
    return kmer_hash_set.contents.num_kmers


# =============================================================================
# B200 additions
# =============================================================================

def version() -> str:
    return lib.tb_version().decode()


def device_count() -> int:
    return int(lib.tb_device_count())


def set_device(device: int) -> None:
    """Select the GPU used by subsequently created sets (one process per GPU)."""
    _check(lib.tb_set_device(device))


def device_info(device: int = 0) -> dict:
    name = C.create_string_buffer(256)
    sms, hbm, l2 = C.c_int(), C.c_uint64(), C.c_uint64()
    _check(lib.tb_device_info(device, name, 256, C.byref(sms), C.byref(hbm), C.byref(l2)))
    return {"name": name.value.decode(), "sm_count": sms.value, "hbm_bytes": hbm.value,
            "l2_bytes": l2.value}


def set_load_factor(load_factor: float) -> None:
    _check(lib.tb_set_load_factor(load_factor))


def create_kmer_hash_set_from_ints(keys, k: int, device: Optional[int] = None) -> "HashSet":
    """Build a set from RAW integer encodings (`kmer_to_int` values); num_kmers = len(keys)."""
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    dev = lib.tb_get_device() if device is None else device
    handle = lib.tb_set_create_from_keys(keys.ctypes.data, len(keys), k, dev)
    if not handle:
        raise RuntimeError(f"create_kmer_hash_set_from_ints: {_error()}")
    return handle


def destroy_kmer_hash_set(kmer_hash_set: "HashSet") -> None:
    """Free the device table (the reference never frees its sets)."""
    lib.tb_set_destroy(kmer_hash_set)


def set_contains_ints(kmer_hash_set: "HashSet", keys) -> np.ndarray:
    """Membership of already-canonicalised lookup keys (testing aid)."""
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    out = np.zeros(len(keys), dtype=np.uint8)
    _check(lib.tb_set_contains_keys(kmer_hash_set, keys.ctypes.data, len(keys), out.ctypes.data))
    return out.astype(bool)


def scaling_factors(kmers_hap_a: "HashSet", kmers_hap_b: "HashSet") -> Tuple[float, float]:
    sa, sb = C.c_double(), C.c_double()
    lib.tb_scaling_factors(kmers_hap_a, kmers_hap_b, C.byref(sa), C.byref(sb))
    return sa.value, sb.value


def pack_reads(reads: Sequence[str]) -> Tuple[np.ndarray, np.ndarray]:
    """Concatenate reads into (uint8 bases, int64 offsets), the batch input format.

    Each read is UTF-8 encoded and cut at its first NUL byte, which is what the
    reference's `read.encode("utf-8")` + C `strlen` see (kmers.py:147, c/kmers.c:285).
    """
    try:
        blob = "".join(reads).encode("ascii")
        lengths = np.fromiter((len(r) for r in reads), dtype=np.int64, count=len(reads))
        if b"\x00" in blob:
            raise UnicodeError
    except UnicodeError:
        parts = [r.encode("utf-8").split(b"\x00", 1)[0] for r in reads]
        blob = b"".join(parts)
        lengths = np.fromiter((len(p) for p in parts), dtype=np.int64, count=len(parts))
    offsets = np.zeros(len(reads) + 1, dtype=np.int64)
    np.cumsum(lengths, out=offsets[1:])
    bases = np.frombuffer(blob, dtype=np.uint8)
    return bases, offsets


def classify_packed(bases: np.ndarray, offsets: np.ndarray, kmers_hap_a: "HashSet",
                    kmers_hap_b: "HashSet", want_scores: bool = True):
    """One batch: returns (count_a, count_b, score_a, score_b, bins) as numpy arrays
    (scores/bins None unless requested).  bins holds ord('A'|'B'|'U')."""
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    n = len(offsets) - 1
    ca = np.zeros(n, dtype=np.int32)
    cb = np.zeros(n, dtype=np.int32)
    sa = np.zeros(n, dtype=np.float64) if want_scores else None
    sb = np.zeros(n, dtype=np.float64) if want_scores else None
    bins = np.zeros(n, dtype=np.uint8) if want_scores else None
    if n:
        _check(lib.tb_classify_batch(
            bases.ctypes.data, offsets.ctypes.data, n, kmers_hap_a, kmers_hap_b,
            ca.ctypes.data, cb.ctypes.data,
            sa.ctypes.data if want_scores else None, sb.ctypes.data if want_scores else None,
            bins.ctypes.data if want_scores else None))
    return ca, cb, sa, sb, bins


def count_kmers_in_reads(reads: Sequence[str], kmers_hap_a: "HashSet",
                         kmers_hap_b: "HashSet") -> List[Tuple[int, int]]:
    """Batched `count_kmers_in_read`: one kernel pipeline for all reads."""
    bases, offsets = pack_reads(reads)
    ca, cb, _, _, _ = classify_packed(bases, offsets, kmers_hap_a, kmers_hap_b, want_scores=False)
    return list(zip(ca.tolist(), cb.tolist()))


def classify_reads(reads: Sequence[str], kmers_hap_a: "HashSet", kmers_hap_b: "HashSet"):
    """Counts, size-normalised scores and the A/B/U call per read, all from the
    device (classify_by_kmers.py:104-115).  Returns a list of
    (count_a, count_b, score_a, score_b, bin) with bin in 'A','B','U'."""
    bases, offsets = pack_reads(reads)
    ca, cb, sa, sb, bins = classify_packed(bases, offsets, kmers_hap_a, kmers_hap_b)
    letters = bins.tobytes().decode("ascii")
    return list(zip(ca.tolist(), cb.tolist(), sa.tolist(), sb.tolist(), letters))


class PinnedBuffer:
    """Page-locked host memory exposed as a numpy array (for asynchronous copies)."""

    def __init__(self, n_items: int, dtype):
        self.dtype = np.dtype(dtype)
        self.nbytes = max(1, int(n_items) * self.dtype.itemsize)
        self.ptr = lib.tb_host_alloc(self.nbytes)
        if not self.ptr:
            raise RuntimeError(f"pinned allocation failed: {_error()}")
        raw = (C.c_ubyte * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(raw, dtype=self.dtype, count=int(n_items))

    def free(self):
        if self.ptr:
            self.array = None
            lib.tb_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:  # noqa: BLE001
            pass


class PackedBatch:
    """Host-side 2-bit packed batch in pinned memory (tb_packed_batch) plus its packer.

    `pack(bases, offsets, k)` runs the multi-threaded AVX-512/AVX2 host packer (host_pack.cpp);
    the arrays are then ready for `Classifier.submit_packed`."""

    def __init__(self, max_bases: int, max_reads: int, pinned: bool = True):
        tiles = (int(max_bases) + 1023) // 1024
        self.cap = dict(words=(tiles + 1) * 32, tiles=tiles + 1,
                        dirty=(tiles + 1) * 32, ub=int(max_reads) + 1,
                        ub_bytes=max(1 << 20, int(max_bases) // 16))
        alloc = (lambda n, dt: PinnedBuffer(n, dt)) if pinned else (lambda n, dt: _Plain(n, dt))
        self.F = alloc(self.cap["words"], np.uint64)
        self.dirty = alloc(self.cap["tiles"], np.uint32)
        self.dirty_idx = alloc(self.cap["dirty"], np.uint32)
        self.e_words = alloc(self.cap["dirty"], np.uint64)
        self.ub_read = alloc(self.cap["ub"], np.uint32)
        self.ub_skip = alloc(self.cap["ub"], np.uint32)
        self.ub_off = alloc(self.cap["ub"] + 1, np.int64)
        self.ub_bytes = alloc(self.cap["ub_bytes"], np.uint8)
        self.offsets = alloc(int(max_reads) + 1, np.int64)
        self.c = _PackedBatch()
        self.c.F, self.c.cap_words = self.F.ptr, self.cap["words"]
        self.c.dirty, self.c.cap_tiles = self.dirty.ptr, self.cap["tiles"]
        self.c.dirty_idx, self.c.e_words, self.c.cap_dirty = self.dirty_idx.ptr, self.e_words.ptr, self.cap["dirty"]
        self.c.ub_read, self.c.ub_skip, self.c.ub_off = self.ub_read.ptr, self.ub_skip.ptr, self.ub_off.ptr
        self.c.ub_bytes, self.c.cap_ub, self.c.cap_ub_bytes = self.ub_bytes.ptr, self.cap["ub"], self.cap["ub_bytes"]

    def pack(self, bases: np.ndarray, offsets: np.ndarray, k: int, threads: int = 0) -> None:
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        n = len(offsets) - 1
        self.offsets.array[: n + 1] = offsets
        rc = lib.tb_pack_reads(bases.ctypes.data, self.offsets.ptr, n, k, threads, C.byref(self.c))
        if rc != 0:
            raise RuntimeError(f"tb_pack_reads failed with code {rc}")

    @property
    def n_reads(self) -> int:
        return int(self.c.n_reads)

    @property
    def h2d_bytes(self) -> int:
        c = self.c
        return int((c.n_tiles + 1) * 32 * 8 + (c.n_tiles + 1) * 4 + (c.n_reads + 1) * 8 + c.n_dirty * 12
                   + c.n_ub * 16 + c.ub_bytes_used)


class _Plain:
    """Pageable stand-in for PinnedBuffer (lets the packer be tested without a GPU)."""

    def __init__(self, n_items, dtype):
        self.array = np.zeros(max(1, int(n_items)), dtype=dtype)
        self.ptr = self.array.ctypes.data

    def free(self):
        pass


class Classifier:
    """Double-buffered asynchronous classifier (tb_classifier_* of the C ABI).

    Each slot owns a CUDA stream, device buffers and pinned host staging/result
    buffers.  `submit(slot, bases, offsets)` copies the batch into the slot's
    pinned buffers and enqueues H2D + kernels + D2H; `result(slot)` waits and
    returns views of the slot's pinned result arrays (valid until the slot is
    submitted again)."""

    def __init__(self, kmers_hap_a, kmers_hap_b, max_bases: int, max_reads: int, n_slots: int = 2):
        self.max_bases, self.max_reads, self.n_slots = int(max_bases), int(max_reads), int(n_slots)
        self._sets = (kmers_hap_a, kmers_hap_b)  # keep alive
        self._h = lib.tb_classifier_create(kmers_hap_a, kmers_hap_b, self.max_bases, self.max_reads, n_slots)
        if not self._h:
            raise RuntimeError(f"Classifier: {_error()}")
        self._slots = []
        for _ in range(n_slots):
            self._slots.append({
                "bases": PinnedBuffer(self.max_bases, np.uint8),
                "offsets": PinnedBuffer(self.max_reads + 1, np.int64),
                "ca": PinnedBuffer(self.max_reads, np.int32),
                "cb": PinnedBuffer(self.max_reads, np.int32),
                "sa": PinnedBuffer(self.max_reads, np.float64),
                "sb": PinnedBuffer(self.max_reads, np.float64),
                "bins": PinnedBuffer(self.max_reads, np.uint8),
                "n": 0,
            })

    def staging(self, slot: int):
        """The slot's pinned (bases, offsets) arrays, to be filled in place."""
        s = self._slots[slot]
        return s["bases"].array, s["offsets"].array

    def submit(self, slot: int, bases: Optional[np.ndarray], offsets: Optional[np.ndarray],
               n_reads: Optional[int] = None, want_scores: bool = True) -> None:
        """Enqueue a batch.  With bases/offsets given they are first copied into the
        pinned staging arrays; with None the staging arrays are used as filled."""
        s = self._slots[slot]
        if offsets is not None:
            n_reads = len(offsets) - 1
            total = int(offsets[-1]) if n_reads else 0
            if n_reads > self.max_reads or total > self.max_bases:
                raise ValueError("batch exceeds classifier capacity")
            s["offsets"].array[: n_reads + 1] = offsets
            s["bases"].array[:total] = bases[:total]
        s["n"] = int(n_reads)
        s["scores"] = want_scores
        _check(lib.tb_classifier_submit(
            self._h, slot, s["bases"].ptr, s["offsets"].ptr, s["n"], s["ca"].ptr, s["cb"].ptr,
            s["sa"].ptr if want_scores else None, s["sb"].ptr if want_scores else None,
            s["bins"].ptr if want_scores else None))

    def submit_packed(self, slot: int, batch: "PackedBatch", want_scores: bool = True) -> None:
        """Enqueue a host-packed batch (PackedBatch.pack) on a slot."""
        s = self._slots[slot]
        s["n"] = batch.n_reads
        s["scores"] = want_scores
        _check(lib.tb_classifier_submit_packed(
            self._h, slot, C.byref(batch.c), batch.offsets.ptr, s["ca"].ptr, s["cb"].ptr,
            s["sa"].ptr if want_scores else None, s["sb"].ptr if want_scores else None,
            s["bins"].ptr if want_scores else None))

    def wait(self, slot: int) -> None:
        _check(lib.tb_classifier_wait(self._h, slot))

    def result(self, slot: int):
        self.wait(slot)
        s, n = self._slots[slot], self._slots[slot]["n"]
        if s.get("scores", True):
            return s["ca"].array[:n], s["cb"].array[:n], s["sa"].array[:n], s["sb"].array[:n], s["bins"].array[:n]
        return s["ca"].array[:n], s["cb"].array[:n], None, None, None

    def mark(self, slot: int, which: int) -> None:
        """Record timing event `which` (0/1) at the current tail of the slot's stream."""
        _check(lib.tb_classifier_mark(self._h, slot, which))

    def elapsed_ms(self, slot_from: int, which_from: int, slot_to: int, which_to: int) -> float:
        ms = C.c_float()
        _check(lib.tb_classifier_elapsed(self._h, slot_from, which_from, slot_to, which_to, C.byref(ms)))
        return ms.value

    def rerun_resident(self, slot: int, iters: int, flush_bytes: int = 0) -> Tuple[float, float]:
        """Re-run the device part of the slot's last batch `iters` times with no
        copies; returns (total ms, ms spent in the classify kernel).  With
        flush_bytes > 0 a scratch buffer of that size is overwritten between steps."""
        total, cls = C.c_float(), C.c_float()
        _check(lib.tb_classifier_rerun_resident_flush(self._h, slot, iters, flush_bytes,
                                                      C.byref(total), C.byref(cls)))
        return total.value, cls.value

    def filter_passes(self, slot: int) -> int:
        n = C.c_uint64()
        _check(lib.tb_classifier_filter_passes(self._h, slot, C.byref(n)))
        return n.value

    def sector_runs(self, slot: int) -> int:
        """Bitmap sector fetches of the slot's last batch with per-lane re-use (roofline aid)."""
        n = C.c_uint64()
        _check(lib.tb_classifier_sector_runs(self._h, slot, C.byref(n)))
        return n.value

    def filter_info(self) -> dict:
        """log2(bits) of the L2-resident pre-filter (0 = direct kernel) and its build time."""
        m, ms = C.c_int(), C.c_float()
        _check(lib.tb_classifier_filter_info(self._h, C.byref(m), C.byref(ms)))
        nb, nd, bits = C.c_uint32(), C.c_uint64(), C.c_int()
        _check(lib.tb_classifier_pair_info(self._h, C.byref(nb), C.byref(nd), C.byref(bits)))
        return {"log2_bits": m.value, "build_ms": ms.value, "pair_buckets": nb.value,
                "pair_distinct": nd.value, "bits_per_key": bits.value if nb.value else 1,
                "minimizer": bool(lib.tb_classifier_minimizer_mode(self._h)),
                "two_level": lib.tb_classifier_minimizer_mode(self._h) == 2}

    def stats(self, slot: int) -> dict:
        w, p, f, launches = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        _check(lib.tb_classifier_stats(self._h, slot, C.byref(w), C.byref(p), C.byref(f), C.byref(launches)))
        return {"windows": w.value, "probes": p.value, "fixup_reads": f.value,
                "kernel_launches": launches.value}

    def close(self) -> None:
        if self._h:
            lib.tb_classifier_destroy(self._h)
            self._h = None
            for s in self._slots:
                for key in ("bases", "offsets", "ca", "cb", "sa", "sb", "bins"):
                    s[key].free()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


def clone_kmer_hash_set(kmer_hash_set: "HashSet", device: int) -> "HashSet":
    """Replica of a set on another GPU (device-to-device copy of the table)."""
    handle = lib.tb_set_clone(kmer_hash_set, device)
    if not handle:
        raise RuntimeError(f"clone_kmer_hash_set: {_error()}")
    return handle


_PY_TEXT_IO = None


def python_text_io_behaviour(directory: str = ".") -> Tuple[int, int]:
    """How THIS interpreter's open(name, "w") buffers, for the native plain-text writer
    (tb_stream_options.py_text_mode / py_buffer_size): the reference's --no-gzip-output files are
    written through two such objects on one file (seq.py:127-130), so their flush points decide
    the bytes.  Probed, not assumed: the rule changed between CPython releases."""
    global _PY_TEXT_IO
    import io
    import tempfile

    if _PY_TEXT_IO is None:
        class _Log(io.BufferedIOBase):
            def __init__(self):
                self.sizes = []

            def writable(self):
                return True

            def write(self, b):
                self.sizes.append(len(b))
                return len(b)

        def calls(sizes):
            log = _Log()
            text = io.TextIOWrapper(log, encoding="utf-8")
            for n in sizes:
                text.write("x" * n)
            text.flush()
            return log.sizes

        if calls([100, 9000]) == [9100]:
            mode = 2                      # never flushes the pending list first
        elif calls([5000, 5000]) == [5000, 5000]:
            mode = 1                      # flushes when list + string would exceed the chunk
        else:
            mode = 0                      # flushes when the string alone is a chunk or more
        assert calls([100, 9000]) == ([9100] if mode == 2 else [100, 9000])
        _PY_TEXT_IO = mode
    # BufferedWriter size (io.open derives it from the file system's block size, differently in
    # different releases): fill a binary file in 512-byte steps; the first write that does not
    # fit makes the buffer flush, and the file is then exactly one buffer long
    buffer_size = 0
    try:
        with tempfile.TemporaryDirectory(dir=directory) as tmp:
            path = os.path.join(tmp, "probe")
            with open(path, "wb") as f:
                piece = b"x" * 512
                for _ in range((16 << 20) // 512 + 2):
                    f.write(piece)
                    buffer_size = os.fstat(f.fileno()).st_size
                    if buffer_size:
                        break
    except OSError:
        buffer_size = 0  # cannot probe there: the writer falls back to the file's st_blksize
    return _PY_TEXT_IO, buffer_size


def classify_fastx_file(reads_path: str, kmers_hap_a, kmers_hap_b, out_a: str, out_b: str, out_u: str,
                        tsv_fd: int = -1, gzip_level: int = 9, batch_bases: int = 256 << 20, threads: int = 0,
                        gzip_output: bool = True, fastq_encoder: bool = True, round_robin: bool = False) -> dict:
    """Native streaming pipeline (tb_classify_fastx_multi): parse, pack to 2 bits, classify on
    the GPU(s), write the three outputs and the TSV, everything in input order.

    kmers_hap_a / kmers_hap_b: one set each, or equally long lists of replicas, one per GPU to use
    (`clone_kmer_hash_set`).  gzip_output=False is the reference's --no-gzip-output mode: plain
    text, hapB's records written through a second handle on hapA's file (seq.py:127-130).
    Returns the pipeline's counters and stage timings."""
    sets_a = list(kmers_hap_a) if isinstance(kmers_hap_a, (list, tuple)) else [kmers_hap_a]
    sets_b = list(kmers_hap_b) if isinstance(kmers_hap_b, (list, tuple)) else [kmers_hap_b]
    if len(sets_a) != len(sets_b) or not sets_a:
        raise ValueError("need one replica of each set per device")
    opt = StreamOptions()
    opt.n_devices = len(sets_a)
    opt.sets_a = (HashSet * len(sets_a))(*sets_a)
    opt.sets_b = (HashSet * len(sets_b))(*sets_b)
    opt.gzip_output, opt.gzip_level, opt.fastq_encoder = int(gzip_output), gzip_level, int(fastq_encoder)
    opt.round_robin, opt.tsv_fd, opt.batch_bases, opt.n_threads = int(round_robin), tsv_fd, batch_bases, threads
    if not gzip_output:
        opt.py_text_mode, opt.py_buffer_size = python_text_io_behaviour(os.path.dirname(os.path.abspath(out_a)))
    stats = StreamStats()
    rc = lib.tb_classify_fastx_multi(os.fsencode(reads_path), C.byref(opt), os.fsencode(out_a), os.fsencode(out_b),
                                     os.fsencode(out_u), C.byref(stats))
    if rc != 0:
        raise RuntimeError(f"tb_classify_fastx failed ({rc}): {_error()}")
    return stats.as_dict()


def route_fastx_file(reads_path: str, bins: str, out_a: str, out_b: str, out_u: str, tsv_fd: int = -1,
                     gzip_output: bool = True, gzip_level: int = 6, fastq_encoder: bool = True, workers: int = 1,
                     round_robin: bool = False, batch_bases: int = 1 << 20, threads: int = 4) -> dict:
    """The host side of the native pipeline without a GPU (testing aid): record i is assigned
    bins[i % len(bins)]; reader, batches, worker threads, in-order router, writers and TSV are
    the product's."""
    stats = StreamStats()
    pattern = bins.encode("ascii")
    mode, bufsize = (0, 0) if gzip_output else python_text_io_behaviour(os.path.dirname(os.path.abspath(out_a)))
    rc = lib.tb_route_fastx(os.fsencode(reads_path), pattern, len(pattern), int(gzip_output), gzip_level,
                            int(fastq_encoder), workers, int(round_robin), mode, bufsize, os.fsencode(out_a), os.fsencode(out_b),
                            os.fsencode(out_u), tsv_fd, batch_bases, threads, C.byref(stats))
    if rc != 0:
        raise RuntimeError(f"tb_route_fastx failed ({rc}): {_error()}")
    return stats.as_dict()


class KmerCounter:
    """Canonical k-mer counting table on the GPU (tb_counter_* of the C ABI): what the reference's
    find-unique-kmers gets from `kmc` (find_unique_kmers.py:62-103).  Keys are MSB-first 2-bit
    encodings of the canonical k-mer (the lexicographically smaller strand)."""

    COUNTER_MAX = 255  # KMC -cs default
    DB_MIN = 2         # KMC -ci default: k-mers seen once are not in a database

    def __init__(self, k: int, n_slots: int, part: int = 0, n_parts: int = 1, device: Optional[int] = None):
        self.k = k
        dev = lib.tb_get_device() if device is None else device
        self._h = lib.tb_counter_create(k, int(n_slots), part, n_parts, dev)
        if not self._h:
            raise RuntimeError(f"KmerCounter: {_error()}")

    def add_reads(self, reads: Sequence[str]) -> None:
        bases, offsets = pack_reads(reads)
        self.add_packed(bases, offsets)

    def add_packed(self, bases: np.ndarray, offsets: np.ndarray) -> None:
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        _check(lib.tb_counter_add_batch(self._h, bases.ctypes.data, offsets.ctypes.data, len(offsets) - 1))

    def add_file(self, path: str, batch_bases: int = 64 << 20, threads: int = 0) -> Tuple[int, int]:
        """Count every record of a FASTA/FASTQ(.gz) file (native reader); returns (reads, bases)."""
        n_reads, n_bases = C.c_uint64(), C.c_uint64()
        rc = lib.tb_counter_add_fastx(self._h, os.fsencode(path), batch_bases, threads, C.byref(n_reads), C.byref(n_bases))
        if rc != 0:
            raise RuntimeError(f"counting k-mers of {path} failed ({rc}): {_error()}")
        return n_reads.value, n_bases.value

    def stats(self) -> dict:
        """Waits for the counting to finish; raises if the table filled up."""
        w, d, s = C.c_uint64(), C.c_uint64(), C.c_uint64()
        rc = lib.tb_counter_stats(self._h, C.byref(w), C.byref(d), C.byref(s))
        out = {"windows": w.value, "distinct": d.value, "slots": s.value}
        if rc != 0:
            raise OverflowError(_error())
        return out

    def histogram(self, n_bins: int = 256) -> np.ndarray:
        """hist[i] = distinct k-mers seen i times; the last bin holds everything >= n_bins-1."""
        hist = np.zeros(n_bins, dtype=np.uint64)
        _check(lib.tb_counter_histogram(self._h, hist.ctypes.data, n_bins))
        return hist

    def select_unique(self, other: Optional["KmerCounter"], min_count: int, max_count: int) -> np.ndarray:
        """Sorted keys of this database that are absent from `other`'s and whose saturated count is in
        [min_count, max_count] (kmers_subtract + kmc_dump -ci -cx, find_unique_kmers.py:171-233)."""
        oh = other._h if other is not None else None
        n = lib.tb_counter_select_unique(self._h, oh, min_count, max_count, self.DB_MIN, self.COUNTER_MAX, None, 0)
        if n < 0:
            raise RuntimeError(_error())
        out = np.zeros(int(n), dtype=np.uint64)
        if n:
            m = lib.tb_counter_select_unique(self._h, oh, min_count, max_count, self.DB_MIN, self.COUNTER_MAX,
                                             out.ctypes.data, int(n))
            if m != n:
                raise RuntimeError(f"selection changed between two passes ({n} then {m}): {_error()}")
        out.sort()
        return out

    def close(self) -> None:
        if self._h:
            lib.tb_counter_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


def float_repr(x: float) -> str:
    """The native TSV writer's rendering of a float (must equal Python's repr)."""
    buf = C.create_string_buffer(64)
    n = lib.tb_float_repr(x, buf, 64)
    if n < 0:
        raise RuntimeError("tb_float_repr: buffer too small")
    return buf.value.decode()


def fastx_dump(reads_path: str):
    """Records of a FASTA/FASTQ(.gz) file as parsed by the native reader: a list of
    (name, seq, qual-or-None) byte strings (testing aid, no GPU needed)."""
    import tempfile

    with tempfile.TemporaryFile() as tmp:
        n = lib.tb_fastx_dump(os.fsencode(reads_path), tmp.fileno())
        if n < 0:
            raise RuntimeError(f"tb_fastx_dump failed ({n}): {_error()}")
        tmp.seek(0)
        data = tmp.read()
    out, pos = [], 0
    while pos < len(data):
        fields = []
        for _ in range(3):  # length-prefixed: int64 little-endian, -1 = None
            length = int.from_bytes(data[pos:pos + 8], "little", signed=True)
            pos += 8
            fields.append(None if length < 0 else data[pos:pos + length])
            pos += max(length, 0)
        out.append(tuple(fields))
    assert len(out) == n
    return out


def fastx_cat(reads_path: str, decode_threads: int = 4) -> bytes:
    """The text of a FASTA/FASTQ(.gz) file as the native reader's input stage delivers it to the
    parser (testing aid, no GPU needed)."""
    import tempfile

    with tempfile.TemporaryFile() as tmp:
        n = lib.tb_fastx_cat(os.fsencode(reads_path), tmp.fileno(), decode_threads)
        if n < 0:
            raise RuntimeError(f"tb_fastx_cat failed ({n}): {_error()}")
        tmp.seek(0)
        data = tmp.read()
    assert len(data) == n
    return data


def gunzip(path: str, chunk: int = 0) -> bytes:
    """Contents of a .gz file as decoded by the native reader's own streaming inflate, read
    `chunk` bytes at a time (testing aid, no GPU needed).  Raises ValueError for input that is
    corrupt, truncated, not gzip at all, or followed by bytes that are neither zero padding nor
    another member -- the cases in which Python's gzip module raises for the reference."""
    import tempfile

    with tempfile.TemporaryFile() as tmp:
        n = lib.tb_gunzip_fd(os.fsencode(path), tmp.fileno(), chunk)
        if n == -2:
            raise ValueError(f"{path}: corrupt or truncated gzip data")
        if n < 0:
            raise OSError(f"tb_gunzip_fd({path}) failed ({n})")
        tmp.seek(0)
        data = tmp.read()
    assert len(data) == n
    return data


def fastq_gzip(text: bytes, verify: bool = True) -> Optional[bytes]:
    """One gzip member for FASTA/FASTQ text by the writer's FASTQ-aware encoder, or None where
    it declines (short lines, tiny input) and the writer would use zlib (testing aid)."""
    cap = len(text) * 2 + 65536
    buf = C.create_string_buffer(cap)
    n = lib.tb_fastq_gzip(text, len(text), buf, cap, 1 if verify else 0)
    if n == -1:
        return None
    if n < 0:
        raise RuntimeError(f"tb_fastq_gzip failed ({n})")
    return buf.raw[:n]


def gzip_copy(in_path: str, out_path: str, level: int = 6, threads: int = 2, fastq_encoder: bool = False,
              piece: int = 1 << 16, bgzf: bool = False) -> int:
    """Compress a file with the native pipeline's output writer: 4 MiB chunks as independent gzip
    members, or BGZF blocks with bgzf=True (testing / benchmarking aid, no GPU needed)."""
    with open(in_path, "rb") as f:
        n = lib.tb_gzip_copy(f.fileno(), os.fsencode(out_path), level, threads,
                             2 if bgzf else (1 if fastq_encoder else 0), piece)
    if n < 0:
        raise RuntimeError(f"tb_gzip_copy failed ({n})")
    return int(n)


def crc32(data: bytes, crc: int = 0) -> int:
    """gzip CRC-32 as computed by the native reader (zlib.crc32 semantics; testing aid)."""
    return int(lib.tb_crc32(crc & 0xFFFFFFFF, data, len(data)))


def crc32_accelerated() -> bool:
    return bool(lib.tb_crc32_accelerated())


def h2d_bench(n_bytes: int = 256 << 20, iters: int = 100) -> float:
    """GB/s of plain pinned-host -> device copies on the current device (tb_h2d_bench)."""
    rate = C.c_double()
    _check(lib.tb_h2d_bench(n_bytes, iters, C.byref(rate)))
    return rate.value


def gather_bench(table_bytes: int, loads_in_flight: int = 8, threads_per_block: int = 256,
                 blocks_per_sm: int = 8, steps: int = 64, variant: int = 0) -> Tuple[float, float]:
    """Random-32-B-sector gather ceiling on this GPU: (sectors per second, ms)."""
    rate, ms = C.c_double(), C.c_float()
    _check(lib.tb_gather_bench(table_bytes, loads_in_flight, threads_per_block, blocks_per_sm,
                               steps, variant, C.byref(rate), C.byref(ms)))
    return rate.value, ms.value
