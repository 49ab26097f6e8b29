"""The host-side 2-bit packer (csrc/host_pack.cpp, tb_pack_reads) against the numpy
restatement of the stream formulation in oracle/oracle.py.  Runs without a GPU: the packer
is host code (format conversion only, no k-mer lookups)."""
import numpy as np
import pytest

from oracle import oracle as orc_mod


@pytest.fixture(scope="module")
def kmers():
    from trio_binning_b200 import kmers as mod

    return mod


def _unpack(words, n):
    out = np.zeros(n, dtype=np.uint8)
    for i in range(n):
        out[i] = (int(words[i // 32]) >> (2 * (i % 32))) & 3
    return out


def _random_batch(rng, n_reads, max_len, bad_rate, k):
    reads = []
    junk = b"NnacgtRY\r\x80\xff-"
    for r in range(n_reads):
        L = int(rng.integers(0, max_len))
        s = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, L)].copy()
        bad = np.nonzero(rng.random(L) < bad_rate)[0]
        s[bad] = np.frombuffer(junk, dtype=np.uint8)[rng.integers(0, len(junk), len(bad))]
        if L and r % 5 == 0:
            s[: int(rng.integers(1, min(L, 3 * k) + 1))] = ord("N")
        reads.append(s.tobytes())
    offsets = np.zeros(n_reads + 1, dtype=np.int64)
    offsets[1:] = np.cumsum([len(r) for r in reads])
    return reads, np.frombuffer(b"".join(reads), dtype=np.uint8), offsets


@pytest.mark.parametrize("k,threads", [(21, 1), (21, 4), (32, 3), (5, 2)])
def test_pack_matches_stream_restatement(kmers, k, threads):
    rng = np.random.default_rng(k + threads)
    reads, bases, offsets = _random_batch(rng, 120, 700, 0.02, k)
    total = int(offsets[-1])
    pb = kmers.PackedBatch(total + 5000, len(reads) + 5, pinned=False)
    pb.pack(bases, offsets, k, threads=threads)
    assert pb.c.total_bases == total and pb.c.n_tiles == (total + 1023) // 1024
    # F: forward codes, non-ACGT -> 0
    want_f = orc_mod._CODE[bases]
    assert (_unpack(pb.F.array, total) == want_f).all()
    # dirty bitmap and E words: flat fill-forward stream (may look back across reads; only
    # positions inside windows that are not left to the fix-up are ever consumed)
    valid = orc_mod._VALID[bases]
    n_words = (total + 31) // 32
    dirty_want = np.zeros(n_words, dtype=bool)
    for w in range(n_words):
        dirty_want[w] = not valid[w * 32:(w + 1) * 32].all()
    dirty_got = np.zeros(n_words, dtype=bool)
    for w in range(n_words):
        dirty_got[w] = (int(pb.dirty.array[w // 32]) >> (w % 32)) & 1
    assert (dirty_got == dirty_want).all()
    idx = np.where(valid, np.arange(total), -1)
    last = np.maximum.accumulate(idx)
    e_flat = np.where(last >= 0, want_f[np.maximum(last, 0)], 0)
    nd = int(pb.c.n_dirty)
    assert nd == int(dirty_want.sum())
    for i in range(nd):
        w = int(pb.dirty_idx.array[i])
        assert dirty_want[w]
        n = min(32, total - w * 32)
        got = _unpack(pb.e_words.array[i:i + 1], n)
        assert (got == e_flat[w * 32:w * 32 + n]).all()
    # Q3-UB list: skip = min(max_j U[j], windows) for reads with a non-ACGT byte in the first k
    want = {}
    for r, s in enumerate(reads):
        W = len(s) - k + 1
        if W <= 0 or all(orc_mod._VALID[b] for b in s[:k]):
            continue
        U = orc_mod.never_written(orc_mod._VALID[np.frombuffer(s, dtype=np.uint8)], k)
        skip = min(int(U.max()), W)
        if skip:
            want[r] = (skip, s[: skip + k - 1])
    got = {}
    for i in range(int(pb.c.n_ub)):
        r, skip, off = int(pb.ub_read.array[i]), int(pb.ub_skip.array[i]), int(pb.ub_off.array[i])
        got[r] = (skip, pb.ub_bytes.array[off:off + skip + k - 1].tobytes())
    assert got == want


def test_pack_empty_and_capacity(kmers):
    pb = kmers.PackedBatch(4096, 8, pinned=False)
    pb.pack(np.zeros(0, dtype=np.uint8), np.zeros(1, dtype=np.int64), 21)
    assert pb.c.n_reads == 0 and pb.c.total_bases == 0
    big = np.full(10000, ord("A"), dtype=np.uint8)
    with pytest.raises(RuntimeError):
        pb.pack(big, np.array([0, 10000], dtype=np.int64), 21)
