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


@pytest.mark.parametrize("threads", [1, 3])
@pytest.mark.parametrize("misalign", [0, 8])
def test_pack_clean_runs_and_block_edges(kmers, threads, misalign):
    """Long clean stretches take the 256-base block loop; junk placed on, before and after block
    and tile edges (and in the last, partial block) must stop it at the right word and hand over
    to the word path, and the loop must resume afterwards.  F is placed at a 64-byte aligned
    address (streaming stores, as with pinned buffers) or 8 bytes past one (plain stores)."""
    rng = np.random.default_rng(77 + threads)
    lens = [300_000, 0, 1, 255, 256, 257, 70_000, 33]
    total = sum(lens)
    bases = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, total)].copy()
    junk_at = [0, 255, 256, 257, 511, 1023, 1024, 1025, 8191, 8192, 100_000, 100_031, 100_032,
               299_999, 300_000, 300_600, total - 300, total - 33, total - 1]
    junk_at += [int(x) for x in rng.integers(0, total, 12)]
    bases[junk_at] = np.frombuffer(b"Nn\x80\xff@Q", dtype=np.uint8)[rng.integers(0, 6, len(junk_at))]
    offsets = np.zeros(len(lens) + 1, dtype=np.int64)
    offsets[1:] = np.cumsum(lens)
    pb = kmers.PackedBatch(total + 4096, len(lens) + 2, pinned=False)
    big = np.zeros(pb.cap["words"] + 16, dtype=np.uint64)
    off = ((misalign - big.ctypes.data) % 64) // 8
    view = big[off:off + pb.cap["words"]]
    assert view.ctypes.data % 64 == misalign
    pb.F.array, pb.F.ptr = view, view.ctypes.data
    pb.c.F = view.ctypes.data
    pb.pack(bases, offsets, 21, threads=threads)
    n_words = (total + 31) // 32
    padded = np.zeros(n_words * 32, dtype=np.uint8)
    padded[:total] = orc_mod._CODE[bases]
    want = (padded.reshape(n_words, 32).astype(np.uint64) << (2 * np.arange(32, dtype=np.uint64))).sum(
        axis=1, dtype=np.uint64)
    assert (pb.F.array[:n_words] == want).all()
    vpad = np.ones(n_words * 32, dtype=bool)
    vpad[:total] = orc_mod._VALID[bases]
    dirty_want = ~vpad.reshape(n_words, 32).all(axis=1)
    bits = np.unpackbits(pb.dirty.array[: (n_words + 31) // 32].view(np.uint8), bitorder="little")[:n_words]
    assert (bits.astype(bool) == dirty_want).all()
    assert pb.c.n_dirty == int(dirty_want.sum())
    assert (pb.dirty_idx.array[: pb.c.n_dirty] == np.nonzero(dirty_want)[0]).all()


@pytest.mark.filterwarnings("ignore::DeprecationWarning")
def test_pack_pool_concurrent_callers_and_fork(kmers):
    """tb_pack_reads keeps persistent worker threads: two callers at once must both get correct
    results (calls serialise), and a forked child (which inherits no workers) must still pack."""
    import os
    import threading

    rng = np.random.default_rng(5)
    n, L = 40, 20_000
    offsets = np.arange(n + 1, dtype=np.int64) * L
    batches = [np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n * L)].copy() for _ in range(2)]

    def want(b):
        return (orc_mod._CODE[b].reshape(-1, 32).astype(np.uint64) << (2 * np.arange(32, dtype=np.uint64))).sum(
            axis=1, dtype=np.uint64)

    pbs = [kmers.PackedBatch(n * L, n, pinned=False) for _ in range(2)]
    errs = []

    def run(i):
        for _ in range(20):
            pbs[i].pack(batches[i], offsets, 21, threads=4)
            if not (pbs[i].F.array[: n * L // 32] == want(batches[i])).all():
                errs.append(i)

    ts = [threading.Thread(target=run, args=(i,)) for i in range(2)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs
    pid = os.fork()
    if pid == 0:
        ok = 1
        try:
            pbs[0].pack(batches[1], offsets, 21, threads=3)
            ok = 0 if (pbs[0].F.array[: n * L // 32] == want(batches[1])).all() else 2
        finally:
            os._exit(ok)
    _, status = os.waitpid(pid, 0)
    assert os.WIFEXITED(status) and os.WEXITSTATUS(status) == 0
