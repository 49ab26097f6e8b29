"""The writer's FASTQ-aware gzip encoder (csrc/fastq_deflate.cpp): its members must be standard
gzip -- Python's gzip module, i.e. what any user of the output files has, must give back the
text byte for byte -- and for long-read text they should be smaller than zlib's.  CPU only."""
import gzip
import zlib

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st


@pytest.fixture(scope="module")
def kmers():
    from trio_binning_b200 import kmers as mod

    return mod


def _fastq(n, length, seed, qual="random"):
    rng = np.random.default_rng(seed)
    parts = []
    for i in range(n):
        L = int(length * (0.5 + rng.random()))
        s = rng.choice(list(b"ACGT"), L).astype(np.uint8).tobytes()
        if qual == "random":
            q = rng.integers(33, 74, L, dtype=np.uint8).tobytes()
        elif qual == "runs":  # HiFi-like: saturated quality with short dips
            a = np.full(L, 126, dtype=np.uint8)
            for p, w in zip(rng.integers(0, L, L // 200 + 1), rng.integers(1, 30, L // 200 + 1)):
                a[p:p + w] = rng.integers(40, 100)
            q = a.tobytes()
        else:
            q = bytes([int(qual)]) * L
        parts.append(b"@read_%d/ccs np=%d\n" % (i, i % 17) + s + b"\n+\n" + q + b"\n")
    return b"".join(parts)


def _check(kmers, text):
    member = kmers.fastq_gzip(text)
    assert member is not None
    assert gzip.decompress(member) == text
    assert kmers.fastq_gzip(text, verify=False) == member  # the self-check never alters the output
    return member


@pytest.mark.parametrize("qual", ["random", "runs", "73"])
def test_long_read_fastq_round_trips_and_beats_zlib(kmers, qual):
    text = _fastq(60, 20_000, 1, qual)
    member = _check(kmers, text)
    assert len(member) < len(zlib.compress(text, 6)), "expected to be smaller than zlib -6 on long reads"
    assert len(member) < 0.75 * len(text)


def test_fasta_and_odd_shapes(kmers):
    rng = np.random.default_rng(2)
    fasta = b"".join(b">contig%d\n" % i + rng.choice(list(b"ACGTN"), 50_000).astype(np.uint8).tobytes() + b"\n" for i in range(8))
    _check(kmers, fasta)
    one_long_line = b"@r\n" + rng.choice(list(b"ACGT"), 900_000).astype(np.uint8).tobytes() + b"\n+\n" + b"I" * 900_000 + b"\n"
    _check(kmers, one_long_line)                      # lines longer than a block are cut mid-line
    _check(kmers, b"A" * 300_000 + b"\n")              # runs far longer than 258
    _check(kmers, bytes(b for b in range(256) if b != 10) * 4000)  # every other byte value, no newline at all
    assert kmers.fastq_gzip(bytes(range(256)) * 4000) is None        # 256-byte lines: left to zlib
    no_final_newline = _fastq(5, 30_000, 3)[:-1]
    _check(kmers, no_final_newline)
    lower = _fastq(5, 30_000, 4).lower()
    _check(kmers, lower)
    crlf = _fastq(5, 30_000, 5).replace(b"\n", b"\r\n")
    _check(kmers, crlf)
    # Fibonacci frequencies: the unrestricted Huffman tree is ~25 deep, the code must be cut to 14 bits
    fib = [1, 2]
    while sum(fib) < 200_000:
        fib.append(fib[-1] + fib[-2])
    skewed = np.concatenate([np.full(f, 40 + i, dtype=np.uint8) for i, f in enumerate(fib)])
    rng.shuffle(skewed)
    _check(kmers, skewed.tobytes() + b"\n")


def test_reference_fixture_reads(kmers):
    import os

    from conftest import GOLDEN

    text = gzip.open(os.path.join(GOLDEN, "ref_data", "test.ccs.fastq.gz"), "rb").read()
    _check(kmers, text)


def test_declines_what_zlib_does_better(kmers):
    short_reads = b"".join(b"@r%d\n" % i + b"ACGTTGCA" * 18 + b"\n+\n" + b"F" * 144 + b"\n" for i in range(2000))
    assert kmers.fastq_gzip(short_reads) is None      # 150-bp reads: lines too short for a block each
    assert kmers.fastq_gzip(b"@r\nACGT\n+\nIIII\n") is None
    assert kmers.fastq_gzip(b"") is None


@settings(max_examples=60, deadline=None)
@given(st.lists(st.tuples(st.sampled_from([b"ACGT", b"ACGTN", b"!\"#$%&'()*+,-./0123456789:;<=>?@ABCDEFGHI", b"~", b"\x00\xff\n"]),
                          st.integers(1, 40_000), st.integers(0, 2**32 - 1)), min_size=1, max_size=8))
def test_random_line_structures(pieces):
    from trio_binning_b200 import kmers

    parts = []
    for alphabet, length, seed in pieces:
        rng = np.random.default_rng(seed)
        parts.append(rng.choice(list(alphabet), length).astype(np.uint8).tobytes() + b"\n")
    text = b"".join(parts)
    member = kmers.fastq_gzip(text)
    if member is not None:
        assert gzip.decompress(member) == text


@pytest.mark.parametrize("fastq_encoder", [False, True])
def test_parallel_member_writer_keeps_order(kmers, tmp_path, fastq_encoder):
    """The output writer cuts the text into 4 MB chunks, compresses them on a thread pool as
    independent gzip members and writes them in order (with back-pressure): the file must
    decompress to exactly the input, whatever the piece size the text arrives in."""
    text = _fastq(700, 20_000, 7, "runs")  # ~28 MB: several chunks per run
    src = tmp_path / "in.fastq"
    src.write_bytes(text)
    for piece, threads, level in ((1 << 16, 3, 1), (4_000_001, 2, 6 if fastq_encoder else 1), (123_457, 1, 1)):
        out = tmp_path / f"out_{piece}.gz"
        assert kmers.gzip_copy(str(src), str(out), level=level, threads=threads, fastq_encoder=fastq_encoder,
                               piece=piece) == len(text)
        blob = out.read_bytes()
        assert gzip.decompress(blob) == text
        assert blob.count(b"\x1f\x8b\x08") >= 3  # several members
        assert kmers.gunzip(str(out)) == text  # and our own reader agrees
    empty = tmp_path / "empty"
    empty.write_bytes(b"")
    out = tmp_path / "empty.gz"
    assert kmers.gzip_copy(str(empty), str(out), fastq_encoder=fastq_encoder) == 0
    assert gzip.decompress(out.read_bytes()) == b""  # Python always writes a valid empty gzip file
