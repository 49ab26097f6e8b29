"""One plain gzip stream decoded on several cores (csrc/gz_stream.cpp: SpanInflate; the coordinator in
csrc/fastx_pipeline.cpp: LineSource::produce_pgzip).  What must hold whatever the span size, the
number of workers or the shape of the stream: the text delivered to the parser is exactly what
Python's gzip module -- the reference's reader, seq.py:88-92 -- returns, and whenever Python's
module raises, so does the reader, with the serial reader's own message.  CPU only."""
import gzip
import os
import zlib

import numpy as np
import pytest


@pytest.fixture(scope="module")
def kmers():
    from trio_binning_b200 import kmers as mod

    return mod


@pytest.fixture
def spans(monkeypatch):
    """Force the parallel path onto small files: spans of a few KiB, no minimum size."""
    def use(span_bytes, enabled=True):
        monkeypatch.setenv("TB_PGZIP", "1" if enabled else "0")
        monkeypatch.setenv("TB_PGZIP_SPAN", str(span_bytes))
        monkeypatch.setenv("TB_PGZIP_MIN", "0")
    return use


def _fastq(n_reads, mean_len, seed, qual_symbols=41):
    rng = np.random.default_rng(seed)
    parts = []
    for i in range(n_reads):
        n = int(mean_len * (0.5 + rng.random()))
        s = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), n).tobytes()
        q = rng.integers(33, 33 + qual_symbols, n, dtype=np.uint8).tobytes()
        parts.append(b"@read%d some comment\n" % i + s + b"\n+\n" + q + b"\n")
    return b"".join(parts)


def _both(kmers, spans, path, span_bytes, threads=4):
    """The text through the parallel and through the serial reader (or the exception text)."""
    out = []
    for enabled in (True, False):
        spans(span_bytes, enabled)
        try:
            out.append(kmers.fastx_cat(path, threads))
        except RuntimeError as e:
            out.append(str(e).split(": ", 1)[-1])
    return out


@pytest.mark.parametrize("level", [1, 6, 9])
@pytest.mark.parametrize("span_bytes", [2048, 20_000, 300_000])
def test_levels_and_span_sizes(kmers, spans, tmp_path, level, span_bytes):
    text = _fastq(120, 9000, level)
    path = str(tmp_path / "reads.fastq.gz")
    with open(path, "wb") as f:
        f.write(gzip.compress(text, level))
    parallel, serial = _both(kmers, spans, path, span_bytes)
    assert parallel == text and serial == text


def test_short_reads_with_many_matches(kmers, spans, tmp_path):
    """150-base reads, binned qualities, names that repeat: most of the stream is matches, and the
    markers of a span survive deep into it."""
    rng = np.random.default_rng(3)
    genome = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 30_000).tobytes()
    parts = []
    for i in range(6000):
        p = int(rng.integers(0, len(genome) - 150))
        q = rng.choice(np.frombuffer(b"#-<F", dtype=np.uint8), 150, p=[0.02, 0.03, 0.15, 0.8]).tobytes()
        parts.append(b"@A00123:45:HXXXXXXX:1:1101:%d:%d 1:N:0:ACGT\n" % (1000 + i, 2000 + i % 97) + genome[p:p + 150] + b"\n+\n" + q + b"\n")
    text = b"".join(parts)
    path = str(tmp_path / "short.fastq.gz")
    open(path, "wb").write(gzip.compress(text, 6))
    for span_bytes in (1500, 12_345):
        parallel, serial = _both(kmers, spans, path, span_bytes)
        assert parallel == text and serial == text


def test_members_padding_flushes_and_other_block_types(kmers, spans, tmp_path):
    a, b, c = _fastq(40, 8000, 10), _fastq(3, 200, 11), _fastq(60, 5000, 12)
    # sync and full flushes leave empty stored blocks and byte-aligned boundaries inside a member
    co = zlib.compressobj(6, zlib.DEFLATED, 31)
    flushed = b"".join([co.compress(a[:100_000]), co.flush(zlib.Z_SYNC_FLUSH), co.compress(a[100_000:200_000]),
                        co.flush(zlib.Z_FULL_FLUSH), co.compress(a[200_000:]), co.flush()])
    stored = gzip.compress(os.urandom(70_000).replace(b"\r", b"\x01"), 6)          # incompressible: stored blocks
    stored_text = gzip.decompress(stored)
    fixed = gzip.compress(b"@r\nACGT\n+\nIIII\n", 9)                                # tiny: one fixed-Huffman block
    huff_only = zlib.compressobj(6, zlib.DEFLATED, 31, 8, zlib.Z_HUFFMAN_ONLY)
    huff = huff_only.compress(c) + huff_only.flush()
    cases = {
        "members": (gzip.compress(a, 6) + gzip.compress(b, 1) + gzip.compress(c, 9), a + b + c),
        "padding": (gzip.compress(a, 6) + b"\0" * 700 + gzip.compress(c, 6) + b"\0" * 5000, a + c),
        "flushes": (flushed + gzip.compress(b), a + b),
        "stored": (gzip.compress(a) + stored + gzip.compress(c), a + stored_text + c),
        "fixed": (fixed * 50 + gzip.compress(a), b"@r\nACGT\n+\nIIII\n" * 50 + a),
        "huffman_only": (huff + gzip.compress(a, 1), c + a),
    }
    own = kmers.fastq_gzip(a)
    assert own is not None
    cases["own_encoder"] = (own + own, a + a)                                       # one distance code of one bit
    for name, (blob, text) in cases.items():
        path = str(tmp_path / (name + ".fq.gz"))
        open(path, "wb").write(blob)
        assert gzip.decompress(blob) == text, name
        for span_bytes in (1024, 7000, 50_000):
            parallel, serial = _both(kmers, spans, path, span_bytes)
            # '\r' never occurs in these texts, so the line layer passes them through unchanged
            assert parallel == text, (name, span_bytes)
            assert serial == text, (name, span_bytes)


def test_header_fields(kmers, spans, tmp_path):
    text = _fastq(50, 6000, 20)
    body = zlib.compress(text, 6)[2:-4]          # raw deflate
    trailer = zlib.crc32(text).to_bytes(4, "little") + (len(text) & 0xFFFFFFFF).to_bytes(4, "little")
    extra = b"AB\x04\x00abcd"
    header = (b"\x1f\x8b\x08" + bytes([4 | 8 | 16 | 2]) + b"\0\0\0\0\x02\xff" + len(extra).to_bytes(2, "little") + extra +
              b"reads.fastq\0" + b"a comment\0")
    header += (zlib.crc32(header) & 0xFFFF).to_bytes(2, "little")
    blob = header + body + trailer
    assert gzip.decompress(blob) == text
    path = str(tmp_path / "fields.fq.gz")
    open(path, "wb").write(blob)
    parallel, serial = _both(kmers, spans, path, 3000)
    assert parallel == text and serial == text


def test_damage_is_reported_like_the_serial_reader_does(kmers, spans, tmp_path):
    text = _fastq(80, 7000, 30)
    blob = gzip.compress(text, 6)
    rng = np.random.default_rng(31)
    variants = {
        "truncated_body": blob[:len(blob) // 2],
        "truncated_trailer": blob[:-3],
        "garbage_after": blob + b"not gzip at all",
        "wrong_crc": blob[:-8] + b"\x00\x11\x22\x33" + blob[-4:],
        "wrong_isize": blob[:-4] + b"\x01\x00\x00\x00",
        "second_member_damaged": blob + gzip.compress(text, 6)[:20_000] + b"\xff" * 300 + gzip.compress(text, 6)[20_300:],
    }
    for i in range(12):  # one flipped byte somewhere in the deflate data
        pos = int(rng.integers(40, len(blob) - 16))
        variants[f"flip_{pos}"] = blob[:pos] + bytes([blob[pos] ^ (1 << int(rng.integers(0, 8)))]) + blob[pos + 1:]
    for name, data in variants.items():
        path = str(tmp_path / (name + ".fq.gz"))
        open(path, "wb").write(data)
        try:
            want = gzip.decompress(data)
        except Exception:  # noqa: BLE001
            want = None
        for span_bytes in (2000, 25_000):
            parallel, serial = _both(kmers, spans, path, span_bytes)
            assert parallel == serial, (name, span_bytes)
            if want is None:
                assert isinstance(serial, str) and "gzip" in serial, (name, serial)
            else:
                assert serial == want, name


def test_workers_and_thread_counts(kmers, spans, tmp_path):
    text = _fastq(300, 12_000, 40)
    path = str(tmp_path / "reads.fastq.gz")
    open(path, "wb").write(gzip.compress(text, 6))
    spans(30_000)
    for threads in (1, 2, 3, 8, 12):
        assert kmers.fastx_cat(path, threads) == text


def test_records_through_the_parser(kmers, spans, tmp_path):
    """End of the chain: the records the parser makes of the parallel reader's text."""
    text = _fastq(200, 5000, 50).replace(b"\n+\n", b"\r\n+\r\n", 40)   # some CRLF line ends too
    path = str(tmp_path / "reads.fq.gz")
    open(path, "wb").write(gzip.compress(text, 6))
    spans(4096, enabled=False)
    want = kmers.fastx_dump(path)
    spans(4096, enabled=True)
    assert kmers.fastx_dump(path) == want and len(want) == 200


def test_random_streams_and_spans(kmers, spans, tmp_path):
    """Randomised: text kinds x zlib levels / strategies / window sizes x member layouts x span sizes."""
    rng = np.random.default_rng(2026)
    kinds = {
        "long": lambda: _fastq(int(rng.integers(5, 40)), int(rng.integers(2000, 12_000)), int(rng.integers(1 << 30))),
        "short": lambda: _fastq(int(rng.integers(200, 1500)), 150, int(rng.integers(1 << 30)), qual_symbols=4),
        "fasta": lambda: b"".join(b">c%d\n" % i + rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), 7000).tobytes() + b"\n"
                                  for i in range(int(rng.integers(3, 30)))),
        "repeats": lambda: (b"@r\n" + b"ACGT" * 500 + b"\n+\n" + b"I" * 2000 + b"\n") * int(rng.integers(20, 300)),
    }
    strategies = [zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED]
    for trial in range(25):
        members, text = [], []
        for _ in range(int(rng.integers(1, 4))):
            t = kinds[str(rng.choice(list(kinds)))]()
            co = zlib.compressobj(int(rng.integers(1, 10)), zlib.DEFLATED, 16 + int(rng.integers(9, 16)), int(rng.integers(1, 10)),
                                  int(rng.choice(strategies)))
            members.append(co.compress(t) + co.flush())
            text.append(t)
            if rng.random() < 0.3:
                members.append(b"\0" * int(rng.integers(1, 3000)))
        blob, want = b"".join(members), b"".join(text)
        assert gzip.decompress(blob) == want
        path = str(tmp_path / f"t{trial}.fq.gz")
        open(path, "wb").write(blob)
        spans(int(rng.integers(1024, 80_000)))
        assert kmers.fastx_cat(path, int(rng.integers(1, 9))) == want, trial
