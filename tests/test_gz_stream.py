"""The native reader's own streaming inflate (csrc/gz_stream.cpp) against Python's gzip/zlib --
what the reference reads its input with (seq.py:88-92).  CPU only.  Semantics pinned here are
those of Python's gzip module: concatenated members, zero padding skipped, anything else that is
not gzip is an error; every CRC-32 / ISIZE verified."""
import gzip
import os
import struct
import zlib

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st


@pytest.fixture(scope="module")
def kmers():
    from trio_binning_b200 import kmers as mod

    return mod


def _member(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, wbits=31, memlevel=8):
    c = zlib.compressobj(level, zlib.DEFLATED, wbits, memlevel, strategy)
    return c.compress(data) + c.flush()


def _check(kmers, tmp_path, blob, want, name="x.gz", chunks=(0, 1 << 20, 4093, 7)):
    p = tmp_path / name
    p.write_bytes(blob)
    for chunk in chunks:
        got = kmers.gunzip(str(p), chunk)
        assert got == want, (name, chunk, len(got), len(want))


def _corpus():
    rng = np.random.default_rng(11)
    fastq = b"".join(b"@r%d desc\n" % i + rng.choice(list(b"ACGT"), 300).astype(np.uint8).tobytes() + b"\n+\n" +
                     bytes(rng.integers(33, 74, 300, dtype=np.uint8)) + b"\n" for i in range(400))
    period = bytes(rng.integers(0, 256, 32768, dtype=np.uint8))
    return {
        "empty": b"",
        "one": b"A",
        "short": b"hello, hello, hello world\n",
        "fastq": fastq,
        "random": bytes(rng.integers(0, 256, 300_000, dtype=np.uint8)),
        "zeros": bytes(1_000_000),
        "run_lengths": b"".join(bytes([i % 251]) * (i % 700 + 1) for i in range(3000)),
        "period_32768": period * 40,           # matches at the maximum distance, across window slides
        "period_32767_plus": (period[:-1]) * 40 + b"tail",
        "two_symbols": rng.choice([65, 67], 200_000).astype(np.uint8).tobytes(),
        "big_mixed": fastq * 30 + bytes(rng.integers(0, 256, 3_000_000, dtype=np.uint8)) + fastq * 30,
    }


@pytest.mark.parametrize("level", [0, 1, 6, 9])
def test_round_trip_levels(kmers, tmp_path, level):
    for name, data in _corpus().items():
        _check(kmers, tmp_path, _member(data, level), data, name=f"{name}_{level}.gz",
               chunks=(0, 4093) if len(data) > 100_000 else (0, 1 << 20, 4093, 7))


@pytest.mark.parametrize("strategy", [zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FILTERED])
def test_round_trip_strategies(kmers, tmp_path, strategy):
    for name, data in _corpus().items():
        if len(data) > 1_500_000:
            continue
        _check(kmers, tmp_path, _member(data, 6, strategy), data, name=f"{name}_s{strategy}.gz", chunks=(0, 4093))


def test_small_memlevel_and_window(kmers, tmp_path):
    data = _corpus()["fastq"]
    _check(kmers, tmp_path, _member(data, 9, memlevel=1), data, chunks=(0,))      # tiny blocks: many headers
    _check(kmers, tmp_path, _member(data, 9, wbits=16 + 9), data, chunks=(0,))    # 512-byte window


def test_python_gzip_module_output(kmers, tmp_path):
    data = _corpus()["fastq"]
    p = tmp_path / "named.fastq.gz"
    with gzip.open(p, "wb", compresslevel=9) as f:  # writes FNAME + MTIME
        f.write(data)
    assert kmers.gunzip(str(p)) == data


def test_members_concatenated_and_trailing_bytes(kmers, tmp_path):
    c = _corpus()
    parts = [c["fastq"], b"", c["short"], c["zeros"][:70_000], c["random"][:50_000]]
    blob = b"".join(_member(x, lvl) for x, lvl in zip(parts, (6, 9, 1, 6, 0)))
    _check(kmers, tmp_path, blob, b"".join(parts))
    # zero padding after (and between) members is skipped, as Python's gzip module does
    _check(kmers, tmp_path, blob + bytes(100), b"".join(parts), name="pad.gz", chunks=(0,))
    two = _member(parts[0]) + bytes(13) + _member(parts[2]) + bytes(1)
    assert gzip.decompress(two) == parts[0] + parts[2]
    _check(kmers, tmp_path, two, parts[0] + parts[2], name="pad_between.gz", chunks=(0, 7))
    # anything else after a member is an error upstream (gzip.BadGzipFile) and here
    for i, junk in enumerate((b"garbage, not gzip", bytes(5) + b"x", b"\x1f", b"\x1f\x8b\x08")):
        p = tmp_path / f"junk{i}.gz"
        p.write_bytes(blob + junk)
        with pytest.raises((gzip.BadGzipFile, EOFError)):
            gzip.decompress(blob + junk)
        with pytest.raises(ValueError):
            kmers.gunzip(str(p))


def test_not_gzip_is_an_error_like_pythons_gzip(kmers, tmp_path):
    """The reference opens *.gz with gzip.open (seq.py:88-92): a file of that name that is not
    gzip raises there and must not be read as text here; an empty file reads as empty in both."""
    _check(kmers, tmp_path, b"", b"", name="empty_file.gz", chunks=(0,))
    assert gzip.decompress(b"") == b""
    for i, data in enumerate((_corpus()["fastq"], b"\x1f", b"A", b"\x00\x00")):
        p = tmp_path / f"plain{i}.gz"
        p.write_bytes(data)
        with pytest.raises((gzip.BadGzipFile, EOFError)):
            with gzip.open(p, "rb") as f:
                f.read()
        with pytest.raises(ValueError):
            kmers.gunzip(str(p))


def test_optional_header_fields(kmers, tmp_path):
    data = _corpus()["short"] * 50
    raw = _member(data, 6, wbits=-15)
    trailer = struct.pack("<II", zlib.crc32(data), len(data) & 0xFFFFFFFF)
    for flg, extra in ((0, b""), (8, b"name.txt\0"), (16, b"a comment\0"), (4, struct.pack("<H", 6) + b"BC\x02\x00\x12\x34"),
                       (2, None), (4 | 8 | 16, struct.pack("<H", 3) + b"xyz" + b"n\0" + b"c\0")):
        head = bytes([0x1f, 0x8b, 8, flg, 0, 0, 0, 0, 0, 255])
        if extra is None:  # FHCRC: low 16 bits of the CRC-32 of the header so far
            extra = struct.pack("<H", zlib.crc32(head) & 0xFFFF)
        _check(kmers, tmp_path, head + extra + raw + trailer, data, name=f"flg{flg}.gz", chunks=(0, 7))


def test_corrupt_and_truncated_inputs_are_errors(kmers, tmp_path):
    data = _corpus()["fastq"]
    blob = bytearray(_member(data, 6))
    p = tmp_path / "bad.gz"

    def fails(b):
        p.write_bytes(bytes(b))
        with pytest.raises(ValueError):
            kmers.gunzip(str(p))

    for cut in (len(blob) - 1, len(blob) - 8, len(blob) - 9, len(blob) // 2, 12, 5):
        fails(blob[:cut])
    for pos in (len(blob) // 3, len(blob) // 2, len(blob) - 12):
        b = bytearray(blob)
        b[pos] ^= 0x5A
        fails(b)
    b = bytearray(blob)
    b[-1] ^= 1  # ISIZE
    fails(b)
    b = bytearray(blob)
    b[-5] ^= 1  # CRC-32
    fails(b)
    b = bytearray(blob)
    b[2] = 7    # unknown compression method in the FIRST header: not a gzip file we can read
    fails(b)
    # a member that reaches back before its own start (distance beyond the data decoded so far)
    bad = zlib.compressobj(6, zlib.DEFLATED, -15)
    raw = bytes([0b01100011, 0b00000000]) + b"\x00" * 4  # fixed block: a match as the first symbol
    fails(bytes([0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 0, 255]) + raw + struct.pack("<II", 0, 0))
    del bad


def test_crc32_matches_zlib(kmers):
    rng = np.random.default_rng(3)
    buf = bytes(rng.integers(0, 256, 200_000, dtype=np.uint8))
    for n in (0, 1, 15, 16, 63, 64, 65, 79, 80, 127, 128, 1000, 4096, 65_537, 200_000):
        for off in (0, 1, 3, 8):
            for seed in (0, 0xFFFFFFFF, 0x12345678):
                assert kmers.crc32(buf[off:off + n], seed) == zlib.crc32(buf[off:off + n], seed), (n, off, seed)
    flags = open("/proc/cpuinfo").read() if os.path.exists("/proc/cpuinfo") else ""
    if "pclmulqdq" in flags and "sse4_1" in flags:
        assert kmers.crc32_accelerated()  # the self-check at start-up accepted the fast path


@settings(max_examples=150, deadline=None)
@given(st.lists(st.tuples(st.binary(min_size=0, max_size=600), st.integers(1, 40)), min_size=0, max_size=12),
       st.integers(0, 9), st.sampled_from([zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_RLE, zlib.Z_HUFFMAN_ONLY]),
       st.integers(1, 9))
def test_random_structures(tmp_path_factory, pieces, level, strategy, memlevel):
    from trio_binning_b200 import kmers

    data = b"".join(p * r for p, r in pieces)
    p = tmp_path_factory.mktemp("gz") / "h.gz"
    p.write_bytes(_member(data, level, strategy, memlevel=memlevel))
    assert kmers.gunzip(str(p), 0) == data
    assert kmers.gunzip(str(p), 13) == data
