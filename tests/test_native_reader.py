"""The native FASTX reader and TSV float formatting of csrc/fastx_pipeline.cpp, on the CPU.

Oracle: trio_binning_b200/seq.py on the same FILES (its open_fastx_read is the reference's
call, `readfq(open(name, "r"))` / `gzip.open(name, "rt")`, so text-mode universal newlines
apply on both sides), and seq.py itself is pinned to the reference's readfq by
tests/test_seq.py.  Records must not contain TABs for the dump format used here."""
import gzip
import os
import random

import pytest


@pytest.fixture(scope="module")
def kmers():
    from trio_binning_b200 import kmers as mod

    return mod


@pytest.fixture(scope="module")
def seq():
    from trio_binning_b200 import seq as mod

    return mod


def _python_records(seq, path):
    return [(r.name.encode(), r.seq.encode(), None if r.qual is None else r.qual.encode())
            for r in seq.open_fastx_read(str(path))]


def test_native_reader_on_golden_edge_cases(kmers, seq, golden_seq, tmp_path):
    for name, case in golden_seq.items():
        text = case["input"].replace("\t", " ")
        for suffix in ("", ".gz"):
            p = tmp_path / f"{name}.fx{suffix}"
            raw = text.encode()
            if suffix:
                with gzip.open(p, "wb") as f:
                    f.write(raw)
            else:
                p.write_bytes(raw)
            assert kmers.fastx_dump(str(p)) == _python_records(seq, p), (name, suffix)


def test_native_reader_newline_flavours(kmers, seq, tmp_path):
    cases = {
        "crlf": b"@r1 c\r\nACGT\r\n+\r\n!!!!\r\n>r2\r\nGG\r\nTT\r\n",
        "cr_only": b">r1\rACGT\r>r2\rGG\r",
        "mixed": b">r1\nAC\r\nGT\rTT\n@q\nAA\n+\n!!\n",
        "cr_at_chunk_edge": b">r1\n" + b"A" * 70 + b"\r\n" + b"C" * 10 + b"\r",
        "utf8_name": ">résumé x\nACGT\n".encode(),
        # Python counts CHARACTERS when it decides where a quality block ends (seq.py:75-80)
        "utf8_seq_and_qual": "@r1\nACéGT\n+\n!!é!!\n@r2\nAC\n+\n!!\n".encode(),
        "utf8_qual_two_lines": "@r1\nACGTAC\n+\né!!\n!!é\n>r2\nGG\n".encode(),
        "utf8_qual_longer_in_bytes": "@r1\nACGT\n+\nééé\né\n@r2\nAC\n+\n!!\n".encode(),
        "fastq_no_final_newline": b"@r1\nACGT\n+\n!!!!",
        "plus_only_at_eof": b"@r1\nACGT\n+",
        "header_only_at_eof": b">",
    }
    for name, raw in cases.items():
        p = tmp_path / f"{name}.fq"
        p.write_bytes(raw)
        assert kmers.fastx_dump(str(p)) == _python_records(seq, p), name


def test_native_reader_random_files(kmers, seq, tmp_path):
    rng = random.Random(5)
    for trial in range(6):
        lines = []
        for r in range(300):
            L = rng.randrange(0, 400)
            s = "".join(rng.choice("ACGTNacgt") for _ in range(L))
            width = rng.choice([60, 80, 10**9])
            chunks = [s[i:i + width] for i in range(0, L, width)] or [""]
            if rng.random() < 0.5:
                lines.append(f">read{r} comment {trial}")
                lines += chunks
            else:
                q = "".join(chr(rng.randrange(33, 74)) for _ in range(L))
                lines.append(f"@read{r}")
                lines += chunks
                lines.append("+" if rng.random() < 0.7 else f"+read{r}")
                lines += [q[i:i + width] for i in range(0, L, width)] or [""]
        text = "\n".join(lines) + ("\n" if trial % 2 == 0 else "")
        p = tmp_path / f"rand{trial}.fastq.gz"
        with gzip.open(p, "wt") as f:
            f.write(text)
        got = kmers.fastx_dump(str(p))
        assert got == _python_records(seq, p), trial
        assert len(got) >= 299


def test_float_repr_matches_python(kmers):
    rng = random.Random(1)
    values = [0.0, 1.0, 4.0, 4 / 3, 8 / 3, 1e15, 1e16, 1.5e16, 123456789012345680.0, 1e22, 1e-4, 1e-5,
              0.00012345, 5e-324, 1.7976931348623157e308, 2.5, 100.0, 1e100, 3 * (7 / 3), 0.1 + 0.2]
    values += [rng.randrange(0, 2**31) * (rng.randrange(1, 10**9) / rng.randrange(1, 10**9)) for _ in range(3000)]
    values += [rng.random() * 10 ** rng.randrange(-30, 30) for _ in range(3000)]
    for v in values:
        assert kmers.float_repr(v) == repr(float(v)), v


def test_corrupt_gzip_input_is_an_error_not_a_short_result(tmp_path):
    """The reference reads through Python's gzip module, which raises on damaged files; the native
    reader must not hand back the records it managed to parse as if the file had ended there."""
    import numpy as np

    from trio_binning_b200 import kmers

    rng = np.random.default_rng(9)
    text = b"".join(b"@r%d\n" % i + rng.choice(list(b"ACGT"), 500).astype(np.uint8).tobytes() + b"\n+\n" + b"I" * 500 + b"\n"
                    for i in range(300))
    good = tmp_path / "ok.fastq.gz"
    good.write_bytes(gzip.compress(text))
    assert len(kmers.fastx_dump(str(good))) == 300
    blob = bytearray(good.read_bytes())
    for name, damaged in (("cut", blob[: len(blob) // 2]), ("flip", blob[:4000] + bytes([blob[4000] ^ 0xFF]) + blob[4001:]),
                          ("no_trailer", blob[:-8])):
        p = tmp_path / f"{name}.fastq.gz"
        p.write_bytes(bytes(damaged))
        with pytest.raises(RuntimeError, match="gzip"):
            kmers.fastx_dump(str(p))


def _bgzf(data, block=65280, level=6, eof_marker=True):
    """bgzip's container (SAM spec 4.1): independent gzip members of <= 64 KB, each with a 'BC'
    extra subfield holding its total size - 1."""
    import struct
    import zlib

    out = bytearray()
    pieces = [data[i:i + block] for i in range(0, len(data), block)] + ([b""] if eof_marker else [])
    for piece in pieces:
        c = zlib.compressobj(level, zlib.DEFLATED, -15)
        raw = c.compress(piece) + c.flush()
        total = 12 + 6 + len(raw) + 8
        out += bytes([0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff]) + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, total - 1)
        out += raw + struct.pack("<II", zlib.crc32(piece), len(piece))
    return bytes(out)


def _fastq_text(n, length, seed):
    import numpy as np

    rng = np.random.default_rng(seed)
    recs, parts = [], []
    for i in range(n):
        s = rng.choice(list(b"ACGT"), length).astype(np.uint8).tobytes()
        q = rng.integers(33, 74, length, dtype=np.uint8).tobytes()
        recs.append((b"r%d" % i, s, q))
        parts.append(b"@r%d some comment\n" % i + s + b"\n+\n" + q + b"\n")
    return recs, b"".join(parts)


def test_bgzf_input_is_decoded_in_parallel_and_in_order(tmp_path):
    """BGZF (bgzip) input: groups of blocks are inflated by worker threads and must come out in
    stream order, records spanning block and task boundaries intact."""
    import gzip

    from trio_binning_b200 import kmers

    recs, text = _fastq_text(700, 3000, 21)  # ~4.2 MB: several 512 KB tasks
    for name, blob in (("bgzf.fastq.gz", _bgzf(text)), ("bgzf_small_blocks.fastq.gz", _bgzf(text, block=777, level=1)),
                       ("bgzf_no_eof.fastq.gz", _bgzf(text, eof_marker=False)),
                       ("bgzf_then_gzip.fastq.gz", _bgzf(text[:2_000_003], eof_marker=False) + gzip.compress(text[2_000_003:])),
                       ("gzip_then_bgzf.fastq.gz", gzip.compress(text[:1_000_001]) + _bgzf(text[1_000_001:])),
                       ("bgzf_padded.fastq.gz", _bgzf(text) + bytes(37))):
        p = tmp_path / name
        p.write_bytes(blob)
        assert gzip.decompress(blob) == text  # (Python's gzip tolerates zero padding after the last member)
        got = kmers.fastx_dump(str(p))
        assert got == recs, name
        assert kmers.gunzip(str(p)) == text  # the serial decoder reads the same container


def test_bgzf_damage_is_an_error(tmp_path):
    from trio_binning_b200 import kmers

    _, text = _fastq_text(300, 3000, 22)
    blob = bytearray(_bgzf(text))
    cases = {
        "flip_payload": lambda b: b[:200_000] + bytes([b[200_000] ^ 0x40]) + b[200_001:],
        "cut_mid_block": lambda b: b[: len(b) // 2],
        "wrong_isize": lambda b: b[:-28 - 1] + bytes([b[-28 - 1] ^ 1]) + b[-28:],
    }
    for name, fn in cases.items():
        p = tmp_path / f"{name}.fastq.gz"
        p.write_bytes(bytes(fn(blob)))
        with pytest.raises(RuntimeError, match="gzip"):
            kmers.fastx_dump(str(p))


def test_bgzf_forged_isize_is_an_error_not_an_abort(tmp_path):
    """blocks claiming ISIZE = 0xFFFFFFFF made a worker thread throw bad_alloc and
    std::terminate killed the process.  A BGZF block holds at most 64 KiB: larger claims go to the
    serial decoder, which reports the corruption through tb_last_error."""
    import struct
    import subprocess
    import sys

    from trio_binning_b200 import kmers

    _, text = _fastq_text(100, 3000, 23)
    blob = bytearray(_bgzf(text, eof_marker=False))
    # every block's trailer: CRC32, ISIZE -- forge ISIZE in all of them
    pos = 0
    while pos < len(blob):
        total = struct.unpack_from("<H", blob, pos + 16)[0] + 1
        for claim in (0xFFFFFFFF,):
            struct.pack_into("<I", blob, pos + total - 4, claim)
        pos += total
    p = tmp_path / "forged.fastq.gz"
    p.write_bytes(bytes(blob))
    # in a child process: the failure mode being guarded against is abort(), not an exception
    code = ("import sys; sys.path.insert(0, %r)\nfrom trio_binning_b200 import kmers\n"
            "try:\n    kmers.fastx_dump(%r)\nexcept RuntimeError as e:\n    print('ERR', e); sys.exit(7)\n"
            % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), str(p)))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 7 and "gzip" in r.stdout, (r.returncode, r.stdout, r.stderr[-500:])
    # a moderately wrong ISIZE (70 000 > 64 KiB) takes the same road
    blob2 = bytearray(_bgzf(text, eof_marker=False))
    total = struct.unpack_from("<H", blob2, 16)[0] + 1
    struct.pack_into("<I", blob2, total - 4, 70_000)
    p2 = tmp_path / "forged2.fastq.gz"
    p2.write_bytes(bytes(blob2))
    with pytest.raises(RuntimeError, match="gzip"):
        kmers.fastx_dump(str(p2))


def test_records_with_tabs_survive_the_dump(kmers, seq, tmp_path):
    """Names are cut at the first SPACE only (seq.py:61), so tabs stay in names; the dump helper
    must not split on them."""
    raw = b">name\twith\ttabs comment\nAC\tGT\n@q\t1\nAC\tG\n+\n!\t!!\n"
    p = tmp_path / "tabs.fa"
    p.write_bytes(raw)
    want = _python_records(seq, p)
    assert want[0][0] == b"name\twith\ttabs" and want[1][1] == b"AC\tG"
    assert kmers.fastx_dump(str(p)) == want


def test_line_endings_across_the_4_mib_chunks_of_the_read_ahead(kmers, seq, tmp_path):
    """The reader hands the parser 4 MiB chunks: a "\\r\\n" split exactly at a chunk edge must still
    be one line end, a lone "\\r" there a line end of its own, and records must survive the cut --
    plain and gzip (the inflate has chunk edges of its own, at other offsets)."""
    edge = 4 << 20
    for flavour in ("crlf_split", "cr_last", "lf_first"):
        head = b">first record\r\n"
        filler_len = edge - len(head) - (1 if flavour != "lf_first" else 0)
        if flavour == "crlf_split":      # ...'\r' is the last byte of chunk 1, '\n' the first of chunk 2
            body = b"A" * filler_len + b"\r\n" + b"CCCC\r\n>second\r\nGGGG\r\n"
        elif flavour == "cr_last":       # a lone '\r' ends chunk 1, chunk 2 starts with sequence
            body = b"A" * filler_len + b"\r" + b"CCCC\r>second\rGGGG\r"
        else:                            # '\n' is the first byte of chunk 2
            body = b"A" * filler_len + b"\n" + b"CCCC\n@third\nGG\n+\n!!\n"
        raw = head + body
        assert raw[edge - 1:edge + 1] in (b"\r\n", b"\rC", b"A\n")
        plain = tmp_path / f"{flavour}.fa"
        plain.write_bytes(raw)
        want = _python_records(seq, plain)
        assert len(want) >= 2
        assert kmers.fastx_dump(str(plain)) == want, flavour
        gz = tmp_path / f"{flavour}.fa.gz"
        with gzip.open(gz, "wb", compresslevel=1) as f:
            f.write(raw)
        assert kmers.fastx_dump(str(gz)) == want, flavour + " (gz)"


def test_bgzf_writer_round_trip(kmers, tmp_path):
    """tb_gzip_copy's BGZF mode (input generator for benchmarks): valid gzip for Python, BGZF for
    the parallel reader, identical text."""
    import gzip

    recs, text = _fastq_text(400, 3000, 31)
    src = tmp_path / "in.fastq"
    src.write_bytes(text)
    out = tmp_path / "out.fastq.gz"
    kmers.gzip_copy(str(src), str(out), level=6, threads=3, bgzf=True, piece=100_000)
    blob = out.read_bytes()
    assert blob[:4] == b"\x1f\x8b\x08\x04" and blob[12:14] == b"BC"
    assert gzip.decompress(blob) == text
    assert kmers.fastx_dump(str(out)) == recs
    empty = tmp_path / "empty.fastq"
    empty.write_bytes(b"")
    kmers.gzip_copy(str(empty), str(tmp_path / "e.gz"), bgzf=True)
    assert gzip.decompress((tmp_path / "e.gz").read_bytes()) == b""


def test_mixed_newlines_and_utf8_everywhere(kmers, seq, tmp_path):
    """A file that mixes "\\n", "\\r\\n" and "\\r" line ends record by record, has tabs in names and
    multi-byte characters in sequences and qualities (Python counts characters, not bytes, when it
    decides where a quality block ends): same records as the Python reader, plain and gzipped."""
    rng = random.Random(5)
    parts = []
    for i in range(3000):
        n = rng.choice([1, 30, 150, 5000])
        s = "".join(rng.choice("ACGTNacgté") for _ in range(n))
        q = "".join(rng.choice("!#5Iéü") for _ in range(n))
        nl = rng.choice(["\n", "\r\n", "\r"])
        parts.append(f">f{i} x{nl}{s}{nl}" if i % 5 == 0 else f"@r{i}\tt c{nl}{s}{nl}+{nl}{q}{nl}")
    raw = "".join(parts).encode()
    plain = tmp_path / "mixed.fastq"
    plain.write_bytes(raw)
    want = _python_records(seq, plain)
    assert len(want) > 2000
    assert kmers.fastx_dump(str(plain)) == want
    gz = tmp_path / "mixed.fastq.gz"
    with gzip.open(gz, "wb", compresslevel=1) as f:
        f.write(raw)
    assert kmers.fastx_dump(str(gz)) == want
