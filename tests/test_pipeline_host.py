"""The host side of the native pipeline without a GPU (tb_route_fastx: the product's reader,
batches, worker threads, in-order router, writers and TSV; the classifier is replaced by a
fixed pattern of bins).  What must hold, against the Python writer that mirrors the reference
(trio_binning_b200/seq.py == reference seq.py:98-136 + Read.print :33-42):

  * any number of workers, either dealing mode, any batch size -> byte-identical outputs
    (the multi-GPU merge logic; reference loop classify_by_kmers.py:99-117 is sequential);
  * gzip mode: the decompressed bytes equal the reference writer's;
  * plain mode (--no-gzip-output): hapA's file, written through TWO handles upstream
    (seq.py:129), and the unclassified file are byte-identical to what CPython's buffered text
    writers produce, for record sizes on both sides of every buffer threshold.
"""
import gzip
import itertools
import os

import numpy as np
import pytest


@pytest.fixture(scope="module")
def kmers():
    from trio_binning_b200 import kmers as mod

    return mod


@pytest.fixture(scope="module")
def seq():
    from trio_binning_b200 import seq as mod

    return mod


def _fastx(path, lengths, seed, fasta_every=0, gz=False):
    """Reads of the given lengths, FASTQ (FASTA for every `fasta_every`-th), names with a comment."""
    rng = np.random.default_rng(seed)
    parts = []
    for i, n in enumerate(lengths):
        s = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), n).tobytes()
        if fasta_every and i % fasta_every == 0:
            parts.append(b">f%d fasta record\n" % i + s + b"\n")
        else:
            q = rng.integers(33, 74, n, dtype=np.uint8).tobytes()
            parts.append(b"@r%d/ccs np=%d\n" % (i, i % 7) + s + b"\n+\n" + q + b"\n")
    data = b"".join(parts)
    with (gzip.open(path, "wb", compresslevel=1) if gz else open(path, "wb")) as f:
        f.write(data)
    return len(lengths)


def _python_writer(seq, reads_path, bins, tmp, gzip_output, ext):
    """What the reference's loop writes for a given assignment of bins."""
    d = os.path.join(tmp, "py")
    os.makedirs(d, exist_ok=True)
    fa, fb, fu = seq.open_outfiles(os.path.join(d, "hapA"), os.path.join(d, "hapB"), os.path.join(d, "hapU"), ext,
                                   gzip_output)
    tsv = []
    for i, read in enumerate(seq.open_fastx_read(reads_path)):
        b = bins[i % len(bins)]
        read.print(file={"A": fa, "B": fb}.get(b, fu))
        n = len(read.seq)
        tsv.append("\t".join(map(str, [read.name, b, (n % 7) * 1.5, (n % 5) / 3.0])))
    fa.close()
    fb.close()
    fu.close()
    return d, "\n".join(tsv) + ("\n" if tsv else "")


def _native(kmers, reads_path, bins, tmp, tag, gzip_output, ext, **kw):
    d = os.path.join(tmp, tag)
    os.makedirs(d, exist_ok=True)
    suffix = ext + (".gz" if gzip_output else "")
    tsv_path = os.path.join(d, "out.tsv")
    with open(tsv_path, "wb") as t:
        stats = kmers.route_fastx_file(reads_path, bins, os.path.join(d, "hapA" + suffix), os.path.join(d, "hapB" + suffix),
                                       os.path.join(d, "hapU" + suffix), tsv_fd=t.fileno(), gzip_output=gzip_output, **kw)
    return d, open(tsv_path).read(), stats


def _content(path, gz):
    raw = open(path, "rb").read()
    return gzip.decompress(raw) if gz else raw


LONG = [20_000, 9_000, 14_000, 40_000, 100, 8_191, 8_192, 8_193, 4_095, 4_096, 4_097, 25_000, 1, 3_000] * 3
SHORT = [150] * 400 + [151, 149, 1, 2] * 10
MIXED = [n for pair in zip(LONG, itertools.cycle([150, 4000, 60])) for n in pair]


@pytest.mark.parametrize("shape,lengths", [("long", LONG), ("short", SHORT), ("mixed", MIXED)])
@pytest.mark.parametrize("bins", ["ABU", "AB", "BAAB", "A", "B", "UUBUA", "ABBBBBBBBA"])
def test_plain_output_equals_cpythons_writers(kmers, seq, tmp_path, shape, lengths, bins):
    """The reference's --no-gzip-output mode incl. its two-handles-on-one-file artefact."""
    tmp = str(tmp_path)
    reads = os.path.join(tmp, "reads.fastq")
    n = _fastx(reads, lengths, seed=len(bins) * 100 + len(lengths), fasta_every=5)
    want_dir, want_tsv = _python_writer(seq, reads, bins, tmp, False, ".fastq")
    got_dir, got_tsv, stats = _native(kmers, reads, bins, tmp, "native", False, ".fastq", batch_bases=50_000)
    assert stats["n_reads"] == n
    assert sorted(os.listdir(want_dir)) == ["hapA.fastq", "hapU.fastq"]  # upstream never creates hapB's file
    assert sorted(f for f in os.listdir(got_dir) if f != "out.tsv") == ["hapA.fastq", "hapU.fastq"]
    for name in ("hapA.fastq", "hapU.fastq"):
        assert _content(os.path.join(got_dir, name), False) == _content(os.path.join(want_dir, name), False), name
    assert got_tsv == want_tsv


@pytest.mark.parametrize("workers,round_robin,batch", [(1, False, 1 << 20), (3, False, 30_000), (4, True, 30_000),
                                                       (2, True, 1), (8, False, 5_000)])
def test_outputs_do_not_depend_on_workers_or_batches(kmers, seq, tmp_path, workers, round_robin, batch):
    tmp = str(tmp_path)
    reads = os.path.join(tmp, "reads.fq.gz")
    n = _fastx(reads, MIXED, seed=7, fasta_every=9, gz=True)
    bins = "ABUUABBAUB"
    want_dir, want_tsv = _python_writer(seq, reads, bins, tmp, True, ".fq")
    for gz_out in (True, False):
        ext = ".fq"
        if not gz_out:
            want_dir, _ = _python_writer(seq, reads, bins, os.path.join(tmp, "plain"), False, ext)
        for encoder in ((True, False) if gz_out else (True,)):
            got_dir, got_tsv, stats = _native(kmers, reads, bins, tmp, f"n_{gz_out}_{encoder}", gz_out, ext, workers=workers,
                                              round_robin=round_robin, batch_bases=batch, fastq_encoder=encoder,
                                              gzip_level=1)
            assert stats["n_reads"] == n and got_tsv == want_tsv
            assert stats["devices_used"] >= 1 and (not round_robin or stats["devices_used"] == min(workers, stats["n_batches"]))
            names = ["hapA", "hapB", "hapU"] if gz_out else ["hapA", "hapU"]
            for name in names:
                suffix = ext + (".gz" if gz_out else "")
                assert _content(os.path.join(got_dir, name + suffix), gz_out) == \
                    _content(os.path.join(want_dir, name + suffix), gz_out), (name, gz_out, encoder)


def test_empty_input_and_single_record(kmers, seq, tmp_path):
    tmp = str(tmp_path)
    for tag, data in (("empty", b""), ("one", b">only\nACGT\n"), ("no_newline", b"@q\nACGT\n+\n!!!!")):
        reads = os.path.join(tmp, f"{tag}.fa")
        open(reads, "wb").write(data)
        for gz_out in (True, False):
            want_dir, want_tsv = _python_writer(seq, reads, "B", os.path.join(tmp, f"{tag}{gz_out}"), gz_out, ".fa")
            got_dir, got_tsv, _ = _native(kmers, reads, "B", os.path.join(tmp, f"{tag}{gz_out}"), "native", gz_out, ".fa", workers=2)
            assert got_tsv == want_tsv
            for f in os.listdir(want_dir):
                assert _content(os.path.join(got_dir, f), gz_out) == _content(os.path.join(want_dir, f), gz_out), (tag, f)


def test_gpu_code_path_of_the_pipeline_with_stubbed_classifier(tmp_path):
    """The pipeline's real multi-GPU path -- packer, tb_classifier_submit_packed / wait with two
    slots per worker, replicas, on-demand and round-robin dealing -- linked against stand-ins for
    the classifier calls (tools/pipeline_sanitize.cpp; the same harness runs under TSan/ASan).
    Regression: workers that held their in-flight batches while waiting for input starved the
    reader of free batches and the run hung with 2+ workers (found with this harness)."""
    import hashlib
    import shutil
    import subprocess

    from conftest import ROOT

    if not shutil.which("g++"):
        pytest.skip("no g++")
    csrc = os.path.join(ROOT, "trio_binning_b200", "csrc")
    exe = str(tmp_path / "pipe_stub")
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", csrc, "-o", exe, os.path.join(ROOT, "tools", "pipeline_sanitize.cpp")] +
                   [os.path.join(csrc, f) for f in ("fastx_pipeline.cpp", "gz_stream.cpp", "fastq_deflate.cpp", "host_pack.cpp")] +
                   ["-lz", "-lpthread"], check=True)
    reads = str(tmp_path / "reads.fastq")
    _fastx(reads, [3000, 150, 9000, 1, 12000, 600] * 150, seed=3, fasta_every=11)
    digests = {}
    for workers in (1, 2, 3, 5, 8):          # odd counts deal round-robin, even ones on demand
        for mode in ("gz", "plain"):
            out = str(tmp_path / f"o_{workers}_{mode}")
            r = subprocess.run([exe, reads, out, "40000", "8", "gpu", str(workers), mode], capture_output=True, text=True,
                               timeout=120)
            assert r.returncode == 0 and "rc=0" in r.stdout, (workers, mode, r.stdout, r.stderr[-300:])
            h = hashlib.sha256(open(out + ".tsv", "rb").read())
            for suffix in (".A", ".B", ".U"):
                path = out + suffix + (".gz" if mode == "gz" else "")
                if os.path.exists(path):
                    raw = open(path, "rb").read()
                    h.update(gzip.decompress(raw) if mode == "gz" else raw)
            digests.setdefault(mode, set()).add(h.hexdigest())
    assert len(digests["gz"]) == 1 and len(digests["plain"]) == 1


def test_plain_writer_random_shapes_against_cpython(kmers, seq, tmp_path):
    """Randomised version of the plain-output test: record sizes drawn around the writers' thresholds
    (the text layer's 8192-byte chunk, the buffer's st_blksize, multiples of both), random bins."""
    import random

    rng = random.Random(20261017)
    marks = [1, 50, 150, 2040, 2048, 4090, 4096, 4100, 8180, 8192, 8200, 12288, 16384, 30000]
    for trial in range(8):
        lengths = [max(1, rng.choice(marks) + rng.randrange(-12, 13)) for _ in range(rng.randrange(5, 120))]
        bins = "".join(rng.choice("ABU" if trial % 2 else "AB") for _ in range(rng.randrange(1, 40)))
        tmp = str(tmp_path / f"t{trial}")
        os.makedirs(tmp)
        reads = os.path.join(tmp, "reads.fq")
        _fastx(reads, lengths, seed=trial, fasta_every=rng.choice([0, 3, 7]))
        want_dir, want_tsv = _python_writer(seq, reads, bins, tmp, False, ".fq")
        got_dir, got_tsv, _ = _native(kmers, reads, bins, tmp, "native", False, ".fq", batch_bases=rng.choice([1, 5000, 1 << 20]),
                                      workers=rng.choice([1, 2, 4]))
        for name in os.listdir(want_dir):
            assert _content(os.path.join(got_dir, name), False) == _content(os.path.join(want_dir, name), False), (trial, name, bins)
        assert got_tsv == want_tsv


@pytest.mark.parametrize("gz_out", [True, False])
def test_a_full_disk_is_an_error_with_its_reason(kmers, tmp_path, gz_out):
    """write() failing on an output file (ENOSPC on /dev/full) must end the run with an error that
    names the reason -- never a silently short output."""
    if not os.path.exists("/dev/full"):
        pytest.skip("no /dev/full")
    reads = str(tmp_path / "reads.fastq")
    _fastx(reads, [20_000] * 40, seed=1)
    with open(tmp_path / "out.tsv", "wb") as t:
        with pytest.raises(RuntimeError, match="No space left"):
            kmers.route_fastx_file(reads, "AB", "/dev/full", str(tmp_path / "b"), str(tmp_path / "u"), tsv_fd=t.fileno(),
                                   gzip_output=gz_out, batch_bases=100_000)
