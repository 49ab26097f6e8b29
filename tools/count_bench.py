"""Throughput of the find-unique-kmers compute on one GPU: two haplotype read
libraries counted into two tables, both histograms, the cut-offs, the two subtractions.

The workload is what the reference feeds KMC (find_unique_kmers.py:62-103): short reads at a
coverage where true k-mers repeat and sequencing errors make most DISTINCT k-mers singletons.
  genome      G random bases; haplotype B = A with one substitution per 200 bases
  libraries   150-base reads, uniform starts, random strand, 0.5 % substitutions, COV x each
The reads are generated in host memory (numpy) and handed to KmerCounter.add_packed in batches,
so the timed region holds the H2D copy of the ASCII, k_pack and k_count_windows; timed once from
pageable numpy arrays (the runtime stages them at 6-7 GB/s) and once from page-locked memory.

  python tools/count_bench.py [--genome 20000000] [--coverage 30] [--k 21] [--cpu-sample 2000000]

Prints one JSON line; the CPU figure beside it is the numpy restatement in oracle/ (a "port":
KMC itself is not in this image) on a bounded sample, and its result is compared with the GPU's
on that sample.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
COMP = np.zeros(256, dtype=np.uint8)
CODE = np.zeros(256, dtype=np.int64)
for i, (a, b) in enumerate(zip(b"ACGT", b"TGCA")):
    COMP[a] = b
    CODE[a] = i


def library(genome: np.ndarray, coverage: float, read_len: int, err: float, rng):
    """(bases uint8 ASCII, offsets int64) of one read library."""
    n_reads = int(len(genome) * coverage / read_len)
    starts = rng.integers(0, len(genome) - read_len, n_reads, dtype=np.int64)
    bases = np.empty(n_reads * read_len, dtype=np.uint8)
    step = 1 << 18
    for lo in range(0, n_reads, step):
        s = starts[lo:lo + step]
        block = genome[s[:, None] + np.arange(read_len, dtype=np.int64)[None, :]]
        flip = rng.random(len(s)) < 0.5
        block[flip] = COMP[block[flip, ::-1]]
        bases[lo * read_len:(lo + len(s)) * read_len] = block.reshape(-1)
    n_err = int(len(bases) * err)
    where = rng.integers(0, len(bases), n_err, dtype=np.int64)
    bases[where] = ACGT[rng.integers(0, 4, n_err)]
    offsets = np.arange(n_reads + 1, dtype=np.int64) * read_len
    return bases, offsets


def add_in_batches(counter, bases, offsets, batch_reads):
    n = len(offsets) - 1
    for lo in range(0, n, batch_reads):
        hi = min(n, lo + batch_reads)
        counter.add_packed(bases[offsets[lo]:offsets[hi]], offsets[lo:hi + 1] - offsets[lo])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome", type=int, default=20_000_000)
    ap.add_argument("--coverage", type=float, default=30.0)
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--k", type=int, default=21)
    ap.add_argument("--error-rate", type=float, default=0.005)
    ap.add_argument("--batch-bases", type=int, default=256 << 20)
    ap.add_argument("--cpu-sample", type=int, default=2_000_000, help="bases of library A the numpy port counts")
    ap.add_argument("--seed", type=int, default=11)
    args = ap.parse_args()

    from trio_binning_b200 import find_unique_kmers as fuk
    from trio_binning_b200 import kmers

    rng = np.random.default_rng(args.seed)
    genome_a = ACGT[rng.integers(0, 4, args.genome)]
    genome_b = genome_a.copy()
    snp = rng.integers(0, args.genome, args.genome // 200)
    genome_b[snp] = ACGT[(rng.integers(1, 4, len(snp)) + CODE[genome_b[snp]]) % 4]
    t0 = time.perf_counter()
    lib_a = library(genome_a, args.coverage, args.read_len, args.error_rate, rng)
    lib_b = library(genome_b, args.coverage, args.read_len, args.error_rate, rng)
    t_gen = time.perf_counter() - t0
    n_bases = len(lib_a[0]) + len(lib_b[0])
    windows_per_read = args.read_len - args.k + 1
    n_windows = (len(lib_a[1]) - 1 + len(lib_b[1]) - 1) * windows_per_read

    # distinct k-mers per library: the genome's plus up to k new ones per error
    expect = args.genome + int(len(lib_a[0]) * args.error_rate * args.k)
    n_slots = int(expect * 1.6)
    batch_reads = max(1, args.batch_bases // args.read_len)

    def run(lib_a, lib_b):
        t = {}
        t0 = time.perf_counter()
        ca = kmers.KmerCounter(args.k, n_slots)
        cb = kmers.KmerCounter(args.k, n_slots)
        t["create_tables"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        add_in_batches(ca, *lib_a, batch_reads)
        add_in_batches(cb, *lib_b, batch_reads)
        sa, sb = ca.stats(), cb.stats()
        t["count"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        ha, hb = ca.histogram(256), cb.histogram(256)
        t["histograms"] = time.perf_counter() - t0
        cut_a, cut_b = fuk.analyze_histogram(fuk.histogram_lines(ha)), fuk.analyze_histogram(fuk.histogram_lines(hb))
        t0 = time.perf_counter()
        ua = ca.select_unique(cb, *cut_a)
        ub = cb.select_unique(ca, *cut_b)
        t["subtract_select_sort"] = time.perf_counter() - t0
        ca.close()
        cb.close()
        return t, sa, sb, cut_a, cut_b, len(ua), len(ub)

    run(lib_a, lib_b)  # warm-up: context, allocations, first-touch of the host arrays
    t_pageable = run(lib_a, lib_b)[0]
    # the same reads in page-locked memory: the copy runs at PCIe speed
    pins = [kmers.PinnedBuffer(len(lib[0]), np.uint8) for lib in (lib_a, lib_b)]
    for pin, lib in zip(pins, (lib_a, lib_b)):
        pin.array[:] = lib[0]
    pin_a, pin_b = (pins[0].array, lib_a[1]), (pins[1].array, lib_b[1])
    run(pin_a, pin_b)
    t, sa, sb, cut_a, cut_b, n_ua, n_ub = run(pin_a, pin_b)

    # bounded CPU sample, checked against the GPU on the same reads
    from oracle import find_unique_oracle as orc

    n_sample_reads = max(1, args.cpu_sample // args.read_len)
    sample = lib_a[0][:n_sample_reads * args.read_len].reshape(n_sample_reads, args.read_len)
    t0 = time.perf_counter()
    keys, counts = orc.count_library((row.tobytes() for row in sample), args.k)
    t_cpu = time.perf_counter() - t0
    with kmers.KmerCounter(args.k, max(1 << 20, 2 * len(keys))) as c:
        c.add_packed(sample.reshape(-1), np.arange(n_sample_reads + 1, dtype=np.int64) * args.read_len)
        st = c.stats()
        hist = c.histogram(256)
    want = np.bincount(np.minimum(counts, 255), minlength=256)
    sample_ok = st["distinct"] == len(keys) and st["windows"] == int(counts.sum()) and np.array_equal(hist[1:], want[1:])

    out = {
        "metric": "find_unique_kmers_read_bases_counted_per_second",
        "value": n_bases / t["count"],
        "unit": "bases/s",
        "k": args.k,
        "workload": f"2 libraries x {args.coverage:g}x of a {args.genome / 1e6:g}-Mbase genome, {args.read_len}-base reads, "
                    f"{args.error_rate * 100:g} % errors",
        "n_bases": n_bases,
        "n_windows": n_windows,
        "windows_counted": sa["windows"] + sb["windows"],
        "distinct": [sa["distinct"], sb["distinct"]],
        "table_slots": n_slots,
        "table_gb": n_slots * 12 / 1e9,
        "seconds": t,
        "cutoffs": [list(cut_a), list(cut_b)],
        "unique_kmers": [n_ua, n_ub],
        "input": "ASCII in page-locked host memory, H2D inside the timed region",
        "value_pageable_input": n_bases / t_pageable["count"],
        "seconds_pageable_input": t_pageable,
        "cpu_port": {"value": sample.size / t_cpu, "unit": "bases/s", "cores": 1, "kind": "port",
                     "sample": f"{sample.size} bases of library A (numpy restatement, oracle/find_unique_oracle.py)",
                     "gpu_equals_port_on_sample": bool(sample_ok)},
        "seconds_generate_reads": t_gen,
    }
    print(json.dumps(out))
    return 0 if sample_ok else 1


if __name__ == "__main__":
    sys.exit(main())
