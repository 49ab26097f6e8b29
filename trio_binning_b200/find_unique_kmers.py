#!/usr/bin/env python3
"""
find-unique-kmers -- Given multiple short-read libraries
find k-mers that are unique to each.
"""
# Host-side mirror of the reference's find-unique-kmers (reference
# src/trio_binning/find_unique_kmers.py): same command line (:25-59), same progress messages,
# same two output files (hapA_only_kmers.txt / hapB_only_kmers.txt, one k-mer per line).  The
# reference shells out to the KMC binaries for everything heavy; here those steps run on the
# GPU (csrc/kmer_count.cu):
#   kmc -k..                          (:62-103)   -> KmerCounter.add_file: canonical k-mer counting
#   kmc_tools transform .. histogram  (:123-130)  -> KmerCounter.histogram
#   the cut-off search                (:132-168)  -> analyze_histogram below (same loop)
#   kmc_tools simple .. kmers_subtract + kmc_dump -ci -cx (:171-233) -> KmerCounter.select_unique
# KMC's documented semantics are kept (canonical = lexicographically smaller strand, windows with
# non-ACGT skipped, counters saturate at 255, k-mers seen once dropped, dump in k-mer order).
# `classifier_sets()` hands the two unique sets to the classifier's set builder as integers,
# without the text files in between.

import argparse
import os
import sys
from typing import List, Optional, Tuple

import numpy as np

from . import kmers


class HistogramError(Exception):
    def __init__(self, histogram_path):
        self.message = (
            "Could not find min and max counts in histogram. "
            + "Try looking at the histogram yourself and manually "
            + "choosing cutoffs: {}".format(histogram_path)
        )
        super().__init__(self.message)


def parse_args(argv=None):
    parser = argparse.ArgumentParser(
        description="Given multiple short-read "
        "libraries, find k-mers that are unique to each library."
    )
    parser.add_argument("-k", "--kmer-size", type=int, required=True)
    parser.add_argument(
        "--path-to-kmc",
        default="kmc",
        help="accepted for compatibility with the reference; k-mers are counted on the GPU, kmc is not run",
    )
    parser.add_argument(
        "-p",
        "--threads",
        type=int,
        default=1,
        help="number of host threads for reading the input [1]",
    )
    parser.add_argument(
        "-o", "--outpath", default=".", help="Prefix to write " "output haplotypes to"
    )
    parser.add_argument(
        "-s",
        "--scratch-dir",
        default=".",
        help="Directory for the count histograms (the reference keeps its large temporary files here)",
    )
    parser.add_argument(
        "--table-gb", type=float, default=0.0,
        help="GPU memory for each of the two counting tables in GB, 0 = a quarter of the free memory "
             "(B200 addition); libraries with more distinct k-mers than fit are counted in several passes",
    )
    parser.add_argument(
        "--cutoffs", default=None,
        help="minA,maxA,minB,maxB: use these count ranges instead of searching the histograms (B200 addition; "
             "what the reference's HistogramError asks the user to do by hand)",
    )
    parser.add_argument(
        "read_files",
        nargs=2,
        help="one comma-separated "
        "list of file paths for both libraries being compared. Files can "
        "be in fasta or fastq format, and uncompressed or gzipped.",
    )
    return parser.parse_args(argv)


def histogram_lines(hist: np.ndarray) -> List[Tuple[int, int]]:
    """(count, number of k-mers) as `kmc_tools transform .. histogram` writes them: from 2 (k-mers
    seen once are not in a KMC database) to the saturated maximum."""
    return [(c, int(hist[c])) for c in range(kmers.KmerCounter.DB_MIN, len(hist))]


def analyze_histogram(lines, histogram_path: Optional[str] = None) -> Tuple[int, int]:
    """Choose minimum and maximum k-mer counts for finding unique k-mers: the first local
    minimum of the histogram, and the first count after it whose bin is lower than the bin at
    that minimum (reference find_unique_kmers.py:132-156, same loop over the same lines)."""
    min_coverage, max_coverage = False, False
    min_coverage_count = None
    last_count = -1
    for coverage, count in lines:
        if coverage != 2:  # don't do anything except record the first entry
            if not min_coverage:
                # looking for a local minimum, i.e., a place where count starts increasing
                if count > last_count:
                    min_coverage = coverage - 1
                    min_coverage_count = last_count
            elif not max_coverage:
                # looking for the place where the count dips below the count at min_coverage
                if count < min_coverage_count:
                    max_coverage = coverage
                    break
        last_count = count

    if not min_coverage or not max_coverage:
        raise HistogramError(histogram_path)

    if max_coverage - min_coverage < 5:
        print(
            "WARNING: min and max coverage not very far apart. This may be "
            "a result of coverage being too low. Try taking a look at "
            'the histogram in "{}" yourself.'.format(histogram_path),
            file=sys.stderr,
        )
    return min_coverage, max_coverage


def write_histogram(path: str, lines) -> None:
    with open(path, "w") as f:
        for coverage, count in lines:
            f.write(f"{coverage}\t{count}\n")


def keys_to_lsb_first(keys: np.ndarray, k: int) -> np.ndarray:
    """MSB-first canonical keys -> the classifier's `kmer_to_int` values (first base in the two
    least significant bits, c/kmers.c:50-72): the digits in reverse order."""
    keys = keys.astype(np.uint64)
    out = np.zeros_like(keys)
    for j in range(k):
        out |= ((keys >> np.uint64(2 * (k - 1 - j))) & np.uint64(3)) << np.uint64(2 * j)
    return out


def write_kmers(path: str, keys: np.ndarray, k: int, chunk: int = 1 << 22) -> int:
    """One k-mer per line, in KMC's dump order (the keys are sorted): find_unique_kmers.py:222-231."""
    letters = np.frombuffer(b"ACGT", dtype=np.uint8)
    with open(path, "wb") as f:
        for i in range(0, len(keys), chunk):
            part = keys[i:i + chunk]
            mat = np.empty((len(part), k + 1), dtype=np.uint8)
            for j in range(k):
                mat[:, j] = letters[((part >> np.uint64(2 * (k - 1 - j))) & np.uint64(3)).astype(np.intp)]
            mat[:, k] = 10
            f.write(mat.tobytes())
    return len(keys)


def _table_slots(table_gb: float, k: int) -> int:
    if table_gb <= 0:
        free = kmers.device_info(kmers.lib.tb_get_device())["hbm_bytes"]
        table_gb = free / 4 / 1e9
    return max(1 << 16, int(table_gb * 1e9 / kmers.lib.tb_counter_slot_bytes(k)))


def _count_pair(files_a, files_b, k, n_slots, part, n_parts, threads):
    """Both libraries' k-mers of one partition; OverflowError if a table gets too full."""
    pair = []
    try:
        for files in (files_a, files_b):
            counter = kmers.KmerCounter(k, n_slots, part, n_parts)
            pair.append(counter)
            for path in files:
                counter.add_file(path, threads=threads)
            st = counter.stats()  # raises OverflowError if the table filled up
            if st["distinct"] > 0.7 * st["slots"]:
                raise OverflowError("table more than 70% full")
    except Exception:
        for counter in pair:
            counter.close()
        raise
    return pair[0], pair[1]


def _run(files_a, files_b, k, n_slots, n_parts, threads, cutoffs, scratch_dir):
    n_bins = kmers.KmerCounter.COUNTER_MAX + 1
    if n_parts == 1:
        ca, cb = _count_pair(files_a, files_b, k, n_slots, 0, 1, threads)
        try:
            ranges = _ranges(ca.histogram(n_bins), cb.histogram(n_bins), cutoffs, scratch_dir)
            print("\033[92mFinding k-mers unique to haplotype A...\033[0m", file=sys.stderr)
            keys_a = ca.select_unique(cb, *ranges[0])
            print("\033[92mFinding k-mers unique to haplotype B...\033[0m", file=sys.stderr)
            keys_b = cb.select_unique(ca, *ranges[1])
        finally:
            ca.close()
            cb.close()
        return keys_a, keys_b, ranges[0], ranges[1]
    # several partitions (passes over the input): the cut-offs need the complete histograms
    # first, so every partition is counted twice
    hist_a = np.zeros(n_bins, dtype=np.uint64)
    hist_b = np.zeros(n_bins, dtype=np.uint64)
    for part in range(n_parts):
        ca, cb = _count_pair(files_a, files_b, k, n_slots, part, n_parts, threads)
        hist_a += ca.histogram(n_bins)
        hist_b += cb.histogram(n_bins)
        ca.close()
        cb.close()
    ranges = _ranges(hist_a, hist_b, cutoffs, scratch_dir)
    out_a, out_b = [], []
    for part in range(n_parts):
        ca, cb = _count_pair(files_a, files_b, k, n_slots, part, n_parts, threads)
        out_a.append(ca.select_unique(cb, *ranges[0]))
        out_b.append(cb.select_unique(ca, *ranges[1]))
        ca.close()
        cb.close()
    return np.sort(np.concatenate(out_a)), np.sort(np.concatenate(out_b)), ranges[0], ranges[1]


def find_unique_kmers(files_a: List[str], files_b: List[str], k: int, table_gb: float = 0.0, threads: int = 1,
                      cutoffs=None, scratch_dir: Optional[str] = None, n_parts: int = 1):
    """The whole computation.  Returns (keys_a_only, keys_b_only, (min_a, max_a), (min_b, max_b)) with
    sorted MSB-first canonical keys.  Libraries with more distinct k-mers than one table holds are
    counted in 2, 4, ... partitions (passes over the input)."""
    n_slots = _table_slots(table_gb, k)
    while True:
        try:
            return _run(files_a, files_b, k, n_slots, n_parts, threads, cutoffs, scratch_dir)
        except OverflowError:
            if n_parts >= 1024:
                raise
            n_parts *= 2
            print(f"k-mer table too small: counting again in {n_parts} passes", file=sys.stderr)


def _ranges(hist_a, hist_b, cutoffs, scratch_dir):
    if cutoffs is not None:
        return (cutoffs[0], cutoffs[1]), (cutoffs[2], cutoffs[3])
    out = []
    for hap, hist in (("A", hist_a), ("B", hist_b)):
        lines = histogram_lines(hist)
        path = None
        if scratch_dir is not None:
            path = os.path.join(scratch_dir, f"haplotype{hap}.histogram")
            write_histogram(path, lines)
        print("\033[92mComputing and analyzing histogram...\033[0m", file=sys.stderr)
        lo, hi = analyze_histogram(lines, path)
        print("\033[92mUsing counts in range [{},{}].\033[0m".format(lo, hi), file=sys.stderr)
        out.append((lo, hi))
    return out[0], out[1]


def main(argv=None):
    args = parse_args(argv)
    k = args.kmer_size
    files = [s.split(",") for s in args.read_files]
    for path in files[0] + files[1]:
        if not os.path.isfile(path):
            raise IOError(f"Specified file {path} does not exist or is not file.")
    cutoffs = None
    if args.cutoffs:
        cutoffs = [int(x) for x in args.cutoffs.split(",")]
        if len(cutoffs) != 4:
            raise SystemExit("--cutoffs wants minA,maxA,minB,maxB")
    print("\033[92mCounting k-mers in haplotypes A and B...\033[0m", file=sys.stderr)
    keys_a, keys_b, _, _ = find_unique_kmers(files[0], files[1], k, args.table_gb, args.threads, cutoffs, args.scratch_dir)
    print("\033[92mDumping k-mers unique to haplotype A...\033[0m", file=sys.stderr)
    hap_a_num_kmers = write_kmers(os.path.join(args.outpath, "hapA_only_kmers.txt"), keys_a, k)
    print("\033[92mDumping k-mers unique to haplotype B...\033[0m", file=sys.stderr)
    hap_b_num_kmers = write_kmers(os.path.join(args.outpath, "hapB_only_kmers.txt"), keys_b, k)
    print("\n\n\033[94m# of unique k-mers in haplotype A: {}\033[0m".format(hap_a_num_kmers), file=sys.stderr)
    print("\033[94m# of unique k-mers in haplotype B: {}\033[0m".format(hap_b_num_kmers), file=sys.stderr)


def classifier_sets(keys_a: np.ndarray, keys_b: np.ndarray, k: int):
    """The two unique sets as classifier hash sets, without writing and re-reading the text files:
    the same sets `create_kmer_hash_set` would build from hapA_only_kmers.txt / hapB_only_kmers.txt
    (num_kmers = number of lines, keys = kmer_to_int of each line, inserted raw)."""
    return (kmers.create_kmer_hash_set_from_ints(keys_to_lsb_first(keys_a, k), k),
            kmers.create_kmer_hash_set_from_ints(keys_to_lsb_first(keys_b, k), k))


if __name__ == "__main__":
    main()
