"""Kernel-only timings of the classify path at a BASELINE config's set sizes (measurement aid).

    python tools/config_bench.py c3 [--reads N] [--iters I] [--env TB_X=1 ...] [--plant-every P]
    python tools/config_bench.py ksweep            # every k 13..31 at 10 M + 10 M

One JSON line per run: sizes, bitmap/pair info, classify-kernel ms (CUDA events, HBM-resident
batch), Gbases/s, bitmap pass rate.  Sets are synth.sorted_disjoint_sets (seconds even at
400 M + 400 M); reads are synth.random_reads with one planted k-mer per `plant_every` bases.
No oracle here: parity at these sizes is tests/test_gpu_large.py."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CONFIGS = {
    # name: (k, n_a, n_b, n_reads, mean_len, fixed)
    "c2": (21, 10_000_000, 10_000_000, 10_000, 20_000, False),
    "c3": (21, 150_000_000, 150_000_000, 10_000, 25_000, False),
    "c4": (31, 400_000_000, 400_000_000, 10_000, 25_000, False),
    "c5": (21, 50_000_000, 50_000_000, 2_000_000, 150, True),
}


def run(name, k, n_a, n_b, n_reads, mean_len, fixed, iters, plant_every, sets=None, label=None, repeat_reads=None):
    from trio_binning_b200 import kmers, synth

    t0 = time.time()
    ka, kb = sets if sets else synth.sorted_disjoint_sets(n_a, n_b, k, seed=1000 + k)
    if repeat_reads is not None:
        bases, offsets = repeat_reads
    else:
        bases, offsets, _ = synth.random_reads(n_reads, mean_len, k, ka, kb, seed=3,
                                               plant_every=plant_every, fixed_len=fixed)
    A = kmers.create_kmer_hash_set_from_ints(ka, k)
    B = kmers.create_kmer_hash_set_from_ints(kb, k)
    total = int(offsets[-1])
    clf = kmers.Classifier(A, B, total, len(offsets) - 1, 1)
    clf.submit(0, bases, offsets)
    ca, cb, *_ = (x.copy() for x in clf.result(0))
    st = clf.stats(0)
    fi = clf.filter_info()
    clf.rerun_resident(0, 3)
    ms, cls = clf.rerun_resident(0, iters)
    passes = clf.filter_passes(0)
    out = {"config": label or name, "k": k, "set_a": len(ka), "set_b": len(kb), "reads": len(offsets) - 1, "bases": total,
           "windows": st["windows"], "filter_log2": fi["log2_bits"], "bits_per_key": fi["bits_per_key"],
           "pair_gb": fi["pair_buckets"] * 32 / 1e9, "pass_fraction": passes / max(1, st["windows"]),
           "hit_fraction": float(ca.sum() + cb.sum()) / max(1, st["windows"]),
           "classify_ms": cls / iters, "step_ms": ms / iters,
           "gbases_per_s_kernel": total / (cls / iters * 1e-3) / 1e9,
           "gbases_per_s_step": total / (ms / iters * 1e-3) / 1e9,
           "env": {k_: v for k_, v in os.environ.items() if k_.startswith("TB_")},
           "setup_s": round(time.time() - t0, 1)}
    print(json.dumps(out), flush=True)
    clf.close()
    kmers.destroy_kmer_hash_set(A)
    kmers.destroy_kmer_hash_set(B)
    return ka, kb


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what")
    ap.add_argument("--reads", type=int, default=0)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--plant-every", type=int, default=0)
    ap.add_argument("--env", nargs="*", default=[])
    args = ap.parse_args()
    for kv in args.env:
        key, _, val = kv.partition("=")
        os.environ[key] = val
    if args.what == "ksweep":
        from trio_binning_b200 import synth

        for k in range(13, 32):
            n = min(10_000_000, 4 ** k // 16)  # k = 13, 14: the k-mer space is too small for 10 M + 10 M
            run("ksweep", k, n, n, args.reads or 10_000, 20_000, False, args.iters,
                args.plant_every or 1000, label=f"k{k}")
        return
    if args.what == "density":
        # throughput against hit density: the sets are ALL k-mers of two 10-Mbase random genomes
        # (canonical orientation: every one findable), the reads carry pieces of their
        # haplotype's genome over a chosen fraction of their length
        from trio_binning_b200 import synth

        k, n_a, n_b, n_reads, mean_len, fixed = CONFIGS["c2"]
        rng = np.random.default_rng(77)
        letters = np.frombuffer(b"ACGT", dtype=np.uint8)
        ga, gb = letters[rng.integers(0, 4, size=n_a + k - 1)], letters[rng.integers(0, 4, size=n_b + k - 1)]
        sets = (synth.genome_kmer_set(ga, k), synth.genome_kmer_set(gb, k))
        for f in (0.0, 0.001, 0.01, 0.1, 0.5, 1.0):
            reads = synth.mosaic_reads(args.reads or n_reads, mean_len, k, ga, gb, seed=5, fraction=f)[:2]
            run("c2", k, n_a, n_b, args.reads or n_reads, mean_len, fixed, args.iters, 0, sets=sets,
                label=f"c2_genome_sets_fraction_{f}", repeat_reads=reads)
        # low-complexity reads: short tandem repeats (period 1..6) -- few distinct minimizers,
        # every window of a repeat maps to the same few bitmap sectors
        rng = np.random.default_rng(5)
        n = args.reads or n_reads
        lens = synth.read_lengths(n, mean_len, rng)
        offsets = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(lens, out=offsets[1:])
        bases = np.empty(int(offsets[-1]), dtype=np.uint8)
        letters = np.frombuffer(b"ACGT", dtype=np.uint8)
        for r in range(n):
            unit = letters[rng.integers(0, 4, size=int(rng.integers(1, 7)))]
            reps = int(lens[r]) // len(unit) + 1
            bases[offsets[r]:offsets[r + 1]] = np.tile(unit, reps)[: int(lens[r])]
        run("c2", k, n_a, n_b, n, mean_len, fixed, args.iters, 0, sets=sets, label="c2_tandem_repeats",
            repeat_reads=(bases, offsets))
        return
    k, n_a, n_b, n_reads, mean_len, fixed = CONFIGS[args.what]
    run(args.what, k, n_a, n_b, args.reads or n_reads, mean_len, fixed, args.iters,
        args.plant_every or (100 if fixed else 1000))


if __name__ == "__main__":
    main()
