"""BASELINE.json configs[2..4] at full SET sizes on one GPU (reads: a slice of the stated shape).
Prints one JSON line per config: sizes, filter/pair info, kernel time, throughput, parity check.
usage: python tests/run_large_configs.py [c3] [c4] [c5]
(kept under tests/ because it checks against the oracle; not collected by pytest: it needs
minutes of GPU-box time and tens of GB of host memory)"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.oracle import Oracle  # noqa: E402  (checker only)
from trio_binning_b200 import kmers, synth  # noqa: E402

CONFIGS = {
    # name: (k, n_a, n_b, n_reads, mean_len, fixed, oracle_reads)
    "c3": (21, 150_000_000, 150_000_000, 10_000, 25_000, False, 60),
    "c4": (31, 400_000_000, 400_000_000, 10_000, 25_000, False, 0),
    "c5": (21, 50_000_000, 50_000_000, 2_000_000, 150, True, 20_000),
}


def run(name):
    k, n_a, n_b, n_reads, mean_len, fixed, oracle_reads = CONFIGS[name]
    t0 = time.time()
    rng = np.random.default_rng(7)
    hi = 1 << (2 * k)
    ka = rng.integers(0, hi, size=n_a, dtype=np.uint64)
    kb = rng.integers(0, hi, size=n_b, dtype=np.uint64)
    bases, offsets, hap = synth.random_reads(n_reads, mean_len, k, ka[:200_000], kb[:200_000], seed=3,
                                             plant_every=1000 if not fixed else 100, fixed_len=fixed)
    t_gen = time.time() - t0
    A = kmers.create_kmer_hash_set_from_ints(ka, k)
    B = kmers.create_kmer_hash_set_from_ints(kb, k)
    total = int(offsets[-1])
    clf = kmers.Classifier(A, B, total, n_reads, 2)
    clf.submit(0, bases, offsets)
    ca, cb, sa, sb, bins = (x.copy() for x in clf.result(0))
    st = clf.stats(0)
    fi = clf.filter_info()
    clf.rerun_resident(0, 3)
    ms, cls = clf.rerun_resident(0, 10)
    passes = clf.filter_passes(0)
    out = {
        "config": name, "k": k, "set_a": n_a, "set_b": n_b, "distinct_a": int(A.contents.n_distinct),
        "distinct_b": int(B.contents.n_distinct), "table_gb_each": A.contents.n_buckets * 32 / 1e9,
        "pair_table_gb": fi["pair_buckets"] * 32 / 1e9, "filter_log2": fi["log2_bits"], "bits_per_key": fi["bits_per_key"],
        "set_build_ms": [A.contents.build_ms, B.contents.build_ms], "pair_filter_build_ms": fi["build_ms"],
        "reads": n_reads, "bases": total, "windows": st["windows"], "filter_pass_fraction": passes / max(1, st["windows"]),
        "step_ms": ms / 10, "classify_ms": cls / 10, "gbases_per_s_kernel": total / (cls / 10 * 1e-3) / 1e9,
        "gbases_per_s_step": total / (ms / 10 * 1e-3) / 1e9, "ref_probes_per_s": st["probes"] / (cls / 10 * 1e-3),
        "bins": {"A": int((bins == 65).sum()), "B": int((bins == 66).sum()), "U": int((bins == 85).sum())},
        "hits": [int(ca.sum()), int(cb.sum())], "gen_s": t_gen,
    }
    # parity
    if oracle_reads:
        t1 = time.time()
        orc = Oracle()
        oa, ob = orc.set_from_ints(ka, k), orc.set_from_ints(kb, k)
        step = max(1, n_reads // oracle_reads)
        idx = list(range(0, n_reads, step))
        bad = 0
        for r in idx:
            if orc.count(bytes(bases[offsets[r]:offsets[r + 1]]), oa, ob) != (ca[r], cb[r]):
                bad += 1
        out["parity"] = {"method": f"oracle (full-size sets) on every {step}th read", "reads_checked": len(idx),
                         "mismatches": bad, "seconds": time.time() - t1}
        assert bad == 0
    else:
        # size-independent property: reverse-complementing every read leaves the counts unchanged
        comp = np.zeros(256, dtype=np.uint8)
        comp[np.frombuffer(b"ACGT", dtype=np.uint8)] = np.frombuffer(b"TGCA", dtype=np.uint8)
        rc = np.empty_like(bases)
        for r in range(n_reads):
            rc[offsets[r]:offsets[r + 1]] = comp[bases[offsets[r]:offsets[r + 1]][::-1]]
        clf.submit(1, rc, offsets)
        ra, rb, *_ = clf.result(1)
        out["parity"] = {"method": "reverse-complement symmetry of all reads + planted haplotype wins",
                         "mismatches": int(((ra != ca) | (rb != cb)).sum())}
        assert out["parity"]["mismatches"] == 0
    # phase-2 width: the other setting must give the same counts; record its kernel time
    clf.close()
    other = 4 if os.environ.get("TB_P2_WIDTH", "0") not in ("4",) else 2
    prev = os.environ.get("TB_P2_WIDTH")
    os.environ["TB_P2_WIDTH"] = str(other)
    clf = kmers.Classifier(A, B, total, n_reads, 1)
    clf.submit(0, bases, offsets)
    wa, wb, *_ = clf.result(0)
    assert (wa == ca).all() and (wb == cb).all(), "phase-2 width changed the counts"
    clf.rerun_resident(0, 3)
    _, cls_w = clf.rerun_resident(0, 10)
    out[f"classify_ms_p2_width_{other}"] = cls_w / 10
    if prev is None:
        del os.environ["TB_P2_WIDTH"]
    else:
        os.environ["TB_P2_WIDTH"] = prev
    planted_a = float((bins[hap == 0] == 65).mean())
    planted_b = float((bins[hap == 1] == 66).mean())
    out["planted_called"] = [planted_a, planted_b]
    print(json.dumps(out), flush=True)
    clf.close()
    kmers.destroy_kmer_hash_set(A)
    kmers.destroy_kmer_hash_set(B)


if __name__ == "__main__":
    for name in (sys.argv[1:] or ["c3", "c4", "c5"]):
        run(name)
