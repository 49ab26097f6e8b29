"""Host-side logic that needs neither a GPU nor the reference: batch packing helpers,
synthetic workload generators, bench.py's reference arm and its JSON contract."""
import json
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT


def test_pack_reads_utf8_and_nul_semantics():
    from trio_binning_b200 import kmers

    bases, offsets = kmers.pack_reads(["ACGT", "", "GG"])
    assert bytes(bases) == b"ACGTGG" and offsets.tolist() == [0, 4, 4, 6]
    # the reference encodes UTF-8 and the C side stops at the first NUL (kmers.py:147, c/kmers.c:285)
    bases, offsets = kmers.pack_reads(["AC\x00GT", "é", "T"])
    assert bytes(bases) == b"AC" + "é".encode() + b"T" and offsets.tolist() == [0, 2, 4, 5]
    bases, offsets = kmers.pack_reads([])
    assert len(bases) == 0 and offsets.tolist() == [0]


def test_synth_is_deterministic_and_well_formed(tmp_path):
    from trio_binning_b200 import synth

    a1, b1 = synth.random_kmer_sets(5000, 4000, 21, seed=3)
    a2, b2 = synth.random_kmer_sets(5000, 4000, 21, seed=3)
    assert (a1 == a2).all() and (b1 == b2).all()
    assert len(np.unique(np.concatenate([a1, b1]))) == 9000           # distinct and disjoint
    ca, _ = synth.random_kmer_sets(1000, 10, 21, seed=3, canonical=True)
    assert (ca <= synth.revcomp_ints(ca, 21)).all()                    # all findable (quirk Q1)
    bases, offsets, hap = synth.random_reads(50, 500, 21, a1, b1, seed=4, plant_every=100, invalid_rate=0.01)
    assert offsets[0] == 0 and offsets[-1] == len(bases) and (np.diff(offsets) >= 22).all()
    assert set(hap.tolist()) <= {0, 1, 2}
    for r in range(50):                                                # junk never inside the first k bases
        head = bytes(bases[offsets[r]:offsets[r] + 21])
        assert set(head) <= set(b"ACGT")
    p = tmp_path / "k.txt"
    synth.write_kmer_file(str(p), a1[:10], 21)
    lines = p.read_text().split("\n")
    assert len(lines) == 11 and lines[-1] == "" and all(len(x) == 21 for x in lines[:10])
    assert synth.total_windows(np.array([0, 10, 40, 61]), 21) == 0 + 10 + 1


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` on the small workload: one JSON line with the contract's keys."""
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload",
                        "c2_small", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "bases/s" and line["value"] > 1e5
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and "workload" in line["config"]
