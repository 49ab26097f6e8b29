"""Sustained end-to-end run at BASELINE configs[2] read volume: 25 Gbases of ASCII reads
(128 batches of ~0.2 Gbases, 4 distinct batches in rotation) through the double-buffered
classifier against C2-size sets; every repetition of a batch must reproduce its first result.
usage: sustained_run.py [n_rounds]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trio_binning_b200 import kmers, synth  # noqa: E402

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 32
k = 21
ka, kb = synth.random_kmer_sets(10_000_000, 10_000_000, k, seed=42)
A = kmers.create_kmer_hash_set_from_ints(ka, k)
B = kmers.create_kmer_hash_set_from_ints(kb, k)
batches = [synth.random_reads(10_000, 20_000, k, ka, kb, seed=100 + i, plant_every=1000)[:2] for i in range(4)]
max_bases = max(int(o[-1]) for _, o in batches)
clf = kmers.Classifier(A, B, max_bases, 10_000, 2)
first = {}
t0 = time.perf_counter()
done_bases = 0
pending = {}
n = 0
for r in range(rounds):
    for i, (bases, offsets) in enumerate(batches):
        slot = n & 1
        if slot in pending:
            j = pending.pop(slot)
            ca, cb, *_ = clf.result(slot)
            key = (ca.sum(), cb.sum(), int(ca[::97].sum()))
            assert first.setdefault(j, key) == key, f"batch {j} changed its result"
        clf.submit(slot, bases, offsets)      # includes the copy into the pinned staging buffers
        pending[slot] = i
        done_bases += int(offsets[-1])
        n += 1
for slot, j in pending.items():
    ca, cb, *_ = clf.result(slot)
    key = (ca.sum(), cb.sum(), int(ca[::97].sum()))
    assert first.setdefault(j, key) == key
dt = time.perf_counter() - t0
print(f"{n} batches, {done_bases / 1e9:.1f} Gbases in {dt:.2f} s wall = {done_bases / dt / 1e9:.1f} Gbases/s "
      f"(host copy into pinned staging included); results stable across {rounds} repetitions: {first}")
