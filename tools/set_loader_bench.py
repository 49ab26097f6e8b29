"""Set loader at scale: text file (one k-mer per line) -> GPU table, wall clock.
usage: set_loader_bench.py [n_kmers] [k]"""
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trio_binning_b200 import kmers, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 150_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 21
rng = np.random.default_rng(3)
keys = rng.integers(0, 1 << (2 * k), size=n, dtype=np.uint64)
tmp = tempfile.mkdtemp(prefix="tb_sets_")
path = os.path.join(tmp, "hap.txt")
t0 = time.time()
synth.write_kmer_file(path, keys, k)
print(f"wrote {os.path.getsize(path) / 1e9:.2f} GB of text in {time.time() - t0:.1f} s", flush=True)
kmers.set_device(0)
for attempt in ("cold-ish (page cache from the write)", "warm"):
    t0 = time.time()
    s = kmers.create_kmer_hash_set(path)
    dt = time.time() - t0
    print(f"{attempt}: {n} lines -> {s.contents.n_distinct} distinct keys in {dt:.2f} s wall "
          f"({n / dt / 1e6:.1f} M k-mers/s; device timeline {s.contents.build_ms:.0f} ms; "
          f"reference: ~2.2 M k-mers/s => {n / 2.2e6:.0f} s)", flush=True)
    kmers.destroy_kmer_hash_set(s)
t0 = time.time()
s = kmers.create_kmer_hash_set_from_ints(keys, k)
print(f"from pre-encoded keys: {time.time() - t0:.2f} s wall", flush=True)
os.remove(path)
