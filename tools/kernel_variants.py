"""Classify-kernel timing under different knobs / shapes (run on the GPU box).
usage: kernel_variants.py <n_keys_per_set> <n_reads> <mean_len> [fixed] [load_factor]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trio_binning_b200 import kmers, synth  # noqa: E402

n_keys, n_reads, mean_len = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
fixed = len(sys.argv) > 4 and sys.argv[4] == "fixed"
lf = float(sys.argv[5]) if len(sys.argv) > 5 else 0.5
k = int(os.environ.get("K", "21"))
t0 = time.time()
rng = np.random.default_rng(1)
# distinctness does not matter for timing: plain random keys (duplicates collapse)
ka = rng.integers(0, 1 << (2 * k), size=n_keys, dtype=np.uint64)
kb = rng.integers(0, 1 << (2 * k), size=n_keys, dtype=np.uint64)
bases, offsets, _ = synth.random_reads(n_reads, mean_len, k, ka[:100000], kb[:100000], seed=2, plant_every=1000, fixed_len=fixed)
kmers.set_load_factor(lf)
A = kmers.create_kmer_hash_set_from_ints(ka, k)
B = kmers.create_kmer_hash_set_from_ints(kb, k)
total = int(offsets[-1])
clf = kmers.Classifier(A, B, total, n_reads, 1)
clf.submit(0, bases, offsets)
ca, cb, *_ = clf.result(0)
st = clf.stats(0)
clf.rerun_resident(0, 3)
ms, cls = clf.rerun_resident(0, 5)
fi = clf.filter_info()
print(f"filter=2^{fi['log2_bits']} pair={int(fi['pair_buckets']>0)} mini={int(fi['minimizer'])} bits={fi['bits_per_key']} passes={clf.filter_passes(0)/max(1,st['windows']):.4f} FU={os.environ.get('TB_FILTER_UNROLL','16')} keys={n_keys} lf={lf} buckets={A.contents.n_buckets} table_MB={A.contents.n_buckets*32/1e6:.0f}x2 "
      f"reads={n_reads}x{mean_len} k={k} UNROLL={os.environ.get('TB_UNROLL','8')} TPB={os.environ.get('TB_TPB','128')} "
      f"L2G={os.environ.get('TB_L2_FETCH_GRANULARITY','32')}: step {ms/5:.3f} ms, classify {cls/5:.3f} ms, "
      f"{total/(cls/5*1e-3)/1e9:.2f} Gbases/s, {st['probes']/(cls/5*1e-3)/1e9:.2f} Gprobes/s "
      f"(build {A.contents.build_ms:.0f}+{B.contents.build_ms:.0f} ms, setup {time.time()-t0:.0f}s) "
      f"counts={int(ca.sum())}/{int(cb.sum())}", flush=True)
