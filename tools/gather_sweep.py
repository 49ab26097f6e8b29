"""Sweep of the random-32-B-sector gather microbenchmark (run on the GPU box)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trio_binning_b200 import kmers  # noqa: E402

print("device", kmers.device_info(0))
print(f"{'table':>10} {'variant':>7} {'mlp':>4} {'tpb':>4} {'bps':>4} {'Gsect/s':>9} {'GB/s':>8} {'ms':>8}")
for table_mb in (64, 320, 4096, 16384):
    for variant in (0, 1, 2):
        for mlp, tpb, bps in ((1, 256, 8), (2, 256, 8), (4, 256, 8), (8, 256, 8), (16, 256, 4),
                              (8, 128, 16), (8, 512, 4), (4, 1024, 2), (16, 128, 8)):
            if variant != 0 and (mlp, tpb, bps) not in ((8, 256, 8), (16, 256, 4)):
                continue
            rate, ms = kmers.gather_bench(table_mb << 20, mlp, tpb, bps, 64, variant)
            print(f"{table_mb:>8}MB {variant:>7} {mlp:>4} {tpb:>4} {bps:>4} {rate / 1e9:>9.2f} {rate * 32 / 1e9:>8.1f} {ms:>8.3f}",
                  flush=True)
