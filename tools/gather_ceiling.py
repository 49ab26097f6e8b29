"""Random-32-B-sector request ceilings of this GPU (the denominator north_star names): the best rate
over load flavours (256-bit no-allocate bucket load, the same with L2::64B, the classify kernel's
32-bit L1-allocating bitmap load), loads in flight and launch shapes, for tables of several sizes.
usage: gather_ceiling.py [table_mb ...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trio_binning_b200 import kmers  # noqa: E402

sizes = [int(x) for x in sys.argv[1:]] or [32, 1024, 4096, 16384]
for mb in sizes:
    best = {}
    for variant, name in ((0, "v4.u64 no_allocate"), (7, "v4.u64 no_allocate L2::64B"), (13, "u32 L1-allocating evict_last")):
        rates = {}
        for mlp, tpb, bps in ((4, 256, 8), (8, 256, 8), (16, 256, 4), (8, 128, 16), (16, 128, 8), (8, 512, 4), (16, 512, 2)):
            try:
                rate, ms = kmers.gather_bench(mb << 20, mlp, tpb, bps, 48 if mb > 256 else 256, variant)
            except RuntimeError:
                continue
            rates[f"mlp{mlp}_tpb{tpb}_bps{bps}"] = round(rate / 1e9, 2)
        top = max(rates, key=rates.get)
        best[name] = {"best_G_per_s": rates[top], "at": top, "all": rates}
    print(json.dumps({"table_mb": mb, "ceilings": best}), flush=True)
