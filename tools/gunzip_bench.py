"""Serial against parallel decoding of one gzip stream through the native reader (CPU only).

  python tools/gunzip_bench.py [reads.fastq.gz] [threads ...]

Without a file, 250 Mbases of synthetic long-read FASTQ (25-kb reads, 41 quality values) are written
and compressed with gzip -6 first.  Prints one JSON line per setting: seconds and MB/s of text
delivered to the parser (tb_fastx_cat: lines only, no records, nothing written)."""
import gzip
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from trio_binning_b200 import kmers  # noqa: E402


def synthetic(path):
    rng = np.random.default_rng(1)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    plain = path[:-3]
    with open(plain, "wb") as f:
        for r in range(10_000):
            n = int(rng.integers(15_000, 35_000))
            f.write(b"@read%d\n" % r + acgt[rng.integers(0, 4, n)].tobytes() + b"\n+\n" +
                    rng.integers(33, 74, n, dtype=np.uint8).tobytes() + b"\n")
    subprocess.run(["gzip", "-6", "-f", plain], check=True)


def main():
    args = sys.argv[1:]
    path = args[0] if args and not args[0].isdigit() else None
    threads = [int(a) for a in args if a.isdigit()] or [2, 4, 8]
    if path is None:
        path = os.path.join(tempfile.mkdtemp(prefix="tb_gunzip_"), "reads.fastq.gz")
        synthetic(path)
    size = os.path.getsize(path)
    with open(os.devnull, "wb") as sink:
        def run(label, n_threads):
            best = None
            for _ in range(3):
                t0 = time.perf_counter()
                n = kmers.lib.tb_fastx_cat(os.fsencode(path), sink.fileno(), n_threads)
                dt = time.perf_counter() - t0
                if n < 0:
                    raise SystemExit(f"tb_fastx_cat failed ({n})")
                best = dt if best is None else min(best, dt)
            print(json.dumps({"reader": label, "threads": n_threads, "seconds": round(best, 3), "text_mb_per_s": round(n / best / 1e6),
                              "compressed_mb": round(size / 1e6), "text_mb": round(n / 1e6), "host_cores": os.cpu_count()}), flush=True)
        os.environ["TB_PGZIP"] = "0"
        run("serial", 1)
        os.environ["TB_PGZIP"] = "1"
        for t in threads:
            run("spans", t)


if __name__ == "__main__":
    main()
