"""End-to-end CLI timing on a synthetic FASTQ(.gz): native pipeline vs Python I/O
(run on the GPU box).  usage: pipeline_bench.py [n_reads] [mean_len] [only] [n_keys]
`only`: comma-separated substrings of the run labels to keep (default: all five runs)"""
import gzip
import io
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from trio_binning_b200 import synth  # noqa: E402

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
mean_len = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
only = [x for x in (sys.argv[3].split(",") if len(sys.argv) > 3 else []) if x]
n_keys = int(sys.argv[4]) if len(sys.argv) > 4 else 10_000_000
k = 21
tmp = tempfile.mkdtemp(prefix="tb_pipe_")
ka, kb, pa, pb = synth.materialise_sets(tmp, n_keys, n_keys, k, seed=42)
bases, offsets, _ = synth.random_reads(n_reads, mean_len, k, ka, kb, seed=43, plant_every=1000)
rng = np.random.default_rng(1)
t0 = time.time()
path_plain = os.path.join(tmp, "reads.fastq")
with open(path_plain, "wb") as f:
    for r in range(n_reads):
        s = bases[offsets[r]:offsets[r + 1]]
        q = rng.integers(33, 74, size=len(s), dtype=np.uint8)  # random qualities: worst case for gzip
        f.write(b"@read%d\n" % r + s.tobytes() + b"\n+\n" + q.tobytes() + b"\n")
path_gz = path_plain + ".gz"
subprocess.run(f"gzip -1 -k -f {path_plain}", shell=True, check=True)
total = int(offsets[-1])
print(f"input: {n_reads} reads, {total/1e6:.0f} Mbases, fastq {os.path.getsize(path_plain)/1e6:.0f} MB, "
      f"gz {os.path.getsize(path_gz)/1e6:.0f} MB (made in {time.time()-t0:.0f}s)", flush=True)
for label, reads, extra in (("native gz-in level9", path_gz, ["--gzip-level", "9"]),
                            ("native gz-in level6", path_gz, ["--gzip-level", "6"]),
                            ("native gz-in level1", path_gz, ["--gzip-level", "1"]),
                            ("native plain-in level1", path_plain, ["--gzip-level", "1"]),
                            ("python-io gz-in", path_gz, ["--python-io"])):
    if only and not any(x in label for x in only):
        continue
    out = os.path.join(tmp, label.replace(" ", "_"))
    os.makedirs(out, exist_ok=True)
    cmd = [sys.executable, "-m", "trio_binning_b200.classify_by_kmers", reads, pa, pb,
           "--haplotype-a-out-prefix", os.path.join(out, "A"), "--haplotype-b-out-prefix", os.path.join(out, "B"),
           "--unclassified-out-prefix", os.path.join(out, "U")] + extra
    t0 = time.time()
    p = subprocess.run(cmd, cwd=ROOT, stdout=open(os.path.join(out, "tsv"), "w"), stderr=subprocess.PIPE, text=True)
    dt = time.time() - t0
    if p.returncode != 0:
        print(label, "FAILED", p.stderr[-400:])
        continue
    sizes = sum(os.path.getsize(os.path.join(out, f)) for f in os.listdir(out) if f.endswith(".gz"))
    print(f"{label:>24}: {dt:6.1f} s wall (incl. ~5 s python/CUDA start-up and set build), "
          f"{total/dt/1e6:8.1f} Mbases/s, outputs {sizes/1e6:.0f} MB", flush=True)
# outputs identical?
import hashlib
ref = None
for d in sorted(os.listdir(tmp)):
    full = os.path.join(tmp, d)
    if os.path.isdir(full) and os.path.exists(os.path.join(full, "tsv")):
        h = hashlib.sha256(open(os.path.join(full, "tsv"), "rb").read())
        for n in ("A.fastq.gz", "B.fastq.gz", "U.fastq.gz"):
            h.update(gzip.decompress(open(os.path.join(full, n), "rb").read()))
        print(d, h.hexdigest()[:16])
