"""Sharded CLI check (run on a multi-GPU box): the same FASTQ through
  (a) one GPU, native pipeline   (b) one GPU, Python I/O   (c) torchrun N ranks, one GPU each
must give identical TSV and identical decompressed outputs.  usage: cli_multi_gpu.py [N]"""
import gzip
import hashlib
import os
import random
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from trio_binning_b200 import synth  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2
k = 21
tmp = tempfile.mkdtemp(prefix="tb_cli_")
ka, kb = synth.random_kmer_sets(200_000, 150_000, k, seed=5)
pa, pb = os.path.join(tmp, "hapA.txt"), os.path.join(tmp, "hapB.txt")
synth.write_kmer_file(pa, ka, k)
synth.write_kmer_file(pb, kb, k)
bases, offsets, _ = synth.random_reads(3000, 8000, k, ka, kb, seed=6, plant_every=500, invalid_rate=0.0005)
rng = random.Random(1)
reads = os.path.join(tmp, "reads.fastq.gz")
with gzip.open(reads, "wb", compresslevel=1) as f:
    for r in range(len(offsets) - 1):
        s = bytes(bases[offsets[r]:offsets[r + 1]])
        f.write(b"@read%d x\n%s\n+\n%s\n" % (r, s, bytes(rng.randrange(33, 74) for _ in range(len(s)))))


def run(label, launcher, extra):
    out = os.path.join(tmp, label)
    os.makedirs(out)
    cmd = launcher + ["-m", "trio_binning_b200.classify_by_kmers", reads, pa, pb,
                      "--haplotype-a-out-prefix", os.path.join(out, "A"), "--haplotype-b-out-prefix", os.path.join(out, "B"),
                      "--unclassified-out-prefix", os.path.join(out, "U"), "--batch-bases", "2000000"] + extra
    t0 = time.time()
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True)
    if p.returncode != 0:
        print(label, "FAILED", p.stderr[-2000:])
        sys.exit(1)
    tsv = "".join(line + "\n" for line in p.stdout.splitlines() if "\t" in line)
    h = hashlib.sha256(tsv.encode())
    for n in ("A.fastq.gz", "B.fastq.gz", "U.fastq.gz"):
        h.update(gzip.decompress(open(os.path.join(out, n), "rb").read()))
    print(f"{label:>12}: {time.time() - t0:5.1f} s, {tsv.count(chr(10))} TSV lines, digest {h.hexdigest()[:16]}", flush=True)
    return h.hexdigest()


py = [sys.executable]
tr = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={N}",
      "--master-addr", "127.0.0.1", "--master-port", "29533"]
digests = [run("native-1gpu", py, []), run("python-1gpu", py, ["--python-io"]), run(f"torchrun-{N}", tr, [])]
assert len(set(digests)) == 1, "outputs differ"
print("identical outputs")
