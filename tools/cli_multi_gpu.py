"""Multi-GPU CLI check (run on a multi-GPU box): the same FASTQ through
  (a) one GPU  (b) N GPUs, batches dealt on demand  (c) the Python reader/writer on one GPU
must give identical TSV and identical decompressed outputs; prints wall times and the pipeline's
stage timings.  usage: cli_multi_gpu.py [N] [n_reads] [mean_len]"""
import gzip
import hashlib
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from trio_binning_b200 import synth  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
mean_len = int(sys.argv[3]) if len(sys.argv) > 3 else 8000
k = 21
tmp = tempfile.mkdtemp(prefix="tb_cli_")
ka, kb = synth.random_kmer_sets(200_000, 150_000, k, seed=5)
pa, pb = os.path.join(tmp, "hapA.txt"), os.path.join(tmp, "hapB.txt")
synth.write_kmer_file(pa, ka, k)
synth.write_kmer_file(pb, kb, k)
bases, offsets, _ = synth.random_reads(n_reads, mean_len, k, ka, kb, seed=6, plant_every=500, invalid_rate=0.0005)
import numpy as np  # noqa: E402

qual_pool = np.random.default_rng(1).integers(33, 74, size=1 << 22, dtype=np.uint8)
reads = os.path.join(tmp, "reads.fastq")
with open(reads, "wb") as f:
    for r in range(len(offsets) - 1):
        s = bases[offsets[r]:offsets[r + 1]].tobytes()
        start = (r * 7919) % (len(qual_pool) - len(s) - 1) if len(s) < len(qual_pool) - 1 else 0
        q = qual_pool[start:start + len(s)].tobytes() if len(s) < len(qual_pool) - 1 else b"I" * len(s)
        f.write(b"@read%d extra\n" % r + s + b"\n+\n" + q + b"\n")


def run(tag, extra):
    d = os.path.join(tmp, tag)
    os.makedirs(d)
    cmd = [sys.executable, "-m", "trio_binning_b200.classify_by_kmers", reads, pa, pb,
           "--haplotype-a-out-prefix", os.path.join(d, "A"), "--haplotype-b-out-prefix", os.path.join(d, "B"),
           "--unclassified-out-prefix", os.path.join(d, "U"), "--batch-bases", str(8 << 20)] + extra
    t0 = time.time()
    res = subprocess.run(cmd, capture_output=True, cwd=ROOT, env=dict(os.environ, TB_VERBOSE="1"))
    dt = time.time() - t0
    assert res.returncode == 0, res.stderr.decode()[-2000:]
    digest = [hashlib.sha256(res.stdout).hexdigest()]
    for name in sorted(os.listdir(d)):
        digest.append(name + ":" + hashlib.sha256(gzip.decompress(open(os.path.join(d, name), "rb").read())).hexdigest())
    stats = [line for line in res.stderr.decode().splitlines() if line.startswith("[tb] native pipeline")]
    print(f"{tag}: {dt:.2f} s  {stats[-1] if stats else ''}", flush=True)
    return digest


digests = [run("native-1gpu", ["--gpus", "1"]), run(f"native-{N}gpu", ["--gpus", str(N)]),
           run("python-1gpu", ["--python-io"])]
assert digests[0] == digests[1] == digests[2], digests
print(f"identical TSV and outputs: 1 GPU, {N} GPUs, Python I/O ({int(offsets[-1])} bases, {n_reads} reads)")
