"""Generate the committed golden fixtures under tests/golden/.

Run in the BUILD container only (needs /root/reference):

    python tests/golden/make_golden.py

What it writes (all small; the GPU box has no /root/reference, so the `-m gpu`
parity tests rely on these files plus the oracle):

  ref_data/*            byte copies of the reference's own DATA fixtures
                        (reference tests/data/: k-mer lists, FASTA/FASTQ inputs
                        and its golden hapA.fastq) -- data, not source code.
  cli_fixture.json      stdout TSV + the output files of the reference CLI run on
                        those fixtures (gz and --no-gzip-output modes), produced
                        by the UNMODIFIED reference built in a /tmp copy.
  sets/*.txt            synthetic k-mer lists (k = 21, 31, 32, 11, 4) with
                        unfindable (non-canonical) k-mers, duplicates and k-mers
                        present in both sets.
  counts_golden.json    per-read (count_A, count_B) for seeded synthetic reads,
                        from oracle/_ref/kmers_ref.so (unmodified reference; reads
                        with non-ACGT only at positions >= k) and from
                        oracle/_ref/kmers_ref_zinit.so (same source, zeroed
                        scratch; non-ACGT anywhere -- the Q3-UB contract).
  kat.json              known answers for kmer_to_int / reverse_complement /
                        hash_function from the compiled reference.
  seq_golden.json       readfq()/Read.__str__ results of the reference's seq.py
                        (imported from /root/reference/src) on edge-case inputs.
"""
import gzip
import io
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.oracle import CompiledReference  # noqa: E402

REF = "/root/reference"
COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}


def revcomp(s):
    return "".join(COMP[c] for c in reversed(s))


def rand_seq(rng, n):
    return "".join("ACGT"[i] for i in rng.integers(0, 4, n))


def make_sets(rng, k, genome_len=6000, snp_every=150):
    """Two haplotypes of one random genome; hap-specific k-mers over the SNPs,
    written in random orientation (=> about half are unfindable, quirk Q1)."""
    g = rand_seq(rng, genome_len)
    hap = {}
    for name in "AB":
        s = list(g)
        for pos in range(int(rng.integers(20, snp_every)), genome_len, snp_every):
            s[pos] = "ACGT"[("ACGT".index(s[pos]) + int(rng.integers(1, 4))) % 4]
        hap[name] = "".join(s)

    def kmers_of(s):
        return {s[i:i + k] for i in range(len(s) - k + 1)}

    def canon_pair(s):
        return {x for x in s} | {revcomp(x) for x in s}

    ka, kb = kmers_of(hap["A"]), kmers_of(hap["B"])
    only_a = sorted(ka - canon_pair(kb))
    only_b = sorted(kb - canon_pair(ka))
    lines = {}
    for name, only in (("A", only_a), ("B", only_b)):
        out = []
        for km in only:
            out.append(km if rng.random() < 0.5 else revcomp(km))
        lines[name] = out
    # k-mers in BOTH sets (quirk Q4): shared genome k-mers
    shared = sorted(ka & kb)
    both = [shared[int(i)] for i in rng.integers(0, len(shared), 25)]
    lines["A"] += both
    lines["B"] += both
    # duplicates (counted twice in num_kmers, quirk Q5)
    for name in "AB":
        dup = [lines[name][int(i)] for i in rng.integers(0, len(lines[name]), 10)]
        lines[name] += dup
        order = rng.permutation(len(lines[name]))
        lines[name] = [lines[name][int(i)] for i in order]
    return hap, lines


def make_reads(rng, hap, k, n_reads, invalid_anywhere):
    reads = []
    junk = "NnacgtRYx-"
    for r in range(n_reads):
        src = hap["AB"[int(rng.integers(0, 2))]]
        L = int(rng.choice([0, 1, k - 1, k, k + 1, 40, 150, 600, 2500]))
        L = max(0, L + int(rng.integers(0, 30)) if L > k + 1 else L)
        start = int(rng.integers(0, max(1, len(src) - L)))
        s = src[start:start + L]
        if rng.random() < 0.5:
            s = revcomp(s)
        s = list(s)
        # sequencing errors
        for pos in np.nonzero(rng.random(len(s)) < 0.01)[0]:
            s[pos] = "ACGT"[int(rng.integers(0, 4))]
        # non-ACGT bytes
        rate = [0.0, 0.0, 0.02, 0.08][int(rng.integers(0, 4))]
        lo = 0 if invalid_anywhere else k
        for pos in np.nonzero(rng.random(len(s)) < rate)[0]:
            if pos >= lo:
                s[pos] = junk[int(rng.integers(0, len(junk)))]
        # occasional long invalid run
        if len(s) > 4 * k and rng.random() < 0.3:
            a = int(rng.integers(lo, len(s) - 1))
            b = min(len(s), a + int(rng.integers(1, 3 * k)))
            for pos in range(a, b):
                s[pos] = "N"
        if invalid_anywhere and len(s) > 0 and rng.random() < 0.5:
            # make sure the UB corner is exercised: junk inside the first k bases
            for _ in range(int(rng.integers(1, 4))):
                pos = int(rng.integers(0, min(len(s), k)))
                s[pos] = junk[int(rng.integers(0, len(junk)))]
            if rng.random() < 0.3:
                n0 = int(rng.integers(1, min(len(s), 2 * k) + 1))
                for pos in range(n0):
                    s[pos] = "N"
        reads.append("".join(s))
    return reads


def gen_counts():
    os.makedirs(os.path.join(HERE, "sets"), exist_ok=True)
    ref = CompiledReference(zinit=False)
    zref = CompiledReference(zinit=True)
    cases = []
    for k, seed in ((21, 1), (31, 2), (32, 3), (11, 4), (4, 5)):
        rng = np.random.default_rng(1000 + seed)
        if k == 4:
            # tiny k: a handful of 4-mers, everything collides
            hap = {"A": rand_seq(rng, 3000), "B": rand_seq(rng, 3000)}
            lines = {"A": ["ACGT", "AAAA", "TTGA", "CCCC", "GATC"], "B": ["TTTT", "ACGT", "GGCA", "CATG"]}
        else:
            hap, lines = make_sets(rng, k)
        if k == 32:
            lines["A"].append("T" * 32)  # key == 2^64-1: the GPU table's empty sentinel
            lines["B"].append("A" * 32)
            hap["A"] = hap["A"][:500] + "A" * 70 + "T" * 70 + hap["A"][500:]
        paths = {}
        for name in "AB":
            p = os.path.join(HERE, "sets", f"k{k}_hap{name}.txt")
            with open(p, "w") as f:
                f.write("".join(x + "\n" for x in lines[name]))
            paths[name] = p
        for label, lib, anywhere in (("ref", ref, False), ("ref_zinit", zref, True)):
            A = lib.set_from_file(paths["A"])
            B = lib.set_from_file(paths["B"])
            reads = make_reads(rng, hap, k, 60, anywhere)
            counts = [list(lib.count(r, A, B)) for r in reads]
            cases.append({
                "name": f"k{k}_{label}", "k": k, "source": label,
                "set_a": f"sets/k{k}_hapA.txt", "set_b": f"sets/k{k}_hapB.txt",
                "num_kmers_a": A.contents.num_kmers, "num_kmers_b": B.contents.num_kmers,
                "reads": reads, "counts": counts,
            })
            nz = sum(1 for c in counts if c != [0, 0])
            print(f"  {cases[-1]['name']}: {len(reads)} reads, {nz} with hits,"
                  f" sum={np.sum(counts, axis=0).tolist()}")
    with open(os.path.join(HERE, "counts_golden.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py", "cases": cases}, f)


def gen_kat():
    ref = CompiledReference()
    rng = np.random.default_rng(7)
    kmers = ["ATGCTAGCTAGAGAGAGAGGA", "T" * 28, "TTTTTTTTTTTTTTTAGGCCCACTTTTT", "A" * 26,
             "GGGAGGGAGGGAGGGAGGGAGGGAGGG", "T" * 32, "ACGTNacgtACGTRYKMACGTT", "C"]
    kmers += [rand_seq(rng, int(rng.integers(1, 33))) for _ in range(40)]
    ints = [0, 1, 696357446508, 2**42 - 1, 2**62 - 1, 2**64 - 1] + [int(x) for x in rng.integers(0, 2**63, 40)]
    out = {
        "kmer_to_int": [[s, ref.kmer_to_int(s)] for s in kmers],
        "reverse_complement": [[s, ref.reverse_complement(s)] for s in kmers],
        "hash_function": [[x, ref.hash_function(x)] for x in ints],
    }
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(out, f, indent=0)


def gen_cli():
    """Run the unmodified reference CLI (built in a /tmp copy) on its fixtures."""
    dst = os.path.join(HERE, "ref_data")
    os.makedirs(dst, exist_ok=True)
    for name in os.listdir(os.path.join(REF, "tests", "data")):
        shutil.copyfile(os.path.join(REF, "tests", "data", name), os.path.join(dst, name))
        os.chmod(os.path.join(dst, name), 0o644)
    tmp = tempfile.mkdtemp(prefix="refbuild_")
    work = os.path.join(tmp, "ref")
    shutil.copytree(REF, work)
    subprocess.run([sys.executable, "setup.py", "-q", "build_ext", "--inplace"], cwd=work, check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    env = dict(os.environ, PYTHONPATH=os.path.join(work, "src"))
    result = {}
    for mode, extra in (("gz", []), ("plain", ["--no-gzip-output"])):
        out_dir = os.path.join(tmp, mode)
        os.makedirs(out_dir)
        cmd = [sys.executable, "-m", "trio_binning.classify_by_kmers",
               os.path.join(dst, "test.ccs.fastq.gz"), os.path.join(dst, "hapA.txt"), os.path.join(dst, "hapB.txt"),
               "--haplotype-a-out-prefix", os.path.join(out_dir, "hapA"),
               "--haplotype-b-out-prefix", os.path.join(out_dir, "hapB"),
               "--unclassified-out-prefix", os.path.join(out_dir, "hapU")] + extra
        p = subprocess.run(cmd, env=env, capture_output=True, text=True, check=True)
        files = {}
        for name in sorted(os.listdir(out_dir)):
            raw = open(os.path.join(out_dir, name), "rb").read()
            if name.endswith(".gz"):
                raw = gzip.decompress(raw)
            import hashlib
            files[name] = {"bytes": len(raw), "sha256": hashlib.sha256(raw).hexdigest(),
                           "n_lines": raw.count(b"\n"),
                           "first_line": raw.split(b"\n", 1)[0].decode()}
        result[mode] = {"stdout": p.stdout, "files": files}
    # a FASTA input and a plain FASTQ input through the CLI as well (ext logic, Q8)
    for name in ("test.fa", "test.fastq"):
        out_dir = os.path.join(tmp, name)
        os.makedirs(out_dir)
        cmd = [sys.executable, "-m", "trio_binning.classify_by_kmers",
               os.path.join(dst, name), os.path.join(dst, "hapA.txt"), os.path.join(dst, "hapB.txt"),
               "--haplotype-a-out-prefix", os.path.join(out_dir, "hapA"),
               "--haplotype-b-out-prefix", os.path.join(out_dir, "hapB"),
               "--unclassified-out-prefix", os.path.join(out_dir, "hapU")]
        p = subprocess.run(cmd, env=env, capture_output=True, text=True, check=True)
        files = {}
        for fn in sorted(os.listdir(out_dir)):
            files[fn] = gzip.decompress(open(os.path.join(out_dir, fn), "rb").read()).decode()
        result[name] = {"stdout": p.stdout, "files_text": files}
    with open(os.path.join(HERE, "cli_fixture.json"), "w") as f:
        json.dump(result, f, indent=1)
    shutil.rmtree(tmp)


def gen_seq():
    sys.path.insert(0, os.path.join(REF, "src"))
    from trio_binning import seq as refseq  # the reference's pure-Python FASTX reader

    inputs = {
        "fasta_multiline": ">r1 comment here\nACGT\nGGCC\n>r2\tx y\nTTTT\n",
        "fastq_basic": "@r1 c\nACGT\n+\n!!!!\n@r2\nGG\n+r2\n##\n",
        "fastq_multiline": "@r1\nACGT\nTTGA\n+\n!!!!\n####\n@r2\nAC\n+\n@@\n",
        "fastq_qual_starts_at": "@r1\nACGT\n+\n@!!!\n@r2\nAC\n+\n!!\n",
        "no_trailing_newline_fasta": ">r1\nACGT\n>r2\nGGCC",
        "no_trailing_newline_fastq": "@r1\nACGT\n+\n!!!!",
        "crlf": ">r1\r\nACGT\r\n>r2\r\nGG\r\n",
        "empty_seq_fastq": "@r1\n\n+\n\n@r2\nAC\n+\n!!\n",
        "leading_garbage": "garbage\nmore\n>r1\nAC\n",
        "mixed": ">f1\nACGT\n@q1\nGGTT\n+\n!!!!\n>f2\nTT\n",
        "truncated_fastq": "@r1\nACGTACGT\n+\n!!!\n",
        "blank_lines": ">r1\nAC\n\nGT\n>r2\n\n",
        "empty": "",
        "header_only": ">r1\n",
        "long_qual": "@r1\nACG\n+\n!!!!!\n@r2\nAA\n+\n!!\n",
    }
    out = {}
    for name, text in inputs.items():
        try:
            recs = [[r.name, r.seq, r.qual, str(r)] for r in refseq.readfq(io.StringIO(text))]
            out[name] = {"input": text, "records": recs}
        except Exception as e:  # the reference crashes on some inputs ("no error checking")
            out[name] = {"input": text, "raises": type(e).__name__}
    with open(os.path.join(HERE, "seq_golden.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    print("counts:")
    gen_counts()
    gen_kat()
    gen_cli()
    gen_seq()
    print("done")
