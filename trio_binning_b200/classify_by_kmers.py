"""Classify reads into bins based on kmers.

This is a script for classifying sequence reads into parental bins
based on the presence of k-mers.
"""
# Host-side mirror of the reference driver (reference
# src/trio_binning/classify_by_kmers.py): same command line (:14-54), same
# scaling factors (:57-77), same three output files and the same TSV on stdout
# (:99-117).  The per-read ctypes call of the reference's loop is replaced by the
# native pipeline (csrc/fastx_pipeline.cpp): one process parses the reads once,
# packs them to 2 bits, deals batches to one worker thread per GPU (each GPU holds
# its own replica of the two sets), and writes records and TSV lines in input
# order -- gzip or plain, whatever the reference's flags ask for.  --python-io keeps
# the reference's own reader/writer code path (seq.py) around the batched GPU call.

import argparse
import os
import sys
from os import path
from typing import Tuple

from . import kmers, seq

DEFAULT_BATCH_BASES = 256 << 20


# The reference's command line (classify_by_kmers.py:14-54): three positionals and four
# options, same names, defaults and help texts; then the B200 additions.
_REFERENCE_ARGUMENTS = (
    (("reads",), dict(help="reads to classify into bins, in fasta/q format. Can be gzipped.")),
    (("haplotype_a_kmers",), dict(type=kmers.create_kmer_hash_set,
                                  help="a list of k-mers unique to haplotype A, one per line")),
    (("haplotype_b_kmers",), dict(type=kmers.create_kmer_hash_set,
                                  help="a list of k-mers unique to haplotype B, one per line")),
    (("--haplotype-a-out-prefix",), dict(default="hapA", help="prefix for haplotype A output file")),
    (("--haplotype-b-out-prefix",), dict(default="hapB", help="prefix for haplotype B output file")),
    (("--unclassified-out-prefix",), dict(default="unclassified", help="prefix for unclassified output file")),
    (("--no-gzip-output",), dict(action="store_true", default=False, help="don't gzip the output")),
)
_B200_ARGUMENTS = (
    (("--batch-bases",), dict(type=int, default=DEFAULT_BATCH_BASES,
                              help="bases per GPU batch (B200 addition; does not change results)")),
    (("--python-io",), dict(action="store_true", default=False,
                            help="read and write FASTA/FASTQ in Python like the reference instead of the "
                                 "native multi-threaded pipeline (B200 addition; same output, one GPU)")),
    (("--gpus",), dict(type=int, default=1,
                       help="number of GPUs to spread the batches over, 0 = all visible ones; the sets are "
                            "replicated on each, the output does not depend on it. One B200 classifies about a "
                            "hundred times faster than one host parses FASTQ, so more GPUs only pay off when the "
                            "reads come from something faster than a file (B200 addition)")),
    (("--gpu-workers",), dict(type=int, default=0,
                              help="device worker threads of the pipeline, 0 = one per GPU used; more than that wrap "
                                   "around the GPUs, each worker with its own replica of the sets (B200 addition; "
                                   "exercises the multi-GPU merge on a single-GPU machine)")),
    (("--gzip-level",), dict(type=int, default=6,
                             help="zlib level for output chunks the FASTQ-aware encoder declines (short reads) "
                                  "or for everything with --no-fastq-gzip: 6 costs a fraction of the time of 9, "
                                  "the level the reference's gzip.open uses, for files about 0.5 percent larger; "
                                  "the decompressed bytes are the same at any level (B200 addition)")),
    (("--no-fastq-gzip",), dict(action="store_true", default=False,
                                help="compress the outputs with zlib only. By default long-read chunks go through "
                                     "the FASTQ-aware gzip encoder: standard gzip files, smaller than zlib level 9 "
                                     "on long reads and many times faster; every block is decoded and compared "
                                     "before it is written (B200 addition)")),
    (("--fastq-gzip",), dict(action="store_true", default=False,
                             help="(default since round 2; kept so that old command lines still parse)")),
    (("--threads",), dict(type=int, default=0,
                          help="host threads for the native pipeline, 0 = all cores (B200 addition)")),
)


def parse_args(argv=None):
    """Parse arguments; the two k-mer sets are built on the GPU while parsing, because the
    reference passes its set constructor as the argparse `type` of those positionals."""
    parser = argparse.ArgumentParser(description=__doc__,
                                     formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    for flags, options in _REFERENCE_ARGUMENTS + _B200_ARGUMENTS:
        parser.add_argument(*flags, **options)
    return parser.parse_args(argv)


def calculate_scaling_factors(haplotype_a_kmers, haplotype_b_kmers) -> Tuple[float, float]:
    """Scaling factors for the k-mer scores: each count is multiplied by
    max(nA, nB) / n_own, with n the number of k-mers (lines) of the set
    (reference classify_by_kmers.py:57-77)."""
    num_a = kmers.get_number_kmers_in_set(haplotype_a_kmers)
    num_b = kmers.get_number_kmers_in_set(haplotype_b_kmers)
    biggest = max(num_a, num_b)
    return 1.0 * biggest / num_a, 1.0 * biggest / num_b


def _emit(read, result, files, out):
    """One read's record into the file of its bin and its TSV line (reference :104-117)."""
    _, _, score_a, score_b, read_bin = result
    file_a, file_b, file_u = files
    if read_bin == "A":
        read.print(file=file_a)
    elif read_bin == "B":
        read.print(file=file_b)
    else:
        read.print(file=file_u)
    print("\t".join(map(str, [read.name, read_bin, score_a, score_b])), file=out)


def _main_native(args):
    """The whole loop (parse, pack, classify on every GPU, route, compress or write plain, TSV)
    in the native pipeline; file names follow the reference (:90 and seq.py:122-134)."""
    extension = path.splitext(args.reads.rstrip(".gz"))[1]  # reference :90 (quirk Q8)
    suffix = extension + ("" if args.no_gzip_output else ".gz")
    names = [prefix + suffix for prefix in
             (args.haplotype_a_out_prefix, args.haplotype_b_out_prefix, args.unclassified_out_prefix)]
    n_gpus = kmers.device_count() if args.gpus <= 0 else min(args.gpus, kmers.device_count())
    first = kmers.lib.tb_get_device()
    devices = [first] + [d for d in range(kmers.device_count()) if d != first][: n_gpus - 1]
    if args.gpu_workers > len(devices):
        devices = [devices[i % len(devices)] for i in range(args.gpu_workers)]
    sets_a = [args.haplotype_a_kmers] + [kmers.clone_kmer_hash_set(args.haplotype_a_kmers, d) for d in devices[1:]]
    sets_b = [args.haplotype_b_kmers] + [kmers.clone_kmer_hash_set(args.haplotype_b_kmers, d) for d in devices[1:]]
    sys.stdout.flush()
    try:
        fd = sys.stdout.fileno()
        borrowed = None
    except (AttributeError, OSError, ValueError):
        # stdout is not a real file (captured): collect the TSV in a temporary file
        import tempfile

        borrowed = tempfile.TemporaryFile()
        fd = borrowed.fileno()
    try:
        stats = kmers.classify_fastx_file(args.reads, sets_a, sets_b, names[0], names[1], names[2], tsv_fd=fd,
                                          gzip_level=args.gzip_level, batch_bases=args.batch_bases,
                                          threads=args.threads, gzip_output=not args.no_gzip_output,
                                          fastq_encoder=not args.no_fastq_gzip)
    finally:
        for replica in sets_a[1:] + sets_b[1:]:
            kmers.destroy_kmer_hash_set(replica)
    if os.environ.get("TB_VERBOSE"):
        print("[tb] native pipeline: " + ", ".join(f"{k}={v:.3f}" if isinstance(v, float) else f"{k}={v}"
                                                   for k, v in stats.items()), file=sys.stderr)
    if borrowed is not None:
        borrowed.seek(0)
        sys.stdout.write(borrowed.read().decode("utf-8"))
        borrowed.close()
    return stats


def _batches(reads, batch_bases):
    """Consecutive reads until a batch holds batch_bases bases."""
    batch, total = [], 0
    for read in reads:
        batch.append(read)
        total += len(read.seq)
        if total >= batch_bases:
            yield batch
            batch, total = [], 0
    if batch:
        yield batch


def main(argv=None):
    """Main method of program"""
    # One process drives every GPU.  Started under torchrun (one process per GPU) only rank 0
    # works -- the others would just parse the same file again.
    if int(os.environ.get("RANK", "0")) != 0:
        print("classify-by-kmers: one process drives all GPUs; rank "
              f"{os.environ['RANK']} has nothing to do", file=sys.stderr)
        return

    args = parse_args(argv)

    if not args.python_io:
        _main_native(args)
        return

    reads = seq.open_fastx_read(args.reads)

    # three locals, released in this order when main() returns, like the reference's
    file_a, file_b, file_u = seq.open_outfiles(
        args.haplotype_a_out_prefix,
        args.haplotype_b_out_prefix,
        args.unclassified_out_prefix,
        # reference :90 -- rstrip strips a CHARACTER SET, kept as is (quirk Q8)
        path.splitext(args.reads.rstrip(".gz"))[1],
        not args.no_gzip_output,
    )

    # The device computes count * scaling in fp64 with the same factors
    # (tb_scaling_factors); this call is kept for API parity and as a cross-check.
    scaling_a, scaling_b = calculate_scaling_factors(args.haplotype_a_kmers, args.haplotype_b_kmers)
    assert (scaling_a, scaling_b) == kmers.scaling_factors(args.haplotype_a_kmers, args.haplotype_b_kmers)

    set_a, set_b = args.haplotype_a_kmers, args.haplotype_b_kmers
    for batch in _batches(reads, args.batch_bases):
        results = kmers.classify_reads([read.seq for read in batch], set_a, set_b)
        for read, result in zip(batch, results):
            _emit(read, result, (file_a, file_b, file_u), sys.stdout)


if __name__ == "__main__":
    main()
