"""trio_binning_b200 -- B200-native `classify-by-kmers` hot path of trio_binning.

Modules mirror the reference package (reference src/trio_binning/):
  kmers              ctypes boundary to the CUDA library (sets, counting, batches)
  seq                FASTA/FASTQ reader and writer
  classify_by_kmers  the command-line driver

Importing `trio_binning_b200.kmers` loads (or builds) the sm_100a shared
library; there is no CPU implementation in this package.
"""
__version__ = "0.1.0"
