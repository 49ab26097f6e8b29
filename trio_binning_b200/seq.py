"""FASTA/FASTQ reading and writing for classify-by-kmers.

Host-side mirror of the reference's sequence I/O (reference
src/trio_binning/seq.py): `Read` (:16-42), `readfq` (:45-83),
`open_fastx_read` (:86-92) and `open_outfiles` (:98-136), with the same
observable behaviour, including the edge cases the reference inherits from
its line handling and its output-file naming (Q8):

* a record's name is the header up to the first SPACE (tabs are kept);
* every line loses its last character unconditionally, so a file without a
  final newline loses its last base (FASTA) or turns its last FASTQ record
  into a FASTA record without qualities;
* multi-line FASTA and FASTQ are accepted, quality lines may start with '@';
* a FASTQ record with an empty sequence prints as FASTA;
* without gzip, the haplotype-B handle is opened on haplotype A's file name
  (seq.py:129) -- the reference's own golden file depends on it.

This is a fresh implementation (an explicit state machine over lines), not a
copy of the reference's generator; tests/test_seq.py pins it to golden
results produced by the reference's code.
"""
import gzip
import sys
from typing import IO, Iterable, Iterator, Optional, Tuple


class Read:
    """A FASTA/FASTQ record: name, sequence and optional quality string."""

    __slots__ = ("name", "seq", "qual")

    def __init__(self, name: str, seq: str, qual: Optional[str] = None):
        self.name = name
        self.seq = seq
        self.qual = qual

    def __repr__(self):
        return f"Read(name={self.name!r}, seq={self.seq!r}, qual={self.qual!r})"

    def __eq__(self, other):
        return (isinstance(other, Read) and self.name == other.name and self.seq == other.seq
                and self.qual == other.qual)

    def __str__(self):
        # seq.py:27-31 -- FASTQ only when the quality string is non-empty
        if self.qual:
            return "@" + self.name + "\n" + self.seq + "\n+\n" + self.qual
        return ">" + self.name + "\n" + self.seq

    def print(self, file: IO = sys.stdout):
        """Write the record (FASTQ if it has qualities, else FASTA) plus a newline."""
        print(self, file=file)


def _body(line: str) -> str:
    """A line without its last character, whatever that character is (Q7)."""
    return line[:-1]


def readfq(fp: Iterable[str]) -> Iterator[Read]:
    """Yield a `Read` for every record of a FASTA/FASTQ text stream.

    Behaviourally equivalent to the reference's readfq (seq.py:45-83); no
    validation is attempted, as there.
    """
    lines = iter(fp)
    header = ""  # pending header line (already chopped); "" = none
    while True:
        # -- find the next header unless the previous record left one behind
        if not header:
            for line in lines:
                if line[0] == ">" or line[0] == "@":
                    header = _body(line)
                    break
            if not header:
                return
        name = header[1:].split(" ", 1)[0]
        header = ""
        # -- sequence lines, up to the next line starting with '@', '+' or '>'
        chunks = []
        for line in lines:
            if line[0] in "@+>":
                header = _body(line)
                break
            chunks.append(_body(line))
        sequence = "".join(chunks)
        if not header or header[0] != "+":
            # FASTA record (or end of input)
            yield Read(name, sequence, None)
            if not header:
                return
            continue
        # -- FASTQ: quality lines until at least len(sequence) characters are read
        quals, have, complete = [], 0, False
        for line in lines:
            quals.append(_body(line))
            have += len(line) - 1
            if have >= len(sequence):
                complete = True
                break
        if complete:
            header = ""
            yield Read(name, sequence, "".join(quals))
        else:
            # input ended inside the quality block: the reference emits the record
            # without qualities and stops
            yield Read(name, sequence, None)
            return


def open_fastx_read(filename: str) -> Iterator[Read]:
    """Records of a FASTA/FASTQ file; gzip-compressed if the name ends in '.gz'."""
    if filename.endswith(".gz"):
        return readfq(gzip.open(filename, "rt"))
    return readfq(open(filename, "r"))


def open_outfiles(haplotype_a_prefix: str, haplotype_b_prefix: str, unclassified_prefix: str,
                  outfile_extension: str, gzip_output: bool) -> Tuple[IO, IO, IO]:
    """Open the three output files (haplotype A, haplotype B, unclassified).

    gzip mode: `<prefix><ext>.gz`, text mode, default level 9 (seq.py:132-134).
    plain mode: `<prefix><ext>` -- and, as in the reference (seq.py:129), the
    haplotype-B handle is a second handle on haplotype A's file name.
    """
    name_a = haplotype_a_prefix + outfile_extension
    name_b = haplotype_b_prefix + outfile_extension
    name_u = unclassified_prefix + outfile_extension
    if gzip_output:
        return (gzip.open(name_a + ".gz", "wt"), gzip.open(name_b + ".gz", "wt"),
                gzip.open(name_u + ".gz", "wt"))
    # bug-compatible with the reference: B writes into A's file
    return open(name_a, "w"), open(name_a, "w"), open(name_u, "w")
