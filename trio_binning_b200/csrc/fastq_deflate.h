// fastq_deflate.h -- a gzip encoder for FASTA/FASTQ text (host side, output writer).
//
// the reference writes its three outputs through Python's gzip at level 9,
// four orders of magnitude below the kernel rate.  What has to match the reference is the
// DECOMPRESSED content; the compressed bytes are free.  General-purpose deflate spends nearly
// all of its time looking for matches that long-read FASTQ does not have (bases and qualities
// barely repeat within 32 KB) and then codes bases and qualities with ONE Huffman code, which
// costs a base 3 bits instead of 2.  This encoder does the opposite: no match search (only runs
// of one repeated byte become matches of distance 1, which binned or saturated quality strings
// are full of), and a new dynamic-Huffman block whenever the text changes between sequence
// lines and everything else, so each kind of line gets its own code.  On long reads that is both
// several times faster than zlib at any level and smaller than zlib -9.
//
// The output is a standard gzip member (RFC 1951/1952).  Every member is decoded again with the
// reader's own inflate (gz_stream.cpp) and compared with the input before it is accepted; the
// caller falls back to zlib if that ever fails or if the text is not worth it (short lines).
// (Short reads were tried -- mixed blocks, read names matched against the previous name -- and
// dropped: on 150-bp reads with binned qualities a Huffman code cannot go below one bit per
// quality character, and the result was 34 % of the input against 31 % for zlib -1 at a similar
// speed.)
#pragma once
#include <stddef.h>

#include <string>

namespace tb {

// Append one gzip member holding data[0..n) to `out`.  Returns false (and leaves `out` as it
// was) if the text does not suit this encoder or the self-check failed; the caller then uses
// zlib.  `verify` = decode and compare before accepting (always on in the pipeline).
bool fastq_gzip_member(const char *data, size_t n, std::string &out, bool verify = true);

}  // namespace tb
