// fastq_deflate.cpp -- see fastq_deflate.h.  RFC 1951 dynamic-Huffman blocks, literals and
// distance-1 runs only, one block per stretch of like lines; RFC 1952 wrapper.
#include "fastq_deflate.h"

#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "gz_stream.h"

namespace tb {

namespace {

constexpr size_t kMinBlock = 2048;         // do not open a new block for less text than this
constexpr size_t kMaxBlock = 256u << 10;   // nor let one grow beyond this (keeps the codes local)
constexpr size_t kLongLine = 64;           // lines at least this long decide the class of a block
constexpr int kMaxLitBits = 14;            // literal/length code lengths: four literals fit one 56-bit put
constexpr int kPairSymbols = 128;          // byte-pair table for blocks with at most this many distinct bytes

const uint16_t LEN_BASE[29] = { 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59,
                                67, 83, 99, 115, 131, 163, 195, 227, 258 };
const uint8_t LEN_EXTRA[29] = { 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0 };
const uint8_t CLEN_ORDER[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };

// match length 3..258 -> length symbol index (0..28) and extra-bit value
struct LenCode {
    uint8_t sym[259], extra_bits[259];
    uint16_t extra_val[259];
    LenCode()
    {
        for (unsigned len = 3; len <= 258; len++) {
            unsigned s = 28;
            while (LEN_BASE[s] > len) s--;
            if (len == 258) s = 28;
            sym[len] = (uint8_t)s;
            extra_bits[len] = LEN_EXTRA[s];
            extra_val[len] = (uint16_t)(len - LEN_BASE[s]);
        }
    }
};
const LenCode &len_code()
{
    static const LenCode lc;
    return lc;
}

// Huffman code lengths of at most `limit` bits for the symbols with freq > 0 (two-queue
// construction on the sorted frequencies; if the tree comes out too deep the frequencies are
// flattened and it is rebuilt -- blocks are small, this almost never triggers).
void huffman_lengths(const uint32_t *freq_in, int n, int limit, uint8_t *lens)
{
    memset(lens, 0, (size_t)n);
    std::vector<uint32_t> freq(freq_in, freq_in + n);
    std::vector<int> used;
    for (int s = 0; s < n; s++)
        if (freq[s]) used.push_back(s);
    if (used.empty()) return;
    if (used.size() == 1) {
        lens[used[0]] = 1;
        return;
    }
    const int m = (int)used.size();
    std::vector<uint64_t> weight((size_t)2 * m);
    std::vector<int> parent((size_t)2 * m), depth((size_t)2 * m);
    for (;;) {
        std::sort(used.begin(), used.end(), [&](int a, int b) { return freq[a] != freq[b] ? freq[a] < freq[b] : a < b; });
        for (int i = 0; i < m; i++) weight[i] = freq[used[i]];
        int leaf = 0, node = m, next = m;  // leaves [0,m), internal nodes [m, 2m-1)
        while (next < 2 * m - 1) {
            int pick[2];
            for (int t = 0; t < 2; t++) {
                if (leaf < m && (node >= next || weight[leaf] <= weight[node])) pick[t] = leaf++;
                else pick[t] = node++;
            }
            weight[next] = weight[pick[0]] + weight[pick[1]];
            parent[pick[0]] = parent[pick[1]] = next;
            next++;
        }
        depth[2 * m - 2] = 0;
        int deepest = 0;
        for (int i = 2 * m - 3; i >= 0; i--) {
            depth[i] = depth[parent[i]] + 1;
            if (i < m && depth[i] > deepest) deepest = depth[i];
        }
        if (deepest <= limit) {
            for (int i = 0; i < m; i++) lens[used[i]] = (uint8_t)depth[i];
            return;
        }
        for (int s : used) freq[s] = (freq[s] + 1) / 2 + 1;  // flatten and retry
    }
}

// canonical codes (RFC 1951 3.2.2), already bit-reversed for the LSB-first bit writer
void canonical_codes(const uint8_t *lens, int n, uint16_t *codes)
{
    unsigned count[16] = { 0 }, next[16];
    for (int s = 0; s < n; s++) count[lens[s]]++;
    count[0] = 0;
    unsigned code = 0;
    for (int l = 1; l <= 15; l++) {
        code = (code + count[l - 1]) << 1;
        next[l] = code;
    }
    for (int s = 0; s < n; s++) {
        const unsigned l = lens[s];
        if (!l) { codes[s] = 0; continue; }
        unsigned c = next[l]++, r = 0;
        for (unsigned i = 0; i < l; i++) r |= ((c >> i) & 1u) << (l - 1 - i);
        codes[s] = (uint16_t)r;
    }
}

// LSB-first bit writer.  Invariant between calls: fewer than 8 bits are pending, so a call may add
// up to 56 bits before it stores -- one unaligned 8-byte store, of which only the whole bytes count.
struct BitWriter {
    unsigned char *p, *end;  // an 8-byte store at p is inside the buffer while p <= end
    uint64_t acc = 0;
    unsigned n = 0;
    bool overflow = false;
    inline void flush()
    {
        if (p > end) { overflow = true; acc = 0; n = 0; return; }
        memcpy(p, &acc, 8);  // little-endian hosts
        p += n >> 3;
        acc >>= n & 56u;
        n &= 7u;
    }
    // len <= 56
    inline void put(uint64_t bits, unsigned len)
    {
        acc |= bits << n;
        n += len;
        flush();
    }
    void finish_byte()
    {
        if (n > 0) {
            if (p > end) { overflow = true; return; }
            *p++ = (unsigned char)acc;
        }
        acc = 0;
        n = 0;
    }
};

// One pass over a block, eight bytes at a time: a word of eight equal bytes starts a run (extended
// byte-wise, emitted as one literal + a distance-1 match of up to 258), any other word is eight
// literals.  Shorter or unaligned runs stay literals -- a small loss for a tokenizer that costs
// one compare per eight bytes.  Both passes (histogram, emit) go through this same function.
inline uint64_t load64(const unsigned char *p)
{
    uint64_t v;
    memcpy(&v, p, 8);
    return v;
}
template <class OnLiterals8, class OnLiteral, class OnRun>
inline void tokenize(const unsigned char *d, size_t n, OnLiterals8 lit8, OnLiteral lit, OnRun run)
{
    size_t i = 0;
    while (i + 8 <= n) {
        const uint64_t w = load64(d + i);
        if (w != (w & 0xFF) * 0x0101010101010101ull) {
            lit8(d + i);
            i += 8;
            continue;
        }
        const unsigned char c = d[i];
        size_t r = 8;
        while (i + r < n && d[i + r] == c && r < 259) r++;
        lit(c);
        run((unsigned)(r - 1));  // 7..258 further copies of c
        i += r;
    }
    for (; i < n; i++) lit(d[i]);
}

bool encode_block(const unsigned char *d, size_t n, bool final_block, BitWriter &bw)
{
    const LenCode &lc = len_code();
    uint32_t freq[286], f1[256], f2[256], f3[256];  // four tables: no store-to-load stalls on repeats
    memset(freq, 0, sizeof freq);
    memset(f1, 0, sizeof f1);
    memset(f2, 0, sizeof f2);
    memset(f3, 0, sizeof f3);
    tokenize(d, n,
             [&](const unsigned char *q) {
                 freq[q[0]]++; f1[q[1]]++; f2[q[2]]++; f3[q[3]]++;
                 freq[q[4]]++; f1[q[5]]++; f2[q[6]]++; f3[q[7]]++;
             },
             [&](unsigned char c) { freq[c]++; },
             [&](unsigned len) { freq[257 + lc.sym[len]]++; });
    for (int s = 0; s < 256; s++) freq[s] += f1[s] + f2[s] + f3[s];
    freq[256] = 1;
    uint8_t ll[286], dl[1] = { 1 };  // one distance code (distance 1) of one bit, used or not
    huffman_lengths(freq, 286, kMaxLitBits, ll);
    uint16_t lcodes[286];
    canonical_codes(ll, 286, lcodes);
    int hlit = 286;
    while (hlit > 257 && ll[hlit - 1] == 0) hlit--;
    // code lengths of both alphabets, run-length coded with the code-length alphabet
    uint8_t seq[287];
    memcpy(seq, ll, (size_t)hlit);
    seq[hlit] = dl[0];
    const int n_seq = hlit + 1;
    struct Cl { uint8_t sym, extra_bits, extra_val; };
    Cl cls[287];
    int n_cl = 0;
    uint32_t cfreq[19];
    memset(cfreq, 0, sizeof cfreq);
    for (int i = 0; i < n_seq;) {
        const uint8_t v = seq[i];
        int r = 1;
        while (i + r < n_seq && seq[i + r] == v) r++;
        int left = r;
        if (v == 0) {
            while (left >= 11) { const int t = std::min(left, 138); cls[n_cl++] = { 18, 7, (uint8_t)(t - 11) }; left -= t; }
            if (left >= 3) { cls[n_cl++] = { 17, 3, (uint8_t)(left - 3) }; left = 0; }
            while (left-- > 0) cls[n_cl++] = { 0, 0, 0 };
        } else {
            cls[n_cl++] = { v, 0, 0 };
            left--;
            while (left >= 3) { const int t = std::min(left, 6); cls[n_cl++] = { 16, 2, (uint8_t)(t - 3) }; left -= t; }
            while (left-- > 0) cls[n_cl++] = { v, 0, 0 };
        }
        i += r;
    }
    for (int i = 0; i < n_cl; i++) cfreq[cls[i].sym]++;
    uint8_t cl_len[19];
    huffman_lengths(cfreq, 19, 7, cl_len);
    uint16_t cl_code[19];
    canonical_codes(cl_len, 19, cl_code);
    int hclen = 19;
    while (hclen > 4 && cl_len[CLEN_ORDER[hclen - 1]] == 0) hclen--;

    bw.put(final_block ? 1u : 0u, 1);
    bw.put(2, 2);
    bw.put((uint32_t)(hlit - 257), 5);
    bw.put(0, 5);  // HDIST = 1
    bw.put((uint32_t)(hclen - 4), 4);
    for (int i = 0; i < hclen; i++) bw.put(cl_len[CLEN_ORDER[i]], 3);
    for (int i = 0; i < n_cl; i++) {
        bw.put(cl_code[cls[i].sym], cl_len[cls[i].sym]);
        if (cls[i].extra_bits) bw.put(cls[i].extra_val, cls[i].extra_bits);
    }
    // the text: literals go out two at a time through a table of byte PAIRS (code and length of both,
    // merged), four pairs per eight bytes and two stores.  Only pairs of symbols that occur in the
    // block are filled in -- a few dozen symbols for any kind of FASTA/FASTQ line; text with more
    // distinct bytes than kPairSymbols takes the one-literal path.
    uint8_t used[256];
    int n_used = 0;
    for (int s = 0; s < 256; s++)
        if (ll[s]) used[n_used++] = (uint8_t)s;
    // indexed by the two bytes as one little-endian 16-bit number (a dense index was tried: its
    // multiply costs more than the spread-out rows do)
    static thread_local std::vector<uint64_t> pair_store;
    const bool pairs = n_used <= kPairSymbols && (size_t)n_used * n_used <= 2 * n;  // filling it must not cost more than using it
    if (pairs) {
        if (pair_store.empty()) pair_store.resize(65536);
        for (int a = 0; a < n_used; a++) {
            const unsigned sa = used[a], la = ll[sa];
            const uint64_t ca = lcodes[sa];
            for (int b = 0; b < n_used; b++) {
                const unsigned sb = used[b];
                pair_store[sa | (sb << 8)] =
                    (ca | ((uint64_t)lcodes[sb] << la)) | ((uint64_t)(la + ll[sb]) << 32);
            }
        }
    }
    const uint64_t *pair = pair_store.data();
    auto pair_of = [&](const unsigned char *q) {
        uint16_t i;
        memcpy(&i, q, 2);
        return pair[i];
    };
    auto one = [&](unsigned char c) { bw.put(lcodes[c], ll[c]); };
    tokenize(d, n,
             [&](const unsigned char *q) {
                 if (!pairs) {
                     for (int t = 0; t < 8; t++) one(q[t]);
                     return;
                 }
                 const uint64_t e0 = pair_of(q), e1 = pair_of(q + 2), e2 = pair_of(q + 4), e3 = pair_of(q + 6);
                 const unsigned l0 = (unsigned)(e0 >> 32), l1 = (unsigned)(e1 >> 32), l2 = (unsigned)(e2 >> 32),
                                l3 = (unsigned)(e3 >> 32);
                 // two pairs are at most 4 * kMaxLitBits = 56 bits
                 bw.put((e0 & 0xFFFFFFFFull) | ((e1 & 0xFFFFFFFFull) << l0), l0 + l1);
                 bw.put((e2 & 0xFFFFFFFFull) | ((e3 & 0xFFFFFFFFull) << l2), l2 + l3);
             },
             one,
             [&](unsigned len) {
                 const unsigned s = 257 + lc.sym[len];
                 // length code, its extra bits, and the one distance code (one bit): <= 14 + 5 + 1
                 const unsigned eb = lc.extra_bits[len];
                 bw.put((uint64_t)lcodes[s] | ((uint64_t)lc.extra_val[len] << ll[s]), ll[s] + eb + 1);
             });
    bw.put(lcodes[256], ll[256]);
    return !bw.overflow;
}

inline bool sequence_like(const unsigned char *p, size_t n)
{
    // first 64 bytes of a long line: nucleotide letters only (IUPAC N included)?
    unsigned hits = 0;
    const size_t m = n < 64 ? n : 64;
    for (size_t i = 0; i < m; i++) {
        const unsigned char c = p[i] & 0xDF;  // upper case
        hits += (c == 'A') | (c == 'C') | (c == 'G') | (c == 'T') | (c == 'N');
    }
    return hits * 16 >= m * 15;
}

}  // namespace

bool fastq_gzip_member(const char *data, size_t n, std::string &out, bool verify)
{
    if (n < 4096 || n >= (1ull << 32)) return false;
    const unsigned char *d = (const unsigned char *)data;
    // worth it only for long lines: one code per kind of line needs lines worth a block
    {
        size_t lines = 0;
        for (const unsigned char *q = d; (q = (const unsigned char *)memchr(q, '\n', (size_t)(d + n - q))) != nullptr; q++) lines++;
        if (n / (lines + 1) < 512) return false;
    }
    const size_t start = out.size();
    out.resize(start + n + n / 8 + (n / kMinBlock + 2) * 512 + 64);
    unsigned char *base = (unsigned char *)&out[start];
    const unsigned char head[10] = { 0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 0, 0xff };
    memcpy(base, head, 10);
    BitWriter bw;
    bw.p = base + 10;
    bw.end = base + (out.size() - start) - 24;  // room for the last store, the final byte and the trailer

    // blocks: cut where the class of (long) lines changes, or when a block is full
    size_t block_start = 0, pos = 0;
    int block_class = -1;
    bool ok = true;
    auto close_block = [&](size_t end_pos, bool final_block) {
        if (ok) ok = encode_block(d + block_start, end_pos - block_start, final_block, bw);
        block_start = end_pos;
        block_class = -1;
    };
    while (pos < n && ok) {
        const unsigned char *nl = (const unsigned char *)memchr(d + pos, '\n', n - pos);
        size_t line_end = nl ? (size_t)(nl - d) + 1 : n;
        // very long lines are cut into pieces so that no block outgrows kMaxBlock
        if (line_end - pos > kMaxBlock) line_end = pos + kMaxBlock;
        const size_t len = line_end - pos;
        if (len >= kLongLine) {
            const int cls = sequence_like(d + pos, len) ? 1 : 0;
            if (block_class >= 0 && cls != block_class && pos - block_start >= kMinBlock) close_block(pos, false);
            block_class = cls;
        }
        if (line_end - block_start > kMaxBlock && pos > block_start) close_block(pos, false);
        pos = line_end;
    }
    if (ok) close_block(n, true);
    if (!ok || bw.overflow) {
        out.resize(start);
        return false;
    }
    bw.finish_byte();
    if (bw.overflow) {
        out.resize(start);
        return false;
    }
    unsigned char *t = bw.p;
    const uint32_t crc = crc32_any(0, d, n), isize = (uint32_t)n;
    for (int i = 0; i < 4; i++) t[i] = (unsigned char)(crc >> (8 * i));
    for (int i = 0; i < 4; i++) t[4 + i] = (unsigned char)(isize >> (8 * i));
    const size_t total = (size_t)(t + 8 - base);
    out.resize(start + total);

    if (verify) {
        // decode what was just written with the reader's inflate and compare byte for byte
        GzStream gz;
        gz.open_memory((const unsigned char *)&out[start], total);
        std::vector<char> back(1u << 20);
        size_t seen = 0;
        bool same = true;
        for (;;) {
            const long got = gz.read(back.data(), back.size());
            if (got < 0) { same = false; break; }
            if (got == 0) break;
            if (seen + (size_t)got > n || memcmp(back.data(), data + seen, (size_t)got) != 0) { same = false; break; }
            seen += (size_t)got;
        }
        if (!same || seen != n) {
            out.resize(start);
            return false;
        }
    }
    return true;
}

}  // namespace tb

// testing aid (include/trio_kmers_b200.h)
extern "C" int64_t tb_fastq_gzip(const void *data, int64_t n, void *out_buf, int64_t out_cap, int verify)
{
    std::string out;
    if (n < 0 || !tb::fastq_gzip_member((const char *)data, (size_t)n, out, verify != 0)) return -1;
    if ((int64_t)out.size() > out_cap) return -2;
    memcpy(out_buf, out.data(), out.size());
    return (int64_t)out.size();
}
