// gz_stream.cpp -- see gz_stream.h.  RFC 1951 (DEFLATE) + RFC 1952 (gzip), streaming.
#include "gz_stream.h"

#include <errno.h>
#include <fcntl.h>
#include <string.h>
#include <unistd.h>
#include <zlib.h>  // crc32() only

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace tb {

namespace {

constexpr size_t IN_CAP = 4u << 20;
constexpr size_t HIST = 32768;
constexpr size_t OUT_CHUNK = 2u << 20;
constexpr size_t OUT_SLACK = 320;  // a match (258) + 8-byte copy overshoot, with room to spare

constexpr unsigned LB = 11;  // primary bits, literal/length table
constexpr unsigned DB = 8;   // primary bits, distance table
constexpr uint32_t LM = (1u << LB) - 1, DM = (1u << DB) - 1;

// table entry: [31:16] payload | [15] literal | [14] second literal in payload bits 8..15 |
//              [13] link to sub-table | [12] exception: end of block, or (payload bit 0) invalid |
//              [11:8] extra bits (or sub-table bits) | [7:0] bits to consume
// FASTQ is literal after literal (bases and qualities barely repeat within 32 KB), and Huffman
// decoding is a serial chain -- table load, shift, next load -- of ~8 cycles per symbol.  So
// wherever two literal codes fit into the primary index together the entry carries both: one
// link of the chain then yields two bytes.
constexpr uint32_t F_LIT = 1u << 15, F_LIT2 = 1u << 14, F_LINK = 1u << 13, F_EXC = 1u << 12;
constexpr uint32_t F_EOB = F_EXC, F_BAD = F_EXC | (1u << 16);

const uint16_t LEN_BASE[29] = { 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59,
                                67, 83, 99, 115, 131, 163, 195, 227, 258 };
const uint8_t LEN_EXTRA[29] = { 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0 };
const uint16_t DIST_BASE[30] = { 1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769,
                                 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577 };
const uint8_t DIST_EXTRA[30] = { 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11,
                                 12, 12, 13, 13 };
const uint8_t CLEN_ORDER[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };

inline uint64_t load64(const unsigned char *p)
{
    uint64_t v;
    memcpy(&v, p, 8);
    return v;  // little-endian hosts only (x86-64, aarch64)
}
inline void store64(unsigned char *p, uint64_t v) { memcpy(p, &v, 8); }

// bit reversal of a code of `len` <= 15 bits through a byte table
struct Rev8 {
    uint8_t t[256];
    Rev8()
    {
        for (unsigned i = 0; i < 256; i++) {
            unsigned r = 0;
            for (unsigned b = 0; b < 8; b++) r |= ((i >> b) & 1u) << (7 - b);
            t[i] = (uint8_t)r;
        }
    }
};
inline uint32_t reverse_bits(uint32_t code, unsigned len)
{
    static const Rev8 rev;
    const uint32_t r16 = ((uint32_t)rev.t[code & 0xFF] << 8) | rev.t[(code >> 8) & 0xFF];
    return r16 >> (16 - len);
}

// Canonical Huffman decoding table with one level of sub-tables.  payload[sym] holds bits 8..31
// of the symbol's entry (flags, extra-bit count, value).  Returns false for an over-subscribed
// code; incomplete codes are accepted and their unused patterns decode as invalid.  A block of
// literal-heavy FASTQ is only ~60 KB long, so this runs thousands of times per second and is
// written to touch each table entry once.
bool build_table(const uint8_t *lens, unsigned n_syms, const uint32_t *payload, unsigned primary_bits,
                 uint32_t *table, size_t table_cap, bool pair_literals = false)
{
    unsigned count[16] = { 0 };
    for (unsigned s = 0; s < n_syms; s++) count[lens[s]]++;
    count[0] = 0;
    uint32_t next_code[17];
    uint32_t code = 0;
    int64_t left = 1;  // Kraft: patterns still free, scaled per length
    for (unsigned l = 1; l <= 15; l++) {
        left = left * 2 - (int64_t)count[l];
        if (left < 0) return false;
        code = (code + (l > 1 ? count[l - 1] : 0)) << 1;
        next_code[l] = code;
    }
    const size_t primary = (size_t)1 << primary_bits;
    if (left != 0)  // incomplete code: unused patterns must decode as invalid
        for (size_t i = 0; i < primary; i++) table[i] = F_BAD | 1u;
    // long codes (rare) first: sub-table width per primary prefix
    struct Long { uint16_t sym; uint16_t rev; };
    Long longs[288];
    unsigned n_long = 0;
    size_t next_free = primary;
    for (unsigned s = 0; s < n_syms; s++) {
        const unsigned l = lens[s];
        if (!l) continue;
        const uint32_t rev = reverse_bits(next_code[l]++, l);
        if (l <= primary_bits) {
            const uint32_t entry = payload[s] | l;
            for (size_t i = rev; i < primary; i += (size_t)1 << l) table[i] = entry;
        } else {
            longs[n_long].sym = (uint16_t)s;
            longs[n_long].rev = (uint16_t)rev;
            n_long++;
        }
    }
    if (n_long) {
        // canonical order: codes sharing a primary prefix are not contiguous in symbol order, so
        // size each prefix's sub-table first
        uint8_t sub_bits[1u << LB];
        memset(sub_bits, 0, primary);
        for (unsigned i = 0; i < n_long; i++) {
            const unsigned l = lens[longs[i].sym];
            uint8_t &b = sub_bits[longs[i].rev & (primary - 1)];
            if (l - primary_bits > b) b = (uint8_t)(l - primary_bits);
        }
        // allocate (a prefix is allocated when first seen; mark with the high bit of sub_bits)
        for (unsigned i = 0; i < n_long; i++) {
            const size_t p = longs[i].rev & (primary - 1);
            if (sub_bits[p] & 0x80) continue;
            const size_t n = (size_t)1 << sub_bits[p];
            if (next_free + n > table_cap) return false;
            table[p] = ((uint32_t)next_free << 16) | F_LINK | ((uint32_t)sub_bits[p] << 8) | primary_bits;
            for (size_t j = 0; j < n; j++) table[next_free + j] = F_BAD | (primary_bits + 1u);
            next_free += n;
            sub_bits[p] |= 0x80;
        }
        for (unsigned i = 0; i < n_long; i++) {
            const unsigned s = longs[i].sym, l = lens[s];
            const size_t p = longs[i].rev & (primary - 1);
            const size_t base = table[p] >> 16, n = (size_t)1 << (sub_bits[p] & 0x7F);
            const uint32_t entry = payload[s] | l;
            for (size_t j = longs[i].rev >> primary_bits; j < n; j += (size_t)1 << (l - primary_bits))
                table[base + j] = entry;
        }
    }
    if (pair_literals) {
        // second pass: where a literal leaves room in the index for a complete second literal
        // code, fuse them (the single-literal entries are read from a copy)
        static thread_local uint32_t single[1u << LB];
        memcpy(single, table, primary * sizeof(uint32_t));
        for (size_t i = 0; i < primary; i++) {
            const uint32_t e1 = single[i];
            if ((e1 & (F_LIT | F_LINK | F_EXC)) != F_LIT) continue;
            const unsigned l1 = e1 & 0xFF;
            const uint32_t e2 = single[i >> l1];
            const unsigned l2 = e2 & 0xFF;
            if ((e2 & (F_LIT | F_LINK | F_EXC)) != F_LIT || l1 + l2 > primary_bits) continue;
            table[i] = (e1 & 0x00FF0000u) | ((e2 & 0x00FF0000u) << 8) | F_LIT | F_LIT2 | (l1 + l2);
        }
    }
    return true;
}

// CRC-32 (the gzip one, reflected 0xEDB88320) by carry-less multiplication: four 128-bit lanes
// folded 64 bytes at a time, then reduced to 32 bits (Gopal et al., "Fast CRC computation for
// generic polynomials using PCLMULQDQ", Intel 2009; the constants are x^n mod P for that
// paper's fold distances).  ~10x the table-driven loop, which otherwise costs a fifth of the
// whole gunzip.  Checked against zlib's crc32() once at start-up; any disagreement (or a CPU
// without PCLMULQDQ) falls back to zlib's.
#if defined(__x86_64__)
__attribute__((target("pclmul,sse4.1"))) uint32_t crc32_clmul(const unsigned char *buf, size_t len, uint32_t crc)
{
    // len: multiple of 16, >= 64; crc: running remainder (already inverted)
    alignas(16) static const uint64_t k1k2[2] = { 0x0154442bd4ull, 0x01c6e41596ull };
    alignas(16) static const uint64_t k3k4[2] = { 0x01751997d0ull, 0x00ccaa009eull };
    alignas(16) static const uint64_t k5k0[2] = { 0x0163cd6124ull, 0x0000000000ull };
    alignas(16) static const uint64_t poly[2] = { 0x01db710641ull, 0x01f7011641ull };
    __m128i x0, x1, x2, x3, x4, x5, x6, x7, x8, y5, y6, y7, y8;
    x1 = _mm_loadu_si128((const __m128i *)(buf + 0x00));
    x2 = _mm_loadu_si128((const __m128i *)(buf + 0x10));
    x3 = _mm_loadu_si128((const __m128i *)(buf + 0x20));
    x4 = _mm_loadu_si128((const __m128i *)(buf + 0x30));
    x1 = _mm_xor_si128(x1, _mm_cvtsi32_si128((int)crc));
    x0 = _mm_load_si128((const __m128i *)k1k2);
    buf += 64;
    len -= 64;
    while (len >= 64) {
        x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
        x6 = _mm_clmulepi64_si128(x2, x0, 0x00);
        x7 = _mm_clmulepi64_si128(x3, x0, 0x00);
        x8 = _mm_clmulepi64_si128(x4, x0, 0x00);
        x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
        x2 = _mm_clmulepi64_si128(x2, x0, 0x11);
        x3 = _mm_clmulepi64_si128(x3, x0, 0x11);
        x4 = _mm_clmulepi64_si128(x4, x0, 0x11);
        y5 = _mm_loadu_si128((const __m128i *)(buf + 0x00));
        y6 = _mm_loadu_si128((const __m128i *)(buf + 0x10));
        y7 = _mm_loadu_si128((const __m128i *)(buf + 0x20));
        y8 = _mm_loadu_si128((const __m128i *)(buf + 0x30));
        x1 = _mm_xor_si128(_mm_xor_si128(x1, x5), y5);
        x2 = _mm_xor_si128(_mm_xor_si128(x2, x6), y6);
        x3 = _mm_xor_si128(_mm_xor_si128(x3, x7), y7);
        x4 = _mm_xor_si128(_mm_xor_si128(x4, x8), y8);
        buf += 64;
        len -= 64;
    }
    // four lanes -> one
    x0 = _mm_load_si128((const __m128i *)k3k4);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, x3), x5);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, x4), x5);
    while (len >= 16) {
        x2 = _mm_loadu_si128((const __m128i *)buf);
        x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
        x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
        x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
        buf += 16;
        len -= 16;
    }
    // 128 -> 64 bits
    x2 = _mm_clmulepi64_si128(x1, x0, 0x10);
    x3 = _mm_setr_epi32(~0, 0, ~0, 0);
    x1 = _mm_srli_si128(x1, 8);
    x1 = _mm_xor_si128(x1, x2);
    x0 = _mm_loadl_epi64((const __m128i *)k5k0);
    x2 = _mm_srli_si128(x1, 4);
    x1 = _mm_and_si128(x1, x3);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_xor_si128(x1, x2);
    // Barrett reduction to 32 bits
    x0 = _mm_load_si128((const __m128i *)poly);
    x2 = _mm_and_si128(x1, x3);
    x2 = _mm_clmulepi64_si128(x2, x0, 0x10);
    x2 = _mm_and_si128(x2, x3);
    x2 = _mm_clmulepi64_si128(x2, x0, 0x00);
    x1 = _mm_xor_si128(x1, x2);
    return (uint32_t)_mm_extract_epi32(x1, 1);
}

bool clmul_crc_usable()
{
    static const bool ok = [] {
        if (!__builtin_cpu_supports("pclmul") || !__builtin_cpu_supports("sse4.1")) return false;
        unsigned char probe[1024 + 7];
        uint32_t x = 0x9E3779B9u;
        for (size_t i = 0; i < sizeof probe; i++) {
            x = x * 1664525u + 1013904223u;
            probe[i] = (unsigned char)(x >> 24);
        }
        for (size_t off = 0; off < 3; off++)
            for (size_t n : { (size_t)64, (size_t)80, (size_t)256, (size_t)1008 }) {
                const uint32_t seed = 0x1234567u * (uint32_t)(n + off);
                if (~crc32_clmul(probe + off, n, ~seed) != (uint32_t)crc32(seed, probe + off, (uInt)n)) return false;
            }
        return true;
    }();
    return ok;
}
#endif

uint32_t fast_crc32(uint32_t crc, const unsigned char *p, size_t n)
{
#if defined(__x86_64__)
    if (n >= 64 && clmul_crc_usable()) {
        const size_t chunk = n & ~(size_t)15;
        crc = ~crc32_clmul(p, chunk, ~crc);
        p += chunk;
        n -= chunk;
    }
#endif
    while (n) {  // zlib takes a 32-bit length
        const size_t m = n < (1u << 30) ? n : (1u << 30);
        crc = (uint32_t)crc32(crc, p, (uInt)m);
        p += m;
        n -= m;
    }
    return crc;
}

struct Payloads {
    uint32_t lit[288], dist[32];
    Payloads()
    {
        for (unsigned s = 0; s < 256; s++) lit[s] = (s << 16) | F_LIT;
        lit[256] = F_EOB;
        for (unsigned s = 257; s < 286; s++)
            lit[s] = ((uint32_t)LEN_BASE[s - 257] << 16) | ((uint32_t)LEN_EXTRA[s - 257] << 8);
        lit[286] = lit[287] = F_BAD;  // codes that exist in the fixed code but never in valid data
        for (unsigned s = 0; s < 30; s++)
            dist[s] = ((uint32_t)DIST_BASE[s] << 16) | ((uint32_t)DIST_EXTRA[s] << 8);
        dist[30] = dist[31] = F_BAD;
    }
};
const Payloads &payloads()
{
    static const Payloads p;
    return p;
}

}  // namespace

uint32_t crc32_any(uint32_t crc, const unsigned char *p, size_t n) { return fast_crc32(crc, p, n); }
bool crc32_is_accelerated()
{
#if defined(__x86_64__)
    return clmul_crc_usable();
#else
    return false;
#endif
}

struct GzStream::Tables {
    uint32_t lit[(1u << LB) + 4608];
    uint32_t dist[(1u << DB) + 4096];
};

GzStream::GzStream() { payloads(); }

GzStream::~GzStream()
{
    if (fd_ >= 0 && own_fd_) ::close(fd_);
    delete tables_;
}

bool GzStream::fail(const char *what)
{
    err_ = what;
    mode_ = FAILED;
    return false;
}

bool GzStream::open(const char *path, std::string &err)
{
    fd_ = ::open(path, O_RDONLY);
    if (fd_ < 0) {
        err = std::string("cannot open ") + path;
        return false;
    }
    reset_state();
    return true;
}

// buffers are allocated once per object; a GzStream can be re-opened on another source
void GzStream::reset_state()
{
    if (in_.empty()) in_.assign(IN_CAP + 16, 0);
    if (out_.empty()) out_.assign(HIST + OUT_CHUNK + OUT_SLACK, 0);
    if (!tables_) tables_ = new Tables();
    eof_ = false;
    in_pos_ = in_len_ = 0;
    in_total_ = 0;
    out_pos_ = out_taken_ = out_base_ = 0;
    member_out_ = 0;
    crc_ = 0;
    crc_from_ = 0;
    bitbuf_ = 0;
    bitcnt_ = 0;
    final_block_ = false;
    stored_left_ = 0;
    mode_ = START;
    first_member_ = true;
    mem_ = nullptr;
    mem_left_ = 0;
    err_.clear();
}

bool GzStream::open_fd(int fd, bool at_start)
{
    reset_state();
    fd_ = fd;
    own_fd_ = false;
    if (!at_start) {
        first_member_ = false;
        mode_ = MEMBER_HEADER;
    }
    return true;
}

bool GzStream::open_memory(const unsigned char *data, size_t len)
{
    reset_state();
    mem_ = data;
    mem_left_ = len;
    first_member_ = false;
    mode_ = MEMBER_HEADER;
    return true;
}

bool GzStream::fill_input()
{
    if (eof_) return false;
    if (in_pos_ > 0) {
        memmove(in_.data(), in_.data() + in_pos_, in_len_ - in_pos_);
        in_len_ -= in_pos_;
        in_pos_ = 0;
    }
    bool got_any = false;
    if (mem_) {
        const size_t n = mem_left_ < IN_CAP - in_len_ ? mem_left_ : IN_CAP - in_len_;
        memcpy(in_.data() + in_len_, mem_, n);
        mem_ += n;
        mem_left_ -= n;
        in_len_ += n;
        in_total_ += n;
        got_any = n > 0;
        if (mem_left_ == 0) eof_ = true;
    }
    while (!mem_ && in_len_ < IN_CAP) {
        const ssize_t n = ::read(fd_, in_.data() + in_len_, IN_CAP - in_len_);
        if (n < 0) {
            if (errno == EINTR) continue;
            fail("read error on the compressed input");
            return false;
        }
        if (n == 0) {
            eof_ = true;
            break;
        }
        in_len_ += (size_t)n;
        in_total_ += (uint64_t)n;
        got_any = true;
    }
    memset(in_.data() + in_len_, 0, 16);
    return got_any;
}

void GzStream::slide_window()
{
    // everything before out_taken_ has been delivered; keep HIST bytes of history
    if (out_pos_ <= HIST) return;
    const size_t s = out_pos_ - HIST;
    if (crc_from_ < out_pos_ && mode_ != PASSTHROUGH) {
        crc_ = fast_crc32(crc_, out_.data() + crc_from_, out_pos_ - crc_from_);
        crc_from_ = out_pos_;
    }
    memmove(out_.data(), out_.data() + s, HIST);
    out_pos_ -= s;
    out_taken_ -= s;
    crc_from_ = crc_from_ > s ? crc_from_ - s : 0;
    if (s > out_base_) member_out_ += s - out_base_;  // member bytes that leave the buffer
    out_base_ = out_base_ > s ? out_base_ - s : 0;
}

// 1 = header parsed, 0 = need more input, -1 = not a gzip member / malformed
int GzStream::parse_member_header()
{
    const unsigned char *p = in_.data() + in_pos_;
    const size_t n = in_len_ - in_pos_;
    if (n < 10) return eof_ ? -1 : 0;
    if (p[0] != 0x1f || p[1] != 0x8b) return -1;
    if (p[2] != 8 || (p[3] & 0xE0)) return -1;
    const unsigned flg = p[3];
    size_t q = 10;
    if (flg & 4) {  // FEXTRA
        if (n < q + 2) return eof_ ? -1 : 0;
        const size_t xlen = p[q] | ((size_t)p[q + 1] << 8);
        q += 2;
        if (n < q + xlen) return eof_ ? -1 : 0;
        q += xlen;
    }
    for (unsigned bit = 8; bit <= 16; bit <<= 1) {  // FNAME, FCOMMENT: zero-terminated
        if (!(flg & bit)) continue;
        const void *z = memchr(p + q, 0, n - q);
        if (!z) return (eof_ || n >= IN_CAP) ? -1 : 0;
        q = (size_t)((const unsigned char *)z - p) + 1;
    }
    if (flg & 2) {  // FHCRC
        if (n < q + 2) return eof_ ? -1 : 0;
        q += 2;
    }
    in_pos_ += q;
    return 1;
}

// careful bit reader for block headers (rare, small): returns false when the input ran out
#define NEED_BITS(nb)                                                        \
    do {                                                                     \
        while (bc < (unsigned)(nb)) {                                        \
            if (in >= in_end) return eof_ ? (fail("truncated deflate stream"), -1) : 0; \
            bb |= (uint64_t)*in++ << bc;                                     \
            bc += 8;                                                         \
        }                                                                    \
    } while (0)
#define TAKE_BITS(var, nb)                         \
    do {                                           \
        NEED_BITS(nb);                             \
        var = (unsigned)(bb & ((1ull << (nb)) - 1)); \
        bb >>= (nb);                               \
        bc -= (nb);                                \
    } while (0)

int GzStream::parse_block_header()
{
    // a dynamic header is at most ~600 bytes; ask for them up front so that a mid-header
    // "need more input" can simply restart from the saved state
    if (!eof_ && in_len_ - in_pos_ < 1024) return 0;
    const unsigned char *in = in_.data() + in_pos_, *in_end = in_.data() + in_len_;
    uint64_t bb = bitbuf_;
    unsigned bc = bitcnt_;
    // the fast path leaves valid look-ahead bits above bc; drop them, the careful reader
    // re-reads those bytes
    bb &= bc ? ((~0ull) >> (64 - bc)) : 0ull;
    unsigned final_bit, type;
    TAKE_BITS(final_bit, 1);
    TAKE_BITS(type, 2);
    final_block_ = final_bit != 0;
    if (type == 0) {
        // stored: discard to the byte boundary, give whole bytes back, read LEN / NLEN
        bb >>= bc & 7;
        bc -= bc & 7;
        in -= bc >> 3;
        bb = 0;
        bc = 0;
        if (in_end - in < 4) return eof_ ? (fail("truncated stored block"), -1) : 0;
        const unsigned len = in[0] | (in[1] << 8), nlen = in[2] | (in[3] << 8);
        if ((len ^ nlen) != 0xFFFF) return fail("stored block length check failed"), -1;
        in += 4;
        stored_left_ = len;
        mode_ = STORED;
    } else if (type == 1) {
        uint8_t lens[288];
        for (unsigned s = 0; s < 144; s++) lens[s] = 8;
        for (unsigned s = 144; s < 256; s++) lens[s] = 9;
        for (unsigned s = 256; s < 280; s++) lens[s] = 7;
        for (unsigned s = 280; s < 288; s++) lens[s] = 8;
        uint8_t dl[32];
        memset(dl, 5, 32);
        if (!build_table(lens, 288, payloads().lit, LB, tables_->lit, sizeof(tables_->lit) / 4, true) ||
            !build_table(dl, 32, payloads().dist, DB, tables_->dist, sizeof(tables_->dist) / 4))
            return fail("internal: fixed tables"), -1;
        mode_ = HUFFMAN;
    } else if (type == 2) {
        unsigned hlit, hdist, hclen;
        TAKE_BITS(hlit, 5);
        TAKE_BITS(hdist, 5);
        TAKE_BITS(hclen, 4);
        hlit += 257;
        hdist += 1;
        hclen += 4;
        if (hlit > 286 || hdist > 30) return fail("too many length or distance symbols"), -1;
        uint8_t cl[19];
        memset(cl, 0, 19);
        for (unsigned i = 0; i < hclen; i++) {
            unsigned v;
            TAKE_BITS(v, 3);
            cl[CLEN_ORDER[i]] = (uint8_t)v;
        }
        uint32_t cl_payload[19], cl_table[1u << 7];
        for (unsigned s = 0; s < 19; s++) cl_payload[s] = s << 16;
        if (!build_table(cl, 19, cl_payload, 7, cl_table, 1u << 7)) return fail("invalid code-length code"), -1;
        uint8_t lens[320];
        unsigned n = 0;
        while (n < hlit + hdist) {
            NEED_BITS(7);
            const uint32_t e = cl_table[bb & 127];
            if (e & F_EXC) return fail("invalid code-length symbol"), -1;
            bb >>= (e & 63);
            bc -= (e & 0xFF);
            const unsigned sym = e >> 16;
            if (sym < 16) {
                lens[n++] = (uint8_t)sym;
                continue;
            }
            unsigned rep, val = 0;
            if (sym == 16) {
                if (n == 0) return fail("repeat without a previous length"), -1;
                val = lens[n - 1];
                TAKE_BITS(rep, 2);
                rep += 3;
            } else if (sym == 17) {
                TAKE_BITS(rep, 3);
                rep += 3;
            } else {
                TAKE_BITS(rep, 7);
                rep += 11;
            }
            if (n + rep > hlit + hdist) return fail("code lengths overrun"), -1;
            while (rep--) lens[n++] = (uint8_t)val;
        }
        if (lens[256] == 0) return fail("no end-of-block code"), -1;
        uint8_t ll[288], dl[32];
        memset(ll, 0, sizeof ll);
        memset(dl, 0, sizeof dl);
        memcpy(ll, lens, hlit);
        memcpy(dl, lens + hlit, hdist);
        if (!build_table(ll, 288, payloads().lit, LB, tables_->lit, sizeof(tables_->lit) / 4, true))
            return fail("invalid literal/length code"), -1;
        if (!build_table(dl, 32, payloads().dist, DB, tables_->dist, sizeof(tables_->dist) / 4))
            return fail("invalid distance code"), -1;
        mode_ = HUFFMAN;
    } else {
        return fail("invalid block type"), -1;
    }
    in_pos_ = (size_t)(in - in_.data());
    bitbuf_ = bb;
    bitcnt_ = bc;
    return 1;
}

// Decode symbols of the current Huffman block.  1 = end of block reached, 0 = stopped because
// the output chunk is full or the input ran low (caller refills), -1 = error.
int GzStream::decode_huffman()
{
    const uint32_t *lt = tables_->lit, *dt = tables_->dist;
    const unsigned char *in = in_.data() + in_pos_, *const in_end = in_.data() + in_len_;
    unsigned char *out = out_.data() + out_pos_;
    unsigned char *const out_limit = out_.data() + HIST + OUT_CHUNK;
    const unsigned char *const hist_lo = out_.data() + out_base_;
    uint64_t bb = bitbuf_;
    unsigned bc = bitcnt_;
    int rc = 0;
    const char *problem = nullptr;

#define LOOKUP_LIT(e)                                                                        \
    do {                                                                                     \
        e = lt[bb & LM];                                                                     \
        if (e & F_LINK) e = lt[(e >> 16) + ((bb >> LB) & ((1u << ((e >> 8) & 15)) - 1))];     \
    } while (0)
#define LOOKUP_DIST(e)                                                                       \
    do {                                                                                     \
        e = dt[bb & DM];                                                                     \
        if (e & F_LINK) e = dt[(e >> 16) + ((bb >> DB) & ((1u << ((e >> 8) & 15)) - 1))];     \
    } while (0)
#define REFILL_FAST()                     \
    do {                                  \
        bb |= load64(in) << bc;           \
        in += (63 - bc) >> 3;             \
        bc |= 56;                         \
    } while (0)

    for (;;) {
        if (out >= out_limit) break;
        if (in_end - in >= 16) {
            // ---- fast iteration: up to three literals, or literals + one match, per pass
            uint32_t e;
            REFILL_FAST();
            LOOKUP_LIT(e);
            if (e & F_LIT) {
                // one or two literals per entry; the second byte of the store is scratch for singles
                bb >>= (e & 63);
                bc -= (e & 0xFF);
                out[0] = (unsigned char)(e >> 16);
                out[1] = (unsigned char)(e >> 24);
                out += 1 + ((e >> 14) & 1);
                LOOKUP_LIT(e);
                if (e & F_LIT) {
                    bb >>= (e & 63);
                    bc -= (e & 0xFF);
                    out[0] = (unsigned char)(e >> 16);
                    out[1] = (unsigned char)(e >> 24);
                    out += 1 + ((e >> 14) & 1);
                    LOOKUP_LIT(e);
                    if (e & F_LIT) {
                        bb >>= (e & 63);
                        bc -= (e & 0xFF);
                        out[0] = (unsigned char)(e >> 16);
                        out[1] = (unsigned char)(e >> 24);
                        out += 1 + ((e >> 14) & 1);
                        continue;
                    }
                }
                REFILL_FAST();  // e was looked up from bits that a refill never changes
            }
            bb >>= (e & 63);
            bc -= (e & 0xFF);
            if (e & F_EXC) {
                const bool bad = (e & F_BAD) == F_BAD;
                if (bad) problem = "invalid literal/length code in the stream";
                rc = bad ? -1 : 1;
                break;
            }
            unsigned x = (e >> 8) & 15;
            unsigned length = (e >> 16) + (unsigned)(bb & ((1u << x) - 1));
            bb >>= x;
            bc -= x;
            uint32_t d;
            LOOKUP_DIST(d);
            if (d & F_EXC) {
                problem = "invalid distance code in the stream";
                rc = -1;
                break;
            }
            bb >>= (d & 63);
            bc -= (d & 0xFF);
            x = (d >> 8) & 15;
            const size_t dist = (d >> 16) + (size_t)(bb & ((1u << x) - 1));
            bb >>= x;
            bc -= x;
            if (dist > (size_t)(out - hist_lo)) {
                problem = "match distance reaches before the start of the member";
                rc = -1;
                break;
            }
            const unsigned char *src = out - dist;
            unsigned char *dst = out;
            out += length;
            if (dist >= 8) {
                do {
                    store64(dst, load64(src));
                    dst += 8;
                    src += 8;
                } while (dst < out);
            } else if (dist == 1) {
                memset(dst, *src, length);
            } else {
                do {
                    *dst++ = *src++;
                } while (dst < out);
            }
            continue;
        }
        // ---- tail of the input: one symbol at a time, every bit accounted for
        if (!eof_) break;  // let the caller refill first
        {
            // drop look-ahead bits above bc (the careful reader re-reads those bytes)
            bb &= bc ? ((~0ull) >> (64 - bc)) : 0ull;
            while (bc <= 56 && in < in_end) {
                bb |= (uint64_t)*in++ << bc;
                bc += 8;
            }
            uint32_t e;
            LOOKUP_LIT(e);
            if ((e & 0xFF) > bc) {
                problem = "truncated deflate stream";
                rc = -1;
                break;
            }
            bb >>= (e & 63);
            bc -= (e & 0xFF);
            if (e & F_LIT) {
                *out++ = (unsigned char)(e >> 16);
                if (e & F_LIT2) *out++ = (unsigned char)(e >> 24);
                continue;
            }
            if (e & F_EXC) {
                const bool bad = (e & F_BAD) == F_BAD;
                if (bad) problem = "invalid literal/length code in the stream";
                rc = bad ? -1 : 1;
                break;
            }
            unsigned x = (e >> 8) & 15;
            if (x > bc) {
                problem = "truncated deflate stream";
                rc = -1;
                break;
            }
            const unsigned length = (e >> 16) + (unsigned)(bb & ((1u << x) - 1));
            bb >>= x;
            bc -= x;
            while (bc <= 56 && in < in_end) {
                bb |= (uint64_t)*in++ << bc;
                bc += 8;
            }
            uint32_t d;
            LOOKUP_DIST(d);
            x = (d >> 8) & 15;
            if ((d & F_EXC) || (d & 0xFF) + x > bc) {
                problem = (d & F_EXC) ? "invalid distance code in the stream" : "truncated deflate stream";
                rc = -1;
                break;
            }
            bb >>= (d & 63);
            bc -= (d & 0xFF);
            const size_t dist = (d >> 16) + (size_t)(bb & ((1u << x) - 1));
            bb >>= x;
            bc -= x;
            if (dist > (size_t)(out - hist_lo)) {
                problem = "match distance reaches before the start of the member";
                rc = -1;
                break;
            }
            const unsigned char *src = out - dist;
            for (unsigned i = 0; i < length; i++) out[i] = src[i];
            out += length;
        }
    }
#undef LOOKUP_LIT
#undef LOOKUP_DIST
#undef REFILL_FAST
    in_pos_ = (size_t)(in - in_.data());
    out_pos_ = (size_t)(out - out_.data());
    bitbuf_ = bb;
    bitcnt_ = bc;
    if (rc < 0) fail(problem ? problem : "corrupt deflate stream");
    return rc;
}

long GzStream::read(char *dst, size_t cap)
{
    if (cap == 0) return 0;
    for (;;) {
        if (out_taken_ < out_pos_) {
            const size_t n = out_pos_ - out_taken_ < cap ? out_pos_ - out_taken_ : cap;
            memcpy(dst, out_.data() + out_taken_, n);
            out_taken_ += n;
            return (long)n;
        }
        if (mode_ == DONE) return 0;
        if (mode_ == FAILED) return -1;
        slide_window();
        const size_t out_cap = HIST + OUT_CHUNK;
        // run the state machine until the chunk is full or the stream ends
        while (out_pos_ < out_cap && mode_ != DONE && mode_ != FAILED) {
            switch (mode_) {
            case START:
                if (in_len_ - in_pos_ < 2 && !eof_) {
                    fill_input();
                    if (mode_ == FAILED) break;
                    if (in_len_ - in_pos_ < 2 && !eof_) continue;
                }
                if (in_len_ - in_pos_ >= 2 && in_[in_pos_] == 0x1f && in_[in_pos_ + 1] == 0x8b)
                    mode_ = MEMBER_HEADER;
                else if (in_len_ == in_pos_)
                    mode_ = DONE;  // an empty file reads as empty (Python's gzip does the same)
                else if (lenient_)
                    mode_ = PASSTHROUGH;  // gzread semantics: not gzip -> as is
                else
                    fail("not a gzip file (no gzip magic at the start)");  // gzip.BadGzipFile upstream
                break;
            case PASSTHROUGH: {
                if (in_pos_ == in_len_) {
                    if (!fill_input()) {
                        if (mode_ != FAILED && eof_) mode_ = DONE;
                        break;
                    }
                }
                size_t n = in_len_ - in_pos_;
                if (n > out_cap - out_pos_) n = out_cap - out_pos_;
                memcpy(out_.data() + out_pos_, in_.data() + in_pos_, n);
                out_pos_ += n;
                in_pos_ += n;
                break;
            }
            case MEMBER_HEADER: {
                if (!first_member_ && !lenient_) {
                    // Python's gzip (what the reference reads with, seq.py:88-92) skips zero
                    // padding after a member; anything else must be another member
                    while (in_pos_ < in_len_ && in_[in_pos_] == 0) in_pos_++;
                    if (in_pos_ == in_len_) {
                        if (!fill_input() && mode_ != FAILED && eof_) mode_ = DONE;
                        break;
                    }
                }
                const int r = parse_member_header();
                if (r == 0) {
                    if (!fill_input() && mode_ != FAILED && eof_ && in_len_ - in_pos_ < 10) {
                        // nothing (or less than a header) left
                        if (first_member_) fail("truncated gzip header");
                        else if (lenient_ || in_len_ == in_pos_) mode_ = DONE;
                        else fail("bytes after the last gzip member are not gzip data");
                    }
                    break;
                }
                if (r < 0) {
                    if (first_member_) fail("malformed gzip header");
                    else if (lenient_) mode_ = DONE;  // gzread: trailing non-member bytes are ignored
                    else fail("bytes after a gzip member are not another gzip member");
                    break;
                }
                crc_ = 0;
                crc_from_ = out_pos_;
                out_base_ = out_pos_;
                member_out_ = 0;
                bitbuf_ = 0;
                bitcnt_ = 0;
                mode_ = BLOCK_HEADER;
                break;
            }
            case BLOCK_HEADER: {
                const int r = parse_block_header();
                if (r == 0 && !fill_input() && mode_ != FAILED && eof_) {
                    // eof_ is set now: the next call parses to the end or reports truncation
                    continue;
                }
                break;
            }
            case STORED: {
                if (stored_left_ == 0) {
                    mode_ = final_block_ ? TRAILER : BLOCK_HEADER;
                    break;
                }
                if (in_pos_ == in_len_ && !fill_input()) {
                    if (mode_ != FAILED) fail("truncated stored block");
                    break;
                }
                size_t n = in_len_ - in_pos_;
                if (n > stored_left_) n = stored_left_;
                if (n > out_cap - out_pos_) n = out_cap - out_pos_;
                memcpy(out_.data() + out_pos_, in_.data() + in_pos_, n);
                out_pos_ += n;
                in_pos_ += n;
                stored_left_ -= (uint32_t)n;
                break;
            }
            case HUFFMAN: {
                const int r = decode_huffman();
                if (r < 0) break;
                if (r == 1) {
                    mode_ = final_block_ ? TRAILER : BLOCK_HEADER;
                    break;
                }
                if (out_pos_ >= out_cap) break;  // chunk full
                fill_input();                    // input ran low: more bytes, or eof_ becomes true
                break;
            }
            case TRAILER: {
                // back to the byte boundary: whole bytes still in the bit buffer go back
                in_pos_ -= bitcnt_ >> 3;
                bitbuf_ = 0;
                bitcnt_ = 0;
                if (in_len_ - in_pos_ < 8) {
                    fill_input();
                    if (mode_ == FAILED) break;
                    if (in_len_ - in_pos_ < 8) {
                        if (eof_) fail("truncated gzip trailer");
                        break;
                    }
                }
                crc_ = fast_crc32(crc_, out_.data() + crc_from_, out_pos_ - crc_from_);
                const unsigned char *t = in_.data() + in_pos_;
                const uint32_t want_crc = t[0] | (t[1] << 8) | (t[2] << 16) | ((uint32_t)t[3] << 24);
                const uint32_t want_len = t[4] | (t[5] << 8) | (t[6] << 16) | ((uint32_t)t[7] << 24);
                in_pos_ += 8;
                if (want_crc != crc_) {
                    fail("gzip CRC-32 mismatch");
                    break;
                }
                if (want_len != (uint32_t)(member_out_ + (out_pos_ - out_base_))) {
                    fail("gzip length mismatch");
                    break;
                }
                first_member_ = false;
                crc_from_ = out_pos_;
                mode_ = MEMBER_HEADER;
                break;
            }
            default:
                break;
            }
        }
        if (mode_ == FAILED) {
            // deliver nothing after an error: partial members are not results
            out_taken_ = out_pos_;
            return -1;
        }
        if (out_taken_ == out_pos_ && mode_ == DONE) return 0;
    }
}

// ---------------------------------------------------------------------------------------------
// SpanInflate: DEFLATE from the middle of a stream, 16-bit symbols with window markers
// (gz_stream.h has the scheme).  Slower per byte than GzStream (wider output, one careful
// accounting of the bits per symbol) -- it exists to run on several cores at once.
// ---------------------------------------------------------------------------------------------
namespace {

struct BitIn {
    const unsigned char *base = nullptr, *p = nullptr, *end = nullptr;
    uint64_t bb = 0;
    unsigned bc = 0;
    inline void refill()
    {
        if (end - p >= 8) {
            bb |= load64(p) << bc;  // bits above bc are look-ahead of the same stream: OR-ing them again is harmless
            p += (63 - bc) >> 3;
            bc |= 56;
        } else {
            while (bc <= 56 && p < end) {
                bb |= (uint64_t)*p++ << bc;
                bc += 8;
            }
        }
    }
    void seek(uint64_t bit)
    {
        p = base + (bit >> 3);
        bb = 0;
        bc = 0;
        refill();
        const unsigned d = (unsigned)(bit & 7);
        if (bc >= d) {
            bb >>= d;
            bc -= d;
        } else {
            bc = 0;
        }
    }
    uint64_t bit_pos() const { return (uint64_t)(p - base) * 8 - bc; }
    // drop the look-ahead and go back to whole bytes (stored blocks, trailers); returns the byte pointer
    const unsigned char *align_to_byte()
    {
        const uint64_t bit = (bit_pos() + 7) & ~7ull;
        p = base + (bit >> 3);
        bb = 0;
        bc = 0;
        return p;
    }
};

// Kraft sum of a set of code lengths: 0 complete, > 0 incomplete, < 0 over-subscribed
int64_t kraft_left(const uint8_t *lens, unsigned n, unsigned *n_used)
{
    unsigned count[16] = { 0 };
    unsigned used = 0;
    for (unsigned s = 0; s < n; s++) {
        count[lens[s]]++;
        used += lens[s] != 0;
    }
    if (n_used) *n_used = used;
    int64_t left = 1;
    for (unsigned l = 1; l <= 15; l++) {
        left = left * 2 - (int64_t)count[l];
        if (left < 0) return -1;
    }
    return left;
}

}  // namespace

struct SpanInflate::Tables {
    uint32_t lit[(1u << LB) + 4608];
    uint32_t dist[(1u << DB) + 4096];
};

SpanInflate::SpanInflate() : t_(new Tables())
{
    payloads();
}
SpanInflate::~SpanInflate() { delete t_; }

size_t SpanInflate::member_header_length(const unsigned char *p, size_t n)
{
    if (n < 10 || p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || (p[3] & 0xE0)) return 0;
    const unsigned flg = p[3];
    size_t q = 10;
    if (flg & 4) {
        if (n < q + 2) return 0;
        const size_t xlen = p[q] | ((size_t)p[q + 1] << 8);
        q += 2;
        if (n < q + xlen) return 0;
        q += xlen;
    }
    for (unsigned bit = 8; bit <= 16; bit <<= 1) {
        if (!(flg & bit)) continue;
        const void *z = memchr(p + q, 0, n - q);
        if (!z) return 0;
        q = (size_t)((const unsigned char *)z - p) + 1;
    }
    if (flg & 2) {
        if (n < q + 2) return 0;
        q += 2;
    }
    return q;
}

namespace {

inline void narrow_slow(const uint16_t *sym, size_t from, size_t to, unsigned char *out, std::vector<uint32_t> &marker_pos,
                        unsigned &any_or, unsigned &any_cr)
{
    for (size_t j = from; j < to; j++) {
        const unsigned s = sym[j];
        if (s < 256) {
            out[j] = (unsigned char)s;
            any_or |= s;
            any_cr |= s == '\r';
        } else {
            out[j] = 0;
            marker_pos.push_back((uint32_t)j);
        }
    }
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) void narrow_avx2(const uint16_t *sym, size_t n, unsigned char *out,
                                                 std::vector<uint32_t> &marker_pos, unsigned &any_or, unsigned &any_cr)
{
    __m256i acc_or = _mm256_setzero_si256(), acc_cr = _mm256_setzero_si256();
    const __m256i cr = _mm256_set1_epi8('\r');
    size_t i = 0;
    for (; i + 32 <= n; i += 32) {
        const __m256i a = _mm256_loadu_si256((const __m256i *)(sym + i)), b = _mm256_loadu_si256((const __m256i *)(sym + i + 16));
        if (!_mm256_testz_si256(_mm256_or_si256(a, b), _mm256_set1_epi16((short)0xFF00))) {
            narrow_slow(sym, i, i + 32, out, marker_pos, any_or, any_cr);
            continue;
        }
        // thirty-two plain bytes: pack (per 128-bit lane), then put the lanes back in order
        const __m256i v = _mm256_permute4x64_epi64(_mm256_packus_epi16(a, b), 0xD8);
        _mm256_storeu_si256((__m256i *)(out + i), v);
        acc_or = _mm256_or_si256(acc_or, v);
        acc_cr = _mm256_or_si256(acc_cr, _mm256_cmpeq_epi8(v, cr));
    }
    narrow_slow(sym, i, n, out, marker_pos, any_or, any_cr);
    any_or |= _mm256_movemask_epi8(acc_or) ? 0x80u : 0u;  // only the high bits matter (ASCII or not)
    any_cr |= _mm256_movemask_epi8(acc_cr) ? 1u : 0u;
}
#endif

}  // namespace

void narrow_symbols(const uint16_t *sym, size_t n, unsigned char *out, std::vector<uint32_t> &marker_pos, bool &has_cr,
                    bool &ascii)
{
    unsigned any_or = 0, any_cr = 0;
    marker_pos.clear();
    if (n >> 32) {  // positions are 32 bits wide; no span comes near
        for (size_t j = 0; j < n; j++) out[j] = 0;
        has_cr = true;
        ascii = false;
        return;
    }
#if defined(__x86_64__)
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2) narrow_avx2(sym, n, out, marker_pos, any_or, any_cr);
    else
#endif
        narrow_slow(sym, 0, n, out, marker_pos, any_or, any_cr);
    has_cr = any_cr != 0;
    ascii = (any_or & 0x80u) == 0;
}

namespace {

// Dynamic block header (after the three header bits) into the tables.  strict: only what a real
// encoder writes is accepted (complete codes) -- the filter of the start search.
// 1 ok, -1 invalid, -2 input ends inside it
int read_dynamic_header(BitIn &in, bool strict, uint32_t *lit_table, size_t lit_cap, uint32_t *dist_table, size_t dist_cap,
                        const char *&problem)
{
#define S_TAKE(var, nb)                                   \
    do {                                                  \
        if (in.bc < (unsigned)(nb)) in.refill();          \
        if (in.bc < (unsigned)(nb)) return -2;            \
        var = (unsigned)(in.bb & ((1ull << (nb)) - 1));   \
        in.bb >>= (nb);                                   \
        in.bc -= (nb);                                    \
    } while (0)
    unsigned hlit, hdist, hclen;
    S_TAKE(hlit, 5);
    S_TAKE(hdist, 5);
    S_TAKE(hclen, 4);
    hlit += 257;
    hdist += 1;
    hclen += 4;
    if (hlit > 286 || hdist > 30) { problem = "too many length or distance symbols"; return -1; }
    uint8_t cl[19];
    memset(cl, 0, 19);
    for (unsigned i = 0; i < hclen; i++) {
        unsigned v;
        S_TAKE(v, 3);
        cl[CLEN_ORDER[i]] = (uint8_t)v;
    }
    if (strict && kraft_left(cl, 19, nullptr) != 0) return -1;
    uint32_t cl_payload[19], cl_table[1u << 7];
    for (unsigned s = 0; s < 19; s++) cl_payload[s] = s << 16;
    if (!build_table(cl, 19, cl_payload, 7, cl_table, 1u << 7)) { problem = "invalid code-length code"; return -1; }
    uint8_t lens[320];
    unsigned n = 0;
    while (n < hlit + hdist) {
        if (in.bc < 14) in.refill();
        const uint32_t e = cl_table[in.bb & 127];
        if (e & F_EXC) { problem = "invalid code-length symbol"; return -1; }
        if ((e & 0xFF) > in.bc) return -2;
        in.bb >>= (e & 63);
        in.bc -= (e & 0xFF);
        const unsigned sym = e >> 16;
        if (sym < 16) {
            lens[n++] = (uint8_t)sym;
            continue;
        }
        unsigned rep, val = 0;
        if (sym == 16) {
            if (n == 0) { problem = "repeat without a previous length"; return -1; }
            val = lens[n - 1];
            S_TAKE(rep, 2);
            rep += 3;
        } else if (sym == 17) {
            S_TAKE(rep, 3);
            rep += 3;
        } else {
            S_TAKE(rep, 7);
            rep += 11;
        }
        if (n + rep > hlit + hdist) { problem = "code lengths overrun"; return -1; }
        while (rep--) lens[n++] = (uint8_t)val;
    }
#undef S_TAKE
    if (lens[256] == 0) { problem = "no end-of-block code"; return -1; }
    uint8_t ll[288], dl[32];
    memset(ll, 0, sizeof ll);
    memset(dl, 0, sizeof dl);
    memcpy(ll, lens, hlit);
    memcpy(dl, lens + hlit, hdist);
    if (strict) {
        if (kraft_left(ll, 288, nullptr) != 0) return -1;
        unsigned used = 0;
        const int64_t left = kraft_left(dl, 32, &used);
        // zlib writes a complete distance code, or a single code of one bit when at most one distance is used
        if (left < 0 || (left != 0 && !(used <= 1))) return -1;
    }
    if (!build_table(ll, 288, payloads().lit, LB, lit_table, lit_cap, true)) { problem = "invalid literal/length code"; return -1; }
    if (!build_table(dl, 32, payloads().dist, DB, dist_table, dist_cap)) { problem = "invalid distance code"; return -1; }
    return 1;
}

// Symbols of one Huffman block appended to sym[0..n).  min_pos: lowest index a match may reach.
// 1 end of block, -1 error (problem set), -2 more than `limit` symbols without an end of block
template <class T>  // uint16_t: bytes and window markers; char: bytes only (no marker within reach any more)
int decode_block_symbols(BitIn &in, const uint32_t *lt, const uint32_t *dt, std::vector<T> &sym, size_t &n_io,
                         size_t min_pos, size_t limit, const char *&problem)
{
    size_t n = n_io;
    T *o = sym.data();
    size_t cap = sym.size();
    int rc = 0;
#define S_LOOKUP(e, table, PB, PM)                                                              \
    do {                                                                                        \
        e = table[in.bb & PM];                                                                  \
        if (e & F_LINK) e = table[(e >> 16) + ((in.bb >> PB) & ((1u << ((e >> 8) & 15)) - 1))];  \
    } while (0)
#define S_CONSUME(nb)                                                               \
    {                                                                               \
        const unsigned nb_ = (nb);                                                  \
        if (nb_ > in.bc) { problem = "truncated deflate stream"; rc = -1; break; }  \
        in.bb >>= nb_;                                                              \
        in.bc -= nb_;                                                               \
    }
#define S_LITERALS(e)                              \
    S_CONSUME(e & 0xFF);                           \
    o[n] = (T)((e >> 16) & 0xFF);                  \
    o[n + 1] = (T)(e >> 24);                       \
    n += 1 + ((e >> 14) & 1);
    for (;;) {
        if (n + 600 > cap) {
            if (n > limit) { rc = -2; break; }
            sym.resize(cap + cap / 2 + 65536);
            o = sym.data();
            cap = sym.size();
        }
        in.refill();
        uint32_t e;
        S_LOOKUP(e, lt, LB, LM);
        if (e & F_LIT) {
            // up to three table entries (six literals) per refill: 3 x 15 bits <= 56
            S_LITERALS(e);
            S_LOOKUP(e, lt, LB, LM);
            if (e & F_LIT) {
                S_LITERALS(e);
                S_LOOKUP(e, lt, LB, LM);
                if (e & F_LIT) {
                    S_LITERALS(e);
                    continue;
                }
            }
            in.refill();
        }
        S_CONSUME(e & 0xFF);
        if (e & F_EXC) {
            if ((e & F_BAD) == F_BAD) { problem = "invalid literal/length code in the stream"; rc = -1; }
            else rc = 1;
            break;
        }
        // a match: <= 15 + 5 bits of length were in reach, <= 15 + 13 of distance follow (48 <= 56)
        unsigned x = (e >> 8) & 15;
        const unsigned length = (e >> 16) + (unsigned)(in.bb & ((1u << x) - 1));
        S_CONSUME(x);
        uint32_t d;
        S_LOOKUP(d, dt, DB, DM);
        if (d & F_EXC) { problem = "invalid distance code in the stream"; rc = -1; break; }
        S_CONSUME(d & 0xFF);
        x = (d >> 8) & 15;
        const size_t dist = (d >> 16) + (size_t)(in.bb & ((1u << x) - 1));
        S_CONSUME(x);
        if (dist > n - min_pos) { problem = "match distance reaches before the start of the member"; rc = -1; break; }
        const T *src = o + n - dist;
        T *dst = o + n;
        if (dist >= length) memcpy(dst, src, (size_t)length * sizeof(T));
        else if (sizeof(T) == 1 && dist == 1) memset(dst, (int)(unsigned char)*src, length);
        else for (unsigned i = 0; i < length; i++) dst[i] = src[i];
        n += length;
    }
#undef S_LOOKUP
#undef S_CONSUME
#undef S_LITERALS
    n_io = n;
    return rc;
}

// symbols one span may decode to (128 MB of them): a megabyte of FASTQ is 4-5 MB of text
constexpr size_t kMaxSpanSymbols = SpanResult::kWindow + (64u << 20);

// index just past the last marker among sym[from..to), or `before` if there is none (coarse: to the
// end of the group of 32 symbols the marker sits in)
size_t last_marker_end(const uint16_t *sym, size_t from, size_t to, size_t before)
{
    size_t last = before, i = from;
#if defined(__x86_64__)
    for (; i + 16 <= to; i += 16) {
        const __m128i a = _mm_loadu_si128((const __m128i *)(sym + i)), b = _mm_loadu_si128((const __m128i *)(sym + i + 8));
        const __m128i hi = _mm_or_si128(_mm_srli_epi16(a, 8), _mm_srli_epi16(b, 8));
        if (_mm_movemask_epi8(_mm_cmpeq_epi8(hi, _mm_setzero_si128())) != 0xFFFF) last = i + 16;
    }
#endif
    for (; i < to; i++)
        if (sym[i] >= 256) last = i + 1;
    return last;
}

void fixed_code_lengths(uint8_t *ll, uint8_t *dl)
{
    for (unsigned s = 0; s < 144; s++) ll[s] = 8;
    for (unsigned s = 144; s < 256; s++) ll[s] = 9;
    for (unsigned s = 256; s < 280; s++) ll[s] = 7;
    for (unsigned s = 280; s < 288; s++) ll[s] = 8;
    memset(dl, 5, 32);
}

}  // namespace

void SpanInflate::decode(const unsigned char *file, uint64_t size, uint64_t start_bit, bool search, uint64_t end_bit,
                         bool fresh_member, SpanResult &r)
{
    constexpr size_t W = SpanResult::kWindow;
    r.start_bit = r.stop_bit = start_bit;
    r.stop_fresh = false;
    r.ends.clear();
    r.tail = false;
    r.tail_off = 0;
    r.found = false;
    r.error.clear();
    if (r.sym.size() < W + (1u << 20)) r.sym.resize(W + (1u << 20));
    for (size_t j = 0; j < W; j++) r.sym[j] = (uint16_t)(256 + j);
    r.n_sym = W;
    const uint64_t total_bits = size * 8;
    BitIn in;
    in.base = file;
    in.end = file + size;
    const char *problem = nullptr;
    size_t min_pos = fresh_member ? W : 0;

    if (search) {
        // the first bit position where a dynamic block header passes the strict checks, the block
        // decodes to its end and something that can be a block header follows
        const size_t trial_limit = W + (8u << 20);
        uint64_t b = start_bit;
        for (; b < end_bit && b + 64 < total_bits; b++) {
            const uint64_t peek = load64(file + (b >> 3)) >> (b & 7);  // >= 56 bits
            if (((peek >> 1) & 3) != 2) continue;                       // BTYPE
            if (((peek >> 3) & 31) > 29 || ((peek >> 8) & 31) > 29) continue;  // HLIT, HDIST
            in.seek(b + 3);
            if (read_dynamic_header(in, true, t_->lit, sizeof(t_->lit) / 4, t_->dist, sizeof(t_->dist) / 4, problem) != 1) continue;
            size_t n = W;
            if (decode_block_symbols(in, t_->lit, t_->dist, r.sym, n, 0, trial_limit, problem) != 1) continue;
            if (n == W) continue;  // an empty block proves nothing
            if (!(peek & 1)) {
                // not the final block: the next three bits must be able to start a block
                if (in.bc < 17) in.refill();
                if (in.bc < 3) continue;
                const unsigned type = (unsigned)(in.bb >> 1) & 3;
                if (type == 3) continue;
                if (type == 2 && in.bc >= 13 && ((((in.bb >> 3) & 31) > 29) || (((in.bb >> 8) & 31) > 29))) continue;
            }
            break;
        }
        if (b >= end_bit || b + 64 >= total_bits) return;  // found stays false
        r.found = true;
        r.start_bit = b;
        problem = nullptr;
    }

    // ---- decode block after block from r.start_bit.  Symbols are 16 bits wide while a marker is
    // ---- within reach of a match (the last 32 KiB); once it is not, the rest of the span is plain
    // ---- bytes, decoded straight into the buffer the parser will read.
    in.seek(r.start_bit);
    bool at_member_start = fresh_member;
    bool wide = !fresh_member;  // the first block of a member cannot reach back at all
    size_t marker_end = W;      // sym index just past the last marker (the prefix is all markers)
    size_t nb = 0;              // bytes in r.bytes once narrow
    size_t min_pos_b = 0;
    r.n_wide = 0;
    r.marker_pos.clear();
    r.has_cr = false;
    r.ascii = true;
    if (r.bytes.size() < (1u << 20)) r.bytes.resize(1u << 20);
    auto go_narrow = [&]() {
        const size_t n1 = r.n_sym - W;
        if (r.bytes.size() < n1 + (1u << 20)) r.bytes.resize(n1 + n1 / 2 + (1u << 20));
        narrow_symbols(r.sym.data() + W, n1, (unsigned char *)r.bytes.data(), r.marker_pos, r.has_cr, r.ascii);
        r.n_wide = n1;
        nb = n1;
        min_pos_b = min_pos > W ? min_pos - W : 0;
        wide = false;
    };
    auto finish = [&]() {
        if (wide) go_narrow();
        // what the line reader wants to know about the bytes that never were symbols
        const unsigned char *q = (const unsigned char *)r.bytes.data() + r.n_wide;
        const size_t m = nb - r.n_wide;
        if (m) {
            if (memchr(q, '\r', m)) r.has_cr = true;
            uint64_t acc = 0;
            size_t i = 0;
            for (; i + 8 <= m; i += 8) acc |= load64(q + i);
            for (; i < m; i++) acc |= q[i];
            if (acc & 0x8080808080808080ull) r.ascii = false;
        }
        r.n_out = nb;
    };
    for (;;) {
        const uint64_t here = in.bit_pos();
        if (here >= end_bit && here > r.start_bit) {
            r.stop_bit = here;
            r.stop_fresh = at_member_start;
            finish();
            return;
        }
        at_member_start = false;
        if (in.bc < 3) in.refill();
        if (in.bc < 3) { r.error = "truncated deflate stream"; return; }
        const unsigned final_bit = (unsigned)in.bb & 1, type = (unsigned)(in.bb >> 1) & 3;
        in.bb >>= 3;
        in.bc -= 3;
        const size_t block_from = r.n_sym;
        if (type == 0) {
            const unsigned char *q = in.align_to_byte();
            if (in.end - q < 4) { r.error = "truncated stored block"; return; }
            const unsigned len = q[0] | (q[1] << 8), nlen = q[2] | (q[3] << 8);
            if ((len ^ nlen) != 0xFFFF) { r.error = "stored block length check failed"; return; }
            q += 4;
            if ((size_t)(in.end - q) < len) { r.error = "truncated stored block"; return; }
            if ((wide ? r.n_sym : nb) > kMaxSpanSymbols) { r.error = "span too large to hold"; return; }
            if (wide) {
                if (r.n_sym + len + 600 > r.sym.size()) r.sym.resize(r.n_sym + len + (1u << 20));
                for (unsigned i = 0; i < len; i++) r.sym[r.n_sym + i] = q[i];
                r.n_sym += len;
            } else {
                if (nb + len + 600 > r.bytes.size()) r.bytes.resize(nb + len + (1u << 20));
                memcpy(r.bytes.data() + nb, q, len);
                nb += len;
            }
            in.p = q + len;
        } else if (type == 3) {
            r.error = "invalid block type";
            return;
        } else {
            if (type == 1) {
                uint8_t ll[288], dl[32];
                fixed_code_lengths(ll, dl);
                build_table(ll, 288, payloads().lit, LB, t_->lit, sizeof(t_->lit) / 4, true);
                build_table(dl, 32, payloads().dist, DB, t_->dist, sizeof(t_->dist) / 4);
            } else {
                const int h = read_dynamic_header(in, false, t_->lit, sizeof(t_->lit) / 4, t_->dist, sizeof(t_->dist) / 4, problem);
                if (h != 1) {
                    r.error = h == -2 ? "truncated deflate stream" : (problem ? problem : "invalid block header");
                    return;
                }
            }
            const int rc = wide ? decode_block_symbols(in, t_->lit, t_->dist, r.sym, r.n_sym, min_pos, kMaxSpanSymbols, problem)
                                : decode_block_symbols(in, t_->lit, t_->dist, r.bytes, nb, min_pos_b, kMaxSpanSymbols, problem);
            if (rc != 1) {
                // -2: text that expands this much is for the serial reader and its bounded memory
                r.error = rc == -2 ? "span too large to hold" : (problem ? problem : "corrupt deflate stream");
                return;
            }
        }
        if (wide) {
            marker_end = last_marker_end(r.sym.data(), block_from, r.n_sym, marker_end);
            if (r.n_sym - marker_end >= W) go_narrow();
        }
        if (!final_bit) continue;
        // ---- end of a member: trailer, then either another member or the serial reader's business
        const unsigned char *q = in.align_to_byte();
        if (in.end - q < 8) { r.error = "truncated gzip trailer"; return; }
        SpanResult::MemberEnd me;
        me.pos = wide ? r.n_sym - W : nb;
        me.crc = q[0] | (q[1] << 8) | (q[2] << 16) | ((uint32_t)q[3] << 24);
        me.isize = q[4] | (q[5] << 8) | (q[6] << 16) | ((uint32_t)q[7] << 24);
        q += 8;
        me.next_off = (uint64_t)(q - file);
        r.ends.push_back(me);
        const size_t hdr = member_header_length(q, (size_t)(in.end - q));
        if (hdr == 0) {
            r.tail = true;
            r.tail_off = (uint64_t)(q - file);
            r.stop_bit = r.tail_off * 8;
            finish();
            return;
        }
        in.p = q + hdr;
        if (wide) min_pos = r.n_sym;  // the new member has no history
        else min_pos_b = nb;
        at_member_start = true;
    }
}

}  // namespace tb

// testing aid (include/trio_kmers_b200.h): decode `path` with GzStream and write the bytes to out_fd
extern "C" int64_t tb_gunzip_fd(const char *path, int out_fd, int64_t chunk)
{
    tb::GzStream gz;
    std::string err;
    if (!gz.open(path, err)) return -1;
    std::vector<char> buf((size_t)(chunk > 0 ? chunk : (1 << 20)));
    int64_t total = 0;
    for (;;) {
        const long n = gz.read(buf.data(), buf.size());
        if (n < 0) return -2;
        if (n == 0) break;
        size_t off = 0;
        while (off < (size_t)n) {
            const ssize_t w = ::write(out_fd, buf.data() + off, (size_t)n - off);
            if (w <= 0) return -3;
            off += (size_t)w;
        }
        total += n;
    }
    return total;
}

// testing aids for the CRC: the value, and whether the carry-less-multiply path is in use
extern "C" uint32_t tb_crc32(uint32_t crc, const void *buf, int64_t n)
{
    return tb::crc32_any(crc, (const unsigned char *)buf, n > 0 ? (size_t)n : 0);
}
extern "C" int tb_crc32_accelerated(void) { return tb::crc32_is_accelerated() ? 1 : 0; }
