// host_pack.cpp -- host side of the read stream: ASCII bases -> pinned 2-bit batches.
//
// north_star: "reads are parsed on the host ... into pinned, 2-bit-packed batches and
// streamed with cudaMemcpyAsync".  This is that packer: the same arithmetic as the
// device kernel k_pack (classify.cu), on the host, multi-threaded and AVX2-vectorised, so
// that 0.25 byte per base crosses PCIe instead of 1.  It produces exactly the layout the
// classify kernels consume (DESIGN.md "Data layout"):
//   F       flat 2-bit forward stream (non-ACGT -> 0; reference c/kmers.c:56-68)
//   dirty   1 bit per 32-base word: the word holds a non-ACGT byte
//   E words fill-forward stream for dirty words only (reference c/kmers.c:78-91, quirk Q3)
//   ub_*    reads with a non-ACGT byte among their first k bases (quirk Q3-UB): number of
//           leading windows to replay literally + the ASCII prefix the replay needs
// No k-mer lookups happen here; this is format conversion only.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "../../include/trio_kmers_b200.h"

namespace {

inline bool base_valid(unsigned char c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }
inline uint32_t base_code(unsigned char c)
{
    return base_valid(c) ? (((uint32_t)c >> 1) ^ ((uint32_t)c >> 2)) & 3u : 0u;
}

// 32 ASCII bytes -> 64-bit word of codes + 32-bit validity mask (scalar)
inline void pack32_scalar(const unsigned char *p, int n, uint64_t &f, uint32_t &valid)
{
    f = 0;
    valid = 0xFFFFFFFFu;  // positions past the end count as valid 'A'
    for (int i = 0; i < n; i++) {
        unsigned char c = p[i];
        if (base_valid(c))
            f |= (uint64_t)base_code(c) << (2 * i);
        else
            valid &= ~(1u << i);
    }
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) inline void pack32_avx2(const unsigned char *p, uint64_t &f,
                                                         uint32_t &valid)
{
    const __m256i x = _mm256_loadu_si256((const __m256i *)p);
    const __m256i va = _mm256_cmpeq_epi8(x, _mm256_set1_epi8('A'));
    const __m256i vc = _mm256_cmpeq_epi8(x, _mm256_set1_epi8('C'));
    const __m256i vg = _mm256_cmpeq_epi8(x, _mm256_set1_epi8('G'));
    const __m256i vt = _mm256_cmpeq_epi8(x, _mm256_set1_epi8('T'));
    const __m256i ok = _mm256_or_si256(_mm256_or_si256(va, vc), _mm256_or_si256(vg, vt));
    valid = (uint32_t)_mm256_movemask_epi8(ok);
    // code = ((c >> 1) ^ (c >> 2)) & 3, zeroed where invalid (16-bit shifts: the bits that
    // leak in from the neighbouring byte are above bit 1 and masked off)
    __m256i code = _mm256_xor_si256(_mm256_srli_epi16(x, 1), _mm256_srli_epi16(x, 2));
    code = _mm256_and_si256(_mm256_and_si256(code, _mm256_set1_epi8(3)), ok);
    // bytes (c0,c1) -> c0 + 4*c1 in 16-bit lanes; 16-bit (a,b) -> a + 16*b in 32-bit lanes
    const __m256i nib = _mm256_maddubs_epi16(code, _mm256_set1_epi16(0x0401));
    const __m256i byt = _mm256_madd_epi16(nib, _mm256_set1_epi32(0x00100001));
    // every 32-bit lane now holds 4 bases in its low byte: gather the 8 low bytes
    const __m256i shuf = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                          0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    const __m256i g = _mm256_shuffle_epi8(byt, shuf);
    const uint32_t lo = (uint32_t)_mm256_extract_epi32(g, 0);
    const uint32_t hi = (uint32_t)_mm256_extract_epi32(g, 4);
    f = ((uint64_t)hi << 32) | lo;
}
#endif

// 64 ASCII bytes -> two 64-bit words + 64-bit validity mask (AVX-512 F+BW)
__attribute__((target("avx512f,avx512bw"))) inline void pack64_avx512(const unsigned char *p, uint64_t &f0,
                                                                      uint64_t &f1, uint64_t &valid)
{
    const __m512i x = _mm512_loadu_si512((const void *)p);
    const __mmask64 ok = _mm512_cmpeq_epi8_mask(x, _mm512_set1_epi8('A')) |
                         _mm512_cmpeq_epi8_mask(x, _mm512_set1_epi8('C')) |
                         _mm512_cmpeq_epi8_mask(x, _mm512_set1_epi8('G')) |
                         _mm512_cmpeq_epi8_mask(x, _mm512_set1_epi8('T'));
    valid = (uint64_t)ok;
    __m512i code = _mm512_xor_si512(_mm512_srli_epi16(x, 1), _mm512_srli_epi16(x, 2));
    code = _mm512_maskz_and_epi64(0xFF, code, _mm512_set1_epi8(3));
    code = _mm512_maskz_mov_epi8(ok, code);
    const __m512i nib = _mm512_maddubs_epi16(code, _mm512_set1_epi16(0x0401));
    const __m512i byt = _mm512_madd_epi16(nib, _mm512_set1_epi32(0x00100001));
    const __m128i out = _mm512_cvtepi32_epi8(byt);  // 16 dwords -> 16 bytes = 64 bases
    f0 = (uint64_t)_mm_extract_epi64(out, 0);
    f1 = (uint64_t)_mm_extract_epi64(out, 1);
}

bool have_avx512()
{
#if defined(__x86_64__)
    static const bool v = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") &&
                          !getenv("TB_NO_AVX512");
    return v;
#else
    return false;
#endif
}

bool have_avx2()
{
#if defined(__x86_64__)
    static const bool v = __builtin_cpu_supports("avx2");
    return v;
#else
    return false;
#endif
}

// code of the nearest valid base strictly before flat position x (0 if none)
uint32_t lookback_code(const unsigned char *ascii, int64_t x)
{
    for (int64_t p = x - 1; p >= 0; p--)
        if (base_valid(ascii[p])) return base_code(ascii[p]);
    return 0;
}

int64_t read_of_position(const int64_t *offsets, int64_t n_reads, int64_t x)
{
    // last r with offsets[r] <= x, skipping empty reads automatically
    const int64_t *it = std::upper_bound(offsets, offsets + n_reads + 1, x);
    return (int64_t)(it - offsets) - 1;
}

struct DirtyWord {
    uint32_t index;
    uint64_t e;
};

// one thread's share: words [w0, w1)
void pack_range(const unsigned char *ascii, int64_t total, const int64_t *offsets, int64_t n_reads,
                int k, int64_t w0, int64_t w1, uint64_t *F, uint32_t *dirty,
                std::vector<DirtyWord> &dirty_out, std::vector<int64_t> &ub_reads)
{
    const bool avx2 = have_avx2();
#if defined(__x86_64__)
    const bool avx512 = have_avx512();
#endif
    for (int64_t w = w0; w < w1; w++) {
        const int64_t p0 = w * 32;
        uint64_t f;
        uint32_t valid;
#if defined(__x86_64__)
        if (avx512 && w + 1 < w1 && p0 + 64 <= total) {
            // common case: two clean words at once
            uint64_t fa, fb, v64;
            pack64_avx512(ascii + p0, fa, fb, v64);
            if (v64 == ~0ull) {
                F[w] = fa;
                F[w + 1] = fb;
                w++;
                continue;
            }
        }
#endif
        if (p0 + 32 <= total) {
#if defined(__x86_64__)
            if (avx2)
                pack32_avx2(ascii + p0, f, valid);
            else
#endif
                pack32_scalar(ascii + p0, 32, f, valid);
        } else if (p0 < total) {
            pack32_scalar(ascii + p0, (int)(total - p0), f, valid);
        } else {
            f = 0;
            valid = 0xFFFFFFFFu;
        }
        F[w] = f;
        if (valid != 0xFFFFFFFFu) {
            // rare path: fill-forward word + Q3-UB flagging (same as k_pack)
            uint32_t carry = lookback_code(ascii, p0);
            uint64_t e = 0;
            int64_t r = -1, r_start = 0, r_end = -1;
            for (int i = 0; i < 32; i++) {
                if ((valid >> i) & 1u) {
                    carry = (uint32_t)(f >> (2 * i)) & 3u;
                } else {
                    const int64_t x = p0 + i;
                    if (x >= r_end) {
                        r = read_of_position(offsets, n_reads, x);
                        r_start = offsets[r];
                        r_end = offsets[r + 1];
                    }
                    if (x - r_start < k && (ub_reads.empty() || ub_reads.back() != r)) ub_reads.push_back(r);
                }
                e |= (uint64_t)carry << (2 * i);
            }
            dirty_out.push_back({ (uint32_t)w, e });
            // the tile's bitmap word is shared by 32 consecutive words; ranges are tile-aligned
            dirty[w >> 5] |= 1u << (w & 31);
        }
    }
}

}  // namespace

extern "C" {

int tb_pack_reads(const char *bases, const int64_t *offsets, int64_t n_reads, int k, int n_threads,
                  tb_packed_batch *out)
{
    if (!out || !offsets || n_reads < 0 || k < 1 || k > 32) return -1;
    const unsigned char *ascii = (const unsigned char *)bases;
    const int64_t total = n_reads ? offsets[n_reads] : 0;
    const int64_t n_tiles = (total + 1023) / 1024;
    const int64_t n_words_padded = (n_tiles + 1) * 32;  // one spare tile, like k_pack writes
    if (n_words_padded > out->cap_words || n_tiles + 1 > out->cap_tiles) return -2;
    if (total >= ((int64_t)1 << 37)) return -2;  // dirty word indices are 32-bit
    out->n_reads = n_reads;
    out->total_bases = total;
    out->n_tiles = n_tiles;
    out->n_dirty = 0;
    out->n_ub = 0;
    out->ub_bytes_used = 0;
    memset(out->dirty, 0, (size_t)(n_tiles + 1) * sizeof(uint32_t));

    int threads = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    if (threads < 1) threads = 1;
    const int64_t tiles_total = n_tiles + 1;
    if ((int64_t)threads > (tiles_total + 255) / 256) threads = (int)((tiles_total + 255) / 256);
    if (threads < 1) threads = 1;
    std::vector<std::vector<DirtyWord>> dirty_lists((size_t)threads);
    std::vector<std::vector<int64_t>> ub_lists((size_t)threads);
    auto work = [&](int t) {
        const int64_t t0 = tiles_total * t / threads, t1 = tiles_total * (t + 1) / threads;
        pack_range(ascii, total, offsets, n_reads, k, t0 * 32, t1 * 32, out->F, out->dirty,
                   dirty_lists[(size_t)t], ub_lists[(size_t)t]);
    };
    if (threads == 1) {
        work(0);
    } else {
        std::vector<std::thread> pool;
        for (int t = 1; t < threads; t++) pool.emplace_back(work, t);
        work(0);
        for (auto &th : pool) th.join();
    }

    // merge the rare-path outputs (thread order == word order)
    for (auto &lst : dirty_lists) {
        for (const DirtyWord &d : lst) {
            if (out->n_dirty >= out->cap_dirty) return -3;
            out->dirty_idx[out->n_dirty] = d.index;
            out->e_words[out->n_dirty] = d.e;
            out->n_dirty++;
        }
    }
    int64_t last = -1;
    for (auto &lst : ub_lists) {
        for (int64_t r : lst) {
            if (r == last) continue;  // a read can be flagged from two adjacent ranges
            last = r;
            const int64_t start = offsets[r], len = offsets[r + 1] - start;
            const int64_t windows = len - k + 1;
            if (windows <= 0) continue;
            // U[j] = (first valid position >= j) - j; skip = min(max_j U[j], windows)
            int64_t max_u = 0, p = 0;
            for (int j = 0; j < k; j++) {
                if (p < j) p = j;
                while (p < len && !base_valid(ascii[start + p])) p++;
                max_u = std::max(max_u, p - j);
            }
            const int64_t skip = std::min(max_u, windows);
            if (skip == 0) continue;
            const int64_t need = skip + k - 1;  // bytes the literal replay reads
            if (out->n_ub >= out->cap_ub || out->ub_bytes_used + need > out->cap_ub_bytes) return -3;
            out->ub_read[out->n_ub] = (uint32_t)r;
            out->ub_skip[out->n_ub] = (uint32_t)skip;
            out->ub_off[out->n_ub] = out->ub_bytes_used;
            memcpy(out->ub_bytes + out->ub_bytes_used, ascii + start, (size_t)need);
            out->ub_bytes_used += need;
            out->n_ub++;
        }
    }
    return 0;
}

}  // extern "C"
