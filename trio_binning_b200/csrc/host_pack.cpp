// host_pack.cpp -- host side of the read stream: ASCII bases -> pinned 2-bit batches.
//
// north_star: "reads are parsed on the host ... into pinned, 2-bit-packed batches and
// streamed with cudaMemcpyAsync".  This is that packer: the same arithmetic as the
// device kernel k_pack (classify.cu), on the host, multi-threaded and AVX-512/AVX2-vectorised, so
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

#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
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

// The common case, a run of 256-base blocks with nothing but ACGT in them: 4 x 64 bytes in, one
// 64-byte line of F out (non-temporal when F is line-aligned: the packed stream is read next by
// the DMA engine, not by this core).  Validity and code both come from a 16-entry table indexed
// by the low nibble of the byte ('A' 0x41, 'C' 0x43, 'T' 0x54, 'G' 0x47): the validity table
// returns the full expected byte, so one compare against the input decides, and bytes >= 0x80
// shuffle to 0 and fail it.  Stops in front of the first block that holds anything else and
// returns the number of blocks done.  A core streams ~1.7x faster through this loop than through
// one word at a time; the prefetch distance (4 KB) was picked by measurement.
__attribute__((target("avx512f,avx512bw"))) int64_t pack_clean_blocks_avx512(const unsigned char *a,
                                                                            int64_t n_blocks, uint64_t *F)
{
    const __m512i tv =
        _mm512_broadcast_i32x4(_mm_setr_epi8(0, 'A', 0, 'C', 'T', 0, 0, 'G', 0, 0, 0, 0, 0, 0, 0, 0));
    const __m512i tc = _mm512_broadcast_i32x4(_mm_setr_epi8(0, 0, 0, 1, 3, 0, 0, 2, 0, 0, 0, 0, 0, 0, 0, 0));
    const __m512i m1 = _mm512_set1_epi16(0x0401), m2 = _mm512_set1_epi32(0x00100001);
    const bool aligned = ((uintptr_t)F & 63) == 0;
    int64_t b = 0;
    for (; b < n_blocks; b++) {
        const unsigned char *p = a + b * 256;
        _mm_prefetch((const char *)p + 4096, _MM_HINT_T0);
        _mm_prefetch((const char *)p + 4096 + 64, _MM_HINT_T0);
        _mm_prefetch((const char *)p + 4096 + 128, _MM_HINT_T0);
        _mm_prefetch((const char *)p + 4096 + 192, _MM_HINT_T0);
        const __m512i x0 = _mm512_loadu_si512(p), x1 = _mm512_loadu_si512(p + 64);
        const __m512i x2 = _mm512_loadu_si512(p + 128), x3 = _mm512_loadu_si512(p + 192);
        const __mmask64 ok = _mm512_cmpeq_epi8_mask(_mm512_shuffle_epi8(tv, x0), x0) &
                             _mm512_cmpeq_epi8_mask(_mm512_shuffle_epi8(tv, x1), x1) &
                             _mm512_cmpeq_epi8_mask(_mm512_shuffle_epi8(tv, x2), x2) &
                             _mm512_cmpeq_epi8_mask(_mm512_shuffle_epi8(tv, x3), x3);
        if (ok != ~0ull) break;
        // 4 codes -> a byte: (c0 + 4 c1) per 16-bit lane, then (lo + 16 hi) per 32-bit lane
        const __m512i c0 = _mm512_madd_epi16(_mm512_maddubs_epi16(_mm512_shuffle_epi8(tc, x0), m1), m2);
        const __m512i c1 = _mm512_madd_epi16(_mm512_maddubs_epi16(_mm512_shuffle_epi8(tc, x1), m1), m2);
        const __m512i c2 = _mm512_madd_epi16(_mm512_maddubs_epi16(_mm512_shuffle_epi8(tc, x2), m1), m2);
        const __m512i c3 = _mm512_madd_epi16(_mm512_maddubs_epi16(_mm512_shuffle_epi8(tc, x3), m1), m2);
        __m512i o = _mm512_castsi128_si512(_mm512_cvtepi32_epi8(c0));
        o = _mm512_inserti32x4(o, _mm512_cvtepi32_epi8(c1), 1);
        o = _mm512_inserti32x4(o, _mm512_cvtepi32_epi8(c2), 2);
        o = _mm512_inserti32x4(o, _mm512_cvtepi32_epi8(c3), 3);
        if (aligned)
            _mm512_stream_si512((__m512i *)(F + b * 8), o);
        else
            _mm512_storeu_si512((void *)(F + b * 8), o);
    }
    _mm_sfence();
    return b;
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
        if (avx512 && (w & 7) == 0) {
            const int64_t n_blocks = std::min((w1 - w) / 8, (total - p0) / 256);
            if (n_blocks > 0) {
                const int64_t done = pack_clean_blocks_avx512(ascii + p0, n_blocks, F + w);
                if (done > 0) {
                    w += done * 8 - 1;  // the loop's w++ steps onto the first word not done
                    continue;
                }
            }
        }
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

// Persistent workers for tb_pack_reads: a batch is packed in 2-3 ms, and starting 15 threads per
// call costs a tenth of that.  One job at a time (callers serialise on run_mutex); the caller
// runs share 0 itself.  Heap-allocated and never destroyed, so no worker ever waits on a
// destructed condition variable at exit; rebuilt in a forked child, which has no workers.
class PackPool {
  public:
    static PackPool &get()
    {
        static std::mutex guard;
        static PackPool *pool = nullptr;
        static pid_t owner = 0;
        std::lock_guard<std::mutex> lk(guard);
        if (!pool || owner != getpid()) {
            pool = new PackPool();
            owner = getpid();
        }
        return *pool;
    }

    void run(int n, const std::function<void(int)> &f)
    {
        std::lock_guard<std::mutex> one(run_mutex_);
        {
            std::unique_lock<std::mutex> lk(m_);
            while ((int)workers_.size() < n - 1) {
                workers_.emplace_back([this] { worker(); });
                workers_.back().detach();
            }
            job_ = &f;
            n_jobs_ = n;
            next_ = 1;
            pending_ = n - 1;
        }
        cv_work_.notify_all();
        f(0);
        std::unique_lock<std::mutex> lk(m_);
        cv_done_.wait(lk, [&] { return pending_ == 0; });
        job_ = nullptr;
        n_jobs_ = next_ = 0;
    }

  private:
    void worker()
    {
        std::unique_lock<std::mutex> lk(m_);
        for (;;) {
            cv_work_.wait(lk, [&] { return next_ < n_jobs_; });
            const int t = next_++;
            const std::function<void(int)> *f = job_;
            lk.unlock();
            (*f)(t);
            lk.lock();
            if (--pending_ == 0) cv_done_.notify_one();
        }
    }

    std::mutex run_mutex_, m_;
    std::condition_variable cv_work_, cv_done_;
    std::vector<std::thread> workers_;
    const std::function<void(int)> *job_ = nullptr;
    int n_jobs_ = 0, next_ = 0, pending_ = 0;
};

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
    const std::function<void(int)> work = [&](int t) {
        const int64_t t0 = tiles_total * t / threads, t1 = tiles_total * (t + 1) / threads;
        pack_range(ascii, total, offsets, n_reads, k, t0 * 32, t1 * 32, out->F, out->dirty,
                   dirty_lists[(size_t)t], ub_lists[(size_t)t]);
    };
    if (threads == 1)
        work(0);
    else
        PackPool::get().run(threads, work);

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
