// kmer_count.cu -- canonical k-mer counting, count histogram and set subtraction on sm_100a:
// the producer of the classifier's inputs.
//
// Replaces what the reference's find-unique-kmers delegates to the external KMC binaries
// (src/trio_binning/find_unique_kmers.py:62-103 `kmc`, :106-168 `kmc_tools transform ...
// histogram`, :171-197 `kmc_tools simple ... kmers_subtract`, :200-233 `kmc_dump -ci -cx`).
// KMC is not part of the reference tree; what is restated here is its documented behaviour
// with the options the reference passes:
//   * k-mers are counted in CANONICAL form: the lexicographically smaller of a k-mer and its
//     reverse complement (A < C < G < T, compared from the first base);
//   * a window that contains anything but A, C, G, T (either case) is not counted;
//   * counters saturate at 255 (-cs default) and k-mers seen once are dropped (-ci default 2)
//     -- both applied when the table is READ (histogram / select), the table keeps full counts;
//   * subtraction is on the k-mer SETS of the two databases (counts >= 2), the survivor keeps
//     its own count; the dump keeps counts within [min, max].
//
// Keys are MSB-first 2-bit encodings of the canonical string (first base in the top digit), so
// integer order is the lexicographic order KMC dumps in.  With fwd / rc the LSB-first values of a
// window and of its reverse complement (the classify kernels' encoding, c/kmers.c:50-72):
//   msb(window) = ~rc & mask,  msb(revcomp) = ~fwd & mask,  canonical = ~max(fwd, rc) & mask.
//
// Table: open addressing, linear probing, empty = ~0 (no canonical key is all ones, its partner
// would be 0).  Two layouts:
//   packed (k <= 22, the reference's default k = 21 included): ONE 64-bit word per slot, the key in
//     its low 2k bits and the count above it; insert = atomicCAS(empty -> key | 1 << 2k), count =
//     one 64-bit atomic add on the same word.  8 bytes per slot and one random sector per window.
//     The count stops near 2^19 (readers clamp at 255, far below) so that it never carries out.
//   split (k >= 23): keys[] (u64) and counts[] (u32), atomicCAS on the key, atomicAdd on the count;
//     12 bytes per slot, two random sectors per window.
// Bound by random 32-byte sector accesses to HBM (each one a 128-byte line fetch); the
// slot size matters beyond the kernel: it sets how many passes over the input a genome needs.
#include "launchers.h"

namespace {

__device__ __forceinline__ uint64_t count_slot_of(uint64_t key, uint64_t n_slots)
{
    return __umul64hi(tb_hash64(key), n_slots);
}

// key and count of an occupied slot in either layout
__device__ __forceinline__ uint64_t slot_key(const tb_count_view &T, unsigned long long word)
{
    return T.count_shift ? word & ((1ull << T.count_shift) - 1ull) : word;
}
__device__ __forceinline__ uint32_t slot_count(const tb_count_view &T, unsigned long long word, uint64_t slot)
{
    return T.count_shift ? (uint32_t)(word >> T.count_shift) : T.counts[slot];
}

__device__ __forceinline__ int64_t read_index_of(const int64_t *__restrict__ offsets, int64_t n_reads, int64_t x)
{
    int64_t lo = 0, hi = n_reads;  // offsets[lo] <= x < offsets[hi]
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (offsets[mid] <= x) lo = mid; else hi = mid;
    }
    return lo;
}

// One thread per 32-base word of the flat 2-bit stream: its 32 window starts.
//   F  2-bit codes (k_pack), V  one validity bit per base (k_pack with b.V set)
__global__ void __launch_bounds__(128)
k_count_windows(const uint64_t *__restrict__ F, const uint32_t *__restrict__ V, const int64_t *__restrict__ offsets,
                int64_t n_reads, int64_t total_bases, int k, tb_count_view T, unsigned long long *stats)
{
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t p0 = w * 32;
    unsigned n_counted = 0, n_new = 0, overflow = 0;
    if (p0 < total_bases) {
        const uint64_t f0 = F[w], f1 = F[w + 1];
        const uint64_t vspan = (uint64_t)V[w] | ((uint64_t)V[w + 1] << 32);
        const uint64_t kmask = k == 32 ? ~0ull : ((1ull << (2 * k)) - 1ull);
        const uint64_t vneed = k == 32 ? 0xFFFFFFFFull : ((1ull << k) - 1ull);
        int64_t r = read_index_of(offsets, n_reads, p0);
        int64_t end = offsets[r + 1];
        for (int j = 0; j < 32; j++) {
            const int64_t pos = p0 + j;
            if (pos >= total_bases) break;
            while (end <= pos) end = offsets[++r + 1];
            if (pos + k > end) continue;                       // the window would leave its read
            if (((vspan >> j) & vneed) != vneed) continue;      // a non-ACGT base inside
            const uint64_t fwd = (j == 0 ? f0 : (f0 >> (2 * j)) | (f1 << (64 - 2 * j))) & kmask;
            const uint64_t rc = tb_revcomp64(fwd) >> (64 - 2 * k);
            const uint64_t key = ~(fwd > rc ? fwd : rc) & kmask;
            if (T.n_parts > 1 && (uint32_t)((tb_hash64(key ^ 0x5bd1e9955bd1e995ull) >> 40) % T.n_parts) != T.part) continue;
            n_counted++;
            uint64_t slot = count_slot_of(key, T.n_slots);
            bool done = false;
            if (T.count_shift) {
                const unsigned long long one = 1ull << T.count_shift;
                for (uint64_t probe = 0; probe < T.n_slots && !done; probe++) {
                    unsigned long long cur = T.keys[slot];
                    if (cur == TB_EMPTY) {
                        cur = atomicCAS(&T.keys[slot], TB_EMPTY, (unsigned long long)key | one);
                        if (cur == TB_EMPTY) {
                            n_new++;
                            done = true;
                        }
                    }
                    if (!done && (cur & kmask) == key) {
                        if ((cur >> T.count_shift) < TB_COUNT_PACKED_SATURATE) atomicAdd(&T.keys[slot], one);
                        done = true;
                    }
                    slot = slot + 1 == T.n_slots ? 0 : slot + 1;
                }
            } else {
                for (uint64_t probe = 0; probe < T.n_slots && !done; probe++) {
                    unsigned long long cur = T.keys[slot];
                    if (cur == TB_EMPTY) {
                        cur = atomicCAS(&T.keys[slot], TB_EMPTY, (unsigned long long)key);
                        if (cur == TB_EMPTY) {
                            n_new++;
                            cur = key;
                        }
                    }
                    if (cur == key) {
                        atomicAdd(&T.counts[slot], 1u);
                        done = true;
                    }
                    slot = slot + 1 == T.n_slots ? 0 : slot + 1;
                }
            }
            if (!done) overflow = 1;
        }
    }
    // per-warp totals
    n_counted = __reduce_add_sync(0xFFFFFFFFu, n_counted);
    n_new = __reduce_add_sync(0xFFFFFFFFu, n_new);
    overflow = __reduce_or_sync(0xFFFFFFFFu, overflow);
    if ((threadIdx.x & 31) == 0) {
        if (n_counted) atomicAdd(&stats[0], (unsigned long long)n_counted);
        if (n_new) atomicAdd(&stats[1], (unsigned long long)n_new);
        if (overflow) atomicOr(&stats[2], 1ull);
    }
}

// hist[c] = number of stored k-mers whose saturated count is c (c = 0 .. n_bins-1)
__global__ void __launch_bounds__(256)
k_count_histogram(tb_count_view T, unsigned long long *hist, int n_bins)
{
    __shared__ unsigned int local[1024];
    for (int i = threadIdx.x; i < n_bins; i += blockDim.x) local[i] = 0;
    __syncthreads();
    for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < T.n_slots; s += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long word = T.keys[s];
        if (word == TB_EMPTY) continue;
        const uint32_t c = slot_count(T, word, s);
        atomicAdd(&local[c < (uint32_t)n_bins ? c : (uint32_t)n_bins - 1], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_bins; i += blockDim.x)
        if (local[i]) atomicAdd(&hist[i], (unsigned long long)local[i]);
}

__device__ __forceinline__ uint32_t count_lookup(const tb_count_view &T, uint64_t key)
{
    uint64_t slot = count_slot_of(key, T.n_slots);
    for (uint64_t probe = 0; probe < T.n_slots; probe++) {
        const unsigned long long cur = T.keys[slot];
        if (cur == TB_EMPTY) return 0;
        if (slot_key(T, cur) == key) return slot_count(T, cur, slot);
        slot = slot + 1 == T.n_slots ? 0 : slot + 1;
    }
    return 0;
}

// keys of A with db_min <= count, min_count <= min(count, counter_max) <= max_count that are not
// in B's database (count in B < db_min); *n_out counts them all, the first `cap` are stored
__global__ void __launch_bounds__(256)
k_count_select(tb_count_view A, tb_count_view B, uint32_t min_count, uint32_t max_count, uint32_t db_min,
               uint32_t counter_max, unsigned long long *out, unsigned long long cap, unsigned long long *n_out)
{
    for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < A.n_slots; s += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long word = A.keys[s];
        if (word == TB_EMPTY) continue;
        const unsigned long long key = slot_key(A, word);
        uint32_t c = slot_count(A, word, s);
        if (c < db_min) continue;
        if (c > counter_max) c = counter_max;
        if (c < min_count || c > max_count) continue;
        if (B.keys && count_lookup(B, key) >= db_min) continue;
        const unsigned long long i = atomicAdd(n_out, 1ull);
        if (i < cap) out[i] = key;
    }
}

}  // namespace

cudaError_t tb_launch_count_windows(const uint64_t *d_F, const uint32_t *d_V, const int64_t *d_offsets, int64_t n_reads,
                                    int64_t total_bases, int k, tb_count_view T, unsigned long long *d_stats,
                                    cudaStream_t stream)
{
    if (total_bases <= 0 || n_reads <= 0) return cudaSuccess;
    const int64_t words = (total_bases + 31) / 32;
    k_count_windows<<<(unsigned)((words + 127) / 128), 128, 0, stream>>>(d_F, d_V, d_offsets, n_reads, total_bases, k, T, d_stats);
    return cudaGetLastError();
}

cudaError_t tb_launch_count_histogram(tb_count_view T, unsigned long long *d_hist, int n_bins, cudaStream_t stream)
{
    if (n_bins < 1 || n_bins > 1024) return cudaErrorInvalidValue;
    k_count_histogram<<<148 * 8, 256, 0, stream>>>(T, d_hist, n_bins);
    return cudaGetLastError();
}

cudaError_t tb_launch_count_select(tb_count_view A, tb_count_view B, uint32_t min_count, uint32_t max_count,
                                   uint32_t db_min, uint32_t counter_max, unsigned long long *d_out,
                                   unsigned long long cap, unsigned long long *d_n_out, cudaStream_t stream)
{
    k_count_select<<<148 * 8, 256, 0, stream>>>(A, B, min_count, max_count, db_min, counter_max, d_out, cap, d_n_out);
    return cudaGetLastError();
}
