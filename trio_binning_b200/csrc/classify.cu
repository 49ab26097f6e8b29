// classify.cu -- the classify-by-kmers hot path on sm_100a.
//
// Replaces, for a whole batch of reads per launch:
//   count_kmers_in_read   c/kmers.c:270-299   (window loop, A then else-B)
//   kmer_in_hash_set      c/kmers.c:245-268   (encode fwd + revcomp, min, probe)
//   kmer_to_int           c/kmers.c:50-72     (2-bit LSB-first encode)
//   reverse_complement    c/kmers.c:74-93     (incl. the "no write for non-ACGT" quirk Q3)
//   score + decision      classify_by_kmers.py:104-115 (fp64, on device)
//
// Formulation (validated on the CPU in oracle/oracle.py::window_keys): with
//   F[x] = code(read[x]), non-ACGT -> 0
//   E[x] = code of the nearest ACGT base at or before x            (fill-forward)
// window i of a read has   fwd = sum_j F[i+j] << 2j,
//                          rc  = sum_j (3-E[i+j]) << 2(k-1-j),   key = min(fwd, rc)
// for every window except the first skip[r] windows of a read whose first k
// bases contain a non-ACGT byte (the reference reads never-written scratch
// there, Q3-UB); those few windows are done literally by k_fixup.
// Windows are therefore independent: the batch is ONE flat 2-bit stream cut
// into tiles of 1024 window-start positions, one warp per tile, one lane per
// 32 consecutive positions (= one aligned 64-bit word of the stream), whatever
// the read lengths are.
#include "launchers.h"

namespace {

// ============================================================================
// k_pack: flat ASCII -> F words, dirty bitmap, sparse E words, Q3-UB flags.
// One thread per 32-base word; a warp covers one tile, so the dirty bitmap word
// of the tile is a single ballot.
// ============================================================================

__device__ __forceinline__ uint32_t valid_mask4(uint32_t x)
{
    // 0xFF in every byte of x that is one of 'A','C','G','T'
    return __vcmpeq4(x, 0x41414141u) | __vcmpeq4(x, 0x43434343u) | __vcmpeq4(x, 0x47474747u) |
           __vcmpeq4(x, 0x54545454u);
}

// 4 ASCII bytes -> 8 bits of codes (byte i -> bits 2i) and 4 validity bits
__device__ __forceinline__ void pack4(uint32_t x, uint32_t &codes, uint32_t &valid)
{
    uint32_t vm = valid_mask4(x);
    uint32_t t = ((x >> 1) ^ (x >> 2)) & 0x03030303u & vm;
    codes = (t | (t >> 6) | (t >> 12) | (t >> 18)) & 0xFFu;
    uint32_t v = vm & 0x01010101u;
    valid = (v | (v >> 7) | (v >> 14) | (v >> 21)) & 0xFu;
}

// read that contains flat position x (offsets non-decreasing, offsets[n]=total > x)
__device__ __forceinline__ int64_t read_of_position(const int64_t *__restrict__ offsets,
                                                     int64_t n_reads, int64_t x)
{
    int64_t lo = 0, hi = n_reads;  // invariant: offsets[lo] <= x < offsets[hi]
    while (hi - lo > 1) {
        int64_t mid = (lo + hi) >> 1;
        if (offsets[mid] <= x) lo = mid; else hi = mid;
    }
    return lo;
}

// Code of the nearest valid base strictly before flat position x0 (0 if none), found by a whole
// warp: x0 is a multiple of 1024, the start of the warp's tile.  Every lane inspects its own 16 bytes of the 512 before
// the current position, so a run of L non-ACGT bases costs L/512 steps per TILE that lies in it
// instead of L/16 steps per WORD (gap-filled, assembly-like inputs).
__device__ __noinline__ uint32_t warp_lookback_code(const unsigned char *__restrict__ ascii, int64_t x0, int lane)
{
    for (int64_t base = x0; base > 0; base -= 512) {
        const int64_t p = base - 16 * (int64_t)(lane + 1);  // lane 0 holds the nearest 16 bytes
        uint4 q = make_uint4(0, 0, 0, 0);
        if (p >= 0) q = *reinterpret_cast<const uint4 *>(ascii + p);
        const bool any = (valid_mask4(q.x) | valid_mask4(q.y) | valid_mask4(q.z) | valid_mask4(q.w)) != 0;
        const unsigned am = __ballot_sync(0xFFFFFFFFu, any);
        if (am) {
            const int nearest = __ffs(am) - 1;
            uint32_t code = 0;
            if (lane == nearest) {
                const uint32_t words[4] = { q.x, q.y, q.z, q.w };
                for (int i = 15; i >= 0; i--) {
                    const unsigned char c = (unsigned char)(words[i >> 2] >> (8 * (i & 3)));
                    if (tb_base_valid(c)) {
                        code = tb_base_code(c);
                        break;
                    }
                }
            }
            return __shfl_sync(0xFFFFFFFFu, code, nearest);
        }
    }
    return 0;
}

__global__ void __launch_bounds__(256)
k_pack(tb_batch_view b, int k)
{
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // word index
    const int64_t p0 = w * 32;
    uint64_t f = 0;
    uint32_t vmask = 0xFFFFFFFFu;  // positions past the end count as valid 'A'
    if (p0 + 32 <= b.total_bases) {
        const uint4 *src = reinterpret_cast<const uint4 *>(b.ascii + p0);
        uint4 q0 = __ldg(src), q1 = __ldg(src + 1);
        uint32_t x[8] = { q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w };
        if (b.fold_case) {
            // k-mer counting: clearing bit 5 turns a, c, g, t into A, C, G, T and no other byte into a base
#pragma unroll
            for (int i = 0; i < 8; i++) x[i] &= 0xDFDFDFDFu;
        }
        // Fast path: take the two code bits of every byte as if it were a base, rebuild the
        // letter those bits stand for (A 0x41, C 0x43, G 0x47, T 0x54 = 0x41 + 2c + 2*[c>=2]
        // + 11*[c==3], four bytes at a time, no carries) and compare with the input: any
        // difference in the 32 bytes sends the word to the exact per-byte path below.
        uint32_t any_diff = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t u = (x[i] >> 1) ^ (x[i] >> 2);
            const uint32_t t = u & 0x03030303u;
            const uint32_t b1 = (u >> 1) & 0x01010101u;
            const uint32_t expect = 0x41414141u + 2u * t + 2u * b1 + 11u * (b1 & u);
            any_diff |= expect ^ x[i];
            f |= (uint64_t)((t | (t >> 6) | (t >> 12) | (t >> 18)) & 0xFFu) << (8 * i);
        }
        if (any_diff) {
            f = 0;
            vmask = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                uint32_t c, v;
                pack4(x[i], c, v);
                f |= (uint64_t)c << (8 * i);
                vmask |= v << (4 * i);
            }
        }
    } else if (p0 < b.total_bases) {
        int n = (int)(b.total_bases - p0);
        for (int i = 0; i < n; i++) {
            unsigned char c = b.ascii[p0 + i];
            if (b.fold_case) c &= 0xDFu;
            if (tb_base_valid(c))
                f |= (uint64_t)tb_base_code(c) << (2 * i);
            else
                vmask &= ~(1u << i);
        }
    }
    b.F[w] = f;
    if (b.V) {
        // k-mer counting wants the validity bits themselves; the fill-forward stream and the
        // Q3-UB bookkeeping below belong to the classify path only
        b.V[w] = vmask;
        return;
    }
    const bool is_dirty = vmask != 0xFFFFFFFFu;
    const unsigned dmask = __ballot_sync(0xFFFFFFFFu, is_dirty);
    uint32_t carry = 0;
    if (dmask) {
        // rare path, warp-uniform part: what a dirty word inherits from before its first base.  A
        // warp covers one tile (32 words), so the nearest valid base is either in a lower lane's
        // word (one shuffle) or before the tile (one look-back for the whole warp).
        const int lane = threadIdx.x & 31;
        const bool has = vmask != 0;
        const uint32_t last_code = has ? (uint32_t)(f >> (2 * (31 - __clz(vmask)))) & 3u : 0u;
        const unsigned hmask = __ballot_sync(0xFFFFFFFFu, has);
        const unsigned lower = hmask & ((1u << lane) - 1u);
        const int src = lower ? 31 - __clz(lower) : -1;
        const uint32_t from_lane = __shfl_sync(0xFFFFFFFFu, last_code, src < 0 ? 0 : src);
        uint32_t tile_carry = 0;
        if (__ballot_sync(0xFFFFFFFFu, is_dirty && src < 0)) tile_carry = warp_lookback_code(b.ascii, (w - lane) * 32, lane);
        carry = src >= 0 ? from_lane : tile_carry;
    }
    if (is_dirty) {
        // rare path: fill-forward stream + Q3-UB flagging
        uint64_t e = 0;
        int64_t r = -1, r_end = -1, r_start = 0;
        for (int i = 0; i < 32; i++) {
            if ((vmask >> i) & 1u) {
                carry = (uint32_t)(f >> (2 * i)) & 3u;
            } else {
                int64_t x = p0 + i;
                if (x >= r_end) {
                    r = read_of_position(b.offsets, b.n_reads, x);
                    r_start = b.offsets[r];
                    r_end = b.offsets[r + 1];
                }
                if (x - r_start < k) atomicOr(&b.flags[r], 2u);
            }
            e |= (uint64_t)carry << (2 * i);
        }
        b.E[w] = e;
    }
    if ((threadIdx.x & 31) == 0) b.dirty[w >> 5] = dmask;
}

// ============================================================================
// k_index: one thread per read.  tile -> first read map, zero the counts, and
// for Q3-UB-flagged reads compute how many leading windows the fix-up kernel
// must redo:  U[j] = (first ACGT position >= j) - j  for slot j < k,
// skip = min(max_j U[j], windows).
// ============================================================================
__global__ void __launch_bounds__(256)
k_index(tb_batch_view b, int k)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r == 0) *b.n_pass = 0;
    if (r >= b.n_reads) return;
    const int64_t start = b.offsets[r], end = b.offsets[r + 1];
    for (int64_t t = (start + 1023) >> 10; (t << 10) < end; t++)
        b.tile_first_read[t] = (uint32_t)r;
    b.count_a[r] = 0;
    b.count_b[r] = 0;
    uint32_t skip = 0;
    if (b.flags[r] & 2u) {
        b.flags[r] = 0;
        const int64_t len = end - start;
        const int64_t windows = len - k + 1;
        if (windows > 0) {
            int64_t max_u = 0;
            int64_t p = 0;  // first valid position >= j is non-decreasing in j
            for (int j = 0; j < k; j++) {
                if (p < j) p = j;
                while (p < len && !tb_base_valid(b.ascii[start + p])) p++;
                if (p - j > max_u) max_u = p - j;
            }
            skip = (uint32_t)(max_u < windows ? max_u : windows);
            if (skip) b.ub_list[atomicAdd(b.ub_count, 1u)] = (uint32_t)r;
        }
    }
    b.skip[r] = skip;
}

// ============================================================================
// host-packed batches: the work k_pack/k_index did from ASCII arrives precomputed
// (host_pack.cpp); what is left is the tile map, zeroed counters, scattering the few
// fill-forward words and installing the Q3-UB list.
// ============================================================================
__global__ void __launch_bounds__(256)
k_index_packed(tb_batch_view b)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r == 0) *b.n_pass = 0;
    if (r >= b.n_reads) return;
    const int64_t start = b.offsets[r], end = b.offsets[r + 1];
    for (int64_t t = (start + 1023) >> 10; (t << 10) < end; t++)
        b.tile_first_read[t] = (uint32_t)r;
    b.count_a[r] = 0;
    b.count_b[r] = 0;
    b.skip[r] = 0;
}

__global__ void __launch_bounds__(256)
k_apply_rare(tb_batch_view b, const uint32_t *__restrict__ dirty_idx,
             const uint64_t *__restrict__ e_words, int64_t n_dirty,
             const uint32_t *__restrict__ ub_read, const uint32_t *__restrict__ ub_skip, int64_t n_ub)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_dirty) b.E[dirty_idx[i]] = e_words[i];
    if (i < n_ub) {
        b.skip[ub_read[i]] = ub_skip[i];
        b.ub_list[i] = ub_read[i];
    }
    if (i == 0) *b.ub_count = (unsigned)n_ub;
}

// ============================================================================
// k_classify (direct) and k_classify_filtered (L2-resident pre-filter)
// ============================================================================

// bits [shift, shift+64) of the 128-bit value hi:lo  (0 <= shift < 128)
__device__ __forceinline__ uint64_t extract128(uint64_t lo, uint64_t hi, int shift)
{
    if (shift >= 64) return hi >> (shift - 64);
    if (shift == 0) return lo;
    return (lo >> shift) | (hi << (64 - shift));
}

// What one lane knows about its 32 window-start positions.
struct LaneCtx {
    uint64_t f0, f1;      // forward stream, 64 bases starting at the lane's first position
    uint64_t r_lo, r_hi;  // reverse complement of the same span of the E stream
    int64_t p0;           // flat position of window 0
    int64_t r, end;       // read containing p0 and its end offset
    uint32_t vmask;       // bit j: window j lies inside one read and is not left to k_fixup
    bool fast;            // all 32 windows valid and in read r
    bool clean;           // neither word of the span holds a non-ACGT base (E == F)
};

template <int K>
__device__ __forceinline__ uint64_t window_key(const LaneCtx &c, int j, int k, uint64_t kmask)
{
    const uint64_t fwd = extract128(c.f0, c.f1, 2 * j) & kmask;
    // window j = bases [j, j+k) of the span = bases [64-j-k, 64-j) of its reverse complement
    const uint64_t rc = extract128(c.r_lo, c.r_hi, 2 * (64 - j - k)) & kmask;
    return fwd < rc ? fwd : rc;
}

// lane_setup in two parts.  lane_stream: the lane's 64 bases of both streams -- all phase 1 of the
// minimizer kernel needs.  lane_bounds: which of its 32 windows lie inside one read (and past the
// read's Q3-UB replay), a chain of dependent loads (tile -> read -> offsets -> skip) whose latency
// that kernel hides behind phase 1 by asking for it afterwards.
__device__ __forceinline__ void lane_stream(const tb_batch_view &b, int64_t tile, int lane, LaneCtx &c)
{
    const int64_t w = tile * 32 + lane;
    c.p0 = w * 32;
    c.f0 = __ldg(b.F + w);
    c.f1 = __ldg(b.F + w + 1);
    const uint32_t d = __ldg(b.dirty + tile), dn = __ldg(b.dirty + tile + 1);
    const bool d0 = (d >> lane) & 1u;
    const bool d1 = lane < 31 ? (d >> (lane + 1)) & 1u : dn & 1u;
    const uint64_t e0 = d0 ? b.E[w] : c.f0;
    const uint64_t e1 = d1 ? b.E[w + 1] : c.f1;
    c.clean = !(d0 | d1);
    // reverse complement of the span: base s of the span sits at base 63-s
    c.r_lo = tb_revcomp64(e1);
    c.r_hi = tb_revcomp64(e0);
    c.r = 0;  // until lane_bounds says otherwise: no window of this lane exists
    c.end = 0;
    c.vmask = 0;
    c.fast = false;
}

__device__ __forceinline__ void lane_bounds(const tb_batch_view &b, int64_t tile, int k, LaneCtx &c)
{
    const int64_t *__restrict__ off = b.offsets;
    c.r = 0;
    c.end = 0;
    c.vmask = 0;
    c.fast = false;
    if (c.p0 < b.total_bases) {
        int64_t r = b.tile_first_read[tile];
        while (off[r + 1] <= c.p0) r++;
        const int64_t start = off[r], end = off[r + 1];
        c.r = r;
        c.end = end;
        c.fast = (c.p0 + 31 + k <= end) && ((uint64_t)(c.p0 - start) >= (uint64_t)b.skip[r]);
        if (c.fast) {
            c.vmask = 0xFFFFFFFFu;
        } else {
            // One step per READ that overlaps the lane's 32 positions (two for 150-bp reads), not
            // one per position: read [s, e) contributes the window starts from s + skip to e - k.
            int64_t rr = r, s = start, e = end;
            const int64_t last = c.p0 + 31;
            for (;;) {
                const int64_t lo = max(s + (int64_t)b.skip[rr], c.p0), hi = min(e - k, last);
                if (hi >= lo) {
                    const uint32_t n = (uint32_t)(hi - lo + 1);
                    c.vmask |= (n >= 32u ? 0xFFFFFFFFu : ((1u << n) - 1u)) << (uint32_t)(lo - c.p0);
                }
                if (e > last || e >= b.total_bases) break;
                rr++;
                s = e;
                e = off[rr + 1];
            }
        }
    }
}

__device__ __forceinline__ void lane_setup(const tb_batch_view &b, int64_t tile, int lane, int k, LaneCtx &c)
{
    lane_stream(b, tile, lane, c);
    lane_bounds(b, tile, k, c);
}

// Add the lane's hits to the per-read counters.
__device__ __forceinline__ void lane_attribute(const tb_batch_view &b, const LaneCtx &c, int lane,
                                               uint32_t hit_a, uint32_t hit_b)
{
    // common case: the whole warp sits inside one read -> one atomic per warp
    const int64_t r0 = __shfl_sync(0xFFFFFFFFu, c.r, 0);
    const bool uniform = __all_sync(0xFFFFFFFFu, c.fast && c.r == r0);
    if (uniform) {
        const int ca = __reduce_add_sync(0xFFFFFFFFu, __popc(hit_a));
        const int cb = __reduce_add_sync(0xFFFFFFFFu, __popc(hit_b));
        if (lane == 0) {
            if (ca) atomicAdd(&b.count_a[r0], ca);
            if (cb) atomicAdd(&b.count_b[r0], cb);
        }
    } else if (hit_a | hit_b) {
        if (c.fast) {
            if (hit_a) atomicAdd(&b.count_a[c.r], __popc(hit_a));
            if (hit_b) atomicAdd(&b.count_b[c.r], __popc(hit_b));
        } else {
            const int64_t *__restrict__ off = b.offsets;
            int64_t rr = c.r, e = c.end;
            uint32_t m = hit_a | hit_b;
            while (m) {
                const int j = __ffs(m) - 1;
                m &= m - 1;
                const int64_t pos = c.p0 + j;
                while (e <= pos) {
                    rr++;
                    e = off[rr + 1];
                }
                if ((hit_a >> j) & 1u)
                    atomicAdd(&b.count_a[rr], 1);
                else
                    atomicAdd(&b.count_b[rr], 1);
            }
        }
    }
}

// membership from an already loaded home bucket
template <int K>
__device__ __forceinline__ bool resolve(const tb_set_view &S, const tb_bucket &bk, uint64_t key,
                                        uint32_t h)
{
    bool hit = tb_bucket_has(bk, key);
    if (!hit && tb_bucket_full(bk)) hit = tb_probe_overflow(S, key, __umulhi(h, S.n_buckets));
    if (K == 0 || K == 32) {
        if (key == TB_EMPTY) hit = S.has_all_ones != 0;
    }
    return hit;
}

// ---- direct kernel: every window probes A, then B where A missed (Q4) -----------
template <int K, int U>
__global__ void __launch_bounds__(128)
k_classify(tb_batch_view b, int k_runtime, tb_set_view A, tb_set_view B)
{
    const int k = K ? K : k_runtime;
    const uint64_t kmask = k == 32 ? ~0ull : ((1ull << (2 * k)) - 1ull);
    const int lane = threadIdx.x & 31;
    const int64_t tile = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (tile >= b.n_tiles) return;  // warp-uniform
    LaneCtx c;
    lane_setup(b, tile, lane, k, c);

    uint32_t hit_a = 0, hit_b = 0;
#pragma unroll
    for (int g = 0; g < 32; g += U) {
        const uint32_t gm = (c.vmask >> g) & ((1u << U) - 1u);
        if (gm == 0) continue;
        uint64_t key[U];
        uint32_t h[U];
        tb_bucket bk[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            key[u] = window_key<K>(c, g + u, k, kmask);
            h[u] = tb_hash32(key[u]);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if ((gm >> u) & 1u)
                bk[u] = tb_load_bucket(A.table + __umulhi(h[u], A.n_buckets));
            else
                bk[u].s[0] = bk[u].s[1] = bk[u].s[2] = bk[u].s[3] = TB_EMPTY;
        }
        uint32_t ga = 0;
#pragma unroll
        for (int u = 0; u < U; u++) ga |= (resolve<K>(A, bk[u], key[u], h[u]) ? 1u : 0u) << u;
        ga &= gm;
        const uint32_t gmb = gm & ~ga;
#pragma unroll
        for (int u = 0; u < U; u++) {
            if ((gmb >> u) & 1u)
                bk[u] = tb_load_bucket(B.table + __umulhi(h[u], B.n_buckets));
            else
                bk[u].s[0] = bk[u].s[1] = bk[u].s[2] = bk[u].s[3] = TB_EMPTY;
        }
        uint32_t gb = 0;
#pragma unroll
        for (int u = 0; u < U; u++) gb |= (resolve<K>(B, bk[u], key[u], h[u]) ? 1u : 0u) << u;
        gb &= gmb;
        hit_a |= ga << g;
        hit_b |= gb << g;
    }
    lane_attribute(b, c, lane, hit_a, hit_b);
}

// ---- filtered kernel ---------------------------------------------------------------
// Phase 1: every window tests ONE bit of a bitmap over (A u B) that is sized to stay
// resident in L2 (<= 64 MB): a clear bit proves the k-mer is in neither set -- the
// common case by far -- without touching HBM.  Phase 2: only the windows whose bit is
// set (true members + false positives) probe the exact tables in HBM, A then B.
// Exactness is untouched: the bitmap can only say "maybe".
// The bitmap must STAY in L2 while table sectors stream through it.  ncu shows that
// L1::no_allocate loads are looked up as evict_first in L2 (lts__t_sectors_*_evict_first),
// which made bitmap and table lines compete as equals; the bitmap is therefore read with
// an explicit evict_last policy (mode 2), plain (evict_normal, mode 1) or as before (0).
__device__ __forceinline__ uint32_t ld_filter(const uint32_t *p, int mode, uint64_t policy)
{
    uint32_t v;
    if (mode == 2)
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(policy));
    else if (mode == 1)
        asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else
        asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

template <int K, int U>
__global__ void __launch_bounds__(128)
k_classify_filtered(tb_batch_view b, int k_runtime, tb_set_view A, tb_set_view B,
                    const uint32_t *__restrict__ filter, uint32_t filter_mask, int filter_mode)
{
    uint64_t keep_policy = 0;
    if (filter_mode == 2)
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep_policy));
    const int k = K ? K : k_runtime;
    const uint64_t kmask = k == 32 ? ~0ull : ((1ull << (2 * k)) - 1ull);
    const int lane = threadIdx.x & 31;
    const int64_t tile = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (tile >= b.n_tiles) return;  // warp-uniform
    LaneCtx c;
    lane_setup(b, tile, lane, k, c);

    // ---- phase 1: one filter bit per window, U loads in flight --------------------
    uint32_t pass = 0;
    if (c.vmask) {
#pragma unroll
        for (int g = 0; g < 32; g += U) {
            uint32_t h[U], wd[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const uint64_t key = window_key<K>(c, g + u, k, kmask);
                h[u] = tb_hash32(key);
                if (K == 0 || K == 32) {
                    // the out-of-band key (2^64-1) is not in the bitmap: force a pass
                    if (key == TB_EMPTY) pass |= 1u << (g + u);
                }
            }
#pragma unroll
            for (int u = 0; u < U; u++) wd[u] = ld_filter(filter + ((h[u] & filter_mask) >> 5), filter_mode, keep_policy);
#pragma unroll
            for (int u = 0; u < U; u++) pass |= ((wd[u] >> (h[u] & 31u)) & 1u) << (g + u);
        }
        pass &= c.vmask;
    }
    {
        const unsigned np = __reduce_add_sync(0xFFFFFFFFu, (unsigned)__popc(pass));
        if (lane == 0 && np) atomicAdd(b.n_pass, (unsigned long long)np);
    }

    // ---- phase 2: exact probes for the survivors, two windows per round -----------
    uint32_t hit_a = 0, hit_b = 0;
    uint32_t m = pass;
    while (m) {
        const int j0 = __ffs(m) - 1;
        m &= m - 1;
        const int j1 = m ? __ffs(m) - 1 : -1;
        if (m) m &= m - 1;
        const uint64_t key0 = window_key<0>(c, j0, k, kmask);
        const uint64_t key1 = window_key<0>(c, j1 < 0 ? j0 : j1, k, kmask);
        const uint32_t h0 = tb_hash32(key0), h1 = tb_hash32(key1);
        // all four buckets in flight at once; a hit in A simply discards the B answer
        const tb_bucket a0 = tb_load_bucket(A.table + __umulhi(h0, A.n_buckets));
        const tb_bucket a1 = tb_load_bucket(A.table + __umulhi(h1, A.n_buckets));
        const tb_bucket b0 = tb_load_bucket(B.table + __umulhi(h0, B.n_buckets));
        const tb_bucket b1 = tb_load_bucket(B.table + __umulhi(h1, B.n_buckets));
        const bool ia0 = resolve<K>(A, a0, key0, h0);
        const bool ib0 = !ia0 && resolve<K>(B, b0, key0, h0);
        hit_a |= (ia0 ? 1u : 0u) << j0;
        hit_b |= (ib0 ? 1u : 0u) << j0;
        if (j1 >= 0) {
            const bool ia1 = resolve<K>(A, a1, key1, h1);
            const bool ib1 = !ia1 && resolve<K>(B, b1, key1, h1);
            hit_a |= (ia1 ? 1u : 0u) << j1;
            hit_b |= (ib1 ? 1u : 0u) << j1;
        }
    }
    lane_attribute(b, c, lane, hit_a, hit_b);
}

// ---- pair kernel -----------------------------------------------------------------------
// Same two phases, with (1) a bitmap that sets NBITS bits per key inside ONE 32-bit word
// (still a single L2 request per window, several times fewer false positives while the
// bitmap is sparse) and (2) the pair table in phase 2: ONE HBM bucket per surviving window
// answers "in A? else in B?" (quirk Q4) instead of one bucket per set.  k <= 31 only (the
// pair table keeps its tags in the two bits a 31-mer leaves free).
template <int NBITS>
__device__ __forceinline__ uint32_t filter_bits(uint64_t h64)
{
    const uint32_t lo = (uint32_t)h64;
    // 1 << (x & 31) as one wrapping funnel shift (no separate mask)
    uint32_t m = __funnelshift_l(0u, 1u, lo);
    if (NBITS == 2) m |= __funnelshift_l(0u, 1u, lo >> 5);
    return m;
}

template <int K, int U, int NBITS>
__global__ void __launch_bounds__(128)
k_classify_pair(tb_batch_view b, int k_runtime, tb_pair_view P, const uint32_t *__restrict__ filter,
                uint32_t filter_word_mask, int filter_mode)
{
    const int k = K ? K : k_runtime;
    const uint64_t kmask = (1ull << (2 * k)) - 1ull;  // k <= 31
    const int lane = threadIdx.x & 31;
    const int64_t tile = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (tile >= b.n_tiles) return;  // warp-uniform
    LaneCtx c;
    lane_setup(b, tile, lane, k, c);
    uint64_t keep_policy = 0;
    if (filter_mode == 2)
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep_policy));

    // ---- phase 1 ---------------------------------------------------------------------
    uint32_t pass = 0;
    if (c.vmask) {
#pragma unroll
        for (int g = 0; g < 32; g += U) {
            uint32_t need[U], wd[U];
            uint32_t widx[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const uint64_t h64 = tb_hash64(window_key<K>(c, g + u, k, kmask));
                need[u] = filter_bits<NBITS>(h64);
                widx[u] = (uint32_t)(h64 >> 32) & filter_word_mask;
            }
#pragma unroll
            for (int u = 0; u < U; u++) wd[u] = ld_filter(filter + widx[u], filter_mode, keep_policy);
#pragma unroll
            for (int u = 0; u < U; u++) pass |= ((wd[u] & need[u]) == need[u] ? 1u : 0u) << (g + u);
        }
        pass &= c.vmask;
    }
    {
        const unsigned np = __reduce_add_sync(0xFFFFFFFFu, (unsigned)__popc(pass));
        if (lane == 0 && np) atomicAdd(b.n_pass, (unsigned long long)np);
    }

    // ---- phase 2: up to four surviving windows per round, one bucket each -------------
    uint32_t hit_a = 0, hit_b = 0;
    uint32_t m = pass;
    while (m) {
        int j[4];
        uint64_t key[4];
        uint32_t bkt[4];
        tb_bucket bk[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            j[q] = m ? __ffs(m) - 1 : -1;
            if (m) m &= m - 1;
            key[q] = window_key<0>(c, j[q] < 0 ? 0 : j[q], k, kmask);
            bkt[q] = tb_pair_bucket_of(tb_hash64(key[q]), P.n_buckets);
        }
#pragma unroll
        for (int q = 0; q < 4; q++)
            if (j[q] >= 0) bk[q] = tb_load_bucket(P.table + bkt[q]);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (j[q] < 0) continue;
            uint32_t tags = tb_pair_tags(bk[q], key[q]);
            if (!tags && tb_pair_bucket_full(bk[q])) tags = tb_pair_overflow(P, key[q], bkt[q]);
            hit_a |= (tags & 1u) << j[q];
            hit_b |= ((tags == 2u) ? 1u : 0u) << j[q];
        }
    }
    lane_attribute(b, c, lane, hit_a, hit_b);
}

template <int NBITS>
__global__ void __launch_bounds__(256)
k_pair_filter_build(const unsigned long long *__restrict__ slots, uint64_t n_slots, uint32_t *filter,
                    uint32_t filter_word_mask)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_slots) return;
    const unsigned long long e = slots[i];
    if (e == 0) return;
    const uint64_t h64 = tb_hash64(e & TB_PAIR_KEYMASK);
    atomicOr(&filter[(uint32_t)(h64 >> 32) & filter_word_mask], filter_bits<NBITS>(h64));
}

// ---- minimizer-addressed bitmap -------------------------------------------------------------
// (Tried and dropped: moving the phase-1 hashes onto the idle FMA pipe with multiply-fold
// mixers and a strand-symmetric k-mer hash that avoids the 64-bit min.  A plain "high half of
// a product" mixer is too linear -- the false-positive rate then grows with the set size --
// and the fold version needs 64-bit products whose register pressure cost more than the
// ALU-pipe slots it freed: 0.76 ms against 0.64 ms at C2.)
// When A u B is so large that no selective bitmap fits in L2, what limits the kernel is HBM
// again: one bitmap sector miss per window.  Consecutive windows of a read overlap in k-1
// bases, so they usually share their MINIMIZER (the canonical m-mer with the smallest hash
// among the w = k-m+1 m-mers of the k-mer).  Here the 32-byte bitmap SECTOR of a key is chosen
// by its minimizer and only the word/bits inside the sector by the k-mer itself: a lane that
// walks 32 consecutive windows then touches a new sector only every (w+1)/2 windows on
// average and finds the others in L1 -- 3-4x fewer L2 requests and HBM sector fetches, which
// also makes bitmaps far beyond the L2 size (2^31..2^33 bits) affordable.  m is large enough
// (15/16) that many minimizer values share a sector, which evens out the skew of minimizer
// popularity.  The minimizer is a pure function of the key VALUE, so stored keys (raw, Q1)
// and looked-up keys agree whatever strand or scratch quirk produced them.
// m: 16 for k >= 22; below that k-6 (seven m-mers per window), but not less than 13 while that
// leaves three m-mers per window -- with fewer than ~4^13 / 2 distinct minimizers the keys of a
// 10 M + 10 M set crowd into a fraction of the bitmap's sectors and a third to nine tenths of all
// windows pass it (measured at k = 15 .. 17 with m = k - 6: profiles/r2_ksweep_before_m13.jsonl)
// -- and never less than k/2, so that the first and the last m-mer of a k-mer together cover all
// of its bases (their hashes also address the word and the bits)
#ifndef TB_MINI_W
#define TB_MINI_W 7   // m-mers per window below k = 22 (measurement knob; capi.cu has the same default)
#endif
__host__ __device__ constexpr int tb_minimizer_m(int k)
{
    if (k >= 22) return 16;
    int m = k - (TB_MINI_W - 1);
    const int floor13 = k - 2 < 13 ? k - 2 : 13;
    if (m < floor13) m = floor13;
    const int half = (k + 1) / 2;
    return m > half ? m : half;
}

#ifndef TB_MIX
#define TB_MIX 0
#endif
__device__ __forceinline__ uint32_t mmer_mix(uint32_t c)
{
#if TB_MIX == 0
    c *= 0x9E3779B1u;
    c ^= c >> 15;
    c *= 0x85EBCA77u;
    c ^= c >> 13;
    return c;
#else
    // multiply-fold: the 64-bit product of the m-mer and an odd constant, high half xor low half,
    // twice.  Each round is one IMAD.WIDE (FMA pipe) and one LOP3 where the xor-shift round above
    // needs a shift and a xor on the ALU pipe, the pipe that bounds the kernel
    uint64_t p = (uint64_t)c * 0x9E3779B1u;
    c = (uint32_t)p ^ (uint32_t)(p >> 32);
    p = (uint64_t)c * 0x85EBCA77u;
    return (uint32_t)p ^ (uint32_t)(p >> 32);
#endif
}

// reverse complement of an m-mer held in the low 2m bits
__device__ __forceinline__ uint32_t mmer_revcomp(uint32_t f, int m)
{
    uint32_t x = __brev(~f);
    x = ((x & 0x55555555u) << 1) | ((x >> 1) & 0x55555555u);
    return x >> (32 - 2 * m);
}

// Bitmap coordinates of a k-mer, as pure functions of its integer key:
//   min_hash  smallest hash over its canonical m-mers (the minimizer) -> the sector
//   fh        hash of the first canonical m-mer + hash of the last one -> word in the sector and
//             the bit positions.  The two m-mers cover the whole k-mer (2m >= k), the sum is
//             symmetric under reverse complement (first and last swap), and in the kernel both
//             terms are already in the sliding window of m-mer hashes: phase 1 needs no k-mer
//             key at all, only m-mers.
// TB_MSCALE=1: the canonical m-mer is hashed as min(f, r) << (32 - 2m) -- the same order, so the
// same minimum -- which lets the kernel drop an m-mer's two unused top bits with a multiply (FMA
// pipe) instead of masking them off (ALU pipe).  m = 16 needs neither.
#ifndef TB_MSCALE
#define TB_MSCALE 1
#endif
struct KmerCoord {
    uint32_t min_hash, fh;
    uint32_t line_hash;  // two-level addressing only: smallest hash over the canonical 13-mers
};

// Two-level ("hier") addressing for bitmaps far beyond L2 (k = 20 .. 24).  Every L1 miss on the
// bitmap makes L2 fetch a whole 128-byte line from HBM, and the kernel's time at C3 / C5 follows
// the number of those line fetches, i.e. the number of minimizer runs.  A second, coarser
// minimizer -- over the k-12 canonical 13-mers of the k-mer instead of its k-m+1 m-mers -- changes
// less often along a read (2 / (k - 11) of the windows instead of 2 / (k - m + 2)); it picks the
// LINE, the m-mer minimizer picks the sector inside the line as before.  Neighbouring runs that
// share their 13-mer minimizer then find their sector in L2 (the line is already there).
#define TB_HIER_M 13
__device__ __forceinline__ uint32_t hier_line_hash(uint64_t key, int k)
{
    uint32_t best = 0xFFFFFFFFu;
    for (int i = 0; i + TB_HIER_M <= k; i++) {
        const uint32_t f = (uint32_t)(key >> (2 * i)) & ((1u << (2 * TB_HIER_M)) - 1u);
        const uint32_t r = mmer_revcomp(f, TB_HIER_M);
        const uint32_t h = mmer_mix((f < r ? f : r) << (32 - 2 * TB_HIER_M));
        best = h < best ? h : best;
    }
    return best;
}

__device__ __forceinline__ KmerCoord coord_of_key(uint64_t key, int k, int m, bool hier = false)
{
    const uint32_t mmask = m == 16 ? 0xFFFFFFFFu : ((1u << (2 * m)) - 1u);
    KmerCoord c;
    c.min_hash = 0xFFFFFFFFu;
    c.line_hash = hier ? hier_line_hash(key, k) : 0u;
    uint32_t first = 0, last = 0;
    for (int i = 0; i + m <= k; i++) {
        const uint32_t f = (uint32_t)(key >> (2 * i)) & mmask;
        const uint32_t r = mmer_revcomp(f, m);
#if TB_MSCALE
        const uint32_t h = mmer_mix((f < r ? f : r) << (32 - 2 * m));
#else
        const uint32_t h = mmer_mix(f < r ? f : r);
#endif
        c.min_hash = h < c.min_hash ? h : c.min_hash;
        if (i == 0) first = h;
        last = h;
    }
    c.fh = first + last;
    return c;
}

// word index of a key's bitmap word: sector from the minimizer, word in the sector from the k-mer.
// The minimum of several hashes is biased small; one odd multiply spreads it over the sectors
// again (the range reduction takes the HIGH half of a product, which depends on every bit of
// min_hash).  All of it runs as IMAD / IMAD.HI: the FMA pipe has room, the ALU pipe is what
// bounds the kernel -- dropping a xor-shift here and taking the word from a product instead of
// a shift+mask (4 ALU ops per window of 21) took the C2 kernel from 0.52 to 0.45 ms at the
// same pass rate (profiles/r1_kernel_variants_alu.txt).
__device__ __forceinline__ uint32_t mini_word_index(uint32_t min_hash, uint32_t fh, uint32_t n_sectors)
{
    return __umulhi(min_hash * 0xC2B2AE35u, n_sectors) * 8u + __umulhi(fh * 0x9E3779B1u, 8u);
}

// two-level form: line (4 sectors) from the 13-mer minimizer, sector in the line from the m-mer one
__device__ __forceinline__ uint32_t hier_word_index(uint32_t line_hash, uint32_t min_hash, uint32_t fh, uint32_t n_sectors)
{
    return __umulhi(line_hash * 0x85EBCA6Bu, n_sectors >> 2) * 32u + __umulhi(min_hash * 0xC2B2AE35u, 4u) * 8u +
           __umulhi(fh * 0x9E3779B1u, 8u);
}

__device__ __forceinline__ uint32_t coord_word_index(const KmerCoord &kc, uint32_t n_sectors, bool hier)
{
    return hier ? hier_word_index(kc.line_hash, kc.min_hash, kc.fh, n_sectors) : mini_word_index(kc.min_hash, kc.fh, n_sectors);
}

// The bits a k-mer sets in its bitmap word.  TB_BITTEST=1 takes the second position from the product
// that already selects the word (bits 24..28 of fh * 0x9E3779B1; bits 29..31 are the word), so that
// the kernel can test "both bits set" with two wrapping shifts of the loaded word and one LOP3
// instead of building a two-bit mask first.
#ifndef TB_BITTEST
#define TB_BITTEST 0
#endif
template <int NBITS>
__device__ __forceinline__ uint32_t mini_need(uint32_t fh)
{
#if TB_BITTEST == 0
    return filter_bits<NBITS>(fh);
#else
    uint32_t m = __funnelshift_l(0u, 1u, fh);
    if (NBITS == 2) m |= __funnelshift_l(0u, 1u, (fh * 0x9E3779B1u) >> 24);
    return m;
#endif
}

__device__ __forceinline__ uint32_t ld_filter_l1(const uint32_t *p, uint64_t policy)
{
    // allocate in L1 (the next windows re-read this sector), keep in L2; a 64-B L2 fetch
    // granularity (L2::64B) was measured and changes nothing here
    uint32_t v;
#if !defined(TB_BITMAP_LD) || TB_BITMAP_LD == 0
    asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(policy));
#elif TB_BITMAP_LD == 1  // measurement variants (tools/variants.sh)
    asm volatile("ld.global.nc.L2::cache_hint.L2::64B.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(policy));
#elif TB_BITMAP_LD == 2
    asm volatile("ld.global.nc.L2::cache_hint.L2::128B.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(policy));
#elif TB_BITMAP_LD == 3
    (void)policy;
    asm volatile("ld.global.nc.L1::evict_last.u32 %0, [%1];" : "=r"(v) : "l"(p));
#else
    (void)policy;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
#endif
    return v;
}

// Tuning notes for k_classify_pair_mini (all measured at C2, profiles/r1_kernel_variants_alu.txt):
// the kernel sits between the ALU pipe (~17 ALU ops per window in phase 1) and load latency, and
// ptxas' schedule decides which side wins -- fewer instructions is not always faster.  Kept:
// product-based sector/word selection (0.52 -> 0.45 ms), 2 survivors per phase-2 round.
// Measured and dropped: sliding minimum from shared 3- and 9-wide partial minima (1 op fewer
// per window at k = 21, 5 fewer at k = 31, but 80 instead of 72 registers: 0.49 ms); bit
// positions from the high bits of the hash without its last xor-shift (0.49 ms); m = 16 at
// k = 21 (0.443 ms, but 14% more sector misses on sets beyond L2); a persistent tile loop with
// the grid sized to the resident warps (0.51-0.56 ms: dynamic block scheduling balances
// better); 7, 8 or 9 blocks per SM through launch bounds (0.51 / 0.53 / 0.81 ms).
#ifndef TB_MINI_MIN_BLOCKS
#define TB_MINI_MIN_BLOCKS 6
#endif
template <int K, int U, int NBITS, int P2, bool HIER = false>
#ifndef TB_HIER_MIN_BLOCKS
#define TB_HIER_MIN_BLOCKS 4
#endif
__global__ void __launch_bounds__(128, HIER ? TB_HIER_MIN_BLOCKS : TB_MINI_MIN_BLOCKS)
k_classify_pair_mini(tb_batch_view b, int k_runtime, tb_pair_view P, const uint32_t *__restrict__ filter,
                     uint32_t n_sectors)
{
    const int k = K ? K : k_runtime;
    const int m = tb_minimizer_m(k);
    constexpr int MC = K ? tb_minimizer_m(K) : 1;        // compile-time m for the unrolled path
    constexpr int WC = K ? K - MC + 1 : 1;               // m-mers per k-mer
    const uint64_t kmask = (1ull << (2 * k)) - 1ull;  // k <= 31
    const int lane = threadIdx.x & 31;
    const int64_t tile = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (tile >= b.n_tiles) return;  // warp-uniform
    LaneCtx c;
    lane_stream(b, tile, lane, c);
    uint64_t keep_policy;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep_policy));
    // the general path below walks the windows one by one and wants to know which exist; the
    // unrolled path computes all 32 and masks afterwards, while the read boundaries arrive
    const bool unrolled = K != 0 && c.clean;
    if (!unrolled) lane_bounds(b, tile, k, c);

    uint32_t pass = 0;
    if (unrolled ? c.p0 < b.total_bases : c.vmask != 0) {
        if (unrolled) {
            // ---- all bases valid: m-mer hashes straight from the two streams, sliding minimum.
            // Two passes of 16 windows over the same unrolled body (the second on the span
            // moved by 16 bases) keep the loop inside the instruction cache.
            constexpr uint32_t MMASK = MC == 16 ? 0xFFFFFFFFu : ((1u << (2 * MC)) - 1u);
            constexpr uint32_t MSCALE = 1u << (32 - 2 * MC);
            (void)MMASK;
            (void)MSCALE;
            LaneCtx h = c;  // half-span view: window j of the half starts at base j of (f0,f1)
#ifndef TB_REVODD
#define TB_REVODD 1  // measured: C3 1.350 -> 1.297 ms, k = 31 at C2 0.543 -> 0.533 ms (profiles/r2_variants_alu.txt)
#endif
#if TB_REVODD
            // Odd lanes walk their 32 windows backwards by working on the other strand: window j of
            // the reverse-complemented span (moved by 33-K bases) is window 31-j of the lane, and the
            // bitmap coordinates do not depend on the strand.  A lane's last window then meets its
            // right neighbour's first window in the SAME load instruction -- they usually share
            // their minimizer, hence their sector -- instead of missing L1 a second time much later.
            if (lane & 1) {
                constexpr int S = K ? 2 * (33 - K) : 2;  // 4 .. 40 bits (K = 0 never takes this path)
                h.f0 = (c.r_lo >> S) | (c.r_hi << (64 - S));
                h.f1 = c.r_hi >> S;
                h.r_hi = (c.f1 << S) | (c.f0 >> (64 - S));
                h.r_lo = c.f0 << S;
            }
#endif
            // two-level addressing: 13-mer hashes of the first W1-1 positions, computed directly;
            // inside the loop the newest 13-mer of a window is the top 26 bits of its newest m-mer
            constexpr int W1 = HIER ? K - TB_HIER_M + 1 : 1;  // 13-mers per k-mer
            uint32_t head1[W1 > 1 ? W1 - 1 : 1];
            if (HIER) {
#pragma unroll
                for (int q = 0; q < W1 - 1; q++) {
                    const uint32_t f = (uint32_t)extract128(h.f0, h.f1, 2 * q) << (32 - 2 * TB_HIER_M);
                    const uint32_t r = (uint32_t)extract128(h.r_lo, h.r_hi, 2 * (64 - q - TB_HIER_M)) << (32 - 2 * TB_HIER_M);
                    head1[q] = mmer_mix(f < r ? f : r);
                }
            }
            // m-mer hashes of the first WC-1 positions of a half; the second half inherits them
            // from the first (its positions 16 .. 16+WC-2)
            uint32_t head[WC > 1 ? WC - 1 : 1];
#pragma unroll
            for (int q = 0; q < WC - 1; q++) {
#if TB_MSCALE
                const uint32_t f = (uint32_t)extract128(h.f0, h.f1, 2 * q) * MSCALE;
                const uint32_t r = (uint32_t)extract128(h.r_lo, h.r_hi, 2 * (64 - q - MC)) * MSCALE;
#else
                const uint32_t f = (uint32_t)extract128(h.f0, h.f1, 2 * q) & MMASK;
                const uint32_t r = (uint32_t)extract128(h.r_lo, h.r_hi, 2 * (64 - q - MC)) & MMASK;
#endif
                head[q] = mmer_mix(f < r ? f : r);
            }
#pragma unroll 1
            for (int half = 0; half < 2; half++) {
                uint32_t hm[16 + WC];
#pragma unroll
                for (int q = 0; q < WC - 1; q++) hm[q] = head[q];
                uint32_t hm1[16 + W1];
                if (HIER) {
#pragma unroll
                    for (int q = 0; q < W1 - 1; q++) hm1[q] = head1[q];
                }
                uint32_t half_pass = 0;
#pragma unroll
                for (int g = 0; g < 16; g += U) {
                    uint32_t need[U], wd[U], widx[U];
#if TB_BITTEST
                    uint32_t need2[U];
#endif
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        const int j = g + u, q = j + WC - 1;
                        {
                            const uint32_t raw_r = (uint32_t)extract128(h.r_lo, h.r_hi, 2 * (64 - q - MC));
#if TB_MSCALE
                            const uint32_t f = (uint32_t)extract128(h.f0, h.f1, 2 * q) * MSCALE;
                            const uint32_t r = raw_r * MSCALE;
#else
                            const uint32_t f = (uint32_t)extract128(h.f0, h.f1, 2 * q) & MMASK;
                            const uint32_t r = raw_r & MMASK;
#endif
                            hm[q] = mmer_mix(f < r ? f : r);
                            if (HIER) {
                                // the last 13 bases of this m-mer: the top 26 bits of the left-aligned
                                // forward m-mer, the low 26 bits of its reverse complement
                                static_assert(!HIER || TB_MSCALE, "two-level addressing reads the left-aligned m-mer");
                                const uint32_t f1 = f & ~((1u << (32 - 2 * TB_HIER_M)) - 1u);
                                const uint32_t r1 = raw_r * (1u << (32 - 2 * TB_HIER_M));
                                hm1[j + W1 - 1] = mmer_mix(f1 < r1 ? f1 : r1);
                            }
                        }
                        uint32_t mn = hm[j];
#pragma unroll
                        for (int t = 1; t + 1 < WC; t += 2) mn = __vimin3_u32(mn, hm[j + t], hm[j + t + 1]);
                        if ((WC & 1) == 0) mn = min(mn, hm[j + WC - 1]);
                        const uint32_t fh = hm[j] + hm[j + WC - 1];  // first + last m-mer of the window
#if TB_BITTEST
                        const uint32_t t = fh * 0x9E3779B1u;
                        need[u] = fh;                   // shift amounts: the funnel shifts use their low 5 bits
                        need2[u] = __umulhi(t, 256u);   // t >> 24 on the FMA pipe
                        widx[u] = __umulhi(mn * 0xC2B2AE35u, n_sectors) * 8u + __umulhi(t, 8u);
#else
                        need[u] = filter_bits<NBITS>(fh);
                        if (HIER) {
                            uint32_t mn1 = hm1[j];
#pragma unroll
                            for (int t = 1; t + 1 < W1; t += 2) mn1 = __vimin3_u32(mn1, hm1[j + t], hm1[j + t + 1]);
                            if ((W1 & 1) == 0) mn1 = min(mn1, hm1[j + W1 - 1]);
                            widx[u] = hier_word_index(mn1, mn, fh, n_sectors);
                        } else {
                            widx[u] = mini_word_index(mn, fh, n_sectors);
                        }
#endif
                    }
#pragma unroll
                    for (int u = 0; u < U; u++) wd[u] = ld_filter_l1(filter + widx[u], keep_policy);
#if TB_BITTEST
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        uint32_t bit = __funnelshift_r(wd[u], 0u, need[u]);
                        if (NBITS == 2) bit &= __funnelshift_r(wd[u], 0u, need2[u]);
                        half_pass += (bit & 1u) << (g + u);
                    }
#else
#pragma unroll
                    for (int u = 0; u < U; u++)
                        if ((need[u] & ~wd[u]) == 0) half_pass |= 1u << (g + u);
#endif
                }
                pass |= half_pass << (16 * half);
#pragma unroll
                for (int q = 0; q < WC - 1; q++) head[q] = hm[16 + q];
                if (HIER) {
#pragma unroll
                    for (int q = 0; q < W1 - 1; q++) head1[q] = hm1[16 + q];
                }
                // next half: drop 16 bases from the front of the span (and of its reverse complement)
                h.f0 = (h.f0 >> 32) | (h.f1 << 32);
                h.f1 >>= 32;
                h.r_hi = (h.r_hi << 32) | (h.r_lo >> 32);
                h.r_lo <<= 32;
            }
#if TB_REVODD
            if (lane & 1) pass = __brev(pass);
#endif
            lane_bounds(b, tile, k, c);
        } else {
            // ---- a non-ACGT base nearby (or an unusual k): minimizer from the key itself
            for (int j = 0; j < 32; j++) {
                if (!((c.vmask >> j) & 1u)) continue;
                const KmerCoord kc = coord_of_key(window_key<0>(c, j, k, kmask), k, m, HIER);
                const uint32_t need = mini_need<NBITS>(kc.fh);
                const uint32_t wd = ld_filter_l1(filter + coord_word_index(kc, n_sectors, HIER), keep_policy);
                pass |= ((wd & need) == need ? 1u : 0u) << j;
            }
        }
        pass &= c.vmask;
    }
    {
        const unsigned np = __reduce_add_sync(0xFFFFFFFFu, (unsigned)__popc(pass));
        if (lane == 0 && np) atomicAdd(b.n_pass, (unsigned long long)np);
    }

    uint32_t hit_a = 0, hit_b = 0;
    uint32_t mm = pass;
    if (P2 == 0) {
        // ---- phase 2, warp-cooperative: the warp's survivors go through a queue in shared memory
        // so that the table probes -- the part that waits for HBM -- are spread evenly over the
        // lanes whatever lane the survivors came from.  With reads that carry real haplotype
        // k-mers a few per cent of all windows survive and they cluster (a heterozygous site makes
        // k consecutive windows hit): lane by lane the warp would run as many rounds as its
        // unluckiest lane needs.  Fill: every lane that has survivors left pushes one key per step,
        // four steps (128 entries at most) per batch.  Probe: entry e is handled by lane e mod 32,
        // two loads in flight.  Read back: every lane collects the tags of what it pushed.
        __shared__ unsigned long long q_key[4][128];
        __shared__ unsigned char q_tag[4][128];
        const int wi = threadIdx.x >> 5;
        const unsigned lt = (1u << lane) - 1u;
        while (__any_sync(0xFFFFFFFFu, mm != 0)) {
            uint32_t pos_pack = 0, j_pack = 0;
            int cnt = 0, n = 0;
#pragma unroll
            for (int step = 0; step < 4; step++) {
                const bool have = mm != 0;
                const unsigned b = __ballot_sync(0xFFFFFFFFu, have);
                if (have) {
                    const int j = __ffs(mm) - 1;
                    mm &= mm - 1;
                    const int pos = n + __popc(b & lt);
                    q_key[wi][pos] = window_key<0>(c, j, k, kmask);
                    pos_pack |= (uint32_t)pos << (8 * cnt);
                    j_pack |= (uint32_t)j << (5 * cnt);
                    cnt++;
                }
                n += __popc(b);
            }
            __syncwarp();
            for (int base = 0; base < n; base += 64) {
                const int e0 = base + lane, e1 = base + 32 + lane;
                const bool v0 = e0 < n, v1 = e1 < n;
                const uint64_t key0 = v0 ? q_key[wi][e0] : 0, key1 = v1 ? q_key[wi][e1] : 0;
                const uint32_t bkt0 = tb_pair_bucket_of(tb_hash64(key0), P.n_buckets);
                const uint32_t bkt1 = tb_pair_bucket_of(tb_hash64(key1), P.n_buckets);
                tb_bucket bk0, bk1;
                if (v0) bk0 = tb_load_bucket(P.table + bkt0);
                if (v1) bk1 = tb_load_bucket(P.table + bkt1);
                if (v0) {
                    uint32_t tags = tb_pair_tags(bk0, key0);
                    if (!tags && tb_pair_bucket_full(bk0)) tags = tb_pair_overflow(P, key0, bkt0);
                    q_tag[wi][e0] = (unsigned char)tags;
                }
                if (v1) {
                    uint32_t tags = tb_pair_tags(bk1, key1);
                    if (!tags && tb_pair_bucket_full(bk1)) tags = tb_pair_overflow(P, key1, bkt1);
                    q_tag[wi][e1] = (unsigned char)tags;
                }
            }
            __syncwarp();
            for (int q = 0; q < cnt; q++) {
                const uint32_t tags = q_tag[wi][(pos_pack >> (8 * q)) & 0xFFu];
                const int j = (j_pack >> (5 * q)) & 31;
                hit_a |= (tags & 1u) << j;
                hit_b |= ((tags == 2u) ? 1u : 0u) << j;
            }
            __syncwarp();
        }
    }
    // ---- phase 2, lane by lane (as in k_classify_pair).  P2 survivors per lane and round: the
    // whole warp pays for every slot, so 2 where ~1% of the windows pass (most lanes hold none or
    // one) and 4 where more do
    while (P2 != 0 && mm) {
        constexpr int PW = P2 ? P2 : 1;
        int j[PW];
        uint64_t key[PW];
        uint32_t bkt[PW];
        tb_bucket bk[PW];
#pragma unroll
        for (int q = 0; q < PW; q++) {
            j[q] = mm ? __ffs(mm) - 1 : -1;
            if (mm) mm &= mm - 1;
            key[q] = window_key<0>(c, j[q] < 0 ? 0 : j[q], k, kmask);
            bkt[q] = tb_pair_bucket_of(tb_hash64(key[q]), P.n_buckets);
        }
#pragma unroll
        for (int q = 0; q < PW; q++)
            if (j[q] >= 0) bk[q] = tb_load_bucket(P.table + bkt[q]);
#pragma unroll
        for (int q = 0; q < PW; q++) {
            if (j[q] < 0) continue;
            uint32_t tags = tb_pair_tags(bk[q], key[q]);
            if (!tags && tb_pair_bucket_full(bk[q])) tags = tb_pair_overflow(P, key[q], bkt[q]);
            hit_a |= (tags & 1u) << j[q];
            hit_b |= ((tags == 2u) ? 1u : 0u) << j[q];
        }
    }
    lane_attribute(b, c, lane, hit_a, hit_b);
}

// Measurement aid (bench.py's roofline; never on the timed path): how many bitmap SECTOR fetches
// the batch needs when every lane re-uses a sector across consecutive windows that share it --
// per lane, the number of times the sector changes along its 32 windows.  Same work split and
// same coordinates as k_classify_pair_mini, computed the slow way from each window's key.
__global__ void __launch_bounds__(128)
k_count_sector_runs(tb_batch_view b, int k, uint32_t n_sectors, int hier, unsigned long long *out)
{
    const int m = tb_minimizer_m(k);
    const uint64_t kmask = (1ull << (2 * k)) - 1ull;
    const int lane = threadIdx.x & 31;
    const int64_t tile = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (tile >= b.n_tiles) return;
    LaneCtx c;
    lane_setup(b, tile, lane, k, c);
    unsigned runs = 0;
    uint32_t prev = 0xFFFFFFFFu;
    for (int j = 0; j < 32; j++) {
        if (!((c.vmask >> j) & 1u)) continue;
        // two-level addressing: what reaches HBM is the LINE (its sectors are found in L2 afterwards)
        const KmerCoord kc = coord_of_key(window_key<0>(c, j, k, kmask), k, m, hier != 0);
        const uint32_t sector = hier ? coord_word_index(kc, n_sectors, true) >> 5 : __umulhi(kc.min_hash * 0xC2B2AE35u, n_sectors);
        if (sector != prev) runs++;
        prev = sector;
    }
    runs = __reduce_add_sync(0xFFFFFFFFu, runs);
    if (lane == 0 && runs) atomicAdd(out, (unsigned long long)runs);
}

template <int NBITS>
__global__ void __launch_bounds__(256)
k_pair_filter_build_mini(const unsigned long long *__restrict__ slots, uint64_t n_slots, uint32_t *filter,
                         uint32_t n_sectors, int k, int hier)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_slots) return;
    const unsigned long long e = slots[i];
    if (e == 0) return;
    const KmerCoord kc = coord_of_key(e & TB_PAIR_KEYMASK, k, tb_minimizer_m(k), hier != 0);
    atomicOr(&filter[coord_word_index(kc, n_sectors, hier != 0)], mini_need<NBITS>(kc.fh));
}

// ---- bitmap construction: one thread per table slot ---------------------------------
__global__ void __launch_bounds__(256)
k_filter_build(const unsigned long long *__restrict__ slots, uint64_t n_slots, uint32_t *filter,
               uint32_t filter_mask)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_slots) return;
    const unsigned long long key = slots[i];
    if (key == TB_EMPTY) return;
    const uint32_t h = tb_hash32(key);
    atomicOr(&filter[(h & filter_mask) >> 5], 1u << (h & 31u));
}

// ============================================================================
// k_fixup: the Q3-UB corner, literally.  For a read with a non-ACGT byte among
// its first k bases the reference's reverse-complement scratch holds slots that
// were never written; the contract is "they read as 0".  One thread
// per flagged read replays windows [0, skip) with a sticky k-digit buffer,
// exactly as c/kmers.c:287-294 does, and adds the hits to the read's counts.
// ============================================================================
__global__ void __launch_bounds__(128)
k_fixup(tb_batch_view b, int k, tb_set_view A, tb_set_view B)
{
    const unsigned n = *b.ub_count;
    const uint64_t kmask = k == 32 ? ~0ull : ((1ull << (2 * k)) - 1ull);
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t r = b.ub_list[i];
        const unsigned char *read = b.ub_ascii_off ? b.ub_ascii + b.ub_ascii_off[i] : b.ascii + b.offsets[r];
        const uint32_t skip = b.skip[r];
        uint64_t rc = 0;  // sticky digits, slot s at bits 2s, zero-initialised
        int ca = 0, cb = 0;
        // A read that STARTS with a long run of non-ACGT bytes: every window that lies entirely
        // inside the run has forward digits 0 and a reverse-complement buffer nobody has written
        // yet, i.e. key 0 -- one lookup for all of them instead of k byte tests per window.
        uint32_t first = 0;
        {
            uint32_t run = 0;
            while (run < skip + (uint32_t)k - 1u && !tb_base_valid(read[run])) run++;
            if (run >= (uint32_t)k) {
                first = run - (uint32_t)k + 1u < skip ? run - (uint32_t)k + 1u : skip;
                if (tb_contains(A, 0))
                    ca += (int)first;
                else if (tb_contains(B, 0))
                    cb += (int)first;
            }
        }
        for (uint32_t win = first; win < skip; win++) {
            uint64_t fwd = 0;
            for (int j = 0; j < k; j++) {
                unsigned char c = read[win + j];
                if (tb_base_valid(c)) {
                    uint64_t code = tb_base_code(c);
                    fwd |= code << (2 * j);
                    const int s = k - 1 - j;
                    rc = (rc & ~(3ull << (2 * s))) | ((3ull - code) << (2 * s));
                }
            }
            const uint64_t key = (fwd < rc ? fwd : rc) & kmask;
            if (tb_contains(A, key))
                ca++;
            else if (tb_contains(B, key))
                cb++;
        }
        if (ca) atomicAdd(&b.count_a[r], ca);
        if (cb) atomicAdd(&b.count_b[r], cb);
    }
}

// ============================================================================
// k_finalize: classify_by_kmers.py:104-115 in IEEE fp64 on device.
// ============================================================================
__global__ void __launch_bounds__(256)
k_finalize(tb_batch_view b, double scale_a, double scale_b)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r == 0) *b.ub_count = 0;  // ready for the next batch on this slot
    if (r >= b.n_reads) return;
    const double a = __dmul_rn((double)b.count_a[r], scale_a);
    const double c = __dmul_rn((double)b.count_b[r], scale_b);
    if (b.score_a) b.score_a[r] = a;
    if (b.score_b) b.score_b[r] = c;
    if (b.bin) b.bin[r] = a > c ? 'A' : (c > a ? 'B' : 'U');
}

int g_unroll = 4;
int g_tpb = 128;
int g_filter_unroll = 16;  // hash-addressed filtered kernel (k = 32 or TB_PAIR=0)
int g_filter_mode = 2;

template <int K>
cudaError_t launch_classify_k(const tb_batch_view &b, int k, tb_set_view A, tb_set_view B,
                              cudaStream_t stream)
{
    const int tpb = g_tpb;
    const int warps = tpb / 32;
    const unsigned blocks = (unsigned)((b.n_tiles + warps - 1) / warps);
    switch (g_unroll) {
    case 4: k_classify<K, 4><<<blocks, tpb, 0, stream>>>(b, k, A, B); break;
    case 16: k_classify<K, 16><<<blocks, tpb, 0, stream>>>(b, k, A, B); break;
    default: k_classify<K, 8><<<blocks, tpb, 0, stream>>>(b, k, A, B); break;
    }
    return cudaGetLastError();
}

template <int K>
cudaError_t launch_filtered_k(const tb_batch_view &b, int k, tb_set_view A, tb_set_view B,
                              const uint32_t *filter, uint32_t filter_mask, cudaStream_t stream)
{
    const int tpb = g_tpb;
    const int warps = tpb / 32;
    const unsigned blocks = (unsigned)((b.n_tiles + warps - 1) / warps);
    switch (g_filter_unroll) {
    case 8: k_classify_filtered<K, 8><<<blocks, tpb, 0, stream>>>(b, k, A, B, filter, filter_mask, g_filter_mode); break;
    case 32: k_classify_filtered<K, 32><<<blocks, tpb, 0, stream>>>(b, k, A, B, filter, filter_mask, g_filter_mode); break;
    default: k_classify_filtered<K, 16><<<blocks, tpb, 0, stream>>>(b, k, A, B, filter, filter_mask, g_filter_mode); break;
    }
    return cudaGetLastError();
}

template <int K>
cudaError_t launch_pair_k(const tb_batch_view &b, int k, tb_pair_view P, const uint32_t *filter,
                          uint32_t word_mask, int n_bits, cudaStream_t stream)
{
    const int tpb = g_tpb;
    const int warps = tpb / 32;
    const unsigned blocks = (unsigned)((b.n_tiles + warps - 1) / warps);
    if (n_bits == 2)
        k_classify_pair<K, 16, 2><<<blocks, tpb, 0, stream>>>(b, k, P, filter, word_mask, g_filter_mode);
    else
        k_classify_pair<K, 16, 1><<<blocks, tpb, 0, stream>>>(b, k, P, filter, word_mask, g_filter_mode);
    return cudaGetLastError();
}

template <int K>
cudaError_t launch_pair_mini_k(const tb_batch_view &b, int k, tb_pair_view P, const uint32_t *filter,
                               uint32_t n_sectors, int n_bits, int p2_width, cudaStream_t stream)
{
    const int tpb = g_tpb;
    const int warps = tpb / 32;
    const unsigned blocks = (unsigned)((b.n_tiles + warps - 1) / warps);
    // 8 bitmap loads in flight per lane; 4 and 16 measured the same (profiles/r1_kernel_variants_minimizer*.txt)
    if (n_bits == 2) {
        if (p2_width >= 4)
            k_classify_pair_mini<K, 8, 2, 4><<<blocks, tpb, 0, stream>>>(b, k, P, filter, n_sectors);
        else if (p2_width == 1 && tpb == 128)  // warp-cooperative phase 2 (its queue is sized for 4 warps)
            k_classify_pair_mini<K, 8, 2, 0><<<blocks, tpb, 0, stream>>>(b, k, P, filter, n_sectors);
        else
            k_classify_pair_mini<K, 8, 2, 2><<<blocks, tpb, 0, stream>>>(b, k, P, filter, n_sectors);
    } else {
        if (p2_width >= 4)
            k_classify_pair_mini<K, 8, 1, 4><<<blocks, tpb, 0, stream>>>(b, k, P, filter, n_sectors);
        else
            k_classify_pair_mini<K, 8, 1, 2><<<blocks, tpb, 0, stream>>>(b, k, P, filter, n_sectors);
    }
    return cudaGetLastError();
}

}  // namespace

void tb_classify_config(int unroll, int threads_per_block, int filter_unroll, int filter_mode)
{
    if (filter_mode >= 0 && filter_mode <= 2) g_filter_mode = filter_mode;
    if (filter_unroll == 8 || filter_unroll == 16 || filter_unroll == 32) g_filter_unroll = filter_unroll;
    if (unroll == 4 || unroll == 8 || unroll == 16) g_unroll = unroll;
    if (threads_per_block == 32 || threads_per_block == 64 || threads_per_block == 128)
        g_tpb = threads_per_block;
}

cudaError_t tb_launch_pack(const tb_batch_view &b, int k, cudaStream_t stream)
{
    if (b.n_tiles == 0) return cudaSuccess;
    // one extra tile so that F[n_tiles*32] and dirty[n_tiles], which lane 31 of the
    // last tile reads, are always (re)written; 256 threads per block = 8 tiles
    const unsigned blocks = (unsigned)(((b.n_tiles + 1) * 32 + 255) / 256);
    k_pack<<<blocks, 256, 0, stream>>>(b, k);
    return cudaGetLastError();
}

cudaError_t tb_launch_index(const tb_batch_view &b, int k, cudaStream_t stream)
{
    if (b.n_reads == 0) return cudaSuccess;
    k_index<<<(unsigned)((b.n_reads + 255) / 256), 256, 0, stream>>>(b, k);
    return cudaGetLastError();
}

cudaError_t tb_launch_index_packed(const tb_batch_view &b, const uint32_t *d_dirty_idx,
                                   const uint64_t *d_e_words, int64_t n_dirty, const uint32_t *d_ub_read,
                                   const uint32_t *d_ub_skip, int64_t n_ub, cudaStream_t stream)
{
    if (b.n_reads == 0) return cudaSuccess;
    k_index_packed<<<(unsigned)((b.n_reads + 255) / 256), 256, 0, stream>>>(b);
    const int64_t n = n_dirty > n_ub ? n_dirty : n_ub;
    // always launched (n >= 1 thread) so that ub_count is set even when nothing is rare
    k_apply_rare<<<(unsigned)((n + 256) / 256), 256, 0, stream>>>(b, d_dirty_idx, d_e_words, n_dirty,
                                                                 d_ub_read, d_ub_skip, n_ub);
    return cudaGetLastError();
}

cudaError_t tb_launch_classify(const tb_batch_view &b, int k, tb_set_view A, tb_set_view B,
                               cudaStream_t stream)
{
    if (b.n_tiles == 0) return cudaSuccess;
    switch (k) {
    case 21: return launch_classify_k<21>(b, k, A, B, stream);
    case 31: return launch_classify_k<31>(b, k, A, B, stream);
    default: return launch_classify_k<0>(b, k, A, B, stream);
    }
}

cudaError_t tb_launch_classify_filtered(const tb_batch_view &b, int k, tb_set_view A, tb_set_view B,
                                        const uint32_t *filter, uint32_t filter_mask,
                                        cudaStream_t stream)
{
    if (b.n_tiles == 0) return cudaSuccess;
    switch (k) {
    case 21: return launch_filtered_k<21>(b, k, A, B, filter, filter_mask, stream);
    case 31: return launch_filtered_k<31>(b, k, A, B, filter, filter_mask, stream);
    default: return launch_filtered_k<0>(b, k, A, B, filter, filter_mask, stream);
    }
}

cudaError_t tb_launch_classify_pair(const tb_batch_view &b, int k, tb_pair_view P,
                                    const uint32_t *filter, uint32_t filter_word_mask, int n_bits,
                                    cudaStream_t stream)
{
    if (b.n_tiles == 0) return cudaSuccess;
    switch (k) {
    case 21: return launch_pair_k<21>(b, k, P, filter, filter_word_mask, n_bits, stream);
    case 31: return launch_pair_k<31>(b, k, P, filter, filter_word_mask, n_bits, stream);
    default: return launch_pair_k<0>(b, k, P, filter, filter_word_mask, n_bits, stream);
    }
}

cudaError_t tb_launch_classify_pair_mini(const tb_batch_view &b, int k, tb_pair_view P,
                                         const uint32_t *filter, uint32_t n_sectors, int n_bits,
                                         int p2_width, int hier, cudaStream_t stream)
{
    if (b.n_tiles == 0) return cudaSuccess;
    if (hier) {
        // two-level addressing (the bitmap was built with it): k = 20 .. 24, default configuration only
        const unsigned blocks = (unsigned)((b.n_tiles + 3) / 4);
        switch (k) {
#define TB_HIER_CASE(KK) \
        case KK: k_classify_pair_mini<KK, 8, 2, 2, true><<<blocks, 128, 0, stream>>>(b, k, P, filter, n_sectors); \
                 return cudaGetLastError();
        TB_HIER_CASE(20) TB_HIER_CASE(21) TB_HIER_CASE(22) TB_HIER_CASE(23) TB_HIER_CASE(24)
#undef TB_HIER_CASE
        default: return cudaErrorInvalidValue;
        }
    }
    switch (k) {
    case 21: return launch_pair_mini_k<21>(b, k, P, filter, n_sectors, n_bits, p2_width, stream);
    case 31: return launch_pair_mini_k<31>(b, k, P, filter, n_sectors, n_bits, p2_width, stream);
    default: break;
    }
    // every other k from 13 to 31 gets the same unrolled phase 1 (all shifts and the number of
    // m-mers per window are compile-time constants) in the default configuration: two bitmap
    // bits per key, two survivors per phase-2 round.  The experiment knobs (TB_FILTER_BITS=1,
    // TB_P2_WIDTH=4) and k < 13 fall back to the runtime-k body, which computes every window's
    // minimizer from its key: same results, several times slower.
    if (n_bits == 2 && p2_width < 4) {
        const unsigned blocks = (unsigned)((b.n_tiles + (g_tpb / 32) - 1) / (g_tpb / 32));
        switch (k) {
#define TB_MINI_CASE(KK) \
        case KK: k_classify_pair_mini<KK, 8, 2, 2><<<blocks, g_tpb, 0, stream>>>(b, k, P, filter, n_sectors); \
                 return cudaGetLastError();
        TB_MINI_CASE(13) TB_MINI_CASE(14) TB_MINI_CASE(15) TB_MINI_CASE(16) TB_MINI_CASE(17)
        TB_MINI_CASE(18) TB_MINI_CASE(19) TB_MINI_CASE(20) TB_MINI_CASE(22) TB_MINI_CASE(23)
        TB_MINI_CASE(24) TB_MINI_CASE(25) TB_MINI_CASE(26) TB_MINI_CASE(27) TB_MINI_CASE(28)
        TB_MINI_CASE(29) TB_MINI_CASE(30)
#undef TB_MINI_CASE
        default: break;
        }
    }
    return launch_pair_mini_k<0>(b, k, P, filter, n_sectors, n_bits, p2_width, stream);
}

cudaError_t tb_launch_count_sector_runs(const tb_batch_view &b, int k, uint32_t n_sectors, int hier,
                                        unsigned long long *d_out, cudaStream_t stream)
{
    if (b.n_tiles == 0) return cudaSuccess;
    k_count_sector_runs<<<(unsigned)((b.n_tiles + 3) / 4), 128, 0, stream>>>(b, k, n_sectors, hier, d_out);
    return cudaGetLastError();
}

cudaError_t tb_launch_pair_filter_build_mini(const uint64_t *d_pair_table, uint64_t n_slots, uint32_t *filter,
                                             uint32_t n_sectors, int n_bits, int k, int hier, cudaStream_t stream)
{
    if (n_slots == 0) return cudaSuccess;
    const unsigned blocks = (unsigned)((n_slots + 255) / 256);
    const unsigned long long *slots = reinterpret_cast<const unsigned long long *>(d_pair_table);
    if (n_bits == 2)
        k_pair_filter_build_mini<2><<<blocks, 256, 0, stream>>>(slots, n_slots, filter, n_sectors, k, hier);
    else
        k_pair_filter_build_mini<1><<<blocks, 256, 0, stream>>>(slots, n_slots, filter, n_sectors, k, hier);
    return cudaGetLastError();
}

cudaError_t tb_launch_pair_filter_build(const uint64_t *d_pair_table, uint64_t n_slots, uint32_t *filter,
                                        uint32_t filter_word_mask, int n_bits, cudaStream_t stream)
{
    if (n_slots == 0) return cudaSuccess;
    const unsigned blocks = (unsigned)((n_slots + 255) / 256);
    const unsigned long long *slots = reinterpret_cast<const unsigned long long *>(d_pair_table);
    if (n_bits == 2)
        k_pair_filter_build<2><<<blocks, 256, 0, stream>>>(slots, n_slots, filter, filter_word_mask);
    else
        k_pair_filter_build<1><<<blocks, 256, 0, stream>>>(slots, n_slots, filter, filter_word_mask);
    return cudaGetLastError();
}

cudaError_t tb_launch_filter_build(const uint64_t *d_table, uint64_t n_slots, uint32_t *filter,
                                   uint32_t filter_mask, cudaStream_t stream)
{
    if (n_slots == 0) return cudaSuccess;
    k_filter_build<<<(unsigned)((n_slots + 255) / 256), 256, 0, stream>>>(
        reinterpret_cast<const unsigned long long *>(d_table), n_slots, filter, filter_mask);
    return cudaGetLastError();
}

cudaError_t tb_launch_fixup(const tb_batch_view &b, int k, tb_set_view A, tb_set_view B,
                            cudaStream_t stream)
{
    if (b.n_reads == 0) return cudaSuccess;
    k_fixup<<<64, 128, 0, stream>>>(b, k, A, B);
    return cudaGetLastError();
}

cudaError_t tb_launch_finalize(const tb_batch_view &b, double scale_a, double scale_b,
                               cudaStream_t stream)
{
    const int64_t n = b.n_reads > 0 ? b.n_reads : 1;
    k_finalize<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(b, scale_a, scale_b);
    return cudaGetLastError();
}
