// kmer_device.cuh -- device-side building blocks shared by the set-build and
// classify kernels (sm_100a).
//
// Table layout (HBM): n_buckets x 32-byte buckets, each 4 x u64 RAW k-mer keys
// (LSB-first 2-bit encoding, reference c/kmers.c:50-72), empty slot = ~0.
// One bucket == one 32-byte DRAM/L2 sector, fetched by a single 256-bit load
// (ld.global.nc.v4.u64 -> LDG.E.ENL2.256.CONSTANT).  Slots inside a bucket fill
// in order 0..3 (insert only ever CASes an EMPTY slot after seeing the earlier
// ones occupied), so "slot 3 is empty" <=> "bucket not full" <=> a miss can
// stop here.  Overflow goes to the next bucket (linear probing over buckets).
//
// Membership does not depend on the hash or layout (the reference's
// hash_function, c/kmers.c:98-103, is not observable), so the device hash is a
// cheap 32-bit mixer and the bucket is mulhi(hash, n_buckets).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define TB_EMPTY 0xFFFFFFFFFFFFFFFFull
#define TB_SLOTS 4

struct __align__(32) tb_bucket {
    unsigned long long s[TB_SLOTS];
};

struct tb_set_view {
    const tb_bucket *table;
    uint32_t n_buckets;
    uint32_t has_all_ones;
};

__host__ __device__ __forceinline__ uint32_t tb_hash32(uint64_t key)
{
    uint32_t lo = (uint32_t)key, hi = (uint32_t)(key >> 32);
    uint32_t h = lo ^ (hi * 0x9E3779B1u);
    h ^= h >> 16;
    h *= 0x7FEB352Du;
    h ^= h >> 15;
    h *= 0x846CA68Bu;
    h ^= h >> 16;
    return h;
}

__device__ __forceinline__ uint32_t tb_bucket_of(uint64_t key, uint32_t n_buckets)
{
    return __umulhi(tb_hash32(key), n_buckets);
}

// One 256-bit read-only load of a bucket.  L1::no_allocate: a probed sector is
// never reused by this SM, keep L1 for the read stream.  L2::64B: by default an L2
// miss on B200 fetches the whole 128-B line from HBM (measured 125 B of DRAM traffic
// per random 32-B request, profiles/r1_gather_cache_qualifiers.txt); the 64-B
// granularity is the smallest the ISA offers and halves that traffic.
__device__ __forceinline__ tb_bucket tb_load_bucket(const tb_bucket *p)
{
    tb_bucket b;
    asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(b.s[0]), "=l"(b.s[1]), "=l"(b.s[2]), "=l"(b.s[3])
                 : "l"(p));
    return b;
}

__device__ __forceinline__ bool tb_bucket_has(const tb_bucket &b, uint64_t key)
{
    return (b.s[0] == key) | (b.s[1] == key) | (b.s[2] == key) | (b.s[3] == key);
}

__device__ __forceinline__ bool tb_bucket_full(const tb_bucket &b)
{
    return b.s[TB_SLOTS - 1] != TB_EMPTY;
}

// Continue a lookup whose home bucket was full and did not hold the key.
static __device__ __noinline__ bool tb_probe_overflow(const tb_set_view &s, uint64_t key, uint32_t bucket)
{
    for (uint32_t n = 1; n < s.n_buckets; n++) {
        bucket = bucket + 1 == s.n_buckets ? 0 : bucket + 1;
        tb_bucket b = tb_load_bucket(s.table + bucket);
        if (tb_bucket_has(b, key)) return true;
        if (!tb_bucket_full(b)) return false;
    }
    return false;
}

// Full scalar lookup (used by the fix-up path, tests and kmer_in_hash_set).
__device__ __forceinline__ bool tb_contains(const tb_set_view &s, uint64_t key)
{
    if (key == TB_EMPTY) return s.has_all_ones != 0;
    uint32_t bucket = tb_bucket_of(key, s.n_buckets);
    tb_bucket b = tb_load_bucket(s.table + bucket);
    if (tb_bucket_has(b, key)) return true;
    if (!tb_bucket_full(b)) return false;
    return tb_probe_overflow(s, key, bucket);
}

// ---- pair table ----------------------------------------------------------------------
// One table for BOTH sets: slot = (tags << 62) | key with tags bit 0 = "in A", bit 1 =
// "in B" (k <= 31, so a key needs at most 62 bits); empty slot = 0.  One bucket load
// answers "A? else B?" (quirk Q4) for a window instead of one load per set.  Buckets and
// the multi-bit pre-filter are addressed with a 64-bit mixer.
#define TB_PAIR_KEYMASK 0x3FFFFFFFFFFFFFFFull

struct tb_pair_view {
    const tb_bucket *table;
    uint32_t n_buckets;
};

__host__ __device__ __forceinline__ uint64_t tb_hash64(uint64_t x)
{
    x *= 0x9E3779B97F4A7C15ull;
    x ^= x >> 32;
    x *= 0xD6E8FEB86659FD93ull;
    x ^= x >> 32;
    return x;
}

__device__ __forceinline__ uint32_t tb_pair_bucket_of(uint64_t h64, uint32_t n_buckets)
{
    return __umulhi((uint32_t)h64, n_buckets);
}

// tags of `key` in an already loaded bucket (0 = not here)
__device__ __forceinline__ uint32_t tb_pair_tags(const tb_bucket &b, uint64_t key)
{
    uint32_t tags = 0;
#pragma unroll
    for (int i = 0; i < TB_SLOTS; i++) {
        const uint64_t t = b.s[i] ^ key;  // same key <=> only the two tag bits differ
        if ((t << 2) == 0) tags |= (uint32_t)(t >> 62);
    }
    return tags;
}

__device__ __forceinline__ bool tb_pair_bucket_full(const tb_bucket &b) { return b.s[TB_SLOTS - 1] != 0; }

static __device__ __noinline__ uint32_t tb_pair_overflow(const tb_pair_view &s, uint64_t key, uint32_t bucket)
{
    for (uint32_t n = 1; n < s.n_buckets; n++) {
        bucket = bucket + 1 == s.n_buckets ? 0 : bucket + 1;
        const tb_bucket b = tb_load_bucket(s.table + bucket);
        const uint32_t tags = tb_pair_tags(b, key);
        if (tags) return tags;
        if (!tb_pair_bucket_full(b)) return 0;
    }
    return 0;
}

// ---- base codes -----------------------------------------------------------
// A0 C1 G2 T3 (c/kmers.c:56-68); only the four upper-case letters are bases.
__host__ __device__ __forceinline__ bool tb_base_valid(unsigned char c)
{
    return c == 'A' || c == 'C' || c == 'G' || c == 'T';
}

__host__ __device__ __forceinline__ uint32_t tb_base_code(unsigned char c)
{
    // bits 1 and 2 of the ASCII code: A 00, C 01, G 11, T 10 -> xor with bit 2
    return tb_base_valid(c) ? (((uint32_t)c >> 1) ^ ((uint32_t)c >> 2)) & 3u : 0u;
}

// Reverse the 32 2-bit digits of x and complement them (3-d): the reverse
// complement of a 32-base word.
__device__ __forceinline__ uint64_t tb_revcomp64(uint64_t x)
{
    x = ~x;
    uint32_t lo = __brev((uint32_t)x), hi = __brev((uint32_t)(x >> 32));
    // brev reversed single bits; swap the two bits of every digit back
    lo = ((lo & 0x55555555u) << 1) | ((lo >> 1) & 0x55555555u);
    hi = ((hi & 0x55555555u) << 1) | ((hi >> 1) & 0x55555555u);
    return ((uint64_t)lo << 32) | hi;
}
