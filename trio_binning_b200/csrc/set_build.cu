// set_build.cu -- k-mer set construction on sm_100a.
//
// Replaces the reference's add_to_hash loop (c/kmers.c:204-221 + :112-122) and
// the per-line kmer_to_int (c/kmers.c:50-72): one thread per line of the k-mer
// text (or per pre-encoded key) encodes the first k bytes LSB-first and inserts
// the RAW key (quirk Q1: never canonicalised) into the bucketed table with
// atomicCAS.  Duplicates collapse (membership is unchanged; num_kmers stays the
// line count and is kept on the host).
#include "launchers.h"

namespace {

// Insert `key`; returns true if this call created a new slot.
__device__ __forceinline__ bool insert_key(tb_bucket *table, uint32_t n_buckets, uint64_t key)
{
    uint32_t bucket = tb_bucket_of(key, n_buckets);
    for (uint32_t n = 0; n < n_buckets; n++) {
        unsigned long long *slots = table[bucket].s;
#pragma unroll
        for (int i = 0; i < TB_SLOTS; i++) {
            unsigned long long cur = slots[i];
            if (cur == key) return false;
            if (cur == TB_EMPTY) {
                unsigned long long old = atomicCAS(&slots[i], TB_EMPTY, (unsigned long long)key);
                if (old == TB_EMPTY) return true;
                if (old == key) return false;
                // lost the race to a different key: this slot is now occupied, go on
            }
        }
        bucket = bucket + 1 == n_buckets ? 0 : bucket + 1;
    }
    return false;  // table full: cannot happen, slots >= 2 * keys
}

__device__ __forceinline__ void account(bool is_new, bool all_ones, bool malformed,
                                        unsigned long long *counters)
{
    unsigned m_new = __ballot_sync(0xFFFFFFFFu, is_new);
    unsigned m_ones = __ballot_sync(0xFFFFFFFFu, all_ones);
    unsigned m_bad = __ballot_sync(0xFFFFFFFFu, malformed);
    if ((threadIdx.x & 31) == 0) {
        if (m_new) atomicAdd(&counters[0], (unsigned long long)__popc(m_new));
        if (m_ones) atomicOr(&counters[1], 1ull);
        if (m_bad) atomicOr(&counters[2], 1ull);
    }
}

// One thread per line.  Line i starts at i*stride; its first k bytes are the
// k-mer.  With check_newline (well-formed text, stride = k+1) the kernel also
// verifies that byte k is '\n' (or end of file) and that no '\n' hides inside
// the k-mer; any violation flags the file as ragged and the host re-does it
// through the line-buffer emulation (quirk Q5).
__global__ void __launch_bounds__(256)
k_insert_text(const char *__restrict__ text, uint64_t n_lines, uint64_t text_bytes, int stride,
              int k, int check_newline, tb_bucket *table, uint32_t n_buckets,
              unsigned long long *counters)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool is_new = false, all_ones = false, bad = false;
    if (i < n_lines) {
        const char *line = text + i * (uint64_t)stride;
        uint64_t avail = text_bytes - i * (uint64_t)stride;
        uint64_t key = 0;
        for (int j = 0; j < k; j++) {
            unsigned char c = (uint64_t)j < avail ? (unsigned char)line[j] : (unsigned char)'\n';
            if (c == '\n') bad = true;
            key |= (uint64_t)tb_base_code(c) << (2 * j);
        }
        if (check_newline) {
            if ((uint64_t)k < avail && line[k] != '\n') bad = true;
        } else {
            bad = false;
        }
        if (key == TB_EMPTY)
            all_ones = true;
        else
            is_new = insert_key(table, n_buckets, key);
    }
    account(is_new, all_ones, bad, counters);
}

__global__ void __launch_bounds__(256)
k_insert_keys(const uint64_t *__restrict__ keys, uint64_t n, uint64_t key_mask, tb_bucket *table,
              uint32_t n_buckets, unsigned long long *counters)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool is_new = false, all_ones = false, out_of_range = false;
    if (i < n) {
        uint64_t key = keys[i];
        if (key & ~key_mask)
            out_of_range = true;  // not the encoding of any k-mer: reported, never stored (the
                                  // pair table keeps its tags in the bits a k-mer leaves free)
        else if (key == TB_EMPTY)
            all_ones = true;
        else
            is_new = insert_key(table, n_buckets, key);
    }
    account(is_new, all_ones, out_of_range, counters);
}

__global__ void __launch_bounds__(256)
k_contains(tb_set_view set, const uint64_t *__restrict__ keys, uint64_t n, unsigned char *out)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = tb_contains(set, keys[i]) ? 1 : 0;
}

// Merge one per-set table into the pair table: every occupied slot of `src` is inserted as
// (tag << 62) | key; a key already present (from the other set, or a racing duplicate) just
// gets the tag OR-ed in.
__global__ void __launch_bounds__(256)
k_pair_insert(const unsigned long long *__restrict__ src, uint64_t n_slots, unsigned long long tag,
              tb_bucket *table, uint32_t n_buckets, unsigned long long *counters)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool is_new = false;
    if (i < n_slots) {
        const unsigned long long key = src[i];
        if (key != TB_EMPTY) {
            const unsigned long long entry = key | (tag << 62);
            uint32_t bucket = tb_pair_bucket_of(tb_hash64(key), n_buckets);
            bool done = false;
            for (uint32_t n = 0; n < n_buckets && !done; n++) {
                unsigned long long *slots = table[bucket].s;
#pragma unroll
                for (int j = 0; j < TB_SLOTS && !done; j++) {
                    unsigned long long cur = slots[j];
                    if (cur == 0) {
                        cur = atomicCAS(&slots[j], 0ull, entry);
                        if (cur == 0) {
                            is_new = true;
                            done = true;
                            break;
                        }
                    }
                    if (((cur ^ key) << 2) == 0) {
                        atomicOr(&slots[j], tag << 62);
                        done = true;
                    }
                }
                bucket = bucket + 1 == n_buckets ? 0 : bucket + 1;
            }
        }
    }
    account(is_new, false, false, counters);
}

inline unsigned blocks_for(uint64_t n, unsigned tpb) { return (unsigned)((n + tpb - 1) / tpb); }

}  // namespace

cudaError_t tb_launch_insert_text(const char *d_text, uint64_t n_lines, uint64_t text_bytes,
                                  int stride, int k, int check_newline, tb_bucket *table,
                                  uint32_t n_buckets, unsigned long long *d_counters,
                                  cudaStream_t stream)
{
    if (n_lines == 0) return cudaSuccess;
    k_insert_text<<<blocks_for(n_lines, 256), 256, 0, stream>>>(
        d_text, n_lines, text_bytes, stride, k, check_newline, table, n_buckets, d_counters);
    return cudaGetLastError();
}

cudaError_t tb_launch_insert_keys(const uint64_t *d_keys, uint64_t n, int k, tb_bucket *table,
                                  uint32_t n_buckets, unsigned long long *d_counters,
                                  cudaStream_t stream)
{
    if (n == 0) return cudaSuccess;
    const uint64_t key_mask = k >= 32 ? ~0ull : ((1ull << (2 * k)) - 1ull);
    k_insert_keys<<<blocks_for(n, 256), 256, 0, stream>>>(d_keys, n, key_mask, table, n_buckets, d_counters);
    return cudaGetLastError();
}

cudaError_t tb_launch_pair_insert(const uint64_t *d_src_table, uint64_t n_slots, int tag,
                                  tb_bucket *pair_table, uint32_t n_buckets,
                                  unsigned long long *d_counters, cudaStream_t stream)
{
    if (n_slots == 0) return cudaSuccess;
    k_pair_insert<<<blocks_for(n_slots, 256), 256, 0, stream>>>(
        reinterpret_cast<const unsigned long long *>(d_src_table), n_slots, (unsigned long long)tag,
        pair_table, n_buckets, d_counters);
    return cudaGetLastError();
}

cudaError_t tb_launch_contains(tb_set_view set, const uint64_t *d_keys, uint64_t n,
                               unsigned char *d_out, cudaStream_t stream)
{
    if (n == 0) return cudaSuccess;
    k_contains<<<blocks_for(n, 256), 256, 0, stream>>>(set, d_keys, n, d_out);
    return cudaGetLastError();
}
