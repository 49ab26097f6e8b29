// capi_count.cu -- C ABI of the k-mer counter (include/trio_kmers_b200.h, "find-unique-kmers").
// Host-side orchestration only: buffers, copies, launch order; the arithmetic is in
// kmer_count.cu and classify.cu (k_pack).  No CPU fallback.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/trio_kmers_b200.h"
#include "launchers.h"

extern "C" void tb_internal_set_error(const char *message);
extern "C" int tb_internal_use_device(int device);

namespace {

int cfail(const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    tb_internal_set_error(buf);
    return -1;
}

#define CKC(call)                                                                          \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess)                                                             \
            return cfail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                         __LINE__);                                                        \
    } while (0)

}  // namespace

struct tb_kmer_counter {
    int k = 0, device = 0;
    tb_count_view T;
    unsigned long long *d_stats = nullptr;  // windows counted, distinct keys, overflow flag
    cudaStream_t stream = nullptr;
    cudaEvent_t copied = nullptr;           // the batch's host buffers have been read
    // batch buffers (grown on demand)
    unsigned char *d_ascii = nullptr;
    int64_t *d_offsets = nullptr;
    uint64_t *d_F = nullptr;
    uint32_t *d_V = nullptr, *d_dirty = nullptr;
    int64_t cap_bases = 0, cap_reads = 0;
};

namespace {

void free_batch_buffers(tb_kmer_counter *c)
{
    void *ptrs[] = { c->d_ascii, c->d_offsets, c->d_F, c->d_V, c->d_dirty };
    for (void *p : ptrs)
        if (p) cudaFree(p);
    c->d_ascii = nullptr;
    c->d_offsets = nullptr;
    c->d_F = nullptr;
    c->d_V = nullptr;
    c->d_dirty = nullptr;
    c->cap_bases = c->cap_reads = 0;
}

int ensure_batch_buffers(tb_kmer_counter *c, int64_t total, int64_t n_reads)
{
    if (total <= c->cap_bases && n_reads <= c->cap_reads) return 0;
    CKC(cudaStreamSynchronize(c->stream));
    const int64_t nb = std::max(total + total / 8, c->cap_bases), nr = std::max(n_reads + n_reads / 8 + 64, c->cap_reads);
    free_batch_buffers(c);
    const int64_t tiles = (nb + 1023) / 1024;
    const int64_t words = (tiles + 9) * 32 + 2;
    CKC(cudaMalloc((void **)&c->d_ascii, (size_t)(tiles * 1024 + 64)));
    CKC(cudaMalloc((void **)&c->d_offsets, (size_t)(nr + 2) * 8));
    CKC(cudaMalloc((void **)&c->d_F, (size_t)words * 8));
    CKC(cudaMalloc((void **)&c->d_V, (size_t)words * 4));
    CKC(cudaMalloc((void **)&c->d_dirty, (size_t)(tiles + 10) * 4));
    c->cap_bases = nb;
    c->cap_reads = nr;
    return 0;
}

}  // namespace

extern "C" {

int tb_counter_slot_bytes(int k) { return k <= TB_COUNT_PACKED_MAX_K ? 8 : 12; }

tb_kmer_counter *tb_counter_create(int k, uint64_t n_slots, int part, int n_parts, int device)
{
    if (k < 1 || k > 32) { cfail("k=%d outside 1..32", k); return nullptr; }
    if (n_slots < 1024) n_slots = 1024;
    if (n_parts < 1 || part < 0 || part >= n_parts) { cfail("bad partition %d of %d", part, n_parts); return nullptr; }
    if (tb_internal_use_device(device)) return nullptr;
    tb_kmer_counter *c = new tb_kmer_counter();
    c->k = k;
    c->device = device;
    memset(&c->T, 0, sizeof c->T);
    c->T.n_slots = n_slots;
    c->T.part = (uint32_t)part;
    c->T.n_parts = (uint32_t)n_parts;
    const bool packed = k <= TB_COUNT_PACKED_MAX_K;
    c->T.count_shift = packed ? 2u * (uint32_t)k : 0u;
    bool ok = cudaMalloc((void **)&c->T.keys, n_slots * 8) == cudaSuccess &&
              (packed || cudaMalloc((void **)&c->T.counts, n_slots * 4) == cudaSuccess) &&
              cudaMalloc((void **)&c->d_stats, 4 * 8) == cudaSuccess &&
              cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreateWithFlags(&c->copied, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaMemset(c->T.keys, 0xFF, n_slots * 8) == cudaSuccess &&
         (packed || cudaMemset(c->T.counts, 0, n_slots * 4) == cudaSuccess) &&
         cudaMemset(c->d_stats, 0, 4 * 8) == cudaSuccess && cudaDeviceSynchronize() == cudaSuccess;
    if (!ok) {
        cfail("k-mer counter: cannot allocate a table of %llu slots (%.1f GB): %s", (unsigned long long)n_slots,
              (double)n_slots * tb_counter_slot_bytes(k) / 1e9, cudaGetErrorString(cudaGetLastError()));
        tb_counter_destroy(c);
        return nullptr;
    }
    return c;
}

void tb_counter_destroy(tb_kmer_counter *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    free_batch_buffers(c);
    if (c->T.keys) cudaFree(c->T.keys);
    if (c->T.counts) cudaFree(c->T.counts);
    if (c->d_stats) cudaFree(c->d_stats);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->copied) cudaEventDestroy(c->copied);
    delete c;
}

int tb_counter_add_batch(tb_kmer_counter *c, const char *bases, const int64_t *offsets, int64_t n_reads)
{
    if (!c) return cfail("null counter");
    if (n_reads < 0) return cfail("n_reads < 0");
    if (n_reads == 0) return 0;
    if (offsets[0] != 0) return cfail("offsets[0] must be 0");
    const int64_t total = offsets[n_reads];
    for (int64_t r = 0; r < n_reads; r++)
        if (offsets[r + 1] < offsets[r]) return cfail("offsets must be non-decreasing (read %lld)", (long long)r);
    if (total == 0) return 0;
    CKC(cudaSetDevice(c->device));
    if (ensure_batch_buffers(c, total, n_reads)) return -1;
    // The caller may reuse its buffers when this returns: pageable sources are staged by the runtime
    // before cudaMemcpyAsync returns, page-locked ones are read by the copy engine later, so those
    // are waited for (the copy only).  The kernels overlap with the parsing of the next batch.
    CKC(cudaMemcpyAsync(c->d_ascii, bases, (size_t)total, cudaMemcpyHostToDevice, c->stream));
    CKC(cudaMemcpyAsync(c->d_offsets, offsets, (size_t)(n_reads + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    CKC(cudaEventRecord(c->copied, c->stream));
    tb_batch_view v;
    memset(&v, 0, sizeof v);
    v.ascii = c->d_ascii;
    v.offsets = c->d_offsets;
    v.n_reads = n_reads;
    v.total_bases = total;
    v.n_tiles = (total + 1023) / 1024;
    v.F = c->d_F;
    v.V = c->d_V;
    v.dirty = c->d_dirty;
    v.fold_case = 1;
    CKC(tb_launch_pack(v, c->k, c->stream));
    CKC(tb_launch_count_windows(c->d_F, c->d_V, c->d_offsets, n_reads, total, c->k, c->T, c->d_stats, c->stream));
    CKC(cudaEventSynchronize(c->copied));
    return 0;
}

int tb_counter_stats(tb_kmer_counter *c, uint64_t *n_windows, uint64_t *n_distinct, uint64_t *n_slots)
{
    if (!c) return cfail("null counter");
    CKC(cudaSetDevice(c->device));
    CKC(cudaStreamSynchronize(c->stream));
    unsigned long long st[4];
    CKC(cudaMemcpy(st, c->d_stats, sizeof st, cudaMemcpyDeviceToHost));
    if (n_windows) *n_windows = st[0];
    if (n_distinct) *n_distinct = st[1];
    if (n_slots) *n_slots = c->T.n_slots;
    if (st[2]) return cfail("k-mer table full (%llu slots): count with more partitions or a larger table",
                            (unsigned long long)c->T.n_slots);
    return 0;
}

int tb_counter_histogram(tb_kmer_counter *c, uint64_t *hist, int n_bins)
{
    if (!c || !hist) return cfail("null counter / histogram");
    if (n_bins < 2 || n_bins > 1024) return cfail("n_bins must be 2..1024");
    CKC(cudaSetDevice(c->device));
    unsigned long long *d = nullptr;
    CKC(cudaMalloc((void **)&d, (size_t)n_bins * 8));
    cudaError_t e = cudaMemsetAsync(d, 0, (size_t)n_bins * 8, c->stream);
    if (e == cudaSuccess) e = tb_launch_count_histogram(c->T, d, n_bins, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(hist, d, (size_t)n_bins * 8, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    if (e != cudaSuccess) return cfail("k-mer histogram failed: %s", cudaGetErrorString(e));
    return 0;
}

int64_t tb_counter_select_unique(tb_kmer_counter *a, tb_kmer_counter *b, uint32_t min_count, uint32_t max_count,
                                 uint32_t db_min, uint32_t counter_max, uint64_t *out_keys, uint64_t cap)
{
    if (!a) return cfail("null counter");
    if (b && (b->device != a->device || b->k != a->k)) return cfail("counters differ in device or k");
    if (cudaSetDevice(a->device) != cudaSuccess) return cfail("cudaSetDevice failed");
    if (cudaStreamSynchronize(a->stream) != cudaSuccess || (b && cudaStreamSynchronize(b->stream) != cudaSuccess))
        return cfail("k-mer counting failed: %s", cudaGetErrorString(cudaGetLastError()));
    unsigned long long *d_out = nullptr, *d_n = nullptr;
    cudaError_t e = cudaMalloc((void **)&d_n, 8);
    if (e == cudaSuccess && cap) e = cudaMalloc((void **)&d_out, cap * 8);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_n, 0, 8, a->stream);
    tb_count_view none;
    memset(&none, 0, sizeof none);
    if (e == cudaSuccess)
        e = tb_launch_count_select(a->T, b ? b->T : none, min_count, max_count, db_min, counter_max, d_out, cap, d_n, a->stream);
    unsigned long long n = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&n, d_n, 8, cudaMemcpyDeviceToHost, a->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(a->stream);
    if (e == cudaSuccess && cap && out_keys)
        e = cudaMemcpy(out_keys, d_out, (size_t)std::min<unsigned long long>(n, cap) * 8, cudaMemcpyDeviceToHost);
    if (d_out) cudaFree(d_out);
    if (d_n) cudaFree(d_n);
    if (e != cudaSuccess) return cfail("k-mer selection failed: %s", cudaGetErrorString(e));
    return (int64_t)n;
}

}  // extern "C"
