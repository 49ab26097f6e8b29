// capi.cu -- the C ABI declared in include/trio_kmers_b200.h.
//
// Host-side orchestration only: file mapping, device buffers, streams, launch
// order, copies.  All k-mer arithmetic on the path runs in the kernels of
// set_build.cu / classify.cu; there is no CPU fallback -- every entry point
// fails loudly (error code / NULL + message) if CUDA is unavailable.
#include <fcntl.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/trio_kmers_b200.h"
#include "launchers.h"

#define TB_MAGIC 0x30324254u  // 'TB20'

namespace {

thread_local std::string g_err;
std::mutex g_mutex;  // serialises the one-shot / legacy paths and global state
int g_device = 0;
double g_load_factor = 0.5;
bool g_device_ready[64] = { false };

int fail(const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return -1;
}

#define CK(call)                                                                         \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess)                                                           \
            return fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                        __LINE__);                                                       \
    } while (0)

// frees device memory / events when a function leaves through one of CK's early returns
struct DeviceScratch {
    std::vector<void *> mem;
    std::vector<cudaEvent_t> events;
    ~DeviceScratch()
    {
        for (void *p : mem)
            if (p) cudaFree(p);
        for (cudaEvent_t e : events)
            if (e) cudaEventDestroy(e);
    }
};

int env_int(const char *name, int dflt)
{
    const char *v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

// Select `device` and apply per-device settings once.
int use_device(int device)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail("no CUDA device available (%s); this library has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= n) return fail("device %d out of range (0..%d)", device, n - 1);
    CK(cudaSetDevice(device));
    if (device < 64 && !g_device_ready[device]) {
        // A random probe needs one 32-byte sector; do not let L2 fetch 64/128 B for it.
        int gran = env_int("TB_L2_FETCH_GRANULARITY", 32);
        if (gran > 0) {
            cudaError_t ge = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)gran);
            size_t got = 0;
            cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
            if (getenv("TB_VERBOSE"))
                fprintf(stderr, "[tb] cudaLimitMaxL2FetchGranularity: asked %d -> %s, now %zu\n", gran,
                        cudaGetErrorString(ge), got);
        }
        // room for evict_last lines (the bitmap); harmless if the hint works without it
        int persist_mb = env_int("TB_L2_PERSIST_MB", 0);  // measured: a carve-out does not help
        if (persist_mb != 0) {
            cudaDeviceProp prop;
            if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) {
                size_t want = persist_mb > 0 ? (size_t)persist_mb << 20 : (size_t)prop.persistingL2CacheMaxSize;
                cudaError_t pe = cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
                if (getenv("TB_VERBOSE"))
                    fprintf(stderr, "[tb] cudaLimitPersistingL2CacheSize %zu -> %s\n", want, cudaGetErrorString(pe));
            }
        }
        cudaGetLastError();
        tb_classify_config(env_int("TB_UNROLL", 4), env_int("TB_TPB", 128), env_int("TB_FILTER_UNROLL", 16),
                            env_int("TB_FILTER_MODE", 2));
        g_device_ready[device] = true;
    }
    return 0;
}

// m of the minimizer scheme (same rule as tb_minimizer_m in classify.cu; checked by a test that
// compares bitmap sizes, and by every parity test: build and look-up share the device function)
#ifndef TB_MINI_W
#define TB_MINI_W 7
#endif
int tb_minimizer_m_host(int k)
{
    if (k >= 22) return 16;
    int m = k - (TB_MINI_W - 1);
    const int floor13 = k - 2 < 13 ? k - 2 : 13;
    if (m < floor13) m = floor13;
    const int half = (k + 1) / 2;
    return m > half ? m : half;
}

tb_set_view view_of(const hash_set *s)
{
    tb_set_view v;
    v.table = reinterpret_cast<const tb_bucket *>(s->d_table);
    v.n_buckets = s->n_buckets;
    v.has_all_ones = s->has_all_ones;
    return v;
}

bool valid_set(const hash_set *s) { return s && s->magic == TB_MAGIC && s->d_table; }

// ---- set construction -------------------------------------------------------

hash_set *new_set_shell(uint64_t n_lines, int k, int device)
{
    hash_set *s = (hash_set *)calloc(1, sizeof(hash_set));
    s->kmers = NULL;
    s->full = NULL;
    s->hash_size = (int)(n_lines * 4 / 3);  // c/kmers.c:166
    s->k = (unsigned char)k;
    s->num_kmers = (int)n_lines;
    s->magic = TB_MAGIC;
    s->device = device;
    return s;
}

int alloc_table(hash_set *s, uint64_t n_keys)
{
    double want = (double)n_keys / (TB_SLOTS * g_load_factor);
    uint64_t nb = (uint64_t)want + 1;
    if (nb > 0xFFFFFFF0ull) return fail("k-mer set too large for 32-bit bucket index");
    s->n_buckets = (uint32_t)nb;
    CK(cudaMalloc((void **)&s->d_table, nb * sizeof(tb_bucket)));
    CK(cudaMemset(s->d_table, 0xFF, nb * sizeof(tb_bucket)));
    return 0;
}

struct Counters {
    unsigned long long *d = nullptr;
    int init()
    {
        CK(cudaMalloc((void **)&d, 4 * sizeof(unsigned long long)));
        CK(cudaMemset(d, 0, 4 * sizeof(unsigned long long)));
        return 0;
    }
    int read(unsigned long long out[4])
    {
        CK(cudaMemcpy(out, d, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        return 0;
    }
    ~Counters() { if (d) cudaFree(d); }
};

// Upload `bytes` of host text in line-aligned chunks and insert every line.
int insert_text(hash_set *s, const char *text, uint64_t bytes, uint64_t n_lines, int stride,
                int check_newline, Counters &cnt)
{
    const uint64_t lines_per_chunk = std::max<uint64_t>(1, (512ull << 20) / (uint64_t)stride);
    const uint64_t chunk_bytes = std::min<uint64_t>(bytes, lines_per_chunk * (uint64_t)stride);
    char *d_text[2] = { nullptr, nullptr };
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    int rc = 0;
    for (int i = 0; i < 2 && !rc; i++)
        if (cudaMalloc((void **)&d_text[i], chunk_bytes + 64) != cudaSuccess)
            rc = fail("cudaMalloc of %llu-byte text staging failed", (unsigned long long)chunk_bytes);
    cudaEvent_t done[2];
    for (int i = 0; i < 2; i++) cudaEventCreate(&done[i]);
    uint64_t line = 0;
    int buf = 0;
    while (!rc && line < n_lines) {
        uint64_t nl = std::min(lines_per_chunk, n_lines - line);
        uint64_t off = line * (uint64_t)stride;
        uint64_t nbytes = std::min(bytes - off, nl * (uint64_t)stride);
        cudaEventSynchronize(done[buf]);  // staging buffer free again?
        cudaError_t e = cudaMemcpyAsync(d_text[buf], text + off, nbytes, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess)
            e = tb_launch_insert_text(d_text[buf], nl, nbytes, stride, s->k, check_newline,
                                      reinterpret_cast<tb_bucket *>(s->d_table), s->n_buckets,
                                      cnt.d, st);
        if (e != cudaSuccess) rc = fail("k-mer text upload/insert failed: %s", cudaGetErrorString(e));
        cudaEventRecord(done[buf], st);
        line += nl;
        buf ^= 1;
    }
    cudaError_t e = cudaStreamSynchronize(st);
    if (!rc && e != cudaSuccess) rc = fail("insert kernel failed: %s", cudaGetErrorString(e));
    for (int i = 0; i < 2; i++) {
        cudaEventDestroy(done[i]);
        if (d_text[i]) cudaFree(d_text[i]);
    }
    cudaStreamDestroy(st);
    return rc;
}

// Quirk Q5 (c/kmers.c:204-206): the reference feeds add_to_hash the getline
// buffer, which is reused between lines, so a line shorter than k exposes its
// '\n', its NUL and stale bytes of earlier lines.  For files that are not
// "k bytes + newline" per line, replay that buffer on the host (text
// normalisation only) into fixed k-byte records; encoding and insertion still
// happen on the device.  The buffer starts zeroed (the reference's is
// uninitialised).
void normalise_ragged(const char *text, uint64_t bytes, int k, std::vector<char> &records)
{
    std::vector<char> buf(1 << 16, 0);
    for (uint64_t p = 0; p < bytes;) {
        const char *nl = (const char *)memchr(text + p, '\n', bytes - p);
        uint64_t len = nl ? (uint64_t)(nl - (text + p)) + 1 : bytes - p;
        if (len + 1 > buf.size()) buf.resize(std::max<uint64_t>(len + 1, buf.size() * 2), 0);
        memcpy(buf.data(), text + p, len);
        buf[len] = '\0';
        records.insert(records.end(), buf.data(), buf.data() + k);
        p += len;
    }
}

hash_set *set_from_text(const char *text, uint64_t bytes, const char *label, int device)
{
    if (use_device(device)) return NULL;
    if (bytes == 0) { fail("k-mer file %s is empty", label); return NULL; }
    const char *nl = (const char *)memchr(text, '\n', bytes);
    uint64_t first_len = nl ? (uint64_t)(nl - text) + 1 : bytes;
    int64_t k = (int64_t)first_len - 1;  // c/kmers.c:132-133
    if (k < 1 || k > 32) {
        fail("k-mer length %lld in %s is outside 1..32 (first line decides k)", (long long)k, label);
        return NULL;
    }
    const int stride = (int)k + 1;
    uint64_t n_lines = (bytes + stride - 1) / stride;  // if well-formed
    if (n_lines >= (1ull << 31)) { fail("%s: more than 2^31-1 lines", label); return NULL; }

    cudaEvent_t ev0, ev1;
    cudaEventCreate(&ev0);
    cudaEventCreate(&ev1);
    cudaEventRecord(ev0, 0);

    hash_set *s = new_set_shell(n_lines, (int)k, device);
    Counters cnt;
    unsigned long long c[4] = { 0, 0, 0, 0 };
    bool ok = !cnt.init() && !alloc_table(s, n_lines) &&
              !insert_text(s, text, bytes, n_lines, stride, 1, cnt) && !cnt.read(c);
    if (ok && c[2]) {
        // ragged file: redo through the line-buffer emulation
        cudaFree(s->d_table);
        s->d_table = NULL;
        std::vector<char> records;
        normalise_ragged(text, bytes, (int)k, records);
        n_lines = records.size() / (uint64_t)k;
        if (n_lines >= (1ull << 31)) { fail("%s: more than 2^31-1 lines", label); ok = false; }
        hash_set *t = new_set_shell(n_lines, (int)k, device);
        free(s);
        s = t;
        ok = ok && !cudaMemset(cnt.d, 0, 4 * sizeof(unsigned long long)) && !alloc_table(s, n_lines) &&
             !insert_text(s, records.data(), records.size(), n_lines, (int)k, 0, cnt) && !cnt.read(c);
    }
    cudaEventRecord(ev1, 0);
    cudaEventSynchronize(ev1);
    float ms = 0;
    cudaEventElapsedTime(&ms, ev0, ev1);
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    if (!ok) {
        if (s->d_table) cudaFree(s->d_table);
        free(s);
        return NULL;
    }
    s->n_distinct = c[0];
    s->has_all_ones = c[1] ? 1u : 0u;
    s->build_ms = ms;
    return s;
}

// ---- classifier ---------------------------------------------------------------

struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[2] = { nullptr, nullptr };
    cudaEvent_t mark[2] = { nullptr, nullptr };  // user marks (tb_classifier_mark)
    unsigned char *d_ascii = nullptr;
    int64_t *d_offsets = nullptr;
    uint64_t *d_F = nullptr, *d_E = nullptr;
    uint32_t *d_dirty = nullptr, *d_tile_first = nullptr, *d_flags = nullptr, *d_skip = nullptr,
             *d_ub_list = nullptr;
    unsigned int *d_ub_count = nullptr;
    unsigned long long *d_n_pass = nullptr;
    // rare-path arrays of host-packed batches
    uint32_t *d_dirty_idx = nullptr, *d_ub_read = nullptr, *d_ub_skip = nullptr;
    uint64_t *d_e_words = nullptr;
    int64_t *d_ub_off = nullptr;
    unsigned char *d_ub_bytes = nullptr;
    int64_t cap_dirty = 0, cap_ub = 0, cap_ub_bytes = 0;
    int32_t *d_ca = nullptr, *d_cb = nullptr;
    double *d_sa = nullptr, *d_sb = nullptr;
    char *d_bin = nullptr;
    tb_batch_view view;
    bool has_batch = false;
    bool packed_resident = false;  // the slot's last batch arrived host-packed (no ASCII on the device)
    int64_t packed_n_dirty = 0, packed_n_ub = 0;
    uint64_t windows = 0;
    int32_t *h_ca = nullptr;  // where the last batch's count_A went (for stats)
    uint64_t fixup_reads = 0;
};

}  // namespace

struct tb_classifier {
    const hash_set *A, *B;
    int device;
    int k;
    double scale_a, scale_b;
    int64_t max_bases, max_reads;
    std::vector<Slot> slots;
    // L2-resident pre-filter over the keys of A and B (nullptr = direct kernel)
    uint32_t *d_filter = nullptr;
    uint32_t filter_mask = 0;
    int filter_log2 = 0;
    float filter_build_ms = 0;
    // pair table (both sets, tagged; k <= 31): phase 2 needs one HBM bucket per window
    tb_bucket *d_pair = nullptr;
    uint32_t pair_buckets = 0;
    uint64_t pair_distinct = 0;
    int filter_bits = 1;
    uint32_t filter_word_mask = 0;  // (bits / 32) - 1
    bool mini = false;              // bitmap sectors addressed by the key's minimizer
    int hier = 0;                   // ... and its 128-byte lines by a coarser (13-mer) minimizer
    int p2_width = 2;               // table probes per lane and phase-2 round (minimizer kernel)
};

namespace {

void free_slot(Slot &s)
{
    void *ptrs[] = { s.d_ascii, s.d_offsets, s.d_F, s.d_E, s.d_dirty, s.d_tile_first, s.d_flags,
                     s.d_skip, s.d_ub_list, s.d_ub_count, s.d_n_pass, s.d_dirty_idx, s.d_ub_read, s.d_ub_skip,
                     s.d_e_words, s.d_ub_off, s.d_ub_bytes, s.d_ca, s.d_cb, s.d_sa, s.d_sb, s.d_bin };
    for (void *p : ptrs)
        if (p) cudaFree(p);
    for (int i = 0; i < 2; i++) {
        if (s.ev[i]) cudaEventDestroy(s.ev[i]);
        if (s.mark[i]) cudaEventDestroy(s.mark[i]);
    }
    if (s.stream) cudaStreamDestroy(s.stream);
    s = Slot();
}

int alloc_slot(Slot &s, int64_t max_bases, int64_t max_reads)
{
    const int64_t tiles = (max_bases + 1023) / 1024;
    const int64_t words = (tiles + 9) * 32 + 2;
    CK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&s.ev[0]));
    CK(cudaEventCreate(&s.ev[1]));
    CK(cudaEventCreate(&s.mark[0]));
    CK(cudaEventCreate(&s.mark[1]));
    CK(cudaMalloc((void **)&s.d_ascii, (size_t)(tiles * 1024 + 64)));
    CK(cudaMalloc((void **)&s.d_offsets, (size_t)(max_reads + 2) * sizeof(int64_t)));
    CK(cudaMalloc((void **)&s.d_F, (size_t)words * 8));
    CK(cudaMalloc((void **)&s.d_E, (size_t)words * 8));
    CK(cudaMalloc((void **)&s.d_dirty, (size_t)(tiles + 10) * 4));
    CK(cudaMalloc((void **)&s.d_tile_first, (size_t)(tiles + 1) * 4));
    CK(cudaMalloc((void **)&s.d_flags, (size_t)(max_reads + 1) * 4));
    CK(cudaMalloc((void **)&s.d_skip, (size_t)(max_reads + 1) * 4));
    CK(cudaMalloc((void **)&s.d_ub_list, (size_t)(max_reads + 1) * 4));
    CK(cudaMalloc((void **)&s.d_ub_count, 2 * sizeof(unsigned int)));
    CK(cudaMalloc((void **)&s.d_n_pass, sizeof(unsigned long long)));
    CK(cudaMemset(s.d_n_pass, 0, sizeof(unsigned long long)));
    CK(cudaMalloc((void **)&s.d_ca, (size_t)(max_reads + 1) * 4));
    CK(cudaMalloc((void **)&s.d_cb, (size_t)(max_reads + 1) * 4));
    CK(cudaMalloc((void **)&s.d_sa, (size_t)(max_reads + 1) * 8));
    CK(cudaMalloc((void **)&s.d_sb, (size_t)(max_reads + 1) * 8));
    CK(cudaMalloc((void **)&s.d_bin, (size_t)(max_reads + 1)));
    s.cap_dirty = words;  // worst case: every word holds a non-ACGT byte
    s.cap_ub = max_reads + 1;
    s.cap_ub_bytes = std::max<int64_t>(1 << 20, max_bases / 16);
    CK(cudaMalloc((void **)&s.d_dirty_idx, (size_t)s.cap_dirty * 4));
    CK(cudaMalloc((void **)&s.d_e_words, (size_t)s.cap_dirty * 8));
    CK(cudaMalloc((void **)&s.d_ub_read, (size_t)s.cap_ub * 4));
    CK(cudaMalloc((void **)&s.d_ub_skip, (size_t)s.cap_ub * 4));
    CK(cudaMalloc((void **)&s.d_ub_off, (size_t)(s.cap_ub + 1) * 8));
    CK(cudaMalloc((void **)&s.d_ub_bytes, (size_t)s.cap_ub_bytes + 64));
    CK(cudaMemset(s.d_flags, 0, (size_t)(max_reads + 1) * 4));
    CK(cudaMemset(s.d_ub_count, 0, 2 * sizeof(unsigned int)));
    CK(cudaMemset(s.d_E, 0, (size_t)words * 8));
    return 0;
}

// enqueue the device part of a batch (5 kernels) on the slot's stream
int enqueue_kernels(tb_classifier *c, Slot &s, cudaEvent_t before_classify, cudaEvent_t after_classify)
{
    const tb_set_view A = view_of(c->A), B = view_of(c->B);
    CK(tb_launch_pack(s.view, c->k, s.stream));
    CK(tb_launch_index(s.view, c->k, s.stream));
    if (before_classify) CK(cudaEventRecord(before_classify, s.stream));
    if (c->d_pair && c->d_filter && c->mini) {
        tb_pair_view P = { c->d_pair, c->pair_buckets };
        CK(tb_launch_classify_pair_mini(s.view, c->k, P, c->d_filter, (c->filter_word_mask >> 3) + 1, c->filter_bits, c->p2_width, c->hier, s.stream));
    } else if (c->d_pair && c->d_filter) {
        tb_pair_view P = { c->d_pair, c->pair_buckets };
        CK(tb_launch_classify_pair(s.view, c->k, P, c->d_filter, c->filter_word_mask, c->filter_bits, s.stream));
    } else if (c->d_filter)
        CK(tb_launch_classify_filtered(s.view, c->k, A, B, c->d_filter, c->filter_mask, s.stream));
    else
        CK(tb_launch_classify(s.view, c->k, A, B, s.stream));
    if (after_classify) CK(cudaEventRecord(after_classify, s.stream));
    CK(tb_launch_fixup(s.view, c->k, A, B, s.stream));
    CK(tb_launch_finalize(s.view, c->scale_a, c->scale_b, s.stream));
    return 0;
}

// the device part of a host-packed batch: index + scatter of the rare words, classify, fix-up, decide
int enqueue_packed_kernels(tb_classifier *c, Slot &s, cudaEvent_t before_classify, cudaEvent_t after_classify)
{
    const tb_set_view A = view_of(c->A), B = view_of(c->B);
    const tb_batch_view &v = s.view;
    cudaStream_t st = s.stream;
    CK(tb_launch_index_packed(v, s.d_dirty_idx, s.d_e_words, s.packed_n_dirty, s.d_ub_read, s.d_ub_skip,
                              s.packed_n_ub, st));
    if (before_classify) CK(cudaEventRecord(before_classify, st));
    if (c->d_pair && c->d_filter && c->mini) {
        tb_pair_view P = { c->d_pair, c->pair_buckets };
        CK(tb_launch_classify_pair_mini(v, c->k, P, c->d_filter, (c->filter_word_mask >> 3) + 1, c->filter_bits, c->p2_width, c->hier, st));
    } else if (c->d_pair && c->d_filter) {
        tb_pair_view P = { c->d_pair, c->pair_buckets };
        CK(tb_launch_classify_pair(v, c->k, P, c->d_filter, c->filter_word_mask, c->filter_bits, st));
    } else if (c->d_filter)
        CK(tb_launch_classify_filtered(v, c->k, A, B, c->d_filter, c->filter_mask, st));
    else
        CK(tb_launch_classify(v, c->k, A, B, st));
    if (after_classify) CK(cudaEventRecord(after_classify, st));
    CK(tb_launch_fixup(v, c->k, A, B, st));
    CK(tb_launch_finalize(v, c->scale_a, c->scale_b, st));
    return 0;
}

// Bitmap pre-filter over (A u B): 2^m bits, m chosen so that the bitmap holds >= 16
// bits per distinct key but never more than 2^29 bits = 64 MB, the size up to which a
// randomly read table stays L2-resident on B200 (profiles/gather_l2curve_r1.txt).
// TB_FILTER_LOG2=0 disables it, another value forces m.
int build_filter(tb_classifier *c)
{
    int m = env_int("TB_FILTER_LOG2", -1);
    if (m == 0) return 0;
    const bool pair_ok = c->k <= 31 && env_int("TB_PAIR", 1) != 0;
    // minimizer-addressed sectors (default whenever the pair table exists): consecutive windows
    // share bitmap sectors, so the bitmap may be far larger than L2 and still cheap to read
    c->mini = pair_ok && env_int("TB_MINIMIZER", 1) != 0;
    {
        const int w = env_int("TB_P2_WIDTH", 0);
        c->p2_width = w >= 4 ? 4 : (w == 1 ? 1 : 2);  // 1 = warp-cooperative phase 2
    }
    if (m < 0) {
        const uint64_t n = c->A->n_distinct + c->B->n_distinct;
        if (c->mini) {
            // Bits per key, measured (profiles/r2_bitmap_size_sweep.txt): a larger bitmap lets fewer
            // non-members through to the table probes, a smaller one keeps more of its sectors in
            // L2.  With few m-mers per window (k <= 23: a sector fetch every ~4 windows) the
            // fetches dominate and ~8-16 bits per key are best (C3 2^32, C5 2^30); with many
            // (k >= 28: a fetch every ~8 windows) the survivors do, and 32+ bits per key win
            // (C4 2^35).  A bitmap that fits L2 anyway (<= 2^29 bits = 64 MB) may as well be
            // sparse: up to 32 bits per key (C2 2^29).
            const int w = c->k - tb_minimizer_m_host(c->k) + 1;
            const uint64_t per_key = w <= 8 ? 8 : (w <= 12 ? 16 : 32);
            m = 16;
            while (m < 35 && (1ull << m) < per_key * n) m++;
            if (m < 29) {
                m = 16;
                while (m < 29 && (1ull << m) < 32 * n) m++;
            }
        } else {
            // hash-addressed words: keep the bitmap L2-resident (<= 2^29 bits = 64 MB); very
            // large sets do better with 2^30 although it spills
            m = 16;
            while (m < 29 && (1ull << m) < 16 * n) m++;
            if (n >= (1ull << 27)) m = 30;
        }
    }
    if (m < 10) m = 10;
    if (m > (c->mini ? 35 : 32)) m = c->mini ? 35 : 32;
    const uint64_t bits = 1ull << m;
    CK(cudaMalloc((void **)&c->d_filter, bits / 8));
    CK(cudaMemset(c->d_filter, 0, bits / 8));
    c->filter_mask = (uint32_t)(bits - 1);  // only used by the two-table kernel (m <= 32)
    c->filter_word_mask = (uint32_t)((bits >> 5) - 1);
    c->filter_log2 = m;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0, 0));
    const uint64_t slots_a = (uint64_t)c->A->n_buckets * TB_SLOTS, slots_b = (uint64_t)c->B->n_buckets * TB_SLOTS;
    if (pair_ok) {
        // one tagged table for both sets
        const uint64_t n = c->A->n_distinct + c->B->n_distinct;
        const uint64_t nb = (uint64_t)((double)n / (TB_SLOTS * g_load_factor)) + 1;
        if (nb > 0xFFFFFFF0ull) return fail("pair table too large");
        c->pair_buckets = (uint32_t)nb;
        CK(cudaMalloc((void **)&c->d_pair, nb * sizeof(tb_bucket)));
        CK(cudaMemset(c->d_pair, 0, nb * sizeof(tb_bucket)));
        Counters cnt;
        if (cnt.init()) return -1;
        CK(tb_launch_pair_insert(c->A->d_table, slots_a, 1, c->d_pair, c->pair_buckets, cnt.d, 0));
        CK(tb_launch_pair_insert(c->B->d_table, slots_b, 2, c->d_pair, c->pair_buckets, cnt.d, 0));
        unsigned long long cc[4];
        if (cnt.read(cc)) return -1;
        c->pair_distinct = cc[0];
        // two bits per key while the bitmap stays sparse (fewer false positives for free)
        int nbits = env_int("TB_FILTER_BITS", 0);
        if (nbits != 1 && nbits != 2) nbits = ((double)n / (double)bits < 0.4) ? 2 : 1;
        c->filter_bits = nbits;
        // two-level addressing (TB_HIER=1; off by default): where the bitmap is far beyond L2 (>= 2^31
        // bits = 256 MB) and the k-mer has more 13-mers than m-mers (k = 20 .. 24).  Measured at C3 /
        // C5: 18 % fewer DRAM bytes, but 35 % more L1 misses and 27 % / 8-24 % MORE time
        // (profiles/r2_variants_hier.txt) -- kept as an experiment, not as the default
        c->hier = (c->mini && nbits == 2 && c->p2_width == 2 && c->k >= 20 && c->k <= 24 && m >= 31 &&
                   env_int("TB_HIER", 0) != 0) ? 1 : 0;
        if (c->mini)
            CK(tb_launch_pair_filter_build_mini(reinterpret_cast<const uint64_t *>(c->d_pair), nb * TB_SLOTS,
                                                c->d_filter, (c->filter_word_mask >> 3) + 1, nbits, c->k, c->hier, 0));
        else
            CK(tb_launch_pair_filter_build(reinterpret_cast<const uint64_t *>(c->d_pair), nb * TB_SLOTS,
                                           c->d_filter, c->filter_word_mask, nbits, 0));
    } else {
        CK(tb_launch_filter_build(c->A->d_table, slots_a, c->d_filter, c->filter_mask, 0));
        CK(tb_launch_filter_build(c->B->d_table, slots_b, c->d_filter, c->filter_mask, 0));
    }
    CK(cudaEventRecord(e1, 0));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&c->filter_build_ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return 0;
}

// one-shot cache used by tb_classify_batch and the legacy per-read call
tb_classifier *g_default = nullptr;

}  // namespace

extern "C" {

const char *tb_last_error(void) { return g_err.c_str(); }
// for the other translation units of the library (not part of the public header)
void tb_internal_set_error(const char *message) { g_err = message ? message : ""; }
int tb_internal_use_device(int device)
{
    std::lock_guard<std::mutex> lk(g_mutex);
    return use_device(device);
}
const char *tb_version(void) { return "trio_kmers_b200 0.1 (sm_100a)"; }

int tb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int tb_set_device(int device)
{
    std::lock_guard<std::mutex> lk(g_mutex);
    if (use_device(device)) return -1;
    g_device = device;
    return 0;
}

int tb_get_device(void) { return g_device; }

int tb_device_info(int device, char *name_buf, int name_len, int *sm_count, uint64_t *hbm_bytes,
                   uint64_t *l2_bytes)
{
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, device));
    if (name_buf && name_len > 0) {
        strncpy(name_buf, p.name, (size_t)name_len - 1);
        name_buf[name_len - 1] = '\0';
    }
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (hbm_bytes) *hbm_bytes = (uint64_t)p.totalGlobalMem;
    if (l2_bytes) *l2_bytes = (uint64_t)p.l2CacheSize;
    return 0;
}

void *tb_host_alloc(uint64_t bytes)
{
    void *p = nullptr;
    if (use_device(g_device)) return nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        fail("cudaHostAlloc(%llu) failed", (unsigned long long)bytes);
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

void tb_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

int tb_set_load_factor(double lf)
{
    if (!(lf > 0.0) || lf > 0.95) return fail("load factor must be in (0, 0.95]");
    g_load_factor = lf;
    return 0;
}

hash_set *tb_set_create_from_file(const char *path, int device)
{
    int fd = open(path, O_RDONLY);
    if (fd < 0) { fail("cannot open k-mer file %s", path); return NULL; }
    struct stat st;
    if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) {
        close(fd);
        fail("%s is not a regular file", path);
        return NULL;
    }
    uint64_t bytes = (uint64_t)st.st_size;
    if (bytes == 0) { close(fd); fail("k-mer file %s is empty", path); return NULL; }
    void *map = mmap(NULL, bytes, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (map == MAP_FAILED) { fail("mmap of %s failed", path); return NULL; }
    madvise(map, bytes, MADV_SEQUENTIAL);
    hash_set *s = set_from_text((const char *)map, bytes, path, device);
    munmap(map, bytes);
    if (s)
        fprintf(stderr, "Found %d %d-mers in %s (%llu distinct, %u buckets on GPU %d, %.1f ms).\n",
                s->num_kmers, (int)s->k, path, (unsigned long long)s->n_distinct, s->n_buckets,
                s->device, s->build_ms);
    return s;
}

hash_set *tb_set_create_from_keys(const uint64_t *keys, uint64_t n, int k, int device)
{
    if (use_device(device)) return NULL;
    if (k < 1 || k > 32) { fail("k=%d outside 1..32", k); return NULL; }
    if (n == 0 || n >= (1ull << 31)) { fail("number of keys must be in 1..2^31-1"); return NULL; }
    cudaEvent_t ev0, ev1;
    cudaEventCreate(&ev0);
    cudaEventCreate(&ev1);
    cudaEventRecord(ev0, 0);
    hash_set *s = new_set_shell(n, k, device);
    Counters cnt;
    unsigned long long c[4] = { 0, 0, 0, 0 };
    bool ok = !cnt.init() && !alloc_table(s, n);
    const uint64_t chunk = 64ull << 20;  // keys per chunk (512 MB)
    uint64_t *d_keys = nullptr;
    if (ok && cudaMalloc((void **)&d_keys, std::min(n, chunk) * 8) != cudaSuccess) {
        fail("cudaMalloc for key staging failed");
        ok = false;
    }
    for (uint64_t i = 0; ok && i < n; i += chunk) {
        uint64_t m = std::min(chunk, n - i);
        cudaError_t e = cudaMemcpy(d_keys, keys + i, m * 8, cudaMemcpyHostToDevice);
        if (e == cudaSuccess)
            e = tb_launch_insert_keys(d_keys, m, k, reinterpret_cast<tb_bucket *>(s->d_table),
                                      s->n_buckets, cnt.d, 0);
        if (e != cudaSuccess) { fail("key upload/insert failed: %s", cudaGetErrorString(e)); ok = false; }
    }
    if (ok && cudaDeviceSynchronize() != cudaSuccess) { fail("insert kernel failed"); ok = false; }
    ok = ok && !cnt.read(c);
    if (ok && c[2]) {
        fail("a key is not below 4^k (k=%d): not the kmer_to_int value of any %d-mer", k, k);
        ok = false;
    }
    if (d_keys) cudaFree(d_keys);
    cudaEventRecord(ev1, 0);
    cudaEventSynchronize(ev1);
    float ms = 0;
    cudaEventElapsedTime(&ms, ev0, ev1);
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    if (!ok) {
        if (s->d_table) cudaFree(s->d_table);
        free(s);
        return NULL;
    }
    s->n_distinct = c[0];
    s->has_all_ones = c[1] ? 1u : 0u;
    s->build_ms = ms;
    return s;
}

hash_set *tb_set_clone(const hash_set *src, int device)
{
    if (!valid_set(src)) { fail("invalid set handle"); return NULL; }
    std::lock_guard<std::mutex> lk(g_mutex);
    if (use_device(device)) return NULL;
    hash_set *s = (hash_set *)malloc(sizeof(hash_set));
    *s = *src;
    s->device = device;
    s->d_table = NULL;
    const size_t bytes = (size_t)src->n_buckets * sizeof(tb_bucket);
    cudaError_t e = cudaMalloc((void **)&s->d_table, bytes);
    if (e == cudaSuccess) {
        if (device != src->device) {
            int can = 0;
            cudaDeviceCanAccessPeer(&can, device, src->device);
            if (can && cudaDeviceEnablePeerAccess(src->device, 0) == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        }
        e = cudaMemcpyPeer(s->d_table, device, src->d_table, src->device, bytes);
    }
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        fail("cloning a set to device %d failed: %s", device, cudaGetErrorString(e));
        if (s->d_table) cudaFree(s->d_table);
        free(s);
        cudaGetLastError();
        return NULL;
    }
    return s;
}

void tb_set_destroy(hash_set *s)
{
    if (!valid_set(s)) return;
    std::lock_guard<std::mutex> lk(g_mutex);
    if (g_default && (g_default->A == s || g_default->B == s)) {
        tb_classifier_destroy(g_default);
        g_default = nullptr;
    }
    cudaSetDevice(s->device);
    cudaFree(s->d_table);
    s->magic = 0;
    s->d_table = NULL;
    free(s);
}

int tb_set_contains_keys(const hash_set *s, const uint64_t *keys, uint64_t n, unsigned char *out)
{
    if (!valid_set(s)) return fail("invalid set handle");
    if (use_device(s->device)) return -1;
    if (n == 0) return 0;
    uint64_t *d_keys = nullptr;
    unsigned char *d_out = nullptr;
    DeviceScratch scratch;
    CK(cudaMalloc((void **)&d_keys, n * 8));
    scratch.mem.push_back(d_keys);
    CK(cudaMalloc((void **)&d_out, n));
    scratch.mem.push_back(d_out);
    CK(cudaMemcpy(d_keys, keys, n * 8, cudaMemcpyHostToDevice));
    CK(tb_launch_contains(view_of(s), d_keys, n, d_out, 0));
    CK(cudaMemcpy(out, d_out, n, cudaMemcpyDeviceToHost));
    return 0;
}

void tb_scaling_factors(const hash_set *A, const hash_set *B, double *scale_A, double *scale_B)
{
    // classify_by_kmers.py:72-76
    int m = A->num_kmers > B->num_kmers ? A->num_kmers : B->num_kmers;
    *scale_A = 1.0 * (double)m / (double)A->num_kmers;
    *scale_B = 1.0 * (double)m / (double)B->num_kmers;
}

tb_classifier *tb_classifier_create(const hash_set *A, const hash_set *B, int64_t max_bases,
                                    int64_t max_reads, int n_slots)
{
    if (!valid_set(A) || !valid_set(B)) { fail("invalid set handle"); return NULL; }
    if (A->device != B->device) { fail("sets live on different devices"); return NULL; }
    if (A->k != B->k) {
        fail("k differs between the sets (%d vs %d); the reference would misread its buffers here",
             (int)A->k, (int)B->k);
        return NULL;
    }
    if (max_bases < 1) max_bases = 1;
    if (max_reads < 1) max_reads = 1;
    if (max_reads >= (1ll << 32)) { fail("max_reads must be < 2^32"); return NULL; }
    if (n_slots < 1 || n_slots > 16) { fail("n_slots must be 1..16"); return NULL; }
    if (use_device(A->device)) return NULL;
    tb_classifier *c = new tb_classifier();
    c->A = A;
    c->B = B;
    c->device = A->device;
    c->k = A->k;
    tb_scaling_factors(A, B, &c->scale_a, &c->scale_b);
    c->max_bases = max_bases;
    c->max_reads = max_reads;
    c->slots.resize((size_t)n_slots);
    if (build_filter(c)) {
        std::string keep = g_err;
        tb_classifier_destroy(c);
        g_err = keep;
        return NULL;
    }
    for (auto &s : c->slots) {
        if (alloc_slot(s, max_bases, max_reads)) {
            std::string keep = g_err;
            tb_classifier_destroy(c);
            g_err = keep;
            return NULL;
        }
    }
    if (c->d_filter && env_int("TB_L2_WINDOW", 0)) {
        // experiment: pin the bitmap with the driver's persisting-L2 mechanism (set-aside +
        // per-stream access policy window) instead of the per-load evict_last hint
        cudaDeviceProp prop;
        cudaGetDeviceProperties(&prop, c->device);
        const size_t bytes = ((size_t)c->filter_mask + 1) / 8;
        const size_t set_aside = std::min<size_t>(bytes, (size_t)prop.persistingL2CacheMaxSize);
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, set_aside);
        for (auto &s : c->slots) {
            cudaStreamAttrValue attr;
            memset(&attr, 0, sizeof attr);
            attr.accessPolicyWindow.base_ptr = c->d_filter;
            attr.accessPolicyWindow.num_bytes = std::min<size_t>(bytes, (size_t)prop.accessPolicyMaxWindowSize);
            attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)set_aside / (double)bytes);
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            cudaError_t e = cudaStreamSetAttribute(s.stream, cudaStreamAttributeAccessPolicyWindow, &attr);
            if (getenv("TB_VERBOSE"))
                fprintf(stderr, "[tb] L2 window %zu bytes, set-aside %zu: %s\n", bytes, set_aside, cudaGetErrorString(e));
        }
        cudaGetLastError();
    }
    return c;
}

void tb_classifier_destroy(tb_classifier *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    for (auto &s : c->slots) {
        if (s.stream) cudaStreamSynchronize(s.stream);
        free_slot(s);
    }
    if (c->d_filter) cudaFree(c->d_filter);
    if (c->d_pair) cudaFree(c->d_pair);
    delete c;
}

int tb_classifier_submit(tb_classifier *c, int slot, const char *bases, const int64_t *offsets,
                         int64_t n_reads, int32_t *count_A, int32_t *count_B, double *score_A,
                         double *score_B, char *bin)
{
    if (!c) return fail("null classifier");
    if (slot < 0 || slot >= (int)c->slots.size()) return fail("slot %d out of range", slot);
    if (n_reads < 0 || n_reads > c->max_reads)
        return fail("batch has %lld reads, classifier capacity is %lld", (long long)n_reads,
                    (long long)c->max_reads);
    if (!count_A || !count_B) return fail("count_A / count_B must not be NULL");
    CK(cudaSetDevice(c->device));
    Slot &s = c->slots[(size_t)slot];
    const int64_t total = n_reads ? offsets[n_reads] : 0;
    if (n_reads && offsets[0] != 0) return fail("offsets[0] must be 0");
    if (total > c->max_bases)
        return fail("batch has %lld bases, classifier capacity is %lld", (long long)total,
                    (long long)c->max_bases);
    uint64_t windows = 0;
    for (int64_t r = 0; r < n_reads; r++) {
        int64_t len = offsets[r + 1] - offsets[r];
        if (len < 0) return fail("offsets must be non-decreasing (read %lld)", (long long)r);
        if (len >= (1ll << 31)) return fail("read %lld longer than 2^31-1 bases", (long long)r);
        if (len >= c->k) windows += (uint64_t)(len - c->k + 1);
    }
    tb_batch_view &v = s.view;
    v.ascii = s.d_ascii;
    v.offsets = s.d_offsets;
    v.n_reads = n_reads;
    v.total_bases = total;
    v.n_tiles = (total + 1023) / 1024;
    v.F = s.d_F;
    v.E = s.d_E;
    v.dirty = s.d_dirty;
    v.tile_first_read = s.d_tile_first;
    v.flags = s.d_flags;
    v.skip = s.d_skip;
    v.ub_list = s.d_ub_list;
    v.ub_count = s.d_ub_count;
    v.n_pass = s.d_n_pass;
    v.ub_ascii = nullptr;
    v.ub_ascii_off = nullptr;
    v.count_a = s.d_ca;
    v.count_b = s.d_cb;
    v.score_a = score_A ? s.d_sa : nullptr;
    v.score_b = score_B ? s.d_sb : nullptr;
    v.bin = bin ? s.d_bin : nullptr;
    s.windows = windows;
    s.h_ca = count_A;
    s.has_batch = true;
    s.packed_resident = false;
    if (n_reads == 0) return 0;
    if (total) CK(cudaMemcpyAsync(s.d_ascii, bases, (size_t)total, cudaMemcpyHostToDevice, s.stream));
    CK(cudaMemcpyAsync(s.d_offsets, offsets, (size_t)(n_reads + 1) * 8, cudaMemcpyHostToDevice, s.stream));
    if (enqueue_kernels(c, s, nullptr, nullptr)) return -1;
    // number of reads that took the fix-up path (d_ub_count[1] is a sticky copy)
    CK(cudaMemcpyAsync(count_A, s.d_ca, (size_t)n_reads * 4, cudaMemcpyDeviceToHost, s.stream));
    CK(cudaMemcpyAsync(count_B, s.d_cb, (size_t)n_reads * 4, cudaMemcpyDeviceToHost, s.stream));
    if (score_A) CK(cudaMemcpyAsync(score_A, s.d_sa, (size_t)n_reads * 8, cudaMemcpyDeviceToHost, s.stream));
    if (score_B) CK(cudaMemcpyAsync(score_B, s.d_sb, (size_t)n_reads * 8, cudaMemcpyDeviceToHost, s.stream));
    if (bin) CK(cudaMemcpyAsync(bin, s.d_bin, (size_t)n_reads, cudaMemcpyDeviceToHost, s.stream));
    return 0;
}

int tb_classifier_capacity(tb_classifier *c, int64_t *max_bases, int64_t *max_reads, int64_t *cap_words,
                           int64_t *cap_tiles, int64_t *cap_dirty, int64_t *cap_ub, int64_t *cap_ub_bytes)
{
    if (!c) return fail("null classifier");
    const Slot &s = c->slots[0];
    const int64_t tiles = (c->max_bases + 1023) / 1024;
    if (max_bases) *max_bases = c->max_bases;
    if (max_reads) *max_reads = c->max_reads;
    if (cap_words) *cap_words = (tiles + 1) * 32;
    if (cap_tiles) *cap_tiles = tiles + 1;
    if (cap_dirty) *cap_dirty = s.cap_dirty;
    if (cap_ub) *cap_ub = s.cap_ub;
    if (cap_ub_bytes) *cap_ub_bytes = s.cap_ub_bytes;
    return 0;
}

int tb_classifier_submit_packed(tb_classifier *c, int slot, const tb_packed_batch *pb,
                                const int64_t *offsets, int32_t *count_A, int32_t *count_B,
                                double *score_A, double *score_B, char *bin)
{
    if (!c || !pb) return fail("null classifier / batch");
    if (slot < 0 || slot >= (int)c->slots.size()) return fail("slot %d out of range", slot);
    if (!count_A || !count_B) return fail("count_A / count_B must not be NULL");
    const int64_t n_reads = pb->n_reads, total = pb->total_bases;
    if (n_reads < 0 || n_reads > c->max_reads || total > c->max_bases)
        return fail("packed batch (%lld reads, %lld bases) exceeds classifier capacity",
                    (long long)n_reads, (long long)total);
    CK(cudaSetDevice(c->device));
    Slot &s = c->slots[(size_t)slot];
    if (pb->n_dirty > s.cap_dirty || pb->n_ub > s.cap_ub || pb->ub_bytes_used > s.cap_ub_bytes)
        return fail("packed batch has too many non-ACGT words for the slot; use the ASCII submit");
    uint64_t windows = 0;
    for (int64_t r = 0; r < n_reads; r++) {
        int64_t len = offsets[r + 1] - offsets[r];
        if (len >= c->k) windows += (uint64_t)(len - c->k + 1);
    }
    tb_batch_view &v = s.view;
    v.ascii = nullptr;
    v.offsets = s.d_offsets;
    v.n_reads = n_reads;
    v.total_bases = total;
    v.n_tiles = pb->n_tiles;
    v.F = s.d_F;
    v.E = s.d_E;
    v.dirty = s.d_dirty;
    v.tile_first_read = s.d_tile_first;
    v.flags = s.d_flags;
    v.skip = s.d_skip;
    v.ub_list = s.d_ub_list;
    v.ub_count = s.d_ub_count;
    v.n_pass = s.d_n_pass;
    v.ub_ascii = s.d_ub_bytes;
    v.ub_ascii_off = s.d_ub_off;
    v.count_a = s.d_ca;
    v.count_b = s.d_cb;
    v.score_a = score_A ? s.d_sa : nullptr;
    v.score_b = score_B ? s.d_sb : nullptr;
    v.bin = bin ? s.d_bin : nullptr;
    s.windows = windows;
    s.h_ca = count_A;
    s.has_batch = false;  // no ASCII on the device: rerun_resident replays the packed form
    if (n_reads == 0) return 0;
    cudaStream_t st = s.stream;
    CK(cudaMemcpyAsync(s.d_F, pb->F, (size_t)(pb->n_tiles + 1) * 32 * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(s.d_dirty, pb->dirty, (size_t)(pb->n_tiles + 1) * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(s.d_offsets, offsets, (size_t)(n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
    if (pb->n_dirty) {
        CK(cudaMemcpyAsync(s.d_dirty_idx, pb->dirty_idx, (size_t)pb->n_dirty * 4, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(s.d_e_words, pb->e_words, (size_t)pb->n_dirty * 8, cudaMemcpyHostToDevice, st));
    }
    if (pb->n_ub) {
        CK(cudaMemcpyAsync(s.d_ub_read, pb->ub_read, (size_t)pb->n_ub * 4, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(s.d_ub_skip, pb->ub_skip, (size_t)pb->n_ub * 4, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(s.d_ub_off, pb->ub_off, (size_t)pb->n_ub * 8, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(s.d_ub_bytes, pb->ub_bytes, (size_t)pb->ub_bytes_used, cudaMemcpyHostToDevice, st));
    }
    s.packed_resident = true;
    s.packed_n_dirty = pb->n_dirty;
    s.packed_n_ub = pb->n_ub;
    if (enqueue_packed_kernels(c, s, nullptr, nullptr)) return -1;
    CK(cudaMemcpyAsync(count_A, s.d_ca, (size_t)n_reads * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(count_B, s.d_cb, (size_t)n_reads * 4, cudaMemcpyDeviceToHost, st));
    if (score_A) CK(cudaMemcpyAsync(score_A, s.d_sa, (size_t)n_reads * 8, cudaMemcpyDeviceToHost, st));
    if (score_B) CK(cudaMemcpyAsync(score_B, s.d_sb, (size_t)n_reads * 8, cudaMemcpyDeviceToHost, st));
    if (bin) CK(cudaMemcpyAsync(bin, s.d_bin, (size_t)n_reads, cudaMemcpyDeviceToHost, st));
    return 0;
}

int tb_classifier_wait(tb_classifier *c, int slot)
{
    if (!c) return fail("null classifier");
    if (slot < 0 || slot >= (int)c->slots.size()) return fail("slot %d out of range", slot);
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->slots[(size_t)slot].stream));
    return 0;
}

int tb_classifier_mark(tb_classifier *c, int slot, int which)
{
    if (!c) return fail("null classifier");
    if (slot < 0 || slot >= (int)c->slots.size()) return fail("slot %d out of range", slot);
    if (which < 0 || which > 1) return fail("mark index must be 0 or 1");
    CK(cudaSetDevice(c->device));
    CK(cudaEventRecord(c->slots[(size_t)slot].mark[which], c->slots[(size_t)slot].stream));
    return 0;
}

int tb_classifier_elapsed(tb_classifier *c, int slot_from, int which_from, int slot_to,
                          int which_to, float *ms)
{
    if (!c) return fail("null classifier");
    const int n = (int)c->slots.size();
    if (slot_from < 0 || slot_from >= n || slot_to < 0 || slot_to >= n) return fail("slot out of range");
    if ((which_from | which_to) & ~1) return fail("mark index must be 0 or 1");
    CK(cudaSetDevice(c->device));
    CK(cudaEventSynchronize(c->slots[(size_t)slot_to].mark[which_to]));
    CK(cudaEventElapsedTime(ms, c->slots[(size_t)slot_from].mark[which_from],
                            c->slots[(size_t)slot_to].mark[which_to]));
    return 0;
}

int tb_classifier_rerun_resident(tb_classifier *c, int slot, int iters, float *ms_total,
                                 float *ms_classify)
{
    return tb_classifier_rerun_resident_flush(c, slot, iters, 0, ms_total, ms_classify);
}

int tb_classifier_rerun_resident_flush(tb_classifier *c, int slot, int iters, uint64_t flush_bytes,
                                       float *ms_total, float *ms_classify)
{
    if (!c) return fail("null classifier");
    if (slot < 0 || slot >= (int)c->slots.size()) return fail("slot %d out of range", slot);
    if (iters < 1 || iters > 4096) return fail("iters must be 1..4096");
    CK(cudaSetDevice(c->device));
    Slot &s = c->slots[(size_t)slot];
    if ((!s.has_batch && !s.packed_resident) || s.view.n_reads == 0) return fail("no batch resident in slot %d", slot);
    void *scratch = nullptr;
    DeviceScratch guard;
    if (flush_bytes) {
        CK(cudaMalloc(&scratch, flush_bytes));
        guard.mem.push_back(scratch);
    }
    // four events per iteration: step begin/end and classify begin/end
    std::vector<cudaEvent_t> &ev = guard.events;
    ev.assign((size_t)iters * 4, nullptr);
    for (auto &e : ev) CK(cudaEventCreate(&e));
    CK(cudaStreamSynchronize(s.stream));
    for (int i = 0; i < iters; i++) {
        // between timed steps: overwrite a buffer larger than L2 so nothing survives in cache
        if (scratch) CK(cudaMemsetAsync(scratch, i & 0xFF, flush_bytes, s.stream));
        CK(cudaEventRecord(ev[(size_t)4 * i], s.stream));
        if (s.packed_resident ? enqueue_packed_kernels(c, s, ev[(size_t)4 * i + 1], ev[(size_t)4 * i + 2])
                              : enqueue_kernels(c, s, ev[(size_t)4 * i + 1], ev[(size_t)4 * i + 2]))
            return -1;
        CK(cudaEventRecord(ev[(size_t)4 * i + 3], s.stream));
    }
    CK(cudaStreamSynchronize(s.stream));
    float total = 0, cls = 0;
    for (int i = 0; i < iters; i++) {
        float t = 0;
        CK(cudaEventElapsedTime(&t, ev[(size_t)4 * i], ev[(size_t)4 * i + 3]));
        total += t;
        CK(cudaEventElapsedTime(&t, ev[(size_t)4 * i + 1], ev[(size_t)4 * i + 2]));
        cls += t;
    }
    if (ms_total) *ms_total = total;
    if (ms_classify) *ms_classify = cls;
    return 0;
}

int tb_classifier_filter_passes(tb_classifier *c, int slot, uint64_t *n_pass)
{
    if (!c) return fail("null classifier");
    if (slot < 0 || slot >= (int)c->slots.size()) return fail("slot %d out of range", slot);
    CK(cudaSetDevice(c->device));
    Slot &s = c->slots[(size_t)slot];
    CK(cudaStreamSynchronize(s.stream));
    unsigned long long v = 0;
    CK(cudaMemcpy(&v, s.d_n_pass, sizeof v, cudaMemcpyDeviceToHost));
    if (n_pass) *n_pass = c->d_filter ? v : 0;
    return 0;
}

int tb_classifier_sector_runs(tb_classifier *c, int slot, uint64_t *n_runs)
{
    if (!c) return fail("null classifier");
    if (slot < 0 || slot >= (int)c->slots.size()) return fail("slot %d out of range", slot);
    CK(cudaSetDevice(c->device));
    Slot &s = c->slots[(size_t)slot];
    if (!s.view.F || s.view.n_reads == 0) return fail("no batch in slot %d", slot);
    CK(cudaStreamSynchronize(s.stream));
    if (!(c->d_pair && c->d_filter && c->mini)) {
        // hash-addressed bitmap or no bitmap: nothing is shared between windows
        if (n_runs) *n_runs = s.windows;
        return 0;
    }
    unsigned long long *d = nullptr;
    DeviceScratch guard;
    CK(cudaMalloc((void **)&d, sizeof *d));
    guard.mem.push_back(d);
    CK(cudaMemsetAsync(d, 0, sizeof *d, s.stream));
    CK(tb_launch_count_sector_runs(s.view, c->k, (c->filter_word_mask >> 3) + 1, c->hier, d, s.stream));
    unsigned long long v = 0;
    CK(cudaMemcpyAsync(&v, d, sizeof v, cudaMemcpyDeviceToHost, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    if (n_runs) *n_runs = v;
    return 0;
}

int tb_classifier_stats(tb_classifier *c, int slot, uint64_t *n_windows, uint64_t *n_probes,
                        uint64_t *n_fixup_reads, uint64_t *n_kernel_launches)
{
    if (!c) return fail("null classifier");
    if (slot < 0 || slot >= (int)c->slots.size()) return fail("slot %d out of range", slot);
    Slot &s = c->slots[(size_t)slot];
    if (!s.has_batch && !s.packed_resident) return fail("no batch in slot %d", slot);
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(s.stream));
    uint64_t hits_a = 0;
    for (int64_t r = 0; r < s.view.n_reads; r++) hits_a += (uint64_t)s.h_ca[r];
    if (n_windows) *n_windows = s.windows;
    // probes(read) = W + (W - hits_A)  (B is only probed when A misses, Q4)
    if (n_probes) *n_probes = 2 * s.windows - hits_a;
    if (n_fixup_reads) {
        // reads whose skip[] is non-zero
        std::vector<uint32_t> skip((size_t)s.view.n_reads);
        if (s.view.n_reads)
            CK(cudaMemcpy(skip.data(), s.d_skip, (size_t)s.view.n_reads * 4, cudaMemcpyDeviceToHost));
        uint64_t n = 0;
        for (uint32_t x : skip) n += x != 0;
        *n_fixup_reads = n;
    }
    if (n_kernel_launches) *n_kernel_launches = s.view.n_reads ? 5 : 0;
    return 0;
}

int tb_classifier_minimizer_mode(tb_classifier *c)
{
    return c && c->mini ? (c->hier ? 2 : 1) : 0;
}

int tb_classifier_pair_info(tb_classifier *c, uint32_t *n_buckets, uint64_t *n_distinct, int *filter_bits)
{
    if (!c) return fail("null classifier");
    if (n_buckets) *n_buckets = c->d_pair ? c->pair_buckets : 0;
    if (n_distinct) *n_distinct = c->pair_distinct;
    if (filter_bits) *filter_bits = c->filter_bits;
    return 0;
}

int tb_classifier_filter_info(tb_classifier *c, int *log2_bits, float *build_ms)
{
    if (!c) return fail("null classifier");
    if (log2_bits) *log2_bits = c->d_filter ? c->filter_log2 : 0;
    if (build_ms) *build_ms = c->filter_build_ms;
    return 0;
}

int tb_classify_batch(const char *bases, const int64_t *offsets, int64_t n_reads,
                      const hash_set *A, const hash_set *B, int32_t *count_A, int32_t *count_B,
                      double *score_A, double *score_B, char *bin)
{
    if (!valid_set(A) || !valid_set(B)) return fail("invalid set handle");
    if (n_reads < 0) return fail("n_reads < 0");
    if (n_reads == 0) return 0;
    std::lock_guard<std::mutex> lk(g_mutex);
    const int64_t total = offsets[n_reads];
    if (!g_default || g_default->A != A || g_default->B != B || g_default->max_bases < total ||
        g_default->max_reads < n_reads) {
        int64_t mb = std::max<int64_t>(total + total / 4, 1 << 16);
        int64_t mr = std::max<int64_t>(n_reads + n_reads / 4, 1 << 10);
        if (g_default) {
            if (g_default->A == A && g_default->B == B) {
                mb = std::max(mb, g_default->max_bases);
                mr = std::max(mr, g_default->max_reads);
            }
            tb_classifier_destroy(g_default);
            g_default = nullptr;
        }
        g_default = tb_classifier_create(A, B, mb, mr, 1);
        if (!g_default) return -1;
    }
    if (tb_classifier_submit(g_default, 0, bases, offsets, n_reads, count_A, count_B, score_A,
                             score_B, bin))
        return -1;
    return tb_classifier_wait(g_default, 0);
}

// ---- legacy symbols (c/kmers.c) -------------------------------------------------

hash_set *create_kmer_hash_set(char *kmer_file_path)
{
    hash_set *s = tb_set_create_from_file(kmer_file_path, g_device);
    if (!s) fprintf(stderr, "create_kmer_hash_set: %s\n", g_err.c_str());
    else fprintf(stderr, "Done!\n");
    return s;
}

void count_kmers_in_read(char *read, hash_set *haplotype_A, hash_set *haplotype_B, int *count_A,
                         int *count_B)
{
    int64_t offsets[2] = { 0, (int64_t)strlen(read) };  // c/kmers.c:285
    int32_t ca = -1, cb = -1;
    if (tb_classify_batch(read, offsets, 1, haplotype_A, haplotype_B, &ca, &cb, NULL, NULL, NULL)) {
        fprintf(stderr, "count_kmers_in_read: %s\n", g_err.c_str());
        ca = cb = -1;
    }
    *count_A = ca;
    *count_B = cb;
}

uint64_t kmer_to_int(char *kmer, unsigned char k)
{
    uint64_t v = 0;
    for (int i = 0; i < k && i < 32; i++)
        v |= (uint64_t)tb_base_code((unsigned char)kmer[i]) << (2 * i);
    return v;
}

void reverse_complement(char *kmer_in, char *kmer_out, unsigned char k)
{
    for (int i = 0; i < k; i++) {
        unsigned char c = (unsigned char)kmer_in[i];
        if (tb_base_valid(c)) kmer_out[k - 1 - i] = "TGCA"[tb_base_code(c)];
    }
}

unsigned int hash_function(uint64_t x)
{
    x = (x ^ (x >> 16)) * 0x45d9f3bull;
    x = (x ^ (x >> 16)) * 0x45d9f3bull;
    return (unsigned int)(x ^ (x >> 16));
}

char kmer_in_hash_set(char *kmer, char *empty_kmer_buffer, hash_set *set)
{
    if (!valid_set(set)) { fail("invalid set handle"); return 0; }
    uint64_t fwd = kmer_to_int(kmer, set->k);
    reverse_complement(kmer, empty_kmer_buffer, set->k);
    uint64_t rc = kmer_to_int(empty_kmer_buffer, set->k);
    uint64_t key = fwd < rc ? fwd : rc;
    unsigned char out = 0;
    if (tb_set_contains_keys(set, &key, 1, &out)) {
        fprintf(stderr, "kmer_in_hash_set: %s\n", g_err.c_str());
        return 0;
    }
    return (char)out;
}

// ---- microbenchmark ----------------------------------------------------------------

int tb_gather_bench(uint64_t table_bytes, int loads_in_flight, int threads_per_block,
                    int blocks_per_sm, int steps, int variant, double *sectors_per_s, float *ms)
{
    if (use_device(g_device)) return -1;
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, g_device));
    const uint64_t n_sectors = table_bytes / 32;
    if (n_sectors == 0) return fail("table too small");
    void *buf = nullptr;
    unsigned long long *sink = nullptr;
    CK(cudaMalloc(&buf, n_sectors * 32));
    CK(cudaMalloc((void **)&sink, 8));
    CK(cudaMemset(buf, 0x5A, n_sectors * 32));
    const int blocks = p.multiProcessorCount * blocks_per_sm;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    // warm-up, then timed
    CK(tb_launch_gather(buf, n_sectors, loads_in_flight, threads_per_block, blocks,
                        std::max(1, steps / 8), variant, sink, 0));
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0, 0));
    CK(tb_launch_gather(buf, n_sectors, loads_in_flight, threads_per_block, blocks, steps, variant,
                        sink, 0));
    CK(cudaEventRecord(e1, 0));
    CK(cudaEventSynchronize(e1));
    float t = 0;
    CK(cudaEventElapsedTime(&t, e0, e1));
    const double sectors = (double)blocks * threads_per_block * (double)steps * loads_in_flight;
    if (sectors_per_s) *sectors_per_s = sectors / (t * 1e-3);
    if (ms) *ms = t;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    cudaFree(sink);
    return 0;
}

int tb_h2d_bench(uint64_t bytes, int iters, double *gb_per_s)
{
    if (use_device(g_device)) return -1;
    if (bytes == 0 || iters < 1) return fail("tb_h2d_bench: nothing to copy");
    void *host = nullptr, *dev = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaError_t e = cudaHostAlloc(&host, bytes, cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaMalloc(&dev, bytes);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&e0);
    if (e == cudaSuccess) e = cudaEventCreate(&e1);
    if (e == cudaSuccess) {
        memset(host, 0x5A, bytes);  // first touch on this thread's NUMA node
        for (int i = 0; i < std::max(2, iters / 10) && e == cudaSuccess; i++)
            e = cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e == cudaSuccess) e = cudaEventRecord(e0, stream);
    for (int i = 0; i < iters && e == cudaSuccess; i++) e = cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaEventRecord(e1, stream);
    if (e == cudaSuccess) e = cudaEventSynchronize(e1);
    float t = 0;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&t, e0, e1);
    if (e == cudaSuccess && gb_per_s) *gb_per_s = (double)bytes * iters / (t * 1e-3) / 1e9;
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (stream) cudaStreamDestroy(stream);
    if (dev) cudaFree(dev);
    if (host) cudaFreeHost(host);
    if (e != cudaSuccess) return fail("tb_h2d_bench: %s", cudaGetErrorString(e));
    return 0;
}

}  // extern "C"
