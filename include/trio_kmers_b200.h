/*
 * trio_kmers_b200.h -- C ABI of the B200-native replacement for trio_binning's
 * native k-mer library (reference: c/kmers.c, loaded by src/trio_binning/kmers.py:30-38
 * through ctypes).
 *
 * The shared library built from trio_binning_b200/csrc/ is called
 * `kmers_c<EXT_SUFFIX>` so that the reference's own loader finds it.  It exports
 *
 *   (1) the LEGACY symbols the reference's kmers.py binds, with identical names,
 *       argument meaning and struct prefix -- each backed by CUDA kernels on
 *       sm_100a, no CPU fallback (a missing / failing device is an error), and
 *   (2) a batched / streaming API (prefix tb_) that the reference does not have
 *       and that a GPU needs: many reads per call, pinned host buffers,
 *       double-buffered asynchronous slots, device selection, explicit destroy.
 *
 * Plain C types only: pointers, sizes, ints, doubles.  No torch types.
 *
 * Error convention for tb_* functions: return 0 on success, a negative code on
 * failure; tb_last_error() returns a thread-local message.  Legacy functions
 * keep the reference's signatures (no return codes): on failure
 * create_kmer_hash_set returns NULL and count_kmers_in_read writes -1/-1 and
 * records the message; they print it to stderr as well.
 */
#ifndef TRIO_KMERS_B200_H
#define TRIO_KMERS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------
 * Set handle.  The first 32 bytes are the reference's `hash_set`
 * (c/kmers.c:12-38; mirrored by kmers.py:53-59): Python only ever reads
 * `.contents.num_kmers` (kmers.py:157-159).  `kmers` / `full` are NULL -- the
 * table lives in HBM -- and nothing on the host dereferences them.
 * ---------------------------------------------------------------------- */
typedef struct hash_set {
    uint64_t *kmers;        /* always NULL on the host side                        */
    unsigned char *full;    /* always NULL                                          */
    int hash_size;          /* reference value num_kmers*4/3 (c/kmers.c:166)        */
    unsigned char k;        /* k-mer length, 1..32                                  */
    int num_kmers;          /* number of LINES of the k-mer file (c/kmers.c:131-136)*/
    /* ---- B200 extension (opaque to callers) ---- */
    uint32_t magic;         /* 'TB20'                                               */
    int device;             /* CUDA device ordinal holding the table                */
    uint64_t *d_table;      /* device: n_buckets x 4 x u64, empty slot = ~0         */
    uint32_t n_buckets;     /* 32-byte buckets                                      */
    uint32_t has_all_ones;  /* key 2^64-1 (k=32 poly-T) stored out of band          */
    uint64_t n_distinct;    /* distinct keys resident in the table                  */
    double build_ms;        /* device time of parse+insert                          */
} hash_set;

/* ------------------------------------------------------------------------
 * (1) Legacy symbols -- drop-in for c/kmers.c
 * ---------------------------------------------------------------------- */

/* Replaces create_kmer_hash_set, c/kmers.c:185-229 (+ peek_at_file :124-146,
 * initialize_hash_set :160-180, add_to_hash :112-122): k = length of the first
 * line, num_kmers = number of lines, keys inserted RAW (not canonicalised) by
 * the sm_100a insert kernel.  Uses the current tb device (default 0). */
hash_set *create_kmer_hash_set(char *kmer_file_path);

/* Replaces count_kmers_in_read, c/kmers.c:270-299 (+ kmer_in_hash_set :245-268):
 * a batch of one read through the CUDA classify kernel. */
void count_kmers_in_read(char *read, hash_set *haplotype_A, hash_set *haplotype_B,
                         int *count_A, int *count_B);

/* Replaces kmer_to_int, c/kmers.c:50-72 (string utility; host arithmetic). */
uint64_t kmer_to_int(char *kmer, unsigned char k);

/* Replaces reverse_complement, c/kmers.c:74-93 (non-ACGT positions of `out`
 * are left untouched, as in the reference). */
void reverse_complement(char *kmer_in, char *kmer_out, unsigned char k);

/* Replaces hash_function, c/kmers.c:98-103.  Kept for API completeness; the
 * device table uses its own hash (membership does not depend on layout). */
unsigned int hash_function(uint64_t x);

/* Replaces kmer_in_hash_set, c/kmers.c:245-268, for one k-mer string: canonical
 * min(fwd, revcomp) lookup in the device table.  `empty_kmer_buffer` receives
 * the reverse complement like the reference's scratch buffer. */
char kmer_in_hash_set(char *kmer, char *empty_kmer_buffer, hash_set *set);

/* ------------------------------------------------------------------------
 * (2) B200 API
 * ---------------------------------------------------------------------- */

const char *tb_last_error(void);
const char *tb_version(void);

/* Devices.  One process per GPU is the intended model (tb_set_device once). */
int tb_device_count(void);
int tb_set_device(int device);
int tb_get_device(void);
/* name into buf; sm count, HBM bytes, L2 bytes returned through pointers (may be NULL) */
int tb_device_info(int device, char *name_buf, int name_len, int *sm_count,
                   uint64_t *hbm_bytes, uint64_t *l2_bytes);

/* Pinned host memory for batch buffers. */
void *tb_host_alloc(uint64_t bytes);
void tb_host_free(void *p);

/* Sets. */
hash_set *tb_set_create_from_file(const char *path, int device);
/* keys: RAW LSB-first 2-bit encodings (c/kmers.c:50-72) in host memory;
 * num_kmers := n (as if the file had n lines).  A key that is not below 4^k is not the
 * kmer_to_int value of any k-mer: the call fails (NULL) instead of storing it. */
hash_set *tb_set_create_from_keys(const uint64_t *keys, uint64_t n, int k, int device);
void tb_set_destroy(hash_set *s);
/* membership of already-encoded lookup keys (host arrays), for tests */
int tb_set_contains_keys(const hash_set *s, const uint64_t *keys, uint64_t n, unsigned char *out);
/* load factor used for new tables (slots = 4*n_buckets); default 0.5, (0,0.95] */
int tb_set_load_factor(double lf);

/* Scaling factors, classify_by_kmers.py:57-77 (fp64). */
void tb_scaling_factors(const hash_set *A, const hash_set *B, double *scale_A, double *scale_B);

/* One-shot batch: reads are concatenated ASCII in `bases` (host memory, pinned or
 * pageable), read r = bases[offsets[r] .. offsets[r+1]); offsets has n_reads+1
 * entries.  Outputs are host arrays of n_reads entries; score_* and bin may be
 * NULL.  bin[r] in {'A','B','U'} (classify_by_kmers.py:104-115, on device, fp64).
 * Per read this equals count_kmers_in_read + the driver's decision. */
int tb_classify_batch(const char *bases, const int64_t *offsets, int64_t n_reads,
                      const hash_set *A, const hash_set *B,
                      int32_t *count_A, int32_t *count_B,
                      double *score_A, double *score_B, char *bin);

/* Double-buffered asynchronous classifier: `n_slots` independent slots, each with
 * its own CUDA stream and device buffers for up to max_bases / max_reads per
 * batch.  submit() enqueues H2D copy + pack + classify + decide + D2H copy on
 * the slot's stream and returns; wait() blocks until that slot's results are in
 * the host arrays given to submit().  Host buffers should come from
 * tb_host_alloc for true asynchrony. */
typedef struct tb_classifier tb_classifier;
tb_classifier *tb_classifier_create(const hash_set *A, const hash_set *B,
                                    int64_t max_bases, int64_t max_reads, int n_slots);
void tb_classifier_destroy(tb_classifier *c);
int tb_classifier_submit(tb_classifier *c, int slot,
                         const char *bases, const int64_t *offsets, int64_t n_reads,
                         int32_t *count_A, int32_t *count_B,
                         double *score_A, double *score_B, char *bin);
int tb_classifier_wait(tb_classifier *c, int slot);
/* ---- host-packed batches (north_star: "pinned, 2-bit-packed batches") -------------
 * tb_pack_reads converts concatenated ASCII reads into the flat 2-bit layout the
 * classify kernels consume, on the host (multi-threaded, AVX2 / AVX-512), so that 0.25 byte per
 * base crosses PCIe.  All arrays are caller-owned (pinned memory from tb_host_alloc for
 * asynchronous copies); cap_* are their capacities, the n_* fields are outputs.
 *   F          cap_words  u64: >= (ceil(total/1024)+1)*32
 *   dirty      cap_tiles  u32: >= ceil(total/1024)+1
 *   dirty_idx / e_words   cap_dirty: words holding a non-ACGT byte and their fill-forward form
 *   ub_*       reads with a non-ACGT byte among their first k bases (quirk Q3-UB): read
 *              index, number of leading windows to replay, offset of its ASCII prefix in
 *              ub_bytes
 * Returns 0, -1 bad arguments, -2 F/dirty too small, -3 rare-path arrays too small. */
typedef struct tb_packed_batch {
    int64_t n_reads, total_bases, n_tiles;
    uint64_t *F;
    int64_t cap_words;
    uint32_t *dirty;
    int64_t cap_tiles;
    uint32_t *dirty_idx;
    uint64_t *e_words;
    int64_t cap_dirty, n_dirty;
    uint32_t *ub_read, *ub_skip;
    int64_t *ub_off;
    unsigned char *ub_bytes;
    int64_t cap_ub, n_ub, cap_ub_bytes, ub_bytes_used;
} tb_packed_batch;
int tb_pack_reads(const char *bases, const int64_t *offsets, int64_t n_reads, int k, int n_threads,
                  tb_packed_batch *out);
/* Array capacities a tb_packed_batch needs to be usable with this classifier's slots. */
int tb_classifier_capacity(tb_classifier *c, int64_t *max_bases, int64_t *max_reads, int64_t *cap_words,
                           int64_t *cap_tiles, int64_t *cap_dirty, int64_t *cap_ub, int64_t *cap_ub_bytes);
/* Enqueue a host-packed batch on a slot: H2D of F/dirty/offsets (+ the rare-path arrays),
 * then index + classify + fix-up + decide and the D2H of the results, like submit(). */
int tb_classifier_submit_packed(tb_classifier *c, int slot, const tb_packed_batch *batch,
                                const int64_t *offsets, int32_t *count_A, int32_t *count_B,
                                double *score_A, double *score_B, char *bin);

/* ---- native streaming pipeline ---------------------------------
 * Classify a whole FASTA/FASTQ(.gz) file: a reader thread parses records with the exact
 * semantics of the reference's readfq (seq.py:45-83) into pinned batches, batches are
 * classified on the GPU with two slots in flight, every record is written to the gzip file
 * of its bin (independent gzip members compressed in parallel; the decompressed bytes
 * equal the reference's output) and one TSV line per read (name, bin, score A, score B;
 * classify_by_kmers.py:117, Python float repr) goes to tsv_fd (-1: none).
 * out_a/out_b/out_u are complete file names.  Returns 0 or a negative error code. */
typedef struct tb_stream_stats {
    uint64_t n_reads, n_bases, n_batches, input_bytes;
    uint64_t reads_a, reads_b, reads_u;
    uint64_t output_raw_bytes, output_gz_bytes;
    double seconds_total, seconds_parse, seconds_gpu_wait, seconds_write;
    /* round 2: where the host time goes.  seconds_parse is the reader thread's wall time (it
     * includes waiting for the inflate and for free batches); seconds_inflate is the time spent
     * inside the gzip decoder, summed over its threads; seconds_pack the 2-bit packing;
     * seconds_deflate_cpu the output compression, summed over the pool's threads;
     * seconds_gpu_wait is summed over the device workers. */
    double seconds_inflate, seconds_pack, seconds_deflate_cpu;
    uint64_t devices_used;
} tb_stream_stats;
/* One process, several GPUs: reads are parsed once, packed to 2 bits on the host and dealt
 * batch by batch to `n_devices` device workers, each with its own replica of the two sets
 * (sets_a[d] / sets_b[d] live on the device that worker drives; tb_set_clone makes replicas).
 * Results are merged in input order, so outputs and TSV are byte-identical for any n_devices.
 * Replaces the reference's per-read loop, classify_by_kmers.py:99-117.
 *   gzip_output 1: three gzip files out_a/out_b/out_u (seq.py:132-134);
 *               0: the reference's plain mode (seq.py:127-130): text files out_a and out_u --
 *                  the B handle is opened on out_a as upstream does (out_b is never created),
 *                  with CPython's buffering restated so the overlap comes out byte-identical;
 *   round_robin 0: a worker takes the next batch whenever it has a free slot; 1: batch i goes to
 *                  device i mod n_devices. */
typedef struct tb_stream_options {
    int n_devices;
    const hash_set *const *sets_a;
    const hash_set *const *sets_b;
    int gzip_output, gzip_level, fastq_encoder, round_robin;
    int tsv_fd;
    int64_t batch_bases;
    int n_threads;
    /* plain output only: which interpreter's text writer to restate -- when TextIOWrapper.write
     * flushes its pending list before taking a new string: 0 = the string alone is >= 8192
     * bytes (CPython 3.9 .. 3.12), 1 = list + string would exceed 8192 bytes (later versions),
     * 2 = never (3.8) -- and the BufferedWriter's size (0 = the file's st_blksize).  The
     * Python caller probes both on the running interpreter. */
    int py_text_mode;
    int64_t py_buffer_size;
} tb_stream_options;
int tb_classify_fastx_multi(const char *reads_path, const tb_stream_options *opt, const char *out_a,
                            const char *out_b, const char *out_u, tb_stream_stats *stats);
/* Replica of a set on another device (table copied device to device, over NVLink where the
 * GPUs are peers); num_kmers, k and the rest of the reference's fields are kept. */
hash_set *tb_set_clone(const hash_set *src, int device);
/* Testing aid (no GPU): the host pipeline -- reader, batches, n_workers stand-ins for the device
 * workers, router, the three writers, TSV -- with record i assigned to bins[i % n_bins]. */
int tb_route_fastx(const char *reads_path, const char *bins, int64_t n_bins, int gzip_output, int gzip_level,
                   int fastq_encoder, int n_workers, int round_robin, int py_text_mode, int64_t py_buffer_size,
                   const char *out_a, const char *out_b, const char *out_u, int tsv_fd, int64_t batch_bases,
                   int n_threads, tb_stream_stats *stats);
int tb_classify_fastx(const char *reads_path, const hash_set *A, const hash_set *B,
                      const char *out_a, const char *out_b, const char *out_u, int gzip_level,
                      int tsv_fd, int64_t batch_bases, int n_threads, tb_stream_stats *stats);
/* Python's repr(float) of x into buf (testing aid for the TSV writer). */
int tb_float_repr(double x, char *buf, int buf_len);
/* Parse a file with the native reader and write every record to out_fd as three
 * length-prefixed fields -- name, sequence, quality; each an int64 (little-endian) length
 * followed by the bytes, length -1 for "no quality" (testing aid; no GPU involved).  Returns the number of records, or < 0: -1 cannot
 * open, -2 unreadable input (corrupt gzip data; message in tb_last_error), -3 out of memory. */
int64_t tb_fastx_dump(const char *reads_path, int out_fd);
/* The lines of a FASTA/FASTQ(.gz) file as the native reader's input stage delivers them (gzip,
 * BGZF or one gzip stream decoded on `decode_threads` cores; line ends translated to '\n'),
 * written to out_fd (testing aid).  Returns the number of bytes, -1 cannot open, -2 unreadable
 * input (message in tb_last_error), -3 write error. */
int64_t tb_fastx_cat(const char *reads_path, int out_fd, int decode_threads);
/* Decode a gzip file with the pipeline's own streaming inflate, `chunk` bytes per read call,
 * and write the result to out_fd (testing aid; no GPU involved).  File semantics are those of
 * Python's gzip module: members concatenated, zero padding after a member skipped, anything
 * else that is not gzip is an error.  Returns the number of bytes written, -1 cannot open,
 * -2 corrupt / non-gzip input, -3 write error.  Replaces what the reference gets from Python's gzip module (seq.py:88-92). */
int64_t tb_gunzip_fd(const char *path, int out_fd, int64_t chunk);
/* The gzip CRC-32 the reader verifies members with (zlib crc32() semantics), and whether its
 * carry-less-multiply path is active on this CPU (testing aids). */
uint32_t tb_crc32(uint32_t crc, const void *buf, int64_t n);
int tb_crc32_accelerated(void);
/* One gzip member for `n` bytes of FASTA/FASTQ text by the writer's FASTQ-aware encoder
 * (csrc/fastq_deflate.h) into out_buf.  Returns the member's size, -1 if the encoder declines
 * the text (short lines, tiny input: the writer uses zlib then) or its self-check failed, -2 if
 * out_cap is too small.  verify != 0: decode and compare before accepting, as the pipeline does
 * (testing aid; no GPU involved). */
int64_t tb_fastq_gzip(const void *data, int64_t n, void *out_buf, int64_t out_cap, int verify);
/* Copy in_fd to the gzip file out_path through the pipeline's output writer: text cut into chunks
 * that n_threads workers compress as independent gzip members, written in order, `piece` bytes
 * appended at a time; fastq_encoder 1 tries the FASTQ-aware encoder before zlib, 2 writes BGZF
 * (<= 64 KiB members with a 'BC' extra field, what bgzip/htslib produce) at `level`.  Returns the
 * number of input bytes or < 0 (testing aid; no GPU involved; the reference writes through
 * Python's gzip.open, seq.py:132-134). */
int64_t tb_gzip_copy(int in_fd, const char *out_path, int level, int n_threads, int fastq_encoder, int64_t piece);

/* ------------------------------------------------------------------------
 * (3) find-unique-kmers on the GPU: what the reference gets from the KMC
 * binaries (src/trio_binning/find_unique_kmers.py:62-233) -- canonical k-mer counting, the
 * count histogram, set subtraction and the count-range dump.  KMC's documented semantics are
 * restated: canonical = the lexicographically smaller of a k-mer and its reverse complement;
 * windows with anything but A/C/G/T (either case) are skipped; counts saturate at
 * `counter_max` (KMC -cs, 255) and k-mers with fewer than `db_min` occurrences (KMC -ci, 2)
 * are not in a database.  Keys are MSB-first 2-bit encodings of the canonical string (first base
 * in the top digit of 2k bits): their integer order is the order KMC dumps in.
 * ---------------------------------------------------------------------- */
typedef struct tb_kmer_counter tb_kmer_counter;
/* A counting table of n_slots (key, count) slots on `device`.  Only keys of partition `part`
 * of `n_parts` are counted (1 partition = everything): libraries whose distinct k-mers exceed
 * one table are counted in several passes over the input.  replaces `kmc -k..` (:62-103) */
tb_kmer_counter *tb_counter_create(int k, uint64_t n_slots, int part, int n_parts, int device);
/* Device bytes one slot of the table takes for this k (8 up to k = 22, where key and count share
 * a word, else 12): what a caller sizes n_slots with for a memory budget. */
int tb_counter_slot_bytes(int k);
void tb_counter_destroy(tb_kmer_counter *c);
/* Count the canonical k-mers of a batch of reads (concatenated ASCII + offsets, host memory).
 * Returns once the batch is on its way to the device; the caller may reuse its buffers. */
int tb_counter_add_batch(tb_kmer_counter *c, const char *bases, const int64_t *offsets, int64_t n_reads);
/* Count a whole FASTA/FASTQ(.gz) file through the native reader (readfq's records). */
int tb_counter_add_fastx(tb_kmer_counter *c, const char *reads_path, int64_t batch_bases, int n_threads,
                         uint64_t *n_reads, uint64_t *n_bases);
/* Waits for the counting; windows counted, distinct k-mers, table size.  < 0 if the table
 * filled up (results are then incomplete). */
int tb_counter_stats(tb_kmer_counter *c, uint64_t *n_windows, uint64_t *n_distinct, uint64_t *n_slots);
/* hist[i] = number of distinct k-mers seen i times, the last bin collecting everything at or
 * above n_bins-1 (n_bins = 256 gives KMC's saturated counters).  replaces `kmc_tools transform
 * .. histogram` (:106-168), whose file starts at count 2 */
int tb_counter_histogram(tb_kmer_counter *c, uint64_t *hist, int n_bins);
/* K-mers of a's database that are not in b's (b may be NULL) and whose saturated count lies in
 * [min_count, max_count]: `kmc_tools simple a b kmers_subtract` + `kmc_dump -ci -cx`
 * (:171-233).  Stores up to `cap` keys (unordered) and returns how many qualify. */
int64_t tb_counter_select_unique(tb_kmer_counter *a, tb_kmer_counter *b, uint32_t min_count, uint32_t max_count,
                                 uint32_t db_min, uint32_t counter_max, uint64_t *out_keys, uint64_t cap);

/* Timing marks: record CUDA event `which` (0 or 1) on the slot's stream at the
 * current point of its queue; elapsed() waits for the `to` mark and returns the
 * device time between two marks (possibly of different slots) in milliseconds. */
int tb_classifier_mark(tb_classifier *c, int slot, int which);
int tb_classifier_elapsed(tb_classifier *c, int slot_from, int which_from, int slot_to,
                          int which_to, float *ms);
/* Re-run the device part (no copies) of the batch last submitted to `slot` -- pack + index +
 * classify + fix-up + decide for an ASCII batch, index + classify + fix-up + decide for a
 * host-packed one -- `iters` times, timed with CUDA events on the slot's
 * stream.  Returns total milliseconds in *ms_total and the classify kernel's
 * share in *ms_classify (measured with its own event pair).  Used by bench.py
 * for the HBM-resident figure and the roofline. */
int tb_classifier_rerun_resident(tb_classifier *c, int slot, int iters,
                                 float *ms_total, float *ms_classify);
/* Same, with `flush_bytes` > 0 overwriting a scratch buffer of that size between the
 * timed steps (outside the timed intervals) so that no step starts with a warm L2. */
int tb_classifier_rerun_resident_flush(tb_classifier *c, int slot, int iters, uint64_t flush_bytes,
                                       float *ms_total, float *ms_classify);
/* Number of windows of the slot's last batch whose pre-filter bit was set (they are the
 * only ones that probed the tables in HBM); 0 when the direct kernel is in use. */
int tb_classifier_filter_passes(tb_classifier *c, int slot, uint64_t *n_pass);
/* Measurement aid for the roofline: the number of bitmap SECTOR fetches the slot's last batch
 * needs when each lane of the classify kernel re-uses a sector across the consecutive windows
 * that share it (minimizer-addressed bitmap; = windows otherwise).  Runs its own kernel. */
int tb_classifier_sector_runs(tb_classifier *c, int slot, uint64_t *n_runs);
/* The classifier's bitmap pre-filter: log2 of its size in bits (0 = none, the direct
 * kernel is used) and the device time spent building it and the pair table. */
int tb_classifier_filter_info(tb_classifier *c, int *log2_bits, float *build_ms);
/* The classifier's pair table (one tagged table for both sets, k <= 31; TB_PAIR=0 turns it
 * off): buckets (0 = none), distinct keys of A u B, bitmap bits set per key. */
int tb_classifier_pair_info(tb_classifier *c, uint32_t *n_buckets, uint64_t *n_distinct, int *filter_bits);
/* 2 if, in addition to the following, the bitmap's 128-byte lines are addressed by a coarser
 * (13-mer) minimizer -- an experiment for bitmaps far beyond L2 at k = 20..24, enabled with
 * TB_HIER=1 (measured slower than the default).
 * 1 if the classifier's bitmap sectors are addressed by the keys' minimizers, so that
 * consecutive windows of a read re-use sectors from L1 (default whenever the pair table
 * exists, i.e. k <= 31; TB_MINIMIZER=0 selects hash-addressed words). */
int tb_classifier_minimizer_mode(tb_classifier *c);
/* Counters of the last batch in a slot: windows, probes as the reference performs
 * them (W + (W - hits_A)), reads that took the Q3-UB fix-up path. */
int tb_classifier_stats(tb_classifier *c, int slot, uint64_t *n_windows,
                        uint64_t *n_probes, uint64_t *n_fixup_reads, uint64_t *n_kernel_launches);

/* Random-32-byte-sector gather microbenchmark (roofline denominator):
 * every thread issues `loads_in_flight` independent 32-B loads per step at
 * pseudo-random sector addresses of a `table_bytes` buffer.  Returns sectors/s. */
int tb_gather_bench(uint64_t table_bytes, int loads_in_flight, int threads_per_block,
                    int blocks_per_sm, int steps, int variant,
                    double *sectors_per_s, float *ms);
/* Plain host-to-device copies of a page-locked buffer on the current device, `iters` of `bytes`
 * each, timed with events (GB/s).  Measurement aid: run by all ranks of bench.py at once it gives
 * the aggregate the box can feed its GPUs with, the ceiling of the end-to-end figure at N > 1. */
int tb_h2d_bench(uint64_t bytes, int iters, double *gb_per_s);

#ifdef __cplusplus
}
#endif
#endif /* TRIO_KMERS_B200_H */
