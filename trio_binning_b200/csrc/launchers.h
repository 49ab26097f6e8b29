// launchers.h -- host-callable launch wrappers of the sm_100a kernels.
// (internal; the public C ABI is include/trio_kmers_b200.h)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kmer_device.cuh"

// ---- set build (set_build.cu) ---------------------------------------------
// counters[0] += newly inserted distinct keys, counters[1] |= saw key 2^64-1,
// counters[2] |= malformed line seen (only when check_newline)
cudaError_t tb_launch_insert_text(const char *d_text, uint64_t n_lines, uint64_t text_bytes,
                                  int stride, int k, int check_newline, tb_bucket *table,
                                  uint32_t n_buckets, unsigned long long *d_counters,
                                  cudaStream_t stream);
// keys that are not below 4^k are not inserted; counters[2] |= 1 reports them
cudaError_t tb_launch_insert_keys(const uint64_t *d_keys, uint64_t n, int k, tb_bucket *table,
                                  uint32_t n_buckets, unsigned long long *d_counters,
                                  cudaStream_t stream);
// merge a per-set table (empty = ~0) into the pair table (empty = 0); tag 1 = A, 2 = B
cudaError_t tb_launch_pair_insert(const uint64_t *d_src_table, uint64_t n_slots, int tag,
                                  tb_bucket *pair_table, uint32_t n_buckets,
                                  unsigned long long *d_counters, cudaStream_t stream);
cudaError_t tb_launch_contains(tb_set_view set, const uint64_t *d_keys, uint64_t n,
                               unsigned char *d_out, cudaStream_t stream);

// ---- classify (classify.cu) -------------------------------------------------
// Flat-stream batch layout in HBM (see DESIGN.md "Data layout"):
//   ascii   : total_bases bytes, reads concatenated, no separators
//   offsets : n_reads+1 int64 base offsets
//   F       : 2-bit forward codes, base i at bits [2(i%32)] of word i/32 (non-ACGT -> 0)
//   E       : same for the fill-forward stream; only words whose dirty bit is set are valid
//   dirty   : one u32 per tile (1024 bases = 32 words), bit l = word l has a non-ACGT base
//   tile_first_read : read that contains flat position 1024*t
//   flags   : per read, bit 1 = a non-ACGT byte inside its first k bases (Q3-UB corner)
//   skip    : per read, number of leading windows left to the fix-up kernel
struct tb_batch_view {
    const unsigned char *ascii;
    const int64_t *offsets;
    int64_t n_reads;
    int64_t total_bases;
    int64_t n_tiles;
    uint64_t *F;
    uint64_t *E;
    uint32_t *dirty;
    uint32_t *tile_first_read;
    uint32_t *flags;
    uint32_t *skip;
    uint32_t *ub_list;
    unsigned int *ub_count;  // device counter; zero before the batch, reset by finalize
    unsigned long long *n_pass;  // windows that passed the pre-filter (reset by k_index)
    // Q3-UB replay source: by default the batch's own ASCII (read r starts at offsets[r]);
    // host-packed batches carry only the prefixes k_fixup needs: entry i of ub_list reads
    // ub_ascii + ub_ascii_off[i]
    const unsigned char *ub_ascii;
    const int64_t *ub_ascii_off;
    int32_t *count_a;
    int32_t *count_b;
    double *score_a;
    double *score_b;
    char *bin;
    // k-mer counting (kmer_count.cu) only: V = one validity bit per base (word i/32, bit i%32),
    // written by k_pack when not null; fold_case: a, c, g, t count as bases (KMC's reading of
    // FASTA/FASTQ; the classify path keeps the reference's upper-case-only rule, c/kmers.c:56-68)
    uint32_t *V;
    int fold_case;
};

// ---- k-mer counting (kmer_count.cu) --------------------------------------------------------
struct tb_count_view {
    unsigned long long *keys;  // canonical k-mers, MSB-first, empty = ~0
    uint32_t *counts;          // null in the packed layout
    uint64_t n_slots;
    uint32_t part, n_parts;    // only keys of partition `part` of `n_parts` are counted
    uint32_t count_shift;      // packed layout (k <= TB_COUNT_PACKED_MAX_K): the count lives in the slot's
                               // word above the key's 2k bits; 0 = separate counts[]
};
// The packed layout needs room for a count that saturates far above every cut-off a reader applies:
// 64 - 2k >= 20 bits.  The saturation test is not atomic with the add, so the field must also absorb
// one add from every thread that can be in flight (< 2^19 on this part) beyond the threshold.
#define TB_COUNT_PACKED_MAX_K 22
#define TB_COUNT_PACKED_SATURATE (1ull << 19)
// stats[0] += windows counted, stats[1] += new distinct keys, stats[2] |= table full
cudaError_t tb_launch_count_windows(const uint64_t *d_F, const uint32_t *d_V, const int64_t *d_offsets, int64_t n_reads,
                                    int64_t total_bases, int k, tb_count_view T, unsigned long long *d_stats,
                                    cudaStream_t stream);
cudaError_t tb_launch_count_histogram(tb_count_view T, unsigned long long *d_hist, int n_bins, cudaStream_t stream);
cudaError_t tb_launch_count_select(tb_count_view A, tb_count_view B, uint32_t min_count, uint32_t max_count,
                                   uint32_t db_min, uint32_t counter_max, unsigned long long *d_out,
                                   unsigned long long cap, unsigned long long *d_n_out, cudaStream_t stream);

cudaError_t tb_launch_pack(const tb_batch_view &b, int k, cudaStream_t stream);
cudaError_t tb_launch_index(const tb_batch_view &b, int k, cudaStream_t stream);
// host-packed batches: index without ASCII, scatter of the dirty E words, Q3-UB list from the host
cudaError_t tb_launch_index_packed(const tb_batch_view &b, const uint32_t *d_dirty_idx,
                                   const uint64_t *d_e_words, int64_t n_dirty, const uint32_t *d_ub_read,
                                   const uint32_t *d_ub_skip, int64_t n_ub, cudaStream_t stream);
cudaError_t tb_launch_classify(const tb_batch_view &b, int k, tb_set_view A, tb_set_view B,
                               cudaStream_t stream);
// filtered variant: `filter` is a bitmap of filter_mask+1 bits over the keys of A and B
cudaError_t tb_launch_classify_filtered(const tb_batch_view &b, int k, tb_set_view A, tb_set_view B,
                                        const uint32_t *filter, uint32_t filter_mask,
                                        cudaStream_t stream);
cudaError_t tb_launch_filter_build(const uint64_t *d_table, uint64_t n_slots, uint32_t *filter,
                                   uint32_t filter_mask, cudaStream_t stream);
// pair-table variant: one bitmap word (n_bits = 1 or 2 bits per key) + one pair bucket per
// surviving window; the bitmap is built from the pair table
cudaError_t tb_launch_classify_pair(const tb_batch_view &b, int k, tb_pair_view P,
                                    const uint32_t *filter, uint32_t filter_word_mask, int n_bits,
                                    cudaStream_t stream);
cudaError_t tb_launch_pair_filter_build(const uint64_t *d_pair_table, uint64_t n_slots, uint32_t *filter,
                                        uint32_t filter_word_mask, int n_bits, cudaStream_t stream);
// minimizer-addressed variant for large sets: sector (8 words) chosen by the key's minimizer;
// p2_width = table probes per lane and phase-2 round (2 or 4)
cudaError_t tb_launch_classify_pair_mini(const tb_batch_view &b, int k, tb_pair_view P,
                                         const uint32_t *filter, uint32_t n_sectors, int n_bits,
                                         int p2_width, int hier, cudaStream_t stream);
// measurement aid: bitmap sector fetches of the batch with per-lane re-use (adds to *d_out)
cudaError_t tb_launch_count_sector_runs(const tb_batch_view &b, int k, uint32_t n_sectors, int hier,
                                        unsigned long long *d_out, cudaStream_t stream);
// hier != 0: two-level addressing (line from the 13-mer minimizer, sector from the m-mer one)
cudaError_t tb_launch_pair_filter_build_mini(const uint64_t *d_pair_table, uint64_t n_slots, uint32_t *filter,
                                             uint32_t n_sectors, int n_bits, int k, int hier, cudaStream_t stream);
cudaError_t tb_launch_fixup(const tb_batch_view &b, int k, tb_set_view A, tb_set_view B,
                            cudaStream_t stream);
cudaError_t tb_launch_finalize(const tb_batch_view &b, double scale_a, double scale_b,
                               cudaStream_t stream);

// classify kernel tuning knobs (read once from the environment by capi.cu)
void tb_classify_config(int unroll, int threads_per_block, int filter_unroll, int filter_mode);

// ---- microbenchmark (gather_bench.cu) ----------------------------------------
cudaError_t tb_launch_gather(const void *d_buf, uint64_t n_sectors, int loads_in_flight,
                             int threads_per_block, int blocks, int steps, int variant,
                             unsigned long long *d_sink, cudaStream_t stream);
