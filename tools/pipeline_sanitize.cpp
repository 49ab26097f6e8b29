// The native streaming pipeline (reader -> batches -> classify -> three gzip writers + TSV) with the
// GPU classifier replaced by a stub, so that its host threads can be run under ThreadSanitizer /
// AddressSanitizer on a machine without a GPU.  The stub calls read r of a batch A, B or U in turn.
//
//   cd trio_binning_b200/csrc
//   g++ -O1 -g -std=c++17 -fsanitize=thread -I. -o /tmp/pipe_tsan ../../tools/pipeline_sanitize.cpp \
//       fastx_pipeline.cpp gz_stream.cpp fastq_deflate.cpp -lz -lpthread
//   TB_FASTQ_GZIP=1 /tmp/pipe_tsan reads.fastq.gz /tmp/out 100000     # batch of 100 kbases
//   (same with -fsanitize=address,undefined)
// Prints the stream statistics; exit code 0 = pipeline returned 0.  Check the outputs with
// `zcat /tmp/out.[ABU].gz | wc` against the input's record count.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>

#include "../include/trio_kmers_b200.h"

extern "C" {
static thread_local std::string g_err;
const char *tb_last_error(void) { return g_err.c_str(); }
void tb_internal_set_error(const char *m) { g_err = m ? m : ""; }
void *tb_host_alloc(uint64_t bytes) { return malloc(bytes ? bytes : 1); }
void tb_host_free(void *p) { free(p); }
struct tb_classifier { int dummy; };
tb_classifier *tb_classifier_create(const hash_set *, const hash_set *, int64_t, int64_t, int) { return new tb_classifier(); }
void tb_classifier_destroy(tb_classifier *c) { delete c; }
int tb_classifier_submit(tb_classifier *, int, const char *, const int64_t *offsets, int64_t n_reads, int32_t *ca,
                         int32_t *cb, double *sa, double *sb, char *bin)
{
    for (int64_t r = 0; r < n_reads; r++) {
        const int64_t len = offsets[r + 1] - offsets[r];
        ca[r] = (int32_t)(len % 7);
        cb[r] = (int32_t)(len % 5);
        if (sa) sa[r] = ca[r] * 1.5;
        if (sb) sb[r] = cb[r] * 1.0;
        if (bin) bin[r] = "ABU"[r % 3];
    }
    return 0;
}
int tb_classifier_wait(tb_classifier *, int) { return 0; }
}

int main(int argc, char **argv)
{
    if (argc < 3) {
        fprintf(stderr, "usage: %s reads.fastx[.gz] out_prefix [batch_bases] [threads] [gzip_level]\n", argv[0]);
        return 2;
    }
    hash_set A, B;
    memset(&A, 0, sizeof A);
    memset(&B, 0, sizeof B);
    A.k = B.k = 21;
    A.num_kmers = 4;
    B.num_kmers = 3;
    const std::string p = argv[2];
    const std::string a = p + ".A.gz", b = p + ".B.gz", u = p + ".U.gz", tsv = p + ".tsv";
    FILE *t = fopen(tsv.c_str(), "wb");
    tb_stream_stats st;
    memset(&st, 0, sizeof st);
    const int rc = tb_classify_fastx(argv[1], &A, &B, a.c_str(), b.c_str(), u.c_str(), argc > 5 ? atoi(argv[5]) : 1,
                                     fileno(t), argc > 3 ? atoll(argv[3]) : 1000000, argc > 4 ? atoi(argv[4]) : 4, &st);
    fclose(t);
    printf("rc=%d reads=%llu bases=%llu batches=%llu A/B/U=%llu/%llu/%llu raw=%llu gz=%llu  %.2f s (parse %.2f, classify wait %.2f, write %.2f)\n", rc,
           (unsigned long long)st.n_reads, (unsigned long long)st.n_bases, (unsigned long long)st.n_batches,
           (unsigned long long)st.reads_a, (unsigned long long)st.reads_b, (unsigned long long)st.reads_u,
           (unsigned long long)st.output_raw_bytes, (unsigned long long)st.output_gz_bytes, st.seconds_total,
           st.seconds_parse, st.seconds_gpu_wait, st.seconds_write);
    return rc ? 1 : 0;
}
