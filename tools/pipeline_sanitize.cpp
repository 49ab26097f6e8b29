// The native streaming pipeline (reader -> packer -> device workers -> in-order router -> three
// writers + TSV) with the GPU classifier replaced by a stub, so that its host threads can be run
// under ThreadSanitizer / AddressSanitizer on a machine without a GPU.  Two modes:
//   stub   the pipeline's own stand-in (tb_route_fastx: record i -> bins[i % n]), N worker threads
//   gpu    the full GPU code path of the pipeline (packer, submit_packed / wait, replicas) against
//          link-time stubs of the classifier calls below
//
//   cd trio_binning_b200/csrc
//   g++ -O1 -g -std=c++17 -mavx2 -fsanitize=thread -I. -o /tmp/pipe_tsan ../../tools/pipeline_sanitize.cpp \
//       fastx_pipeline.cpp gz_stream.cpp fastq_deflate.cpp host_pack.cpp -lz -lpthread
//   /tmp/pipe_tsan reads.fastq.gz /tmp/out 100000 8 gpu 3 gz      # 100-kbase batches, 8 threads, 3 workers
//   (same with -fsanitize=address,undefined)
// Prints the stream statistics; exit code 0 = pipeline returned 0.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../include/trio_kmers_b200.h"

extern "C" {
static thread_local std::string g_err;
const char *tb_last_error(void) { return g_err.c_str(); }
void tb_internal_set_error(const char *m) { g_err = m ? m : ""; }
void *tb_host_alloc(uint64_t bytes) { return malloc(bytes ? bytes : 1); }
void tb_host_free(void *p) { free(p); }
struct tb_classifier {
    int32_t *ca[4], *cb[4];
    double *sa[4], *sb[4];
    char *bin[4];
    std::vector<int64_t> off[4];
};
tb_classifier *tb_classifier_create(const hash_set *, const hash_set *, int64_t, int64_t, int) { return new tb_classifier(); }
void tb_classifier_destroy(tb_classifier *c) { delete c; }
static void remember(tb_classifier *c, int slot, const int64_t *offsets, int64_t n, int32_t *ca, int32_t *cb, double *sa,
                     double *sb, char *bin)
{
    c->off[slot].assign(offsets, offsets + n + 1);
    c->ca[slot] = ca;
    c->cb[slot] = cb;
    c->sa[slot] = sa;
    c->sb[slot] = sb;
    c->bin[slot] = bin;
}
int tb_classifier_submit(tb_classifier *c, int slot, const char *, const int64_t *offsets, int64_t n_reads, int32_t *ca,
                         int32_t *cb, double *sa, double *sb, char *bin)
{
    remember(c, slot, offsets, n_reads, ca, cb, sa, sb, bin);
    return 0;
}
int tb_classifier_submit_packed(tb_classifier *c, int slot, const tb_packed_batch *pb, const int64_t *offsets, int32_t *ca,
                                int32_t *cb, double *sa, double *sb, char *bin)
{
    // touch what a real submit would copy to the device
    volatile uint64_t sink = 0;
    for (int64_t i = 0; i < (pb->n_tiles + 1) * 32; i += 97) sink += pb->F[i];
    remember(c, slot, offsets, pb->n_reads, ca, cb, sa, sb, bin);
    return 0;
}
// the results appear at wait(), like the D2H copies of the real thing
int tb_classifier_wait(tb_classifier *c, int slot)
{
    const std::vector<int64_t> &off = c->off[slot];
    for (size_t r = 0; r + 1 < off.size(); r++) {
        const int64_t len = off[r + 1] - off[r];
        c->ca[slot][r] = (int32_t)(len % 7);
        c->cb[slot][r] = (int32_t)(len % 5);
        c->sa[slot][r] = c->ca[slot][r] * 1.5;
        c->sb[slot][r] = c->cb[slot][r] / 3.0;
        c->bin[slot][r] = "ABU"[len % 3];
    }
    return 0;
}
int tb_counter_add_batch(tb_kmer_counter *, const char *, const int64_t *, int64_t) { return 0; }
}

int main(int argc, char **argv)
{
    if (argc < 3) {
        fprintf(stderr, "usage: %s reads.fastx[.gz] out_prefix [batch_bases] [threads] [stub|gpu] [workers] [gz|plain]\n", argv[0]);
        return 2;
    }
    const int64_t batch = argc > 3 ? atoll(argv[3]) : 1000000;
    const int threads = argc > 4 ? atoi(argv[4]) : 4;
    const bool gpu_path = argc > 5 && strcmp(argv[5], "gpu") == 0;
    const int workers = argc > 6 ? atoi(argv[6]) : 2;
    const bool gz = !(argc > 7 && strcmp(argv[7], "plain") == 0);
    const std::string p = argv[2];
    const std::string a = p + ".A" + (gz ? ".gz" : ""), b = p + ".B" + (gz ? ".gz" : ""), u = p + ".U" + (gz ? ".gz" : ""),
                      tsv = p + ".tsv";
    FILE *t = fopen(tsv.c_str(), "wb");
    tb_stream_stats st;
    memset(&st, 0, sizeof st);
    int rc;
    if (gpu_path) {
        hash_set A, B;
        memset(&A, 0, sizeof A);
        memset(&B, 0, sizeof B);
        A.k = B.k = 21;
        A.num_kmers = 4;
        B.num_kmers = 3;
        std::vector<const hash_set *> sa((size_t)workers, &A), sb((size_t)workers, &B);
        tb_stream_options opt;
        memset(&opt, 0, sizeof opt);
        opt.n_devices = workers;
        opt.sets_a = sa.data();
        opt.sets_b = sb.data();
        opt.gzip_output = gz;
        opt.gzip_level = 1;
        opt.fastq_encoder = 1;
        opt.round_robin = workers & 1;
        opt.tsv_fd = fileno(t);
        opt.batch_bases = batch;
        opt.n_threads = threads;
        rc = tb_classify_fastx_multi(argv[1], &opt, a.c_str(), b.c_str(), u.c_str(), &st);
    } else {
        rc = tb_route_fastx(argv[1], "ABUUAB", 6, gz, 1, 1, workers, workers & 1, 0, 0, a.c_str(), b.c_str(), u.c_str(),
                            fileno(t), batch, threads, &st);
    }
    fclose(t);
    printf("rc=%d reads=%llu bases=%llu batches=%llu A/B/U=%llu/%llu/%llu raw=%llu out=%llu devices=%llu  %.2f s "
           "(parse %.2f, pack %.2f, wait %.2f, write %.2f)\n", rc,
           (unsigned long long)st.n_reads, (unsigned long long)st.n_bases, (unsigned long long)st.n_batches,
           (unsigned long long)st.reads_a, (unsigned long long)st.reads_b, (unsigned long long)st.reads_u,
           (unsigned long long)st.output_raw_bytes, (unsigned long long)st.output_gz_bytes,
           (unsigned long long)st.devices_used, st.seconds_total, st.seconds_parse, st.seconds_pack,
           st.seconds_gpu_wait, st.seconds_write);
    return rc ? 1 : 0;
}
