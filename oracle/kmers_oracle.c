/*
 * oracle/kmers_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, CPU restatement of the reference's classify-by-kmers arithmetic
 * (esrice/trio_binning, c/kmers.c).  It exists only so that tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg can check / time the
 * CUDA path against it.  Nothing under trio_binning_b200/ may import, link or
 * call anything in this directory.
 *
 * Parity status: PINNED.  This restatement is checked (tests/test_oracle_*.py)
 * against (1) every known-answer vector of the reference's own tests
 * (reference tests/test_kmers.py:8-53), (2) the fixture run of
 * tests/test_classify_by_kmers.py:19-57, and (3) the reference itself compiled
 * from /root/reference/c/kmers.c into oracle/_ref/ (see oracle/build_oracle.py)
 * on randomised inputs; golden vectors generated from that compiled reference
 * are committed under tests/golden/.
 *
 * The restatement is written in "digit space" (2-bit codes), not as a copy of
 * the reference's char loops; each function cites the reference lines whose
 * behaviour it restates.
 *
 * Defined divergences from the reference (both are undefined behaviour or
 * non-termination upstream; see DESIGN.md "Quirks"):
 *   Q3-UB  the reverse-complement scratch buffer is treated as ZERO-INITIALISED
 *          (the reference reads uninitialised malloc memory when one of the
 *          first k bases of a read is not ACGT, c/kmers.c:279 + :78-91).  This
 *          equals the reference compiled with calloc in place of malloc, which
 *          oracle/_ref/kmers_ref_zinit.so is.
 *   Q6     lookups in a 100 %-full table (num_kmers in {1,2}) terminate after
 *          hash_size slots instead of spinning forever (c/kmers.c:258-265).
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ codes */

/* Forward code of a byte: A0 C1 G2 T3, anything else contributes 0
 * (c/kmers.c:56-68: the switch has no default).  Validity is separate. */
static inline int base_code(unsigned char c)
{
    return c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 0;
}

static inline int base_valid(unsigned char c)
{
    return c == 'A' || c == 'C' || c == 'G' || c == 'T';
}

/* c/kmers.c:50-72 -- base i lands at bits [2i, 2i+1] (least-significant first). */
uint64_t orc_kmer_to_int(const char *kmer, unsigned char k)
{
    uint64_t v = 0;
    for (int i = 0; i < k; i++)
        v |= (uint64_t)base_code((unsigned char)kmer[i]) << (2 * i);
    return v;
}

/* c/kmers.c:74-93 -- out[k-1-i] = complement(in[i]) for ACGT, untouched otherwise. */
void orc_reverse_complement(const char *in, char *out, unsigned char k)
{
    static const char comp[4] = { 'T', 'G', 'C', 'A' };
    for (int i = 0; i < k; i++) {
        unsigned char c = (unsigned char)in[i];
        if (base_valid(c))
            out[k - 1 - i] = comp[base_code(c)];
    }
}

/* c/kmers.c:98-103 -- two rounds of xorshift16 * 0x45d9f3b on 64 bits, then
 * one more xorshift16, truncated to 32 bits. */
unsigned int orc_hash_function(uint64_t x)
{
    for (int round = 0; round < 2; round++)
        x = (x ^ (x >> 16)) * UINT64_C(0x45d9f3b);
    x ^= x >> 16;
    return (unsigned int)x;
}

/* -------------------------------------------------------------------- set */

/* Same field order / offsets as the reference struct (c/kmers.c:12-38) so a
 * ctypes mirror of the reference struct can read either. */
typedef struct {
    uint64_t *kmers;
    unsigned char *full;
    int hash_size;
    unsigned char k;
    int num_kmers;
    /* ---- oracle-only extension (after the reference's 32 bytes) ----
     * sorted != NULL: the set's keys live in a caller-owned ascending array and membership is
     * a binary search instead of the reference's probe sequence.  Membership does not depend
     * on the table layout, so every count is unchanged; this form exists because
     * building the reference's table for 400 M keys takes minutes of random DRAM stores.
     * Pinned against the hash-table form in tests/test_oracle.py. */
    const uint64_t *sorted;
    uint64_t n_sorted;
} orc_hash_set;

/* c/kmers.c:160-180 -- hash_size = num_kmers*4/3 computed in 64 bits, stored in int. */
orc_hash_set *orc_set_new(int k, uint64_t num_kmers)
{
    orc_hash_set *s = (orc_hash_set *)calloc(1, sizeof *s);
    s->k = (unsigned char)k;
    s->num_kmers = (int)num_kmers;
    s->hash_size = (int)(num_kmers * 4 / 3);
    size_t n = s->hash_size > 0 ? (size_t)s->hash_size : 1;
    s->kmers = (uint64_t *)calloc(n, sizeof(uint64_t));
    s->full = (unsigned char *)calloc(n, 1);
    return s;
}

void orc_set_free(orc_hash_set *s)
{
    if (!s) return;
    free(s->kmers);
    free(s->full);
    free(s);
}

/* c/kmers.c:112-122 -- RAW key (no canonicalisation, quirk Q1), linear probe
 * to the first free slot, no duplicate check.  Returns 0 if the table is
 * full (cannot happen for tables sized by orc_set_new). */
int orc_set_add_int(orc_hash_set *s, uint64_t key)
{
    if (s->hash_size <= 0) return 0;
    unsigned int pos = orc_hash_function(key) % (unsigned int)s->hash_size;
    for (int n = 0; n < s->hash_size; n++) {
        if (!s->full[pos]) {
            s->full[pos] = 1;
            s->kmers[pos] = key;
            return 1;
        }
        pos = (pos + 1) % (unsigned int)s->hash_size;
    }
    return 0;
}

/* c/kmers.c:257-265 -- probe until an empty slot; Q6: stop after hash_size. */
int orc_set_contains_int(const orc_hash_set *s, uint64_t key)
{
    if (s->sorted) {
        uint64_t lo = 0, hi = s->n_sorted; /* first index with sorted[i] >= key */
        while (lo < hi) {
            uint64_t mid = lo + (hi - lo) / 2;
            if (s->sorted[mid] < key) lo = mid + 1; else hi = mid;
        }
        return lo < s->n_sorted && s->sorted[lo] == key;
    }
    if (s->hash_size <= 0) return 0;
    unsigned int pos = orc_hash_function(key) % (unsigned int)s->hash_size;
    for (int n = 0; n < s->hash_size && s->full[pos]; n++) {
        if (s->kmers[pos] == key) return 1;
        pos = (pos + 1) % (unsigned int)s->hash_size;
    }
    return 0;
}

/* c/kmers.c:124-146 + :185-229.  Two passes over the file:
 *   pass 1: k = (length of first line including its '\n') - 1; num_kmers =
 *           number of getline() records (a final line without '\n' counts).
 *   pass 2: every record contributes the first k bytes of the line buffer.
 * The line buffer is REUSED between records (quirk Q5): a record shorter than
 * k exposes its '\n', its NUL terminator and then stale bytes of earlier
 * records.  Restated with an explicit persistent buffer, zero-initialised
 * (the reference's is malloc'd; the difference is only visible when a record
 * is shorter than k AND no earlier record was as long -- undefined upstream).
 */
orc_hash_set *orc_set_from_file(const char *path)
{
    FILE *fp = fopen(path, "rb");
    if (!fp) return NULL;
    fseek(fp, 0, SEEK_END);
    long size = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    char *text = (char *)malloc((size_t)size + 1);
    if (size > 0 && fread(text, 1, (size_t)size, fp) != (size_t)size) {
        fclose(fp);
        free(text);
        return NULL;
    }
    fclose(fp);

    /* pass 1 */
    uint64_t n_lines = 0;
    long first_len = -1; /* includes '\n' if present */
    for (long p = 0; p < size;) {
        char *nl = (char *)memchr(text + p, '\n', (size_t)(size - p));
        long len = nl ? (nl - (text + p)) + 1 : size - p;
        if (first_len < 0) first_len = len;
        n_lines++;
        p += len;
    }
    /* Empty file: getline fails at once -> characters = (size_t)-1, k = -2 and
     * num_kmers stays 1 upstream; meaningless, we report k=0/num=1, no keys. */
    int k = first_len < 0 ? 0 : (int)first_len - 1;
    if (n_lines == 0) n_lines = 1;
    if (k < 0) k = 0;
    if (k > 255) k = 255;

    orc_hash_set *s = orc_set_new(k, n_lines);

    /* pass 2 */
    size_t cap = 1 << 16;
    char *buf = (char *)calloc(cap, 1);
    for (long p = 0; p < size;) {
        char *nl = (char *)memchr(text + p, '\n', (size_t)(size - p));
        long len = nl ? (nl - (text + p)) + 1 : size - p;
        if ((size_t)len + 1 > cap) {
            size_t ncap = cap;
            while ((size_t)len + 1 > ncap) ncap *= 2;
            buf = (char *)realloc(buf, ncap);
            memset(buf + cap, 0, ncap - cap);
            cap = ncap;
        }
        memcpy(buf, text + p, (size_t)len);
        buf[len] = '\0';
        orc_set_add_int(s, orc_kmer_to_int(buf, (unsigned char)k));
        p += len;
    }
    free(buf);
    free(text);
    return s;
}

/* Build a set directly from integer keys (used by the big synthetic configs,
 * where writing GBs of text is pointless).  num_kmers := n, as if the file had
 * n lines. */
orc_hash_set *orc_set_from_ints(const uint64_t *keys, uint64_t n, int k)
{
    orc_hash_set *s = orc_set_new(k, n);
    for (uint64_t i = 0; i < n; i++)
        orc_set_add_int(s, keys[i]);
    return s;
}

/* Same set, keys given as an ASCENDING array that the caller keeps alive (duplicates allowed:
 * num_kmers := n, as if the file had n lines).  No table is built. */
orc_hash_set *orc_set_from_sorted(const uint64_t *keys, uint64_t n, int k)
{
    orc_hash_set *s = (orc_hash_set *)calloc(1, sizeof *s);
    s->k = (unsigned char)k;
    s->num_kmers = (int)n;
    s->hash_size = (int)(n * 4 / 3);
    s->sorted = keys;
    s->n_sorted = n;
    return s;
}

int orc_set_num_kmers(const orc_hash_set *s) { return s->num_kmers; }
int orc_set_k(const orc_hash_set *s) { return s->k; }
int orc_set_hash_size(const orc_hash_set *s) { return s->hash_size; }

/* ----------------------------------------------------------- per-read count */

/* c/kmers.c:270-299 + :245-268, restated in digit space.
 *
 *   fwd(i)  = sum_j code(read[i+j]) << 2j              (non-ACGT -> 0)
 *   rcbuf   = k sticky digits.  Window i writes rcbuf[k-1-j] = 3-code(read[i+j])
 *             for every VALID read[i+j] and leaves the digit alone otherwise
 *             (quirk Q3).  Digits start at 0 (Q3-UB contract).
 *   rc(i)   = sum_s rcbuf[s] << 2s
 *   key(i)  = min(fwd(i), rc(i));  A hit -> count_A++, else B hit -> count_B++
 *             (quirk Q4: B is only consulted when A misses).
 *
 * k is taken from set A for the window length (c/kmers.c:287-289); each set
 * encodes with its own k (c/kmers.c:251-253 uses set->k).  When the two differ
 * the reference reads the first B->k bytes of a buffer holding an A->k window;
 * that is restated literally below through the shared `win`/`rcbuf` arrays.
 */
static int window_in_set(const unsigned char *win_code, const unsigned char *win_valid,
                         unsigned char *rcbuf, const orc_hash_set *s)
{
    int k = s->k;
    uint64_t fwd = 0, rc = 0;
    for (int j = 0; j < k; j++) {
        fwd |= (uint64_t)win_code[j] << (2 * j);
        if (win_valid[j])
            rcbuf[k - 1 - j] = (unsigned char)(3 - win_code[j]);
    }
    for (int sidx = 0; sidx < k; sidx++)
        rc |= (uint64_t)rcbuf[sidx] << (2 * sidx);
    return orc_set_contains_int(s, fwd < rc ? fwd : rc);
}

void orc_count_kmers_in_read_n(const char *read, long read_length,
                               const orc_hash_set *A, const orc_hash_set *B,
                               int *count_A, int *count_B)
{
    int k = A->k;
    /* scratch sized for the larger k so a mismatched B->k stays in bounds;
     * win[k] is the NUL the reference plants at kmer[k] (c/kmers.c:280). */
    unsigned char win_code[257] = { 0 }, win_valid[257] = { 0 }, rcbuf[257] = { 0 };
    int ca = 0, cb = 0;
    for (long i = 0; i + k <= read_length; i++) {
        for (int j = 0; j < k; j++) {
            unsigned char c = (unsigned char)read[i + j];
            win_code[j] = (unsigned char)base_code(c);
            win_valid[j] = (unsigned char)base_valid(c);
        }
        if (window_in_set(win_code, win_valid, rcbuf, A))
            ca++;
        else if (window_in_set(win_code, win_valid, rcbuf, B))
            cb++;
    }
    *count_A = ca;
    *count_B = cb;
}

/* The reference takes a NUL-terminated string and uses strlen (c/kmers.c:285). */
void orc_count_kmers_in_read(const char *read, const orc_hash_set *A, const orc_hash_set *B,
                             int *count_A, int *count_B)
{
    orc_count_kmers_in_read_n(read, (long)strlen(read), A, B, count_A, count_B);
}

/* Batched form used by tests and by bench.py's cpu_baseline: reads are
 * concatenated in `bases`, read r is bases[offsets[r] .. offsets[r+1]). */
void orc_count_kmers_in_reads(const char *bases, const int64_t *offsets, int64_t n_reads,
                              const orc_hash_set *A, const orc_hash_set *B,
                              int *count_A, int *count_B)
{
    for (int64_t r = 0; r < n_reads; r++)
        orc_count_kmers_in_read_n(bases + offsets[r], (long)(offsets[r + 1] - offsets[r]),
                                  A, B, &count_A[r], &count_B[r]);
}

/* ------------------------------------------------------ score and decision */

/* src/trio_binning/classify_by_kmers.py:72-76 (scaling) and :104-115 (decision),
 * IEEE fp64: s_a = 1.0*max/n_a; score = count*s; A if a>b, B if b>a, else U. */
void orc_scaling_factors(int num_a, int num_b, double *sa, double *sb)
{
    int m = num_a > num_b ? num_a : num_b;
    *sa = 1.0 * (double)m / (double)num_a;
    *sb = 1.0 * (double)m / (double)num_b;
}

void orc_decide(const int *count_A, const int *count_B, int64_t n, double sa, double sb,
                double *score_A, double *score_B, char *bin)
{
    for (int64_t r = 0; r < n; r++) {
        double a = (double)count_A[r] * sa, b = (double)count_B[r] * sb;
        score_A[r] = a;
        score_B[r] = b;
        bin[r] = a > b ? 'A' : (b > a ? 'B' : 'U');
    }
}
