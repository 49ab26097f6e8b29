// fastx_pipeline.cpp -- native host pipeline around the GPU classifier
// (FASTX input and fastq.gz output at more than Python speed).
//
//   reader thread : file / gzip stream -> lines (universal newlines, like Python text mode)
//                   -> FASTA/FASTQ records with the exact semantics of the reference's
//                   readfq (src/trio_binning/seq.py:45-83, quirk Q7) -> batches whose
//                   sequences are concatenated straight into pinned memory
//   main thread   : batch -> tb_classifier_submit on one of n slots (H2D + kernels + D2H
//                   overlap the parsing of the next batches) -> wait -> route every record
//                   to its bin (classify_by_kmers.py:104-115) -> TSV line (:117)
//   deflate pool  : each bin's text is cut into chunks that are compressed in parallel as
//                   independent gzip members and written in order (a concatenation of
//                   members is a valid .gz; the DECOMPRESSED bytes equal the reference's
//                   gzip.open(..., "wt") output, seq.py:132-134)
//
// Only gzip output is handled natively; --no-gzip-output keeps the Python writer because
// the reference's plain mode depends on Python's buffered-writer behaviour (two handles on
// one file, seq.py:129).
#include <fcntl.h>
#include <errno.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <zlib.h>

#include "fastq_deflate.h"
#include "gz_stream.h"

#include <algorithm>
#include <charconv>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <functional>
#include <future>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/trio_kmers_b200.h"

extern "C" void tb_internal_set_error(const char *message);  // capi.cu

namespace {

using Clock = std::chrono::steady_clock;
double seconds_since(Clock::time_point t0)
{
    return std::chrono::duration<double>(Clock::now() - t0).count();
}

// ---------------------------------------------------------------------------------------
// input: bytes -> lines with Python's universal-newline translation ("\r\n", "\r" -> "\n")
// ---------------------------------------------------------------------------------------
class LineSource {
public:
    // host threads that may decode in parallel (BGZF input only); call before the first next()
    void set_threads(int n) { decode_threads_ = n < 1 ? 1 : (n > 8 ? 8 : n); }
    bool open(const char *path, std::string &err)
    {
        const size_t n = strlen(path);
        gz_ = n >= 3 && strcmp(path + n - 3, ".gz") == 0;  // seq.py:88: by file name only
        fd_ = ::open(path, O_RDONLY);
        if (fd_ < 0) { err = std::string("cannot open ") + path; return false; }
        if (gz_) {
            unsigned char head[64];
            const ssize_t got = ::pread(fd_, head, sizeof head, 0);
            size_t total = 0;
            bgzf_ = got >= 18 && bgzf_block_size(head, (size_t)got, total);
        }
        return true;
    }
    ~LineSource()
    {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_.notify_all();
        if (producer_.joinable()) producer_.join();
        for (auto &w : workers_) w.join();
        if (fd_ >= 0) ::close(fd_);
    }
    // Next line into `line` INCLUDING its terminator translated to '\n' (absent only at
    // end of input).  Returns false at end of input (or on a read error: see error()).
    bool next(std::string &line)
    {
        line.clear();
        for (;;) {
            if (pos_ == len_ && !fill()) return !line.empty();
            if (pending_cr_) {
                // a '\r' ended the previous line; swallow the '\n' of a "\r\n" pair
                pending_cr_ = false;
                if (cur_[pos_] == '\n') {
                    pos_++;
                    continue;
                }
            }
            const char *p = cur_ + pos_;
            const size_t avail = len_ - pos_;
            // first '\n' or '\r' (two vectorised scans; '\r' is looked for only up to the '\n')
            const char *nl = (const char *)memchr(p, '\n', avail);
            size_t i = nl ? (size_t)(nl - p) : avail;
            if (const char *cr = (const char *)memchr(p, '\r', i)) i = (size_t)(cr - p);
            line.append(p, i);
            pos_ += i;
            if (i < avail) {
                if (p[i] == '\r') pending_cr_ = true;
                pos_++;
                line.push_back('\n');
                return true;
            }
        }
    }
    uint64_t bytes_in() const { return bytes_in_; }
    const std::string &error() const { return error_; }
    bool parallel_bgzf() const { return bgzf_ && decode_threads_ > 1; }

private:
    // Decompression (or the read syscalls) run ahead of the parser: a producer thread hands
    // over chunks of decoded text in stream order.  Plain gzip is one serial stream (GzStream
    // on the producer thread).  BGZF input (bgzip, htslib: independent <= 64 KB gzip members
    // that carry their compressed size in a 'BC' extra field) is cut into groups of blocks that
    // worker threads decode side by side; anything that stops looking like BGZF is handed to the
    // serial decoder from that offset on.
    static constexpr size_t kChunkBytes = 4u << 20;
    static constexpr size_t kTaskBytes = 512u << 10;  // compressed bytes per BGZF task
    struct Chunk {
        std::vector<char> data;
        size_t len = 0;
        bool done = false, failed = false;
        std::vector<unsigned char> comp;  // BGZF task: the compressed blocks
        size_t expect = 0;                //            and the sum of their ISIZE fields
        std::string why;
    };
    using ChunkPtr = std::shared_ptr<Chunk>;

    // size of the BGZF block starting at p, if it is one (SAM spec 4.1: FEXTRA with a BC subfield)
    static bool bgzf_block_size(const unsigned char *p, size_t n, size_t &total)
    {
        if (n < 18 || p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || !(p[3] & 4)) return false;
        const size_t xlen = p[10] | ((size_t)p[11] << 8);
        if (n < 12 + xlen) return false;
        for (size_t q = 12; q + 4 <= 12 + xlen;) {
            const size_t slen = p[q + 2] | ((size_t)p[q + 3] << 8);
            if (p[q] == 'B' && p[q + 1] == 'C' && slen == 2 && q + 6 <= 12 + xlen) {
                total = (size_t)(p[q + 4] | (p[q + 5] << 8)) + 1;
                return total >= 12 + xlen + 8 + 2;
            }
            q += 4 + slen;
        }
        return false;
    }

    ChunkPtr fresh_chunk()
    {
        std::lock_guard<std::mutex> lk(m_);
        if (!free_.empty()) {
            ChunkPtr c = free_.back();
            free_.pop_back();
            c->len = c->expect = 0;
            c->done = c->failed = false;
            c->comp.clear();
            c->why.clear();
            return c;
        }
        return std::make_shared<Chunk>();
    }
    // append to the ordered queue (blocks while too many chunks are in flight); false = stopping
    bool enqueue(const ChunkPtr &c, bool as_work)
    {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return stop_ || ordered_.size() < max_in_flight_; });
        if (stop_) return false;
        ordered_.push_back(c);
        if (as_work) work_.push_back(c);
        cv_.notify_all();
        return true;
    }
    void finish(const ChunkPtr &c)
    {
        std::lock_guard<std::mutex> lk(m_);
        c->done = true;
        cv_.notify_all();
    }
    void push_end(bool failed, const std::string &why)
    {
        ChunkPtr c = fresh_chunk();
        c->failed = failed;
        c->why = why;
        c->done = true;
        enqueue(c, false);
    }

    // serial stream: the producer thread decodes (or reads) chunk after chunk
    void produce_serial(bool continue_on_fd)
    {
        if (gz_) {
            gzs_.open_fd(fd_, !continue_on_fd);
        }
        for (;;) {
            ChunkPtr c = fresh_chunk();
            c->data.resize(kChunkBytes);
            long got;
            if (gz_) {
                got = gzs_.read(c->data.data(), c->data.size());
                bytes_in_ = bgzf_bytes_ + gzs_.compressed_bytes();
            } else {
                do got = (long)::read(fd_, c->data.data(), c->data.size());
                while (got < 0 && errno == EINTR);
                if (got > 0) bytes_in_ += (uint64_t)got;
            }
            if (got <= 0) {
                push_end(got < 0, gz_ ? ("corrupt gzip input: " + gzs_.error()) : std::string("read error on the reads file"));
                return;
            }
            c->len = (size_t)got;
            c->done = true;
            if (!enqueue(c, false)) return;
        }
    }

    // BGZF: walk the block headers, hand groups of whole blocks to the workers
    void produce_bgzf()
    {
        std::vector<unsigned char> buf(8u << 20);
        size_t have = 0;
        uint64_t off = 0;  // file offset of buf[0]
        bool eof = false;
        for (;;) {
            while (!eof && have < buf.size()) {
                const ssize_t n = ::pread(fd_, buf.data() + have, buf.size() - have, (off_t)(off + have));
                if (n < 0) {
                    if (errno == EINTR) continue;
                    push_end(true, "read error on the reads file");
                    return;
                }
                if (n == 0) eof = true;
                have += (size_t)n;
            }
            size_t p = 0;
            bool not_bgzf = false;
            ChunkPtr task = fresh_chunk();
            for (;;) {
                size_t total = 0;
                if (have - p < 18) break;
                if (!bgzf_block_size(buf.data() + p, have - p, total)) {
                    not_bgzf = true;
                    break;
                }
                if (have - p < total) break;
                const unsigned char *t = buf.data() + p + total - 4;
                const size_t isize = (size_t)t[0] | ((size_t)t[1] << 8) | ((size_t)t[2] << 16) | ((size_t)t[3] << 24);
                if (isize > 65536) {
                    // a BGZF block never holds more than 64 KiB; a larger (untrusted) ISIZE means
                    // this is not BGZF or it is damaged: the serial decoder reports which
                    not_bgzf = true;
                    break;
                }
                task->comp.insert(task->comp.end(), buf.data() + p, buf.data() + p + total);
                task->expect += isize;
                p += total;
                if (task->comp.size() >= kTaskBytes) {
                    if (!enqueue(task, true)) return;
                    task = fresh_chunk();
                }
            }
            if (!task->comp.empty() && !enqueue(task, true)) return;
            memmove(buf.data(), buf.data() + p, have - p);
            have -= p;
            off += p;
            bgzf_bytes_ = off;
            bytes_in_ = off;
            if (not_bgzf || (eof && have > 0) || (have == buf.size())) {
                // not (or no longer) BGZF, or a partial block at the end: the serial decoder
                // takes over from here and decides what it is (member, garbage, truncation)
                if (::lseek(fd_, (off_t)off, SEEK_SET) < 0) {
                    push_end(true, "cannot seek in the reads file");
                    return;
                }
                produce_serial(true);
                return;
            }
            if (eof) {
                push_end(false, "");
                return;
            }
        }
    }

    void work()
    {
        tb::GzStream gz;  // one decoder per worker, re-opened on every task
        for (;;) {
            ChunkPtr c;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return stop_ || !work_.empty(); });
                if (stop_) return;
                c = work_.front();
                work_.pop_front();
            }
            size_t n = 0;
            bool bad = false;
            try {
                c->data.resize(c->expect ? c->expect : 1);
            } catch (const std::exception &) {  // an exception must not leave a std::thread
                bad = true;
                c->why = "out of memory while decoding BGZF blocks";
            }
            gz.open_memory(c->comp.data(), c->comp.size());
            while (!bad) {
                // one spare byte of room shows an overlong stream instead of truncating it
                char extra;
                const long got = n < c->expect ? gz.read(c->data.data() + n, c->expect - n) : gz.read(&extra, 1);
                if (got < 0) { bad = true; c->why = "corrupt gzip input: " + gz.error(); break; }
                if (got == 0) break;
                if (n >= c->expect) { bad = true; c->why = "corrupt gzip input: BGZF block longer than its ISIZE"; break; }
                n += (size_t)got;
            }
            if (!bad && n != c->expect) {
                bad = true;
                c->why = "corrupt gzip input: BGZF block shorter than its ISIZE";
            }
            c->len = n;
            c->failed = bad;
            finish(c);
        }
    }

    void start()
    {
        started_ = true;
        if (parallel_bgzf()) {
            max_in_flight_ = 2 + 2 * (size_t)decode_threads_;
            for (int i = 0; i < decode_threads_; i++) workers_.emplace_back([this] { work(); });
            producer_ = std::thread([this] { produce_bgzf(); });
        } else {
            producer_ = std::thread([this] { produce_serial(false); });
        }
    }

    bool fill()
    {
        if (eof_) return false;
        if (!started_) start();
        std::unique_lock<std::mutex> lk(m_);
        if (cur_chunk_) {  // hand the consumed chunk back
            free_.push_back(cur_chunk_);
            cur_chunk_.reset();
        }
        for (;;) {
            cv_.wait(lk, [&] { return !ordered_.empty() && ordered_.front()->done; });
            ChunkPtr c = ordered_.front();
            ordered_.pop_front();
            cv_.notify_all();
            if (c->failed) {
                error_ = c->why;
                eof_ = true;
                return false;
            }
            if (c->len == 0) {
                if (c->comp.empty()) {  // the end marker
                    eof_ = true;
                    return false;
                }
                free_.push_back(c);  // a task of empty blocks (BGZF end-of-file marker): next
                continue;
            }
            cur_chunk_ = c;
            cur_ = c->data.data();
            pos_ = 0;
            len_ = c->len;
            return true;
        }
    }

    bool gz_ = false, bgzf_ = false, eof_ = false, pending_cr_ = false, started_ = false, stop_ = false;
    int decode_threads_ = 2;
    tb::GzStream gzs_;
    int fd_ = -1;
    std::thread producer_;
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<ChunkPtr> ordered_, work_;
    std::vector<ChunkPtr> free_;
    size_t max_in_flight_ = 3;
    ChunkPtr cur_chunk_;
    const char *cur_ = nullptr;
    size_t pos_ = 0, len_ = 0;
    uint64_t bytes_in_ = 0, bgzf_bytes_ = 0;
    std::string error_;
};

// number of characters (code points) in a UTF-8 byte string: Python counts characters
inline size_t n_chars(const char *p, size_t n)
{
    size_t c = 0;
    for (size_t i = 0; i < n; i++) c += ((unsigned char)p[i] & 0xC0) != 0x80;
    return c;
}

// Python's line[:-1]: drop the last CHARACTER whatever it is
inline size_t chop_len(const std::string &line)
{
    size_t n = line.size();
    if (n == 0) return 0;
    n--;
    while (n > 0 && ((unsigned char)line[n] & 0xC0) == 0x80) n--;
    return n;
}

// ---------------------------------------------------------------------------------------
// batches
// ---------------------------------------------------------------------------------------
struct Rec {
    uint64_t name_off;
    uint32_t name_len;
    uint64_t qual_off;
    uint32_t qual_len;
    bool has_qual;  // qual is not None (it may still be empty)
};

struct Batch {
    bool pinned = true;     // page-locked (GPU pipeline) or plain malloc (parser-only use)
    char *bases = nullptr;
    int64_t cap_bases = 0;
    std::vector<int64_t> offsets;  // n+1
    std::vector<Rec> recs;
    std::string text;  // names and quality strings
    // results (pinned, sized on demand)
    int32_t *ca = nullptr, *cb = nullptr;
    double *sa = nullptr, *sb = nullptr;
    char *bin = nullptr;
    int64_t cap_reads = 0;
    int slot = -1;
    bool last = false;

    int64_t n() const { return (int64_t)recs.size(); }
    int64_t total() const { return offsets.empty() ? 0 : offsets.back(); }
    void reset()
    {
        offsets.assign(1, 0);
        recs.clear();
        text.clear();
        last = false;
    }
    void *alloc(uint64_t n) { return pinned ? tb_host_alloc(n) : malloc(n ? n : 1); }
    void release(void *p)
    {
        if (!p) return;
        if (pinned) tb_host_free(p); else free(p);
    }
    // make room for `need` bytes, keeping the first `used` bytes
    bool grow_bases(int64_t need, int64_t used = 0)
    {
        int64_t cap = std::max<int64_t>(need + need / 4, 1 << 20);
        char *nb = (char *)alloc((uint64_t)cap);
        if (!nb) return false;
        if (bases) {
            memcpy(nb, bases, (size_t)used);
            release(bases);
        }
        bases = nb;
        cap_bases = cap;
        return true;
    }
    bool ensure_results(int64_t n_reads)
    {
        if (n_reads <= cap_reads) return true;
        free_results();
        cap_reads = n_reads + n_reads / 4 + 1024;
        ca = (int32_t *)alloc((uint64_t)cap_reads * 4);
        cb = (int32_t *)alloc((uint64_t)cap_reads * 4);
        sa = (double *)alloc((uint64_t)cap_reads * 8);
        sb = (double *)alloc((uint64_t)cap_reads * 8);
        bin = (char *)alloc((uint64_t)cap_reads);
        return ca && cb && sa && sb && bin;
    }
    void free_results()
    {
        release(ca);
        release(cb);
        release(sa);
        release(sb);
        release(bin);
        ca = cb = nullptr;
        sa = sb = nullptr;
        bin = nullptr;
        cap_reads = 0;
    }
    ~Batch()
    {
        free_results();
        release(bases);
    }
};

// bounded queue of batch pointers
class BatchQueue {
public:
    void push(Batch *b)
    {
        std::unique_lock<std::mutex> lk(m_);
        q_.push_back(b);
        cv_.notify_all();
    }
    Batch *pop()
    {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return !q_.empty(); });
        Batch *b = q_.front();
        q_.pop_front();
        return b;
    }

private:
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<Batch *> q_;
};

// ---------------------------------------------------------------------------------------
// reader: the reference's readfq as an explicit state machine over lines
// ---------------------------------------------------------------------------------------
struct Reader {
    LineSource src;
    int64_t batch_bases;
    BatchQueue *free_q, *full_q;
    std::string error;
    double seconds = 0;
    uint64_t n_reads = 0, n_bases = 0;

    Batch *cur = nullptr;

    void run()
    {
        const auto t0 = Clock::now();
        cur = free_q->pop();
        cur->reset();
        int64_t seq_start = 0;  // where the current record's sequence starts in cur->bases
        int64_t fill = 0;       // bytes of cur->bases in use (incl. the record being read)

        std::string line, header;
        bool have_header = false;  // a pending header line (already chopped, non-empty)
        bool ok = true;
        auto emit = [&](const std::string &name, int64_t seq_len, bool has_qual, const std::string &qual) {
            // close the batch first if this record would overflow a non-empty batch
            Rec r;
            r.name_off = cur->text.size();
            r.name_len = (uint32_t)name.size();
            cur->text.append(name);
            r.qual_off = cur->text.size();
            r.qual_len = (uint32_t)qual.size();
            r.has_qual = has_qual;
            cur->text.append(qual);
            cur->recs.push_back(r);
            cur->offsets.push_back(seq_start + seq_len);
            n_reads++;
            n_bases += (uint64_t)seq_len;
            if (cur->total() >= batch_bases) {
                full_q->push(cur);
                cur = free_q->pop();
                cur->reset();
                fill = 0;
            }
        };
        std::string name, qual;
        while (ok) {
            // ---- find the next header unless one is pending
            if (!have_header) {
                bool found = false;
                while (src.next(line)) {
                    if (line[0] == '>' || line[0] == '@') {
                        header.assign(line, 0, chop_len(line));
                        found = true;
                        break;
                    }
                }
                if (!found || header.empty()) break;
            }
            have_header = false;
            {
                const size_t sp = header.find(' ', 1);
                name.assign(header, 1, sp == std::string::npos ? std::string::npos : sp - 1);
            }
            // ---- sequence lines until a line starting with '@', '+' or '>'
            seq_start = fill = cur->total();
            bool got_marker = false;
            while (src.next(line)) {
                const char c = line[0];
                if (c == '@' || c == '+' || c == '>') {
                    header.assign(line, 0, chop_len(line));
                    got_marker = true;
                    break;
                }
                const size_t n = chop_len(line);
                if (fill + (int64_t)n > cur->cap_bases) {
                    if (!cur->grow_bases(fill + (int64_t)n, fill)) { error = "pinned allocation failed"; ok = false; break; }
                }
                memcpy(cur->bases + fill, line.data(), n);
                fill += (int64_t)n;
            }
            if (!ok) break;
            const int64_t seq_len = fill - seq_start;
            const bool pending = got_marker && !header.empty();
            if (!pending || header[0] != '+') {
                // FASTA record (or end of input)
                emit(name, seq_len, false, std::string());
                if (!pending) break;
                have_header = true;
                continue;
            }
            // ---- FASTQ: quality lines until at least len(seq) CHARACTERS were read
            const size_t seq_chars = n_chars(cur->bases + seq_start, (size_t)seq_len);
            qual.clear();
            size_t have = 0;
            bool complete = false;
            while (src.next(line)) {
                const size_t n = chop_len(line);
                qual.append(line, 0, n);
                have += n_chars(line.data(), line.size()) - 1;
                if (have >= seq_chars) {
                    complete = true;
                    break;
                }
            }
            if (complete) {
                emit(name, seq_len, true, qual);
            } else {
                // input ended inside the quality block: record without qualities, then stop
                emit(name, seq_len, false, std::string());
                break;
            }
        }
        if (!src.error().empty() && error.empty()) error = src.error();  // never a silently short result
        cur->last = true;
        full_q->push(cur);
        seconds = seconds_since(t0);
    }
};

// ---------------------------------------------------------------------------------------
// Python's repr(float) (float_repr_style 'short'): shortest round-trip digits; fixed
// notation when -4 < decpt <= 16, else d.ddde+XX with at least two exponent digits.
// ---------------------------------------------------------------------------------------
void py_float_repr(double x, std::string &out)
{
    if (x != x) { out += "nan"; return; }
    if (x == 1.0 / 0.0) { out += "inf"; return; }
    if (x == -1.0 / 0.0) { out += "-inf"; return; }
    char buf[64];
    auto res = std::to_chars(buf, buf + sizeof buf, x, std::chars_format::scientific);
    std::string s(buf, res.ptr);  // [-]d[.ddd]e[+-]XX
    size_t i = 0;
    if (s[0] == '-') { out += '-'; i = 1; }
    const size_t e = s.find('e');
    std::string digits;
    for (size_t j = i; j < e; j++)
        if (s[j] != '.') digits += s[j];
    const int exp10 = atoi(s.c_str() + e + 1);
    const int decpt = exp10 + 1;
    const int nd = (int)digits.size();
    if (decpt > -4 && decpt <= 16) {
        if (decpt <= 0) {
            out += "0.";
            out.append((size_t)(-decpt), '0');
            out += digits;
        } else if (decpt >= nd) {
            out += digits;
            out.append((size_t)(decpt - nd), '0');
            out += ".0";
        } else {
            out.append(digits, 0, (size_t)decpt);
            out += '.';
            out.append(digits, (size_t)decpt, std::string::npos);
        }
    } else {
        out += digits[0];
        if (nd > 1) {
            out += '.';
            out.append(digits, 1, std::string::npos);
        }
        out += 'e';
        const int ex = decpt - 1;
        out += ex < 0 ? '-' : '+';
        char eb[16];
        snprintf(eb, sizeof eb, "%02d", ex < 0 ? -ex : ex);
        out += eb;
    }
}

// ---------------------------------------------------------------------------------------
// output: parallel gzip members, written in order
// ---------------------------------------------------------------------------------------
bool gzip_member(const std::string &in, int level, std::string &out)
{
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (deflateInit2(&zs, level, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) return false;
    out.resize(deflateBound(&zs, (uLong)in.size()) + 64);
    zs.next_in = (Bytef *)in.data();
    zs.avail_in = (uInt)in.size();
    zs.next_out = (Bytef *)&out[0];
    zs.avail_out = (uInt)out.size();
    const int rc = deflate(&zs, Z_FINISH);
    out.resize(zs.total_out);
    deflateEnd(&zs);
    return rc == Z_STREAM_END;
}

class Pool {
public:
    explicit Pool(int n)
    {
        for (int i = 0; i < std::max(1, n); i++)
            threads_.emplace_back([this] {
                for (;;) {
                    std::function<void()> job;
                    {
                        std::unique_lock<std::mutex> lk(m_);
                        cv_.wait(lk, [&] { return stop_ || !jobs_.empty(); });
                        if (jobs_.empty()) return;
                        job = std::move(jobs_.front());
                        jobs_.pop_front();
                    }
                    job();
                }
            });
    }
    ~Pool()
    {
        {
            std::unique_lock<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : threads_) t.join();
    }
    void submit(std::function<void()> job)
    {
        {
            std::unique_lock<std::mutex> lk(m_);
            jobs_.push_back(std::move(job));
        }
        cv_.notify_one();
    }

private:
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<std::function<void()>> jobs_;
    std::vector<std::thread> threads_;
    bool stop_ = false;
};

struct GzFileWriter {
    int fd = -1;
    int level = 9;
    bool fastq_encoder = false;  // try the FASTQ-aware encoder first, zlib where it declines
    Pool *pool = nullptr;
    std::string pending;  // uncompressed text not yet dispatched
    std::deque<std::shared_future<std::shared_ptr<std::string>>> inflight;
    uint64_t raw_bytes = 0, gz_bytes = 0;
    bool any_member = false, failed = false;
    static constexpr size_t kChunk = 4 << 20;

    bool open(const std::string &path)
    {
        fd = ::open(path.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
        return fd >= 0;
    }
    void dispatch(std::string &&text)
    {
        raw_bytes += text.size();
        auto in = std::make_shared<std::string>(std::move(text));
        auto prom = std::make_shared<std::promise<std::shared_ptr<std::string>>>();
        inflight.push_back(prom->get_future().share());
        const int lvl = level;
        const bool fq = fastq_encoder;
        pool->submit([in, prom, lvl, fq] {
            auto out = std::make_shared<std::string>();
            if (!(fq && tb::fastq_gzip_member(in->data(), in->size(), *out, true)) && !gzip_member(*in, lvl, *out))
                out->clear();
            prom->set_value(out);
        });
        any_member = true;
    }
    size_t max_inflight = 64;  // back-pressure: at most this many chunks being compressed
    void append(const char *p, size_t n)
    {
        if (pending.capacity() < kChunk) pending.reserve(kChunk + (kChunk >> 2));  // no regrowth copies
        pending.append(p, n);
        if (pending.size() >= kChunk) {
            dispatch(std::move(pending));
            pending.clear();
            while (inflight.size() > max_inflight) {
                inflight.front().wait();
                drain(false);
            }
        }
    }
    // write whatever has finished compressing, in order; block for all if `all`
    void drain(bool all)
    {
        while (!inflight.empty()) {
            auto &f = inflight.front();
            if (!all && f.wait_for(std::chrono::seconds(0)) != std::future_status::ready) break;
            auto out = f.get();
            if (out->empty()) failed = true;
            size_t off = 0;
            while (off < out->size()) {
                ssize_t w = ::write(fd, out->data() + off, out->size() - off);
                if (w <= 0) { failed = true; break; }
                off += (size_t)w;
            }
            gz_bytes += out->size();
            inflight.pop_front();
        }
    }
    void finish()
    {
        if (!pending.empty() || !any_member) {
            // Python always produces a valid (possibly empty) gzip file
            dispatch(std::move(pending));
            pending.clear();
        }
        drain(true);
        if (fd >= 0) ::close(fd);
        fd = -1;
    }
};

}  // namespace

extern "C" {

int tb_classify_fastx(const char *reads_path, const hash_set *A, const hash_set *B,
                      const char *out_a, const char *out_b, const char *out_u, int gzip_level,
                      int tsv_fd, int64_t batch_bases, int n_threads, tb_stream_stats *stats)
{
    const auto t_all = Clock::now();
    if (!reads_path || !A || !B || !out_a || !out_b || !out_u) return -1;
    if (batch_bases < 1) batch_bases = 1;
    if (gzip_level < 0 || gzip_level > 9) gzip_level = 9;
    int threads = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    if (threads < 2) threads = 2;
    const int n_slots = 2, n_batches = n_slots + 2;

    Reader reader;
    reader.batch_bases = batch_bases;
    // a quarter of the host threads may decode BGZF blocks; the parser takes ~2 GB/s, which four
    // to six decoders fill
    reader.src.set_threads(std::max(1, std::min(6, threads / 4)));
    if (!reader.src.open(reads_path, reader.error)) {
        fprintf(stderr, "tb_classify_fastx: %s\n", reader.error.c_str());
        return -2;
    }
    BatchQueue free_q, full_q;
    reader.free_q = &free_q;
    reader.full_q = &full_q;
    std::vector<std::unique_ptr<Batch>> batches;
    for (int i = 0; i < n_batches; i++) {
        batches.emplace_back(new Batch());
        if (!batches.back()->grow_bases(std::min<int64_t>(batch_bases, 64 << 20) + (1 << 20))) return -3;
        free_q.push(batches.back().get());
    }

    Pool pool(std::max(1, threads - 1));
    GzFileWriter wa, wb, wu;
    // TB_FASTQ_GZIP=1 (CLI: --fastq-gzip): chunks of long-line text go through the FASTQ-aware
    // encoder (fastq_deflate.cpp; smaller than zlib -9 on long reads at a fraction of the time,
    // every member decoded and compared before it is written), everything else through zlib
    const char *fq_env = getenv("TB_FASTQ_GZIP");
    const bool fastq_encoder = fq_env && *fq_env && strcmp(fq_env, "0") != 0;
    for (auto *w : { &wa, &wb, &wu }) {
        w->pool = &pool;
        w->level = gzip_level;
        w->fastq_encoder = fastq_encoder;
    }
    if (!wa.open(out_a) || !wb.open(out_b) || !wu.open(out_u)) {
        fprintf(stderr, "tb_classify_fastx: cannot create output files\n");
        return -4;
    }

    std::thread reader_thread([&] { reader.run(); });

    tb_classifier *clf = nullptr;
    int64_t clf_bases = 0, clf_reads = 0;
    std::deque<Batch *> inflight;
    std::string tsv, rec_text;
    double gpu_wait_s = 0, write_s = 0;
    uint64_t n_a = 0, n_b = 0, n_u = 0, n_batches_done = 0;
    int rc = 0;
    int next_slot = 0;

    auto finish_batch = [&](Batch *b) {
        auto t0 = Clock::now();
        if (b->n() > 0 && tb_classifier_wait(clf, b->slot) != 0) rc = -5;
        gpu_wait_s += seconds_since(t0);
        t0 = Clock::now();
        tsv.clear();
        for (int64_t i = 0; i < b->n() && rc == 0; i++) {
            const Rec &r = b->recs[(size_t)i];
            const char *name = b->text.data() + r.name_off;
            const char *seq = b->bases + b->offsets[(size_t)i];
            const size_t seq_len = (size_t)(b->offsets[(size_t)i + 1] - b->offsets[(size_t)i]);
            // Read.__str__ (seq.py:27-31): FASTQ only when the quality string is non-empty
            rec_text.clear();
            if (r.has_qual && r.qual_len > 0) {
                rec_text += '@';
                rec_text.append(name, r.name_len);
                rec_text += '\n';
                rec_text.append(seq, seq_len);
                rec_text += "\n+\n";
                rec_text.append(b->text.data() + r.qual_off, r.qual_len);
                rec_text += '\n';
            } else {
                rec_text += '>';
                rec_text.append(name, r.name_len);
                rec_text += '\n';
                rec_text.append(seq, seq_len);
                rec_text += '\n';
            }
            const char bin = b->bin[i];
            GzFileWriter &w = bin == 'A' ? wa : (bin == 'B' ? wb : wu);
            (bin == 'A' ? n_a : (bin == 'B' ? n_b : n_u))++;
            w.append(rec_text.data(), rec_text.size());
            tsv.append(name, r.name_len);
            tsv += '\t';
            tsv += bin;
            tsv += '\t';
            py_float_repr(b->sa[i], tsv);
            tsv += '\t';
            py_float_repr(b->sb[i], tsv);
            tsv += '\n';
        }
        if (tsv_fd >= 0) {
            size_t off = 0;
            while (off < tsv.size()) {
                ssize_t w = ::write(tsv_fd, tsv.data() + off, tsv.size() - off);
                if (w <= 0) { rc = -6; break; }
                off += (size_t)w;
            }
        }
        wa.drain(false);
        wb.drain(false);
        wu.drain(false);
        write_s += seconds_since(t0);
        n_batches_done++;
        free_q.push(b);
    };

    bool done = false;
    while (!done) {
        Batch *b = full_q.pop();
        done = b->last;
        if (rc == 0 && b->n() > 0) {
            if (!clf || b->total() > clf_bases || b->n() > clf_reads) {
                // (re)create the classifier with room to spare; drain what is in flight first
                while (!inflight.empty()) {
                    finish_batch(inflight.front());
                    inflight.pop_front();
                }
                if (clf) tb_classifier_destroy(clf);
                clf_bases = std::max<int64_t>(b->total() + b->total() / 4, batch_bases + (8 << 20));
                clf_reads = std::max<int64_t>(b->n() + b->n() / 4, 1 << 16);
                clf = tb_classifier_create(A, B, clf_bases, clf_reads, n_slots);
                if (!clf) {
                    fprintf(stderr, "tb_classify_fastx: %s\n", tb_last_error());
                    rc = -7;
                }
            }
            if (rc == 0 && (int)inflight.size() == n_slots) {
                finish_batch(inflight.front());
                inflight.pop_front();
            }
            if (rc == 0) {
                if (!b->ensure_results(b->n())) rc = -3;
            }
            if (rc == 0) {
                b->slot = next_slot;
                next_slot = (next_slot + 1) % n_slots;
                if (tb_classifier_submit(clf, b->slot, b->bases, b->offsets.data(), b->n(), b->ca, b->cb,
                                         b->sa, b->sb, b->bin) != 0) {
                    fprintf(stderr, "tb_classify_fastx: %s\n", tb_last_error());
                    rc = -8;
                }
            }
        }
        if (rc == 0 && b->n() > 0)
            inflight.push_back(b);
        else
            free_q.push(b);
    }
    while (!inflight.empty()) {
        if (rc == 0)
            finish_batch(inflight.front());
        else
            free_q.push(inflight.front());
        inflight.pop_front();
    }
    reader_thread.join();
    const auto t_fin = Clock::now();
    wa.finish();
    wb.finish();
    wu.finish();
    write_s += seconds_since(t_fin);
    if (wa.failed || wb.failed || wu.failed) rc = rc ? rc : -9;
    if (clf) tb_classifier_destroy(clf);
    if (!reader.error.empty()) {
        fprintf(stderr, "tb_classify_fastx: %s\n", reader.error.c_str());
        if (!rc) tb_internal_set_error(reader.error.c_str());
        rc = rc ? rc : -10;
    }
    if (stats) {
        stats->n_reads = reader.n_reads;
        stats->n_bases = reader.n_bases;
        stats->n_batches = n_batches_done;
        stats->input_bytes = reader.src.bytes_in();
        stats->reads_a = n_a;
        stats->reads_b = n_b;
        stats->reads_u = n_u;
        stats->output_raw_bytes = wa.raw_bytes + wb.raw_bytes + wu.raw_bytes;
        stats->output_gz_bytes = wa.gz_bytes + wb.gz_bytes + wu.gz_bytes;
        stats->seconds_total = seconds_since(t_all);
        stats->seconds_parse = reader.seconds;
        stats->seconds_gpu_wait = gpu_wait_s;
        stats->seconds_write = write_s;
    }
    return rc;
}

// exposed for tests: Python's repr(float) as implemented above
int tb_float_repr(double x, char *buf, int buf_len)
{
    std::string s;
    py_float_repr(x, s);
    if ((int)s.size() + 1 > buf_len) return -1;
    memcpy(buf, s.c_str(), s.size() + 1);
    return (int)s.size();
}

// exposed for tests: parse a FASTA/FASTQ file with the native reader and write, for every
// record, "name\tseq\tqual-or-'-'<None>'\n" to out_fd.  Returns the number of records.
int64_t tb_fastx_dump(const char *reads_path, int out_fd)
{
    Reader reader;
    reader.batch_bases = 1 << 20;
    if (!reader.src.open(reads_path, reader.error)) return -1;
    BatchQueue free_q, full_q;
    reader.free_q = &free_q;
    reader.full_q = &full_q;
    std::vector<std::unique_ptr<Batch>> batches;
    for (int i = 0; i < 2; i++) {
        batches.emplace_back(new Batch());
        batches.back()->pinned = false;  // parser only: no GPU, no page-locked memory
        if (!batches.back()->grow_bases(1 << 20)) return -3;
        free_q.push(batches.back().get());
    }
    std::thread t([&] { reader.run(); });
    int64_t n = 0;
    std::string out;
    for (;;) {
        Batch *b = full_q.pop();
        out.clear();
        for (int64_t i = 0; i < b->n(); i++) {
            // three length-prefixed fields (int64 little-endian; -1 = None): names keep tabs, and
            // carriage returns or tabs may sit in sequences, so no separator byte is safe
            const Rec &r = b->recs[(size_t)i];
            const int64_t seq_len = b->offsets[(size_t)i + 1] - b->offsets[(size_t)i];
            const int64_t lens[3] = { (int64_t)r.name_len, seq_len, r.has_qual ? (int64_t)r.qual_len : -1 };
            const char *ptrs[3] = { b->text.data() + r.name_off, b->bases + b->offsets[(size_t)i],
                                    b->text.data() + r.qual_off };
            for (int f = 0; f < 3; f++) {
                out.append(reinterpret_cast<const char *>(&lens[f]), 8);
                if (lens[f] > 0) out.append(ptrs[f], (size_t)lens[f]);
            }
            n++;
        }
        size_t off = 0;
        while (off < out.size()) {
            ssize_t w = ::write(out_fd, out.data() + off, out.size() - off);
            if (w <= 0) break;
            off += (size_t)w;
        }
        const bool last = b->last;
        free_q.push(b);
        if (last) break;
    }
    t.join();
    if (!reader.error.empty()) {
        tb_internal_set_error(reader.error.c_str());
        return -2;
    }
    return n;
}


// exposed for tests: copy in_fd to the gzip file out_path through the pipeline's writer (chunks
// compressed in parallel as independent members, written in order), feeding it `piece` bytes at a
// time.  fastq_encoder != 0 tries the FASTQ-aware encoder first.  Returns the number of input
// bytes, or < 0.
int64_t tb_gzip_copy(int in_fd, const char *out_path, int level, int n_threads, int fastq_encoder, int64_t piece)
{
    if (!out_path || piece < 1) return -1;
    if (level < 0 || level > 9) level = 6;
    Pool pool(std::max(1, n_threads));
    GzFileWriter w;
    w.pool = &pool;
    w.level = level;
    w.fastq_encoder = fastq_encoder != 0;
    w.max_inflight = 3;  // exercise the back-pressure path
    if (!w.open(out_path)) return -2;
    std::vector<char> buf((size_t)piece);
    int64_t total = 0;
    for (;;) {
        const ssize_t n = ::read(in_fd, buf.data(), buf.size());
        if (n < 0) {
            if (errno == EINTR) continue;
            return -3;
        }
        if (n == 0) break;
        w.append(buf.data(), (size_t)n);
        w.drain(false);
        total += n;
    }
    w.finish();
    return w.failed ? -4 : total;
}

}  // extern "C"
