// fastx_pipeline.cpp -- native host pipeline around the GPU classifier
// (FASTX input and fastq.gz output at more than Python speed).
//
//   reader thread : file / gzip stream -> lines (universal newlines, like Python text mode)
//                   -> FASTA/FASTQ records with the exact semantics of the reference's
//                   readfq (src/trio_binning/seq.py:45-83, quirk Q7) -> batches whose
//                   sequences are concatenated straight into pinned memory
//   packer thread : batch -> pinned 2-bit arrays (tb_pack_reads): 0.25 byte per base crosses PCIe
//   device workers: one thread per GPU, two slots each, every GPU with its own replica of the
//                   sets: tb_classifier_submit_packed (H2D + kernels + D2H overlap the parsing
//                   and packing of the next batches) -> wait
//   router        : completed batches back into input order -> every record to its bin
//                   (classify_by_kmers.py:104-115) -> TSV line (:117)
//   deflate pool  : each bin's text is cut into chunks that are compressed in parallel as
//                   independent gzip members and written in order (a concatenation of
//                   members is a valid .gz; the DECOMPRESSED bytes equal the reference's
//                   gzip.open(..., "wt") output, seq.py:132-134)
//   plain output  : the reference's --no-gzip-output mode, CPython's buffered text writers
//                   restated so that the two handles the reference opens on hapA's file
//                   (seq.py:129) overwrite each other exactly as they do upstream.
#include <fcntl.h>
#include <errno.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <zlib.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "fastq_deflate.h"
#include "gz_stream.h"

#include <sys/mman.h>
#include <sys/stat.h>

#include <algorithm>
#include <atomic>
#include <map>
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
    // ... and for ONE plain gzip stream cut into spans: its decoder is slower per core than the
    // serial one, so it gets more of them
    void set_gzip_threads(int n) { gzip_threads_ = n < 2 ? 2 : (n > 12 ? 12 : n); }
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
            // a plain gzip stream in a regular file that is worth it: decoded on several cores
            // (TB_PGZIP=0 turns that off; TB_PGZIP_SPAN / TB_PGZIP_MIN: span and smallest file, bytes)
            const char *on = getenv("TB_PGZIP"), *span = getenv("TB_PGZIP_SPAN"), *least = getenv("TB_PGZIP_MIN");
            if (span && atoll(span) >= 1024) span_bytes_ = (uint64_t)atoll(span);
            const uint64_t min_size = least ? (uint64_t)atoll(least) : 16 * span_bytes_;
            struct stat st;
            if (!bgzf_ && !(on && on[0] == '0') && fstat(fd_, &st) == 0 && S_ISREG(st.st_mode) && (uint64_t)st.st_size >= min_size &&
                st.st_size > 64) {
                void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd_, 0);
                if (m != MAP_FAILED) {
                    map_ = (const unsigned char *)m;
                    map_size_ = (uint64_t)st.st_size;
                    pgzip_ok_ = true;
                }
            }
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
        if (map_) munmap((void *)map_, (size_t)map_size_);
        if (fd_ >= 0) ::close(fd_);
    }
    // Next line INCLUDING its terminator translated to '\n' (absent only at end of input), as a
    // view: *data points into the current chunk when the whole line lies inside it (the common
    // case: no copy), else into a spill buffer that collects the pieces.  The view stays valid
    // until the next call.  Returns false at end of input (or on a read error: see error()).
    bool next(const char *&data, size_t &len)
    {
        bool spilled = false;
        line_ascii_ = true;
        for (;;) {
            if (pos_ == len_ && !fill()) {
                if (!spilled || spill_.empty()) return false;
                data = spill_.data();
                len = spill_.size();
                return true;
            }
            line_ascii_ = line_ascii_ && cur_ascii_;
            if (pending_cr_) {
                // a '\r' ended the previous line; swallow the '\n' of a "\r\n" pair
                pending_cr_ = false;
                if (cur_[pos_] == '\n') {
                    pos_++;
                    continue;
                }
            }
            char *p = cur_ + pos_;
            const size_t avail = len_ - pos_;
            // first '\n' or '\r' (two vectorised scans; '\r' is looked for only up to the '\n')
            const char *nl = (const char *)memchr(p, '\n', avail);
            size_t i = nl ? (size_t)(nl - p) : avail;
            if (cur_has_cr_)
                if (const char *cr = (const char *)memchr(p, '\r', i)) i = (size_t)(cr - p);
            pos_ += i;
            if (i < avail) {
                if (p[i] == '\r') {
                    pending_cr_ = true;
                    p[i] = '\n';  // the chunk is ours: translate in place
                }
                pos_++;
                if (!spilled) {
                    data = p;
                    len = i + 1;
                } else {
                    spill_.append(p, i + 1);
                    data = spill_.data();
                    len = spill_.size();
                }
                return true;
            }
            // the line runs past this chunk: collect its pieces
            if (!spilled) {
                spill_.clear();
                spilled = true;
            }
            spill_.append(p, i);
        }
    }
    // true if every byte of the line next() returned last is a character of its own (no UTF-8
    // multi-byte sequences): its character count is its length
    bool line_ascii() const { return line_ascii_; }
    uint64_t bytes_in() const { return bytes_in_; }
    // time spent inside the inflate (summed over the decoding threads; 0 for plain input)
    double inflate_seconds() const { return (double)inflate_ns_.load() * 1e-9; }
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
        // what the parser may skip in this chunk: the search for '\r' line ends and the UTF-8
        // character count (every byte is a character).  Scanned once per chunk -- by the thread
        // that produced it when that thread has time to spare (plain files, BGZF workers), else by
        // the parser when it takes the chunk (a serial inflate is the bottleneck already)
        bool flags_known = false, has_cr = true, ascii = false;
        void scan_flags()
        {
            const char *p = data.data();
            has_cr = memchr(p, '\r', len) != nullptr;
            uint64_t acc = 0;
            size_t i = 0;
            for (; i + 8 <= len; i += 8) {
                uint64_t w;
                memcpy(&w, p + i, 8);
                acc |= w;
            }
            for (; i < len; i++) acc |= (uint64_t)(unsigned char)p[i];
            ascii = (acc & 0x8080808080808080ull) == 0;
            flags_known = true;
        }
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
            c->flags_known = false;
            c->comp.clear();
            c->why.clear();
            return c;
        }
        return std::make_shared<Chunk>();
    }
    void free_chunk(const ChunkPtr &c)
    {
        std::lock_guard<std::mutex> lk(m_);
        free_.push_back(c);
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
    void produce_serial(bool continue_on_fd, bool already_open = false)
    {
        if (gz_ && !already_open) {
            gzs_.open_fd(fd_, !continue_on_fd);
        }
        for (;;) {
            ChunkPtr c = fresh_chunk();
            c->data.resize(kChunkBytes);
            long got;
            if (gz_) {
                const auto t0 = Clock::now();
                got = gzs_.read(c->data.data(), c->data.size());
                inflate_ns_.fetch_add((uint64_t)(seconds_since(t0) * 1e9));
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
            if (!gz_) c->scan_flags();  // read() leaves this thread idle most of the time
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
            const auto t_dec = Clock::now();
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
            inflate_ns_.fetch_add((uint64_t)(seconds_since(t_dec) * 1e9));
            c->len = n;
            c->failed = bad;
            if (!bad) c->scan_flags();
            finish(c);
        }
    }

    // ---- one plain gzip stream on several cores (gz_stream.h: SpanInflate) ----------------------
    // The file is cut into spans of span_bytes_ compressed bytes.  Workers decode spans side by side,
    // each from the first block header it can verify, into 16-bit symbols with markers for the text
    // before the span.  This thread takes the results in file order: a span counts only if it starts
    // at the bit where its predecessor stopped (else it is decoded again from there, now knowing
    // where), its markers are resolved against the 32 KiB before it, members' CRC-32 and ISIZE are
    // checked, and the text goes to the parser.  Whatever is not a member followed by another member
    // (padding, garbage, the end) is left to the serial reader, and so is any failure: it decodes
    // the current member again from its start, skips what was delivered, and reports -- or, should
    // the failure have been this code's, simply carries on.
    struct SpanJob {
        uint64_t start_bit = 0, end_bit = 0;
        bool search = false;
        bool done = false, taken = false;
        tb::SpanResult res;
    };
    bool parallel_gzip() const { return gz_ && !bgzf_ && pgzip_ok_; }

    void work_spans()
    {
        tb::SpanInflate inflate;
        for (;;) {
            SpanJob *job = nullptr;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return stop_ || spans_closed_ || !span_work_.empty(); });
                if (stop_ || (span_work_.empty() && spans_closed_)) return;
                job = span_work_.front();
                span_work_.pop_front();
                job->taken = true;
            }
            const auto t0 = Clock::now();
            try {
                inflate.decode(map_, map_size_, job->start_bit, job->search, job->end_bit, !job->search, job->res);
            } catch (const std::exception &) {
                job->res.found = false;
                job->res.error = "out of memory while decoding";
            }
            inflate_ns_.fetch_add((uint64_t)(seconds_since(t0) * 1e9));
            {
                std::lock_guard<std::mutex> lk(m_);
                job->done = true;
            }
            cv_.notify_all();
        }
    }

    void close_spans()
    {
        {
            std::lock_guard<std::mutex> lk(m_);
            spans_closed_ = true;
            span_work_.clear();
        }
        cv_.notify_all();
    }

    // the serial reader takes the current member again from its start and skips what was delivered
    void pgzip_fallback(uint64_t member_start, uint64_t delivered)
    {
        close_spans();
        if (::lseek(fd_, (off_t)member_start, SEEK_SET) < 0) {
            push_end(true, "cannot seek in the reads file");
            return;
        }
        bgzf_bytes_ = member_start;
        gzs_.open_fd(fd_, member_start == 0);
        std::vector<char> skip(1u << 20);
        while (delivered > 0) {
            const long got = gzs_.read(skip.data(), (size_t)std::min<uint64_t>(delivered, skip.size()));
            if (got <= 0) {
                push_end(true, got < 0 ? "corrupt gzip input: " + gzs_.error() : std::string("corrupt gzip input: stream ends early"));
                return;
            }
            delivered -= (uint64_t)got;
        }
        produce_serial(true, true);
    }

    void produce_pgzip()
    {
        const unsigned char *file = map_;
        const uint64_t size = map_size_;
        constexpr size_t W = tb::SpanResult::kWindow;
        const size_t hdr = tb::SpanInflate::member_header_length(file, (size_t)std::min<uint64_t>(size, 1u << 20));
        if (hdr == 0) {  // whatever this is, the serial reader says it
            close_spans();
            produce_serial(false);
            return;
        }
        const uint64_t S = span_bytes_;
        const uint64_t n_spans = (size - hdr + S - 1) / S;
        const size_t ahead = 2 * (size_t)gzip_threads_ + 1;
        auto &ring = span_ring_;  // a member: workers may still be writing into a job when this function returns
        for (size_t i = 0; i < ahead; i++) ring.emplace_back(new SpanJob());
        auto issue = [&](uint64_t k) {
            SpanJob *j = ring[k % ahead].get();
            j->search = k > 0;
            j->start_bit = (hdr + k * S) * 8;
            j->end_bit = (hdr + (k + 1) * S) * 8;
            j->done = j->taken = false;
            std::lock_guard<std::mutex> lk(m_);
            span_work_.push_back(j);
        };
        uint64_t issued = 0;
        std::vector<unsigned char> window(W, 0);
        uint64_t expected = hdr * 8;      // where the next span must start
        bool expected_fresh = true;
        uint64_t member_start = 0, member_delivered = 0, member_len = 0;
        uint32_t crc = 0;
        int bad_streak = 0;
        bool speculate = true;
        uint64_t n_accepted = 0, n_redone = 0, n_covered = 0;
        double resolve_s = 0, wait_s = 0;
        const bool debug = getenv("TB_PGZIP_DEBUG") != nullptr;
        auto report = [&](const char *how) {
            if (debug)
                fprintf(stderr, "pgzip: %llu spans of %llu bytes on %d threads: %llu as guessed, %llu decoded again, %llu covered "
                        "by a predecessor; this thread: %.2f s resolving, %.2f s waiting for spans; %s\n", (unsigned long long)n_spans,
                        (unsigned long long)S, gzip_threads_, (unsigned long long)n_accepted, (unsigned long long)n_redone,
                        (unsigned long long)n_covered, resolve_s, wait_s, how);
        };
        tb::SpanInflate own;               // for spans that have to be decoded again, in order
        SpanJob own_job;

        for (uint64_t k = 0; k < n_spans; k++) {
            while (speculate && issued < n_spans && issued < k + ahead) issue(issued++);
            cv_.notify_all();
            const uint64_t end_bit = (hdr + (k + 1) * S) * 8;
            if (expected >= end_bit && k + 1 < n_spans) {
                // the predecessor's last block reached past this whole span
                if (speculate) {
                    std::unique_lock<std::mutex> lk(m_);
                    SpanJob *j = ring[k % ahead].get();
                    cv_.wait(lk, [&] { return stop_ || j->done; });
                    if (stop_) return;
                }
                n_covered++;
                continue;
            }
            SpanJob *job = nullptr;
            if (speculate) {
                SpanJob *j = ring[k % ahead].get();
                {
                    const auto tw = Clock::now();
                    std::unique_lock<std::mutex> lk(m_);
                    cv_.wait(lk, [&] { return stop_ || j->done; });
                    if (stop_) return;
                    wait_s += seconds_since(tw);
                }
                if (k == 0 || (j->res.found && j->res.start_bit == expected)) {
                    job = j;
                    bad_streak = 0;
                    n_accepted++;
                } else if (++bad_streak >= 3) {
                    // no dynamic blocks to be found (stored data?) or a format this search does not
                    // see through: stop guessing, decode the rest in order on this thread
                    speculate = false;
                    close_spans();
                }
            }
            if (!job) {
                const auto t0 = Clock::now();
                try {
                    own.decode(file, size, expected, false, end_bit, expected_fresh, own_job.res);
                } catch (const std::exception &) {
                    own_job.res.error = "out of memory while decoding";
                }
                inflate_ns_.fetch_add((uint64_t)(seconds_since(t0) * 1e9));
                job = &own_job;
                n_redone++;
            }
            tb::SpanResult *res = &job->res;
            if (!res->error.empty()) {
                report(("serial reader takes the member again: " + res->error).c_str());
                pgzip_fallback(member_start, member_delivered);
                return;
            }
            // ---- the markers the worker left open get their bytes from the window; then, member by
            // ---- member, CRC-32 and ISIZE as the serial reader checks them
            const auto t0 = Clock::now();
            const size_t n = res->n_out;
            bool ok = true, has_cr = res->has_cr, ascii = res->ascii;
            unsigned char *out = (unsigned char *)res->bytes.data();
            for (const uint32_t pos : res->marker_pos) {
                const unsigned m = res->sym[W + pos] - 256u;
                if (m >= W) { ok = false; break; }
                const unsigned char b = window[m];
                out[pos] = b;
                has_cr = has_cr || b == '\r';
                ascii = ascii && b < 0x80;
            }
            size_t from = 0;
            uint64_t next_member_start = member_start, delivered_after = member_delivered + n;
            for (size_t e = 0; ok && e < res->ends.size(); e++) {
                const size_t to = res->ends[e].pos;
                crc = tb::crc32_any(crc, out + from, to - from);
                member_len += to - from;
                if (crc != res->ends[e].crc || (uint32_t)member_len != res->ends[e].isize) ok = false;
                crc = 0;
                member_len = 0;
                from = to;
                next_member_start = res->ends[e].next_off;
                delivered_after = n - to;
            }
            if (!ok) {
                report("serial reader takes the member again: checksum, length or marker out of range");
                pgzip_fallback(member_start, member_delivered);
                return;
            }
            crc = tb::crc32_any(crc, out + from, n - from);
            member_len += n - from;
            member_start = next_member_start;
            member_delivered = delivered_after;
            if (n >= W) {
                memcpy(window.data(), out + n - W, W);
            } else if (n > 0) {
                memmove(window.data(), window.data() + n, W - n);
                memcpy(window.data() + W - n, out, n);
            }
            expected = res->stop_bit;
            expected_fresh = res->stop_fresh;
            bytes_in_ = expected / 8;
            resolve_s += seconds_since(t0);
            inflate_ns_.fetch_add((uint64_t)(seconds_since(t0) * 1e9));
            if (n > 0) {
                ChunkPtr c = fresh_chunk();
                c->data.swap(res->bytes);  // the chunk's old buffer serves a later span
                c->len = n;
                c->has_cr = has_cr;
                c->ascii = ascii;
                c->flags_known = true;
                c->done = true;
                if (!enqueue(c, false)) return;
            }
            if (res->tail) {
                // padding, garbage or the end of the file: the serial reader's rules apply
                close_spans();
                if (::lseek(fd_, (off_t)res->tail_off, SEEK_SET) < 0) {
                    push_end(true, "cannot seek in the reads file");
                    return;
                }
                bgzf_bytes_ = res->tail_off;  // compressed bytes before the serial reader's
                report("last member done, the serial reader looks at what follows");
                produce_serial(true);
                return;
            }
        }
        // the spans ran out inside a member: truncated input -- the serial reader reports it
        report("input ends inside a member: serial reader takes it again");
        pgzip_fallback(member_start, member_delivered);
    }

    void start()
    {
        started_ = true;
        if (parallel_gzip()) {
            max_in_flight_ = 3;  // spans are tens of megabytes of text each
            for (int i = 0; i < gzip_threads_; i++) workers_.emplace_back([this] { work_spans(); });
            producer_ = std::thread([this] { produce_pgzip(); });
        } else if (parallel_bgzf()) {
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
            lk.unlock();
            if (!c->flags_known) c->scan_flags();
            cur_has_cr_ = c->has_cr;
            cur_ascii_ = c->ascii;
            return true;
        }
    }

    bool gz_ = false, bgzf_ = false, eof_ = false, pending_cr_ = false, started_ = false, stop_ = false;
    int decode_threads_ = 2, gzip_threads_ = 4;
    tb::GzStream gzs_;
    int fd_ = -1;
    std::thread producer_;
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<ChunkPtr> ordered_, work_;
    std::deque<SpanJob *> span_work_;
    std::vector<std::unique_ptr<SpanJob>> span_ring_;
    bool spans_closed_ = false, pgzip_ok_ = false;
    const unsigned char *map_ = nullptr;
    uint64_t map_size_ = 0, span_bytes_ = 1u << 20;
    std::vector<ChunkPtr> free_;
    size_t max_in_flight_ = 3;
    ChunkPtr cur_chunk_;
    std::string spill_;  // a line that straddles chunks
    bool cur_has_cr_ = true, cur_ascii_ = false, line_ascii_ = false;
    char *cur_ = nullptr;
    size_t pos_ = 0, len_ = 0;
    uint64_t bytes_in_ = 0, bgzf_bytes_ = 0;
    std::atomic<uint64_t> inflate_ns_{ 0 };
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
inline size_t chop_len(const char *line, size_t n)
{
    if (n == 0) return 0;
    n--;
    while (n > 0 && ((unsigned char)line[n] & 0xC0) == 0x80) n--;
    return n;
}

// ---------------------------------------------------------------------------------------
// batches
// ---------------------------------------------------------------------------------------
// A record's sequence and quality string (tens of kilobytes for long reads) are copied from the
// input chunk, which is in cache, into a batch that is far larger than any cache and is read again
// only much later, by other threads.  Ordinary stores would first pull every destination line into
// the cache (read for ownership) and write it back later: three bytes of memory traffic per byte
// copied.  Streaming stores make it two, which is worth a third of the copy's time on the one
// thread that parses (measured: 7.0 -> 9.2 GB/s for 25-kB pieces).
#if defined(__x86_64__)
__attribute__((target("avx2"))) inline void copy_streaming_avx2(char *dst, const char *src, size_t n)
{
    size_t head = (32 - ((uintptr_t)dst & 31)) & 31;
    if (head > n) head = n;
    memcpy(dst, src, head);
    dst += head;
    src += head;
    n -= head;
    size_t i = 0;
    for (; i + 128 <= n; i += 128) {
        const __m256i a = _mm256_loadu_si256((const __m256i *)(src + i)), b = _mm256_loadu_si256((const __m256i *)(src + i + 32)),
                      c = _mm256_loadu_si256((const __m256i *)(src + i + 64)), d = _mm256_loadu_si256((const __m256i *)(src + i + 96));
        _mm256_stream_si256((__m256i *)(dst + i), a);
        _mm256_stream_si256((__m256i *)(dst + i + 32), b);
        _mm256_stream_si256((__m256i *)(dst + i + 64), c);
        _mm256_stream_si256((__m256i *)(dst + i + 96), d);
    }
    memcpy(dst + i, src + i, n - i);
    _mm_sfence();
}
#endif
inline void copy_to_batch(char *dst, const char *src, size_t n)
{
#if defined(__x86_64__)
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2 && n >= 4096) {
        copy_streaming_avx2(dst, src, n);
        return;
    }
#endif
    memcpy(dst, src, n);
}

// names and quality strings of a batch: a plain growing byte buffer (std::string would zero-fill
// on resize, and its append cannot use copy_to_batch)
struct TextBuf {
    char *p = nullptr;
    size_t n = 0, cap = 0;
    TextBuf() = default;
    TextBuf(const TextBuf &) = delete;
    TextBuf &operator=(const TextBuf &) = delete;
    ~TextBuf() { free(p); }
    const char *data() const { return p; }
    size_t size() const { return n; }
    void clear() { n = 0; }
    bool reserve_more(size_t extra)
    {
        if (n + extra <= cap) return true;
        size_t want = cap ? cap : (1u << 20);
        while (want < n + extra) want += want / 2;
        char *q = (char *)realloc(p, want);
        if (!q) return false;
        p = q;
        cap = want;
        return true;
    }
    bool append(const char *src, size_t len)
    {
        if (!reserve_more(len)) return false;
        copy_to_batch(p + n, src, len);
        n += len;
        return true;
    }
};

struct Rec {
    uint64_t name_off;
    uint32_t name_len;
    uint64_t qual_off;
    uint32_t qual_len;
    bool has_qual;  // qual is not None (it may still be empty)
};

struct Batch {
    bool pinned = true;     // sequences in page-locked memory, or plain malloc (they stay on the host)
    bool pinned_results = true;
    void *owner = nullptr;  // the pipeline's wrapper around this batch
    uint64_t first_record = 0;  // index of the batch's first record in the input
    char *bases = nullptr;
    int64_t cap_bases = 0;
    std::vector<int64_t> offsets;  // n+1
    std::vector<Rec> recs;
    TextBuf text;      // names and quality strings
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
    void *alloc(uint64_t n, bool pin) { return pin ? tb_host_alloc(n) : malloc(n ? n : 1); }
    void release(void *p, bool pin)
    {
        if (!p) return;
        if (pin) tb_host_free(p); else free(p);
    }
    // make room for `need` bytes, keeping the first `used` bytes
    bool grow_bases(int64_t need, int64_t used = 0)
    {
        int64_t cap = std::max<int64_t>(need + need / 4, 1 << 20);
        char *nb = (char *)alloc((uint64_t)cap, pinned);
        if (!nb) return false;
        if (bases) {
            memcpy(nb, bases, (size_t)used);
            release(bases, pinned);
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
        const bool pin = pinned_results;
        ca = (int32_t *)alloc((uint64_t)cap_reads * 4, pin);
        cb = (int32_t *)alloc((uint64_t)cap_reads * 4, pin);
        sa = (double *)alloc((uint64_t)cap_reads * 8, pin);
        sb = (double *)alloc((uint64_t)cap_reads * 8, pin);
        bin = (char *)alloc((uint64_t)cap_reads, pin);
        return ca && cb && sa && sb && bin;
    }
    void free_results()
    {
        release(ca, pinned_results);
        release(cb, pinned_results);
        release(sa, pinned_results);
        release(sb, pinned_results);
        release(bin, pinned_results);
        ca = cb = nullptr;
        sa = sb = nullptr;
        bin = nullptr;
        cap_reads = 0;
    }
    ~Batch()
    {
        free_results();
        release(bases, pinned);
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
        cur->first_record = 0;
        int64_t seq_start = 0;  // where the current record's sequence starts in cur->bases
        int64_t fill = 0;       // bytes of cur->bases in use (incl. the record being read)

        const char *lp = nullptr;  // the current line (view, valid until the next src.next)
        size_t ln = 0;
        std::string header;
        bool have_header = false;  // a pending header line (already chopped, non-empty)
        bool ok = true;
        // append a record whose name (and, for FASTQ, quality) already sit in cur->text
        auto emit = [&](uint64_t name_off, uint32_t name_len, int64_t seq_len, bool has_qual, uint64_t qual_off,
                        uint32_t qual_len) {
            Rec r;
            r.name_off = name_off;
            r.name_len = name_len;
            r.qual_off = qual_off;
            r.qual_len = qual_len;
            r.has_qual = has_qual;
            cur->recs.push_back(r);
            cur->offsets.push_back(seq_start + seq_len);
            n_reads++;
            n_bases += (uint64_t)seq_len;
            if (cur->total() >= batch_bases) {
                full_q->push(cur);
                cur = free_q->pop();
                cur->reset();
                cur->first_record = n_reads;
                fill = 0;
            }
        };
        while (ok) {
            // ---- find the next header unless one is pending
            if (!have_header) {
                bool found = false;
                while (src.next(lp, ln)) {
                    if (lp[0] == '>' || lp[0] == '@') {
                        header.assign(lp, chop_len(lp, ln));
                        found = true;
                        break;
                    }
                }
                if (!found || header.empty()) break;
            }
            have_header = false;
            const uint64_t name_off = cur->text.size();
            {
                const size_t sp = header.find(' ', 1);
                const size_t name_end = sp == std::string::npos ? header.size() : sp;
                if (!cur->text.append(header.data() + 1, name_end - 1)) { error = "allocation of a batch failed"; break; }
            }
            const uint32_t name_len = (uint32_t)(cur->text.size() - name_off);
            // ---- sequence lines until a line starting with '@', '+' or '>'
            seq_start = fill = cur->total();
            bool got_marker = false, seq_ascii = true;
            while (src.next(lp, ln)) {
                const char c = lp[0];
                if (c == '@' || c == '+' || c == '>') {
                    header.assign(lp, chop_len(lp, ln));
                    got_marker = true;
                    break;
                }
                const size_t n = chop_len(lp, ln);
                if (fill + (int64_t)n > cur->cap_bases) {
                    if (!cur->grow_bases(fill + (int64_t)n, fill)) { error = "allocation of a batch failed"; ok = false; break; }
                }
                copy_to_batch(cur->bases + fill, lp, n);
                fill += (int64_t)n;
                seq_ascii = seq_ascii && src.line_ascii();
            }
            if (!ok) break;
            const int64_t seq_len = fill - seq_start;
            const bool pending = got_marker && !header.empty();
            if (!pending || header[0] != '+') {
                // FASTA record (or end of input)
                emit(name_off, name_len, seq_len, false, cur->text.size(), 0);
                if (!pending) break;
                have_header = true;
                continue;
            }
            // ---- FASTQ: quality lines, straight into the batch's text, until at least len(seq)
            // ---- CHARACTERS were read
            const size_t seq_chars = seq_ascii ? (size_t)seq_len : n_chars(cur->bases + seq_start, (size_t)seq_len);
            const uint64_t qual_off = cur->text.size();
            size_t have = 0;
            bool complete = false;
            while (src.next(lp, ln)) {
                const size_t n = chop_len(lp, ln);
                if (!cur->text.append(lp, n)) { error = "allocation of a batch failed"; ok = false; break; }
                have += (src.line_ascii() ? ln : n_chars(lp, ln)) - 1;
                if (have >= seq_chars) {
                    complete = true;
                    break;
                }
            }
            if (complete) {
                emit(name_off, name_len, seq_len, true, qual_off, (uint32_t)(cur->text.size() - qual_off));
            } else {
                // input ended inside the quality block: record without qualities, then stop
                // (the partial quality text stays behind in the batch, unreferenced)
                emit(name_off, name_len, seq_len, false, cur->text.size(), 0);
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

// One chunk of text as BGZF (SAM spec 4.1; what bgzip/htslib write): independent gzip members of at
// most 64 KiB of text, each with a 'BC' extra subfield that holds its compressed size - 1.  Used to
// make BGZF test and benchmark inputs (the reader decodes such files on several threads).
bool bgzf_blocks(const std::string &in, int level, std::string &out)
{
    out.clear();
    std::vector<unsigned char> buf(70000);
    for (size_t off = 0; off < in.size() || off == 0; off += 65280) {
        const size_t n = std::min<size_t>(65280, in.size() - off);
        z_stream zs;
        memset(&zs, 0, sizeof zs);
        if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return false;
        zs.next_in = (Bytef *)in.data() + off;
        zs.avail_in = (uInt)n;
        zs.next_out = buf.data();
        zs.avail_out = (uInt)buf.size();
        const int rc = deflate(&zs, Z_FINISH);
        const size_t clen = zs.total_out;
        deflateEnd(&zs);
        if (rc != Z_STREAM_END || clen + 26 > 65536) return false;
        const size_t total = 18 + clen + 8;
        const unsigned char head[18] = { 0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0,
                                         (unsigned char)((total - 1) & 0xff), (unsigned char)((total - 1) >> 8) };
        out.append((const char *)head, 18);
        out.append((const char *)buf.data(), clen);
        const uint32_t crc = (uint32_t)crc32(0L, (const Bytef *)in.data() + off, (uInt)n), isize = (uint32_t)n;
        for (uint32_t v : { crc, isize })
            for (int b = 0; b < 4; b++) out.push_back((char)((v >> (8 * b)) & 0xff));
        if (in.empty()) break;
    }
    return true;
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

// ---------------------------------------------------------------------------------------
// bin writers.  record() receives str(read) (seq.py:27-31) WITHOUT the newline print() adds:
// Read.print() is print(self, file=f), i.e. f.write(str(self)) followed by f.write("\n") --
// two writes, which matters for the plain-text writer below.
// ---------------------------------------------------------------------------------------
struct BinWriter {
    virtual ~BinWriter() {}
    virtual bool open(const std::string &path) = 0;
    virtual void record(const char *text, size_t n) = 0;
    // the same record given as pieces (a writer that only collects bytes need not see them joined)
    virtual void record_parts(const char *const *parts, const size_t *lens, int n_parts)
    {
        joined_.clear();
        for (int i = 0; i < n_parts; i++) joined_.append(parts[i], lens[i]);
        record(joined_.data(), joined_.size());
    }
    virtual void drain(bool all) = 0;
    virtual void finish() = 0;
    uint64_t raw_bytes = 0, out_bytes = 0;
    bool failed = false;
    int fail_errno = 0;

private:
    std::string joined_;
};

struct GzFileWriter : BinWriter {
    int fd = -1;
    int level = 9;
    bool fastq_encoder = false;  // try the FASTQ-aware encoder first, zlib where it declines
    bool bgzf = false;           // write BGZF blocks (input generator for tests and benchmarks)
    Pool *pool = nullptr;
    std::atomic<uint64_t> *deflate_ns = nullptr;  // CPU time spent compressing, summed over the pool
    std::string pending;  // uncompressed text not yet dispatched
    std::deque<std::shared_future<std::shared_ptr<std::string>>> inflight;
    bool any_member = false;
    static constexpr size_t kChunk = 4 << 20;

    bool open(const std::string &path) override
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
        const bool fq = fastq_encoder, bg = bgzf;
        std::atomic<uint64_t> *ns = deflate_ns;
        pool->submit([in, prom, lvl, fq, bg, ns] {
            const auto t0 = Clock::now();
            auto out = std::make_shared<std::string>();
            if (bg) {
                if (!bgzf_blocks(*in, lvl, *out)) out->clear();
            } else if (!(fq && tb::fastq_gzip_member(in->data(), in->size(), *out, true)) && !gzip_member(*in, lvl, *out))
                out->clear();
            if (ns) ns->fetch_add((uint64_t)(seconds_since(t0) * 1e9));
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
    void record(const char *text, size_t n) override
    {
        append(text, n);
        append("\n", 1);
    }
    void record_parts(const char *const *parts, const size_t *lens, int n_parts) override
    {
        // straight into the chunk being filled; a chunk ends only after a whole record
        if (pending.capacity() < kChunk) pending.reserve(kChunk + (kChunk >> 2));
        for (int i = 0; i < n_parts; i++) pending.append(parts[i], lens[i]);
        append("\n", 1);
    }
    // write whatever has finished compressing, in order; block for all if `all`
    void drain(bool all) override
    {
        while (!inflight.empty()) {
            auto &f = inflight.front();
            if (!all && f.wait_for(std::chrono::seconds(0)) != std::future_status::ready) break;
            auto out = f.get();
            if (out->empty()) failed = true;
            size_t off = 0;
            while (off < out->size()) {
                ssize_t w = ::write(fd, out->data() + off, out->size() - off);
                if (w < 0 && errno == EINTR) continue;
                if (w <= 0) { failed = true; fail_errno = errno; break; }
                off += (size_t)w;
            }
            out_bytes += out->size();
            inflight.pop_front();
        }
    }
    void finish() override
    {
        if (!pending.empty() || !any_member) {
            // Python always produces a valid (possibly empty) gzip file
            dispatch(std::move(pending));
            pending.clear();
        }
        drain(true);
        if (fd >= 0 && ::close(fd) != 0) { failed = true; fail_errno = errno; }  // deferred write errors (quota, NFS) surface here
        fd = -1;
    }
};

// The reference's --no-gzip-output mode writes through open(name, "w") (seq.py:127-130): a
// TextIOWrapper over a BufferedWriter over its own file descriptor.  Plain mode opens the B
// handle on A's file name (seq.py:129), so two such objects, each with its own file position,
// write over each other and WHEN each flushes decides the bytes that survive -- the reference's
// own golden file (tests/data/hapA.fastq) is such an overlap.  This is CPython's (3.9 .. 3.12)
// write path restated:
//   TextIOWrapper.write(s): a pending list collects encoded strings and is flushed once it holds
//     >= 8192 bytes (Modules/_io/textio.c, _io_TextIOWrapper_write_impl).  WHEN a list that is
//     not empty is flushed before s joins it depends on the interpreter (text_mode, probed by the
//     Python caller on the running interpreter): 0 = when s alone is >= 8192 bytes (CPython 3.9
//     .. 3.12), 1 = when list + s would exceed 8192 bytes (later versions), 2 = never (3.8);
//   BufferedWriter.write(b): b is copied into the buffer if it fits in what is left of it;
//     otherwise the buffer is written out, then b itself goes straight to the descriptor while
//     more than a buffer's worth remains (one write() takes all of it), and a short rest is
//     buffered (Modules/_io/bufferedio.c, _io_BufferedWriter_write_impl);
//   the buffer size is the file's st_blksize (io.open) unless the caller passes the size its
//     interpreter uses;
//   close() flushes the pending list, then the buffer.
struct PyTextFileWriter : BinWriter {
    int fd = -1;
    static constexpr size_t kTextChunk = 8192;
    std::string pending;
    bool has_pending = false;
    size_t pending_count = 0;
    std::vector<char> buf;
    size_t pos = 0;
    int text_mode = 0;
    size_t buffer_size = 0;  // 0: st_blksize

    bool open(const std::string &path) override
    {
        fd = ::open(path.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0666);
        if (fd < 0) return false;
        struct stat st;
        size_t bs = 8192;
        if (buffer_size > 0) bs = buffer_size;
        else if (fstat(fd, &st) == 0 && st.st_blksize > 1) bs = (size_t)st.st_blksize;
        buf.assign(bs, 0);
        return true;
    }
    void raw_write(const char *p, size_t n)
    {
        out_bytes += n;
        while (n > 0) {
            const ssize_t w = ::write(fd, p, n);
            if (w < 0 && errno == EINTR) continue;
            if (w <= 0) { failed = true; fail_errno = errno; return; }
            p += w;
            n -= (size_t)w;
        }
    }
    void buffer_flush()
    {
        if (pos) raw_write(buf.data(), pos);
        pos = 0;
    }
    void buffered_write(const char *p, size_t n)
    {
        if (n <= buf.size() - pos) {
            memcpy(buf.data() + pos, p, n);
            pos += n;
            return;
        }
        buffer_flush();
        if (n > buf.size()) {
            raw_write(p, n);
            return;
        }
        memcpy(buf.data(), p, n);
        pos = n;
    }
    void text_flush()
    {
        if (!has_pending) return;
        has_pending = false;
        pending_count = 0;
        buffered_write(pending.data(), pending.size());
        pending.clear();
    }
    void text_write(const char *p, size_t n)
    {
        raw_bytes += n;
        if (!has_pending) {
            pending.assign(p, n);
            has_pending = true;
        } else if ((text_mode == 0 && n >= kTextChunk) || (text_mode == 1 && pending_count + n > kTextChunk)) {
            text_flush();
            pending.assign(p, n);
            has_pending = true;
        } else {
            pending.append(p, n);
        }
        pending_count += n;
        if (pending_count >= kTextChunk) text_flush();
    }
    void record(const char *text, size_t n) override
    {
        text_write(text, n);
        text_write("\n", 1);
    }
    void drain(bool) override {}
    void finish() override
    {
        text_flush();
        buffer_flush();
        if (fd >= 0 && ::close(fd) != 0) { failed = true; fail_errno = errno; }
        fd = -1;
    }
};

// ---------------------------------------------------------------------------------------
// the pipeline
//
//   reader (1 thread)  file -> records -> Batch (sequences concatenated in ordinary host memory,
//                      names/qualities beside them), batches numbered in input order
//   packer (1 thread + the tb_pack_reads workers)  Batch -> pinned 2-bit arrays (F, dirty, the
//                      rare fill-forward words, Q3-UB prefixes): what crosses PCIe is 0.25 byte
//                      per base, never the ASCII
//   device workers (1 thread per GPU, 2 slots each)  tb_classifier_submit_packed / wait on the
//                      worker's own replica of the sets; batches are dealt on demand (a worker
//                      takes the next batch when it has a free slot) or strictly round-robin
//   router (calling thread)  completed batches back into input order -> every record to the
//                      writer of its bin + one TSV line (classify_by_kmers.py:104-117)
// ---------------------------------------------------------------------------------------
struct PackedHost {
    tb_packed_batch pb;
    int64_t *offsets = nullptr;
    int64_t cap_bases = -1, cap_reads = -1;
    PackedHost() { memset(&pb, 0, sizeof pb); }
    void release()
    {
        void *ptrs[] = { pb.F, pb.dirty, pb.dirty_idx, pb.e_words, pb.ub_read, pb.ub_skip, pb.ub_off, pb.ub_bytes, offsets };
        for (void *p : ptrs)
            if (p) tb_host_free(p);
        memset(&pb, 0, sizeof pb);
        offsets = nullptr;
        cap_bases = cap_reads = -1;
    }
    bool ensure(int64_t total_bases, int64_t n_reads)
    {
        if (total_bases <= cap_bases && n_reads <= cap_reads) return true;
        const int64_t nb = std::max(total_bases + total_bases / 8, cap_bases);
        const int64_t nr = std::max(n_reads + n_reads / 8 + 64, cap_reads);
        release();
        const int64_t tiles = (nb + 1023) / 1024;
        pb.cap_words = (tiles + 1) * 32;
        pb.cap_tiles = tiles + 1;
        pb.cap_dirty = pb.cap_words / 8 + 1024;  // more non-ACGT than that: the batch goes as ASCII
        pb.cap_ub = nr + 1;
        pb.cap_ub_bytes = std::max<int64_t>(1 << 20, nb / 16);
        pb.F = (uint64_t *)tb_host_alloc((uint64_t)pb.cap_words * 8);
        pb.dirty = (uint32_t *)tb_host_alloc((uint64_t)pb.cap_tiles * 4);
        pb.dirty_idx = (uint32_t *)tb_host_alloc((uint64_t)pb.cap_dirty * 4);
        pb.e_words = (uint64_t *)tb_host_alloc((uint64_t)pb.cap_dirty * 8);
        pb.ub_read = (uint32_t *)tb_host_alloc((uint64_t)pb.cap_ub * 4);
        pb.ub_skip = (uint32_t *)tb_host_alloc((uint64_t)pb.cap_ub * 4);
        pb.ub_off = (int64_t *)tb_host_alloc((uint64_t)(pb.cap_ub + 1) * 8);
        pb.ub_bytes = (unsigned char *)tb_host_alloc((uint64_t)pb.cap_ub_bytes);
        offsets = (int64_t *)tb_host_alloc((uint64_t)(nr + 2) * 8);
        if (!pb.F || !pb.dirty || !pb.dirty_idx || !pb.e_words || !pb.ub_read || !pb.ub_skip || !pb.ub_off ||
            !pb.ub_bytes || !offsets) {
            release();
            return false;
        }
        cap_bases = nb;
        cap_reads = nr;
        return true;
    }
    ~PackedHost() { release(); }
};

struct PipeBatch {
    Batch b;
    PackedHost packed;
    bool packed_ok = false;
    uint64_t seq = 0;
    bool failed = false;
};

template <class T>
class Channel {
public:
    void push(T v)
    {
        {
            std::lock_guard<std::mutex> lk(m_);
            q_.push_back(v);
        }
        cv_.notify_all();
    }
    T pop()
    {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return !q_.empty(); });
        T v = q_.front();
        q_.pop_front();
        return v;
    }
    bool try_pop(T &v)
    {
        std::lock_guard<std::mutex> lk(m_);
        if (q_.empty()) return false;
        v = q_.front();
        q_.pop_front();
        return true;
    }

private:
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<T> q_;
};

struct PipelineConfig {
    const char *reads_path = nullptr;
    std::vector<const hash_set *> sets_a, sets_b;  // one replica per device
    const char *stub_bins = nullptr;               // testing aid: no GPU, record i gets stub_bins[i % n]
    int64_t n_stub = 0;
    bool gzip_output = true;
    int gzip_level = 6;
    bool fastq_encoder = false;
    bool round_robin = false;  // deal batch i to device i mod G instead of on demand
    int py_text_mode = 0;      // plain output: which CPython's TextIOWrapper to restate
    int64_t py_buffer_size = 0;  //             and its BufferedWriter's size (0: st_blksize)
    int tsv_fd = -1;
    int64_t batch_bases = 256 << 20;
    int threads = 0;
    std::string out_a, out_b, out_u;
};

int run_pipeline(const PipelineConfig &cfg, tb_stream_stats *stats)
{
    const auto t_all = Clock::now();
    const int n_dev = cfg.stub_bins ? std::max<int>(1, (int)cfg.sets_a.size()) : (int)cfg.sets_a.size();
    if (n_dev < 1) return -1;
    const int64_t batch_bases = std::max<int64_t>(1, cfg.batch_bases);
    int threads = cfg.threads > 0 ? cfg.threads : (int)std::thread::hardware_concurrency();
    if (threads < 2) threads = 2;
    const int n_slots = 2;
    const int n_batches = n_dev * n_slots + 4;
    const int k = cfg.stub_bins ? 21 : (int)cfg.sets_a[0]->k;

    Reader reader;
    reader.batch_bases = batch_bases;
    // a quarter of the host threads may decode BGZF blocks; the parser takes ~2 GB/s, which four
    // to six decoders fill
    reader.src.set_threads(std::max(1, std::min(6, threads / 4)));
    reader.src.set_gzip_threads(threads / 2);
    if (!reader.src.open(cfg.reads_path, reader.error)) {
        fprintf(stderr, "tb_classify_fastx: %s\n", reader.error.c_str());
        tb_internal_set_error(reader.error.c_str());
        return -2;
    }
    BatchQueue free_q, parsed_q;
    reader.free_q = &free_q;
    reader.full_q = &parsed_q;
    std::vector<std::unique_ptr<PipeBatch>> batches;
    for (int i = 0; i < n_batches; i++) {
        batches.emplace_back(new PipeBatch());
        Batch &b = batches.back()->b;
        b.pinned = false;           // the ASCII stays on the host: ordinary memory
        b.pinned_results = !cfg.stub_bins;
        b.owner = batches.back().get();
        if (!b.grow_bases(std::min<int64_t>(batch_bases, 64 << 20) + (1 << 20))) return -3;
        free_q.push(&b);
    }

    Pool pool(std::max(1, threads - 1));
    std::atomic<uint64_t> deflate_ns{ 0 };
    std::unique_ptr<BinWriter> wa, wb, wu;
    if (cfg.gzip_output) {
        for (auto *slot : { &wa, &wb, &wu }) {
            auto *w = new GzFileWriter();
            w->pool = &pool;
            w->level = cfg.gzip_level;
            w->fastq_encoder = cfg.fastq_encoder;
            w->deflate_ns = &deflate_ns;
            slot->reset(w);
        }
        if (!wa->open(cfg.out_a) || !wb->open(cfg.out_b) || !wu->open(cfg.out_u)) {
            fprintf(stderr, "tb_classify_fastx: cannot create output files\n");
            return -4;
        }
    } else {
        for (auto *slot : { &wa, &wb, &wu }) {
            auto *w = new PyTextFileWriter();
            w->text_mode = cfg.py_text_mode;
            w->buffer_size = cfg.py_buffer_size > 0 ? (size_t)cfg.py_buffer_size : 0;
            slot->reset(w);
        }
        // seq.py:127-130: A, then B ON A's NAME (sic), then U -- each open() truncates
        if (!wa->open(cfg.out_a) || !wb->open(cfg.out_a) || !wu->open(cfg.out_u)) {
            fprintf(stderr, "tb_classify_fastx: cannot create output files\n");
            return -4;
        }
    }

    std::atomic<int> rc{ 0 };
    auto set_rc = [&](int v) { int z = 0; rc.compare_exchange_strong(z, v); };
    std::mutex err_m;
    std::string err_text;
    auto set_err = [&](const char *what) {
        std::lock_guard<std::mutex> lk(err_m);
        if (err_text.empty()) err_text = what ? what : "";
    };

    std::thread reader_thread([&] { reader.run(); });

    // ---- packer --------------------------------------------------------------------------
    // per-device queues in round-robin mode, one shared queue otherwise
    std::vector<std::unique_ptr<Channel<PipeBatch *>>> ready;
    for (int d = 0; d < (cfg.round_robin ? n_dev : 1); d++) ready.emplace_back(new Channel<PipeBatch *>());
    Channel<PipeBatch *> done_q;
    double pack_s = 0;
    const int pack_threads = std::max(1, std::min(16, threads / 2));
    std::thread packer([&] {
        uint64_t seq = 0;
        for (;;) {
            Batch *b = parsed_q.pop();
            PipeBatch *pbatch = static_cast<PipeBatch *>(b->owner);
            pbatch->seq = seq;
            pbatch->failed = false;
            pbatch->packed_ok = false;
            const bool last = b->last;
            if (!cfg.stub_bins && rc.load() == 0 && b->n() > 0) {
                const auto t0 = Clock::now();
                if (!pbatch->packed.ensure(b->total(), b->n())) {
                    set_rc(-3);
                    set_err("pinned allocation for a packed batch failed");
                } else {
                    memcpy(pbatch->packed.offsets, b->offsets.data(), (size_t)(b->n() + 1) * 8);
                    const int prc = tb_pack_reads(b->bases, pbatch->packed.offsets, b->n(), k, pack_threads,
                                                  &pbatch->packed.pb);
                    pbatch->packed_ok = prc == 0;  // -3: too many non-ACGT words -> this batch goes as ASCII
                    if (prc != 0 && prc != -3) {
                        set_rc(-11);
                        set_err("tb_pack_reads failed");
                    }
                }
                pack_s += seconds_since(t0);
            }
            ready[cfg.round_robin ? (size_t)(seq % (uint64_t)n_dev) : 0]->push(pbatch);
            seq++;
            if (last) break;
        }
        for (auto &q : ready)
            for (int d = 0; d < n_dev; d++) q->push(nullptr);  // one end marker per possible consumer
    });

    // ---- device workers --------------------------------------------------------------------
    std::atomic<uint64_t> gpu_wait_ns{ 0 };
    std::vector<int> used((size_t)n_dev, 0);
    auto device_worker = [&](int d) {
        Channel<PipeBatch *> &in = *ready[cfg.round_robin ? (size_t)d : 0];
        tb_classifier *clf = nullptr;
        int64_t clf_bases = 0, clf_reads = 0;
        std::deque<PipeBatch *> inflight;
        int next_slot = 0;
        auto complete = [&](PipeBatch *p) {
            const auto t0 = Clock::now();
            if (!p->failed && tb_classifier_wait(clf, p->b.slot) != 0) {
                set_rc(-5);
                set_err(tb_last_error());
                p->failed = true;
            }
            gpu_wait_ns.fetch_add((uint64_t)(seconds_since(t0) * 1e9));
            done_q.push(p);
        };
        for (;;) {
            // A worker never sits on finished work: with nothing new to submit it completes what it
            // has in flight and hands it to the router, then blocks.  (Holding batches while waiting
            // for input starves the reader of free batches: the router cannot release later batches
            // before the held one, and the pool runs dry.)
            PipeBatch *p = nullptr;
            if (!in.try_pop(p)) {
                if (!inflight.empty()) {
                    complete(inflight.front());
                    inflight.pop_front();
                    continue;
                }
                p = in.pop();
            }
            if (!p) break;
            Batch &b = p->b;
            if (cfg.stub_bins) {
                // testing aid: bins from the caller's pattern, counts from the read lengths
                if (b.n() > 0 && !b.ensure_results(b.n())) set_rc(-3);
                for (int64_t r = 0; r < b.n() && rc.load() == 0; r++) {
                    const int64_t len = b.offsets[(size_t)r + 1] - b.offsets[(size_t)r];
                    b.ca[r] = (int32_t)(len % 7);
                    b.cb[r] = (int32_t)(len % 5);
                    b.sa[r] = b.ca[r] * 1.5;
                    b.sb[r] = b.cb[r] / 3.0;
                    b.bin[r] = cfg.stub_bins[(size_t)((b.first_record + (uint64_t)r) % (uint64_t)cfg.n_stub)];
                }
                used[(size_t)d] = 1;
                done_q.push(p);
                continue;
            }
            if (rc.load() != 0 || b.n() == 0) {
                p->failed = rc.load() != 0;
                // keep the order of completions of this worker: older batches first
                while (!inflight.empty()) {
                    complete(inflight.front());
                    inflight.pop_front();
                }
                done_q.push(p);
                continue;
            }
            used[(size_t)d] = 1;
            if (!clf || b.total() > clf_bases || b.n() > clf_reads) {
                while (!inflight.empty()) {
                    complete(inflight.front());
                    inflight.pop_front();
                }
                if (clf) tb_classifier_destroy(clf);
                clf_bases = std::max<int64_t>(b.total() + b.total() / 4, batch_bases + (8 << 20));
                clf_reads = std::max<int64_t>(b.n() + b.n() / 4, 1 << 16);
                clf = tb_classifier_create(cfg.sets_a[(size_t)d], cfg.sets_b[(size_t)d], clf_bases, clf_reads, n_slots);
                if (!clf) {
                    set_rc(-7);
                    set_err(tb_last_error());
                }
            }
            if (clf && (int)inflight.size() == n_slots) {
                complete(inflight.front());
                inflight.pop_front();
            }
            if (clf && !b.ensure_results(b.n())) {
                set_rc(-3);
                set_err("pinned allocation for the results failed");
            }
            if (clf && rc.load() == 0) {
                b.slot = next_slot;
                next_slot = (next_slot + 1) % n_slots;
                int src;
                if (p->packed_ok)
                    src = tb_classifier_submit_packed(clf, b.slot, &p->packed.pb, p->packed.offsets, b.ca, b.cb, b.sa,
                                                      b.sb, b.bin);
                else
                    src = tb_classifier_submit(clf, b.slot, b.bases, b.offsets.data(), b.n(), b.ca, b.cb, b.sa, b.sb,
                                               b.bin);
                if (src != 0) {
                    set_rc(-8);
                    set_err(tb_last_error());
                    p->failed = true;
                    done_q.push(p);
                } else {
                    inflight.push_back(p);
                }
            } else {
                p->failed = true;
                done_q.push(p);
            }
        }
        while (!inflight.empty()) {
            complete(inflight.front());
            inflight.pop_front();
        }
        if (clf) tb_classifier_destroy(clf);
    };
    std::vector<std::thread> workers;
    for (int d = 0; d < n_dev; d++) workers.emplace_back(device_worker, d);

    // ---- router ------------------------------------------------------------------------------
    std::string tsv;
    double write_s = 0;
    uint64_t n_a = 0, n_b = 0, n_u = 0, n_batches_done = 0;
    auto emit = [&](PipeBatch *p) {
        const auto t0 = Clock::now();
        Batch *b = &p->b;
        tsv.clear();
        for (int64_t i = 0; i < b->n() && rc.load() == 0 && !p->failed; i++) {
            const Rec &r = b->recs[(size_t)i];
            const char *name = b->text.data() + r.name_off;
            const char *seq = b->bases + b->offsets[(size_t)i];
            const size_t seq_len = (size_t)(b->offsets[(size_t)i + 1] - b->offsets[(size_t)i]);
            // Read.__str__ (seq.py:27-31): FASTQ only when the quality string is non-empty
            const bool fastq = r.has_qual && r.qual_len > 0;
            const char *parts[6] = { fastq ? "@" : ">", name, "\n", seq, "\n+\n", b->text.data() + r.qual_off };
            const size_t lens[6] = { 1, r.name_len, 1, seq_len, 3, r.qual_len };
            const char bin = b->bin[i];
            BinWriter &w = bin == 'A' ? *wa : (bin == 'B' ? *wb : *wu);
            (bin == 'A' ? n_a : (bin == 'B' ? n_b : n_u))++;
            w.record_parts(parts, lens, fastq ? 6 : 4);
            tsv.append(name, r.name_len);
            tsv += '\t';
            tsv += bin;
            tsv += '\t';
            py_float_repr(b->sa[i], tsv);
            tsv += '\t';
            py_float_repr(b->sb[i], tsv);
            tsv += '\n';
        }
        if (cfg.tsv_fd >= 0) {
            size_t off = 0;
            while (off < tsv.size()) {
                ssize_t w = ::write(cfg.tsv_fd, tsv.data() + off, tsv.size() - off);
                if (w <= 0) { set_rc(-6); break; }
                off += (size_t)w;
            }
        }
        wa->drain(false);
        wb->drain(false);
        wu->drain(false);
        if (wa->failed || wb->failed || wu->failed) set_rc(-9);  // disk full, quota: stop reading, do not classify on
        write_s += seconds_since(t0);
        n_batches_done++;
    };
    {
        std::map<uint64_t, PipeBatch *> held;
        uint64_t next_seq = 0;
        bool finished = false;
        while (!finished) {
            PipeBatch *p = done_q.pop();
            held[p->seq] = p;
            for (auto it = held.find(next_seq); it != held.end(); it = held.find(next_seq)) {
                PipeBatch *q = it->second;
                held.erase(it);
                emit(q);
                finished = q->b.last;
                next_seq++;
                free_q.push(&q->b);
                if (finished) break;
            }
        }
    }
    reader_thread.join();
    packer.join();
    for (auto &w : workers) w.join();
    const auto t_fin = Clock::now();
    // the reference's three locals are released in this order when main() returns
    wa->finish();
    wb->finish();
    wu->finish();
    write_s += seconds_since(t_fin);
    int out_rc = rc.load();
    if (wa->failed || wb->failed || wu->failed) {
        const int e = wa->fail_errno ? wa->fail_errno : (wb->fail_errno ? wb->fail_errno : wu->fail_errno);
        const std::string why = std::string("writing an output file failed: ") + (e ? strerror(e) : "compression error");
        set_err(why.c_str());
        out_rc = out_rc ? out_rc : -9;
    }
    if (!err_text.empty()) {
        fprintf(stderr, "tb_classify_fastx: %s\n", err_text.c_str());
        tb_internal_set_error(err_text.c_str());
    }
    if (!reader.error.empty()) {
        fprintf(stderr, "tb_classify_fastx: %s\n", reader.error.c_str());
        if (!out_rc) tb_internal_set_error(reader.error.c_str());
        out_rc = out_rc ? out_rc : -10;
    }
    if (stats) {
        memset(stats, 0, sizeof *stats);
        stats->n_reads = reader.n_reads;
        stats->n_bases = reader.n_bases;
        stats->n_batches = n_batches_done;
        stats->input_bytes = reader.src.bytes_in();
        stats->reads_a = n_a;
        stats->reads_b = n_b;
        stats->reads_u = n_u;
        stats->output_raw_bytes = wa->raw_bytes + wb->raw_bytes + wu->raw_bytes;
        stats->output_gz_bytes = wa->out_bytes + wb->out_bytes + wu->out_bytes;
        stats->seconds_total = seconds_since(t_all);
        stats->seconds_parse = reader.seconds;
        stats->seconds_gpu_wait = (double)gpu_wait_ns.load() * 1e-9;
        stats->seconds_write = write_s;
        stats->seconds_inflate = reader.src.inflate_seconds();
        stats->seconds_pack = pack_s;
        stats->seconds_deflate_cpu = (double)deflate_ns.load() * 1e-9;
        stats->devices_used = 0;
        for (int u : used) stats->devices_used += (uint64_t)u;
    }
    return out_rc;
}

}  // namespace

extern "C" {

int tb_classify_fastx_multi(const char *reads_path, const tb_stream_options *opt, const char *out_a,
                            const char *out_b, const char *out_u, tb_stream_stats *stats)
{
    if (!reads_path || !opt || !out_a || !out_b || !out_u) return -1;
    if (opt->n_devices < 1 || opt->n_devices > 64 || !opt->sets_a || !opt->sets_b) return -1;
    PipelineConfig cfg;
    cfg.reads_path = reads_path;
    for (int d = 0; d < opt->n_devices; d++) {
        if (!opt->sets_a[d] || !opt->sets_b[d]) return -1;
        cfg.sets_a.push_back(opt->sets_a[d]);
        cfg.sets_b.push_back(opt->sets_b[d]);
    }
    cfg.gzip_output = opt->gzip_output != 0;
    cfg.gzip_level = (opt->gzip_level < 0 || opt->gzip_level > 9) ? 9 : opt->gzip_level;
    cfg.fastq_encoder = opt->fastq_encoder != 0;
    cfg.round_robin = opt->round_robin != 0;
    cfg.py_text_mode = opt->py_text_mode;
    cfg.py_buffer_size = opt->py_buffer_size;
    cfg.tsv_fd = opt->tsv_fd;
    cfg.batch_bases = opt->batch_bases;
    cfg.threads = opt->n_threads;
    cfg.out_a = out_a;
    cfg.out_b = out_b;
    cfg.out_u = out_u;
    return run_pipeline(cfg, stats);
}

int tb_classify_fastx(const char *reads_path, const hash_set *A, const hash_set *B,
                      const char *out_a, const char *out_b, const char *out_u, int gzip_level,
                      int tsv_fd, int64_t batch_bases, int n_threads, tb_stream_stats *stats)
{
    if (!reads_path || !A || !B || !out_a || !out_b || !out_u) return -1;
    // TB_FASTQ_GZIP=0 turns the FASTQ-aware encoder off (zlib at gzip_level for every chunk)
    const char *fq_env = getenv("TB_FASTQ_GZIP");
    tb_stream_options opt;
    memset(&opt, 0, sizeof opt);
    opt.n_devices = 1;
    opt.sets_a = &A;
    opt.sets_b = &B;
    opt.gzip_output = 1;
    opt.gzip_level = gzip_level;
    opt.fastq_encoder = !(fq_env && strcmp(fq_env, "0") == 0);
    opt.tsv_fd = tsv_fd;
    opt.batch_bases = batch_bases;
    opt.n_threads = n_threads;
    return tb_classify_fastx_multi(reads_path, &opt, out_a, out_b, out_u, stats);
}

// testing aid (no GPU): the whole host pipeline -- reader, batches, router, the three writers
// and the TSV -- with the classifier replaced by a pattern: record i goes to bins[i % n_bins]
int tb_route_fastx(const char *reads_path, const char *bins, int64_t n_bins, int gzip_output, int gzip_level,
                   int fastq_encoder, int n_workers, int round_robin, int py_text_mode, int64_t py_buffer_size,
                   const char *out_a, const char *out_b, const char *out_u, int tsv_fd, int64_t batch_bases,
                   int n_threads, tb_stream_stats *stats)
{
    if (!reads_path || !bins || n_bins < 1 || !out_a || !out_b || !out_u) return -1;
    PipelineConfig cfg;
    cfg.reads_path = reads_path;
    cfg.stub_bins = bins;
    cfg.n_stub = n_bins;
    cfg.sets_a.assign((size_t)std::max(1, n_workers), nullptr);
    cfg.sets_b = cfg.sets_a;
    cfg.gzip_output = gzip_output != 0;
    cfg.gzip_level = (gzip_level < 0 || gzip_level > 9) ? 9 : gzip_level;
    cfg.fastq_encoder = fastq_encoder != 0;
    cfg.round_robin = round_robin != 0;
    cfg.py_text_mode = py_text_mode;
    cfg.py_buffer_size = py_buffer_size;
    cfg.tsv_fd = tsv_fd;
    cfg.batch_bases = batch_bases;
    cfg.threads = n_threads;
    cfg.out_a = out_a;
    cfg.out_b = out_b;
    cfg.out_u = out_u;
    return run_pipeline(cfg, stats);
}

// find-unique-kmers: every record of a FASTA/FASTQ(.gz) file through the k-mer counter
int tb_counter_add_fastx(tb_kmer_counter *counter, const char *reads_path, int64_t batch_bases, int n_threads,
                         uint64_t *n_reads, uint64_t *n_bases)
{
    if (!counter || !reads_path) return -1;
    Reader reader;
    reader.batch_bases = std::max<int64_t>(1, batch_bases);
    int threads = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    reader.src.set_threads(std::max(1, std::min(6, threads / 2)));
    reader.src.set_gzip_threads(threads - 2);  // counting has no compressor competing for the cores
    if (!reader.src.open(reads_path, reader.error)) {
        tb_internal_set_error(reader.error.c_str());
        return -2;
    }
    BatchQueue free_q, full_q;
    reader.free_q = &free_q;
    reader.full_q = &full_q;
    std::vector<std::unique_ptr<Batch>> batches;
    for (int i = 0; i < 3; i++) {
        batches.emplace_back(new Batch());
        batches.back()->pinned = false;
        if (!batches.back()->grow_bases(std::min<int64_t>(reader.batch_bases, 64 << 20) + (1 << 20))) return -3;
        free_q.push(batches.back().get());
    }
    std::thread t([&] { reader.run(); });
    int rc = 0;
    for (;;) {
        Batch *b = full_q.pop();
        if (rc == 0 && b->n() > 0 && tb_counter_add_batch(counter, b->bases, b->offsets.data(), b->n()) != 0) rc = -5;
        const bool last = b->last;
        free_q.push(b);
        if (last) break;
    }
    t.join();
    if (!reader.error.empty()) {
        tb_internal_set_error(reader.error.c_str());
        rc = rc ? rc : -10;
    }
    if (n_reads) *n_reads = reader.n_reads;
    if (n_bases) *n_bases = reader.n_bases;
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


// exposed for tests: the lines of a (possibly compressed) file as the parser sees them, written to
// out_fd one after the other (line ends already translated to '\n').  Returns the number of bytes.
int64_t tb_fastx_cat(const char *reads_path, int out_fd, int decode_threads)
{
    LineSource src;
    std::string err;
    src.set_threads(decode_threads);
    src.set_gzip_threads(decode_threads);
    if (!src.open(reads_path, err)) {
        tb_internal_set_error(err.c_str());
        return -1;
    }
    std::string out;
    int64_t total = 0;
    const char *lp;
    size_t ln;
    auto flush = [&]() {
        size_t off = 0;
        while (off < out.size()) {
            const ssize_t w = ::write(out_fd, out.data() + off, out.size() - off);
            if (w < 0 && errno == EINTR) continue;
            if (w <= 0) return false;
            off += (size_t)w;
        }
        out.clear();
        return true;
    };
    while (src.next(lp, ln)) {
        out.append(lp, ln);
        total += (int64_t)ln;
        if (out.size() >= (1u << 20) && !flush()) return -3;
    }
    if (!flush()) return -3;
    if (!src.error().empty()) {
        tb_internal_set_error(src.error().c_str());
        return -2;
    }
    return total;
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
    w.fastq_encoder = fastq_encoder == 1;
    w.bgzf = fastq_encoder == 2;
    w.max_inflight = (size_t)std::max(3, 2 * n_threads);  // small: exercises the back-pressure path
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
