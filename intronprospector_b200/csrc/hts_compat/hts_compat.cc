/* hts_compat.cc -- self-contained SAM / BAM (BGZF) / FASTA+fai reader and writer.
 *
 * htslib is not available in the build image, so this file supplies the small
 * subset of its public API that the junction-extraction path needs (see the
 * headers under htslib/ in this directory).  It is written from the SAM/BAM
 * specification (SAMv1: record layout, BGZF framing, reg2bin, aux encodings),
 * not from htslib sources.  Both the product's host driver and the oracle build
 * of the unmodified reference link against it, so the reference's golden tests
 * (pass-through identity diffs in particular) are what validate it.
 *
 * Reading goes through zlib's gzread(), which transparently handles plain text,
 * gzip, concatenated gzip members (= BGZF) and an uncompressed "raw" BAM stream
 * (a file that starts with "BAM\1" without the BGZF wrapper; used by bench.py to
 * take inflate time out of the CPU-baseline measurement).
 */
#include "htslib/hts.h"
#include "htslib/sam.h"
#include "htslib/faidx.h"
#include "htslib/kstring.h"

#include <zlib.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cerrno>
#include <algorithm>
#include <cctype>
#include <string>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>
#include <unordered_map>

namespace {

/* ---------- buffered reader over gzread ---------- */
/* BGZF is a series of independent gzip members of at most 64 KiB each, so a BAM file inflates in parallel: one thread
 * reads the compressed blocks in file order, a few workers inflate them, and the consumer takes the results back in
 * order.  (The reference gets the same from htslib's thread pool when asked; here it is the default for BGZF input.) */
class ParallelBgzf {
    struct Slot {
        std::vector<uint8_t> in, out;
        size_t out_len = 0, out_pos = 0;
        int state = 0; /* 0 free, 1 filled (compressed), 2 being inflated, 3 inflated */
        bool bad = false;
    };
    FILE *fp_ = nullptr;
    std::vector<uint8_t> prefix_; /* bytes of the stream somebody already took from fp_ (input that cannot seek) */
    size_t prefix_pos_ = 0;
    std::vector<Slot> ring_;
    std::vector<std::thread> threads_;
    std::mutex mu_;
    std::condition_variable cv_;
    uint64_t n_read_ = 0;     /* blocks the reader thread has put into the ring */
    uint64_t n_claimed_ = 0;  /* blocks handed to inflate workers */
    uint64_t n_consumed_ = 0; /* blocks the consumer is done with */
    bool read_done_ = false, stop_ = false, read_error_ = false;

    size_t take(void *dst, size_t n) { /* reader thread only */
        size_t done = 0;
        if (prefix_pos_ < prefix_.size()) {
            done = std::min(n, prefix_.size() - prefix_pos_);
            memcpy(dst, prefix_.data() + prefix_pos_, done);
            prefix_pos_ += done;
        }
        if (done < n) done += fread((uint8_t *)dst + done, 1, n - done, fp_);
        return done;
    }

    void reader_main() {
        for (;;) {
            uint8_t hdr[18];
            size_t got = take(hdr, 18);
            bool ok = got == 18 && hdr[0] == 0x1f && hdr[1] == 0x8b && hdr[2] == 8 && (hdr[3] & 4) && hdr[12] == 'B' && hdr[13] == 'C';
            size_t bsize = ok ? (size_t)(hdr[16] | (hdr[17] << 8)) + 1 : 0;
            std::unique_lock<std::mutex> lk(mu_);
            if (got == 0 || !ok || bsize < 26) {
                if (got != 0) read_error_ = true; /* not a BGZF block where one must start */
                read_done_ = true;
                cv_.notify_all();
                return;
            }
            cv_.wait(lk, [&] { return stop_ || n_read_ - n_consumed_ < ring_.size(); });
            if (stop_) return;
            Slot &s = ring_[n_read_ % ring_.size()];
            lk.unlock();
            s.in.resize(bsize);
            memcpy(s.in.data(), hdr, 18);
            if (take(s.in.data() + 18, bsize - 18) != bsize - 18) {
                lk.lock();
                read_error_ = true;
                read_done_ = true;
                cv_.notify_all();
                return;
            }
            lk.lock();
            s.state = 1;
            s.bad = false;
            n_read_++;
            cv_.notify_all();
        }
    }

    void worker_main() {
        z_stream zs;
        memset(&zs, 0, sizeof(zs));
        if (inflateInit2(&zs, -15) != Z_OK) return;
        for (;;) {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [&] { return stop_ || n_claimed_ < n_read_ || (read_done_ && n_claimed_ >= n_read_); });
            if (stop_ || n_claimed_ >= n_read_) break;
            Slot &s = ring_[n_claimed_ % ring_.size()];
            n_claimed_++;
            s.state = 2;
            lk.unlock();
            const size_t bsize = s.in.size();
            uint32_t isize;
            memcpy(&isize, s.in.data() + bsize - 4, 4);
            s.out.resize(isize ? isize : 1);
            inflateReset(&zs);
            zs.next_in = s.in.data() + 18;
            zs.avail_in = (uInt)(bsize - 18 - 8);
            zs.next_out = s.out.data();
            zs.avail_out = (uInt)isize;
            int rc = isize ? inflate(&zs, Z_FINISH) : Z_STREAM_END;
            const bool bad = rc != Z_STREAM_END || zs.avail_out != 0;
            lk.lock();
            s.bad = bad;
            s.out_len = isize;
            s.out_pos = 0;
            s.state = 3;
            cv_.notify_all();
        }
        inflateEnd(&zs);
    }

   public:
    static bool looks_like_bgzf(const char *fn) {
        FILE *f = fopen(fn, "rb");
        if (!f) return false;
        uint8_t h[16];
        bool ok = fread(h, 1, 16, f) == 16 && h[0] == 0x1f && h[1] == 0x8b && h[2] == 8 && (h[3] & 4) && h[12] == 'B' && h[13] == 'C';
        fclose(f);
        return ok;
    }
    static bool is_bgzf_header(const uint8_t *h, size_t n) {
        return n >= 16 && h[0] == 0x1f && h[1] == 0x8b && h[2] == 8 && (h[3] & 4) && h[12] == 'B' && h[13] == 'C';
    }
    /* a stream that is already open and of which `prefix` was already read (takes the FILE over) */
    bool open_stream(FILE *f, const uint8_t *prefix, size_t n_prefix, int n_workers) {
        fp_ = f;
        prefix_.assign(prefix, prefix + n_prefix);
        prefix_pos_ = 0;
        ring_.resize((size_t)n_workers * 8);
        threads_.emplace_back([this] { reader_main(); });
        for (int i = 0; i < n_workers; i++) threads_.emplace_back([this] { worker_main(); });
        return true;
    }
    bool open(const char *fn, int n_workers) {
        fp_ = fopen(fn, "rb");
        if (!fp_) return false;
        setvbuf(fp_, nullptr, _IOFBF, 1 << 22);
        ring_.resize((size_t)n_workers * 8);
        threads_.emplace_back([this] { reader_main(); });
        for (int i = 0; i < n_workers; i++) threads_.emplace_back([this] { worker_main(); });
        return true;
    }
    /* up to `want` inflated bytes, in file order; 0 at the end of the file, -1 on a damaged block */
    int read(uint8_t *dst, int want) {
        int done = 0;
        while (done < want) {
            std::unique_lock<std::mutex> lk(mu_);
            Slot &s = ring_[n_consumed_ % ring_.size()];
            cv_.wait(lk, [&] { return (n_consumed_ < n_read_ && s.state == 3) || (read_done_ && n_consumed_ >= n_read_); });
            if (n_consumed_ >= n_read_) return read_error_ && done == 0 ? -1 : done;
            if (s.bad) return -1;
            lk.unlock();
            const size_t take = std::min<size_t>(s.out_len - s.out_pos, (size_t)(want - done));
            memcpy(dst + done, s.out.data() + s.out_pos, take);
            s.out_pos += take;
            done += (int)take;
            if (s.out_pos == s.out_len) {
                lk.lock();
                s.state = 0;
                n_consumed_++;
                cv_.notify_all();
            } else {
                break; /* `want` is satisfied */
            }
            if (done > 0 && done >= want) break;
        }
        return done;
    }
    void close() {
        {
            std::lock_guard<std::mutex> g(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : threads_) t.join();
        threads_.clear();
        if (fp_) fclose(fp_);
        fp_ = nullptr;
    }
    ~ParallelBgzf() { close(); }
};

/* Sequential input that cannot be reopened or rewound (a pipe on stdin) and of which a few bytes were already read to
 * see what it is: those bytes first, then the rest of the FILE -- as they are, or inflated when they are gzip members
 * (any gzip, BGZF included).  What gzread does, minus the need to hand zlib a descriptor it reads from the start. */
class StreamSource {
    FILE *fp_ = nullptr;
    std::vector<uint8_t> in_;
    size_t beg_ = 0, end_ = 0;
    bool gz_ = false, z_ready_ = false, in_eof_ = false, member_done_ = true;
    z_stream zs_;

    void refill() {
        if (beg_ < end_ || in_eof_) return;
        beg_ = 0;
        end_ = fread(in_.data(), 1, in_.size(), fp_);
        if (end_ == 0) in_eof_ = true;
    }

   public:
    bool gzip() const { return gz_; }
    bool open(FILE *f, const uint8_t *prefix, size_t n_prefix) {
        fp_ = f;
        in_.resize(1 << 20);
        memcpy(in_.data(), prefix, n_prefix);
        beg_ = 0;
        end_ = n_prefix;
        gz_ = n_prefix >= 2 && prefix[0] == 0x1f && prefix[1] == 0x8b;
        if (gz_) {
            memset(&zs_, 0, sizeof(zs_));
            if (inflateInit2(&zs_, 15 + 16) != Z_OK) return false;
            z_ready_ = true;
        }
        return true;
    }
    int read(uint8_t *dst, int want) {
        int done = 0;
        if (!gz_) {
            while (done < want) {
                refill();
                if (beg_ == end_) break;
                const size_t n = std::min<size_t>(end_ - beg_, (size_t)(want - done));
                memcpy(dst + done, in_.data() + beg_, n);
                beg_ += n;
                done += (int)n;
            }
            return done;
        }
        while (done < want) {
            refill();
            if (beg_ == end_) return member_done_ || done > 0 ? done : -1; /* the input ends inside a member: truncated */
            if (member_done_) {
                inflateReset(&zs_); /* the next member */
                member_done_ = false;
            }
            zs_.next_in = in_.data() + beg_;
            zs_.avail_in = (uInt)(end_ - beg_);
            zs_.next_out = dst + done;
            zs_.avail_out = (uInt)(want - done);
            const int rc = inflate(&zs_, Z_NO_FLUSH);
            beg_ = end_ - zs_.avail_in;
            done = want - (int)zs_.avail_out;
            if (rc == Z_STREAM_END) member_done_ = true;
            else if (rc != Z_OK && rc != Z_BUF_ERROR) return -1;
        }
        return done;
    }
    void close() {
        if (z_ready_) inflateEnd(&zs_);
        z_ready_ = false;
        if (fp_) fclose(fp_);
        fp_ = nullptr;
    }
    ~StreamSource() { close(); }
};

struct Reader {
    gzFile gz = nullptr;
    StreamSource *stream = nullptr; /* input that cannot seek (stdin), not handled by the BGZF decoder pool */
    ParallelBgzf *pbgzf = nullptr; /* BGZF input with more than one decode thread */
    bool direct = false;           /* the input is not compressed */
    std::vector<uint8_t> buf;
    size_t beg = 0, end = 0;
    bool eof = false;
    bool err = false;

    static int decode_threads() {
        if (const char *e = getenv("IPX_DECODE_THREADS")) return atoi(e);
        unsigned hw = std::thread::hardware_concurrency();
        return (int)std::min<unsigned>(8u, hw > 2 ? hw / 2 : 1u);
    }
    bool open(const char *fn) {
        if (strcmp(fn, "-") == 0) fn = "/dev/stdin";
        const int threads = decode_threads();
        if (strcmp(fn, "/dev/stdin") == 0) {
            /* a pipe cannot be looked at and then handed to zlib from the start: read the first bytes here and pass
             * them on -- to the BGZF decoder pool when that is what they are, else to a plain / gzip stream reader */
            FILE *f = fopen(fn, "rb");
            if (!f) return false;
            setvbuf(f, nullptr, _IOFBF, 1 << 22);
            uint8_t head[18];
            const size_t n_head = fread(head, 1, sizeof(head), f);
            buf.resize(1 << 20);
            if (threads > 1 && ParallelBgzf::is_bgzf_header(head, n_head)) {
                pbgzf = new ParallelBgzf();
                return pbgzf->open_stream(f, head, n_head, threads);
            }
            stream = new StreamSource();
            if (!stream->open(f, head, n_head)) return false;
            direct = !stream->gzip();
            return true;
        }
        if (threads > 1 && ParallelBgzf::looks_like_bgzf(fn)) {
            pbgzf = new ParallelBgzf();
            if (!pbgzf->open(fn, threads)) {
                delete pbgzf;
                pbgzf = nullptr;
                return false;
            }
            buf.resize(1 << 20);
            return true;
        }
        gz = gzopen(fn, "rb");
        if (!gz) return false;
        gzbuffer(gz, 1 << 20);
        buf.resize(1 << 20);
        direct = gzdirect(gz) != 0;
        return true;
    }
    void close() {
        if (gz) gzclose(gz);
        gz = nullptr;
        delete pbgzf;
        pbgzf = nullptr;
        delete stream;
        stream = nullptr;
    }
    int raw_read(uint8_t *dst, int want) {
        return pbgzf ? pbgzf->read(dst, want) : (stream ? stream->read(dst, want) : gzread(gz, dst, want));
    }
    /* make at least n bytes available if possible; returns bytes available */
    size_t ensure(size_t n) {
        if (end - beg >= n) return end - beg;
        if (beg > 0) {
            memmove(buf.data(), buf.data() + beg, end - beg);
            end -= beg;
            beg = 0;
        }
        if (buf.size() < n) buf.resize(n + (1 << 16));
        while (end - beg < n && !eof) {
            int want = (int)std::min<size_t>(buf.size() - end, 1u << 30);
            int got = raw_read(buf.data() + end, want);
            if (got < 0) { err = true; eof = true; break; }
            if (got == 0) { eof = true; break; }
            end += got;
        }
        return end - beg;
    }
    bool read_exact(void *dst, size_t n) {
        if (ensure(n) < n) return false;
        memcpy(dst, buf.data() + beg, n);
        beg += n;
        return true;
    }
    /* returns pointer to n contiguous bytes and consumes them, or NULL */
    const uint8_t *take(size_t n) {
        if (ensure(n) < n) return nullptr;
        const uint8_t *p = buf.data() + beg;
        beg += n;
        return p;
    }
    int peek() {
        if (ensure(1) < 1) return -1;
        return buf[beg];
    }
    /* read one line without the terminator; false at EOF with nothing read */
    bool getline(std::string &line) {
        line.clear();
        for (;;) {
            if (end == beg) {
                if (ensure(1) < 1) return !line.empty();
            }
            const uint8_t *p = buf.data() + beg;
            const uint8_t *nl = (const uint8_t *)memchr(p, '\n', end - beg);
            if (nl) {
                line.append((const char *)p, nl - p);
                beg += (nl - p) + 1;
                if (!line.empty() && line.back() == '\r') line.pop_back();
                return true;
            }
            line.append((const char *)p, end - beg);
            beg = end;
        }
    }
};

/* ---------- BGZF block writer ---------- */
struct BgzfWriter {
    FILE *fp = nullptr;
    std::vector<uint8_t> in;
    int level = Z_DEFAULT_COMPRESSION;
    static const size_t BLOCK = 0xff00;

    bool write_block(const uint8_t *data, size_t n) {
        uint8_t out[65536 + 64];
        z_stream zs;
        memset(&zs, 0, sizeof(zs));
        if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return false;
        zs.next_in = const_cast<uint8_t *>(data);
        zs.avail_in = (uInt)n;
        zs.next_out = out + 18;
        zs.avail_out = sizeof(out) - 18 - 8;
        int rc = deflate(&zs, Z_FINISH);
        if (rc != Z_STREAM_END) { deflateEnd(&zs); return false; }
        size_t clen = zs.total_out;
        deflateEnd(&zs);
        static const uint8_t hdr[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
        memcpy(out, hdr, 16);
        size_t total = 18 + clen + 8;
        uint16_t bsize = (uint16_t)(total - 1);
        out[16] = bsize & 0xff;
        out[17] = bsize >> 8;
        uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), data, (uInt)n);
        uint32_t isize = (uint32_t)n;
        memcpy(out + 18 + clen, &crc, 4);
        memcpy(out + 18 + clen + 4, &isize, 4);
        return fwrite(out, 1, total, fp) == total;
    }
    bool write(const void *p, size_t n) {
        const uint8_t *s = (const uint8_t *)p;
        while (n > 0) {
            size_t room = BLOCK - in.size();
            size_t k = n < room ? n : room;
            in.insert(in.end(), s, s + k);
            s += k;
            n -= k;
            if (in.size() == BLOCK) {
                if (!write_block(in.data(), in.size())) return false;
                in.clear();
            }
        }
        return true;
    }
    bool flush() {
        if (!in.empty()) {
            if (!write_block(in.data(), in.size())) return false;
            in.clear();
        }
        return true;
    }
    bool finish() {
        if (!flush()) return false;
        static const uint8_t eofb[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0,
                                         0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        return fwrite(eofb, 1, 28, fp) == 28;
    }
};

}  // namespace

struct ipx_hts_impl {
    Reader rd;
    std::string pending_line;  /* first alignment line seen while reading a SAM header */
    bool have_pending = false;
    FILE *out = nullptr;       /* text or raw-binary output */
    BgzfWriter bgzf;
    bool raw_bam = false;      /* write uncompressed BAM without BGZF framing */
    std::unordered_map<std::string, int> name2tid;
    bool name_map_built = false;
    std::string scratch;
    uint64_t bam_header_bytes = 0; /* uncompressed bytes of a BAM file before its first record */
};

/* ---------- open / close ---------- */
static htsFile *alloc_file(const char *fn) {
    htsFile *fp = (htsFile *)calloc(1, sizeof(htsFile));
    fp->fn = strdup(fn);
    fp->impl = new ipx_hts_impl();
    return fp;
}

static void free_file(htsFile *fp) {
    delete fp->impl;
    free(fp->fn);
    free(fp);
}

extern "C" htsFile *hts_open_format(const char *fn, const char *mode, const htsFormat *fmt) {
    bool writing = strchr(mode, 'w') != nullptr;
    htsFile *fp = alloc_file(fn);
    if (!writing) {
        if (!fp->impl->rd.open(fn)) {
            free_file(fp);
            return nullptr;
        }
        Reader &rd = fp->impl->rd;
        size_t n = rd.ensure(4);
        fp->format.category = sequence_data;
        if (n >= 4 && memcmp(rd.buf.data() + rd.beg, "BAM\1", 4) == 0) {
            fp->format.format = bam;
            fp->format.compression = rd.direct ? no_compression : bgzf;
            fp->is_bin = 1;
        } else {
            fp->format.format = sam;
            fp->format.compression = rd.direct ? no_compression : gzip;
        }
        return fp;
    }
    fp->is_write = 1;
    enum htsExactFormat ef = sam;
    if (fmt != nullptr) {
        ef = fmt->format;
    } else if (strchr(mode, 'b') != nullptr) {
        ef = bam;
    }
    if (ef != sam && ef != bam) {
        free_file(fp);
        return nullptr;
    }
    fp->format.category = sequence_data;
    fp->format.format = ef;
    fp->is_bin = (ef == bam);
    FILE *out = (strcmp(fn, "-") == 0) ? stdout : fopen(fn, "wb");
    if (!out) {
        free_file(fp);
        return nullptr;
    }
    static const size_t OBUF = 1 << 20;
    setvbuf(out, nullptr, _IOFBF, OBUF);
    fp->impl->out = out;
    fp->impl->bgzf.fp = out;
    if (ef == bam) {
        fp->format.compression = bgzf;
        /* mode digit selects compression level, 'u' / '0' => stored blocks */
        for (const char *m = mode; *m; ++m) {
            if (*m >= '0' && *m <= '9') fp->impl->bgzf.level = *m - '0';
            if (*m == 'u') fp->impl->bgzf.level = 0;
            if (*m == 'R') fp->impl->raw_bam = true;  /* extension: no BGZF wrapper */
        }
    }
    return fp;
}

extern "C" htsFile *hts_open(const char *fn, const char *mode) {
    return hts_open_format(fn, mode, nullptr);
}

extern "C" int hts_close(htsFile *fp) {
    int rc = 0;
    if (fp->is_write) {
        if (fp->is_bin && !fp->impl->raw_bam) {
            if (!fp->impl->bgzf.finish()) rc = -1;
        }
        if (fp->impl->out) {
            if (fflush(fp->impl->out) != 0) rc = -1;
            if (fp->impl->out != stdout) fclose(fp->impl->out);
        }
    } else {
        fp->impl->rd.close();
    }
    free_file(fp);
    return rc;
}

extern "C" const htsFormat *hts_get_format(htsFile *fp) {
    return fp ? &fp->format : nullptr;
}

/* ---------- header ---------- */
extern "C" sam_hdr_t *sam_hdr_init(void) {
    return (sam_hdr_t *)calloc(1, sizeof(sam_hdr_t));
}

extern "C" void sam_hdr_destroy(sam_hdr_t *h) {
    if (!h) return;
    for (int i = 0; i < h->n_targets; i++) free(h->target_name[i]);
    free(h->target_name);
    free(h->target_len);
    free(h->text);
    free(h);
}

static void hdr_add_target(sam_hdr_t *h, const char *name, size_t name_len, uint32_t len) {
    int n = h->n_targets;
    h->target_name = (char **)realloc(h->target_name, sizeof(char *) * (n + 1));
    h->target_len = (uint32_t *)realloc(h->target_len, sizeof(uint32_t) * (n + 1));
    h->target_name[n] = strndup(name, name_len);
    h->target_len[n] = len;
    h->n_targets = n + 1;
}

/* pull SN/LN out of the @SQ lines of header text */
static void hdr_parse_sq(sam_hdr_t *h) {
    const char *p = h->text;
    const char *e = h->text + h->l_text;
    while (p < e) {
        const char *nl = (const char *)memchr(p, '\n', e - p);
        const char *le = nl ? nl : e;
        if (le - p >= 3 && p[0] == '@' && p[1] == 'S' && p[2] == 'Q') {
            const char *sn = nullptr;
            size_t sn_len = 0;
            uint32_t ln = 0;
            const char *f = p + 3;
            while (f < le) {
                if (*f == '\t') { f++; continue; }
                const char *fe = (const char *)memchr(f, '\t', le - f);
                if (!fe) fe = le;
                if (fe - f >= 3 && f[2] == ':') {
                    if (f[0] == 'S' && f[1] == 'N') { sn = f + 3; sn_len = fe - sn; }
                    else if (f[0] == 'L' && f[1] == 'N') ln = (uint32_t)strtoul(f + 3, nullptr, 10);
                }
                f = fe;
            }
            if (sn) hdr_add_target(h, sn, sn_len, ln);
        }
        p = nl ? nl + 1 : e;
    }
}

extern "C" sam_hdr_t *sam_hdr_read(samFile *fp) {
    ipx_hts_impl *im = fp->impl;
    Reader &rd = im->rd;
    sam_hdr_t *h = sam_hdr_init();
    if (fp->format.format == bam) {
        uint8_t magic[4];
        int32_t l_text, n_ref;
        if (!rd.read_exact(magic, 4) || memcmp(magic, "BAM\1", 4) != 0 || !rd.read_exact(&l_text, 4) || l_text < 0) {
            sam_hdr_destroy(h);
            return nullptr;
        }
        h->text = (char *)malloc((size_t)l_text + 1);
        if (!rd.read_exact(h->text, l_text)) { sam_hdr_destroy(h); return nullptr; }
        h->text[l_text] = 0;
        h->l_text = strlen(h->text);  /* text may be NUL padded */
        uint64_t consumed = 8 + (uint64_t)l_text + 4;
        if (!rd.read_exact(&n_ref, 4) || n_ref < 0) { sam_hdr_destroy(h); return nullptr; }
        for (int i = 0; i < n_ref; i++) {
            int32_t l_name, l_ref;
            if (!rd.read_exact(&l_name, 4) || l_name <= 0) { sam_hdr_destroy(h); return nullptr; }
            std::string nm(l_name, 0);
            if (!rd.read_exact(&nm[0], l_name) || !rd.read_exact(&l_ref, 4)) { sam_hdr_destroy(h); return nullptr; }
            hdr_add_target(h, nm.c_str(), strlen(nm.c_str()), (uint32_t)l_ref);
            consumed += 8 + (uint64_t)l_name;
        }
        fp->impl->bam_header_bytes = consumed;
        return h;
    }
    std::string text, line;
    while (rd.peek() == '@') {
        rd.getline(line);
        text += line;
        text += '\n';
    }
    h->text = strdup(text.c_str());
    h->l_text = text.size();
    hdr_parse_sq(h);
    return h;
}

extern "C" int sam_hdr_name2tid(sam_hdr_t *h, const char *ref) {
    for (int i = 0; i < h->n_targets; i++)
        if (strcmp(h->target_name[i], ref) == 0) return i;
    return -1;
}

extern "C" int sam_hdr_write(samFile *fp, const sam_hdr_t *h) {
    ipx_hts_impl *im = fp->impl;
    if (fp->format.format == sam) {
        if (h->l_text && fwrite(h->text, 1, h->l_text, im->out) != h->l_text) return -1;
        return 0;
    }
    std::string b;
    b.append("BAM\1", 4);
    int32_t l_text = (int32_t)h->l_text;
    b.append((const char *)&l_text, 4);
    b.append(h->text ? h->text : "", h->l_text);
    int32_t n_ref = h->n_targets;
    b.append((const char *)&n_ref, 4);
    for (int i = 0; i < n_ref; i++) {
        int32_t l_name = (int32_t)strlen(h->target_name[i]) + 1;
        b.append((const char *)&l_name, 4);
        b.append(h->target_name[i], l_name);
        int32_t l_ref = (int32_t)h->target_len[i];
        b.append((const char *)&l_ref, 4);
    }
    if (im->raw_bam) return fwrite(b.data(), 1, b.size(), im->out) == b.size() ? 0 : -1;
    if (!im->bgzf.write(b.data(), b.size())) return -1;
    return im->bgzf.flush() ? 0 : -1;  /* header ends its own block, as samtools does */
}

/* ---------- records ---------- */
extern "C" bam1_t *bam_init1(void) {
    return (bam1_t *)calloc(1, sizeof(bam1_t));
}

extern "C" void bam_destroy1(bam1_t *b) {
    if (!b) return;
    free(b->data);
    free(b);
}

static int reserve_data(bam1_t *b, size_t n) {
    if (n > b->m_data) {
        size_t m = n + (n >> 1) + 64;
        uint8_t *p = (uint8_t *)realloc(b->data, m);
        if (!p) return -1;
        b->data = p;
        b->m_data = (uint32_t)m;
    }
    return 0;
}

extern "C" bam1_t *bam_copy1(bam1_t *dst, const bam1_t *src) {
    if (reserve_data(dst, src->l_data) < 0) return nullptr;
    memcpy(dst->data, src->data, src->l_data);
    dst->l_data = src->l_data;
    dst->core = src->core;
    dst->id = src->id;
    return dst;
}

extern "C" hts_pos_t bam_cigar2rlen(int n_cigar, const uint32_t *cigar) {
    hts_pos_t l = 0;
    for (int k = 0; k < n_cigar; ++k)
        if (bam_cigar_type(bam_cigar_op(cigar[k])) & 2) l += bam_cigar_oplen(cigar[k]);
    return l;
}

extern "C" hts_pos_t bam_endpos(const bam1_t *b) {
    hts_pos_t rlen = (b->core.flag & BAM_FUNMAP) ? 0 : bam_cigar2rlen(b->core.n_cigar, bam_get_cigar(b));
    if (rlen == 0) rlen = 1;
    return b->core.pos + rlen;
}

static int reg2bin(int64_t beg, int64_t end) {
    --end;
    if (beg >> 14 == end >> 14) return (int)(((1 << 15) - 1) / 7 + (beg >> 14));
    if (beg >> 17 == end >> 17) return (int)(((1 << 12) - 1) / 7 + (beg >> 17));
    if (beg >> 20 == end >> 20) return (int)(((1 << 9) - 1) / 7 + (beg >> 20));
    if (beg >> 23 == end >> 23) return (int)(((1 << 6) - 1) / 7 + (beg >> 23));
    if (beg >> 26 == end >> 26) return (int)(((1 << 3) - 1) / 7 + (beg >> 26));
    return 0;
}

/* move a real CIGAR held in a CG:B:I tag back in place of the <l_seq>S<rlen>N placeholder */
static void restore_long_cigar(bam1_t *b) {
    bam1_core_t *c = &b->core;
    if (c->n_cigar != 2 || c->tid < 0 || c->pos < 0) return;
    const uint32_t *cig = bam_get_cigar(b);
    if (bam_cigar_op(cig[0]) != BAM_CSOFT_CLIP || (int32_t)bam_cigar_oplen(cig[0]) != c->l_qseq) return;
    if (bam_cigar_op(cig[1]) != BAM_CREF_SKIP) return;
    uint8_t *cg = bam_aux_get(b, "CG");
    if (!cg || cg[0] != 'B' || (cg[1] != 'I' && cg[1] != 'i')) return;
    uint32_t n;
    memcpy(&n, cg + 2, 4);
    if (n == 0) return;
    std::vector<uint32_t> real(n);
    memcpy(real.data(), cg + 6, (size_t)n * 4);
    /* drop the tag */
    size_t tag_beg = (cg - 2) - b->data, tag_end = tag_beg + 2 + 6 + (size_t)n * 4;
    memmove(b->data + tag_beg, b->data + tag_end, b->l_data - tag_end);
    b->l_data -= (int)(tag_end - tag_beg);
    /* widen the cigar area */
    size_t cig_off = c->l_qname;
    size_t extra = ((size_t)n - 2) * 4;
    reserve_data(b, b->l_data + extra);
    memmove(b->data + cig_off + (size_t)n * 4, b->data + cig_off + 8, b->l_data - cig_off - 8);
    memcpy(b->data + cig_off, real.data(), (size_t)n * 4);
    b->l_data += (int)extra;
    c->n_cigar = n;
}

static int bam_read_record(ipx_hts_impl *im, bam1_t *b) {
    Reader &rd = im->rd;
    int32_t block_len = 0;
    if (rd.ensure(4) < 4) return rd.err ? -2 : ((rd.end - rd.beg) == 0 ? -1 : -4);
    rd.read_exact(&block_len, 4);
    if (block_len < 32) return -4;
    const uint8_t *x = rd.take(32);
    if (!x) return -4;
    bam1_core_t *c = &b->core;
    int32_t refid, pos, l_seq, mtid, mpos, isize;
    uint32_t bin_mq_nl, flag_nc;
    memcpy(&refid, x, 4);
    memcpy(&pos, x + 4, 4);
    memcpy(&bin_mq_nl, x + 8, 4);
    memcpy(&flag_nc, x + 12, 4);
    memcpy(&l_seq, x + 16, 4);
    memcpy(&mtid, x + 20, 4);
    memcpy(&mpos, x + 24, 4);
    memcpy(&isize, x + 28, 4);
    c->tid = refid;
    c->pos = pos;
    c->bin = bin_mq_nl >> 16;
    c->qual = (bin_mq_nl >> 8) & 0xff;
    uint32_t l_qname = bin_mq_nl & 0xff;
    c->flag = flag_nc >> 16;
    c->n_cigar = flag_nc & 0xffff;
    c->l_qseq = l_seq;
    c->mtid = mtid;
    c->mpos = mpos;
    c->isize = isize;
    c->l_extranul = (l_qname % 4 != 0) ? (4 - l_qname % 4) : 0;
    size_t rest = (size_t)block_len - 32;
    if (rest < l_qname) return -4;
    size_t new_l = rest + c->l_extranul;
    if (reserve_data(b, new_l) < 0) return -4;
    const uint8_t *src = rd.take(rest);
    if (!src) return -4;
    memcpy(b->data, src, l_qname);
    memset(b->data + l_qname, 0, c->l_extranul);
    memcpy(b->data + l_qname + c->l_extranul, src + l_qname, rest - l_qname);
    c->l_qname = (uint16_t)(l_qname + c->l_extranul);
    b->l_data = (int)new_l;
    if ((size_t)c->l_qname + (size_t)c->n_cigar * 4 + (size_t)((l_seq + 1) / 2) + (size_t)l_seq > new_l) return -4;
    restore_long_cigar(b);
    return 0;
}

extern "C" int ipx_bam_peek(samFile *fp, const uint8_t **rec, uint32_t *len) {
    if (!fp || fp->is_write || fp->format.format != bam) return 0;
    Reader &rd = fp->impl->rd;
    if (rd.ensure(4) < 4) return 0;
    int32_t block_len;
    memcpy(&block_len, rd.buf.data() + rd.beg, 4);
    if (block_len < 32) return 0;
    if (rd.ensure(4 + (size_t)block_len) < 4 + (size_t)block_len) return 0;
    const uint8_t *x = rd.buf.data() + rd.beg + 4;
    int32_t l_seq;
    memcpy(&l_seq, x + 16, 4);
    uint16_t n_cigar;
    memcpy(&n_cigar, x + 12, 2);
    const size_t l_qname = x[8], rest = (size_t)block_len - 32;
    if (l_seq < 0 || rest < l_qname) return 0;
    if (l_qname + (size_t)n_cigar * 4 + (size_t)((l_seq + 1) / 2) + (size_t)l_seq > rest) return 0;
    if (n_cigar == 2) {
        /* "<l_seq>S<ref_len>N": the real CIGAR may sit in a CG:B,I tag (restore_long_cigar) */
        uint32_t c0;
        memcpy(&c0, x + 32 + l_qname, 4);
        if ((c0 & 15u) == BAM_CSOFT_CLIP && (int32_t)(c0 >> 4) == l_seq) return 0;
    }
    *rec = x;
    *len = (uint32_t)block_len;
    return 1;
}

extern "C" uint64_t ipx_bam_header_bytes(samFile *fp) { return (fp && fp->impl) ? fp->impl->bam_header_bytes : 0; }

extern "C" void ipx_bam_consume(samFile *fp) {
    Reader &rd = fp->impl->rd;
    int32_t block_len;
    memcpy(&block_len, rd.buf.data() + rd.beg, 4);
    rd.beg += 4 + (size_t)block_len;
}

static const uint8_t *nt16_table() {
    static uint8_t tab[256];
    static bool init = false;
    if (!init) {
        memset(tab, 15, sizeof(tab));
        const char *codes = "=ACMGRSVTWYHKDBN";
        for (int i = 0; i < 16; i++) {
            tab[(uint8_t)codes[i]] = (uint8_t)i;
            tab[(uint8_t)tolower(codes[i])] = (uint8_t)i;
        }
        init = true;
    }
    return tab;
}

static inline void put_bytes(std::string &d, const void *p, size_t n) {
    d.append((const char *)p, n);
}

/* encode one TAG:TYPE:VALUE text field; returns false on malformed input */
static bool aux_text_to_bin(const char *f, size_t n, std::string &d) {
    if (n < 5 || f[2] != ':' || f[4] != ':') return false;
    d.append(f, 2);
    char type = f[3];
    const char *v = f + 5;
    size_t vn = n - 5;
    std::string val(v, vn);
    switch (type) {
        case 'A':
            d.push_back('A');
            d.push_back(vn ? v[0] : 0);
            break;
        case 'i': {
            errno = 0;
            long long x = strtoll(val.c_str(), nullptr, 10);
            if (x < 0) {
                if (x >= -128) { d.push_back('c'); int8_t y = (int8_t)x; put_bytes(d, &y, 1); }
                else if (x >= -32768) { d.push_back('s'); int16_t y = (int16_t)x; put_bytes(d, &y, 2); }
                else { d.push_back('i'); int32_t y = (int32_t)x; put_bytes(d, &y, 4); }
            } else {
                if (x <= 255) { d.push_back('C'); uint8_t y = (uint8_t)x; put_bytes(d, &y, 1); }
                else if (x <= 65535) { d.push_back('S'); uint16_t y = (uint16_t)x; put_bytes(d, &y, 2); }
                else { d.push_back('I'); uint32_t y = (uint32_t)x; put_bytes(d, &y, 4); }
            }
            break;
        }
        case 'f': {
            float y = strtof(val.c_str(), nullptr);
            d.push_back('f');
            put_bytes(d, &y, 4);
            break;
        }
        case 'd': {
            double y = strtod(val.c_str(), nullptr);
            d.push_back('d');
            put_bytes(d, &y, 8);
            break;
        }
        case 'Z':
        case 'H':
            d.push_back(type);
            d.append(v, vn);
            d.push_back(0);
            break;
        case 'B': {
            if (vn < 1) return false;
            char sub = v[0];
            int32_t cnt = 0;
            for (size_t i = 1; i < vn; i++)
                if (v[i] == ',') cnt++;
            d.push_back('B');
            d.push_back(sub);
            put_bytes(d, &cnt, 4);
            const char *p = val.c_str() + 1;
            for (int32_t i = 0; i < cnt; i++) {
                ++p;  /* skip ',' */
                char *q;
                switch (sub) {
                    case 'c': { int8_t y = (int8_t)strtol(p, &q, 10); put_bytes(d, &y, 1); break; }
                    case 'C': { uint8_t y = (uint8_t)strtoul(p, &q, 10); put_bytes(d, &y, 1); break; }
                    case 's': { int16_t y = (int16_t)strtol(p, &q, 10); put_bytes(d, &y, 2); break; }
                    case 'S': { uint16_t y = (uint16_t)strtoul(p, &q, 10); put_bytes(d, &y, 2); break; }
                    case 'i': { int32_t y = (int32_t)strtol(p, &q, 10); put_bytes(d, &y, 4); break; }
                    case 'I': { uint32_t y = (uint32_t)strtoul(p, &q, 10); put_bytes(d, &y, 4); break; }
                    case 'f': { float y = strtof(p, &q); put_bytes(d, &y, 4); break; }
                    default: return false;
                }
                p = q;
            }
            break;
        }
        default:
            return false;
    }
    return true;
}

static int sam_parse_line(ipx_hts_impl *im, sam_hdr_t *h, const std::string &line, bam1_t *b) {
    /* split the 11 mandatory fields */
    const char *s = line.c_str();
    const char *e = s + line.size();
    const char *fld[11];
    size_t fln[11];
    const char *p = s;
    for (int i = 0; i < 11; i++) {
        if (p > e) return -2;
        const char *t = (const char *)memchr(p, '\t', e - p);
        if (!t) {
            if (i < 10) return -2;
            t = e;
        }
        fld[i] = p;
        fln[i] = t - p;
        p = t + 1;
    }
    if (!im->name_map_built) {
        for (int i = 0; i < h->n_targets; i++) im->name2tid.emplace(h->target_name[i], i);
        im->name_map_built = true;
    }
    bam1_core_t *c = &b->core;
    std::string &d = im->scratch;
    d.clear();
    /* qname, NUL padded to 4 bytes */
    d.append(fld[0], fln[0]);
    d.push_back(0);
    c->l_extranul = 0;
    while (d.size() % 4 != 0) { d.push_back(0); c->l_extranul++; }
    c->l_qname = (uint16_t)d.size();
    c->flag = (uint16_t)strtoul(fld[1], nullptr, 0);
    auto lookup = [&](const char *nm, size_t n) -> int {
        if (n == 1 && nm[0] == '*') return -1;
        auto it = im->name2tid.find(std::string(nm, n));
        return it == im->name2tid.end() ? -1 : it->second;
    };
    c->tid = lookup(fld[2], fln[2]);
    c->pos = strtoll(fld[3], nullptr, 10) - 1;
    /* htslib's sam_parse1: a mapped query cannot have a zero coordinate -- it is treated as unmapped; a record without a
     * (known) reference is flagged unmapped */
    if (c->pos < 0 && c->tid >= 0) c->tid = -1;
    if (c->tid < 0) c->flag |= BAM_FUNMAP;
    c->qual = (uint8_t)strtoul(fld[4], nullptr, 10);
    /* cigar */
    c->n_cigar = 0;
    if (!(fln[5] == 1 && fld[5][0] == '*')) {
        const char *q = fld[5];
        const char *qe = fld[5] + fln[5];
        while (q < qe) {
            char *r;
            unsigned long len = strtoul(q, &r, 10);
            if (r == q || r >= qe) return -2;
            const char *ops = BAM_CIGAR_STR;
            const char *o = strchr(ops, *r);
            if (!o || !*r) return -2;
            uint32_t w = (uint32_t)(len << BAM_CIGAR_SHIFT) | (uint32_t)(o - ops);
            put_bytes(d, &w, 4);
            c->n_cigar++;
            q = r + 1;
        }
    }
    if (fln[6] == 1 && fld[6][0] == '=') c->mtid = c->tid;
    else c->mtid = lookup(fld[6], fln[6]);
    c->mpos = strtoll(fld[7], nullptr, 10) - 1;
    if (c->mpos < 0 && c->mtid >= 0) c->mtid = -1; /* (the same for the mate) */
    c->isize = strtoll(fld[8], nullptr, 10);
    /* seq / qual */
    if (fln[9] == 1 && fld[9][0] == '*') {
        c->l_qseq = 0;
    } else {
        const uint8_t *tab = nt16_table();
        c->l_qseq = (int32_t)fln[9];
        size_t off = d.size();
        d.resize(off + (fln[9] + 1) / 2, 0);
        for (size_t i = 0; i < fln[9]; i++)
            d[off + (i >> 1)] |= (char)(tab[(uint8_t)fld[9][i]] << ((~i & 1) << 2));
        if (fln[10] == 1 && fld[10][0] == '*' && fln[9] != 1) {
            d.append(fln[9], (char)0xff);
        } else if (fln[10] == 1 && fld[10][0] == '*') {
            d.append(1, (char)0xff);
        } else {
            if (fln[10] != fln[9]) return -2;
            size_t qoff = d.size();
            d.append(fld[10], fln[10]);
            for (size_t i = 0; i < fln[10]; i++) d[qoff + i] = (char)(d[qoff + i] - 33);
        }
    }
    /* aux */
    while (p < e) {
        const char *t = (const char *)memchr(p, '\t', e - p);
        if (!t) t = e;
        if (t > p && !aux_text_to_bin(p, t - p, d)) return -2;
        p = t + 1;
    }
    if (reserve_data(b, d.size()) < 0) return -2;
    memcpy(b->data, d.data(), d.size());
    b->l_data = (int)d.size();
    hts_pos_t rlen = 0;
    if (c->n_cigar && !(c->flag & BAM_FUNMAP)) rlen = bam_cigar2rlen(c->n_cigar, bam_get_cigar(b));
    if (rlen == 0) rlen = 1;
    c->bin = (uint16_t)reg2bin(c->pos, c->pos + rlen);
    return 0;
}

extern "C" int sam_read1(samFile *fp, sam_hdr_t *h, bam1_t *b) {
    ipx_hts_impl *im = fp->impl;
    if (fp->format.format == bam) return bam_read_record(im, b);
    std::string line;
    for (;;) {
        if (!im->rd.getline(line)) return im->rd.err ? -2 : -1;
        if (!line.empty()) break;
    }
    fp->lineno++;
    return sam_parse_line(im, h, line, b);
}

static void append_int(std::string &s, long long v) {
    char tmp[32];
    int n = snprintf(tmp, sizeof(tmp), "%lld", v);
    s.append(tmp, n);
}

static void append_float(std::string &s, double v) {
    char tmp[64];
    int n = snprintf(tmp, sizeof(tmp), "%g", v);
    s.append(tmp, n);
}

/* size in bytes of one aux value starting at its type byte; 0 if malformed */
static size_t aux_value_size(const uint8_t *t, const uint8_t *end) {
    if (t >= end) return 0;
    switch (*t) {
        case 'A': case 'c': case 'C': return 2;
        case 's': case 'S': return 3;
        case 'i': case 'I': case 'f': return 5;
        case 'd': return 9;
        case 'Z': case 'H': {
            const uint8_t *z = (const uint8_t *)memchr(t + 1, 0, end - (t + 1));
            return z ? (size_t)(z - t) + 1 : 0;
        }
        case 'B': {
            if (end - t < 6) return 0;
            uint32_t n;
            memcpy(&n, t + 2, 4);
            size_t es;
            switch (t[1]) {
                case 'c': case 'C': es = 1; break;
                case 's': case 'S': es = 2; break;
                case 'i': case 'I': case 'f': es = 4; break;
                default: return 0;
            }
            return 6 + (size_t)n * es;
        }
        default: return 0;
    }
}

static int sam_format_record(const sam_hdr_t *h, const bam1_t *b, std::string &s) {
    const bam1_core_t *c = &b->core;
    s.clear();
    s.append(bam_get_qname(b));
    s.push_back('\t');
    append_int(s, c->flag);
    s.push_back('\t');
    if (c->tid >= 0 && c->tid < h->n_targets) s.append(h->target_name[c->tid]); else s.push_back('*');
    s.push_back('\t');
    append_int(s, c->pos + 1);
    s.push_back('\t');
    append_int(s, c->qual);
    s.push_back('\t');
    if (c->n_cigar == 0) {
        s.push_back('*');
    } else {
        const uint32_t *cig = bam_get_cigar(b);
        for (uint32_t i = 0; i < c->n_cigar; i++) {
            append_int(s, bam_cigar_oplen(cig[i]));
            s.push_back(bam_cigar_opchr(cig[i]));
        }
    }
    s.push_back('\t');
    if (c->mtid < 0) s.push_back('*');
    else if (c->mtid == c->tid) s.push_back('=');
    else if (c->mtid < h->n_targets) s.append(h->target_name[c->mtid]);
    else s.push_back('*');
    s.push_back('\t');
    append_int(s, c->mpos + 1);
    s.push_back('\t');
    append_int(s, c->isize);
    s.push_back('\t');
    if (c->l_qseq == 0) {
        s.append("*\t*");
    } else {
        const uint8_t *seq = bam_get_seq(b);
        for (int i = 0; i < c->l_qseq; i++) s.push_back("=ACMGRSVTWYHKDBN"[bam_seqi(seq, i)]);
        s.push_back('\t');
        const uint8_t *q = bam_get_qual(b);
        if (q[0] == 0xff) s.push_back('*');
        else for (int i = 0; i < c->l_qseq; i++) s.push_back((char)(q[i] + 33));
    }
    const uint8_t *a = bam_get_aux(b);
    const uint8_t *end = b->data + b->l_data;
    while (a + 2 < end) {
        const uint8_t *t = a + 2;
        size_t vs = aux_value_size(t, end);
        if (vs == 0 || t + vs > end) return -1;
        s.push_back('\t');
        s.append((const char *)a, 2);
        s.push_back(':');
        switch (*t) {
            case 'A': s.append("A:"); s.push_back((char)t[1]); break;
            case 'c': s.append("i:"); append_int(s, (int8_t)t[1]); break;
            case 'C': s.append("i:"); append_int(s, t[1]); break;
            case 's': { int16_t v; memcpy(&v, t + 1, 2); s.append("i:"); append_int(s, v); break; }
            case 'S': { uint16_t v; memcpy(&v, t + 1, 2); s.append("i:"); append_int(s, v); break; }
            case 'i': { int32_t v; memcpy(&v, t + 1, 4); s.append("i:"); append_int(s, v); break; }
            case 'I': { uint32_t v; memcpy(&v, t + 1, 4); s.append("i:"); append_int(s, v); break; }
            case 'f': { float v; memcpy(&v, t + 1, 4); s.append("f:"); append_float(s, v); break; }
            case 'd': { double v; memcpy(&v, t + 1, 8); s.append("d:"); append_float(s, v); break; }
            case 'Z': s.append("Z:"); s.append((const char *)t + 1); break;
            case 'H': s.append("H:"); s.append((const char *)t + 1); break;
            case 'B': {
                uint32_t n;
                memcpy(&n, t + 2, 4);
                s.append("B:");
                s.push_back((char)t[1]);
                const uint8_t *v = t + 6;
                for (uint32_t i = 0; i < n; i++) {
                    s.push_back(',');
                    switch (t[1]) {
                        case 'c': append_int(s, (int8_t)v[i]); break;
                        case 'C': append_int(s, v[i]); break;
                        case 's': { int16_t y; memcpy(&y, v + 2 * i, 2); append_int(s, y); break; }
                        case 'S': { uint16_t y; memcpy(&y, v + 2 * i, 2); append_int(s, y); break; }
                        case 'i': { int32_t y; memcpy(&y, v + 4 * i, 4); append_int(s, y); break; }
                        case 'I': { uint32_t y; memcpy(&y, v + 4 * i, 4); append_int(s, y); break; }
                        case 'f': { float y; memcpy(&y, v + 4 * i, 4); append_float(s, y); break; }
                    }
                }
                break;
            }
        }
        a = t + vs;
    }
    s.push_back('\n');
    return 0;
}

extern "C" int sam_write1(samFile *fp, const sam_hdr_t *h, const bam1_t *b) {
    ipx_hts_impl *im = fp->impl;
    if (fp->format.format == sam) {
        if (sam_format_record(h, b, im->scratch) < 0) return -1;
        if (fwrite(im->scratch.data(), 1, im->scratch.size(), im->out) != im->scratch.size()) return -1;
        return (int)im->scratch.size();
    }
    const bam1_core_t *c = &b->core;
    uint32_t l_qname = c->l_qname - c->l_extranul;
    int32_t block_len = 32 + b->l_data - c->l_extranul;
    std::string &d = im->scratch;
    d.clear();
    uint32_t n_cigar = c->n_cigar;
    if (n_cigar > 0xffff) return -1;  /* writing CG-tag long cigars is not needed by this path */
    int32_t x[8];
    x[0] = c->tid;
    x[1] = (int32_t)c->pos;
    x[2] = (int32_t)(((uint32_t)c->bin << 16) | ((uint32_t)c->qual << 8) | l_qname);
    x[3] = (int32_t)(((uint32_t)c->flag << 16) | n_cigar);
    x[4] = c->l_qseq;
    x[5] = c->mtid;
    x[6] = (int32_t)c->mpos;
    x[7] = (int32_t)c->isize;
    put_bytes(d, &block_len, 4);
    put_bytes(d, x, 32);
    put_bytes(d, b->data, l_qname);
    put_bytes(d, b->data + c->l_qname, b->l_data - c->l_qname);
    if (im->raw_bam) return fwrite(d.data(), 1, d.size(), im->out) == d.size() ? (int)d.size() : -1;
    return im->bgzf.write(d.data(), d.size()) ? (int)d.size() : -1;
}

/* ---------- aux access ---------- */
extern "C" uint8_t *bam_aux_get(const bam1_t *b, const char tag[2]) {
    uint8_t *a = bam_get_aux(b);
    const uint8_t *end = b->data + b->l_data;
    while (a + 2 < end) {
        uint8_t *t = a + 2;
        size_t vs = aux_value_size(t, end);
        if (vs == 0 || t + vs > end) return nullptr;
        if (a[0] == (uint8_t)tag[0] && a[1] == (uint8_t)tag[1]) return t;
        a = t + vs;
    }
    return nullptr;
}

extern "C" int64_t bam_aux2i(const uint8_t *s) {
    switch (*s) {
        case 'c': return (int8_t)s[1];
        case 'C': return s[1];
        case 's': { int16_t v; memcpy(&v, s + 1, 2); return v; }
        case 'S': { uint16_t v; memcpy(&v, s + 1, 2); return v; }
        case 'i': { int32_t v; memcpy(&v, s + 1, 4); return v; }
        case 'I': { uint32_t v; memcpy(&v, s + 1, 4); return v; }
        default: errno = EINVAL; return 0;
    }
}

extern "C" double bam_aux2f(const uint8_t *s) {
    switch (*s) {
        case 'f': { float v; memcpy(&v, s + 1, 4); return v; }
        case 'd': { double v; memcpy(&v, s + 1, 8); return v; }
        default: return (double)bam_aux2i(s);
    }
}

extern "C" char bam_aux2A(const uint8_t *s) {
    if (*s == 'A') return (char)s[1];
    errno = EINVAL;
    return 0;
}

extern "C" char *bam_aux2Z(const uint8_t *s) {
    if (*s == 'Z' || *s == 'H') return (char *)(s + 1);
    errno = EINVAL;
    return nullptr;
}

extern "C" int bam_aux_append(bam1_t *b, const char tag[2], char type, int len, const uint8_t *data) {
    size_t n = (size_t)b->l_data + 3 + len;
    if (reserve_data(b, n) < 0) return -1;
    uint8_t *p = b->data + b->l_data;
    p[0] = tag[0];
    p[1] = tag[1];
    p[2] = type;
    memcpy(p + 3, data, len);
    b->l_data = (int)n;
    return 0;
}

extern "C" int bam_aux_del(bam1_t *b, uint8_t *s) {
    const uint8_t *end = b->data + b->l_data;
    size_t vs = aux_value_size(s, end);
    if (vs == 0) return -1;
    uint8_t *tag_beg = s - 2;
    uint8_t *tag_end = s + vs;
    memmove(tag_beg, tag_end, end - tag_end);
    b->l_data -= (int)(tag_end - tag_beg);
    return 0;
}

/* ---------- FASTA index ---------- */
struct fai_entry {
    std::string name;
    int64_t len, offset;
    int32_t line_blen, line_len;
};

struct faidx_t {
    FILE *fp = nullptr;
    std::vector<fai_entry> seqs;
    std::unordered_map<std::string, int> by_name;
    /* a bgzip-compressed FASTA (fn.gz + fn.gz.fai + fn.gz.gzi, the form the reference's documentation uses): inflated
     * whole -- the .fai offsets count uncompressed bytes, and the one caller here reads every sequence anyway */
    std::vector<unsigned char> mem;
    bool in_memory = false;
};

/* every member of a gzip / BGZF file, inflated into `out`; false when the data is damaged */
static bool inflate_members(FILE *in, std::vector<unsigned char> &out) {
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, 15 + 32) != Z_OK) return false;
    std::vector<unsigned char> ibuf(1 << 20), obuf(4 << 20);
    bool ok = true, member_open = false;
    for (;;) {
        const size_t n = fread(ibuf.data(), 1, ibuf.size(), in);
        if (n == 0) break;
        zs.next_in = ibuf.data();
        zs.avail_in = (uInt)n;
        while (zs.avail_in > 0) {
            zs.next_out = obuf.data();
            zs.avail_out = (uInt)obuf.size();
            member_open = true;
            const int rc = inflate(&zs, Z_NO_FLUSH);
            out.insert(out.end(), obuf.data(), obuf.data() + (obuf.size() - zs.avail_out));
            if (rc == Z_STREAM_END) {
                member_open = false;
                if (inflateReset(&zs) != Z_OK) { ok = false; break; }
            } else if (rc != Z_OK && rc != Z_BUF_ERROR) {
                ok = false;
                break;
            } else if (rc == Z_BUF_ERROR && zs.avail_out != 0) {
                break; /* needs more input */
            }
        }
        if (!ok) break;
    }
    inflateEnd(&zs);
    return ok && !member_open;
}

extern "C" int fai_build(const char *fn) {
    FILE *in = fopen(fn, "rb");
    if (!in) return -1;
    std::string out_fn = std::string(fn) + ".fai";
    std::vector<fai_entry> es;
    fai_entry cur;
    bool have = false;
    int64_t off = 0;
    char *line = nullptr;
    size_t cap = 0;
    ssize_t n;
    while ((n = getline(&line, &cap, in)) >= 0) {
        int64_t line_off = off;
        off += n;
        ssize_t blen = n;
        while (blen > 0 && (line[blen - 1] == '\n' || line[blen - 1] == '\r')) blen--;
        if (n > 0 && line[0] == '>') {
            if (have) es.push_back(cur);
            cur = fai_entry();
            size_t k = 1;
            while (k < (size_t)blen && !isspace((unsigned char)line[k])) k++;
            cur.name.assign(line + 1, k - 1);
            cur.len = 0;
            cur.offset = off;
            cur.line_blen = 0;
            cur.line_len = 0;
            have = true;
        } else if (have) {
            if (cur.line_len == 0 && blen > 0) {
                cur.line_blen = (int32_t)blen;
                cur.line_len = (int32_t)n;
            }
            cur.len += blen;
        }
        (void)line_off;
    }
    if (have) es.push_back(cur);
    free(line);
    fclose(in);
    FILE *out = fopen(out_fn.c_str(), "w");
    if (!out) return -1;
    for (const fai_entry &e : es)
        fprintf(out, "%s\t%lld\t%lld\t%d\t%d\n", e.name.c_str(), (long long)e.len, (long long)e.offset, e.line_blen, e.line_len);
    fclose(out);
    return 0;
}

extern "C" faidx_t *fai_load3(const char *fn, const char *fnfai, const char *fngzi, int flags) {
    (void)fngzi;
    std::string idx = fnfai ? std::string(fnfai) : std::string(fn) + ".fai";
    FILE *fi = fopen(idx.c_str(), "r");
    if (!fi) {
        if (!(flags & FAI_CREATE)) {
            fprintf(stderr, "[E::fai_load3_core] Failed to open FASTA index %s: %s\n", idx.c_str(), strerror(errno));
            return nullptr;
        }
        if (fai_build(fn) < 0) return nullptr;
        fi = fopen(idx.c_str(), "r");
        if (!fi) return nullptr;
    }
    faidx_t *fai = new faidx_t();
    char *line = nullptr;
    size_t cap = 0;
    while (getline(&line, &cap, fi) >= 0) {
        char *save = nullptr;
        char *nm = strtok_r(line, "\t", &save);
        char *l = strtok_r(nullptr, "\t", &save);
        char *o = strtok_r(nullptr, "\t", &save);
        char *lb = strtok_r(nullptr, "\t", &save);
        char *lw = strtok_r(nullptr, "\t\n", &save);
        if (!nm || !l || !o || !lb || !lw) continue;
        fai_entry e;
        e.name = nm;
        e.len = strtoll(l, nullptr, 10);
        e.offset = strtoll(o, nullptr, 10);
        e.line_blen = (int32_t)strtol(lb, nullptr, 10);
        e.line_len = (int32_t)strtol(lw, nullptr, 10);
        fai->by_name.emplace(e.name, (int)fai->seqs.size());
        fai->seqs.push_back(e);
    }
    free(line);
    fclose(fi);
    fai->fp = fopen(fn, "rb");
    if (!fai->fp) {
        delete fai;
        return nullptr;
    }
    unsigned char magic[18] = {0};
    const size_t got = fread(magic, 1, sizeof(magic), fai->fp);
    rewind(fai->fp);
    if (got >= 2 && magic[0] == 0x1f && magic[1] == 0x8b) {
        /* htslib takes compressed FASTA only in BGZF form, and only with its .gzi index beside it */
        const bool is_bgzf = got >= 18 && (magic[3] & 4) && magic[12] == 'B' && magic[13] == 'C';
        if (!is_bgzf) {
            fprintf(stderr, "[E::fai_load3_core] Cannot index files compressed with gzip, please use bgzip\n");
            fai_destroy(fai);
            return nullptr;
        }
        const std::string gzi = fngzi ? std::string(fngzi) : std::string(fn) + ".gzi";
        FILE *fg = fopen(gzi.c_str(), "rb");
        if (!fg) {
            fprintf(stderr, "[E::fai_load3_core] Failed to open BGZF index %s: %s\n", gzi.c_str(), strerror(errno));
            fai_destroy(fai);
            return nullptr;
        }
        fclose(fg);
        if (!inflate_members(fai->fp, fai->mem)) {
            fprintf(stderr, "[E::fai_load3_core] Failed to inflate %s\n", fn);
            fai_destroy(fai);
            return nullptr;
        }
        fai->in_memory = true;
    }
    return fai;
}

extern "C" int ipx_fai_in_memory(const faidx_t *fai) { return fai && fai->in_memory ? 1 : 0; }

extern "C" faidx_t *fai_load(const char *fn) {
    return fai_load3(fn, nullptr, nullptr, FAI_CREATE);
}

extern "C" void fai_destroy(faidx_t *fai) {
    if (!fai) return;
    if (fai->fp) fclose(fai->fp);
    delete fai;
}

extern "C" int faidx_has_seq(const faidx_t *fai, const char *seq) {
    return fai->by_name.count(seq) ? 1 : 0;
}

extern "C" int faidx_nseq(const faidx_t *fai) {
    return (int)fai->seqs.size();
}

extern "C" const char *faidx_iseq(const faidx_t *fai, int i) {
    return fai->seqs[i].name.c_str();
}

extern "C" int faidx_seq_len(const faidx_t *fai, const char *seq) {
    auto it = fai->by_name.find(seq);
    return it == fai->by_name.end() ? -1 : (int)fai->seqs[it->second].len;
}

extern "C" char *faidx_fetch_seq(const faidx_t *fai, const char *c_name, int p_beg_i, int p_end_i, int *len) {
    auto it = fai->by_name.find(c_name);
    if (it == fai->by_name.end()) {
        fprintf(stderr, "[W::fai_get_val] Reference %s not found in file, returning empty sequence\n", c_name);
        *len = -2;
        return nullptr;
    }
    const fai_entry &e = fai->seqs[it->second];
    int64_t beg = p_beg_i, end = p_end_i;
    if (end < beg - 1) end = beg - 1;
    if (beg < 0) beg = 0;
    if (beg >= e.len) beg = e.len;
    if (end < 0) end = 0; else if (end >= e.len) end = e.len - 1;
    int64_t want = end - beg + 1;
    if (want < 0) want = 0;
    char *s = (char *)malloc((size_t)want + 1);
    int64_t got = 0;
    if (want > 0 && e.line_blen > 0) {
        int64_t off = e.offset + beg / e.line_blen * e.line_len + beg % e.line_blen;
        if (fai->in_memory) {
            for (size_t i = (size_t)std::max<int64_t>(off, 0); i < fai->mem.size() && got < want; i++) {
                const unsigned char ch = fai->mem[i];
                s[got] = (char)ch;
                got += isgraph(ch) ? 1 : 0;
            }
            s[got] = 0;
            *len = (int)got;
            return s;
        }
        if (fseeko(fai->fp, off, SEEK_SET) != 0) { free(s); *len = -1; return nullptr; }
        /* the bases are what is left of the bytes from `off` on once non-printing characters (line ends) are dropped;
         * read in blocks -- a getc per base costs tens of nanoseconds, a whole genome minutes */
        const int64_t with_line_ends = want + (want / e.line_blen + 2) * (e.line_len - e.line_blen) + 16; /* small fetches stay small */
        std::vector<unsigned char> block((size_t)std::min<int64_t>(1 << 20, with_line_ends));
        while (got < want) {
            const size_t n = fread(block.data(), 1, block.size(), fai->fp);
            if (n == 0) break;
            for (size_t i = 0; i < n && got < want; i++) {
                const unsigned char ch = block[i];
                s[got] = (char)ch; /* one byte of slack past `want` is the terminator's */
                got += isgraph(ch) ? 1 : 0;
            }
        }
    }
    s[got] = 0;
    *len = (int)got;
    return s;
}
