/* bam_pipeline.cc -- see bam_pipeline.h */
#include "bam_pipeline.h"

#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <algorithm>

#include "record_source.h"

namespace ipx {

static const size_t CHUNK_COMPRESSED = 1u << 20; /* compressed bytes of BGZF blocks per chunk (2 - 4 MB inflated) */
static const size_t CHUNK_RAW = 4u << 20;        /* bytes per chunk of an uncompressed BAM stream */

struct BamPipeline::Slot {
    std::vector<uint8_t> in;   /* BGZF blocks as read */
    std::vector<uint8_t> data; /* the chunk's stretch of the uncompressed stream */
    size_t data_len = 0;
    size_t first = 0, tail = 0;      /* whole records that start in this chunk lie in data[first, tail) */
    std::vector<uint8_t> straddle;   /* the record that starts at `tail` and ends in a later chunk, complete */
    bool inflated = false, packed = false;
    bool bad_in = false;             /* a block of this chunk did not inflate */
    bool bad_after = false;          /* the stream is damaged / truncated after this chunk's records */
    PackedChunk out;
};

BamPipeline::BamPipeline() {}
BamPipeline::~BamPipeline() { close(); }

bool BamPipeline::take(void *dst, size_t n, size_t *got) {
    *got = fread(dst, 1, n, fp_);
    return *got == n;
}

bool BamPipeline::open(const char *path, uint64_t skip, int threads) {
    fp_ = fopen(path, "rb");
    if (!fp_) return false;
    setvbuf(fp_, nullptr, _IOFBF, 1 << 22);
    uint8_t h[16];
    const size_t n = fread(h, 1, sizeof(h), fp_);
    bgzf_ = n == 16 && h[0] == 0x1f && h[1] == 0x8b && h[2] == 8 && (h[3] & 4) && h[12] == 'B' && h[13] == 'C';
    if (fseek(fp_, 0, SEEK_SET) != 0) {
        fclose(fp_);
        fp_ = nullptr;
        return false;
    }
    skip_ = skip;
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    if (threads < 1) threads = 1;
    if (threads > 64) threads = 64;
    ring_.resize((size_t)threads * 2 + 4);
    for (auto &s : ring_) s = new Slot();
    threads_.emplace_back([this] { reader_main(); });
    threads_.emplace_back([this] { framer_main(); });
    for (int i = 0; i < threads; i++) threads_.emplace_back([this] { worker_main(); });
    return true;
}

void BamPipeline::close() {
    {
        std::lock_guard<std::mutex> g(mu_);
        stop_ = true;
    }
    cv_.notify_all();
    for (auto &t : threads_) t.join();
    threads_.clear();
    for (auto *s : ring_) delete s;
    ring_.clear();
    if (fp_) fclose(fp_);
    fp_ = nullptr;
}

/* ---- reader: chunks of whole BGZF blocks (or of raw bytes), in file order ---- */
void BamPipeline::reader_main() {
    bool eof = false;
    while (!eof) {
        Slot *s;
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [&] { return stop_ || n_read_ - n_released_ < ring_.size(); });
            if (stop_) return;
            s = ring_[n_read_ % ring_.size()];
        }
        s->in.clear();
        s->inflated = s->packed = s->bad_in = s->bad_after = false;
        s->straddle.clear();
        bool bad = false;
        if (bgzf_) {
            size_t inflated = 0;
            while (s->in.size() < CHUNK_COMPRESSED) {
                uint8_t hdr[18];
                size_t got;
                if (!take(hdr, 18, &got)) {
                    if (got != 0) bad = true; /* the file ends inside a block header */
                    eof = true;
                    break;
                }
                const bool ok = hdr[0] == 0x1f && hdr[1] == 0x8b && hdr[2] == 8 && (hdr[3] & 4) && hdr[12] == 'B' && hdr[13] == 'C';
                const size_t bsize = ok ? (size_t)(hdr[16] | (hdr[17] << 8)) + 1 : 0;
                if (!ok || bsize < 26) {
                    bad = eof = true;
                    break;
                }
                const size_t at = s->in.size();
                s->in.resize(at + bsize);
                memcpy(s->in.data() + at, hdr, 18);
                if (!take(s->in.data() + at + 18, bsize - 18, &got)) {
                    s->in.resize(at);
                    bad = eof = true;
                    break;
                }
                uint32_t isize;
                memcpy(&isize, s->in.data() + at + bsize - 4, 4);
                inflated += isize;
            }
            s->data_len = inflated;
        } else {
            s->data.resize(CHUNK_RAW);
            size_t got;
            if (!take(s->data.data(), CHUNK_RAW, &got)) eof = true;
            s->data_len = got;
            s->inflated = true;
        }
        std::lock_guard<std::mutex> g(mu_);
        if (bad) read_bad_ = true;
        if (s->data_len > 0 || !s->in.empty()) n_read_++;
        if (eof) read_done_ = true;
        cv_.notify_all();
    }
}

void BamPipeline::inflate_slot(Slot &s) {
    if (s.data.size() < s.data_len) s.data.resize(s.data_len);
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, -15) != Z_OK) {
        s.bad_in = true;
        return;
    }
    size_t at = 0, out = 0;
    while (at < s.in.size()) {
        const size_t bsize = (size_t)(s.in[at + 16] | (s.in[at + 17] << 8)) + 1;
        uint32_t isize;
        memcpy(&isize, s.in.data() + at + bsize - 4, 4);
        if (isize) {
            inflateReset(&zs);
            zs.next_in = s.in.data() + at + 18;
            zs.avail_in = (uInt)(bsize - 18 - 8);
            zs.next_out = s.data.data() + out;
            zs.avail_out = (uInt)isize;
            const int rc = inflate(&zs, Z_FINISH);
            if (rc != Z_STREAM_END || zs.avail_out != 0) {
                /* what inflated before the damaged block still counts: the records in it are processed first */
                s.bad_in = true;
                break;
            }
            out += isize;
        }
        at += bsize;
    }
    inflateEnd(&zs);
    if (s.bad_in) s.data_len = out;
}

/* ---- framer: where the records of every chunk start ---- */
void BamPipeline::framer_main() {
    uint64_t skip = skip_;
    bool dead = false; /* the stream was damaged in an earlier chunk: nothing after it is looked at */
    for (uint64_t k = 0;; k++) {
        Slot *s;
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [&] { return stop_ || (k < n_read_ && ring_[k % ring_.size()]->inflated) || (read_done_ && k >= n_read_); });
            if (stop_ || k >= n_read_) return;
            s = ring_[k % ring_.size()];
        }
        const size_t len = s->data_len;
        s->first = s->tail = len;
        if (dead) {
            s->bad_after = false;
        } else if (skip >= len) {
            skip -= len; /* all of it belongs to the header or to a record that started earlier */
            if (s->bad_in) {
                s->bad_after = true;
                dead = true;
            }
        } else {
            size_t p = (size_t)skip;
            skip = 0;
            s->first = p;
            bool damaged = false;
            while (p + 4 <= len) {
                int32_t bs;
                memcpy(&bs, s->data.data() + p, 4);
                if (bs < 32) {
                    damaged = true;
                    break;
                }
                if (p + 4 + (size_t)bs > len) break;
                p += 4 + (size_t)bs;
            }
            s->tail = p;
            if (damaged || s->bad_in) {
                s->bad_after = true;
                dead = true;
            } else if (p < len) {
                /* a record starts here and ends in a later chunk: gather it */
                s->straddle.assign(s->data.begin() + (long)p, s->data.begin() + (long)len);
                size_t need = 0; /* total bytes of the record once its size is known */
                uint64_t m = k + 1;
                size_t used_of_m = 0;
                for (;;) {
                    if (s->straddle.size() >= 4 && need == 0) {
                        int32_t bs;
                        memcpy(&bs, s->straddle.data(), 4);
                        if (bs < 32) {
                            damaged = true;
                            break;
                        }
                        need = 4 + (size_t)bs;
                    }
                    if (need && s->straddle.size() >= need) break;
                    Slot *nx;
                    {
                        std::unique_lock<std::mutex> lk(mu_);
                        cv_.wait(lk, [&] { return stop_ || (m < n_read_ && ring_[m % ring_.size()]->inflated) || (read_done_ && m >= n_read_); });
                        if (stop_) return;
                        if (m >= n_read_) {
                            damaged = true; /* the file ends inside a record */
                            break;
                        }
                        nx = ring_[m % ring_.size()];
                    }
                    const size_t want = need ? need - s->straddle.size() : 4 - s->straddle.size();
                    const size_t avail = nx->data_len - used_of_m;
                    const size_t n = std::min(want, avail);
                    s->straddle.insert(s->straddle.end(), nx->data.begin() + (long)used_of_m, nx->data.begin() + (long)(used_of_m + n));
                    used_of_m += n;
                    skip += n;
                    if (used_of_m == nx->data_len) {
                        if (nx->bad_in && (need == 0 || s->straddle.size() < need)) {
                            damaged = true;
                            break;
                        }
                        m++;
                        used_of_m = 0;
                    }
                }
                if (damaged) {
                    s->straddle.clear();
                    s->bad_after = true;
                    dead = true;
                }
            }
        }
        std::lock_guard<std::mutex> g(mu_);
        n_framed_ = k + 1;
        cv_.notify_all();
    }
}

/* ---- pack: records -> columns (the same fields ColumnarBatch::add_raw takes) ---- */
static bool pack_record(const uint8_t *rec, size_t len, PackedChunk &o) {
    /* rec = the fixed 32-byte part onwards, len = block_size */
    int32_t tid, p, l_seq;
    uint16_t n_cigar16, fl;
    memcpy(&tid, rec, 4);
    memcpy(&p, rec + 4, 4);
    memcpy(&n_cigar16, rec + 12, 2);
    memcpy(&fl, rec + 14, 2);
    memcpy(&l_seq, rec + 16, 4);
    const size_t l_qname = rec[8];
    if (l_seq < 0 || len < 32 + l_qname) return false;
    const size_t fixed = 32 + l_qname + (size_t)n_cigar16 * 4 + (size_t)((l_seq + 1) / 2) + (size_t)l_seq;
    if (fixed > len) return false;
    const uint8_t *cig = rec + 32 + l_qname;
    const uint8_t *aux = rec + fixed, *end = rec + len;
    uint32_t n_cigar = n_cigar16;
    bool has_intron = false;
    if (n_cigar16 == 2 && tid >= 0 && p >= 0) {
        /* "<l_seq>S<ref_len>N": the real CIGAR may sit in a CG:B:I tag (more than 65535 operations) */
        uint32_t c0, c1;
        memcpy(&c0, cig, 4);
        memcpy(&c1, cig + 4, 4);
        if ((c0 & 15u) == 4u && (int32_t)(c0 >> 4) == l_seq && (c1 & 15u) == 3u) {
            const uint8_t *cg = find_aux(aux, end, 'C', 'G');
            if (cg && cg[0] == 'B' && (cg[1] == 'I' || cg[1] == 'i') && cg + 6 <= end) {
                uint32_t n;
                memcpy(&n, cg + 2, 4);
                if (n != 0 && cg + 6 + (size_t)n * 4 <= end) {
                    cig = cg + 6;
                    n_cigar = n;
                }
            }
        }
    }
    const size_t at = o.cigar.size();
    o.cigar.resize(at + n_cigar);
    memcpy(o.cigar.data() + at, cig, 4 * (size_t)n_cigar);
    if (n_cigar <= 16) {
        for (uint32_t k = 0; k < n_cigar; k++) has_intron |= (o.cigar[at + k] & 15u) == 3u;
    } else {
        has_intron = true; /* not worth a second pass: look at the tags */
    }
    /* NH and XS only ever matter for a read that yields a junction candidate, and that takes an N operation */
    uint8_t nh = IPX_NH_NONE, x = IPX_XS_ABSENT;
    if (has_intron) scan_tag_range(aux, end, &nh, &x);
    if (o.runs.empty() || o.runs.back().tid != tid) o.runs.push_back(TidRun{tid, (uint32_t)o.pos.size(), 0u, (uint64_t)at, 0ull});
    o.runs.back().n++;
    o.runs.back().n_cigar += n_cigar;
    o.pos.push_back(p);
    o.flag.push_back(fl);
    o.nh_class.push_back(nh);
    o.xs.push_back(x);
    o.n_ops.push_back(n_cigar);
    if (n_cigar > o.max_ops) o.max_ops = n_cigar;
    return true;
}

void BamPipeline::pack_slot(Slot &s) {
    PackedChunk &o = s.out;
    o.pos.clear(); o.flag.clear(); o.nh_class.clear(); o.xs.clear(); o.n_ops.clear(); o.cigar.clear(); o.runs.clear();
    o.max_ops = 0;
    o.bad = s.bad_after;
    size_t p = s.first;
    while (p < s.tail) {
        int32_t bs;
        memcpy(&bs, s.data.data() + p, 4);
        if (!pack_record(s.data.data() + p + 4, (size_t)bs, o)) {
            o.bad = true; /* a record whose fields contradict its size: the reader gives up here */
            return;
        }
        p += 4 + (size_t)bs;
    }
    if (!s.straddle.empty()) {
        if (!pack_record(s.straddle.data() + 4, s.straddle.size() - 4, o)) o.bad = true;
    }
}

void BamPipeline::worker_main() {
    for (;;) {
        Slot *s = nullptr;
        bool pack = false;
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [&] { return stop_ || n_pack_ < n_framed_ || (bgzf_ && n_inflate_ < n_read_) || (read_done_ && n_pack_ >= n_read_); });
            if (stop_) return;
            if (n_pack_ < n_framed_) {
                s = ring_[n_pack_ % ring_.size()];
                n_pack_++;
                pack = true;
            } else if (bgzf_ && n_inflate_ < n_read_) {
                s = ring_[n_inflate_ % ring_.size()];
                n_inflate_++;
            } else {
                return; /* everything read has been packed */
            }
        }
        if (pack) pack_slot(*s); else inflate_slot(*s);
        std::lock_guard<std::mutex> g(mu_);
        if (pack) s->packed = true; else s->inflated = true;
        cv_.notify_all();
    }
}

PackedChunk *BamPipeline::next() {
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [&] { return stop_ || (n_taken_ < n_read_ && ring_[n_taken_ % ring_.size()]->packed) || (read_done_ && n_taken_ >= n_read_); });
    if (stop_ || n_taken_ >= n_read_) {
        if (read_bad_ && !stop_) {
            /* the file ended in the middle of a block: report it once, after everything before it */
            read_bad_ = false;
            static thread_local PackedChunk damaged;
            damaged = PackedChunk();
            damaged.bad = true;
            return &damaged;
        }
        return nullptr;
    }
    Slot *s = ring_[n_taken_ % ring_.size()];
    n_taken_++;
    return &s->out;
}

void BamPipeline::release(PackedChunk *c) {
    std::lock_guard<std::mutex> g(mu_);
    for (auto *s : ring_)
        if (&s->out == c) {
            n_released_++;
            break;
        }
    cv_.notify_all();
}

}  // namespace ipx
