/* bam_pipeline.h -- BAM file -> packed record columns on every host core (host side of ipx_push_batch_compact).
 *
 * The reference reads its input one record at a time through sam_read1 (read_align, junctions_extractor.cc:620-634).
 * Once the junction path runs at device speed that loop is what a user of the tool waits for, so the default input --
 * a BGZF-compressed BAM file -- goes through a pipeline that keeps every core busy:
 *
 *   reader thread     reads the file and cuts it into chunks of whole BGZF blocks (block headers carry their size)
 *   worker threads    inflate a chunk's blocks into one contiguous buffer            (BGZF blocks are independent)
 *   framer thread     walks the record sizes of the chunks in stream order -- the one sequential step, a 4-byte read
 *                     per record -- so that every chunk knows where its first record starts; a record that straddles
 *                     chunks is copied aside for the chunk it starts in
 *   worker threads    pack a chunk's records into columns: position, flag, CIGAR words, NH class and XS code (the tags
 *                     are only looked at for reads with an N operation), runs of equal target index
 *   consumer          takes the chunks back in stream order (next) and appends their columns to the pinned batch
 *
 * Nothing here interprets records beyond what the packer of record_source.h does; run-mode logic (which records are
 * processed, the sortedness error) stays with the caller, who sees the stream as runs of equal target index.
 */
#ifndef IPX_BAM_PIPELINE_H
#define IPX_BAM_PIPELINE_H
#include <stdint.h>
#include <stdio.h>

#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace ipx {

/* consecutive records with the same target index, as they lie in a chunk's columns */
struct TidRun {
    int32_t tid;
    uint32_t first, n;        /* records [first, first + n) of the chunk */
    uint64_t cigar_first, n_cigar; /* their CIGAR words in the chunk's cigar column */
};

struct PackedChunk {
    /* columns of the records that start in this chunk, stream order */
    std::vector<int32_t> pos;
    std::vector<uint16_t> flag;
    std::vector<uint8_t> nh_class, xs;
    std::vector<uint32_t> n_ops; /* CIGAR words per record (32 bits: a CG:B:I tag may hold more than 65535) */
    std::vector<uint32_t> cigar;
    std::vector<TidRun> runs;
    uint32_t max_ops = 0;
    bool bad = false; /* the stream is damaged at the end of this chunk's records: report after them */
    size_t size() const { return pos.size(); }
};

class BamPipeline {
   public:
    BamPipeline();
    ~BamPipeline();
    /* `path`: a BAM file, BGZF-compressed or not; `skip` = bytes of the uncompressed stream before the first record
     * (the header, which the caller has parsed).  threads <= 0: one per core. */
    bool open(const char *path, uint64_t skip, int threads);
    /* the next chunk in stream order, or NULL at the end of the file; it stays valid until release() */
    PackedChunk *next();
    void release(PackedChunk *c);
    void close();

   private:
    struct Slot;
    void reader_main();
    void worker_main();
    void framer_main();
    void inflate_slot(Slot &s);
    void pack_slot(Slot &s);
    bool take(void *dst, size_t n, size_t *got);

    FILE *fp_ = nullptr;
    bool bgzf_ = false;
    uint64_t skip_ = 0;
    std::vector<Slot *> ring_;
    std::vector<std::thread> threads_;
    std::mutex mu_;
    std::condition_variable cv_;
    uint64_t n_read_ = 0;     /* chunks the reader has filled */
    uint64_t n_inflate_ = 0;  /* chunks handed to inflate workers */
    uint64_t n_framed_ = 0;   /* chunks the framer has finished */
    uint64_t n_pack_ = 0;     /* chunks handed to pack workers */
    uint64_t n_taken_ = 0;    /* chunks given to the consumer */
    uint64_t n_released_ = 0; /* chunks the consumer is done with */
    bool read_done_ = false, read_bad_ = false, stop_ = false;
};

}  // namespace ipx
#endif
