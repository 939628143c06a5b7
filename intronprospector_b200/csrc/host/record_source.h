/* record_source.h -- alignment file -> columnar batch packer (host side of ipx_push_batch). */
#ifndef IPX_RECORD_SOURCE_H
#define IPX_RECORD_SOURCE_H
#include <stdint.h>
#include <string>
#include <vector>
#include "htslib/sam.h"
#include "ipx.h"

namespace ipx {

/* The two aux tags the path looks at, found in one pass over the aux block.
 *   nh_class: get_num_aligns (junctions_extractor.cc:49-56) -- first NH tag; bam_aux2i yields 0 for non-integer types
 *   xs:       get_junction_strand_XS (junctions_extractor.cc:322-330) -- first XS tag; bam_aux2A yields 0 unless type 'A' */
void scan_tags(const bam1_t *b, uint8_t *nh_class, uint8_t *xs);
/* the same over the aux block [a, end) of a BAM record where it lies in a buffer */
void scan_tag_range(const uint8_t *a, const uint8_t *end, uint8_t *nh_class, uint8_t *xs);
/* the first tag named t0 t1 in an aux block: pointer to its type byte, or NULL */
const uint8_t *find_aux(const uint8_t *a, const uint8_t *end, char t0, char t1);

/* growable structure-of-arrays batch */
struct ColumnarBatch {
    std::vector<int32_t> tid, pos;
    std::vector<uint16_t> flag;
    std::vector<uint8_t> nh_class, xs;
    std::vector<uint32_t> cigar_off, cigar;
    /* the compact form (ipx_batch_compact), kept alongside: what goes over the host link when every record fits it */
    std::vector<uint16_t> n_cigar_ops;
    std::vector<int32_t> run_tid;
    std::vector<uint64_t> run_start;
    bool compact_ok = true; /* no record with more than 65535 operations so far */
    uint64_t seq_base = 0;
    void clear(uint64_t next_seq_base) {
        tid.clear(); pos.clear(); flag.clear(); nh_class.clear(); xs.clear(); cigar.clear();
        cigar_off.assign(1, 0u);
        n_cigar_ops.clear(); run_tid.clear(); run_start.clear();
        compact_ok = true;
        seq_base = next_seq_base;
    }
    size_t size() const { return tid.size(); }
    void add(const bam1_t *b, int32_t mapped_tid);
    /* the same straight from a BAM record in the reader's buffer (ipx_bam_peek: `rec` = the fixed 32-byte part,
     * `len` = block_size), without building a bam1_t: position, flag, CIGAR words, and a scan of the tags for NH / XS */
    void add_raw(const uint8_t *rec, uint32_t len, int32_t mapped_tid);
    ipx_batch view() const;
    ipx_batch_compact view_compact() const;
};

}  // namespace ipx
#endif
