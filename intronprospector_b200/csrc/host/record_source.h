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

/* growable structure-of-arrays batch */
struct ColumnarBatch {
    std::vector<int32_t> tid, pos;
    std::vector<uint16_t> flag;
    std::vector<uint8_t> nh_class, xs;
    std::vector<uint32_t> cigar_off, cigar;
    uint64_t seq_base = 0;
    void clear(uint64_t next_seq_base) {
        tid.clear(); pos.clear(); flag.clear(); nh_class.clear(); xs.clear(); cigar.clear();
        cigar_off.assign(1, 0u);
        seq_base = next_seq_base;
    }
    size_t size() const { return tid.size(); }
    void add(const bam1_t *b, int32_t mapped_tid);
    ipx_batch view() const;
};

}  // namespace ipx
#endif
