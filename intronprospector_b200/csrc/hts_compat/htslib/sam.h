/* sam.h -- minimal stand-in for <htslib/sam.h>; see hts.h in this directory.
 * Record layout (bam1_t.data = qname | cigar | seq | qual | aux, qname NUL padded
 * to a 4-byte boundary) and every name below follow htslib's public API; only
 * what the junction-extraction path and its pass-through writer use is provided
 * (reference call sites: junctions_extractor.cc:49-111,287-290,335-336,556-633).
 */
#ifndef IPX_COMPAT_SAM_H
#define IPX_COMPAT_SAM_H
#include <stdint.h>
#include "hts.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct sam_hdr_t {
    int32_t n_targets, ignore_sam_err;
    size_t l_text;
    uint32_t *target_len;
    const int8_t *cigar_tab;
    char **target_name;
    char *text;
    void *sdict;
    void *hrecs;
    uint32_t ref_count;
} sam_hdr_t;
typedef sam_hdr_t bam_hdr_t;

#define BAM_CMATCH      0
#define BAM_CINS        1
#define BAM_CDEL        2
#define BAM_CREF_SKIP   3
#define BAM_CSOFT_CLIP  4
#define BAM_CHARD_CLIP  5
#define BAM_CPAD        6
#define BAM_CEQUAL      7
#define BAM_CDIFF       8
#define BAM_CBACK       9

#define BAM_CIGAR_STR   "MIDNSHP=XB"
#define BAM_CIGAR_SHIFT 4
#define BAM_CIGAR_MASK  0xf
#define BAM_CIGAR_TYPE  0x3C1A7

#define bam_cigar_op(c) ((c)&BAM_CIGAR_MASK)
#define bam_cigar_oplen(c) ((c)>>BAM_CIGAR_SHIFT)
#define bam_cigar_opchr(c) (BAM_CIGAR_STR "??????" [bam_cigar_op(c)])
#define bam_cigar_gen(l, o) ((l)<<BAM_CIGAR_SHIFT|(o))
#define bam_cigar_type(o) (BAM_CIGAR_TYPE>>((o)<<1)&3)

#define BAM_FPAIRED        1
#define BAM_FPROPER_PAIR   2
#define BAM_FUNMAP         4
#define BAM_FMUNMAP        8
#define BAM_FREVERSE      16
#define BAM_FMREVERSE     32
#define BAM_FREAD1        64
#define BAM_FREAD2       128
#define BAM_FSECONDARY   256
#define BAM_FQCFAIL      512
#define BAM_FDUP        1024
#define BAM_FSUPPLEMENTARY 2048

typedef struct bam1_core_t {
    hts_pos_t pos;
    int32_t tid;
    uint16_t bin;
    uint8_t qual;
    uint8_t l_extranul;
    uint16_t flag;
    uint16_t l_qname;
    uint32_t n_cigar;
    int32_t l_qseq;
    int32_t mtid;
    hts_pos_t mpos;
    hts_pos_t isize;
} bam1_core_t;

typedef struct bam1_t {
    bam1_core_t core;
    uint64_t id;
    uint8_t *data;
    int l_data;
    uint32_t m_data;
    uint32_t mempolicy : 2, : 30;
} bam1_t;

#define bam_is_rev(b) (((b)->core.flag&BAM_FREVERSE) != 0)
#define bam_is_mrev(b) (((b)->core.flag&BAM_FMREVERSE) != 0)
#define bam_get_qname(b) ((char*)(b)->data)
#define bam_get_cigar(b) ((uint32_t*)((b)->data + (b)->core.l_qname))
#define bam_get_seq(b)   ((b)->data + ((b)->core.n_cigar<<2) + (b)->core.l_qname)
#define bam_get_qual(b)  ((b)->data + ((b)->core.n_cigar<<2) + (b)->core.l_qname + (((b)->core.l_qseq + 1)>>1))
#define bam_get_aux(b)   ((b)->data + ((b)->core.n_cigar<<2) + (b)->core.l_qname + (((b)->core.l_qseq + 1)>>1) + (b)->core.l_qseq)
#define bam_get_l_aux(b) ((b)->l_data - ((b)->core.n_cigar<<2) - (b)->core.l_qname - (b)->core.l_qseq - (((b)->core.l_qseq + 1)>>1))
#define bam_seqi(s, i) ((s)[(i)>>1] >> ((~(i)&1)<<2) & 0xf)

typedef htsFile samFile;
#define sam_open(fn, mode) (hts_open((fn), (mode)))
#define sam_open_format(fn, mode, fmt) (hts_open_format((fn), (mode), (fmt)))
#define sam_close(fp) hts_close(fp)

sam_hdr_t *sam_hdr_init(void);
sam_hdr_t *sam_hdr_read(samFile *fp);
int sam_hdr_write(samFile *fp, const sam_hdr_t *h);
void sam_hdr_destroy(sam_hdr_t *h);
#define bam_hdr_destroy(h) sam_hdr_destroy(h)
int sam_hdr_name2tid(sam_hdr_t *h, const char *ref);

bam1_t *bam_init1(void);
void bam_destroy1(bam1_t *b);
bam1_t *bam_copy1(bam1_t *bdst, const bam1_t *bsrc);
hts_pos_t bam_cigar2rlen(int n_cigar, const uint32_t *cigar);
hts_pos_t bam_endpos(const bam1_t *b);

int sam_read1(samFile *fp, sam_hdr_t *h, bam1_t *b);
int sam_write1(samFile *fp, const sam_hdr_t *h, const bam1_t *b);

uint8_t *bam_aux_get(const bam1_t *b, const char tag[2]);
int64_t bam_aux2i(const uint8_t *s);
double bam_aux2f(const uint8_t *s);
char bam_aux2A(const uint8_t *s);
char *bam_aux2Z(const uint8_t *s);
int bam_aux_append(bam1_t *b, const char tag[2], char type, int len, const uint8_t *data);
int bam_aux_del(bam1_t *b, uint8_t *s);

/* ---- not htslib: an extension of this subset for the drop-in CLI's record packer ----
 * The next BAM record where it lies in the reader's buffer.  ipx_bam_peek returns 1 and sets *rec to the record's
 * fixed 32-byte part (the bytes after block_size) and *len to block_size when the whole record is buffered and
 * well-formed; 0 when the caller has to use sam_read1 instead (not a BAM file, end of file, a truncated or malformed
 * record, a CIGAR kept in the CG tag).  Nothing is consumed; the pointer is valid until the next read on fp.
 * ipx_bam_consume steps over the record the last ipx_bam_peek returned. */
int ipx_bam_peek(samFile *fp, const uint8_t **rec, uint32_t *len);
/* uncompressed bytes of a BAM file before its first record (after sam_hdr_read) */
uint64_t ipx_bam_header_bytes(samFile *fp);
void ipx_bam_consume(samFile *fp);

#ifdef __cplusplus
}
#endif
#endif
