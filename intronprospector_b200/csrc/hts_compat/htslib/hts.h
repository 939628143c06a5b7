/* hts.h -- minimal stand-in for the part of htslib's <htslib/hts.h> that the
 * junction-extraction path needs.  htslib itself is not available in the build
 * image (no headers, no library), so this project carries its own SAM/BAM/FASTA
 * reader and writer; the names below follow htslib's public API so that code
 * written against htslib (e.g. the reference's junctions_extractor.cc:573-613)
 * compiles unchanged.  Only the members that are actually used are provided.
 */
#ifndef IPX_COMPAT_HTS_H
#define IPX_COMPAT_HTS_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int64_t hts_pos_t;

enum htsFormatCategory { unknown_category, sequence_data, variant_data, index_file, region_list };
enum htsExactFormat {
    unknown_format, binary_format, text_format,
    sam, bam, bai, cram, crai, vcf, bcf, csi, gzi, tbi, bed, fasta_format, fastq_format
};
enum htsCompression { no_compression, gzip, bgzf, custom };

typedef struct htsFormat {
    enum htsFormatCategory category;
    enum htsExactFormat format;
    struct { short major, minor; } version;
    enum htsCompression compression;
    short compression_level;
    void *specific;
} htsFormat;

struct ipx_hts_impl;  /* private reader/writer state */

typedef struct htsFile {
    uint32_t is_bin : 1, is_write : 1, is_be : 1, is_cram : 1, is_bgzf : 1, dummy : 27;
    int64_t lineno;
    char *fn;
    htsFormat format;
    struct ipx_hts_impl *impl;
} htsFile;

htsFile *hts_open(const char *fn, const char *mode);
htsFile *hts_open_format(const char *fn, const char *mode, const htsFormat *fmt);
int hts_close(htsFile *fp);
const htsFormat *hts_get_format(htsFile *fp);

#ifdef __cplusplus
}
#endif
#endif
