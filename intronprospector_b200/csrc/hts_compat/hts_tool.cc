/* hts_tool.cc -- tiny stand-in for the samtools sub-commands the reference's
 * test suite shells out to (tests/Makefile.in:88,264-267,273-274), built on
 * hts_compat.cc:
 *   hts_tool view-bam IN OUT      SAM/BAM -> BAM (BGZF)
 *   hts_tool view-rawbam IN OUT   SAM/BAM -> uncompressed BAM stream without BGZF framing
 *   hts_tool view-sam IN OUT      SAM/BAM -> SAM text with header
 *   hts_tool faidx FASTA          write FASTA.fai
 *   hts_tool count IN             read every record and print the count (decode-only timing)
 */
#include "htslib/sam.h"
#include "htslib/faidx.h"
#include <cstdio>
#include <cstring>

static int convert(const char *in, const char *out, const char *mode) {
    samFile *fi = sam_open(in, "r");
    if (!fi) { fprintf(stderr, "can't open %s\n", in); return 1; }
    sam_hdr_t *h = sam_hdr_read(fi);
    if (!h) { fprintf(stderr, "can't read header of %s\n", in); return 1; }
    samFile *fo = sam_open(out, mode);
    if (!fo) { fprintf(stderr, "can't open %s\n", out); return 1; }
    if (sam_hdr_write(fo, h) < 0) return 1;
    bam1_t *b = bam_init1();
    int rc;
    while ((rc = sam_read1(fi, h, b)) >= 0)
        if (sam_write1(fo, h, b) < 0) { fprintf(stderr, "write error\n"); return 1; }
    if (rc < -1) { fprintf(stderr, "read error %d in %s\n", rc, in); return 1; }
    bam_destroy1(b);
    sam_hdr_destroy(h);
    sam_close(fi);
    return sam_close(fo) < 0 ? 1 : 0;
}

int main(int argc, char **argv) {
    if (argc >= 4 && strcmp(argv[1], "view-bam") == 0) return convert(argv[2], argv[3], "wb");
    if (argc >= 4 && strcmp(argv[1], "view-rawbam") == 0) return convert(argv[2], argv[3], "wbR");
    if (argc >= 4 && strcmp(argv[1], "view-sam") == 0) return convert(argv[2], argv[3], "w");
    if (argc >= 3 && strcmp(argv[1], "faidx") == 0) return fai_build(argv[2]) < 0 ? 1 : 0;
    if (argc >= 3 && strcmp(argv[1], "count") == 0) {
        samFile *fi = sam_open(argv[2], "r");
        if (!fi) return 1;
        sam_hdr_t *h = sam_hdr_read(fi);
        bam1_t *b = bam_init1();
        long n = 0;
        long ncig = 0;
        int rc;
        while ((rc = sam_read1(fi, h, b)) >= 0) { n++; ncig += b->core.n_cigar; }
        printf("%ld records %ld cigar ops\n", n, ncig);
        return rc < -1;
    }
    fprintf(stderr, "usage: hts_tool view-bam|view-rawbam|view-sam IN OUT | faidx FASTA | count IN\n");
    return 2;
}
