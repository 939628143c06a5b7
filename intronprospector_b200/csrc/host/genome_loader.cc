/* genome_loader.cc -- see genome_loader.h.  Sequence text is what faidx_fetch_seq would return for the whole
 * sequence: the bytes between the .fai offset and the end of the sequence with non-printing characters dropped,
 * original letter case kept (upper-casing happens on the device, get_splice_sites junctions_extractor.cc:419). */
#include "genome_loader.h"
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <atomic>
#include <stdexcept>
#include <thread>
#include <vector>
#include "htslib/faidx.h"

namespace ipx {

void load_genome(const std::string &fasta, LoadedGenome &out) {
    faidx_t *fai = fai_load3(fasta.c_str(), NULL, NULL, 0);
    if (fai == NULL) throw std::runtime_error("can't open genome FASTA file or associated index: " + fasta);
    const int n = faidx_nseq(fai);
    out.names.resize(n);
    out.seqs.resize(n);
    out.lengths.resize(n);
    for (int i = 0; i < n; i++) out.names[i] = faidx_iseq(fai, i);
    auto fetch = [&](faidx_t *f, int i) {
        int len = faidx_seq_len(f, out.names[i].c_str());
        int got = 0;
        char *s = len > 0 ? faidx_fetch_seq(f, out.names[i].c_str(), 0, len - 1, &got) : NULL;
        if (s) {
            out.seqs[i].assign(s, (size_t)got);
            free(s);
        }
        /* a short read (truncated file) leaves the tail unfetchable, as it would be for faidx_fetch_seq */
        out.lengths[i] = out.seqs[i].size();
    };
    /* the sequences are independent: a few threads, each with an index handle (and file position) of its own */
    unsigned threads = std::min<unsigned>(std::min<unsigned>(8u, std::thread::hardware_concurrency()), (unsigned)n);
    if (const char *e = getenv("IPX_GENOME_THREADS")) threads = (unsigned)atoi(e);
    if (threads <= 1u) {
        for (int i = 0; i < n; i++) fetch(fai, i);
    } else if (ipx_fai_in_memory(fai)) {
        /* a bgzip-compressed FASTA was inflated whole by fai_load3: the threads share the one handle */
        std::atomic<int> next(0);
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < threads; t++)
            pool.emplace_back([&]() {
                for (int i; (i = next.fetch_add(1)) < n;) fetch(fai, i);
            });
        for (auto &t : pool) t.join();
    } else {
        std::atomic<int> next(0);
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < threads; t++)
            pool.emplace_back([&]() {
                faidx_t *mine = fai_load3(fasta.c_str(), NULL, NULL, 0);
                if (mine == NULL) return; /* the others, or the caller below, take its share */
                for (int i; (i = next.fetch_add(1)) < n;) fetch(mine, i);
                fai_destroy(mine);
            });
        for (auto &t : pool) t.join();
        for (int i; (i = next.fetch_add(1)) < n;) fetch(fai, i);
    }
    fai_destroy(fai);
    out.name_ptrs.resize(n);
    out.seq_ptrs.resize(n);
    for (int i = 0; i < n; i++) {
        out.name_ptrs[i] = out.names[i].c_str();
        out.seq_ptrs[i] = out.seqs[i].data();
    }
}

}  // namespace ipx
