/* genome_loader.h -- FASTA + .fai -> whole sequences in memory (host side of ipx_set_genome). */
#ifndef IPX_GENOME_LOADER_H
#define IPX_GENOME_LOADER_H
#include <stdint.h>
#include <string>
#include <vector>

namespace ipx {
struct LoadedGenome {
    std::vector<std::string> names;
    std::vector<std::string> seqs;
    std::vector<const char *> name_ptrs, seq_ptrs;
    std::vector<uint64_t> lengths;
};
/* throws std::runtime_error with the reference's message (genome.cc:47-49) when the FASTA or its index cannot be opened */
void load_genome(const std::string &fasta, LoadedGenome &out);
}  // namespace ipx
#endif
