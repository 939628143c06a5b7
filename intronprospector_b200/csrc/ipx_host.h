/* ipx_host.h -- host finalizer entry points (ipx_host.cc) used by the C ABI. */
#ifndef IPX_HOST_H
#define IPX_HOST_H
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/ipx.h"

namespace ipx {
void compute_confidences(uint64_t n, const uint64_t *read_counts, const uint64_t *hist_off, const uint32_t *hist_cnt,
                         float *out);
void sort_rows(const ipx_table *t, const char *const *names, bool by_anchor, uint64_t begin, uint64_t end,
               std::vector<uint64_t> &idx);
int format_rows(const ipx_table *t, const char *const *names, int format, float min_confidence, uint64_t begin,
                uint64_t end, std::string &out);
const char *tsv_header();
}  // namespace ipx
#endif
