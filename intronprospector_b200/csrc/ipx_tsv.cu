/* ipx_tsv.cu -- intron-call TSV rows tokenised on the device (sm_100a): SURVEY.md 8(f) row 4.
 *
 * Reference: Tsv::next_row / split_line (tsv.hh:27-38,62-74), read_junction (intronProspectorMerge.cc:164-177),
 * string_to_uint / string_to_float (type_ops.hh:18-43).  The bytes of one file (header line removed) go to the device;
 *   k_tsv_count / k_tsv_block_scan / k_tsv_newlines   positions of the line ends (a counting pass, a scan of the block
 *                                                     counts, a writing pass)
 *   k_tsv_parse                                       a thread per line: fields by tabs, integers, the confidence, the
 *                                                     places of the two strings
 * and the row columns come back.  The device parses the PLAIN form only: unsigned decimal integers of at most ten digits
 * that fit 32 bits, decimal fractions without sign or exponent.  Anything else -- a sign, white space, an exponent, too
 * many digits, a wrong number of columns, a carriage return -- raises the row's flag, and the caller parses that line
 * with the reference's own rules (strtoul / strtof and their error texts): the device never guesses.
 *
 * The confidence: strtof rounds the decimal value to the nearest float once.  Here it is m / 10^k computed in double and
 * rounded to float, which is the same float whenever m < 2^53 and k <= 8: m / 10^k differs from the midpoint a / 2^b of
 * two floats (a < 2^25) by at least 1 / (a 10^k) relative, and the double nearest to it can only fall on that midpoint if
 * that is below 2^-53, i.e. for 10^k > 2^28.  More fractional digits raise the flag.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <utility>
#include <vector>

#include "../../include/ipx.h"

namespace {

enum { TSV_BLOCK = 256, TSV_PER_THREAD = 16, TSV_TILE = TSV_BLOCK * TSV_PER_THREAD };
enum { F_CHROM = 0, F_START, F_END, F_STRAND, F_UNIQ, F_MULTI, F_UNSURE, F_LEFT, F_RIGHT, F_CONF, F_SITES, N_FIELDS };

struct TsvOut {
    uint32_t *i0, *i1, *a0, *a1, *counts;
    float *conf;
    char *strand;
    unsigned long long *sites8;
    uint8_t *sites_len;
    uint32_t *chrom_off;
    uint16_t *chrom_len;
    uint32_t *line_off, *line_len;
    uint8_t *flags;
};

__device__ __forceinline__ uint32_t count_newlines(const char *bytes, uint64_t n, uint64_t at) {
    uint32_t c = 0;
    if (at + TSV_PER_THREAD <= n) {
        const uint4 v = *reinterpret_cast<const uint4 *>(bytes + at);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int b = 0; b < 4; ++b) c += ((w[k] >> (8 * b)) & 0xFFu) == (uint32_t)'\n';
    } else {
        for (uint64_t i = at; i < n; ++i) c += bytes[i] == '\n';
    }
    return c;
}

/* newlines per tile of TSV_TILE bytes */
__global__ void __launch_bounds__(TSV_BLOCK) k_tsv_count(const char *bytes, uint64_t n, uint32_t *tile_count) {
    __shared__ uint32_t s_warp[TSV_BLOCK / 32];
    const uint64_t at = ((uint64_t)blockIdx.x * TSV_BLOCK + threadIdx.x) * TSV_PER_THREAD;
    uint32_t c = at < n ? count_newlines(bytes, n, at) : 0u;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, d);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < TSV_BLOCK / 32; ++w) t += s_warp[w];
        tile_count[blockIdx.x] = t;
    }
}

/* exclusive scan of the tile counts by one block; total[0] = newlines, total[1] = lines */
__global__ void __launch_bounds__(1024) k_tsv_block_scan(uint32_t *tile_count, uint32_t n_tiles, const char *bytes, uint64_t n,
                                                         unsigned long long *total) {
    __shared__ uint32_t s_part[1024];
    const uint32_t per = (n_tiles + 1023u) / 1024u;
    const uint32_t lo = threadIdx.x * per, hi = lo + per < n_tiles ? lo + per : n_tiles;
    uint32_t sum = 0;
    for (uint32_t i = lo; i < hi; ++i) sum += tile_count[i];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int i = 0; i < 1024; ++i) {
            const uint32_t v = s_part[i];
            s_part[i] = run;
            run += v;
        }
        total[0] = run;
        total[1] = (unsigned long long)run + ((n > 0 && bytes[n - 1] != '\n') ? 1ull : 0ull); /* a last line without its newline */
    }
    __syncthreads();
    uint32_t run = s_part[threadIdx.x];
    for (uint32_t i = lo; i < hi; ++i) {
        const uint32_t v = tile_count[i];
        tile_count[i] = run;
        run += v;
    }
}

/* nl[j] = position of the j-th newline */
__global__ void __launch_bounds__(TSV_BLOCK) k_tsv_newlines(const char *bytes, uint64_t n, const uint32_t *tile_base, uint32_t *nl) {
    __shared__ uint32_t s_warp[TSV_BLOCK / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t at = ((uint64_t)blockIdx.x * TSV_BLOCK + threadIdx.x) * TSV_PER_THREAD;
    const uint32_t c = at < n ? count_newlines(bytes, n, at) : 0u;
    uint32_t incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += y;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t before = tile_base[blockIdx.x];
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    uint32_t o = before + incl - c;
    if (c) {
        const uint64_t end = at + TSV_PER_THREAD < n ? at + TSV_PER_THREAD : n;
        for (uint64_t i = at; i < end; ++i)
            if (bytes[i] == '\n') nl[o++] = (uint32_t)i;
    }
}

struct TsvCols {
    int32_t n_columns;
    int32_t col[N_FIELDS];
};

__device__ __forceinline__ bool parse_uint(const char *p, uint32_t len, uint32_t &out) {
    if (len == 0u || len > 10u) return false;
    unsigned long long v = 0ull;
    for (uint32_t i = 0; i < len; ++i) {
        const uint32_t d = (uint32_t)(uint8_t)p[i] - (uint32_t)'0';
        if (d > 9u) return false;
        v = v * 10ull + d;
    }
    if (v > 0xFFFFFFFFull) return false; /* strtoul would not stop at 32 bits: the host's business */
    out = (uint32_t)v;
    return true;
}

__device__ __forceinline__ bool parse_conf(const char *p, uint32_t len, float &out) {
    const double p10[9] = {1.0, 10.0, 100.0, 1000.0, 10000.0, 100000.0, 1000000.0, 10000000.0, 100000000.0};
    if (len == 0u) return false;
    unsigned long long m = 0ull;
    uint32_t digits = 0u, sig = 0u, frac = 0u;
    bool dot = false;
    for (uint32_t i = 0; i < len; ++i) {
        const char ch = p[i];
        if (ch == '.') {
            if (dot) return false;
            dot = true;
            continue;
        }
        const uint32_t d = (uint32_t)(uint8_t)ch - (uint32_t)'0';
        if (d > 9u) return false;
        digits++;
        if (m != 0ull || d != 0u) sig++;
        if (sig > 15u) return false;
        m = m * 10ull + d;
        if (dot) frac++;
    }
    if (digits == 0u || frac > 8u) return false;
    out = (float)((double)m / p10[frac]);
    return true;
}

__global__ void __launch_bounds__(256) k_tsv_parse(const char *bytes, uint64_t n, const uint32_t *nl, const unsigned long long *total,
                                                   TsvCols cols, TsvOut o) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t n_nl = total[0], n_lines = total[1];
    if (j >= n_lines) return;
    const uint64_t beg = j == 0 ? 0ull : (uint64_t)nl[j - 1] + 1ull;
    const uint64_t end = j < n_nl ? (uint64_t)nl[j] : n;
    o.line_off[j] = (uint32_t)beg;
    o.line_len[j] = (uint32_t)(end - beg);
    /* the fields this row needs: where they start and how long they are */
    uint32_t f_off[N_FIELDS], f_len[N_FIELDS];
#pragma unroll
    for (int f = 0; f < N_FIELDS; ++f) f_off[f] = f_len[f] = 0xFFFFFFFFu;
    int32_t column = 0;
    uint64_t field_beg = beg;
    for (uint64_t i = beg; i <= end; ++i) {
        if (i == end || bytes[i] == '\t') {
#pragma unroll
            for (int f = 0; f < N_FIELDS; ++f)
                if (cols.col[f] == column) {
                    f_off[f] = (uint32_t)field_beg;
                    f_len[f] = (uint32_t)(i - field_beg);
                }
            column++;
            field_beg = i + 1;
        }
    }
    bool ok = column == cols.n_columns;
#pragma unroll
    for (int f = 0; f < N_FIELDS; ++f) ok = ok && f_len[f] != 0xFFFFFFFFu;
    uint32_t start = 0, stop = 0, left = 0, right = 0, c0 = 0, c1 = 0, c2 = 0;
    float conf = 0.f;
    if (ok) {
        ok = parse_uint(bytes + f_off[F_START], f_len[F_START], start) && parse_uint(bytes + f_off[F_END], f_len[F_END], stop) &&
             parse_uint(bytes + f_off[F_LEFT], f_len[F_LEFT], left) && parse_uint(bytes + f_off[F_RIGHT], f_len[F_RIGHT], right) &&
             parse_uint(bytes + f_off[F_UNIQ], f_len[F_UNIQ], c0) && parse_uint(bytes + f_off[F_MULTI], f_len[F_MULTI], c1) &&
             parse_uint(bytes + f_off[F_UNSURE], f_len[F_UNSURE], c2) && parse_conf(bytes + f_off[F_CONF], f_len[F_CONF], conf);
        ok = ok && f_len[F_SITES] <= 8u && f_len[F_CHROM] <= 0xFFFFu;
    }
    if (ok) {
        /* a carriage return or any other control character in a string field is taken as it is by the reference too, but
         * the plain form has none: leave such rows to the host */
        for (uint32_t i = 0; i < f_len[F_SITES]; ++i) ok = ok && (uint8_t)bytes[f_off[F_SITES] + i] >= 0x20u;
        for (uint32_t i = 0; i < f_len[F_CHROM]; ++i) ok = ok && (uint8_t)bytes[f_off[F_CHROM] + i] >= 0x20u;
    }
    o.flags[j] = ok ? 0 : 1;
    if (!ok) return;
    o.i0[j] = start;
    o.i1[j] = stop;
    o.a0[j] = start - left; /* uint32 arithmetic, as in the reference */
    o.a1[j] = stop + right;
    o.counts[3 * j + 0] = c0;
    o.counts[3 * j + 1] = c1;
    o.counts[3 * j + 2] = c2;
    o.conf[j] = conf;
    o.strand[j] = f_len[F_STRAND] ? bytes[f_off[F_STRAND]] : '\0';
    unsigned long long s8 = 0ull;
    for (uint32_t i = 0; i < f_len[F_SITES]; ++i) s8 |= (unsigned long long)(uint8_t)bytes[f_off[F_SITES] + i] << (8 * i);
    o.sites8[j] = s8;
    o.sites_len[j] = (uint8_t)f_len[F_SITES];
    o.chrom_off[j] = f_off[F_CHROM];
    o.chrom_len[j] = (uint16_t)f_len[F_CHROM];
}

struct Owned {
    ipx_tsv_rows rows;
    void *block; /* one pinned block holds every column */
    size_t block_cap;
};

/* pinned blocks for the columns on their way back, reused from call to call (a tool tokenises dozens of files from a
 * pool of threads; pinning a fresh block costs more than tokenising a file) */
struct BlockPool {
    std::mutex mu;
    std::vector<std::pair<void *, size_t>> idle;
    void *take(size_t bytes, size_t *cap) {
        {
            std::lock_guard<std::mutex> g(mu);
            for (size_t i = 0; i < idle.size(); i++)
                if (idle[i].second >= bytes) {
                    void *p = idle[i].first;
                    *cap = idle[i].second;
                    idle.erase(idle.begin() + i);
                    return p;
                }
        }
        const size_t want = bytes + bytes / 4 + (4u << 20);
        void *p = nullptr;
        if (cudaMallocHost(&p, want) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        *cap = want;
        return p;
    }
    void give(void *p, size_t cap) {
        std::lock_guard<std::mutex> g(mu);
        if (idle.size() < 32) idle.emplace_back(p, cap); else cudaFreeHost(p);
    }
};
BlockPool g_blocks;

size_t up256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace

namespace {
/* what a thread needs to tokenise a file, kept from file to file: a tool runs dozens of files through a pool of threads,
 * and creating streams and (un)mapping device memory per file cost more than the tokenising */
struct TokCtx {
    int device = -1;
    cudaStream_t st = nullptr;
    char *d_bytes = nullptr, *d_out = nullptr;
    uint32_t *d_tiles = nullptr, *d_nl = nullptr;
    unsigned long long *d_total = nullptr;
    size_t bytes_cap = 0, out_cap = 0, tiles_cap = 0, nl_cap = 0;
    void release() {
        if (device < 0) return;
        cudaSetDevice(device);
        cudaFree(d_bytes); cudaFree(d_out); cudaFree(d_tiles); cudaFree(d_nl); cudaFree(d_total);
        if (st) cudaStreamDestroy(st);
        cudaGetLastError(); /* (at process exit the runtime may be gone already) */
        *this = TokCtx();
    }
    ~TokCtx() { release(); }
    template <class T>
    static bool room(T **p, size_t *cap, size_t want) {
        if (*cap >= want) return true;
        cudaFree(*p);
        *p = nullptr;
        *cap = 0;
        const size_t take = want + want / 4 + 4096;
        if (cudaMalloc((void **)p, take) != cudaSuccess) return false;
        *cap = take;
        return true;
    }
    bool open(int dev) {
        if (device == dev) return true;
        release();
        if (cudaSetDevice(dev) != cudaSuccess || cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess ||
            cudaMalloc((void **)&d_total, 16) != cudaSuccess) {
            cudaGetLastError();
            st = nullptr;
            return false;
        }
        device = dev;
        return true;
    }
};
thread_local TokCtx t_ctx;
}  // namespace

extern "C" int ipx_tsv_tokenize(int32_t device, const char *bytes, uint64_t n_bytes, const ipx_tsv_columns *columns, ipx_tsv_rows **out) {
    if (!bytes || !columns || !out) return IPX_ERR_ARG;
    *out = nullptr;
    if (n_bytes >= (1ull << 32) || columns->n_columns < 1) return IPX_ERR_ARG; /* (positions are 32 bits: larger files are the host's) */
    TsvCols cols;
    cols.n_columns = columns->n_columns;
    const int32_t want[N_FIELDS] = {columns->chrom, columns->intron_start, columns->intron_end, columns->strand, columns->uniq_mapped_count,
                                    columns->multi_mapped_count, columns->unsure_mapped_count, columns->max_left_overhang,
                                    columns->max_right_overhang, columns->confidence, columns->splice_sites};
    for (int f = 0; f < N_FIELDS; ++f) {
        if (want[f] < 0 || want[f] >= columns->n_columns) return IPX_ERR_ARG;
        cols.col[f] = want[f];
    }
    TokCtx &c = t_ctx;
    if (!c.open(device)) return IPX_ERR_CUDA;
    cudaSetDevice(device);
    Owned *own = (Owned *)calloc(1, sizeof(Owned));
    if (!own) return IPX_ERR_ARG;
    if (n_bytes == 0) {
        *out = &own->rows;
        return IPX_OK;
    }
    cudaStream_t st = c.st;
    auto fail = [&](int code) {
        cudaStreamSynchronize(st);
        if (own->block) g_blocks.give(own->block, own->block_cap);
        free(own);
        cudaGetLastError();
        return code;
    };
    const uint32_t n_tiles = (uint32_t)((n_bytes + TSV_TILE - 1) / TSV_TILE);
    if (!TokCtx::room(&c.d_bytes, &c.bytes_cap, n_bytes + 16) || !TokCtx::room(&c.d_tiles, &c.tiles_cap, (size_t)n_tiles * 4)) return fail(IPX_ERR_CUDA);
    if (cudaMemcpyAsync(c.d_bytes, bytes, n_bytes, cudaMemcpyHostToDevice, st) != cudaSuccess) return fail(IPX_ERR_CUDA);
    k_tsv_count<<<n_tiles, TSV_BLOCK, 0, st>>>(c.d_bytes, n_bytes, c.d_tiles);
    k_tsv_block_scan<<<1, 1024, 0, st>>>(c.d_tiles, n_tiles, c.d_bytes, n_bytes, c.d_total);
    unsigned long long total[2] = {0, 0};
    if (cudaMemcpyAsync(total, c.d_total, 16, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess)
        return fail(IPX_ERR_CUDA);
    const uint64_t n_nl = total[0], n = total[1];
    /* every column in one device block and one pinned block: one copy back */
    size_t at = 0;
    struct Piece { size_t off, width; } pc[14];
    const size_t widths[14] = {4, 4, 4, 4, 12, 4, 1, 8, 1, 4, 2, 4, 4, 1}; /* i0 i1 a0 a1 counts conf strand sites8 sites_len chrom_off chrom_len line_off line_len flags */
    for (int k = 0; k < 14; ++k) {
        pc[k].off = at;
        pc[k].width = widths[k];
        at += up256((size_t)(n ? n : 1) * widths[k]);
    }
    const size_t out_bytes = at;
    if (!TokCtx::room(&c.d_nl, &c.nl_cap, (size_t)(n_nl ? n_nl : 1) * 4) || !TokCtx::room(&c.d_out, &c.out_cap, out_bytes)) return fail(IPX_ERR_CUDA);
    k_tsv_newlines<<<n_tiles, TSV_BLOCK, 0, st>>>(c.d_bytes, n_bytes, c.d_tiles, c.d_nl);
    own->block = g_blocks.take(out_bytes, &own->block_cap);
    if (!own->block) return fail(IPX_ERR_CUDA);
    char *d_out = c.d_out;
    auto dev = [&](int k) { return d_out + pc[k].off; };
    TsvOut o{(uint32_t *)dev(0), (uint32_t *)dev(1), (uint32_t *)dev(2), (uint32_t *)dev(3), (uint32_t *)dev(4), (float *)dev(5), dev(6),
             (unsigned long long *)dev(7), (uint8_t *)dev(8), (uint32_t *)dev(9), (uint16_t *)dev(10), (uint32_t *)dev(11), (uint32_t *)dev(12),
             (uint8_t *)dev(13)};
    if (n) k_tsv_parse<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(c.d_bytes, n_bytes, c.d_nl, c.d_total, cols, o);
    if (cudaGetLastError() != cudaSuccess) return fail(IPX_ERR_CUDA);
    if (cudaMemcpyAsync(own->block, d_out, out_bytes, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess)
        return fail(IPX_ERR_CUDA);
    char *hb = (char *)own->block;
    ipx_tsv_rows &r = own->rows;
    r.n_rows = n;
    r.intron_start = (const uint32_t *)(hb + pc[0].off);
    r.intron_end = (const uint32_t *)(hb + pc[1].off);
    r.anchor_start = (const uint32_t *)(hb + pc[2].off);
    r.anchor_end = (const uint32_t *)(hb + pc[3].off);
    r.counts = (const uint32_t *)(hb + pc[4].off);
    r.confidence = (const float *)(hb + pc[5].off);
    r.strand = hb + pc[6].off;
    r.sites8 = (const uint64_t *)(hb + pc[7].off);
    r.sites_len = (const uint8_t *)(hb + pc[8].off);
    r.chrom_off = (const uint32_t *)(hb + pc[9].off);
    r.chrom_len = (const uint16_t *)(hb + pc[10].off);
    r.line_off = (const uint32_t *)(hb + pc[11].off);
    r.line_len = (const uint32_t *)(hb + pc[12].off);
    r.flags = (const uint8_t *)(hb + pc[13].off);
    *out = &own->rows;
    return IPX_OK;
}

extern "C" void ipx_tsv_rows_free(ipx_tsv_rows *r) {
    if (!r) return;
    Owned *own = reinterpret_cast<Owned *>(r); /* (rows is the first member) */
    if (own->block) g_blocks.give(own->block, own->block_cap);
    free(own);
}
