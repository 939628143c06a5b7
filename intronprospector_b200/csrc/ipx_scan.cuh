/* ipx_scan.cuh -- single-pass exclusive prefix sum (decoupled look-back) used wherever the path needs offsets:
 * CIGAR offsets of a compact batch (ipx_extract.cu), histogram offsets of the finished table (ipx_finalize.cu).
 *
 * A block takes tiles of SCAN_TILE items in ticket order, publishes the tile's sum (flag 1), finds its exclusive
 * prefix by looking back over the predecessors' published words 32 at a time until it meets one that carries an
 * inclusive prefix (flag 2), then publishes its own.  One word per tile: value << 2 | flag, so value and flag become
 * visible together.  The caller zeroes `state` (scan_state_words(count) words) before every launch.
 */
#ifndef IPX_SCAN_CUH
#define IPX_SCAN_CUH
#include <cuda_runtime.h>
#include <stdint.h>

namespace ipx {

static const int SCAN_THREADS = 256;
static const int SCAN_ITEMS = 16;
static const int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

inline uint64_t scan_tiles(uint64_t count) { return (count + SCAN_TILE - 1) / SCAN_TILE; }
/* one word per tile plus the ticket counter */
inline uint64_t scan_state_words(uint64_t count) { return scan_tiles(count) + 1; }

template <class In, class OutT>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_exclusive(In in, OutT *out, uint64_t count, unsigned long long *state) {
    __shared__ unsigned long long s_warp[SCAN_THREADS / 32];
    __shared__ unsigned long long s_prefix;
    __shared__ unsigned long long s_tile;
    const uint64_t n_tiles = (count + SCAN_TILE - 1) / SCAN_TILE;
    unsigned long long *ticket = state + n_tiles;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1ull);
    __syncthreads();
    const uint64_t tile = s_tile;
    if (tile >= n_tiles) return;
    const uint64_t i0 = tile * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
    unsigned long long v[SCAN_ITEMS];
    unsigned long long mine = 0ull;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = i0 + k < count ? (unsigned long long)in(i0 + k) : 0ull;
        mine += v[k];
    }
    unsigned long long incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += y;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned long long warp_base = 0ull, block_sum = 0ull;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; ++w) {
        const unsigned long long s = s_warp[w];
        if (w < warp) warp_base += s;
        block_sum += s;
    }
    if (warp == 0) {
        unsigned long long prefix = 0ull;
        if (tile == 0) {
            if (lane == 0) atomicExch(&state[0], (block_sum << 2) | 2ull);
        } else {
            if (lane == 0) atomicExch(&state[tile], (block_sum << 2) | 1ull);
            /* look back over the predecessors, 32 at a time, nearest first */
            long long hi = (long long)tile - 1;
            for (;;) {
                const long long idx = hi - lane;
                unsigned long long w = 2ull; /* lanes before tile 0: an empty inclusive prefix */
                if (idx >= 0) {
                    do {
                        w = *((volatile unsigned long long *)&state[idx]);
                    } while ((w & 3ull) == 0ull);
                }
                const unsigned done = __ballot_sync(0xFFFFFFFFu, (w & 3ull) == 2ull);
                const int stop = done ? __ffs(done) - 1 : 31; /* nearest predecessor that knows its inclusive prefix */
                unsigned long long part = lane <= stop ? (w >> 2) : 0ull;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, d);
                prefix += part;
                if (done) break;
                hi -= 32;
            }
            if (lane == 0) atomicExch(&state[tile], ((prefix + block_sum) << 2) | 2ull);
        }
        if (lane == 0) s_prefix = prefix;
    }
    __syncthreads();
    unsigned long long run = s_prefix + warp_base + (incl - mine);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (i0 + k < count) out[i0 + k] = (OutT)run;
        run += v[k];
    }
}

/* enqueue: zero the tile states, then scan `count` items */
template <class In, class OutT>
inline cudaError_t scan_exclusive(In in, OutT *out, uint64_t count, unsigned long long *state, cudaStream_t st) {
    if (count == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(state, 0, scan_state_words(count) * sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    k_scan_exclusive<In, OutT><<<(unsigned)scan_tiles(count), SCAN_THREADS, 0, st>>>(in, out, count, state);
    return cudaGetLastError();
}

}  // namespace ipx
#endif
