// Scan / radix sort / run reduction (see prims.cuh).  Hand-written for sm_100a; no CUB.
#include "prims.cuh"

namespace dm {
namespace prims {

// ------------------------------------------------------------------------------------ //
// exclusive scan: tile sums -> single-block scan of tile sums -> tile scan + offset
// ------------------------------------------------------------------------------------ //

__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums(const uint32_t* __restrict__ in,
                                                               const int64_t* __restrict__ n_dev,
                                                               uint32_t* __restrict__ tile_sum) {
    const int64_t n = *n_dev;
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    if (base >= n) return;
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        int64_t idx = base + i * SCAN_THREADS + threadIdx.x;
        if (idx < n) s += in[idx];
    }
    __shared__ uint32_t sm[SCAN_THREADS / 32 + 1];
    uint32_t total;
    block_excl_scan<SCAN_THREADS>(s, sm, total);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) scan_tile_offsets(uint32_t* __restrict__ tile_sum,
                                                          const int64_t* __restrict__ n_dev,
                                                          int64_t* __restrict__ total_dev) {
    const int64_t n = *n_dev;
    const int64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    __shared__ uint32_t sm[1024 / 32 + 1];
    uint32_t carry = 0;
    for (int64_t b0 = 0; b0 < tiles; b0 += 1024) {
        int64_t b = b0 + threadIdx.x;
        uint32_t v = b < tiles ? tile_sum[b] : 0;
        uint32_t total;
        uint32_t ex = block_excl_scan<1024>(v, sm, total);
        if (b < tiles) tile_sum[b] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0 && total_dev) *total_dev = (int64_t)carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_apply(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                           const int64_t* __restrict__ n_dev,
                                                           const uint32_t* __restrict__ tile_off) {
    const int64_t n = *n_dev;
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    if (base >= n) return;
    // blocked arrangement: thread t owns items [t*ITEMS, (t+1)*ITEMS)
    uint32_t v[SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        int64_t idx = base + (int64_t)threadIdx.x * SCAN_ITEMS + i;
        v[i] = idx < n ? in[idx] : 0;
        s += v[i];
    }
    __shared__ uint32_t sm[SCAN_THREADS / 32 + 1];
    uint32_t total;
    uint32_t ex = block_excl_scan<SCAN_THREADS>(s, sm, total) + tile_off[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        int64_t idx = base + (int64_t)threadIdx.x * SCAN_ITEMS + i;
        if (idx < n) out[idx] = ex;
        ex += v[i];
    }
}

size_t scan_ws_bytes(int64_t cap) { return align_up((size_t)(ceil_div(cap, SCAN_TILE) + 1) * sizeof(uint32_t), 256); }

int scan_exclusive_u32(const uint32_t* in, uint32_t* out, const int64_t* n_dev, int64_t cap, int64_t* total_dev,
                       void* ws, cudaStream_t s) {
    if (cap <= 0) {
        if (total_dev) DM_CUDA(cudaMemsetAsync(total_dev, 0, sizeof(int64_t), s));
        return DM_OK;
    }
    uint32_t* tile_sum = (uint32_t*)ws;
    const unsigned tiles = (unsigned)ceil_div(cap, SCAN_TILE);
    DM_COUNT_LAUNCH(); scan_tile_sums<<<tiles, SCAN_THREADS, 0, s>>>(in, n_dev, tile_sum);
    DM_COUNT_LAUNCH(); scan_tile_offsets<<<1, 1024, 0, s>>>(tile_sum, n_dev, total_dev);
    DM_COUNT_LAUNCH(); scan_apply<<<tiles, SCAN_THREADS, 0, s>>>(in, out, n_dev, tile_sum);
    DM_LAUNCH_CHECK();
    return DM_OK;
}

// ------------------------------------------------------------------------------------ //
// LSD radix sort, 9-bit digits, 3 kernels per pass (tile histogram, offsets, scatter)
// ------------------------------------------------------------------------------------ //
constexpr int RADIX_BITS = 9;
constexpr int RADIX = 1 << RADIX_BITS;

__device__ __forceinline__ unsigned digit_of(uint64_t key, int id_bits, int shift) {
    uint64_t ck = ((key >> 32) << id_bits) | (key & ((1ull << id_bits) - 1));
    return (unsigned)(ck >> shift) & (unsigned)(RADIX - 1);
}

// hist is digit-major: hist[d * tiles_cap + tile], so that the per-digit scan over tiles is coalesced
__global__ void __launch_bounds__(SORT_THREADS) radix_hist(const uint64_t* __restrict__ keys,
                                                           const int64_t* __restrict__ n_dev, int id_bits, int shift,
                                                           uint32_t* __restrict__ hist, int tiles_cap) {
    const int64_t n = *n_dev;
    const int64_t base = (int64_t)blockIdx.x * SORT_TILE;
    if (base >= n) return;
    __shared__ uint32_t h[RADIX];
    for (int i = threadIdx.x; i < RADIX; i += SORT_THREADS) h[i] = 0;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        int64_t idx = base + i * SORT_THREADS + threadIdx.x;
        if (idx < n) atomicAdd(&h[digit_of(keys[idx], id_bits, shift)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < RADIX; i += SORT_THREADS) hist[(size_t)i * tiles_cap + blockIdx.x] = h[i];
}

// One warp per digit: turns the per-tile counts of that digit into an exclusive prefix over tiles
// (in place) and writes the digit's total to totals[d].  The digit bases (exclusive scan of the
// totals) are computed by every scatter block itself.
__global__ void __launch_bounds__(1024) radix_offsets(uint32_t* __restrict__ hist, const int64_t* __restrict__ n_dev,
                                                      uint32_t* __restrict__ totals, int tiles_cap) {
    const int64_t n = *n_dev;
    const int tiles = (int)((n + SORT_TILE - 1) / SORT_TILE);
    const int lane = threadIdx.x & 31;
    const int d = blockIdx.x * 32 + (threadIdx.x >> 5);
    uint32_t* row = hist + (size_t)d * tiles_cap;
    uint32_t carry = 0;
    for (int t0 = 0; t0 < tiles; t0 += 32) {
        const int t = t0 + lane;
        const uint32_t c = t < tiles ? row[t] : 0;
        uint32_t inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        if (t < tiles) row[t] = carry + inc - c;
        carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) totals[d] = carry;
}

__global__ void __launch_bounds__(SORT_THREADS) radix_scatter(const uint64_t* __restrict__ kin,
                                                              const uint32_t* __restrict__ vin,
                                                              uint64_t* __restrict__ kout, uint32_t* __restrict__ vout,
                                                              const int64_t* __restrict__ n_dev, int id_bits, int shift,
                                                              const uint32_t* __restrict__ hist,
                                                              const uint32_t* __restrict__ totals, int tiles_cap) {
    const int64_t n = *n_dev;
    const int64_t tile_base = (int64_t)blockIdx.x * SORT_TILE;
    if (tile_base >= n) return;
    constexpr int NW = SORT_THREADS / 32;
    constexpr int PER = RADIX / SORT_THREADS;             // digits per thread in the block-wide steps
    __shared__ uint32_t wcnt[NW][RADIX];
    __shared__ uint32_t goff[RADIX];
    __shared__ uint32_t sm_scan[SORT_THREADS / 32 + 1];
    for (int i = threadIdx.x; i < NW * RADIX; i += SORT_THREADS) (&wcnt[0][0])[i] = 0;
    {   // digit base = exclusive scan of the digit totals (thread t owns digits [t*PER, (t+1)*PER))
        uint32_t t[PER], sum = 0;
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            t[k] = totals[threadIdx.x * PER + k];
            sum += t[k];
        }
        uint32_t tot;
        uint32_t base = block_excl_scan<SORT_THREADS>(sum, sm_scan, tot);
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int d = threadIdx.x * PER + k;
            goff[d] = base + hist[(size_t)d * tiles_cap + blockIdx.x];
            base += t[k];
        }
    }
    __syncthreads();
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t wbase = tile_base + (int64_t)warp * (32 * SORT_ITEMS);
    uint64_t key[SORT_ITEMS];
    uint32_t rank[SORT_ITEMS];
    unsigned dig[SORT_ITEMS];
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        const int64_t idx = wbase + i * 32 + lane;
        key[i] = idx < n ? kin[idx] : 0;
    }
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        const int64_t idx = wbase + i * 32 + lane;
        const bool valid = idx < n;
        const unsigned vmask = __ballot_sync(0xffffffffu, valid);
        dig[i] = 0;
        rank[i] = 0;
        if (valid) {
            dig[i] = digit_of(key[i], id_bits, shift);
            const unsigned peers = __match_any_sync(vmask, dig[i]);
            const int leader = __ffs(peers) - 1;
            uint32_t pre = 0;
            if ((int)lane == leader) {
                pre = wcnt[warp][dig[i]];
                wcnt[warp][dig[i]] = pre + __popc(peers);
            }
            pre = __shfl_sync(peers, pre, leader);
            rank[i] = pre + __popc(peers & lanemask_lt());
        }
        __syncwarp();
    }
    __syncthreads();
    for (int d = threadIdx.x; d < RADIX; d += SORT_THREADS) {   // exclusive prefix over warps, per digit
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            uint32_t c = wcnt[w][d];
            wcnt[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        const int64_t idx = wbase + i * 32 + lane;
        if (idx < n) {
            const uint32_t pos = goff[dig[i]] + wcnt[warp][dig[i]] + rank[i];
            kout[pos] = key[i];
            if (vin) vout[pos] = vin[idx];
        }
    }
}

__global__ void copy_pairs(const uint64_t* __restrict__ kin, const uint32_t* __restrict__ vin, uint64_t* __restrict__ kout,
                           uint32_t* __restrict__ vout, const int64_t* __restrict__ n_dev) {
    const int64_t n = *n_dev;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        kout[i] = kin[i];
        if (vin) vout[i] = vin[i];
    }
}

size_t sort_ws_bytes(int64_t cap) {
    if (cap < 1) cap = 1;
    size_t b = 0;
    b += align_up((size_t)cap * sizeof(uint64_t), 256);
    b += align_up((size_t)cap * sizeof(uint32_t), 256);
    b += align_up((size_t)ceil_div(cap, SORT_TILE) * RADIX * sizeof(uint32_t), 256);
    b += align_up(RADIX * sizeof(uint32_t), 256);
    return b;
}

int sort_pairs(uint64_t* keys, uint32_t* vals, const int64_t* n_dev, int64_t cap, int id_bits, int key_bits, void* ws,
               cudaStream_t s) {
    if (cap <= 0) return DM_OK;
    if (id_bits < 1 || id_bits > 32 || key_bits < 1 || key_bits > 2 * id_bits) return DM_ERR_BAD_ARG;
    Carver c(ws);
    uint64_t* k2 = c.take<uint64_t>(cap);
    uint32_t* v2 = c.take<uint32_t>(cap);
    const unsigned tiles = (unsigned)ceil_div(cap, SORT_TILE);
    uint32_t* hist = c.take<uint32_t>((size_t)tiles * RADIX);
    uint32_t* totals = c.take<uint32_t>(RADIX);
    const int passes = (key_bits + RADIX_BITS - 1) / RADIX_BITS;
    uint64_t *ka = keys, *kb = k2;
    uint32_t *va = vals, *vb = vals ? v2 : nullptr;
    for (int p = 0; p < passes; ++p) {
        DM_COUNT_LAUNCH(); radix_hist<<<tiles, SORT_THREADS, 0, s>>>(ka, n_dev, id_bits, RADIX_BITS * p, hist, (int)tiles);
        DM_COUNT_LAUNCH(); radix_offsets<<<RADIX / 32, 1024, 0, s>>>(hist, n_dev, totals, (int)tiles);
        DM_COUNT_LAUNCH(); radix_scatter<<<tiles, SORT_THREADS, 0, s>>>(ka, va, kb, vb, n_dev, id_bits, RADIX_BITS * p, hist, totals,
                                                    (int)tiles);
        uint64_t* tk = ka; ka = kb; kb = tk;
        uint32_t* tv = va; va = vb; vb = tv;
    }
    if (ka != keys) {
        const unsigned g = (unsigned)imin64(ceil_div(cap, 256), 148 * 8);
        DM_COUNT_LAUNCH(); copy_pairs<<<g, 256, 0, s>>>(ka, va, keys, vals, n_dev);
    }
    DM_LAUNCH_CHECK();
    return DM_OK;
}

// ------------------------------------------------------------------------------------ //
// run reduction on sorted keys
// ------------------------------------------------------------------------------------ //

__global__ void unique_flags(const uint64_t* __restrict__ keys, const int64_t* __restrict__ n_dev, uint64_t sentinel,
                             uint32_t* __restrict__ flags) {
    const int64_t n = *n_dev;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t k = keys[i];
        flags[i] = (k != sentinel && (i == 0 || keys[i - 1] != k)) ? 1u : 0u;
    }
}

__global__ void unique_scatter(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ perm,
                               const uint32_t* __restrict__ lens_in, const float* __restrict__ scores_in,
                               const int64_t* __restrict__ n_dev, uint64_t sentinel, const uint32_t* __restrict__ flags,
                               const uint32_t* __restrict__ excl, uint64_t* __restrict__ out_keys,
                               uint32_t* __restrict__ out_lens, float* __restrict__ out_scores) {
    const int64_t n = *n_dev;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t k = keys[i];
        if (k == sentinel) continue;
        const uint32_t f = flags[i];
        const uint32_t run = excl[i] + f - 1;
        const uint32_t src = perm ? perm[i] : (uint32_t)i;
        if (f) {
            out_keys[run] = k;
            if (scores_in) out_scores[run] = scores_in[src];
        }
        atomicAdd(&out_lens[run], lens_in[src]);
    }
}

size_t unique_ws_bytes(int64_t cap) {
    if (cap < 1) cap = 1;
    return 2 * align_up((size_t)cap * sizeof(uint32_t), 256) + scan_ws_bytes(cap);
}

int unique_reduce(const uint64_t* keys, const uint32_t* perm, const uint32_t* lens_in, const float* scores_in,
                  const int64_t* n_dev, int64_t cap, uint64_t sentinel, uint64_t* out_keys, uint32_t* out_lens,
                  float* out_scores, int64_t* n_out_dev, void* ws, cudaStream_t s) {
    if (cap <= 0) {
        DM_CUDA(cudaMemsetAsync(n_out_dev, 0, sizeof(int64_t), s));
        return DM_OK;
    }
    Carver c(ws);
    uint32_t* flags = c.take<uint32_t>(cap);
    uint32_t* excl = c.take<uint32_t>(cap);
    void* scan_ws = c.take<char>(scan_ws_bytes(cap));
    const unsigned g = (unsigned)imin64(ceil_div(cap, 256), 148 * 8);
    DM_CUDA(cudaMemsetAsync(out_lens, 0, (size_t)cap * sizeof(uint32_t), s));
    DM_COUNT_LAUNCH(); unique_flags<<<g, 256, 0, s>>>(keys, n_dev, sentinel, flags);
    DM_TRY(scan_exclusive_u32(flags, excl, n_dev, cap, n_out_dev, scan_ws, s));
    DM_COUNT_LAUNCH(); unique_scatter<<<g, 256, 0, s>>>(keys, perm, lens_in, scores_in, n_dev, sentinel, flags, excl, out_keys, out_lens,
                                     out_scores);
    DM_LAUNCH_CHECK();
    return DM_OK;
}

}  // namespace prims
}  // namespace dm
