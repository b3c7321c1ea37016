// Device-wide primitives used by the graph stages: exclusive scan, stable LSD radix
// sort of packed edge keys, and sorted-run reduction (unique + sum).  All take their
// element count from device memory so that stages chain without host round trips.
#pragma once
#include "common.cuh"

namespace dm {
namespace prims {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 8;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;

size_t scan_ws_bytes(int64_t cap);
int scan_exclusive_u32(const uint32_t* in, uint32_t* out, const int64_t* n_dev, int64_t cap,
                       int64_t* total_dev, void* ws, cudaStream_t s);

size_t sort_ws_bytes(int64_t cap);
// Stable sort of (keys, vals) in place by the low key_bits bits of the compacted key
// ((hi32 << id_bits) | lo32); both 32-bit halves of every key must be < 2^id_bits.
// key_bits = 2*id_bits sorts by (hi, lo); key_bits = id_bits sorts by lo only.  vals may be null.
// allow_fused: the whole sort may run as ONE cooperative launch (persistent blocks + grid barriers); callers that
// sort beside a kernel filling every SM (the pooling chain beside the raster pass) pass false.
// n_ids > 0 (with key_bits = id_bits or 2 id_bits): the primary half of every key is an id < n_ids, and the sort may
// run as bucket + rank in one cooperative launch (bucket_sort_fused in prims.cu).  csr_offsets / csr_ids (need
// n_ids > 0 and a cooperative launch; DM_ERR_UNSUPPORTED otherwise, the caller then takes its own path): also
// offsets[id] = first sorted position of primary id `id`, for id < n_ids, and the sorted values as int32.
int sort_pairs(uint64_t* keys, uint32_t* vals, const int64_t* n_dev, int64_t cap, int id_bits, int key_bits,
               void* ws, cudaStream_t s, bool allow_fused = true, int64_t n_ids = 0, int64_t* csr_offsets = nullptr,
               int32_t* csr_ids = nullptr);
bool sort_fused_available();
int sort_fused_mode();   // DM_SORT_FUSED: 0 off, 1 edge sorts, 2 (default) also the member sort of dm_merge_apply

size_t unique_ws_bytes(int64_t cap);
// keys sorted.  For every run of equal keys (runs of `sentinel` are dropped):
//   out_keys[r] = key, out_lens[r] = sum lens_in[perm ? perm[i] : i],
//   out_scores[r] = scores_in[perm ? perm[i] : i] of the run head (when scores_in != null).
int unique_reduce(const uint64_t* keys, const uint32_t* perm, const uint32_t* lens_in, const float* scores_in,
                  const int64_t* n_dev, int64_t cap, uint64_t sentinel, uint64_t* out_keys, uint32_t* out_lens,
                  float* out_scores, int64_t* n_out_dev, void* ws, cudaStream_t s);

// sort_pairs + unique_reduce (+ an optional copy of the reduced list over back_keys / back_lens / back_scores and
// its length to back_n) -- one launch when the fused kernel is available.  n = min(*n_dev, cap) pairs are read.
// gather_lens == null: the (sorted) values are the lengths.  Otherwise the values are a permutation and lengths /
// scores are gathered through it (gather_lens[vals[i]], gather_scores[vals[i]]).
int sort_unique(uint64_t* keys, uint32_t* vals, const int64_t* n_dev, int64_t cap, int id_bits, int key_bits, void* sort_ws,
                const uint32_t* gather_lens, const float* gather_scores, uint64_t sentinel, uint64_t* out_keys,
                uint32_t* out_lens, float* out_scores, int64_t* n_out_dev, void* unique_ws, uint64_t* back_keys,
                uint32_t* back_lens, float* back_scores, int64_t* back_n, cudaStream_t s, int64_t n_ids = 0);
// n_ids > 0: both halves of every key that is not the sentinel are ids < n_ids and only the reduced list is wanted
// (keys / vals keep their contents): the run reduction may then go through a hash table + per-id buckets instead of
// sorting the raw entries (edge_unique_hashed in prims.cu; DM_EDGE_HASH=0 switches it off).

// ---- device helpers -------------------------------------------------------------------

// Block-wide exclusive scan of one uint32 per thread.  smem needs NT/32 + 1 words.
template <int NT>
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* smem, uint32_t& total) {
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (unsigned)o) inc += t;
    }
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < NT / 32 ? smem[lane] : 0;
        uint32_t winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= (unsigned)o) winc += t;
        }
        if (lane < NT / 32) smem[lane] = winc - w;
        if (lane == 31) smem[NT / 32] = winc;
    }
    __syncthreads();
    uint32_t r = smem[warp] + inc - v;
    total = smem[NT / 32];
    __syncthreads();
    return r;
}

}  // namespace prims
}  // namespace dm
