// R1: region adjacency graph from a label raster, fused with per-region band pooling.
//
// One persistent CTA per SM walks DOWN a 256-pixel-wide column of the raster, tile by
// tile (TH rows).  A producer warp streams label tiles (with a 1-row / 4-column halo) and
// image tiles into a multi-stage shared-memory ring with TMA (cp.async.bulk.tensor,
// mbarrier complete_tx); compute warps own a 128-pixel strip x BR-row band of each tile:
// every lane owns 4 consecutive pixels (one 128-bit LDS of labels, one of image bytes)
// and walks down the band keeping
//   * a 2-entry register cache of per-label accumulators (area, border sides, C band sums
//     and sums of squares computed 4 pixels at a time with PRMT + DP4A), and
//   * a 1-entry run cache of the current (min,max) edge key with its pair count.
// Cache evictions go to CTA-wide shared-memory hash tables (labels -> accumulators, edge
// key -> count) that persist while the CTA walks down its column, so a region or an edge
// costs a handful of global atomics / one appended entry per CTA instead of per pixel.
// The appended (key,count) entries are then radix sorted and run-reduced (prims.cu).
//
// HBM traffic: labels 4 B/px + image C B/px, read once (halo re-reads hit L2).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "prims.cuh"

namespace dm {
namespace rag {

constexpr int STRIP_W = 128;            // pixels per warp row: 32 lanes x 4
constexpr int LAB_PITCH = STRIP_W + 4;  // + halo columns (TMA inner box must be a multiple of 16 B)
constexpr int RSLOTS = 512;             // region table slots (power of two)
constexpr int ESLOTS = 1024;            // edge table slots (power of two)
constexpr int MAX_PROBE = 24;
constexpr int FLUSH_ROWS = 256;         // forced table flush period: 256 px * 256 rows * 255^2 < 2^32
constexpr int EMPTY_LABEL = -1;
constexpr unsigned long long EMPTY_KEY = ~0ull;

constexpr int align128(int x) { return (x + 127) / 128 * 128; }

template <int C_, int TH_, int STRIPS_, int BANDS_, int STAGES_>
struct Cfg {
    static constexpr int C = C_, TH = TH_, STRIPS = STRIPS_, BANDS = BANDS_, STAGES = STAGES_;
    static constexpr int CW = C_ > 0 ? C_ : 1;            // words of image bytes per lane-row
    static constexpr int BR = TH / BANDS;
    static constexpr int NCW = STRIPS * BANDS;             // compute warps
    static constexpr int NCT = NCW * 32;                   // compute threads
    static constexpr int THREADS = NCT + 32;               // + producer warp
    static constexpr int TILE_W = STRIPS * STRIP_W;
    static constexpr int LAB_BOX = align128((TH + 1) * LAB_PITCH * 4);
    static constexpr int IMG_ROW_WORDS = STRIP_W * C / 4;  // 32*C
    static constexpr int IMG_BOX = align128(TH * IMG_ROW_WORDS * 4);
    static constexpr int STAGE_BYTES = STRIPS * (LAB_BOX + IMG_BOX);
    static constexpr int TABLE_WORDS = RSLOTS * (3 + 2 * C) + ESLOTS * 3;
    static constexpr int SMEM_BYTES = 128 + STAGES * STAGE_BYTES + TABLE_WORDS * 4 + 256;
    static constexpr int FLUSH_TILES = FLUSH_ROWS / TH > 0 ? FLUSH_ROWS / TH : 1;
};

struct Params {
    const int32_t* labels;
    int64_t ld;
    const uint8_t* image;
    int64_t image_pitch;
    int rows_own, rows_avail, W;
    int n_regions;
    int top_border, bottom_border;
    unsigned long long* area;
    unsigned long long* border;
    unsigned long long* bsum;
    unsigned long long* bsq;
    unsigned long long* raw_keys;
    uint32_t* raw_cnt;
    long long capacity;
    unsigned long long* counters;   // [1] raw entries, [2] overflow, [3] bad label / internal error
    int tiles_x, tiles_y, tiles_per_cta;
};

// ------------------------------------------------------------------------------------ //
// mbarrier / TMA wrappers
// ------------------------------------------------------------------------------------ //
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a TMA that never lands must become an error, not a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity, unsigned long long* counters) {
    unsigned spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 24)) {
            atomicExch(&counters[3], 2ull);
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst)),
        "l"((uint64_t)map), "r"(x), "r"(y), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void compute_bar(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }

// ------------------------------------------------------------------------------------ //
// shared-memory hash tables
// ------------------------------------------------------------------------------------ //
template <int C>
struct Tables {
    int* rkey;                  // [RSLOTS]
    unsigned* rarea;            // [RSLOTS]
    unsigned* rborder;          // [RSLOTS]
    unsigned* rsum;             // [C][RSLOTS]
    unsigned* rsq;              // [C][RSLOTS]
    unsigned long long* ekey;   // [ESLOTS]
    unsigned* ecnt;             // [ESLOTS]
    unsigned* used;             // [0] region slots used, [1] edge slots used
};

__device__ __forceinline__ int region_slot(int* rkey, unsigned* used, int label) {
    unsigned h = ((unsigned)label * 0x9E3779B1u) >> (32 - 9);
    static_assert(RSLOTS == 512, "hash shift");
#pragma unroll 1
    for (int p = 0; p < MAX_PROBE; ++p) {
        int k = rkey[h];
        if (k == label) return (int)h;
        if (k == EMPTY_LABEL) {
            int old = atomicCAS(&rkey[h], EMPTY_LABEL, label);
            if (old == EMPTY_LABEL) {
                atomicAdd(&used[0], 1u);
                return (int)h;
            }
            if (old == label) return (int)h;
        }
        h = (h + 1) & (RSLOTS - 1);
    }
    return -1;
}

__device__ __forceinline__ void raw_append(const Params& P, unsigned long long key, unsigned cnt) {
    if ((long long)key_hi(key) >= P.n_regions) {
        atomicExch(&P.counters[3], 1ull);
        return;
    }
    unsigned long long i = atomicAdd(&P.counters[1], 1ull);
    if ((long long)i < P.capacity) {
        P.raw_keys[i] = key;
        P.raw_cnt[i] = cnt;
    } else {
        atomicExch(&P.counters[2], 1ull);
    }
}

template <int C>
__device__ __forceinline__ void edge_add(const Tables<C>& T, const Params& P, unsigned long long key, unsigned cnt) {
    unsigned h = (((unsigned)(key >> 32) * 0x9E3779B1u) ^ ((unsigned)key * 0x85EBCA6Bu)) >> (32 - 10);
    static_assert(ESLOTS == 1024, "hash shift");
#pragma unroll 1
    for (int p = 0; p < MAX_PROBE; ++p) {
        unsigned long long k = T.ekey[h];
        if (k == EMPTY_KEY) {
            k = atomicCAS(&T.ekey[h], EMPTY_KEY, key);
            if (k == EMPTY_KEY) {
                atomicAdd(&T.used[1], 1u);
                k = key;
            }
        }
        if (k == key) {
            atomicAdd(&T.ecnt[h], cnt);
            return;
        }
        h = (h + 1) & (ESLOTS - 1);
    }
    raw_append(P, key, cnt);   // table saturated: straight to the global list
}

// ------------------------------------------------------------------------------------ //
// per-thread register caches
// ------------------------------------------------------------------------------------ //
template <int C>
struct Acc {
    int label;
    unsigned area, border;
    unsigned s[C > 0 ? C : 1], q[C > 0 ? C : 1];
    __device__ __forceinline__ void reset(int l) {
        label = l;
        area = border = 0;
#pragma unroll
        for (int c = 0; c < C; ++c) s[c] = q[c] = 0;
    }
};

template <int C>
__device__ __forceinline__ void global_region_add(const Params& P, int label, unsigned area, unsigned border,
                                                  const unsigned* s, const unsigned* q) {
    if ((unsigned)label >= (unsigned)P.n_regions) {
        atomicExch(&P.counters[3], 1ull);
        return;
    }
    if (area) atomicAdd(&P.area[label], (unsigned long long)area);
    if (border) atomicAdd(&P.border[label], (unsigned long long)border);
    if (C > 0 && area) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            atomicAdd(&P.bsum[(size_t)label * C + c], (unsigned long long)s[c]);
            atomicAdd(&P.bsq[(size_t)label * C + c], (unsigned long long)q[c]);
        }
    }
}

template <int C>
__device__ __forceinline__ void acc_flush(const Tables<C>& T, const Params& P, Acc<C>& a) {
    if (a.label < 0 || (a.area | a.border) == 0) return;
    const int slot = region_slot(T.rkey, T.used, a.label);
    if (slot >= 0) {
        if (a.area) atomicAdd(&T.rarea[slot], a.area);
        if (a.border) atomicAdd(&T.rborder[slot], a.border);
        if (C > 0 && a.area) {
#pragma unroll
            for (int c = 0; c < C; ++c) {
                atomicAdd(&T.rsum[c * RSLOTS + slot], a.s[c]);
                atomicAdd(&T.rsq[c * RSLOTS + slot], a.q[c]);
            }
        }
    } else {
        global_region_add<C>(P, a.label, a.area, a.border, a.s, a.q);
    }
    a.area = a.border = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) a.s[c] = a.q[c] = 0;
}

// byte mask (0xFF per matching pixel) of the 4 pixels whose label equals L
__device__ __forceinline__ unsigned match4(const int4& a, int L) {
    unsigned m = (a.x == L) ? 0x000000ffu : 0u;
    m |= (a.y == L) ? 0x0000ff00u : 0u;
    m |= (a.z == L) ? 0x00ff0000u : 0u;
    m |= (a.w == L) ? 0xff000000u : 0u;
    return m;
}
__device__ __forceinline__ unsigned neg4(const int4& a) {
    unsigned m = (a.x < 0) ? 0x000000ffu : 0u;
    m |= (a.y < 0) ? 0x0000ff00u : 0u;
    m |= (a.z < 0) ? 0x00ff0000u : 0u;
    m |= (a.w < 0) ? 0xff000000u : 0u;
    return m;
}
__device__ __forceinline__ int pick4(const int4& a, int i) { return i == 0 ? a.x : i == 1 ? a.y : i == 2 ? a.z : a.w; }

template <int C>
__device__ __forceinline__ void acc_pixels(Acc<C>& a, unsigned bm, const unsigned* T) {
    a.area += __popc(bm) >> 3;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const unsigned w = T[c] & bm;
        a.s[c] = __dp4a(w, 0x01010101u, a.s[c]);
        a.q[c] = __dp4a(w, w, a.q[c]);
    }
}

// T[c] = the 4 pixels' values of band c, one per byte, from the 4*C interleaved bytes W[]
template <int C>
__device__ __forceinline__ void band_transpose(const unsigned* W, unsigned* T) {
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int i0 = 0 * C + c, i1 = 1 * C + c, i2 = 2 * C + c, i3 = 3 * C + c;
        const unsigned t01 = __byte_perm(W[i0 >> 2], W[i1 >> 2], (i0 & 3) | ((4 + (i1 & 3)) << 4));
        const unsigned t23 = __byte_perm(W[i2 >> 2], W[i3 >> 2], (i2 & 3) | ((4 + (i3 & 3)) << 4));
        T[c] = __byte_perm(t01, t23, 0x5410);
    }
}

template <int C>
struct Thread {
    Acc<C> c0, c1;
    unsigned long long ekey;
    unsigned ecnt;

    __device__ __forceinline__ void init() {
        c0.reset(EMPTY_LABEL);
        c1.reset(EMPTY_LABEL);
        ekey = EMPTY_KEY;
        ecnt = 0;
    }
    __device__ __forceinline__ void edge_flush(const Tables<C>& T, const Params& P) {
        if (ecnt) edge_add<C>(T, P, ekey, ecnt);
        ecnt = 0;
    }
    __device__ __forceinline__ void flush_all(const Tables<C>& T, const Params& P) {
        acc_flush<C>(T, P, c0);
        acc_flush<C>(T, P, c1);
        edge_flush(T, P);
    }
    __device__ __forceinline__ void border_add(const Tables<C>& T, const Params& P, int v, unsigned n) {
        if (v == c0.label) c0.border += n;
        else if (v == c1.label) c1.border += n;
        else {
            const int slot = region_slot(T.rkey, T.used, v);
            if (slot >= 0) atomicAdd(&T.rborder[slot], n);
            else global_region_add<C>(P, v, 0, n, nullptr, nullptr);
        }
    }
    // a pixel pair with different labels
    __device__ __forceinline__ void pair(const Tables<C>& T, const Params& P, int a, int b, unsigned n) {
        if ((a | b) >= 0) {
            const unsigned long long k = pack_key(a, b);
            if (k != ekey) {
                edge_flush(T, P);
                ekey = k;
            }
            ecnt += n;
        } else {
            const int v = a >= 0 ? a : b;   // the side of a valid pixel facing nodata
            if (v >= 0) border_add(T, P, v, n);
        }
    }
};

// ------------------------------------------------------------------------------------ //
// the kernel
// ------------------------------------------------------------------------------------ //
template <typename CF, bool USE_TMA>
__global__ void __launch_bounds__(CF::THREADS, 1)
rag_pool_kernel(const __grid_constant__ CUtensorMap mapL, const __grid_constant__ CUtensorMap mapI, const Params P) {
    constexpr int C = CF::C;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    unsigned char* stage_base = smem;
    unsigned* tab = (unsigned*)(smem + CF::STAGES * CF::STAGE_BYTES);
    Tables<C> T;
    T.rkey = (int*)tab;
    T.rarea = tab + RSLOTS;
    T.rborder = tab + 2 * RSLOTS;
    T.rsum = tab + 3 * RSLOTS;
    T.rsq = tab + (3 + C) * RSLOTS;
    T.ekey = (unsigned long long*)(tab + (3 + 2 * C) * RSLOTS);
    T.ecnt = tab + (3 + 2 * C) * RSLOTS + 2 * ESLOTS;
    unsigned* ctrl = tab + CF::TABLE_WORDS;             // 64 words of control space
    T.used = ctrl;                                       // [0],[1]
    uint64_t* full_bar = (uint64_t*)(ctrl + 8);          // [STAGES]
    uint64_t* empty_bar = full_bar + CF::STAGES;         // [STAGES]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = P.tiles_x * P.tiles_y;
    const int t_begin = min(total_tiles, (int)blockIdx.x * P.tiles_per_cta);
    const int t_end = min(total_tiles, t_begin + P.tiles_per_cta);
    const int my_tiles = t_end - t_begin;

    // ---- init -------------------------------------------------------------------------
    for (int i = threadIdx.x; i < RSLOTS; i += CF::THREADS) {
        T.rkey[i] = EMPTY_LABEL;
        T.rarea[i] = 0;
        T.rborder[i] = 0;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            T.rsum[c * RSLOTS + i] = 0;
            T.rsq[c * RSLOTS + i] = 0;
        }
    }
    for (int i = threadIdx.x; i < ESLOTS; i += CF::THREADS) {
        T.ekey[i] = EMPTY_KEY;
        T.ecnt[i] = 0;
    }
    if (threadIdx.x == 0) {
        T.used[0] = T.used[1] = 0;
        if (USE_TMA) {
            for (int s = 0; s < CF::STAGES; ++s) {
                mbar_init(&full_bar[s], 1);
                mbar_init(&empty_bar[s], CF::NCW);
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    __syncthreads();

    if (warp == CF::NCW) {
        // ================================ producer warp ================================
        if (USE_TMA && lane == 0) {
            for (int i = 0; i < my_tiles; ++i) {
                const int t = t_begin + i;
                const int tx = t / P.tiles_y, ty = t - tx * P.tiles_y;
                const int st = i % CF::STAGES;
                const unsigned ph = (unsigned)(i / CF::STAGES) & 1u;
                mbar_wait(&empty_bar[st], ph ^ 1u, P.counters);
                int strips = 0;
#pragma unroll
                for (int s = 0; s < CF::STRIPS; ++s) strips += (tx * CF::TILE_W + s * STRIP_W < P.W) ? 1 : 0;
                mbar_expect_tx(&full_bar[st], (unsigned)(strips * ((CF::TH + 1) * LAB_PITCH * 4 +
                                                                   (C > 0 ? CF::TH * CF::IMG_ROW_WORDS * 4 : 0))));
                unsigned char* sb = stage_base + (size_t)st * CF::STAGE_BYTES;
#pragma unroll
                for (int s = 0; s < CF::STRIPS; ++s) {
                    const int x0 = tx * CF::TILE_W + s * STRIP_W;
                    if (x0 < P.W) {
                        tma_load_2d(sb + s * CF::LAB_BOX, &mapL, x0, ty * CF::TH, &full_bar[st]);
                        if (C > 0)
                            tma_load_2d(sb + CF::STRIPS * CF::LAB_BOX + s * CF::IMG_BOX, &mapI, x0 * C / 4, ty * CF::TH,
                                        &full_bar[st]);
                    }
                }
            }
        }
        return;
    }

    // ================================== compute warps ==================================
    const int strip = warp % CF::STRIPS, band = warp / CF::STRIPS;
    Thread<C> th;
    th.init();
    int tiles_since_flush = 0;

    for (int i = 0; i < my_tiles; ++i) {
        const int t = t_begin + i;
        const int tx = t / P.tiles_y, ty = t - tx * P.tiles_y;
        const int st = USE_TMA ? i % CF::STAGES : 0;
        unsigned char* sb = stage_base + (size_t)st * CF::STAGE_BYTES;
        const int tile_x0 = tx * CF::TILE_W, tile_y0 = ty * CF::TH;

        if (USE_TMA) {
            mbar_wait(&full_bar[st], (unsigned)(i / CF::STAGES) & 1u, P.counters);
        } else {
            // fallback staging for rasters whose pitch/base TMA cannot describe
            for (int s = 0; s < CF::STRIPS; ++s) {
                int* L = (int*)(sb + s * CF::LAB_BOX);
                const int x0 = tile_x0 + s * STRIP_W;
                for (int k = threadIdx.x; k < (CF::TH + 1) * LAB_PITCH; k += CF::NCT) {
                    const int r = k / LAB_PITCH, cidx = k - r * LAB_PITCH;
                    const int gy = tile_y0 + r, gx = x0 + cidx;
                    L[k] = (gy < P.rows_avail && gx < P.W) ? P.labels[(int64_t)gy * P.ld + gx] : 0;
                }
                if constexpr (C > 0) {
                    unsigned char* I = sb + CF::STRIPS * CF::LAB_BOX + s * CF::IMG_BOX;
                    for (int k = threadIdx.x; k < CF::TH * CF::IMG_ROW_WORDS * 4; k += CF::NCT) {
                        const int r = k / (CF::IMG_ROW_WORDS * 4), bidx = k - r * (CF::IMG_ROW_WORDS * 4);
                        const int gy = tile_y0 + r;
                        const int64_t gb = (int64_t)x0 * C + bidx;
                        I[k] = (gy < P.rows_own && gb < (int64_t)P.W * C) ? P.image[(int64_t)gy * P.image_pitch + gb] : 0;
                    }
                }
            }
            compute_bar(CF::NCT);
        }

        const int x0 = tile_x0 + strip * STRIP_W + 4 * lane;   // first of this lane's 4 pixels
        if (tile_x0 + strip * STRIP_W < P.W) {
            const int* L = (const int*)(sb + strip * CF::LAB_BOX);
            const unsigned* I = (const unsigned*)(sb + CF::STRIPS * CF::LAB_BOX + strip * CF::IMG_BOX);
            const int nin = min(4, max(0, P.W - x0));          // pixels of this lane inside the image
            const int r0 = band * CF::BR;
            int4 own = *(const int4*)(L + r0 * LAB_PITCH + 4 * lane);
#pragma unroll 1
            for (int r = r0; r < r0 + CF::BR; ++r) {
                const int y = tile_y0 + r;
                if (y >= P.rows_own) break;
                const int4 dn = *(const int4*)(L + (r + 1) * LAB_PITCH + 4 * lane);
                int right = __shfl_down_sync(0xffffffffu, own.x, 1);
                if (lane == 31) right = L[r * LAB_PITCH + STRIP_W];
                unsigned W[CF::CW], TB[CF::CW];
                if (C > 0) {
                    if (C == 4) {
                        const uint4 v = *(const uint4*)(I + r * CF::IMG_ROW_WORDS + 4 * lane);
                        W[0] = v.x; W[1 % CF::CW] = v.y; W[2 % CF::CW] = v.z; W[3 % CF::CW] = v.w;
                    } else {
#pragma unroll
                        for (int c = 0; c < C; ++c) W[c] = I[r * CF::IMG_ROW_WORDS + C * lane + c];
                    }
                    band_transpose<C>(W, TB);
                }
                int4 a = own;
                const bool interior = (nin == 4) && (x0 + 4 < P.W) && (x0 > 0) && (y > 0) && (y + 1 < P.rows_avail);
                if (!interior) {   // out-of-image pixels behave like "no pixel": never counted, never paired
                    if (nin < 4) a.w = EMPTY_LABEL;
                    if (nin < 3) a.z = EMPTY_LABEL;
                    if (nin < 2) a.y = EMPTY_LABEL;
                    if (nin < 1) a.x = EMPTY_LABEL;
                }
                // ---- accumulate the 4 pixels into the 2-entry label cache -----------------
                unsigned m0 = match4(a, th.c0.label), m1 = match4(a, th.c1.label);
                unsigned covered = m0 | m1 | neg4(a);
                while (covered != 0xffffffffu) {               // rare: a label outside the cache
                    const int i4 = (__ffs(~covered) - 1) >> 3;
                    const int Lb = pick4(a, i4);
                    if (m0 == 0 || th.c0.label < 0) {
                        acc_flush<C>(T, P, th.c0);
                        th.c0.reset(Lb);
                        m0 = match4(a, Lb);
                        covered |= m0;
                    } else if (m1 == 0 || th.c1.label < 0) {
                        acc_flush<C>(T, P, th.c1);
                        th.c1.reset(Lb);
                        m1 = match4(a, Lb);
                        covered |= m1;
                    } else {                                   // >2 labels in 4 pixels: uncached add
                        Acc<C> one;
                        one.reset(Lb);
                        const unsigned bm = 0xffu << (8 * i4);
                        acc_pixels<C>(one, bm, TB);
                        acc_flush<C>(T, P, one);
                        covered |= bm;
                    }
                }
                acc_pixels<C>(th.c0, m0, TB);
                acc_pixels<C>(th.c1, m1, TB);
                // ---- neighbour pairs ----------------------------------------------------------
                if (interior) {
                    if (a.x != a.y) th.pair(T, P, a.x, a.y, 1);
                    if (a.y != a.z) th.pair(T, P, a.y, a.z, 1);
                    if (a.z != a.w) th.pair(T, P, a.z, a.w, 1);
                    if (a.w != right) th.pair(T, P, a.w, right, 1);
                    const bool urow = (a.x == a.y) & (a.y == a.z) & (a.z == a.w);
                    const bool udn = (dn.x == dn.y) & (dn.y == dn.z) & (dn.z == dn.w);
                    if (urow & udn) {
                        if (a.x != dn.x) th.pair(T, P, a.x, dn.x, 4);
                    } else {
                        if (a.x != dn.x) th.pair(T, P, a.x, dn.x, 1);
                        if (a.y != dn.y) th.pair(T, P, a.y, dn.y, 1);
                        if (a.z != dn.z) th.pair(T, P, a.z, dn.z, 1);
                        if (a.w != dn.w) th.pair(T, P, a.w, dn.w, 1);
                    }
                } else if (nin > 0) {
                    // image borders, partial lanes, first/last rows: pixel by pixel
                    const bool has_dn = (y + 1 < P.rows_avail);
                    const bool top = (y == 0) && P.top_border;
                    const bool bot = (!has_dn) && (y == P.rows_own - 1) && P.bottom_border;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (k < nin) {
                            const int l = pick4(a, k);
                            const int x = x0 + k;
                            if (x + 1 < P.W) {
                                const int rn = (k == 3) ? right : pick4(a, k + 1);
                                if (l != rn) th.pair(T, P, l, rn, 1);
                            } else if (l >= 0) {
                                th.border_add(T, P, l, 1);      // right image border
                            }
                            if (x == 0 && l >= 0) th.border_add(T, P, l, 1);
                            if (has_dn) {
                                const int d = pick4(dn, k);
                                if (l != d) th.pair(T, P, l, d, 1);
                            } else if (bot && l >= 0) {
                                th.border_add(T, P, l, 1);
                            }
                            if (top && l >= 0) th.border_add(T, P, l, 1);
                        }
                    }
                }
                own = dn;
            }
        }
        th.flush_all(T, P);
        if (USE_TMA) {
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[st]);
        }

        // ---- table maintenance (all compute warps) -----------------------------------------
        ++tiles_since_flush;
        compute_bar(CF::NCT);
        const bool need = (T.used[0] > RSLOTS / 2) || (T.used[1] > ESLOTS / 2) || (tiles_since_flush >= CF::FLUSH_TILES) ||
                          (i + 1 == my_tiles);
        compute_bar(CF::NCT);
        if (need) {
            for (int k = threadIdx.x; k < RSLOTS; k += CF::NCT) {
                const int label = T.rkey[k];
                if (label != EMPTY_LABEL) {
                    unsigned s[CF::CW], q[CF::CW];
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        s[c] = T.rsum[c * RSLOTS + k];
                        q[c] = T.rsq[c * RSLOTS + k];
                        T.rsum[c * RSLOTS + k] = 0;
                        T.rsq[c * RSLOTS + k] = 0;
                    }
                    global_region_add<C>(P, label, T.rarea[k], T.rborder[k], s, q);
                    T.rkey[k] = EMPTY_LABEL;
                    T.rarea[k] = 0;
                    T.rborder[k] = 0;
                }
            }
            for (int k0 = 0; k0 < ESLOTS; k0 += CF::NCT) {
                const int k = k0 + threadIdx.x;
                unsigned long long key = k < ESLOTS ? T.ekey[k] : EMPTY_KEY;
                unsigned cnt = 0;
                if (key != EMPTY_KEY) {
                    cnt = T.ecnt[k];
                    T.ekey[k] = EMPTY_KEY;
                    T.ecnt[k] = 0;
                    if ((long long)key_hi(key) >= P.n_regions) {   // label outside [0, n_regions)
                        atomicExch(&P.counters[3], 1ull);
                        key = EMPTY_KEY;
                    }
                }
                const bool has = key != EMPTY_KEY;
                const unsigned bal = __ballot_sync(0xffffffffu, has);
                if (bal) {
                    unsigned long long base = 0;
                    if (lane == 0) base = atomicAdd(&P.counters[1], (unsigned long long)__popc(bal));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (has) {
                        const unsigned long long idx = base + __popc(bal & lanemask_lt());
                        if ((long long)idx < P.capacity) {
                            P.raw_keys[idx] = key;
                            P.raw_cnt[idx] = cnt;
                        } else {
                            atomicExch(&P.counters[2], 1ull);
                        }
                    }
                }
            }
            if (threadIdx.x == 0) T.used[0] = T.used[1] = 0;
            tiles_since_flush = 0;
            compute_bar(CF::NCT);
        }
    }
}

// ------------------------------------------------------------------------------------ //
// host side
// ------------------------------------------------------------------------------------ //
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

static bool make_map_2d(CUtensorMap* m, const void* base, uint64_t dim0, uint64_t dim1, uint64_t pitch_bytes,
                        uint32_t box0, uint32_t box1) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {dim0, dim1};
    cuuint64_t strides[1] = {pitch_bytes};
    cuuint32_t box[2] = {box0, box1};
    cuuint32_t estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, const_cast<void*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <typename CF>
static int launch(const Params& Pin, bool allow_tma, cudaStream_t s) {
    Params P = Pin;
    P.tiles_x = (int)ceil_div(P.W, CF::TILE_W);
    P.tiles_y = (int)ceil_div(P.rows_own, CF::TH);
    const int total = P.tiles_x * P.tiles_y;
    if (total == 0) return DM_OK;
    const int grid = min(total, num_sms());
    P.tiles_per_cta = (int)ceil_div(total, grid);
    const int grid2 = (int)ceil_div(total, P.tiles_per_cta);

    CUtensorMap mapL, mapI;
    memset(&mapL, 0, sizeof(mapL));
    memset(&mapI, 0, sizeof(mapI));
    // TMA needs 16-byte aligned bases and row pitches; anything else takes the ld.global staging path
    bool tma = allow_tma && ((uintptr_t)P.labels % 16 == 0) && ((P.ld * 4) % 16 == 0) && P.ld >= P.W;
    if (CF::C > 0)
        tma = tma && ((uintptr_t)P.image % 16 == 0) && (P.image_pitch % 16 == 0) && (((int64_t)P.W * CF::C) % 4 == 0);
    if (tma)
        tma = make_map_2d(&mapL, P.labels, (uint64_t)P.W, (uint64_t)P.rows_avail, (uint64_t)P.ld * 4, LAB_PITCH, CF::TH + 1);
    if (tma && CF::C > 0)
        tma = make_map_2d(&mapI, P.image, (uint64_t)P.W * CF::C / 4, (uint64_t)P.rows_own, (uint64_t)P.image_pitch,
                          CF::IMG_ROW_WORDS, CF::TH);
    if (tma) {
        auto k = rag_pool_kernel<CF, true>;
        DM_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::SMEM_BYTES));
        DM_COUNT_LAUNCH(); k<<<grid2, CF::THREADS, CF::SMEM_BYTES, s>>>(mapL, mapI, P);
    } else {
        auto k = rag_pool_kernel<CF, false>;
        DM_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::SMEM_BYTES));
        DM_COUNT_LAUNCH(); k<<<grid2, CF::THREADS, CF::SMEM_BYTES, s>>>(mapL, mapI, P);
    }
    DM_LAUNCH_CHECK();
    return DM_OK;
}

__global__ void clamp_count_kernel(const int64_t* raw, int64_t capacity, int64_t* out) {
    *out = *raw < capacity ? *raw : capacity;
}
__global__ void copy_back_kernel(const uint64_t* __restrict__ k, const uint32_t* __restrict__ l,
                                 const int64_t* __restrict__ n_dev, uint64_t* __restrict__ ko, uint32_t* __restrict__ lo) {
    const int64_t n = *n_dev;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        ko[e] = k[e];
        lo[e] = l[e];
    }
}

// DM_RAG_NO_TMA=1 forces the ld.global staging path (tests exercise both).
static bool allow_tma_env() {
    const char* e = getenv("DM_RAG_NO_TMA");
    return !(e && e[0] == '1');
}

int run(const Params& P, int C, cudaStream_t s) {
    const bool tma = allow_tma_env();
    switch (C) {
        case 0: return launch<Cfg<0, 32, 2, 4, 3>>(P, tma, s);
        case 1: return launch<Cfg<1, 32, 2, 4, 3>>(P, tma, s);
        case 2: return launch<Cfg<2, 32, 2, 4, 2>>(P, tma, s);
        case 3: return launch<Cfg<3, 32, 2, 4, 2>>(P, tma, s);
        case 4: return launch<Cfg<4, 32, 2, 4, 2>>(P, tma, s);
        default: return DM_ERR_BAD_ARG;
    }
}

}  // namespace rag
}  // namespace dm

using namespace dm;

extern "C" size_t dm_rag_workspace_bytes(int64_t capacity) {
    const int64_t cap = capacity < 1 ? 1 : capacity;
    return align_up((size_t)cap * 8, 256) + align_up((size_t)cap * 4, 256) + 256 + prims::sort_ws_bytes(cap) +
           prims::unique_ws_bytes(cap);
}

namespace {
struct RagWs {
    uint64_t* raw_keys;
    uint32_t* raw_cnt;
    int64_t* n_raw;
    void* sws;
    void* uws;
};
RagWs carve_rag_ws(void* ws, int64_t capacity) {
    const int64_t cap = capacity < 1 ? 1 : capacity;
    Carver c(ws);
    RagWs w;
    w.raw_keys = c.take<uint64_t>(cap);
    w.raw_cnt = c.take<uint32_t>(cap);
    w.n_raw = c.take<int64_t>(1);
    w.sws = c.take<char>(prims::sort_ws_bytes(cap));
    w.uws = c.take<char>(prims::unique_ws_bytes(cap));
    return w;
}
}  // namespace

extern "C" int dm_rag_scan(const int32_t* labels, int64_t rows_own, int64_t rows_avail, int64_t W, int64_t ld,
                           const uint8_t* image, int64_t C, int64_t image_pitch, int64_t n_regions, int top_border,
                           int bottom_border, int64_t* area, int64_t* border, uint64_t* band_sum, uint64_t* band_sumsq,
                           int64_t capacity, int64_t* counts, void* ws, size_t ws_bytes, dm_stream_t stream) {
    if (rows_own < 0 || W < 0 || ld < W || n_regions < 0 || n_regions > 0x7fffffff || capacity < 0) return DM_ERR_BAD_ARG;
    if (rows_avail != rows_own && rows_avail != rows_own + 1) return DM_ERR_BAD_ARG;
    if (rows_own > 0x7ffffff0 || W > 0x7ffffff0) return DM_ERR_BAD_ARG;
    if (!counts) return DM_ERR_BAD_ARG;
    if (!image) C = 0;
    if (C < 0 || C > 4) return DM_ERR_BAD_ARG;
    if (C > 0 && (!band_sum || !band_sumsq || image_pitch < W * C)) return DM_ERR_BAD_ARG;
    cudaStream_t s = S(stream);
    DM_CUDA(cudaMemsetAsync(counts, 0, 4 * sizeof(int64_t), s));
    if (rows_own == 0 || W == 0) return DM_OK;
    if (!labels || !area || !border || !ws) return DM_ERR_BAD_ARG;
    if (ws_bytes < dm_rag_workspace_bytes(capacity)) return DM_ERR_WORKSPACE;
    RagWs w = carve_rag_ws(ws, capacity);
    rag::Params P;
    P.labels = labels;
    P.ld = ld;
    P.image = image;
    P.image_pitch = image_pitch;
    P.rows_own = (int)rows_own;
    P.rows_avail = (int)rows_avail;
    P.W = (int)W;
    P.n_regions = (int)n_regions;
    P.top_border = top_border;
    P.bottom_border = bottom_border;
    P.area = (unsigned long long*)area;
    P.border = (unsigned long long*)border;
    P.bsum = (unsigned long long*)band_sum;
    P.bsq = (unsigned long long*)band_sumsq;
    P.raw_keys = (unsigned long long*)w.raw_keys;
    P.raw_cnt = w.raw_cnt;
    P.capacity = capacity;
    P.counters = (unsigned long long*)counts;
    P.tiles_x = P.tiles_y = P.tiles_per_cta = 0;
    return rag::run(P, (int)C, s);
}

extern "C" int dm_rag_finish(uint64_t* edge_keys, uint32_t* boundary_len, int64_t capacity, int64_t n_regions,
                             int64_t* counts, void* ws, size_t ws_bytes, dm_stream_t stream) {
    if (capacity < 0 || n_regions < 0 || !counts) return DM_ERR_BAD_ARG;
    if (capacity == 0) return DM_OK;
    if (!edge_keys || !boundary_len || !ws) return DM_ERR_BAD_ARG;
    if (ws_bytes < dm_rag_workspace_bytes(capacity)) return DM_ERR_WORKSPACE;
    cudaStream_t s = S(stream);
    RagWs w = carve_rag_ws(ws, capacity);
    // n_raw = min(raw entries, capacity): what actually sits in the raw list
    DM_COUNT_LAUNCH(); rag::clamp_count_kernel<<<1, 1, 0, s>>>(counts + 1, capacity, w.n_raw);
    const int b = bits_for(n_regions);
    DM_TRY(prims::sort_pairs(w.raw_keys, w.raw_cnt, w.n_raw, capacity, b, 2 * b, w.sws, s));
    DM_TRY(prims::unique_reduce(w.raw_keys, nullptr, w.raw_cnt, nullptr, w.n_raw, capacity, ~0ull, edge_keys, boundary_len,
                                nullptr, counts, w.uws, s));
    DM_LAUNCH_CHECK();
    return DM_OK;
}

extern "C" int dm_rag_build(const int32_t* labels, int64_t rows_own, int64_t rows_avail, int64_t W, int64_t ld,
                            const uint8_t* image, int64_t C, int64_t image_pitch, int64_t n_regions, int top_border,
                            int bottom_border, int64_t* area, int64_t* border, uint64_t* band_sum, uint64_t* band_sumsq,
                            uint64_t* edge_keys, uint32_t* boundary_len, int64_t capacity, int64_t* counts, void* ws,
                            size_t ws_bytes, dm_stream_t stream) {
    if (capacity > 0 && (!edge_keys || !boundary_len)) return DM_ERR_BAD_ARG;
    DM_TRY(dm_rag_scan(labels, rows_own, rows_avail, W, ld, image, C, image_pitch, n_regions, top_border, bottom_border,
                       area, border, band_sum, band_sumsq, capacity, counts, ws, ws_bytes, stream));
    if (rows_own == 0 || W == 0) return DM_OK;
    return dm_rag_finish(edge_keys, boundary_len, capacity, n_regions, counts, ws, ws_bytes, stream);
}

extern "C" size_t dm_edges_unique_workspace_bytes(int64_t capacity) {
    const int64_t cap = capacity < 1 ? 1 : capacity;
    return align_up((size_t)cap * 8, 256) + align_up((size_t)cap * 4, 256) + 256 + prims::sort_ws_bytes(cap) +
           prims::unique_ws_bytes(cap);
}

extern "C" int dm_edges_sort_unique(uint64_t* keys, uint32_t* lens, const int64_t* n_in, int64_t capacity, int64_t n_regions,
                                    int64_t* n_out, void* ws, size_t ws_bytes, dm_stream_t stream) {
    if (capacity < 0 || n_regions < 0 || !n_out) return DM_ERR_BAD_ARG;
    cudaStream_t s = S(stream);
    if (capacity == 0) {
        DM_CUDA(cudaMemsetAsync(n_out, 0, sizeof(int64_t), s));
        return DM_OK;
    }
    if (!keys || !lens || !n_in || !ws) return DM_ERR_BAD_ARG;
    if (ws_bytes < dm_edges_unique_workspace_bytes(capacity)) return DM_ERR_WORKSPACE;
    Carver c(ws);
    uint64_t* ok = c.take<uint64_t>(capacity);
    uint32_t* ol = c.take<uint32_t>(capacity);
    int64_t* n_new = c.take<int64_t>(1);
    void* sws = c.take<char>(prims::sort_ws_bytes(capacity));
    void* uws = c.take<char>(prims::unique_ws_bytes(capacity));
    const int b = bits_for(n_regions);
    DM_TRY(prims::sort_pairs(keys, lens, n_in, capacity, b, 2 * b, sws, s));
    DM_TRY(prims::unique_reduce(keys, nullptr, lens, nullptr, n_in, capacity, ~0ull, ok, ol, nullptr, n_new, uws, s));
    DM_COUNT_LAUNCH(); rag::copy_back_kernel<<<(unsigned)imax64(1, imin64(ceil_div(capacity, 256), (int64_t)num_sms() * 8)), 256, 0, s>>>(
        ok, ol, n_new, keys, lens);
    DM_CUDA(cudaMemcpyAsync(n_out, n_new, sizeof(int64_t), cudaMemcpyDeviceToDevice, s));
    DM_LAUNCH_CHECK();
    return DM_OK;
}
