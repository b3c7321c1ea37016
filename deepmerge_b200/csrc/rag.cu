// R1: region adjacency graph from a label raster, fused with per-region band pooling.
//
// Warp-autonomous streaming design.  The raster is cut into 128-pixel-wide strips; every
// warp of a persistent grid (one 16-warp CTA per SM) owns a contiguous run of rows of one
// strip and walks straight down it:
//   * lane 0 of the warp feeds the warp's OWN ring of shared-memory stages with TMA
//     (cp.async.bulk.tensor boxes of TH+1 label rows x 132 columns -- 1-row / 4-column halo
//     -- and TH image rows), completion signalled on the warp's own mbarriers.  No producer
//     warp, no block-level barrier anywhere after start-up: warps never wait for each other.
//   * every lane owns 4 consecutive pixels (one 128-bit LDS of labels, one of image bytes)
//     and keeps a 2-entry register cache of per-label accumulators: area, border sides, C
//     band sums and sums of squares, fed 4 pixels at a time with PRMT + DP4A on byte masks.
//     Because the warp moves down contiguous rows, a cache entry lives for the whole height
//     of a region.  Pixel pairs straddling the two cached labels are counted with the same
//     byte masks (no per-pair work, no divergence); a third label nearby takes a slow path.
//   * evictions go to the warp's private shared-memory hash tables (label -> accumulators,
//     edge key -> pair count), which are drained to global memory with 64-bit atomics /
//     appended (key,count) entries when half full.  The appended entries are then radix
//     sorted and run-reduced (prims.cu) into the sorted unique edge list.
//
// HBM traffic: labels 4 B/px + image C B/px read once (halo re-reads hit L2).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "prims.cuh"

namespace dm {
namespace rag {

constexpr int STRIP_W = 128;            // pixels per warp row: 32 lanes x 4
constexpr int LAB_PITCH = STRIP_W + 4;  // + halo columns (TMA inner box must be a multiple of 16 B)
constexpr int NWARPS = 16;              // warps per CTA, each an independent pipeline
constexpr int RS = 32;                  // region table slots per warp (power of two)
constexpr int ES = 64;                  // edge table slots per warp (power of two)
constexpr int RQ = 32;                  // region eviction queue entries per warp (one per lane when drained)
constexpr int EQ = 64;                  // edge eviction queue entries per warp
constexpr int FLUSH_ROWS = 256;         // forced drain period: 128 px * 256 rows * 255^2 < 2^32
constexpr int EMPTY_LABEL = -1;
constexpr unsigned long long EMPTY_KEY = ~0ull;
constexpr int SLOT_UNKNOWN = -2, SLOT_NONE = -1;

#if defined(DM_RAG_STATS) && !defined(DM_RAG_TIMING_ONLY)
#define STAT(i) atomicAdd(&P.counters[8 + (i)], 1ull)   // counts[] must hold >= 24 entries in a stats build
#define STAT_LAST(l, i) do { if ((l) == th.last_evicted) STAT(i); } while (0)
#else
#define STAT(i) ((void)0)
#define STAT_LAST(l, i) ((void)0)
#endif
// stats: 7 prefetch for right, 8 second prefetch in a row, 9/10 prefetch/own-miss of the label evicted last
// stats: 0 own-miss evict, 1 uncached add, 2 prefetch evict, 3 fast lane-rows, 4 slow lane-rows, 5 border lane-rows, 6 drains

constexpr int align128(int x) { return (x + 127) / 128 * 128; }

template <int C_, int TH_, int NS_>
struct Cfg {
    static constexpr int C = C_, TH = TH_, NS = NS_;
    static constexpr int CW = C_ > 0 ? C_ : 1;               // words of image bytes per lane-row
    static constexpr int THREADS = NWARPS * 32;
    static constexpr int LAB_BOX = align128((TH + 1) * LAB_PITCH * 4);
    static constexpr int IMG_ROW_WORDS = STRIP_W * C / 4;    // 32*C
    static constexpr int IMG_BOX = align128(TH * IMG_ROW_WORDS * 4);
    static constexpr int STAGE_BYTES = LAB_BOX + IMG_BOX;
    static constexpr int QUEUE_WORDS = RQ * (3 + 2 * C) + EQ * 3 + 2 + 2;   // eviction queues + counts[2] + pad
    static constexpr int TABLE_WORDS = RS * (3 + 2 * C) + ES * 3 + 2 + 2 + QUEUE_WORDS;   // + used[2] + pad
    static constexpr int TABLE_BYTES = align128(TABLE_WORDS * 4 + NS * 8);   // + full barriers
    static constexpr int WARP_BYTES = NS * STAGE_BYTES + TABLE_BYTES;
    static constexpr int SMEM_BYTES = 128 + NWARPS * WARP_BYTES;
    static constexpr int TX_BYTES = (TH + 1) * LAB_PITCH * 4 + (C > 0 ? TH * IMG_ROW_WORDS * 4 : 0);
    static constexpr int FLUSH_UNITS = FLUSH_ROWS / TH > 0 ? FLUSH_ROWS / TH : 1;
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
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
    int tiles_x, tiles_y, tiles_per_cta;   // strips, row blocks per strip, units per warp
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
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(100000u)   // suspend-time hint (ns): sleep in HW, do not spin
        : "memory");
    return ok != 0;
}
// Bounded wait: a TMA that never lands must become an error, not a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity, unsigned long long* counters) {
    unsigned spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 20)) {
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

// ------------------------------------------------------------------------------------ //
// per-warp shared-memory hash tables
// ------------------------------------------------------------------------------------ //
template <int C>
struct Tables {
    int* rkey;                  // [RS]
    unsigned* rarea;            // [RS]
    unsigned* rborder;          // [RS]
    unsigned* rsum;             // [C][RS]
    unsigned* rsq;              // [C][RS]
    unsigned long long* ekey;   // [ES]
    unsigned* ecnt;             // [ES]
    unsigned* used;             // [0] region slots used, [1] edge slots used
    // eviction queues: lanes push evicted accumulators with plain stores; the warp drains them together
    int* qlabel;                // [RQ]
    unsigned* qarea;            // [RQ]
    unsigned* qborder;          // [RQ]
    unsigned* qsum;             // [C][RQ]
    unsigned* qsq;              // [C][RQ]
    unsigned long long* qekey;  // [EQ]
    unsigned* qecnt;            // [EQ]
    unsigned* qn;               // [0] region entries, [1] edge entries
    unsigned* base;             // start of the warp's table arena (what the out-of-line slow paths rebuild from)

    __device__ __forceinline__ static Tables from(unsigned* tab) {
        Tables T;
        T.base = tab;
        T.rkey = (int*)tab;
        T.rarea = tab + RS;
        T.rborder = tab + 2 * RS;
        T.rsum = tab + 3 * RS;
        T.rsq = tab + (3 + C) * RS;
        T.ekey = (unsigned long long*)(tab + (3 + 2 * C) * RS);
        T.ecnt = tab + (3 + 2 * C) * RS + 2 * ES;
        T.used = tab + (3 + 2 * C) * RS + 3 * ES;
        unsigned* qb = T.used + 4;                                       // after used[2] + pad
        T.qlabel = (int*)qb;
        T.qarea = qb + RQ;
        T.qborder = qb + 2 * RQ;
        T.qsum = qb + 3 * RQ;
        T.qsq = qb + (3 + C) * RQ;
        T.qekey = (unsigned long long*)(qb + (3 + 2 * C) * RQ);
        T.qecnt = qb + (3 + 2 * C) * RQ + 2 * EQ;
        T.qn = qb + (3 + 2 * C) * RQ + 3 * EQ;
        return T;
    }
};

__device__ __forceinline__ int region_slot(int* rkey, unsigned* used, int label) {
    unsigned h = ((unsigned)label * 0x9E3779B1u) >> (32 - 5);
    static_assert(RS == 32, "hash shift");
#pragma unroll 1
    for (int p = 0; p < RS; ++p) {
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
        h = (h + 1) & (RS - 1);
    }
    return SLOT_NONE;
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
    unsigned h = (((unsigned)(key >> 32) * 0x9E3779B1u) ^ ((unsigned)key * 0x85EBCA6Bu)) >> (32 - 6);
    static_assert(ES == 64, "hash shift");
#pragma unroll 1
    for (int p = 0; p < ES / 2; ++p) {
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
        h = (h + 1) & (ES - 1);
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
    __device__ __forceinline__ void clear() {
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

// accumulators of one label -> the warp's region table (or straight to global when it is full)
template <int C>
__device__ __forceinline__ void table_region_add(const Tables<C>& T, const Params& P, int label, unsigned area,
                                                 unsigned border, const unsigned* s, const unsigned* q) {
    const int slot = region_slot(T.rkey, T.used, label);
    if (slot >= 0) {
        if (area) atomicAdd(&T.rarea[slot], area);
        if (border) atomicAdd(&T.rborder[slot], border);
        if (C > 0 && area) {
#pragma unroll
            for (int c = 0; c < C; ++c) {
                atomicAdd(&T.rsum[c * RS + slot], s[c]);
                atomicAdd(&T.rsq[c * RS + slot], q[c]);
            }
        }
    } else {
        global_region_add<C>(P, label, area, border, s, q);
    }
}

// Out-of-line slow paths (queue overflow, uncached border sides): kept out of the hot loop's code.
template <int C>
__device__ __noinline__ void slow_edge_add(unsigned* tab, const Params* P, unsigned long long key, unsigned cnt) {
    edge_add<C>(Tables<C>::from(tab), *P, key, cnt);
}
template <int C>
__device__ __noinline__ void slow_region_add(unsigned* tab, const Params* P, int label, unsigned area, unsigned border,
                                             uint4 s4, uint4 q4) {
    const unsigned s[4] = {s4.x, s4.y, s4.z, s4.w}, q[4] = {q4.x, q4.y, q4.z, q4.w};
    table_region_add<C>(Tables<C>::from(tab), *P, label, area, border, s, q);
}
template <int C>
__device__ __forceinline__ void slow_region_add_acc(const Tables<C>& T, const Params& P, int label, unsigned area,
                                                    unsigned border, const unsigned* s, const unsigned* q) {
    uint4 s4 = make_uint4(0, 0, 0, 0), q4 = make_uint4(0, 0, 0, 0);
    if (C > 0 && s) {
        s4.x = s[0]; q4.x = q[0];
        if (C > 1) { s4.y = s[1 % (C > 0 ? C : 1)]; q4.y = q[1 % (C > 0 ? C : 1)]; }
        if (C > 2) { s4.z = s[2 % (C > 0 ? C : 1)]; q4.z = q[2 % (C > 0 ? C : 1)]; }
        if (C > 3) { s4.w = s[3 % (C > 0 ? C : 1)]; q4.w = q[3 % (C > 0 ? C : 1)]; }
    }
    slow_region_add<C>(T.base, &P, label, area, border, s4, q4);
}

// Evicted accumulators are only PUSHED by the (few, divergent) evicting lanes; the expensive
// hash probe + atomics happen later in drain_queues with every queued entry on its own lane.
template <int C>
__device__ __forceinline__ void acc_push(const Tables<C>& T, const Params& P, Acc<C>& a) {
    if (a.label < 0 || (a.area | a.border) == 0) return;
    const unsigned p = atomicAdd(&T.qn[0], 1u);
    if (p < RQ) {
        T.qlabel[p] = a.label;
        T.qarea[p] = a.area;
        T.qborder[p] = a.border;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            T.qsum[c * RQ + p] = a.s[c];
            T.qsq[c * RQ + p] = a.q[c];
        }
    } else {                      // queue full (it is drained every unit): rare, do it the slow way
        slow_region_add_acc<C>(T, P, a.label, a.area, a.border, a.s, a.q);
    }
    a.clear();
}

template <int C>
__device__ __forceinline__ void edge_push(const Tables<C>& T, const Params& P, unsigned long long key, unsigned cnt) {
    const unsigned p = atomicAdd(&T.qn[1], 1u);
    if (p < EQ) {
        T.qekey[p] = key;
        T.qecnt[p] = cnt;
    } else {
        slow_edge_add<C>(T.base, &P, key, cnt);
    }
}

// Whole warp, convergent: queued entries -> hash tables.
template <int C>
__device__ __forceinline__ void drain_queues(const Tables<C>& T, const Params& P, int lane) {
    constexpr int CW = C > 0 ? C : 1;
    __syncwarp();
    const unsigned nr = min(T.qn[0], (unsigned)RQ), ne = min(T.qn[1], (unsigned)EQ);
    if (nr | ne) {
        if ((unsigned)lane < nr) {
            unsigned s[CW], q[CW];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                s[c] = T.qsum[c * RQ + lane];
                q[c] = T.qsq[c * RQ + lane];
            }
            table_region_add<C>(T, P, T.qlabel[lane], T.qarea[lane], T.qborder[lane], s, q);
        }
        for (unsigned k = lane; k < ne; k += 32) edge_add<C>(T, P, T.qekey[k], T.qecnt[k]);
        __syncwarp();
        if (lane == 0) T.qn[0] = T.qn[1] = 0;
        __syncwarp();
    }
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
    Acc<C> c0, c1;              // 2-entry label cache with per-label accumulators
    unsigned e01;               // pixel pairs seen between c0.label and c1.label (fast path)
    unsigned long long ekey;    // 1-entry run cache of the slow path
    unsigned ecnt;
#if defined(DM_RAG_STATS) && !defined(DM_RAG_TIMING_ONLY)
    int last_evicted = -7;
#endif

    __device__ __forceinline__ void init() {
        c0.reset(EMPTY_LABEL);
        c1.reset(EMPTY_LABEL);
        e01 = 0;
        ekey = EMPTY_KEY;
        ecnt = 0;
    }
    __device__ __forceinline__ void edge_flush(const Tables<C>& T, const Params& P) {
        if (ecnt) edge_push<C>(T, P, ekey, ecnt);
        ecnt = 0;
    }
    __device__ __forceinline__ void e01_flush(const Tables<C>& T, const Params& P) {
        if (e01) edge_push<C>(T, P, pack_key(c0.label, c1.label), e01);
        e01 = 0;
    }
    // replace cache entry `which` (0/1) by label l
    __device__ __forceinline__ void evict(const Tables<C>& T, const Params& P, int which, int l) {
#if defined(DM_RAG_STATS) && !defined(DM_RAG_TIMING_ONLY)
        last_evicted = which == 0 ? c0.label : c1.label;
#endif
        e01_flush(T, P);
        if (which == 0) {
            acc_push<C>(T, P, c0);
            c0.label = l;
        } else {
            acc_push<C>(T, P, c1);
            c1.label = l;
        }
    }
    __device__ __forceinline__ void flush_all(const Tables<C>& T, const Params& P) {
        e01_flush(T, P);
        acc_push<C>(T, P, c0);
        acc_push<C>(T, P, c1);
        edge_flush(T, P);
    }
    __device__ __forceinline__ void border_add(const Tables<C>& T, const Params& P, int v, unsigned n) {
        if (v == c0.label) c0.border += n;
        else if (v == c1.label) c1.border += n;
        else slow_region_add_acc<C>(T, P, v, 0, n, nullptr, nullptr);
    }
    // slow path: one pixel pair with different labels
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

// One lane-row (4 pixels) of the walk.  SPECIAL = the unit touches an image border (first / last
// strip, first / last rows): only then are pixels masked with `vm` and border sides counted.
template <int C, bool SPECIAL>
__device__ __forceinline__ void process_row(Thread<C>& th, int4& own, unsigned& m0, unsigned& m1, unsigned& use0,
                                            unsigned& use1, const int4 dn, const int right, const unsigned* TB,
                                            const Tables<C>& T, const Params& P, const unsigned vm, const bool left_edge,
                                            const int last_k, const unsigned edge_rows, const int nin) {
    const int4 a = own;
    const bool any_neg = (a.x | a.y | a.z | a.w | right | dn.x | dn.y | dn.z | dn.w) < 0;
    const unsigned negm = any_neg ? neg4(a) : 0u;
    // masks of the lower row / right neighbour against the two cached labels (d* become the next row's m*)
    unsigned d0 = match4(dn, th.c0.label), d1 = match4(dn, th.c1.label);
    unsigned r0 = right == th.c0.label ? 0xff000000u : 0u, r1 = right == th.c1.label ? 0xff000000u : 0u;
    unsigned xtra = 0;                 // own pixels accounted for without the cache (>2 labels in the lane)
    if (((m0 | m1 | negm) & (d0 | d1 | (any_neg ? 0xffffffffu : 0u))) != 0xffffffffu || (!any_neg && !(r0 | r1))) {
        // Some label around is not cached: load it.  One code instance serves the lane's own pixels,
        // the right neighbour and the row below (the region that starts there is then already cached
        // when the walk reaches it).  Entries (re)loaded in this row are pinned.
        unsigned pinned = 0;
#pragma unroll 1
        for (int it = 0; it < 6; ++it) {
            const unsigned c_own = m0 | m1 | negm | xtra;
            int cand, i4 = 0;
            bool for_own = false;
            if (c_own != 0xffffffffu) {
                i4 = (__ffs(~c_own) - 1) >> 3;
                cand = pick4(a, i4);
                for_own = true;
            } else if (any_neg) {
                break;                                    // nodata around: neighbours go through the slow path
            } else if (!(r0 | r1)) {
                cand = right;
            } else if ((d0 | d1) != 0xffffffffu) {
                cand = pick4(dn, (__ffs(~(d0 | d1)) - 1) >> 3);
            } else {
                break;
            }
            int which = -1;
            if (th.c0.label < 0 && !(pinned & 1u)) which = 0;          // an empty entry first
            else if (th.c1.label < 0 && !(pinned & 2u)) which = 1;
            else if (m0 == 0 && !(pinned & 1u)) which = 0;             // else one this row's pixels do not use
            else if (m1 == 0 && !(pinned & 2u)) which = 1;
            if (which < 0) {
                if (!for_own) break;                                   // three labels meet here: slow path below
                STAT(1);
                Acc<C> one;
                one.reset(cand);
                const unsigned bm = (0xffu << (8 * i4)) & vm;
                acc_pixels<C>(one, bm, TB);
                if (SPECIAL && bm)
                    one.border = (left_edge && i4 == 0 ? 1u : 0u) + (i4 == last_k ? 1u : 0u) + edge_rows;
                acc_push<C>(T, P, one);
                xtra |= 0xffu << (8 * i4);
                continue;
            }
            STAT(for_own ? 0 : 2);
            th.evict(T, P, which, cand);
            const unsigned mm = match4(a, cand), dd = match4(dn, cand), rr = right == cand ? 0xff000000u : 0u;
            if (which == 0) { m0 = mm; d0 = dd; r0 = rr; pinned |= 1u; }
            else            { m1 = mm; d1 = dd; r1 = rr; pinned |= 2u; }
        }
    }
    if (any_neg) {                                     // cache labels are >= 0 or EMPTY(-1): keep nodata out
        m0 &= ~negm;
        m1 &= ~negm;
    }
    // ---- accumulate the 4 pixels ---------------------------------------------------------------------
    {
        const unsigned v0 = SPECIAL ? (m0 & vm) : m0, v1 = SPECIAL ? (m1 & vm) : m1;
        acc_pixels<C>(th.c0, v0, TB);
        acc_pixels<C>(th.c1, v1, TB);
        if (SPECIAL) {
            if (left_edge) {
                th.c0.border += v0 & 1u;
                th.c1.border += v1 & 1u;
            }
            if ((unsigned)last_k < 4u) {
                th.c0.border += (v0 >> (8 * last_k)) & 1u;
                th.c1.border += (v1 >> (8 * last_k)) & 1u;
            }
            if (edge_rows) {                            // warp-uniform: first / last raster row
                th.c0.border += edge_rows * (__popc(v0) >> 3);
                th.c1.border += edge_rows * (__popc(v1) >> 3);
            }
        }
    }
    // ---- neighbour pairs -----------------------------------------------------------------------------
    use0 |= m0 | d0 | r0;
    use1 |= m1 | d1 | r1;
    if (!any_neg && ((m0 | m1) & (d0 | d1)) == 0xffffffffu && (r0 | r1)) {
        // every label around is one of the two cached ones: pairs that straddle them are counted
        // with byte-mask logic, no per-pair work and no divergence
        const unsigned n0 = (m0 >> 8) | r0, n1 = (m1 >> 8) | r1;
        const unsigned cross_h = (m0 & n1) | (m1 & n0);
        unsigned cross_v = (m0 & d1) | (m1 & d0);
        if (SPECIAL) cross_v &= vm;                     // padded copies right of the image do not pair
        th.e01 += (__popc(cross_h) + __popc(cross_v)) >> 3;
        STAT(3);
    } else {
        // a third label or nodata nearby: the 8 pairs one by one (bits 0-3 = (x,x+1), 4-7 = (y,y+1))
        unsigned pm = (a.x != a.y ? 1u : 0u) | (a.y != a.z ? 2u : 0u) | (a.z != a.w ? 4u : 0u) |
                      (a.w != right ? 8u : 0u) | (a.x != dn.x ? 16u : 0u) | (a.y != dn.y ? 32u : 0u) |
                      (a.z != dn.z ? 64u : 0u) | (a.w != dn.w ? 128u : 0u);
        if (SPECIAL) pm &= 0x0fu | (((1u << nin) - 1u) << 4);         // no vertical pairs right of the image
        STAT(4);
        while (pm) {
            const int b = __ffs(pm) - 1;
            pm &= pm - 1;
            const int k = b & 3;
            const int pa = pick4(a, k);
            const int pb = b < 4 ? (k == 0 ? a.y : k == 1 ? a.z : k == 2 ? a.w : right) : pick4(dn, k);
            th.pair(T, P, pa, pb, 1);
        }
    }
    own = dn;
    m0 = d0;
    m1 = d1;
    if (any_neg) {                                     // EMPTY(-1) may have matched nodata pixels of the lower row
        const unsigned nd = ~neg4(dn);
        m0 &= nd;
        m1 &= nd;
    }
}

// Drain the warp's tables to global memory (whole warp, convergent).
template <int C>
__device__ __forceinline__ void drain_tables(const Tables<C>& T, const Params& P, int lane) {
    constexpr int CW = C > 0 ? C : 1;
    static_assert(RS == 32 && ES == 64, "one / two slots per lane");
    {
        const int label = T.rkey[lane];
        if (label != EMPTY_LABEL) {
            unsigned s[CW], q[CW];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                s[c] = T.rsum[c * RS + lane];
                q[c] = T.rsq[c * RS + lane];
                T.rsum[c * RS + lane] = 0;
                T.rsq[c * RS + lane] = 0;
            }
            global_region_add<C>(P, label, T.rarea[lane], T.rborder[lane], s, q);
            T.rkey[lane] = EMPTY_LABEL;
            T.rarea[lane] = 0;
            T.rborder[lane] = 0;
        }
    }
#pragma unroll
    for (int k0 = 0; k0 < ES; k0 += 32) {
        const int k = k0 + lane;
        unsigned long long key = T.ekey[k];
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
    if (lane == 0) T.used[0] = T.used[1] = 0;
    __syncwarp();
}

// ------------------------------------------------------------------------------------ //
// the kernel
// ------------------------------------------------------------------------------------ //
template <typename CF, bool USE_TMA>
__global__ void __launch_bounds__(CF::THREADS, 1)
rag_pool_kernel(const __grid_constant__ CUtensorMap mapL, const __grid_constant__ CUtensorMap mapI, const Params P) {
    constexpr int C = CF::C;
    constexpr int TH = CF::TH, NS = CF::NS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* wbase = smem + (size_t)warp * CF::WARP_BYTES;       // this warp's private arena
    unsigned* tab = (unsigned*)(wbase + NS * CF::STAGE_BYTES);
    const Tables<C> T = Tables<C>::from(tab);
    uint64_t* full_bar = (uint64_t*)(tab + ((CF::TABLE_WORDS + 1) & ~1));

    // ---- this warp's run of units (unit = TH rows of one strip, column-major order) -------
    const long long total_units = (long long)P.tiles_x * P.tiles_y;
    const long long gw = (long long)blockIdx.x * NWARPS + warp;
    const long long u_begin = min(total_units, gw * (long long)P.tiles_per_cta);
    const long long u_end = min(total_units, u_begin + P.tiles_per_cta);
    const int my_units = (int)(u_end - u_begin);

    // ---- init (warp-private, no block barrier needed) ---------------------------------------
    T.rkey[lane] = EMPTY_LABEL;
    T.rarea[lane] = 0;
    T.rborder[lane] = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        T.rsum[c * RS + lane] = 0;
        T.rsq[c * RS + lane] = 0;
    }
    T.ekey[lane] = EMPTY_KEY;
    T.ekey[lane + 32] = EMPTY_KEY;
    T.ecnt[lane] = 0;
    T.ecnt[lane + 32] = 0;
    if (lane == 0) {
        T.used[0] = T.used[1] = 0;
        T.qn[0] = T.qn[1] = 0;
        if (USE_TMA) {
            for (int s = 0; s < NS; ++s) mbar_init(&full_bar[s], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    __syncwarp();
    if (my_units == 0) return;

    auto issue = [&](int k) {     // lane 0: TMA loads of this warp's k-th unit into stage k % NS
        const long long u = u_begin + k;
        const int sx = (int)(u / P.tiles_y), j = (int)(u - (long long)sx * P.tiles_y);
        const int st = k % NS;
        unsigned char* sb = wbase + (size_t)st * CF::STAGE_BYTES;
        mbar_expect_tx(&full_bar[st], (unsigned)CF::TX_BYTES);
        tma_load_2d(sb, &mapL, sx * STRIP_W, j * TH, &full_bar[st]);
        if (C > 0) tma_load_2d(sb + CF::LAB_BOX, &mapI, sx * STRIP_W * C / 4, j * TH, &full_bar[st]);
    };
    if (USE_TMA && lane == 0) {
        for (int k = 0; k < NS && k < my_units; ++k) issue(k);
    }

#ifdef DM_RAG_STATS
    unsigned long long t_start;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
#endif
    Thread<C> th;
    th.init();
    int units_since_drain = 0;
    int4 own = make_int4(0, 0, 0, 0);
    unsigned m0 = 0, m1 = 0;            // byte masks of `own` against the two cached labels
    unsigned use0 = 0, use1 = 0;        // cache entry touched during the current unit (see the GC below)
    // (strip, row block) of the current unit, advanced incrementally (column-major order)
    int sx = (int)(u_begin / P.tiles_y), j = (int)(u_begin - (long long)sx * P.tiles_y);
    bool contiguous = false;

    for (int i = 0; i < my_units; ++i) {
        const int st = USE_TMA ? i % NS : 0;
        unsigned char* sb = wbase + (size_t)st * CF::STAGE_BYTES;
        const int strip_x0 = sx * STRIP_W, unit_y0 = j * TH;
        int* Lw = (int*)sb;
        if (USE_TMA) {
            mbar_wait(&full_bar[st], (unsigned)(i / NS) & 1u, P.counters);
        } else {
            // fallback staging for rasters whose pitch/base TMA cannot describe
            __syncwarp();
            for (int k = lane; k < (TH + 1) * LAB_PITCH; k += 32) {
                const int r = k / LAB_PITCH, cidx = k - r * LAB_PITCH;
                const int gy = unit_y0 + r, gx = strip_x0 + cidx;
                Lw[k] = (gy < P.rows_avail && gx < P.W) ? P.labels[(int64_t)gy * P.ld + gx] : 0;
            }
            if constexpr (C > 0) {
                unsigned char* Ib = sb + CF::LAB_BOX;
                for (int k = lane; k < TH * CF::IMG_ROW_WORDS * 4; k += 32) {
                    const int r = k / (CF::IMG_ROW_WORDS * 4), bidx = k - r * (CF::IMG_ROW_WORDS * 4);
                    const int gy = unit_y0 + r;
                    const int64_t gb = (int64_t)strip_x0 * C + bidx;
                    Ib[k] = (gy < P.rows_own && gb < (int64_t)P.W * C) ? P.image[(int64_t)gy * P.image_pitch + gb] : 0;
                }
            }
            __syncwarp();
        }

        const int* L = Lw;
        const unsigned* I = (const unsigned*)(sb + CF::LAB_BOX);
        const int x0 = strip_x0 + 4 * lane;                 // first of this lane's 4 pixels
        // a unit is "special" when it touches an image border: only those pay for border logic
        const bool special = (sx == 0) || (sx == P.tiles_x - 1) || (unit_y0 == 0) || (unit_y0 + TH >= P.rows_own);
        if (!special) {
            if (!contiguous) {
                own = *(const int4*)(L + 4 * lane);
                m0 = match4(own, th.c0.label);
                m1 = match4(own, th.c1.label);
            }
#pragma unroll 1
            for (int r = 0; r < TH; ++r) {
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
                process_row<C, false>(th, own, m0, m1, use0, use1, dn, right, TB, T, P, 0xffffffffu, false, -1, 0u, 4);
            }
        } else {
            // Image borders without a separate per-pixel path: pixels right of the image are replaced
            // by copies of the row's last pixel (so they never differ from a neighbour) and masked out
            // of the accumulation with `vm`; border sides are added arithmetically.
            const int nin = min(4, max(0, P.W - x0));       // pixels of this lane inside the image
            const unsigned vm = nin >= 4 ? 0xffffffffu : ((1u << (8 * nin)) - 1u);
            const bool left_edge = (x0 == 0);
            const int last_k = P.W - 1 - x0;                // in [0,3] for the lane holding the last column
            const int last_col = min(P.W - 1 - strip_x0, STRIP_W - 1);
            if (!contiguous) {
                own = *(const int4*)(L + 4 * lane);
                if (nin < 4) {
                    const int e = L[last_col];
                    if (nin < 1) own.x = e;
                    if (nin < 2) own.y = e;
                    if (nin < 3) own.z = e;
                    own.w = e;
                }
                m0 = match4(own, th.c0.label);
                m1 = match4(own, th.c1.label);
                const unsigned nd = ~neg4(own);
                m0 &= nd;
                m1 &= nd;
            }
#pragma unroll 1
            for (int r = 0; r < TH; ++r) {
                const int y = unit_y0 + r;
                if (y >= P.rows_own) break;
                const bool has_dn = (y + 1 < P.rows_avail);     // warp-uniform
                int4 dn = *(const int4*)(L + (r + 1) * LAB_PITCH + 4 * lane);
                int right = __shfl_down_sync(0xffffffffu, own.x, 1);
                if (lane == 31) right = L[r * LAB_PITCH + STRIP_W];
                if (nin < 4) {                                  // only lanes of the last strip
                    const int e = L[(r + 1) * LAB_PITCH + last_col];
                    if (nin < 1) dn.x = e;
                    if (nin < 2) dn.y = e;
                    if (nin < 3) dn.z = e;
                    dn.w = e;
                }
                if (!has_dn) dn = own;                          // last raster row: no vertical pairs
                if (x0 + 4 >= P.W) right = own.w;               // nothing to the right of the last column
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
                const unsigned edge_rows = ((y == 0 && P.top_border) ? 1u : 0u) +
                                           ((!has_dn && y == P.rows_own - 1 && P.bottom_border) ? 1u : 0u);
                process_row<C, true>(th, own, m0, m1, use0, use1, dn, right, TB, T, P, vm, left_edge, last_k, edge_rows, nin);
            }
        }

        // ---- convergent garbage collection of the label cache --------------------------------------
        // An entry nothing touched during this unit is pushed out now, by all lanes together, so
        // that the next miss of such a lane finds an empty entry instead of evicting on its own.
        if ((!use0 && th.c0.label >= 0) || (!use1 && th.c1.label >= 0)) {
            th.e01_flush(T, P);
            if (!use0 && th.c0.label >= 0) {
                acc_push<C>(T, P, th.c0);
                th.c0.label = EMPTY_LABEL;
            }
            if (!use1 && th.c1.label >= 0) {
                acc_push<C>(T, P, th.c1);
                th.c1.label = EMPTY_LABEL;
            }
        }
        use0 = use1 = 0;
        // next unit of this warp's run
        contiguous = (j + 1 < P.tiles_y);
        if (contiguous) ++j;
        else { j = 0; ++sx; }

        // ---- recycle the stage: this warp is its only reader, so it refills it itself ----------
        __syncwarp();
        if (USE_TMA && lane == 0 && i + NS < my_units) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads before async writes
            issue(i + NS);
        }

        // ---- queued evictions -> tables (whole warp, every unit); tables -> global when they fill up ----
        ++units_since_drain;
        const bool last = (i + 1 == my_units);
        const bool forced = units_since_drain >= CF::FLUSH_UNITS || last;
        if (forced) th.flush_all(T, P);
        drain_queues<C>(T, P, lane);
        const unsigned ur = T.used[0], ue = T.used[1];
        if (ur > RS / 2 || ue > ES / 2 || forced) {
            if (lane == 0) STAT(6);
            drain_tables<C>(T, P, lane);
            units_since_drain = 0;
        }
    }
#ifdef DM_RAG_STATS
    if (lane == 0) {       // per-warp wall time (ns): counts[] must hold 32 + 2 * total_warps entries
        unsigned long long t_end;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
        P.counters[32 + 2 * gw] = t_start;
        P.counters[32 + 2 * gw + 1] = t_end;
    }
#endif
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

static thread_local int g_last_path = -1;     // 1 = TMA staging, 0 = ld.global staging
static thread_local int g_last_encode_error = 0;
int last_path() { return g_last_path; }
int last_encode_error() { return g_last_encode_error; }

static bool make_map_2d(CUtensorMap* m, const void* base, uint64_t dim0, uint64_t dim1, uint64_t pitch_bytes,
                        uint32_t box0, uint32_t box1) {
    EncodeTiledFn enc = get_encode();
    if (!enc) {
        g_last_encode_error = -1;
        return false;
    }
    cuuint64_t dims[2] = {dim0, dim1};
    cuuint64_t strides[1] = {pitch_bytes};
    cuuint32_t box[2] = {box0, box1};
    cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    g_last_encode_error = (int)r;
    return r == CUDA_SUCCESS;
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

template <typename CF>
static int launch(const Params& Pin, bool allow_tma, cudaStream_t s) {
    Params P = Pin;
    P.tiles_x = (int)ceil_div(P.W, STRIP_W);            // strips
    P.tiles_y = (int)ceil_div(P.rows_own, CF::TH);      // row blocks per strip
    const long long total = (long long)P.tiles_x * P.tiles_y;
    if (total == 0) return DM_OK;
    // one persistent 16-warp CTA per SM; every warp takes one contiguous run of units
    const long long max_warps = (long long)num_sms() * NWARPS;
    const long long per = ceil_div(total, max_warps);
    if (per > 0x7fffffff) return DM_ERR_BAD_ARG;
    P.tiles_per_cta = (int)per;
    const int grid = (int)ceil_div(ceil_div(total, per), NWARPS);

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
    g_last_path = tma ? 1 : 0;
    if (tma) {
        auto k = rag_pool_kernel<CF, true>;
        DM_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::SMEM_BYTES));
        DM_COUNT_LAUNCH(); k<<<grid, CF::THREADS, CF::SMEM_BYTES, s>>>(mapL, mapI, P);
    } else {
        auto k = rag_pool_kernel<CF, false>;
        DM_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::SMEM_BYTES));
        DM_COUNT_LAUNCH(); k<<<grid, CF::THREADS, CF::SMEM_BYTES, s>>>(mapL, mapI, P);
    }
    DM_LAUNCH_CHECK();
    return DM_OK;
}

// DM_RAG_NO_TMA=1 forces the ld.global staging path (tests exercise both).
static bool allow_tma_env() {
    const char* e = getenv("DM_RAG_NO_TMA");
    return !(e && e[0] == '1');
}

int run(const Params& P, int C, cudaStream_t s) {
    const bool tma = allow_tma_env();
    switch (C) {
        case 0: return launch<Cfg<0, 8, 2>>(P, tma, s);
        case 1: return launch<Cfg<1, 4, 3>>(P, tma, s);
        case 2: return launch<Cfg<2, 4, 2>>(P, tma, s);
        case 3: return launch<Cfg<3, 4, 2>>(P, tma, s);
        case 4: return launch<Cfg<4, 4, 2>>(P, tma, s);
        default: return DM_ERR_BAD_ARG;
    }
}

}  // namespace rag
}  // namespace dm

using namespace dm;

extern "C" int dm_rag_last_path(void) { return rag::last_path(); }
extern "C" int dm_rag_last_encode_error(void) { return rag::last_encode_error(); }

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
