// R1: region adjacency graph from a label raster, fused with per-region band pooling.
//
// Warp-autonomous streaming design.  The raster is cut into 128-pixel-wide strips; every
// warp of a persistent grid (one 16-warp CTA per SM) owns a contiguous run of rows of one
// strip and walks straight down it:
//   * lane 0 of the warp feeds the warp's OWN ring of shared-memory stages with TMA
//     (cp.async.bulk.tensor boxes of TH label rows x 132 columns -- a 4-column halo for the
//     right neighbour -- and TH image rows), completion signalled on the warp's own
//     mbarriers.  No producer warp, no block-level barrier anywhere after start-up: warps
//     never wait for each other.  The row ABOVE a unit is carried in registers from the
//     previous unit (read from global memory only at the start of a run / a strip).
//   * every lane owns 4 consecutive pixels (one 128-bit LDS of labels, one of image bytes)
//     and keeps a 2-entry register cache of per-label accumulators: area, border sides, C
//     band sums and sums of squares, fed 4 pixels at a time with PRMT + DP4A on byte masks.
//     Because the warp moves down contiguous rows, a cache entry lives for the whole height
//     of a region.  Vertical pairs are (row above, own row), so every label involved is
//     already cached: two pixels of one entry are no pair, two pixels of different entries
//     are a pair of the key (c0, c1), counted by popcount -- no label comparison, no
//     divergence.  Junctions of three regions and nodata take a generic per-pair path.
//   * evictions are pushed to a per-warp shared-memory queue, drained convergently into the
//     warp's private hash tables (label -> accumulators, edge key -> pair count), which are
//     drained to global memory with 64-bit atomics / appended (key,count) entries when half
//     full.  The appended entries are then radix sorted and run-reduced (prims.cu) into the
//     sorted unique edge list.
// Things that matter for speed here (measured): shared-memory pointers must
// keep their address space (no integer round trips: generic LD/ST/ATOM are far slower), runtime
// picks must be selp chains (?: compiles to divergent branches), and every rare divergent block is
// followed by __syncwarp() (otherwise the rest of the row runs once per divergent group).
//
// HBM traffic: labels 4 B/px + image C B/px read once.
#include "rag_common.cuh"
#include "prims.cuh"

namespace dm {
namespace rag {

constexpr int RQ = 32;                  // region eviction queue entries per warp
constexpr int EQ = 64;                  // edge eviction queue entries per warp

template <int C_, int TH_, int NS_, int NW_>
struct Cfg {
    static constexpr int C = C_, TH = TH_, NS = NS_;
    static constexpr int NWARPS = NW_;                       // warps per CTA, each an independent pipeline
    static constexpr int CW = C_ > 0 ? C_ : 1;               // words of image bytes per lane-row
    static constexpr int THREADS = NWARPS * 32;
    static constexpr int LAB_BOX = align128(TH * LAB_PITCH * 4);
    static constexpr int IMG_ROW_WORDS = STRIP_W * C / 4;    // 32*C
    static constexpr int IMG_BOX = align128(TH * IMG_ROW_WORDS * 4);
    static constexpr int STAGE_BYTES = LAB_BOX + IMG_BOX;
    static constexpr int QUEUE_WORDS = RQ * (3 + 2 * C) + EQ * 3 + 2 + 2;   // eviction queues + counts[2] + pad
    static constexpr int TABLE_WORDS = RS * (3 + 2 * C) + ES * 3 + 2 + 2 + QUEUE_WORDS;   // + used[2] + pad
    static constexpr int TABLE_BYTES = align128(TABLE_WORDS * 4 + NS * 8);   // + full barriers
    static constexpr int WARP_BYTES = NS * STAGE_BYTES + TABLE_BYTES;
    static constexpr int SMEM_BYTES = 128 + NWARPS * WARP_BYTES;
    static constexpr int TX_BYTES = TH * LAB_PITCH * 4 + (C > 0 ? TH * IMG_ROW_WORDS * 4 : 0);
    static constexpr int FLUSH_UNITS = FLUSH_ROWS / TH > 0 ? FLUSH_ROWS / TH : 1;
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

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
// branch-free select (the compiler turns ?: chains on a runtime index into divergent branches)
__device__ __forceinline__ int selp(int a, int b, bool p) {
    int r;
    asm("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %3, 0;\n\tselp.s32 %0, %1, %2, q;\n\t}" : "=r"(r) : "r"(a), "r"(b), "r"((int)p));
    return r;
}
__device__ __forceinline__ int pick4(const int4& a, int i) {
    return selp(selp(a.x, a.y, i == 0), selp(a.z, a.w, i == 2), i < 2);
}

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
    unsigned e01;               // pixel pairs seen between c0.label and c1.label (convergent fast path)
    int4 up;                    // the 4 labels of the row above (vertical pairs are (up, own))
    unsigned u0, u1;            // byte masks of `up` against the two cached labels
    int rp, rq;                 // run of the pair that crosses to the next lane: (a.w, right) = (rp, rq)
    unsigned rcnt;
    unsigned long long ekey;    // 1-entry run cache of the generic pair path
    unsigned ecnt;

    __device__ __forceinline__ void init() {
        c0.reset(EMPTY_LABEL);
        c1.reset(EMPTY_LABEL);
        e01 = 0;
        up = make_int4(0, 0, 0, 0);
        u0 = u1 = 0;
        rp = rq = EMPTY_LABEL;
        rcnt = 0;
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
    __device__ __forceinline__ void r_flush(const Tables<C>& T, const Params& P) {
        if (rcnt) edge_push<C>(T, P, pack_key(rp, rq), rcnt);
        rcnt = 0;
    }
    // replace cache entry `which` (0/1) by label l
    __device__ __forceinline__ void evict(const Tables<C>& T, const Params& P, int which, int l) {
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
        r_flush(T, P);
        acc_push<C>(T, P, c0);
        acc_push<C>(T, P, c1);
        edge_flush(T, P);
    }
    __device__ __forceinline__ void border_add(const Tables<C>& T, const Params& P, int v, unsigned n) {
        if (v == c0.label) c0.border += n;
        else if (v == c1.label) c1.border += n;
        else slow_region_add_acc<C>(T, P, v, 0, n, nullptr, nullptr);
    }
    // generic path: n pixel pairs between labels a and b (a != b)
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
    // (re)load the row above (first unit of a run / of a strip)
    __device__ __forceinline__ void set_up(const int4& row) {
        up = row;
        const unsigned nd = ~neg4(row);
        u0 = match4(row, c0.label) & nd;
        u1 = match4(row, c1.label) & nd;
    }
};

// image-border sides of the pixels in byte mask v (border units only)
__device__ __forceinline__ unsigned border_sides(unsigned v, bool left_edge, int last_k, unsigned edge_rows) {
    unsigned n = edge_rows * (__popc(v) >> 3);
    if (left_edge) n += v & 1u;
    if ((unsigned)last_k < 4u) n += (v >> (8 * last_k)) & 1u;
    return n;
}

__device__ __forceinline__ unsigned ne4(const int4& a, int b0, int b1, int b2, int b3) {
    unsigned m = (a.x != b0) ? 0x000000ffu : 0u;
    m |= (a.y != b1) ? 0x0000ff00u : 0u;
    m |= (a.z != b2) ? 0x00ff0000u : 0u;
    m |= (a.w != b3) ? 0xff000000u : 0u;
    return m;
}

// One lane-row (4 pixels) of the walk: `a` = the lane's 4 labels, th.up = the 4 labels above them,
// `right` = the label right of a.w.  The lane owns the 4 pixels' statistics, the 4 horizontal pairs
// (a.x,a.y) (a.y,a.z) (a.z,a.w) (a.w,right) and the 4 vertical pairs (up.k, a.k).
//   * In the common case the two cached labels cover every label of `a` and `up` (a region interior,
//     or the boundary between two regions running through the lane).  Then everything is byte-mask
//     arithmetic without divergence: masked DP4A accumulation of the 4 pixels into the two entries,
//     and -- no label comparison needed -- two pixels of one entry are no pair, two pixels of
//     different entries are a pair of the key (c0, c1), counted by popcount into e01.
//   * A label of `a` that is not cached replaces the entry neither this row nor the row above uses
//     (its accumulators go to the warp's eviction queue).  A third label inside the 4 pixels is
//     pushed directly.
//   * Pairs with a pixel outside the cache (junctions of three regions, nodata) are compared label
//     by label and go through the generic path.  The pair that crosses to the next lane is a run
//     of its own (rp, rq, rcnt): a region boundary that falls between two lanes costs one compare
//     per row.
// SPECIAL = the unit touches an image border: statistics are masked with `vs`, vertical pairs with
// `vp`, and border sides are counted; no_h = a halo row below the tile (vertical pairs only).
template <int C, bool SPECIAL>
__device__ __forceinline__ void process_row(Thread<C>& th, const int4 a, const int right, const unsigned* TB,
                                            const Tables<C>& T, const Params& P, const unsigned vs, const unsigned vp,
                                            const bool left_edge, const int last_k, const unsigned edge_rows,
                                            const bool no_h) {
    const int4 up = th.up;
    unsigned m0 = match4(a, th.c0.label), m1 = match4(a, th.c1.label);
    unsigned u0 = th.u0, u1 = th.u1;
    unsigned cov = m0 | m1;
    if ((a.x | a.y | a.z | a.w) < 0) {                  // nodata among the own pixels (EMPTY is negative too)
        const unsigned negm = neg4(a);
        m0 &= ~negm;
        m1 &= ~negm;
        cov = m0 | m1 | negm;
    }
    // ---- labels that are not cached -----------------------------------------------------------------------
    if (cov != 0xffffffffu) {
#pragma unroll 1
        do {
            const int cand = pick4(a, (__ffs(~cov) - 1) >> 3);
            const unsigned mm = match4(a, cand);
            cov |= mm;
            // victim: an entry neither this row nor the row above uses, else one this row does not use
            const int which = (m0 | u0) == 0 ? 0 : (m1 | u1) == 0 ? 1 : m0 == 0 ? 0 : m1 == 0 ? 1 : -1;
            if (which < 0) {                                         // a third label inside the 4 pixels
                Acc<C> one;
                one.reset(cand);
                const unsigned v = SPECIAL ? (mm & vs) : mm;
                acc_pixels<C>(one, v, TB);
                if (SPECIAL) one.border = border_sides(v, left_edge, last_k, edge_rows);
                acc_push<C>(T, P, one);
            } else {
                th.evict(T, P, which, cand);
                const unsigned uu = match4(up, cand);
                if (which == 0) { m0 = mm; u0 = uu; }
                else            { m1 = mm; u1 = uu; }
            }
        } while (cov != 0xffffffffu);
    }
    __syncwarp();                                       // reconverge before the common part
    // ---- statistics ---------------------------------------------------------------------------------------
    {
        const unsigned v0 = SPECIAL ? (m0 & vs) : m0, v1 = SPECIAL ? (m1 & vs) : m1;
        acc_pixels<C>(th.c0, v0, TB);
        acc_pixels<C>(th.c1, v1, TB);
        if (SPECIAL) {
            th.c0.border += border_sides(v0, left_edge, last_k, edge_rows);
            th.c1.border += border_sides(v1, left_edge, last_k, edge_rows);
        }
    }
    // ---- neighbour pairs ----------------------------------------------------------------------------------
    const unsigned acov = m0 | m1, ucov = u0 | u1;
    unsigned cv = (u0 & m1) | (u1 & m0);                             // vertical pairs (up.k, a.k) across the two labels
    unsigned ch = (m0 & (m1 >> 8)) | (m1 & (m0 >> 8));               // horizontal pairs (k, k+1), k = 0..2
    unsigned lv = ~(acov & ucov);                                    // pairs with a pixel outside the cache
    unsigned lh = ~(acov & (acov >> 8)) & 0x00ffffffu;
    bool rdiff = a.w != right;
    if (SPECIAL) {
        cv &= vp;
        lv &= vp;
        if (no_h) {
            ch = lh = 0;
            rdiff = false;
        }
    }
    th.e01 += (__popc(ch) + __popc(cv)) >> 3;
    const bool rsame = rdiff && a.w == th.rp && right == th.rq;
    th.rcnt += rsame ? 1u : 0u;
    const bool rnew = rdiff && !rsame;
    if (lv | lh | (rnew ? 1u : 0u)) {
        if (lv) {
            lv &= ne4(a, up.x, up.y, up.z, up.w);
#pragma unroll 1
            while (lv) {
                const int k = (__ffs(lv) - 1) >> 3;
                const int pa = pick4(up, k), pb = pick4(a, k);
                const unsigned same = match4(up, pa) & match4(a, pb) & lv;
                lv &= ~same;
                th.pair(T, P, pa, pb, __popc(same) >> 3);
            }
        }
        if (lh) {
            lh &= ne4(a, a.y, a.z, a.w, a.w);
#pragma unroll 1
            while (lh) {
                const int k = (__ffs(lh) - 1) >> 3;
                lh &= ~(0xffu << (8 * k));
                th.pair(T, P, pick4(a, k), selp(selp(a.y, a.z, k == 0), a.w, k < 2), 1);
            }
        }
        if (rnew) {
            th.r_flush(T, P);
            if ((a.w | right) >= 0) {
                th.rp = a.w;
                th.rq = right;
                th.rcnt = 1;
            } else {
                th.pair(T, P, a.w, right, 1);
            }
        }
    }
    // (no __syncwarp here: only register moves follow, and the next row's full-mask shuffle reconverges the warp)
    th.up = a;
    th.u0 = m0;
    th.u1 = m1;
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
    // align by OFFSET (no integer round trip) so that every derived pointer keeps the shared address space
    unsigned char* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* wbase = smem + (size_t)warp * CF::WARP_BYTES;       // this warp's private arena
    unsigned* tab = (unsigned*)(wbase + NS * CF::STAGE_BYTES);
    const Tables<C> T = Tables<C>::from(tab);
    uint64_t* full_bar = (uint64_t*)(tab + ((CF::TABLE_WORDS + 1) & ~1));

    // ---- this warp's run of units (unit = TH rows of one strip, column-major order) -------
    const long long total_units = (long long)P.tiles_x * P.tiles_y;
    const long long gw = (long long)blockIdx.x * CF::NWARPS + warp;
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

    // (strip, row block) of the next unit to fetch, advanced incrementally (column-major order)
    int isx = (int)(u_begin / P.tiles_y), ij = (int)(u_begin - (long long)isx * P.tiles_y), issued = 0;
    auto issue = [&]() {          // TMA loads of this warp's next unit into stage issued % NS (lane 0 issues)
        if (lane == 0) {
            const int st = issued % NS;
            unsigned char* sb = wbase + (size_t)st * CF::STAGE_BYTES;
            mbar_expect_tx(&full_bar[st], (unsigned)CF::TX_BYTES);
            tma_load_2d(sb, &mapL, isx * STRIP_W, ij * TH, &full_bar[st]);
            if (C > 0) tma_load_2d(sb + CF::LAB_BOX, &mapI, isx * STRIP_W * C / 4, ij * TH, &full_bar[st]);
        }
        ++issued;
        if (ij + 1 < P.tiles_y) ++ij;
        else { ij = 0; ++isx; }
    };
    if (USE_TMA) {
        for (int k = 0; k < NS && k < my_units; ++k) issue();
    }

    Thread<C> th;
    th.init();
    int units_since_flush = 0;
    // (strip, row block) of the current unit, advanced incrementally (column-major order)
    int sx = (int)(u_begin / P.tiles_y), j = (int)(u_begin - (long long)sx * P.tiles_y);
    bool contiguous = false;            // th.up / u0 / u1 carry over from the previous unit

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
            for (int k = lane; k < TH * LAB_PITCH; k += 32) {
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
        if (!contiguous) {
            // first unit of the run / of a strip: the row above comes straight from global memory
            // (pixels right of the image are copies of the row's last pixel, as below)
            int4 v = make_int4(0, 0, 0, 0);
            if (unit_y0 > 0) {
                const int32_t* row = P.labels + (int64_t)(unit_y0 - 1) * P.ld;
                v.x = row[min(x0, P.W - 1)];
                v.y = row[min(x0 + 1, P.W - 1)];
                v.z = row[min(x0 + 2, P.W - 1)];
                v.w = row[min(x0 + 3, P.W - 1)];
            }
            th.set_up(v);
        }
        // a unit is "special" when it touches an image border: only those pay for border logic
        const bool special = (sx == 0) || (sx == P.tiles_x - 1) || (unit_y0 == 0) || (unit_y0 + TH >= P.rows_own);
        if (!special) {
            const int* Lp = L + 4 * lane;                       // running row pointers: no per-row index arithmetic
            const unsigned* Ip = I + CF::CW * lane;
            const int* Lend = Lp + TH * LAB_PITCH;
#pragma unroll 1
            for (; Lp != Lend; Lp += LAB_PITCH, Ip += CF::IMG_ROW_WORDS) {
                const int4 a = *(const int4*)Lp;
                int right = __shfl_down_sync(0xffffffffu, a.x, 1);
                if (lane == 31) right = Lp[4];                  // the strip's 4-column halo starts right after lane 31's pixels
                unsigned W[CF::CW], TB[CF::CW];
                if (C > 0) {
                    if (C == 4) {
                        const uint4 v = *(const uint4*)Ip;
                        W[0] = v.x; W[1 % CF::CW] = v.y; W[2 % CF::CW] = v.z; W[3 % CF::CW] = v.w;
                    } else {
#pragma unroll
                        for (int c = 0; c < C; ++c) W[c] = Ip[c];
                    }
                    band_transpose<C>(W, TB);
                }
                process_row<C, false>(th, a, right, TB, T, P, 0xffffffffu, 0xffffffffu, false, -1, 0u, false);
            }
        } else {
            // Image borders without a separate per-pixel path: pixels right of the image are replaced
            // by copies of the row's last pixel (so they never differ from a neighbour) and masked out
            // of the accumulation / the vertical pairs with `vm`; border sides are added arithmetically.
            // A halo row below the owned rows (row-tile sharding) only contributes its vertical pairs.
            const int nin = min(4, max(0, P.W - x0));       // pixels of this lane inside the image
            const unsigned vm = nin >= 4 ? 0xffffffffu : ((1u << (8 * nin)) - 1u);
            const bool left_edge = (x0 == 0);
            const int last_k = P.W - 1 - x0;                // in [0,3] for the lane holding the last column
            const int last_col = min(P.W - 1 - strip_x0, STRIP_W - 1);
#pragma unroll 1
            for (int r = 0; r < TH; ++r) {
                const int y = unit_y0 + r;
                if (y >= P.rows_avail) break;
                const bool halo = (y >= P.rows_own);            // warp-uniform
                int4 a = *(const int4*)(L + r * LAB_PITCH + 4 * lane);
                if (nin < 4) {                                  // only lanes of the last strip
                    const int e = L[r * LAB_PITCH + last_col];
                    if (nin < 1) a.x = e;
                    if (nin < 2) a.y = e;
                    if (nin < 3) a.z = e;
                    a.w = e;
                }
                int right = __shfl_down_sync(0xffffffffu, a.x, 1);
                if (lane == 31) right = L[r * LAB_PITCH + STRIP_W];
                if (x0 + 4 >= P.W) right = a.w;                 // nothing to the right of the last column
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
                                           ((y == P.rows_own - 1 && P.rows_avail == P.rows_own && P.bottom_border) ? 1u : 0u);
                process_row<C, true>(th, a, right, TB, T, P, halo ? 0u : vm, y == 0 ? 0u : vm, left_edge, last_k,
                                     halo ? 0u : edge_rows, halo);
            }
        }
        contiguous = (j + 1 < P.tiles_y);

        // next unit of this warp's run
        if (j + 1 < P.tiles_y) ++j;
        else { j = 0; ++sx; }

        // ---- recycle the stage: this warp is its only reader, so it refills it itself ----------
        __syncwarp();
        if (USE_TMA && i + NS < my_units) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads before async writes
            issue();
        }

        // ---- queued evictions -> tables when the queues fill up; tables -> global when they fill up ----
        ++units_since_flush;
        const bool forced = units_since_flush >= CF::FLUSH_UNITS || (i + 1 == my_units);
        if (forced) {
            th.flush_all(T, P);
            units_since_flush = 0;
        }
        __syncwarp();
        if (forced || T.qn[0] > RQ / 2 || T.qn[1] > EQ / 2) {
            drain_queues<C>(T, P, lane);
            if (forced || T.used[0] > RS / 2 || T.used[1] > ES / 2) drain_tables<C>(T, P, lane);
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

static thread_local int g_last_path = -1;     // 1 = TMA staging, 0 = ld.global staging
static thread_local int g_last_encode_error = 0;
int last_path() { return g_last_path; }
void set_last_path(int path) { g_last_path = path; }
int last_encode_error() { return g_last_encode_error; }

bool make_map_2d(CUtensorMap* m, const void* base, uint64_t dim0, uint64_t dim1, uint64_t pitch_bytes,
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

template <typename CF>
static int launch(const Params& Pin, bool allow_tma, cudaStream_t s) {
    Params P = Pin;
    P.tiles_x = (int)ceil_div(P.W, STRIP_W);            // strips
    P.tiles_y = (int)ceil_div(P.rows_avail, CF::TH);    // row blocks per strip (a halo row below counts)
    const long long total = (long long)P.tiles_x * P.tiles_y;
    if (total == 0) return DM_OK;
    // one persistent 16-warp CTA per SM; every warp takes one contiguous run of units
    const long long max_warps = (long long)num_sms() * CF::NWARPS;
    const long long per = ceil_div(total, max_warps);
    if (per > 0x7fffffff) return DM_ERR_BAD_ARG;
    P.tiles_per_cta = (int)per;
    const int grid = (int)ceil_div(ceil_div(total, per), CF::NWARPS);

    CUtensorMap mapL, mapI;
    memset(&mapL, 0, sizeof(mapL));
    memset(&mapI, 0, sizeof(mapI));
    // TMA needs 16-byte aligned bases and row pitches; anything else takes the ld.global staging path
    bool tma = allow_tma && ((uintptr_t)P.labels % 16 == 0) && ((P.ld * 4) % 16 == 0) && P.ld >= P.W;
    if (CF::C > 0)
        tma = tma && ((uintptr_t)P.image % 16 == 0) && (P.image_pitch % 16 == 0) && (((int64_t)P.W * CF::C) % 4 == 0);
    if (tma)
        tma = make_map_2d(&mapL, P.labels, (uint64_t)P.W, (uint64_t)P.rows_avail, (uint64_t)P.ld * 4, LAB_PITCH, CF::TH);
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

// Kernel shapes (rows per unit, pipeline stages, warps per CTA) per band count.  Measured alternatives for C = 4 on a
// B200 (10k x 10k, ~1000-pixel regions): <4,4,2,16> 0.39 ms; <4,2,2,20> 0.42; <4,4,1,20> 0.38-0.43; <4,8,1,16> 0.38-0.45;
// <4,2,3,16> 0.43; 24-28 warps spill (0.43-0.54 ms).  See DESIGN.md section 6.
int run(const Params& P, int C, cudaStream_t s) {
    const bool tma = allow_tma_env();
    const char* e = getenv("DM_RAG_KERNEL");            // "v1": the label-cache kernel of round 1 (A/B measurements)
    if (e && e[0] == 's') return run_split(P, C, tma, s);    // "split": the warp-specialised kernel
    if (!(e && e[0] == 'v' && e[1] == '1')) return run_blocks(P, C, tma, s);
    switch (C) {
        case 0: return launch<Cfg<0, 8, 2, 16>>(P, tma, s);
        case 1: return launch<Cfg<1, 4, 3, 16>>(P, tma, s);
        case 2: return launch<Cfg<2, 4, 2, 16>>(P, tma, s);
        case 3: return launch<Cfg<3, 4, 2, 16>>(P, tma, s);
        case 4: return launch<Cfg<4, 4, 2, 16>>(P, tma, s);
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
    const int b = bits_for(n_regions);
    const int64_t* n_raw = counts + 1;                      // the fused kernel clamps to capacity itself
    if (!prims::sort_fused_available()) {
        // n_raw = min(raw entries, capacity): what actually sits in the raw list
        DM_COUNT_LAUNCH(); rag::clamp_count_kernel<<<1, 1, 0, s>>>(counts + 1, capacity, w.n_raw);
        n_raw = w.n_raw;
    }
    DM_TRY(prims::sort_unique(w.raw_keys, w.raw_cnt, n_raw, capacity, b, 2 * b, w.sws, nullptr, nullptr, ~0ull, edge_keys,
                              boundary_len, nullptr, counts, w.uws, nullptr, nullptr, nullptr, nullptr, s));
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

namespace dm {
namespace rag {
__global__ void concat_kernel(const uint64_t* __restrict__ sk, const uint32_t* __restrict__ sl,
                              const int64_t* __restrict__ counts, int n_lists, int64_t slot_cap,
                              uint64_t* __restrict__ dk, uint32_t* __restrict__ dl, int64_t dst_cap,
                              int64_t* __restrict__ n_out) {
    // every thread recomputes the (few) slot offsets; lists are short compared with the grid-stride work
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (int64_t)n_lists * slot_cap;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int g = (int)(i / slot_cap);
        const int64_t j = i - (int64_t)g * slot_cap;
        if (j >= counts[g]) continue;
        int64_t off = 0;
        for (int h = 0; h < g; ++h) off += imin64(counts[h], slot_cap);
        if (off + j < dst_cap) {
            dk[off + j] = sk[i];
            dl[off + j] = sl[i];
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int64_t tot = 0, bad = 0;
        for (int h = 0; h < n_lists; ++h) {
            if (counts[h] > slot_cap) bad = 1;
            tot += imin64(counts[h], slot_cap);
        }
        if (tot > dst_cap) { bad = 1; tot = dst_cap; }
        n_out[0] = tot;
        n_out[1] = bad;
    }
}
}  // namespace rag
}  // namespace dm

extern "C" int dm_edges_concat(const uint64_t* src_keys, const uint32_t* src_lens, const int64_t* counts, int64_t n_lists,
                               int64_t slot_capacity, uint64_t* dst_keys, uint32_t* dst_lens, int64_t dst_capacity,
                               int64_t* n_out_dev, dm_stream_t stream) {
    if (n_lists < 0 || n_lists > 4096 || slot_capacity < 0 || dst_capacity < 0 || !n_out_dev) return DM_ERR_BAD_ARG;
    cudaStream_t s = S(stream);
    if (n_lists == 0 || slot_capacity == 0) {
        DM_CUDA(cudaMemsetAsync(n_out_dev, 0, 2 * sizeof(int64_t), s));
        return DM_OK;
    }
    if (!src_keys || !src_lens || !counts || !dst_keys || !dst_lens) return DM_ERR_BAD_ARG;
    const unsigned g = (unsigned)imax64(1, imin64(ceil_div(n_lists * slot_capacity, 256), (int64_t)num_sms() * 8));
    DM_COUNT_LAUNCH(); rag::concat_kernel<<<g, 256, 0, s>>>(src_keys, src_lens, counts, (int)n_lists, slot_capacity, dst_keys,
                                                         dst_lens, dst_capacity, n_out_dev);
    DM_LAUNCH_CHECK();
    return DM_OK;
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
    DM_TRY(prims::sort_unique(keys, lens, n_in, capacity, b, 2 * b, sws, nullptr, nullptr, ~0ull, ok, ol, nullptr, n_new, uws,
                              keys, lens, nullptr, n_out, s));
    DM_LAUNCH_CHECK();
    return DM_OK;
}
