// Per-warp shared-memory hash tables of the raster kernels (rag_blocks.cu, rag_split.cu): region -> accumulators,
// edge key -> pixel-pair count; their drains to global memory; the sink process_item() writes through.
#pragma once
#include "rag_common.cuh"
#include "rag_core.cuh"

namespace dm {
namespace rag {
namespace blk {

// ------------------------------------------------------------------------------------ //
// per-warp shared-memory hash tables
// ------------------------------------------------------------------------------------ //
template <int C>
struct Tab {
    int* rkey;                  // [RS]
    unsigned* rarea;            // [RS]
    unsigned* rborder;          // [RS]
    unsigned* rsum;             // [C][RS]
    unsigned* rsq;              // [C][RS]
    unsigned long long* ekey;   // [ES]
    unsigned* ecnt;             // [ES]
    unsigned* used;             // [0] region slots used, [1] edge slots used
    unsigned* base;

    __device__ __forceinline__ static Tab from(unsigned* tab) {
        Tab T;
        T.base = tab;
        T.rkey = (int*)tab;
        T.rarea = tab + RS;
        T.rborder = tab + 2 * RS;
        T.rsum = tab + 3 * RS;
        T.rsq = tab + (3 + C) * RS;
        T.ekey = (unsigned long long*)(tab + (3 + 2 * C) * RS);
        T.ecnt = tab + (3 + 2 * C) * RS + 2 * ES;
        T.used = tab + (3 + 2 * C) * RS + 3 * ES;
        return T;
    }
};

// (a ^ b) | (b ^ c) and a | b | c as single LOP3s (written as PTX: left to itself the compiler turns the 24-value
// equality test into one serial chain of 24 dependent ISETP.NE.OR)
__device__ __forceinline__ unsigned eq3(int a, int b, int c) {
    unsigned d;
    asm("lop3.b32 %0, %1, %2, %3, 0x7E;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ unsigned or3(unsigned a, unsigned b, unsigned c) {
    unsigned d;
    asm("lop3.b32 %0, %1, %2, %3, 0xFE;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ void red_add_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("red.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

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
__device__ __forceinline__ void edge_add(const Tab<C>& T, const Params& P, unsigned long long key, unsigned cnt) {
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

template <int C>
__device__ __forceinline__ void global_region_add(const Params& P, int label, unsigned area, unsigned border,
                                                  const unsigned* s, const unsigned* q) {
    if ((unsigned)label >= (unsigned)P.n_regions) {
        atomicExch(&P.counters[3], 1ull);
        return;
    }
    if (area) red_add_u64(&P.area[label], (unsigned long long)area);
    if (border) red_add_u64(&P.border[label], (unsigned long long)border);
    if (C > 0 && area) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            red_add_u64(&P.bsum[(size_t)label * C + c], (unsigned long long)s[c]);
            red_add_u64(&P.bsq[(size_t)label * C + c], (unsigned long long)q[c]);
        }
    }
}

template <int C>
__device__ __noinline__ void global_region_add_slow(const Params* P, int label, unsigned area, unsigned border, uint4 s4,
                                                    uint4 q4) {
    const unsigned s[4] = {s4.x, s4.y, s4.z, s4.w}, q[4] = {q4.x, q4.y, q4.z, q4.w};
    global_region_add<C>(*P, label, area, border, s, q);
}

// accumulators of one label -> the warp's region table (or straight to global when it is full: rare, out of line)
template <int C>
__device__ __forceinline__ void table_region_add(const Tab<C>& T, const Params& P, int label, unsigned area,
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
        uint4 s4 = make_uint4(0, 0, 0, 0), q4 = make_uint4(0, 0, 0, 0);
        if (C > 0) { s4.x = s[0]; q4.x = q[0]; }
        if (C > 1) { s4.y = s[1 % (C > 0 ? C : 1)]; q4.y = q[1 % (C > 0 ? C : 1)]; }
        if (C > 2) { s4.z = s[2 % (C > 0 ? C : 1)]; q4.z = q[2 % (C > 0 ? C : 1)]; }
        if (C > 3) { s4.w = s[3 % (C > 0 ? C : 1)]; q4.w = q[3 % (C > 0 ? C : 1)]; }
        global_region_add_slow<C>(&P, label, area, border, s4, q4);
    }
}

// out-of-line versions for the pixel-by-pixel path of process_item (windows with more than four labels)
template <int C>
__device__ __noinline__ void slow_edge_add(unsigned* tab, const Params* P, int a, int b, unsigned cnt) {
    edge_add<C>(Tab<C>::from(tab), *P, pack_key(a, b), cnt);
}
template <int C>
__device__ __noinline__ void slow_region_add(unsigned* tab, const Params* P, int label, unsigned area, unsigned border,
                                             uint4 s4, uint4 q4) {
    const unsigned s[4] = {s4.x, s4.y, s4.z, s4.w}, q[4] = {q4.x, q4.y, q4.z, q4.w};
    table_region_add<C>(Tab<C>::from(tab), *P, label, area, border, s, q);
}

template <int C>
struct Sink {
    const Tab<C>& T;
    const Params& P;
    __device__ __forceinline__ void region(int l, unsigned area, unsigned sides, const unsigned* s, const unsigned* q) const {
        table_region_add<C>(T, P, l, area, sides, s, q);
    }
    __device__ __forceinline__ void edge(int a, int b, unsigned n) const { edge_add<C>(T, P, pack_key(a, b), n); }
    __device__ __forceinline__ void region_slow(int l, unsigned area, unsigned sides, const unsigned* s,
                                                const unsigned* q) const {
        uint4 s4 = make_uint4(0, 0, 0, 0), q4 = make_uint4(0, 0, 0, 0);
        if (C > 0) { s4.x = s[0]; q4.x = q[0]; }
        if (C > 1) { s4.y = s[1 % (C > 0 ? C : 1)]; q4.y = q[1 % (C > 0 ? C : 1)]; }
        if (C > 2) { s4.z = s[2 % (C > 0 ? C : 1)]; q4.z = q[2 % (C > 0 ? C : 1)]; }
        if (C > 3) { s4.w = s[3 % (C > 0 ? C : 1)]; q4.w = q[3 % (C > 0 ? C : 1)]; }
        slow_region_add<C>(T.base, &P, l, area, sides, s4, q4);
    }
    __device__ __forceinline__ void edge_slow(int a, int b, unsigned n) const { slow_edge_add<C>(T.base, &P, a, b, n); }
};

// label at window position p of item j, read back from the item buffer (run-time p: no register indexing)
template <int ICAP>
struct ItemPick {
    const int* words;   // the item buffer as words
    int j;
    __device__ __forceinline__ int operator()(int p) const {
        const int r = (p * 13) >> 6;             // p / 5 for p < 24
        const int k = p - 5 * r;
        const int vec = p >= 20 ? 0 : (k == 4 ? 5 : 1 + r);
        const int w = p >= 20 ? p - 20 : (k == 4 ? r : k);
        return words[(vec * ICAP + j) * 4 + w];
    }
};

// Drain the warp's tables to global memory (whole warp, convergent).
template <int C>
__device__ __forceinline__ void drain_tables_impl(unsigned* tab, const Params* Pp, int lane) {
    constexpr int CW = C > 0 ? C : 1;
    const Tab<C> T = Tab<C>::from(tab);
    const Params& P = *Pp;
    static_assert(RS == 32 && ES == 64, "one / two slots per lane");
    __syncwarp();
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

template <int C>
__device__ __noinline__ void drain_tables(unsigned* tab, const Params* Pp, int lane) {
    drain_tables_impl<C>(tab, Pp, lane);
}

// process_item sink with every operation inline (the warp-specialised kernel runs under per-role register budgets and
// makes no calls)
template <int C>
struct InlineSink {
    const Tab<C>& T;
    const Params& P;
    __device__ __forceinline__ void region(int l, unsigned area, unsigned sides, const unsigned* s, const unsigned* q) const {
        const int slot = region_slot(T.rkey, T.used, l);
        if (slot >= 0) {
            if (area) atomicAdd(&T.rarea[slot], area);
            if (sides) atomicAdd(&T.rborder[slot], sides);
            if (C > 0 && area) {
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    atomicAdd(&T.rsum[c * RS + slot], s[c]);
                    atomicAdd(&T.rsq[c * RS + slot], q[c]);
                }
            }
        } else {
            global_region_add<C>(P, l, area, sides, s, q);
        }
    }
    __device__ __forceinline__ void edge(int a, int b, unsigned n) const { edge_add<C>(T, P, pack_key(a, b), n); }
    __device__ __forceinline__ void region_slow(int l, unsigned area, unsigned sides, const unsigned* s,
                                                const unsigned* q) const {
        region(l, area, sides, s, q);
    }
    __device__ __forceinline__ void edge_slow(int a, int b, unsigned n) const { edge(a, b, n); }
};

}  // namespace blk
}  // namespace rag
}  // namespace dm
