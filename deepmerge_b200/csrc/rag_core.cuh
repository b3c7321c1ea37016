// Per-block arithmetic of the raster pass (R1 + band pooling), host/device.
//
// The raster kernel (rag.cu) cuts every 128-pixel strip into blocks of 4 x 4 pixels, one per lane.  A block whose
// whole window -- its 16 pixels, the 4 pixels above and the 4 pixels to the right -- carries one label is handled
// by the kernel's convergent fast path (block_stats below).  Every other block is copied out as a self-contained
// ITEM (24 labels, 16 pixels of image bytes, its position) and items are processed 32 at a time, one per lane, by
// process_item(): all the border / nodata / pair logic of SURVEY.md section 8(a) R1 lives here, nowhere else.
//
// The functions are plain integer arithmetic, so the same source is compiled by g++ in the CPU test suite
// (tests/rag_core_host.cpp walks a raster exactly as the kernel does, lane by lane) and checked bit for bit against
// the oracle; what is left to the GPU tests is the staging / work-list / hash-table machinery around them.
//
// Ownership of pixel pairs (the same as oracle_np.build_rag): a block owns, for each of its pixels p, the horizontal
// pair (p, right neighbour of p) and the vertical pair (upper neighbour of p, p).  A row tile owns rows
// [0, rows_own); row rows_own (when rows_avail == rows_own + 1) is a halo that only closes vertical pairs.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define DM_HD __host__ __device__ __forceinline__
#else
#define DM_HD inline
#endif

namespace dm {
namespace ragcore {

// ---- portable integer intrinsics ------------------------------------------------------------------------------------
DM_HD unsigned popc32(unsigned x) {
#if defined(__CUDA_ARCH__)
    return (unsigned)__popc(x);
#else
    return (unsigned)__builtin_popcount(x);
#endif
}
DM_HD int ffs32(unsigned x) {   // 1-based index of the lowest set bit, 0 when x == 0
#if defined(__CUDA_ARCH__)
    return __ffs((int)x);
#else
    return x ? __builtin_ctz(x) + 1 : 0;
#endif
}
DM_HD unsigned dp4a_u8(unsigned a, unsigned b, unsigned c) {
#if defined(__CUDA_ARCH__)
    return __dp4a(a, b, c);
#else
    for (int i = 0; i < 4; ++i) c += ((a >> (8 * i)) & 0xffu) * ((b >> (8 * i)) & 0xffu);
    return c;
#endif
}
DM_HD unsigned byte_perm(unsigned a, unsigned b, unsigned s) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, b, s);
#else
    const unsigned long long v = ((unsigned long long)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; ++i) r |= (unsigned)((v >> (8 * ((s >> (4 * i)) & 7))) & 0xffu) << (8 * i);
    return r;
#endif
}

// ---- window layout ---------------------------------------------------------------------------------------------------
// 24 window positions = bit / array index: own pixel (row r, column k) -> 5 r + k; right neighbour of row r -> 5 r + 4;
// pixel above column k -> 20 + k.
constexpr int WIN = 24;
constexpr unsigned OWN_BITS = 0x7BDEFu;     // k = 0..3 of the four rows
constexpr unsigned COL0_BITS = 0x08421u;    // k = 0 of the four rows
constexpr unsigned RIGHT_BITS = 0x84210u;   // k = 4 of the four rows
DM_HD unsigned row_bits(int n) { return (1u << (5 * n)) - 1u; }   // all positions of rows 0 .. n-1, n in [0, 4]

struct Geo {
    int W, rows_own, rows_avail, top_border, bottom_border;
};

// masks of one block at (x0, y0): which positions / pairs exist
struct BlockMasks {
    unsigned PV;      // window positions that take part in anything
    unsigned SV;      // own pixels whose statistics count (owned rows, columns inside the image)
    unsigned HV;      // horizontal pairs (bit = left pixel)
    unsigned VV;      // vertical pairs (bit = lower pixel)
    unsigned SIDES[4];   // own pixels with an image-border side: left, right, top, bottom
};

DM_HD int clamp04(int v) { return v < 0 ? 0 : (v > 4 ? 4 : v); }

DM_HD BlockMasks block_masks(int x0, int y0, const Geo& g) {
    BlockMasks m;
    const int ncols = clamp04(g.W - x0);              // columns of the block inside the image
    const int nh = clamp04(g.W - x0 - 1);             // horizontal pairs per row
    const int nown = clamp04(g.rows_own - y0);        // owned rows
    const int nav = clamp04(g.rows_avail - y0);       // rows that exist (a halo row counts)
    const unsigned cols = ((1u << ncols) - 1u) * COL0_BITS;
    m.SV = cols & row_bits(nown);
    const unsigned ownv = cols & row_bits(nav);
    const unsigned rightv = (x0 + 4 < g.W) ? (RIGHT_BITS & row_bits(nown)) : 0u;
    const unsigned upv = (y0 > 0) ? (((1u << ncols) - 1u) << 20) : 0u;
    m.PV = ownv | rightv | upv;
    m.HV = (((1u << nh) - 1u) * COL0_BITS) & row_bits(nown);
    m.VV = ownv & ~(y0 > 0 ? 0u : 0x1Fu);
    const int klast = g.W - 1 - x0, rlast = g.rows_own - 1 - y0;
    m.SIDES[0] = x0 == 0 ? COL0_BITS : 0u;
    m.SIDES[1] = (klast >= 0 && klast < 4) ? (COL0_BITS << klast) : 0u;
    m.SIDES[2] = (y0 == 0 && g.top_border) ? 0xFu : 0u;
    m.SIDES[3] = (rlast >= 0 && rlast < 4 && g.rows_avail == g.rows_own && g.bottom_border) ? (0xFu << (5 * rlast)) : 0u;
    return m;
}

// does a block have to be an item whatever its labels are (it touches an image border / the end of the tile)?
DM_HD bool block_forced(int x0, int y0, const Geo& g) {
    return x0 == 0 || x0 + 4 >= g.W || y0 == 0 || y0 + 4 >= g.rows_own;
}

// the positions above the positions of M (own rows shifted down one row, the row above the block for row 0)
DM_HD unsigned upper_of(unsigned M) { return (((M & OWN_BITS) << 5) & OWN_BITS) | ((M >> 20) & 0xFu); }

// pixel pairs between the position sets A and B (disjoint)
DM_HD unsigned pairs_between(unsigned A, unsigned B, const BlockMasks& m) {
    const unsigned h = ((A & (B >> 1)) | (B & (A >> 1))) & m.HV;
    const unsigned v = ((A & upper_of(B)) | (B & upper_of(A))) & m.VV;
    return popc32(h) + popc32(v);
}

// 4 bits -> 4 byte masks
DM_HD unsigned nibble_to_bytes(unsigned nib) { return ((nib * 0x00204081u) & 0x01010101u) * 0xFFu; }

// T[c] = band c of the 4 pixels of one lane-row, one pixel per byte, from their 4*C interleaved bytes
template <int C>
DM_HD void band_transpose(const unsigned* W, unsigned* T) {
    if constexpr (C == 4) {
        const unsigned t0 = byte_perm(W[0], W[1], 0x5140), t1 = byte_perm(W[0], W[1], 0x7362);
        const unsigned t2 = byte_perm(W[2], W[3], 0x5140), t3 = byte_perm(W[2], W[3], 0x7362);
        T[0] = byte_perm(t0, t2, 0x5410);
        T[1] = byte_perm(t0, t2, 0x7632);
        T[2] = byte_perm(t1, t3, 0x5410);
        T[3] = byte_perm(t1, t3, 0x7632);
    } else {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int c = 0; c < C; ++c) {
            const int i0 = c, i1 = C + c, i2 = 2 * C + c, i3 = 3 * C + c;
            const unsigned t01 = byte_perm(W[i0 >> 2], W[i1 >> 2], (unsigned)((i0 & 3) | ((4 + (i1 & 3)) << 4)));
            const unsigned t23 = byte_perm(W[i2 >> 2], W[i3 >> 2], (unsigned)((i2 & 3) | ((4 + (i3 & 3)) << 4)));
            T[c] = byte_perm(t01, t23, 0x5410);
        }
    }
}

// unmasked statistics of one lane-row (fast path)
template <int C>
DM_HD void row_stats(const unsigned* W, unsigned* s, unsigned* q) {
    unsigned T[C > 0 ? C : 1];
    band_transpose<C>(W, T);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int c = 0; c < C; ++c) {
        s[c] = dp4a_u8(T[c], 0x01010101u, s[c]);
        q[c] = dp4a_u8(T[c], T[c], q[c]);
    }
}

// Is the whole window one label?  (d == 0)
DM_HD unsigned window_spread(const int* lab) {
    unsigned d = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int p = 1; p < WIN; p += 2) {
        const unsigned a = (unsigned)lab[p - 1], b = (unsigned)lab[p], c = (unsigned)lab[(p + 1) % WIN];
        d |= (a ^ b) | (b ^ c);
    }
    return d;
}

// ---- one item -------------------------------------------------------------------------------------------------------
// Sink: region(label, area, sides, s[C], q[C]) adds to a region's accumulators; edge(a, b, n) adds n pixel pairs
// between labels a != b (both >= 0).  region_slow / edge_slow are the same operations, kept out of line on the device.
// pick(p) returns lab[p] for a run-time p (the device reads it back from the item buffer instead of indexing registers).
template <int C, class Sink, class Pick>
DM_HD void process_item(const int* lab, const unsigned* img /* [4][CW] */, int x0, int y0, const Geo& g, Sink& sink,
                        const Pick& pick) {
    constexpr int CW = C > 0 ? C : 1;
    const BlockMasks m = block_masks(x0, y0, g);

    // nodata positions
    unsigned MN = 0;
    {
        int any = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int p = 0; p < WIN; ++p) any |= lab[p];
        if (any < 0) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int p = 0; p < WIN; ++p) MN |= (lab[p] < 0 ? 1u : 0u) << p;
            MN &= m.PV;
        }
    }

    // band bytes of the 16 pixels, transposed once
    unsigned T[4][CW];
    if (C > 0) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int r = 0; r < 4; ++r) band_transpose<C>(img + r * CW, T[r]);
    }

    // The first MAXL distinct labels, one per trip of a compact loop: position mask, masked statistics, pairs with the
    // (up to three) labels found before, whose masks wait in a rotating register queue.
    constexpr int MAXL = 4;
    unsigned rem = m.PV & ~MN;
    int Lq1 = 0, Lq2 = 0, Lq3 = 0;
    unsigned Mq1 = 0, Mq2 = 0, Mq3 = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int i = 0; i < MAXL && rem != 0; ++i) {
        const int l = pick(ffs32(rem) - 1);
        unsigned mm = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int p = 0; p < WIN; ++p) mm |= (lab[p] == l ? 1u : 0u) << p;
        mm &= rem;
        rem &= ~mm;
        {
            const unsigned ms = mm & m.SV;
            unsigned sides = popc32(ms & m.SIDES[0]) + popc32(ms & m.SIDES[1]) + popc32(ms & m.SIDES[2]) +
                             popc32(ms & m.SIDES[3]);
            if (MN) sides += pairs_between(mm, MN, m);
            unsigned s[CW], q[CW];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int c = 0; c < CW; ++c) s[c] = q[c] = 0;
            if (C > 0) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                for (int r = 0; r < 4; ++r) {
                    const unsigned bm = nibble_to_bytes((ms >> (5 * r)) & 0xFu);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                    for (int c = 0; c < C; ++c) {
                        const unsigned w = T[r][c] & bm;
                        s[c] = dp4a_u8(w, 0x01010101u, s[c]);
                        q[c] = dp4a_u8(w, w, q[c]);
                    }
                }
            }
            const unsigned area = popc32(ms);
            if (area | sides) sink.region(l, area, sides, s, q);
        }
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int j = 0; j < 3; ++j) {                   // one copy of the pair code: the queue rotates past it
            if (Mq1) {
                const unsigned n = pairs_between(mm, Mq1, m);
                if (n) sink.edge(l, Lq1, n);
            }
            const int tl = Lq1;
            const unsigned tm = Mq1;
            Lq1 = Lq2; Mq1 = Mq2;
            Lq2 = Lq3; Mq2 = Mq3;
            Lq3 = tl; Mq3 = tm;
        }
        Lq3 = Lq2; Mq3 = Mq2;
        Lq2 = Lq1; Mq2 = Mq1;
        Lq1 = l; Mq1 = mm;
    }
    if (rem == 0) return;

    // More than MAXL labels in the window (junctions of tiny regions).  The positions of `rem` are still open: their
    // pixels' statistics and every pair with a side in `rem` are done pixel by pixel, pair by pair (rows unrolled so
    // that T[r] is indexed statically; columns and neighbours at run time through pick).
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 4; ++r) {
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int k = 0; k < 4; ++k) {
            const int p = 5 * r + k;
            const unsigned bit = 1u << p;
            const int l = pick(p);
            if ((m.SV & rem & bit) && l >= 0) {
                unsigned s[CW], q[CW];
                for (int c = 0; c < CW; ++c) s[c] = q[c] = 0;
                if (C > 0) {
                    const unsigned bm = 0xFFu << (8 * k);
                    for (int c = 0; c < C; ++c) {
                        const unsigned w = T[r][c] & bm;
                        s[c] = dp4a_u8(w, 0x01010101u, 0u);
                        q[c] = dp4a_u8(w, w, 0u);
                    }
                }
                const unsigned sides = ((m.SIDES[0] & bit) ? 1u : 0u) + ((m.SIDES[1] & bit) ? 1u : 0u) +
                                       ((m.SIDES[2] & bit) ? 1u : 0u) + ((m.SIDES[3] & bit) ? 1u : 0u);
                sink.region_slow(l, 1u, sides, s, q);
            }
            // the two pairs this pixel owns: (p, right of p) and (above p, p)
            for (int which = 0; which < 2; ++which) {
                const bool ok = which == 0 ? (m.HV & bit) != 0 : (m.VV & bit) != 0;
                if (!ok) continue;
                const int po = which == 0 ? p + 1 : (r == 0 ? 20 + k : p - 5);
                if (((bit | (1u << po)) & rem) == 0) continue;      // both sides were handled by the masks
                const int o = pick(po);
                if (o == l) continue;
                if ((o | l) >= 0) {
                    sink.edge_slow(l, o, 1u);
                } else {
                    const int v = l >= 0 ? l : o;
                    if (v >= 0) {
                        unsigned z[CW];
                        for (int c = 0; c < CW; ++c) z[c] = 0;
                        sink.region_slow(v, 0u, 1u, z, z);
                    }
                }
            }
        }
    }
}

}  // namespace ragcore
}  // namespace dm
