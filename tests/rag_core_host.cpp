// Host harness around deepmerge_b200/csrc/rag_core.cuh (TEST INFRASTRUCTURE): walks a label raster exactly as
// rag_blocks_kernel does -- 128-pixel strips, one 4 x 4 block per lane, fast path for one-label windows with per-lane
// accumulators carried down the strip, every other block through process_item() -- with plain host containers in place
// of the shared-memory tables.  Built by tests/test_rag_core_cpu.py with g++.
#include <map>
#include <vector>
#include <stdint.h>
#include <string.h>
#include "../deepmerge_b200/csrc/rag_core.cuh"

using namespace dm::ragcore;

namespace {
template <int C>
struct HostSink {
    static constexpr int CW = C > 0 ? C : 1;
    int64_t *area, *border;
    uint64_t *bsum, *bsq;
    long n_regions;
    std::map<uint64_t, uint64_t>* edges;
    int bad = 0;
    void region(int l, unsigned a, unsigned sides, const unsigned* s, const unsigned* q) {
        if (l < 0 || l >= n_regions) { bad = 1; return; }
        area[l] += a;
        border[l] += sides;
        for (int c = 0; c < C; ++c) {
            bsum[(size_t)l * C + c] += s[c];
            bsq[(size_t)l * C + c] += q[c];
        }
    }
    void edge(int a, int b, unsigned n) {
        const uint64_t lo = a < b ? a : b, hi = a < b ? b : a;
        if ((long)hi >= n_regions) { bad = 1; return; }
        (*edges)[(lo << 32) | hi] += n;
    }
    void region_slow(int l, unsigned a, unsigned sides, const unsigned* s, const unsigned* q) { region(l, a, sides, s, q); }
    void edge_slow(int a, int b, unsigned n) { edge(a, b, n); }
};

struct ArrayPick {
    const int* lab;
    int operator()(int p) const { return lab[p]; }
};

template <int C>
int walk(const int32_t* labels, long rows_own, long rows_avail, long W, long ld, const uint8_t* image, long pitch,
         long n_regions, int top, int bot, int64_t* area, int64_t* border, uint64_t* bsum, uint64_t* bsq,
         std::map<uint64_t, uint64_t>& edges, long* n_items, long* n_fast) {
    constexpr int CW = C > 0 ? C : 1;
    HostSink<C> sink{area, border, bsum, bsq, n_regions, &edges};
    Geo g{(int)W, (int)rows_own, (int)rows_avail, top, bot};
    auto lab_at = [&](long y, long x) -> int { return (y >= 0 && y < rows_avail && x >= 0 && x < W) ? labels[y * ld + x] : 0; };
    auto byte_at = [&](long y, long xb) -> unsigned {
        return (C > 0 && y >= 0 && y < rows_own && xb < W * C) ? image[y * pitch + xb] : 0u;
    };
    for (long sx0 = 0; sx0 < W; sx0 += 128) {
        for (int lane = 0; lane < 32; ++lane) {
            const long x0 = sx0 + 4 * lane;
            int cur = -1;
            unsigned ar = 0, s[CW], q[CW];
            for (int c = 0; c < CW; ++c) s[c] = q[c] = 0;
            auto flush = [&]() {
                if (cur >= 0 && ar) sink.region(cur, ar, 0, s, q);
                ar = 0;
                for (int c = 0; c < CW; ++c) s[c] = q[c] = 0;
            };
            for (long y0 = 0; y0 < rows_avail; y0 += 4) {
                if (x0 >= W) continue;
                int lab[WIN];
                unsigned img[4][CW];
                for (int r = 0; r < 4; ++r) {
                    for (int k = 0; k < 5; ++k) lab[5 * r + k] = lab_at(y0 + r, x0 + k);
                    for (int w = 0; w < CW; ++w) {
                        unsigned v = 0;
                        for (int b = 0; b < 4; ++b) v |= byte_at(y0 + r, x0 * C + 4 * w + b) << (8 * b);
                        img[r][w] = v;
                    }
                }
                for (int k = 0; k < 4; ++k) lab[20 + k] = lab_at(y0 - 1, x0 + k);
                const bool item = block_forced((int)x0, (int)y0, g) || window_spread(lab) != 0;
                if (!item) {
                    ++*n_fast;
                    if (lab[0] != cur) {
                        flush();
                        cur = lab[0];
                    }
                    ar += 16;
                    for (int r = 0; r < 4; ++r) row_stats<C>(img[r], s, q);
                } else {
                    ++*n_items;
                    process_item<C>(lab, &img[0][0], (int)x0, (int)y0, g, sink, ArrayPick{lab});
                }
            }
            flush();
        }
    }
    return sink.bad;
}
}  // namespace

extern "C" long rag_core_host(const int32_t* labels, long rows_own, long rows_avail, long W, long ld, const uint8_t* image,
                              int C, long pitch, long n_regions, int top, int bot, int64_t* area, int64_t* border,
                              uint64_t* bsum, uint64_t* bsq, uint64_t* keys_out, uint32_t* cnt_out, long cap, long* stats) {
    std::map<uint64_t, uint64_t> edges;
    long n_items = 0, n_fast = 0;
    int bad;
    switch (C) {
        case 0: bad = walk<0>(labels, rows_own, rows_avail, W, ld, image, pitch, n_regions, top, bot, area, border, bsum, bsq, edges, &n_items, &n_fast); break;
        case 1: bad = walk<1>(labels, rows_own, rows_avail, W, ld, image, pitch, n_regions, top, bot, area, border, bsum, bsq, edges, &n_items, &n_fast); break;
        case 2: bad = walk<2>(labels, rows_own, rows_avail, W, ld, image, pitch, n_regions, top, bot, area, border, bsum, bsq, edges, &n_items, &n_fast); break;
        case 3: bad = walk<3>(labels, rows_own, rows_avail, W, ld, image, pitch, n_regions, top, bot, area, border, bsum, bsq, edges, &n_items, &n_fast); break;
        case 4: bad = walk<4>(labels, rows_own, rows_avail, W, ld, image, pitch, n_regions, top, bot, area, border, bsum, bsq, edges, &n_items, &n_fast); break;
        default: return -1;
    }
    if (bad) return -2;
    long n = 0;
    for (auto& kv : edges) {
        if (n < cap) {
            keys_out[n] = kv.first;
            cnt_out[n] = (uint32_t)kv.second;
        }
        ++n;
    }
    stats[0] = n_items;
    stats[1] = n_fast;
    return n;
}
