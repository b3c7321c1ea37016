/* dm_oracle.c -- plain C restatement of the integer stages of the region-merging path that the reference has no
 * code for (R1 raster RAG, band statistics, R9 component roots).  TEST INFRASTRUCTURE ONLY:
 * a second, independent implementation beside oracle_np.py (PARITY UNPINNED stages are cross-checked numpy vs C vs
 * CUDA).  Built by oracle/build.py into oracle/_build/libdm_oracle.so; nothing in deepmerge_b200 links it.
 *
 * Semantics follow the written spec: 4-adjacency, labels < 0 are nodata, an edge key is (min << 32) | max, a tile
 * owns the pixel pairs (y, y+1) of its own rows, perimeter = sides facing another label, nodata or the image border. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static int cmp_u64(const void* a, const void* b) {
    const uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
    return x < y ? -1 : x > y;
}

static uint64_t pack(int32_t a, int32_t b) {
    const uint32_t lo = (uint32_t)(a < b ? a : b), hi = (uint32_t)(a < b ? b : a);
    return ((uint64_t)lo << 32) | hi;
}

/* labels [rows_avail, W] (pitch W); the tile owns rows [0, rows_own).  Outputs: area / perim int64 [R] (zeroed here),
 * keys / blen (capacity cap).  Returns the number of unique edges, or -1 when cap is too small / a label >= R. */
int64_t dmo_build_rag(const int32_t* L, int64_t rows_own, int64_t rows_avail, int64_t W, int64_t R, int top_border,
                      int bottom_border, int64_t* area, int64_t* perim, uint64_t* keys, uint32_t* blen, int64_t cap) {
    memset(area, 0, (size_t)R * sizeof(int64_t));
    memset(perim, 0, (size_t)R * sizeof(int64_t));
    int64_t n = 0;
    uint64_t* raw = (uint64_t*)malloc((size_t)(2 * rows_own * W + 1) * sizeof(uint64_t));
    if (!raw) return -1;
    for (int64_t y = 0; y < rows_own; ++y)
        for (int64_t x = 0; x < W; ++x) {
            const int32_t a = L[y * W + x];
            if (a >= R) { free(raw); return -1; }
            if (a >= 0) {
                area[a] += 1;
                if (x == 0) perim[a] += 1;
                if (x == W - 1) perim[a] += 1;
                if (y == 0 && top_border) perim[a] += 1;
                if (y == rows_own - 1 && rows_avail == rows_own && bottom_border) perim[a] += 1;
            }
            for (int d = 0; d < 2; ++d) {                   /* right neighbour, lower neighbour */
                const int64_t xx = x + (d == 0), yy = y + (d == 1);
                if (xx >= W || yy >= rows_avail) continue;
                const int32_t b = L[yy * W + xx];
                if (b >= R) { free(raw); return -1; }
                if (a == b) continue;
                if (a >= 0) perim[a] += 1;                  /* the pair's owner counts the side for both pixels */
                if (b >= 0) perim[b] += 1;
                if (a >= 0 && b >= 0) raw[n++] = pack(a, b);
            }
        }
    qsort(raw, (size_t)n, sizeof(uint64_t), cmp_u64);
    int64_t e = 0;
    for (int64_t i = 0; i < n;) {
        int64_t j = i;
        while (j < n && raw[j] == raw[i]) ++j;
        if (e >= cap) { free(raw); return -1; }
        keys[e] = raw[i];
        blen[e] = (uint32_t)(j - i);
        ++e;
        i = j;
    }
    free(raw);
    return e;
}

/* per-region band sums and sums of squares of a uint8 [rows, W, C] image (exact integers) */
void dmo_pool_bands(const int32_t* L, const uint8_t* img, int64_t rows, int64_t W, int64_t C, int64_t R, uint64_t* sum,
                    uint64_t* sumsq) {
    memset(sum, 0, (size_t)(R * C) * sizeof(uint64_t));
    memset(sumsq, 0, (size_t)(R * C) * sizeof(uint64_t));
    for (int64_t p = 0; p < rows * W; ++p) {
        const int32_t a = L[p];
        if (a < 0 || a >= R) continue;
        for (int64_t c = 0; c < C; ++c) {
            const uint64_t v = img[p * C + c];
            sum[a * C + c] += v;
            sumsq[a * C + c] += v * v;
        }
    }
}

/* root[x] = minimum id of x's connected component over the n edges (u[i], v[i]): union-find with path halving,
 * smaller root wins */
static int32_t find(int32_t* p, int32_t x) {
    while (p[x] != x) {
        p[x] = p[p[x]];
        x = p[x];
    }
    return x;
}
void dmo_min_roots(int64_t R, const int32_t* u, const int32_t* v, int64_t n, int32_t* root) {
    for (int64_t i = 0; i < R; ++i) root[i] = (int32_t)i;
    for (int64_t i = 0; i < n; ++i) {
        int32_t a = find(root, u[i]), b = find(root, v[i]);
        if (a == b) continue;
        if (a < b) root[b] = a;
        else root[a] = b;
    }
    for (int64_t i = 0; i < R; ++i) root[i] = find(root, (int32_t)i);
}

/* labels'[p] = root[labels[p]], nodata kept */
void dmo_relabel(const int32_t* L, int64_t n, const int32_t* root, int64_t R, int32_t* out) {
    for (int64_t p = 0; p < n; ++p) out[p] = (L[p] >= 0 && L[p] < R) ? root[L[p]] : L[p];
}
