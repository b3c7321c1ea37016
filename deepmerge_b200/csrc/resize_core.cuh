// Arithmetic of dm_resize_area (OpenCV's INTER_AREA for square uint8 planes), one output value per call.
// Plain C++ that compiles for the device (resize.cu) and for the host: tests/test_resize_core_cpu.py builds this very
// header with g++ and checks it bit for bit against the oracle, so that only the CUDA indexing around it is left to
// the GPU tests.  Every float operation is a single IEEE multiply, add or divide (no fused multiply-add): on the
// device through the _rn intrinsics, on the host through -ffp-contract=off.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define DM_RESIZE_FN __host__ __device__ __forceinline__
#else
#include <math.h>
#define DM_RESIZE_FN static inline
#endif

namespace dm {
namespace resize {

DM_RESIZE_FN float rs_mul(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}
DM_RESIZE_FN float rs_add(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}
DM_RESIZE_FN int rs_round(float a) {          // round half to even (cvRound)
#if defined(__CUDA_ARCH__)
    return __float2int_rn(a);
#else
    return (int)lrintf(a);
#endif
}
DM_RESIZE_FN float rs_unit(int v) {           // uint8 -> float32 / 255 (resize_data's last line)
    v = v < 0 ? 0 : (v > 255 ? 255 : v);
#if defined(__CUDA_ARCH__)
    return __fdiv_rn((float)v, 255.0f);
#else
    return (float)v / 255.0f;
#endif
}
DM_RESIZE_FN int rs_clamp(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// mode 0: integer shrink factor k = s / t
DM_RESIZE_FN float shrink_integer(const uint8_t* P, int s, int t, int dy, int dx) {
    const int k = s / t;
    int sum = 0;
    for (int v = 0; v < k; ++v)
        for (int u = 0; u < k; ++u) sum += P[(dy * k + v) * s + dx * k + u];
    if (k == 1) return rs_unit(sum);
    if (k == 2) return rs_unit((sum + 2) >> 2);
    const float inv = 1.0f / (float)(k * k);
    return rs_unit(rs_round(rs_mul((float)sum, inv)));
}

// mode 1, first pass: source row sy reduced along x into destination column dx (table order)
DM_RESIZE_FN float shrink_row(const uint8_t* P, int s, int t, int sy, int dx, const int32_t* start, const int32_t* src,
                              const float* w) {
    float acc = 0.0f;
    for (int e = start[dx]; e < start[dx + 1]; ++e) acc = rs_add(acc, rs_mul((float)P[sy * s + rs_clamp(src[e], 0, s - 1)], w[e]));
    return acc;
}
// mode 1, second pass: the reduced rows buf[s][t] combined along y
DM_RESIZE_FN float shrink_col(const float* buf, int s, int t, int dy, int dx, const int32_t* start, const int32_t* src,
                              const float* w) {
    const int e0 = start[dy], e1 = start[dy + 1];
    float acc = 0.0f;
    if (e0 < e1) acc = rs_mul(w[e0], buf[rs_clamp(src[e0], 0, s - 1) * t + dx]);
    for (int e = e0 + 1; e < e1; ++e) acc = rs_add(acc, rs_mul(w[e], buf[rs_clamp(src[e], 0, s - 1) * t + dx]));
    return rs_unit(rs_round(acc));
}

// mode 2: enlargement; ti = sx[t] ++ a0[t] ++ a1[t] ++ xmax
DM_RESIZE_FN float enlarge(const uint8_t* P, int s, int t, int dy, int dx, const int32_t* ti) {
    const int32_t *sx = ti, *a0 = ti + t, *a1 = ti + 2 * t;
    const int xmax = ti[3 * t];
    const int x0 = rs_clamp(sx[dx], 0, s - 1), x1 = x0 + 1 < s ? x0 + 1 : s - 1;
    const int y0 = rs_clamp(sx[dy], 0, s - 1), y1 = y0 + 1 < s ? y0 + 1 : s - 1;
    int S0, S1;
    if (dx >= xmax) {
        S0 = (int)P[y0 * s + x0] * 2048;
        S1 = (int)P[y1 * s + x0] * 2048;
    } else {
        S0 = (int)P[y0 * s + x0] * a0[dx] + (int)P[y0 * s + x1] * a1[dx];
        S1 = (int)P[y1 * s + x0] * a0[dx] + (int)P[y1 * s + x1] * a1[dx];
    }
    return rs_unit((((a0[dy] * (S0 >> 4)) >> 16) + ((a1[dy] * (S1 >> 4)) >> 16) + 2) >> 2);
}

}  // namespace resize
}  // namespace dm
