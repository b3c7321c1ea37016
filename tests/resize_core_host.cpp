// Host harness around deepmerge_b200/csrc/resize_core.cuh (TEST INFRASTRUCTURE): the loops of resize_area_kernel run
// sequentially on the CPU over the same per-value functions the kernel calls.  Built by tests/test_resize_core_cpu.py
// with g++ -ffp-contract=off.
#include <stdlib.h>
#include "../deepmerge_b200/csrc/resize_core.cuh"

extern "C" int resize_planes_host(const uint8_t* patches, long n_planes, int s, int t, int mode, const int32_t* ti,
                                  const float* tf, float* out) {
    using namespace dm::resize;
    float* buf = mode == 1 ? (float*)malloc(sizeof(float) * (size_t)s * t) : 0;
    for (long p = 0; p < n_planes; ++p) {
        const uint8_t* P = patches + (size_t)p * s * s;
        float* O = out + (size_t)p * t * t;
        if (mode == 0) {
            for (int o = 0; o < t * t; ++o) O[o] = shrink_integer(P, s, t, o / t, o % t);
        } else if (mode == 1) {
            const int32_t *start = ti, *src = ti + t + 1;
            for (int o = 0; o < s * t; ++o) buf[o] = shrink_row(P, s, t, o / t, o % t, start, src, tf);
            for (int o = 0; o < t * t; ++o) O[o] = shrink_col(buf, s, t, o / t, o % t, start, src, tf);
        } else {
            for (int o = 0; o < t * t; ++o) O[o] = enlarge(P, s, t, o / t, o % t, ti);
        }
    }
    free(buf);
    return 0;
}
