// Shared helpers for the deepmerge_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/deepmerge_b200.h"

namespace dm {

extern thread_local int g_last_cuda_error;
extern long long g_launch_count;   // kernels launched by this library in this process
#define DM_COUNT_LAUNCH() (__atomic_add_fetch(&dm::g_launch_count, 1, __ATOMIC_RELAXED))

inline int cuda_fail(cudaError_t e) {
    g_last_cuda_error = (int)e;
    return DM_ERR_CUDA;
}

#define DM_CUDA(expr)                                   \
    do {                                                \
        cudaError_t _e = (expr);                        \
        if (_e != cudaSuccess) return dm::cuda_fail(_e); \
    } while (0)

#define DM_LAUNCH_CHECK() DM_CUDA(cudaPeekAtLastError())

#define DM_TRY(expr)             \
    do {                         \
        int _r = (expr);         \
        if (_r != DM_OK) return _r; \
    } while (0)

inline cudaStream_t S(dm_stream_t s) { return (cudaStream_t)s; }

__host__ __device__ inline long long imin64(long long a, long long b) { return a < b ? a : b; }
__host__ __device__ inline long long imax64(long long a, long long b) { return a > b ? a : b; }

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

int num_sms();

// Carves a caller-provided workspace into 256-byte aligned pieces.
struct Carver {
    char* base;
    size_t off = 0;
    explicit Carver(void* p) : base((char*)p) {}
    template <typename T>
    T* take(size_t n) {
        off = align_up(off, 256);
        T* r = (T*)(base + off);
        off += n * sizeof(T);
        return r;
    }
    size_t used() const { return align_up(off, 256); }
};

__device__ __forceinline__ uint64_t pack_key(int a, int b) {
    unsigned lo = (unsigned)min(a, b), hi = (unsigned)max(a, b);
    return ((uint64_t)lo << 32) | hi;
}
__device__ __forceinline__ int key_lo(uint64_t k) { return (int)(k >> 32); }
__device__ __forceinline__ int key_hi(uint64_t k) { return (int)(k & 0xffffffffu); }

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// streaming 128-bit global accesses that do not allocate in L1
__device__ __forceinline__ int4 ldg_stream(const int4* p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream(int4* p, const int4& v) {
    asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// number of bits needed to represent ids in [0, n)
inline int bits_for(int64_t n) {
    int b = 1;
    while (b < 32 && ((int64_t)1 << b) < n) ++b;
    return b;
}

}  // namespace dm
