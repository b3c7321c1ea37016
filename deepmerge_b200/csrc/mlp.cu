// R8: pair-MLP scorer (Nets.py:28-35 layer structure) -- placeholder entry points; the
// tcgen05 implementation replaces this file.
#include "common.cuh"
using namespace dm;
extern "C" size_t dm_mlp_packed_bytes(int64_t, int64_t, int64_t) { return 0; }
extern "C" int dm_mlp_pack(const float*, const float*, const float*, const float*, const float*, const float*, int64_t,
                           int64_t, int64_t, void*, dm_stream_t) { return DM_ERR_UNSUPPORTED; }
extern "C" int dm_score_mlp_bf16(const float*, int64_t, const uint64_t*, const int64_t*, int64_t, const void*, int64_t,
                                 int64_t, int64_t, float*, float*, dm_stream_t) { return DM_ERR_UNSUPPORTED; }
extern "C" int dm_mlp_forward_bf16(const float*, int64_t, const void*, int64_t, int64_t, int64_t, float*, float*,
                                   dm_stream_t) { return DM_ERR_UNSUPPORTED; }
