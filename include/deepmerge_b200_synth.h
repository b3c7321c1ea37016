/*
 * deepmerge_b200_synth -- device-side generator of the synthetic scenes of SURVEY.md section 8(d).
 *
 * BENCH AND TEST UTILITY, not part of the product ABI (include/deepmerge_b200.h) and not part of the reference's path:
 * libdeepmerge_b200_synth.so is a separate library that bench.py, the tools and the GPU tests load to build their inputs on
 * the device, bit-identically to oracle/oracle_np.py's generator (tests/test_gpu_parity.py::test_synth_matches_oracle).
 * Same conventions as the product header: device pointers, caller-owned buffers, stream last, 0 / negative return codes.
 */
#ifndef DEEPMERGE_B200_SYNTH_H
#define DEEPMERGE_B200_SYNTH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* dm_stream_t; /* cudaStream_t */

int dm_synth_labels(int32_t* labels, int64_t y0, int64_t rows, int64_t H, int64_t W, int64_t ld, int64_t pitch_g,
                    uint32_t seed, dm_stream_t stream);
int dm_synth_region_objects(int32_t* region_obj, int64_t H, int64_t W, int64_t pitch_g, uint32_t seed,
                            dm_stream_t stream);
int dm_synth_image(uint8_t* image, const int32_t* labels, int64_t y0, int64_t rows, int64_t W, int64_t ld,
                   int64_t C, const int32_t* region_obj, uint32_t seed, dm_stream_t stream);
int dm_synth_points(int32_t* xs, int32_t* ys, int64_t H, int64_t W, int64_t pitch_g, int64_t P, uint32_t seed,
                    dm_stream_t stream);
/* point_ids (nullable): global ids of the n_points rows (a row-tile shard passes the ids of its own points) */
int dm_synth_feats(float* feats, const int32_t* region_of_point, const int32_t* region_obj, const int64_t* point_ids,
                   int64_t n_points, int64_t D, uint32_t seed, dm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DEEPMERGE_B200_SYNTH_H */
