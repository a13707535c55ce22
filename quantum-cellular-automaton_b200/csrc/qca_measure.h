// Fused (one read per tile pass) measurement of a single-plane state: internal interface of csrc/qca_measure.cu.
#pragma once
#include <cuda_runtime.h>

#include "qca_common.cuh"

namespace qca {

constexpr int kMeasureTileVals = 27;   // doubles per block in d_partials

// Sums (s0, s1, w, 0) of every cell whose index bit varies inside the tiles of pass `ps` (pass 0: local bits
// 0..12, later passes: their strided high bits), written to d_sums[4 * cell ..].  d_partials must hold
// max_blocks * kMeasureTileVals doubles.
int32_t measure_tiles(const double* re, unsigned long long namps, const qca_pass_t& ps, const ShardMap& shard, int ncells,
                      double* d_partials, int max_blocks, double* d_sums, cudaStream_t stream);

}  // namespace qca
