// Whole-step and measurement kernels for registers of at most kSmallMaxBits qubits on one GPU
// (csrc/qca_small.cu): internal interface.
#pragma once
#include <cuda_runtime.h>

#include "qca_common.cuh"

namespace qca {

constexpr int kSmallMaxBits = 13;
constexpr int kSmallMaxTerms = 256;   // Chebyshev terms of one step (R t <= ~200)

struct SmallStepArgs {
    const double* src[2];   // resident state, per plane
    double* dst[2];         // evolved state (may alias src)
    int nbits, distance, nterms;
    unsigned interval_mask;
    double gamma;           // sgn * 2 / R
    double gamma_last;      // sgn / R
    double coef[kSmallMaxTerms];
};

int32_t launch_small_step(const SmallStepArgs& a, int nplanes, cudaStream_t stream);
int32_t launch_small_measure(const double* re, const double* im, int nbits, double* d_sums, cudaStream_t stream);

}  // namespace qca
