// Internal launch interface of the DMMA contraction kernel (csrc/qca_zgemm.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qca {

struct ZgemmArgs {
    const double2* a; const double2* b; double2* c;
    long long a_sg, a_ss, a_sm, a_sk;
    long long b_sg, b_ss, b_sk;
    long long c_sg, c_sm;
    int M, N, K, S, G;
    int conj_a;   // use conj(A)
    int nsplit;   // split-K: blockIdx.z = g * nsplit + split; partial sums go to c + split * c_ssplit
    long long c_ssplit;
    // Structural zeros of the MPO (G <= 4, S <= 32, M / chan_len <= 32 when used):
    // seg_mask[g] bit s clear: segment s of batch g contributes nothing (its A block is identically zero) and is skipped;
    // chan_mask[g] bit c clear: rows [c*chan_len, (c+1)*chan_len) of C_g are never read: tiles inside them are not computed.
    int use_masks;
    unsigned seg_mask[4], chan_mask[4];
    int chan_len;
};

// validates and launches (event-timed while qca_zgemm_profile is on)
int32_t zgemm_launch(const ZgemmArgs& p, cudaStream_t stream);

}  // namespace qca
