// Instantiations of the cluster tile-pass kernel (qca_pass3.cuh) for 32-bit amplitude indices.
#include "qca_pass3.cuh"

namespace qca {

template <typename I, int L, bool FLIP_LOW, int NUNC>
static Pass3Kernel by_cluster(int cb) {
    switch (cb) {
        case 0: return pass_kernel_v3<I, L, FLIP_LOW, NUNC, 0>;
        case 1: return pass_kernel_v3<I, L, FLIP_LOW, NUNC, 1>;
        case 2: return pass_kernel_v3<I, L, FLIP_LOW, NUNC, 2>;
        case 3: return pass_kernel_v3<I, L, FLIP_LOW, NUNC, 3>;
        default: return nullptr;
    }
}

template <typename I>
static Pass3Kernel pick(int low_bits, int nunc, int cb) {
    if (low_bits == kTile3) {   // pass 0: no operand (test hook), one or both recurrence operands
        switch (nunc) {
            case 0: return by_cluster<I, kTile3, true, 0>(cb);
            case 1: return by_cluster<I, kTile3, true, 1>(cb);
            case 2: return by_cluster<I, kTile3, true, 2>(cb);
            default: return nullptr;
        }
    }
    if (nunc != 1) return nullptr;   // later passes: the one recurrence operand c = out
    switch (low_bits) {
        case 13: return by_cluster<I, 13, false, 1>(cb);
        case 12: return by_cluster<I, 12, false, 1>(cb);
        case 11: return by_cluster<I, 11, false, 1>(cb);
        case 10: return by_cluster<I, 10, false, 1>(cb);
        case 9: return by_cluster<I, 9, false, 1>(cb);
        case 8: return by_cluster<I, 8, false, 1>(cb);
        case 7: return by_cluster<I, 7, false, 1>(cb);
        case 6: return by_cluster<I, 6, false, 1>(cb);
        case 5: return by_cluster<I, 5, false, 1>(cb);
        case 4: return by_cluster<I, 4, false, 1>(cb);
        default: return nullptr;
    }
}

#ifndef QCA_PASS3_WIDE
Pass3Kernel pass3_kernel_u32(int low_bits, int nunc, int cluster_bits) { return pick<unsigned int>(low_bits, nunc, cluster_bits); }
#else
Pass3Kernel pass3_kernel_u64(int low_bits, int nunc, int cluster_bits) { return pick<unsigned long long>(low_bits, nunc, cluster_bits); }
#endif

}  // namespace qca
