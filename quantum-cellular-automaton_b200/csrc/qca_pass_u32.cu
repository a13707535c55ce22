// Instantiations of the fast tile-pass kernel for 32-bit amplitude indices.
#include "qca_pass.cuh"

namespace qca {

// later passes: one recurrence operand (c = out) plus up to two remote slots
template <int L>
static PassKernel later_pass(int nunc, int nrem, int rd) {
    if (nunc != 1) return nullptr;
    if (nrem == 1 && rd == 6) return pass_kernel_v2<unsigned int, L, false, 1, 1, 6>;
    switch (nrem) {
        case 0: return pass_kernel_v2<unsigned int, L, false, 1, 0>;
        case 1: return pass_kernel_v2<unsigned int, L, false, 1, 1>;
        case 2: return pass_kernel_v2<unsigned int, L, false, 1, 2>;
        default: return nullptr;
    }
}

PassKernel fast_pass_kernel_u32(int low_bits, int nunc, int nrem, int rd) {
    switch (low_bits) {
        case 13:  // pass 0: no operand (test hook) or one / both recurrence operands, up to two remote slots
            if (nunc == 2 && nrem == 1 && rd == 6) return pass_kernel_v2<unsigned int, 13, true, 2, 1, 6>;
            if (nunc == 0 && nrem == 0) return pass_kernel_v2<unsigned int, 13, true, 0, 0>;
            if (nunc == 0 && nrem == 1) return pass_kernel_v2<unsigned int, 13, true, 0, 1>;
            if (nunc == 0 && nrem == 2) return pass_kernel_v2<unsigned int, 13, true, 0, 2>;
            if (nunc == 1 && nrem == 0) return pass_kernel_v2<unsigned int, 13, true, 1, 0>;
            if (nunc == 1 && nrem == 1) return pass_kernel_v2<unsigned int, 13, true, 1, 1>;
            if (nunc == 1 && nrem == 2) return pass_kernel_v2<unsigned int, 13, true, 1, 2>;
            if (nunc == 2 && nrem == 0) return pass_kernel_v2<unsigned int, 13, true, 2, 0>;
            if (nunc == 2 && nrem == 1) return pass_kernel_v2<unsigned int, 13, true, 2, 1>;
            if (nunc == 2 && nrem == 2) return pass_kernel_v2<unsigned int, 13, true, 2, 2>;
            return nullptr;
        case 12: return later_pass<12>(nunc, nrem, rd);
        case 11: return later_pass<11>(nunc, nrem, rd);
        case 10: return later_pass<10>(nunc, nrem, rd);
        case 9: return later_pass<9>(nunc, nrem, rd);
        case 8: return later_pass<8>(nunc, nrem, rd);
        case 7: return later_pass<7>(nunc, nrem, rd);
        case 6: return later_pass<6>(nunc, nrem, rd);
        case 5: return later_pass<5>(nunc, nrem, rd);
        case 4: return later_pass<4>(nunc, nrem, rd);
        default: return nullptr;
    }
}


// persistent variant (operand rings run across tile boundaries): sharded launches only
template <int L, bool FLIP_LOW, int NUNC>
static PassKernel persistent_by_slots(int nrem, int rd) {
    if (nrem == 1) return rd == 8 ? pass_kernel_v2p<unsigned int, L, FLIP_LOW, NUNC, 1, (NUNC == 2 ? 4 : 8)> : pass_kernel_v2p<unsigned int, L, FLIP_LOW, NUNC, 1, 4>;
    if (nrem == 2) return pass_kernel_v2p<unsigned int, L, FLIP_LOW, NUNC, 2, 4>;
    return nullptr;
}

PassKernel persistent_pass_kernel_u32(int low_bits, int nunc, int nrem, int rd) {
    if (low_bits == 13) {
        switch (nunc) {
            case 0: return persistent_by_slots<13, true, 0>(nrem, rd);
            case 1: return persistent_by_slots<13, true, 1>(nrem, rd);
            case 2: return persistent_by_slots<13, true, 2>(nrem, rd);
            default: return nullptr;
        }
    }
    if (nunc != 1) return nullptr;
    switch (low_bits) {
        case 12: return persistent_by_slots<12, false, 1>(nrem, rd);
        case 11: return persistent_by_slots<11, false, 1>(nrem, rd);
        case 10: return persistent_by_slots<10, false, 1>(nrem, rd);
        case 9: return persistent_by_slots<9, false, 1>(nrem, rd);
        case 8: return persistent_by_slots<8, false, 1>(nrem, rd);
        case 7: return persistent_by_slots<7, false, 1>(nrem, rd);
        case 6: return persistent_by_slots<6, false, 1>(nrem, rd);
        case 5: return persistent_by_slots<5, false, 1>(nrem, rd);
        case 4: return persistent_by_slots<4, false, 1>(nrem, rd);
        default: return nullptr;
    }
}

PassKernel generic_pass_kernel(bool wide) {
    return wide ? pass_kernel_generic<unsigned long long> : pass_kernel_generic<unsigned int>;
}

}  // namespace qca
