// Instantiations of the fast tile-pass kernel for 32-bit amplitude indices.
#include "qca_pass.cuh"

namespace qca {

// later passes: one recurrence operand (c = out) plus up to three remote terms
template <int L>
static PassKernel later_pass(int nunc, int ncond) {
    if (nunc != 1) return nullptr;
    switch (ncond) {
        case 0: return pass_kernel_v2<unsigned int, L, false, 1, 0>;
        case 1: return pass_kernel_v2<unsigned int, L, false, 1, 1>;
        case 2: return pass_kernel_v2<unsigned int, L, false, 1, 2>;
        case 3: return pass_kernel_v2<unsigned int, L, false, 1, 3>;
        default: return nullptr;
    }
}

PassKernel fast_pass_kernel_u32(int low_bits, int nunc, int ncond) {
    switch (low_bits) {
        case 13:  // pass 0: no operand (test hook) or both recurrence operands, at most one remote term
            if (nunc == 0 && ncond == 0) return pass_kernel_v2<unsigned int, 13, true, 0, 0>;
            if (nunc == 0 && ncond == 1) return pass_kernel_v2<unsigned int, 13, true, 0, 1>;
            if (nunc == 1 && ncond == 0) return pass_kernel_v2<unsigned int, 13, true, 1, 0>;
            if (nunc == 1 && ncond == 1) return pass_kernel_v2<unsigned int, 13, true, 1, 1>;
            if (nunc == 2 && ncond == 0) return pass_kernel_v2<unsigned int, 13, true, 2, 0>;
            if (nunc == 2 && ncond == 1) return pass_kernel_v2<unsigned int, 13, true, 2, 1>;
            return nullptr;
        case 12: return later_pass<12>(nunc, ncond);
        case 11: return later_pass<11>(nunc, ncond);
        case 10: return later_pass<10>(nunc, ncond);
        case 9: return later_pass<9>(nunc, ncond);
        case 8: return later_pass<8>(nunc, ncond);
        case 7: return later_pass<7>(nunc, ncond);
        case 6: return later_pass<6>(nunc, ncond);
        case 5: return later_pass<5>(nunc, ncond);
        case 4: return later_pass<4>(nunc, ncond);
        default: return nullptr;
    }
}

PassKernel generic_pass_kernel(bool wide) {
    return wide ? pass_kernel_generic<unsigned long long> : pass_kernel_generic<unsigned int>;
}

}  // namespace qca
