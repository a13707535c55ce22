// Instantiations of the fast tile-pass kernel for 64-bit amplitude indices.
#include "qca_pass.cuh"

namespace qca {

template <int L>
static PassKernel later_pass(int nops) {
    switch (nops) {
        case 1: return pass_kernel_v2<unsigned long long, L, false, 1>;
        case 2: return pass_kernel_v2<unsigned long long, L, false, 2>;
        case 3: return pass_kernel_v2<unsigned long long, L, false, 3>;
        case 4: return pass_kernel_v2<unsigned long long, L, false, 4>;
        default: return nullptr;
    }
}

PassKernel fast_pass_kernel_u64(int low_bits, int nstreams) {
    switch (low_bits) {
        case 13:
            switch (nstreams) {
                case 0: return pass_kernel_v2<unsigned long long, 13, true, 0>;
                case 1: return pass_kernel_v2<unsigned long long, 13, true, 1>;
                case 2: return pass_kernel_v2<unsigned long long, 13, true, 2>;
                default: return nullptr;
            }
        case 12: return later_pass<12>(nstreams);
        case 11: return later_pass<11>(nstreams);
        case 10: return later_pass<10>(nstreams);
        case 9: return later_pass<9>(nstreams);
        case 8: return later_pass<8>(nstreams);
        case 7: return later_pass<7>(nstreams);
        case 6: return later_pass<6>(nstreams);
        case 5: return later_pass<5>(nstreams);
        case 4: return later_pass<4>(nstreams);
        default: return nullptr;
    }
}

}  // namespace qca
