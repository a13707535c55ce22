// Shared helpers of the qca_b200 CUDA library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "qca_b200.h"

namespace qca {

void set_error(const char* fmt, ...);

#define QCA_CUDA(call)                                                                   \
    do {                                                                                 \
        cudaError_t _e = (call);                                                         \
        if (_e != cudaSuccess) {                                                         \
            qca::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,                 \
                           cudaGetErrorString(_e));                                      \
            return (_e == cudaErrorMemoryAllocation) ? QCA_ERR_NOMEM : QCA_ERR_CUDA;     \
        }                                                                                \
    } while (0)

#define QCA_CHECK(expr)                    \
    do {                                   \
        int32_t _r = (expr);               \
        if (_r != QCA_OK) return _r;       \
    } while (0)

#define QCA_REQUIRE(cond, code, ...)       \
    do {                                   \
        if (!(cond)) {                     \
            qca::set_error(__VA_ARGS__);   \
            return (code);                 \
        }                                  \
    } while (0)

// ---------------------------------------------------------------------------
// Rule evaluation.  For a configuration x (bit g <-> cell ncells-1-g) the
// returned word has bit g set iff the number of alive neighbours of that cell
// within `distance` lies in the activation interval (mpo.py:126-149 encode the
// same count test into the automaton).  Bits at or above ncells are garbage and
// must be masked by the caller.  Bit-sliced ripple counters: 2*distance shifted
// copies of x are summed into 4 one-bit planes (counts up to 15 -> distance <= 7).
// ---------------------------------------------------------------------------
template <typename I>
__host__ __device__ __forceinline__ I activity_word(I x, int distance, uint32_t interval_mask) {
    I c0 = 0, c1 = 0, c2 = 0, c3 = 0;
#pragma unroll 1
    for (int o = 1; o <= distance; ++o) {
        I s = x << o;
        I k0 = c0 & s;  c0 ^= s;
        I k1 = c1 & k0; c1 ^= k0;
        I k2 = c2 & k1; c2 ^= k1;
        c3 ^= k2;
        s = x >> o;
        k0 = c0 & s;  c0 ^= s;
        k1 = c1 & k0; c1 ^= k0;
        k2 = c2 & k1; c2 ^= k1;
        c3 ^= k2;
    }
    I act = 0;
    const int top = 2 * distance;
#pragma unroll 1
    for (int c = 0; c <= top; ++c) {
        if ((interval_mask >> c) & 1u) {
            I m = (c & 1) ? c0 : ~c0;
            m &= (c & 2) ? c1 : ~c1;
            m &= (c & 4) ? c2 : ~c2;
            m &= (c & 8) ? c3 : ~c3;
            act |= m;
        }
    }
    return act;
}

// How a rank-local amplitude index expands to the global basis-state index when the register is
// sharded: the sharded qubits sit at global bit positions pos[0] < pos[1] < ... (rank bit j <-> pos[j]);
// the local bits fill the remaining positions in order.
struct ShardMap {
    int nins;
    int pos[3];
    unsigned long long rank_or;  // this rank's values of the sharded qubits, already in place
};

__host__ __device__ __forceinline__ unsigned long long expand_index(unsigned long long x, const ShardMap& m) {
    for (int i = 0; i < m.nins; ++i) {
        const int p = m.pos[i];
        x = ((x >> p) << (p + 1)) | (x & ((1ull << p) - 1ull));
    }
    return x | m.rank_or;
}

// global bit position of local bit g
__host__ __device__ __forceinline__ int global_pos(int g, const ShardMap& m) {
    for (int i = 0; i < m.nins; ++i)
        if (g >= m.pos[i]) ++g;
    return g;
}

__host__ __device__ __forceinline__ uint32_t interval_mask_of(int lo, int hi) {
    uint32_t m = 0;
    for (int c = (lo < 0 ? 0 : lo); c < hi && c < 32; ++c) m |= (1u << c);
    return m;
}

}  // namespace qca
