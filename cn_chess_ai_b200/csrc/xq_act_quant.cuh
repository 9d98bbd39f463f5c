// xq_act_quant.cuh -- the scale of the acting path's fixed-point layer-0 table (xq_act_l0.cuh, act_quant_kernel in xq_dqn_fast.cu).
//
// k is chosen per weight version from the largest |W0|, |b0| so that a sum of 92 terms (90 squares + bias + slack) stays below 2^30 in
// magnitude: W0Q = rint(W0 * 2^k) in int32, z0 = sum * 2^-k.  Host-compilable (tests/hostsim) so that the CPU suite can check the range and
// the accuracy of the fixed-point sum against FP64 over weight scales from 1e-30 to 1e30 and the non-finite corner cases.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "xq_rules.cuh"

namespace xq {

// max_bits = the IEEE bits of max |w| (sign cleared; a NaN sorts above inf).  Returns k in [-100, 124]: 2^k and 2^-k are normal floats.
XQ_HD int act_quant_shift(uint32_t max_bits) {
    float m;
#if defined(__CUDA_ARCH__)
    m = __uint_as_float(max_bits);
#else
    memcpy(&m, &max_bits, sizeof m);
#endif
    const float top = 92.0f * m;
    int e = 0;
    if (top > 0.0f && top < INFINITY) frexpf(top, &e);      // top < 2^e
    else e = (top > 0.0f || top != top) ? 130 : -60;          // inf / NaN: the smallest scale; all-zero weights: any scale will do
    const int k = 30 - e;                                     // |sum| * 2^k < 2^30
    return k < -100 ? -100 : (k > 124 ? 124 : k);
}

// one table entry; out-of-range products (only with non-finite or > 2^120 weights) saturate
XQ_HD int32_t act_quantize(float w, int k) {
    const float v = w * ldexpf(1.0f, k);
#if defined(__CUDA_ARCH__)
    return __float2int_rn(v);
#else
    if (v != v) return 0;
    if (v >= 2147483648.0f) return INT32_MAX;
    if (v <= -2147483648.0f) return INT32_MIN;
    return (int32_t)lrintf(v);
#endif
}

}  // namespace xq
