// xq_act_l0.cuh -- the fixed-point layer-0 state of the acting path, shared by l0_act_kernel (xq_dqn_fast.cu: builds / repairs the
// sums) and the tail of act_team_kernel (xq_act_team.cu: updates them from the move it has just applied).
//
// z0 = b0 + sum of the rows of W0^T a board selects (src/chessai.cpp:268-289) is kept per env as rint(z0 * 2^k) in int32 (wrap-around
// arithmetic; the scale follows max |W0| per weight version, act_quant_*_kernel), so that "previous sum - rows that left + rows that
// entered" equals the gather over all rows bit for bit.  h = tanh(z0) leaves as BF16 hi + lo, the A operand of q90_gemm_kernel.
#pragma once
#include <cuda_bf16.h>

#include "xq_common.cuh"

namespace xq {

// what act_team_kernel needs to carry the sums across a ply (Z == nullptr: no carrying)
struct ActCarry {
    const int32_t* W0Q = nullptr;      // [(1260 + 1)][128] fixed-point W0^T, row 1260 = zeros
    const int32_t* zOpen = nullptr;    // [128] the sum of the opening position (ChessBoard::reset)
    const float* inv_scale = nullptr;  // 2^-k
    int32_t* Z = nullptr;              // [n][128] the sums
    uint32_t* Prev = nullptr;          // [n][12] the board words Z belongs to
    __nv_bfloat16* Hhi = nullptr;      // [n][128] h(s) hi / lo
    __nv_bfloat16* Hlo = nullptr;
};

__device__ __forceinline__ uint32_t act_sanitize(uint32_t w) {      // code 15 is not a piece (getStateRepresentation has no channel for it)
    const uint32_t t = w & (w >> 1) & (w >> 2) & (w >> 3) & 0x11111111u;
    return w & ~(t * 15u);
}

// tanh as 1 - 2 / (1 + 2^(2 log2(e) |z|)) on the SFU: absolute error ~1e-7, below the 2^-17 relative error of the hi + lo split
__device__ __forceinline__ float act_tanh(float z) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fabsf(z) * 2.8853900817779268f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    return copysignf(fmaf(-2.0f, r, 1.0f), z);
}

// h of 4 consecutive hidden units (uint4 number `idx4` of env e) from their fixed-point sums -> Hhi / Hlo
__device__ __forceinline__ void act_emit_h(const uint4& z, float is, __nv_bfloat16* __restrict__ Hhi, __nv_bfloat16* __restrict__ Hlo, int64_t e, int idx4) {
    const float h0 = act_tanh((float)(int)z.x * is), h1 = act_tanh((float)(int)z.y * is), h2 = act_tanh((float)(int)z.z * is), h3 = act_tanh((float)(int)z.w * is);
    const __nv_bfloat162 hi01 = __floats2bfloat162_rn(h0, h1), hi23 = __floats2bfloat162_rn(h2, h3);
    const __nv_bfloat162 lo01 = __floats2bfloat162_rn(h0 - __bfloat162float(hi01.x), h1 - __bfloat162float(hi01.y)),
                         lo23 = __floats2bfloat162_rn(h2 - __bfloat162float(hi23.x), h3 - __bfloat162float(hi23.y));
    uint2 ph, pl;
    ph.x = *reinterpret_cast<const uint32_t*>(&hi01); ph.y = *reinterpret_cast<const uint32_t*>(&hi23);
    pl.x = *reinterpret_cast<const uint32_t*>(&lo01); pl.y = *reinterpret_cast<const uint32_t*>(&lo23);
    reinterpret_cast<uint2*>(Hhi + e * 128)[idx4] = ph;
    reinterpret_cast<uint2*>(Hlo + e * 128)[idx4] = pl;
}

}  // namespace xq
