// xq_dqn_fast.cu -- batched tensor-core path of the {1260,128,8100} self-play Q-network.
//
// What the reference does per ply (src/chessai.cpp:106-131): up to 3 forwards + 1 backward of a dense
// FP64 MLP at batch 1.  Here a TD update runs on a batch of B transitions (s, a, r, s', done):
//   layer 0   h = tanh(W0 x + b0): x is one-hot with <= 32 ones (src/chessai.cpp:268-289), so the
//             pre-activation is a gather-sum of <= 32 rows of W0^T read straight from the packed board
//             (FP32, exact; the 1260-wide one-hot vector is never materialised)            [l0_forward_kernel]
//   layer 1   z = W1 h + b1 over ALL 8100 outputs is the one dense contraction:
//             [B x 128] x [128 x 8100] on tcgen05 tensor cores, BF16 operands staged by TMA (128-B swizzle),
//             FP32 accumulators in TMEM, W1 tile stationary in shared memory, epilogue warps reduce the
//             row max straight out of TMEM (max_a tanh(z_a) = tanh(max_a z_a): no Q matrix, no 33M tanh)   [l1_gemm_kernel]
//   TD error  the live loop's target equals Q(s) except at index `to` (src/chessai.cpp:122-128), so delta1 is
//             one-hot per sample: q(s)[to] is a 128-long dot product, dW1 touches row `to` only, delta0 is one
//             row of W1 (as written: src/dqn.cu:406-423; or corrected)            [td_delta_kernel]
//   dW0       sum_b delta0_b (x) x_b accumulated per board square in shared memory            [dw0_kernel]
//   SGD       W -= lr * sum of per-sample gradients (B = 1 reproduces one reference step)     [apply_kernel]
// FP32 master weights; BF16 only as MMA operands.  Tolerance vs the FP64 oracle: |dQ| <= 2e-3.
#include <cuda.h>
#include <math.h>

#include "xq_dqn_internal.cuh"
#include "xq_tc.cuh"

namespace xq {

constexpr int kIn = XQ_STATE_SIZE, kHid = 128, kOut = 8100, kQRows = 90;   // Q is indexed by `to` < 90 (src/dqn.cpp:47)
constexpr int kGradW0 = 0, kGradB0 = kIn * kHid, kGradW1 = kGradB0 + kHid, kGradB1 = kGradW1 + kQRows * kHid;
constexpr int kGradSize = kGradB1 + kQRows;   // 173,018 floats: the only non-zero gradient entries of a TD step

// ---- GEMM tile configuration ----------------------------------------------------------------
constexpr int BM = 128;            // rows (samples) per MMA = TMEM lanes
constexpr int BN = 224;            // outputs per CTA tile: 37 tiles x 4 row splits = 148 CTAs = one per SM
constexpr int BK = 64;             // BF16 elements per 128-byte swizzle row
constexpr int kKBlocks = kHid / BK;                 // 2
constexpr int kAStages = 3;
constexpr int kNTiles = (kOut + BN - 1) / BN;       // 37
constexpr uint32_t kABytes = BM * kHid * 2;         // 32 KB: one A tile (both k-blocks)
constexpr uint32_t kBBytes = BN * kHid * 2;         // 56 KB: the stationary W1 tile
constexpr uint32_t kTmemCols = 512;                 // 2 accumulator stages x BN columns (power of two >= 448)
constexpr int kGemmThreads = 256;                   // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 bias, 4-7 epilogue
constexpr size_t kGemmSmem = 1024 + kBBytes + kAStages * kABytes + BN * 4 + 256;

struct Fast {
    int64_t cap = 0;                                   // batch capacity of the workspace
    float *W0T = nullptr, *b0 = nullptr, *W1 = nullptr, *b1 = nullptr;          // online, FP32 master
    __nv_bfloat16* W1bf = nullptr;
    float *tW0T = nullptr, *tb0 = nullptr, *tW1 = nullptr, *tb1 = nullptr;      // target network
    __nv_bfloat16* tW1bf = nullptr;
    bool target_current = false;
    float* grad = nullptr;                             // kGradSize
    // batch workspace
    xq_env_rec* boards = nullptr;                      // staging for xq_dqn_forward_boards
    __nv_bfloat16 *Hbf = nullptr, *H2bf = nullptr;     // h(s), h(s') as MMA A operands [cap][128]
    float* Hf = nullptr;                               // h(s) FP32 [cap][128]
    float* zpart = nullptr;                            // [kNTiles][cap] row-max partials
    float* delta0 = nullptr;                           // [cap][128]
    float* q = nullptr;                                // [cap][8100] (debug path only, allocated on demand)
    int64_t q_cap = 0;
    float* info = nullptr;                             // 4 floats
    CUtensorMap tmW1, tmTW1, tmH, tmH2;
    int64_t tm_rows = 0;
};

// ---------------------------------------------------------------------------------------------
// layer 0: one warp per board, lane owns 4 hidden units; <= 32 coalesced 512-byte row reads of W0^T
__global__ void __launch_bounds__(256) l0_forward_kernel(const uint8_t* __restrict__ boards, int64_t stride_bytes, int64_t n,
                                                        const float* __restrict__ W0T, const float* __restrict__ b0,
                                                        __nv_bfloat16* __restrict__ Hbf, float* __restrict__ Hf) {
    const int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (s >= n) return;
    const uint32_t* w = reinterpret_cast<const uint32_t*>(boards + s * stride_bytes);
    const uint32_t word = lane < 12 ? w[lane] : 0u;
    float4 acc = reinterpret_cast<const float4*>(b0)[lane];
    for (int wi = 0; wi < 12; ++wi) {
        uint32_t v = __shfl_sync(0xFFFFFFFFu, word, wi);
        while (v) {                                         // warp-uniform: every lane sees the same board
            const int nib = (__ffs((int)v) - 1) >> 2;
            const int code = (v >> (4 * nib)) & 15;
            v &= ~(15u << (4 * nib));
            const int row = (wi * 8 + nib) * 14 + code - 1;      // getStateRepresentation index, src/chessai.cpp:278-282
            if (code < 15 && row < kIn) {
                const float4 r = reinterpret_cast<const float4*>(W0T + (size_t)row * kHid)[lane];
                acc.x += r.x; acc.y += r.y; acc.z += r.z; acc.w += r.w;
            }
        }
    }
    const float4 h = make_float4(tanhf(acc.x), tanhf(acc.y), tanhf(acc.z), tanhf(acc.w));
    if (Hf) reinterpret_cast<float4*>(Hf + s * kHid)[lane] = h;
    __nv_bfloat162 lo = __floats2bfloat162_rn(h.x, h.y), hi = __floats2bfloat162_rn(h.z, h.w);
    uint2 packed;
    packed.x = *reinterpret_cast<uint32_t*>(&lo); packed.y = *reinterpret_cast<uint32_t*>(&hi);
    reinterpret_cast<uint2*>(Hbf + s * kHid)[lane] = packed;
}

// ---------------------------------------------------------------------------------------------
// layer 1 on tcgen05: Z[M x 8100] = H[M x 128] * W1[8100 x 128]^T (+ b1), W1 tile stationary.
// CTA (n_tile, split): loads its [BN x 128] BF16 slice of W1 once, then streams the A tiles of its row range
// through a 3-stage TMA ring; one elected thread issues 8 UMMA (M128 x N224 x K16) per tile into one of two
// TMEM accumulator stages; 4 epilogue warps (one per TMEM lane quarter) drain the other stage meanwhile.
enum { EPI_ROWMAX = 0, EPI_STORE_TANH = 1 };

template <int MODE>
__global__ void __launch_bounds__(kGemmThreads, 1) l1_gemm_kernel(const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmW1,
                                                                 const float* __restrict__ b1, int M, int m_tiles, int n_splits,
                                                                 float* __restrict__ zpart, int64_t zstride, float* __restrict__ Q) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);   // SW128 tiles need 1024-B alignment
    uint8_t* sB = smem;                                    // [kKBlocks][BN][64] bf16
    uint8_t* sA = smem + kBBytes;                          // [kAStages][kKBlocks][BM][64] bf16
    float* sBias = reinterpret_cast<float*>(sA + kAStages * kABytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + BN);
    uint64_t* b_full = bars;              // 1
    uint64_t* a_full = bars + 1;          // kAStages
    uint64_t* a_empty = bars + 4;         // kAStages
    uint64_t* acc_full = bars + 7;        // 2
    uint64_t* acc_empty = bars + 9;       // 2
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tile = blockIdx.x, split = blockIdx.y;
    const int n0 = n_tile * BN;
    // rows of this CTA: m-tiles split, split + n_splits, ...
    const int my_tiles = (m_tiles - split + n_splits - 1) / n_splits;

    if (threadIdx.x == 0) {
        tc::mbar_init(b_full, 1);
        for (int i = 0; i < kAStages; ++i) { tc::mbar_init(a_full + i, 1); tc::mbar_init(a_empty + i, 1); }
        for (int i = 0; i < 2; ++i) { tc::mbar_init(acc_full + i, 1); tc::mbar_init(acc_empty + i, 4); }
        tc::fence_barrier_init();
    }
    if (warp == 2) tc::tmem_alloc<kTmemCols>(tmem_slot);
    if (warp == 3) for (int i = lane; i < BN; i += 32) sBias[i] = (n0 + i < kOut) ? b1[n0 + i] : -INFINITY;   // padded outputs never win the max
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {   // ===== TMA producer =====
            tc::prefetch_tmap(&tmH); tc::prefetch_tmap(&tmW1);
            tc::mbar_expect_tx(b_full, kBBytes);
            for (int kb = 0; kb < kKBlocks; ++kb) tc::tma_load_2d(sB + kb * (BN * BK * 2), &tmW1, kb * BK, n0, b_full);
            for (int i = 0; i < my_tiles; ++i) {
                const int st = i % kAStages;
                tc::mbar_wait(a_empty + st, ((i / kAStages) & 1) ^ 1);
                tc::mbar_expect_tx(a_full + st, kABytes);
                const int row0 = (split + i * n_splits) * BM;
                for (int kb = 0; kb < kKBlocks; ++kb) tc::tma_load_2d(sA + st * kABytes + kb * (BM * BK * 2), &tmH, kb * BK, row0, a_full + st);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {   // ===== MMA issuer =====
            constexpr uint32_t idesc = tc::umma_idesc_bf16(BM, BN);
            tc::mbar_wait(b_full, 0);
            for (int i = 0; i < my_tiles; ++i) {
                const int st = i % kAStages, acc = i & 1;
                tc::mbar_wait(acc_empty + acc, ((i >> 1) & 1) ^ 1);
                tc::mbar_wait(a_full + st, (i / kAStages) & 1);
                tc::tc_fence_after();
#pragma unroll
                for (int kb = 0; kb < kKBlocks; ++kb)
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t da = tc::umma_desc_sw128(tc::smem_u32(sA + st * kABytes + kb * (BM * BK * 2)) + k * 32);
                        const uint64_t db = tc::umma_desc_sw128(tc::smem_u32(sB + kb * (BN * BK * 2)) + k * 32);
                        tc::umma_bf16(tmem_base + acc * BN, da, db, idesc, (kb | k) != 0);
                    }
                tc::umma_commit(a_empty + st);      // A stage free once these MMAs have read it
                tc::umma_commit(acc_full + acc);    // accumulator ready for the epilogue
            }
        }
    } else if (warp >= 4) {   // ===== epilogue: warp w drains TMEM lanes 32*(w-4) .. +31 =====
        const int quarter = warp - 4;
        for (int i = 0; i < my_tiles; ++i) {
            const int acc = i & 1;
            const int row = (split + i * n_splits) * BM + quarter * 32 + lane;
            tc::mbar_wait(acc_full + acc, (i >> 1) & 1);
            tc::tc_fence_after();
            float best = -INFINITY;
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                float v[32];
                tc::tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + c * 32, v);
                if (MODE == EPI_ROWMAX) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) best = fmaxf(best, v[j] + sBias[c * 32 + j]);
                } else if (row < M) {
                    float* out = Q + (size_t)row * kOut + n0 + c * 32;
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (n0 + c * 32 + j < kOut) out[j] = tanhf(v[j] + sBias[c * 32 + j]);
                }
            }
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(acc_empty + acc);
            if (MODE == EPI_ROWMAX && row < M) zpart[(int64_t)n_tile * zstride + row] = best;
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) { tc::tc_fence_after(); tc::tmem_dealloc<kTmemCols>(tmem_base); }
}

// ---------------------------------------------------------------------------------------------
// Transition record of the replay buffer / TD batch (include/xq.h: xq_transition, 128 B)
struct Transition {
    uint32_t s[12], s2[12];
    uint16_t action; uint8_t mover, done;
    int32_t reward;
    uint32_t pad[6];
};
static_assert(sizeof(Transition) == 128 && sizeof(xq_transition) == 128, "transition record must be 128 bytes");

// TD error, one warp per transition (src/chessai.cpp:121-131 + src/dqn.cu:288-308 specialised to a one-hot delta1)
__global__ void __launch_bounds__(256) td_delta_kernel(const Transition* __restrict__ batch, int64_t n, const float* __restrict__ Hf,
                                                      const float* __restrict__ W1, const float* __restrict__ b1,
                                                      const float* __restrict__ zpart, int64_t zstride, int n_tiles, float gamma, int mode,
                                                      float* __restrict__ delta0, float* __restrict__ grad, float* __restrict__ info) {
    const int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (s >= n) return;
    const Transition* t = batch + s;
    const int to = XQ_ACTION_TO(t->action);                // the Q index of the taken action is action.to (:124,:127)
    const float4 h = reinterpret_cast<const float4*>(Hf + s * kHid)[lane];
    const float4 w = reinterpret_cast<const float4*>(W1 + (size_t)to * kHid)[lane];
    float z = h.x * w.x + h.y * w.y + h.z * w.z + h.w * w.w;
    float zmax = -INFINITY;
    if (!t->done) for (int i = lane; i < n_tiles; i += 32) zmax = fmaxf(zmax, zpart[(int64_t)i * zstride + s]);
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) { z += __shfl_xor_sync(0xFFFFFFFFu, z, k); zmax = fmaxf(zmax, __shfl_xor_sync(0xFFFFFFFFu, zmax, k)); }
    const float q = tanhf(z + b1[to]);
    const float target = t->done ? (float)t->reward : (float)t->reward + gamma * tanhf(zmax);
    const float d1 = (q - target) * (1.0f - q * q);        // outputLayerDeltaKernel, src/dqn.cu:288-295
    // hidden delta: as written W1flat[i*1260 + j] with i = to (only i < 128 is summed and delta1 is one-hot), or W1[to][j]
    const float* wrow = mode == XQ_DQN_AS_WRITTEN ? W1 + (size_t)to * kIn : W1 + (size_t)to * kHid;
    const float4 wd = reinterpret_cast<const float4*>(wrow)[lane];
    float4 d0;
    d0.x = wd.x * d1 * (1.0f - h.x * h.x); d0.y = wd.y * d1 * (1.0f - h.y * h.y);
    d0.z = wd.z * d1 * (1.0f - h.z * h.z); d0.w = wd.w * d1 * (1.0f - h.w * h.w);
    reinterpret_cast<float4*>(delta0 + s * kHid)[lane] = d0;
    float* gw1 = grad + kGradW1 + to * kHid + 4 * lane;    // dW1[to][:] += delta1 * h   (updateWeightsBiasesKernel, :310-319)
    atomicAdd(gw1 + 0, d1 * h.x); atomicAdd(gw1 + 1, d1 * h.y); atomicAdd(gw1 + 2, d1 * h.z); atomicAdd(gw1 + 3, d1 * h.w);
    if (lane == 0) {
        atomicAdd(grad + kGradB1 + to, d1);
        atomicAdd(info + 0, 0.5f * (q - target) * (q - target));
        atomicAdd(info + 1, q);
        atomicAdd(info + 2, target);
    }
}

// dW0^T[feature][:] += delta0_b for every set feature of x_b, and db0 += delta0_b.
// CTA (square q in 0..90, chunk c): 128 threads = hidden units; the 14 feature rows of square q (or the bias
// row for the pseudo-square 90) are accumulated in registers over the chunk's samples without atomics, then
// merged into the gradient with one atomicAdd per non-zero (row, unit).
constexpr int kDw0Chunk = 256;
__global__ void __launch_bounds__(kHid) dw0_kernel(const Transition* __restrict__ batch, int64_t n, const float* __restrict__ delta0,
                                                  float* __restrict__ grad) {
    __shared__ uint8_t s_code[kDw0Chunk];
    const int q = blockIdx.x, j = threadIdx.x;
    const int64_t b0 = (int64_t)blockIdx.y * kDw0Chunk;
    const int cnt = (int)min((int64_t)kDw0Chunk, n - b0);
    for (int i = j; i < cnt; i += kHid) s_code[i] = q == 90 ? 1 : (uint8_t)((batch[b0 + i].s[q >> 3] >> (4 * (q & 7))) & 15);
    __syncthreads();
    float acc[14];
#pragma unroll
    for (int c = 0; c < 14; ++c) acc[c] = 0.0f;
    for (int i = 0; i < cnt; ++i) {
        const int code = s_code[i];                         // block-uniform
        if (code == 0 || code == 15) continue;
        const float d = delta0[(b0 + i) * kHid + j];
#pragma unroll
        for (int c = 0; c < 14; ++c) if (code == c + 1) acc[c] += d;
    }
    if (q == 90) { if (acc[0] != 0.0f) atomicAdd(grad + kGradB0 + j, acc[0]); return; }
#pragma unroll
    for (int c = 0; c < 14; ++c)
        if (acc[c] != 0.0f) atomicAdd(grad + kGradW0 + (size_t)(q * 14 + c) * kHid + j, acc[c]);
}

// SGD: W -= lr * grad on the compact gradient (src/dqn.cu:310-319), refresh the BF16 operand rows, clear the gradient
__global__ void __launch_bounds__(256) apply_kernel(float* __restrict__ W0T, float* __restrict__ b0, float* __restrict__ W1, float* __restrict__ b1,
                                                   __nv_bfloat16* __restrict__ W1bf, float* __restrict__ grad, float lr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kGradSize) return;
    const float g = grad[i];
    grad[i] = 0.0f;
    if (g == 0.0f) return;
    if (i < kGradB0) W0T[i] -= lr * g;
    else if (i < kGradW1) b0[i - kGradB0] -= lr * g;
    else if (i < kGradB1) { const int e = i - kGradW1; const float v = W1[e] - lr * g; W1[e] = v; W1bf[e] = __float2bfloat16_rn(v); }
    else b1[i - kGradB1] -= lr * g;
}

// ---- FP64 (reference layout) <-> fast-path layouts ----------------------------------------------
__global__ void f64_to_fast_kernel(const double* __restrict__ w, const double* __restrict__ b, float* __restrict__ W0T, float* __restrict__ b0,
                                   float* __restrict__ W1, float* __restrict__ b1, __nv_bfloat16* __restrict__ W1bf) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (int64_t)kIn * kHid) { const int o = (int)(i / kIn), in = (int)(i % kIn); W0T[(size_t)in * kHid + o] = (float)w[i]; }   // W0[o][in] -> W0T[in][o]
    if (i < (int64_t)kOut * kHid) { const float v = (float)w[(size_t)kIn * kHid + i]; W1[i] = v; W1bf[i] = __float2bfloat16_rn(v); }
    if (i < kHid) b0[i] = (float)b[i];
    if (i < kOut) b1[i] = (float)b[kHid + i];
}
__global__ void fast_to_f64_kernel(double* __restrict__ w, double* __restrict__ b, const float* __restrict__ W0T, const float* __restrict__ b0,
                                   const float* __restrict__ W1, const float* __restrict__ b1) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (int64_t)kIn * kHid) { const int o = (int)(i / kIn), in = (int)(i % kIn); w[i] = (double)W0T[(size_t)in * kHid + o]; }
    if (i < (int64_t)kOut * kHid) w[(size_t)kIn * kHid + i] = (double)W1[i];
    if (i < kHid) b[i] = (double)b0[i];
    if (i < kOut) b[kHid + i] = (double)b1[i];
}

// ---- host side ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}
// [rows][128] BF16 row-major, box = 64 columns x box_rows, 128-byte swizzle, out-of-range rows read as zero
static int make_tmap(CUtensorMap* m, const void* base, int64_t rows, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(XQ_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[2] = {(cuuint64_t)kHid, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)kHid * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(XQ_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return XQ_OK;
}

static inline unsigned blocks(int64_t n, int per) { return (unsigned)((n + per - 1) / per); }

void dqn_fast_destroy(xq_dqn_s* h) {
    Fast* f = h->fast;
    if (!f) return;
    cudaFree(f->W0T); cudaFree(f->b0); cudaFree(f->W1); cudaFree(f->b1); cudaFree(f->W1bf);
    cudaFree(f->tW0T); cudaFree(f->tb0); cudaFree(f->tW1); cudaFree(f->tb1); cudaFree(f->tW1bf);
    cudaFree(f->grad); cudaFree(f->boards); cudaFree(f->Hbf); cudaFree(f->H2bf); cudaFree(f->Hf); cudaFree(f->zpart);
    cudaFree(f->delta0); cudaFree(f->q); cudaFree(f->info);
    delete f;
    h->fast = nullptr;
}

static int fast_init(xq_dqn_s* h) {
    if (h->fast) return XQ_OK;
    if (h->layers.size() != 3 || h->layers[0] != kIn || h->layers[1] != kHid || h->layers[2] != kOut)
        return fail(XQ_ERR_INVALID, "the batched tensor-core path is specialised to the {1260,128,8100} network (src/chessai.cpp:395-404)");
    Fast* f = new (std::nothrow) Fast();
    if (!f) return fail(XQ_ERR_NOMEM, "out of host memory");
    h->fast = f;
    XQ_CUDA(cudaMalloc(&f->W0T, sizeof(float) * kIn * kHid)); XQ_CUDA(cudaMalloc(&f->b0, sizeof(float) * kHid));
    XQ_CUDA(cudaMalloc(&f->W1, sizeof(float) * kOut * kHid)); XQ_CUDA(cudaMalloc(&f->b1, sizeof(float) * kOut));
    XQ_CUDA(cudaMalloc(&f->W1bf, sizeof(__nv_bfloat16) * kOut * kHid));
    XQ_CUDA(cudaMalloc(&f->tW0T, sizeof(float) * kIn * kHid)); XQ_CUDA(cudaMalloc(&f->tb0, sizeof(float) * kHid));
    XQ_CUDA(cudaMalloc(&f->tW1, sizeof(float) * kOut * kHid)); XQ_CUDA(cudaMalloc(&f->tb1, sizeof(float) * kOut));
    XQ_CUDA(cudaMalloc(&f->tW1bf, sizeof(__nv_bfloat16) * kOut * kHid));
    XQ_CUDA(cudaMalloc(&f->grad, sizeof(float) * kGradSize)); XQ_CUDA(cudaMalloc(&f->info, sizeof(float) * 4));
    XQ_CUDA(cudaMemsetAsync(f->grad, 0, sizeof(float) * kGradSize, h->stream));
    if (int rc = make_tmap(&f->tmW1, f->W1bf, kOut, BN)) return rc;
    if (int rc = make_tmap(&f->tmTW1, f->tW1bf, kOut, BN)) return rc;
    XQ_CUDA(cudaFuncSetAttribute(l1_gemm_kernel<EPI_ROWMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem));
    XQ_CUDA(cudaFuncSetAttribute(l1_gemm_kernel<EPI_STORE_TANH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem));
    return XQ_OK;
}

static int fast_reserve(xq_dqn_s* h, int64_t n) {
    Fast* f = h->fast;
    if (n <= f->cap) return XQ_OK;
    cudaFree(f->boards); cudaFree(f->Hbf); cudaFree(f->H2bf); cudaFree(f->Hf); cudaFree(f->zpart); cudaFree(f->delta0);
    f->boards = nullptr; f->Hbf = f->H2bf = nullptr; f->Hf = f->zpart = f->delta0 = nullptr; f->cap = 0;
    const int64_t rows = (n + BM - 1) / BM * BM;
    XQ_CUDA(cudaMalloc(&f->boards, sizeof(Transition) * n));
    XQ_CUDA(cudaMalloc(&f->Hbf, sizeof(__nv_bfloat16) * rows * kHid)); XQ_CUDA(cudaMalloc(&f->H2bf, sizeof(__nv_bfloat16) * rows * kHid));
    XQ_CUDA(cudaMalloc(&f->Hf, sizeof(float) * rows * kHid)); XQ_CUDA(cudaMalloc(&f->zpart, sizeof(float) * kNTiles * rows));
    XQ_CUDA(cudaMalloc(&f->delta0, sizeof(float) * rows * kHid));
    if (int rc = make_tmap(&f->tmH, f->Hbf, n, BM)) return rc;
    if (int rc = make_tmap(&f->tmH2, f->H2bf, n, BM)) return rc;
    f->cap = n; f->tm_rows = n;
    return XQ_OK;
}

// the fast path's copies follow the FP64 parameters whenever those were modified last
static int ensure_fast(xq_dqn_s* h) {
    if (int rc = fast_init(h)) return rc;
    Fast* f = h->fast;
    if (!h->fast_current) {
        f64_to_fast_kernel<<<blocks((int64_t)kOut * kHid, 256), 256, 0, h->stream>>>(h->d_w, h->d_b, f->W0T, f->b0, f->W1, f->b1, f->W1bf);
        XQ_LAUNCH_CHECK();
        h->fast_current = true;
    }
    if (!f->target_current) {
        f64_to_fast_kernel<<<blocks((int64_t)kOut * kHid, 256), 256, 0, h->stream>>>(h->d_tw, h->d_tb, f->tW0T, f->tb0, f->tW1, f->tb1, f->tW1bf);
        XQ_LAUNCH_CHECK();
        f->target_current = true;
    }
    return XQ_OK;
}

int dqn_ensure_f64(xq_dqn_s* h) {
    if (h->f64_current || !h->fast) return XQ_OK;
    Fast* f = h->fast;
    fast_to_f64_kernel<<<blocks((int64_t)kOut * kHid, 256), 256, 0, h->stream>>>(h->d_w, h->d_b, f->W0T, f->b0, f->W1, f->b1);
    XQ_LAUNCH_CHECK();
    h->f64_current = true;
    return XQ_OK;
}
int dqn_fast_weights(xq_dqn_s* h, FastWeights* out) {
    if (int rc = ensure_fast(h)) return rc;
    out->W0T = h->fast->W0T; out->b0 = h->fast->b0; out->W1 = h->fast->W1; out->b1 = h->fast->b1;
    return XQ_OK;
}
void dqn_target_changed(xq_dqn_s* h) { if (h->fast) h->fast->target_current = false; }

static int launch_gemm(xq_dqn_s* h, int mode, const CUtensorMap& tmA, const CUtensorMap& tmB, const float* b1, int64_t n, float* q) {
    Fast* f = h->fast;
    const int m_tiles = (int)((n + BM - 1) / BM);
    int n_splits = 148 / kNTiles;                  // 4 row splits x 37 column tiles = 148 CTAs
    if (n_splits > m_tiles) n_splits = m_tiles;
    const dim3 grid(kNTiles, n_splits);
    const int64_t zstride = (f->cap + BM - 1) / BM * BM;
    if (mode == EPI_ROWMAX)
        l1_gemm_kernel<EPI_ROWMAX><<<grid, kGemmThreads, kGemmSmem, h->stream>>>(tmA, tmB, b1, (int)n, m_tiles, n_splits, f->zpart, zstride, nullptr);
    else
        l1_gemm_kernel<EPI_STORE_TANH><<<grid, kGemmThreads, kGemmSmem, h->stream>>>(tmA, tmB, b1, (int)n, m_tiles, n_splits, nullptr, zstride, q);
    XQ_LAUNCH_CHECK();
    return XQ_OK;
}

}  // namespace xq

using namespace xq;

#define XQ_DQN_ENTER(h)                                                      \
    if (!(h)) return fail(XQ_ERR_INVALID, "%s: null handle", __func__);      \
    XQ_CUDA(cudaSetDevice((h)->device))

extern "C" {

int xq_dqn_forward_boards(xq_dqn_t h, const xq_env_rec* boards_host, int64_t n, float* q_host) {
    XQ_DQN_ENTER(h);
    if (!boards_host || !q_host || n <= 0) return fail(XQ_ERR_INVALID, "xq_dqn_forward_boards: bad arguments");
    if (int rc = ensure_fast(h)) return rc;
    if (int rc = fast_reserve(h, n)) return rc;
    Fast* f = h->fast;
    if (n > f->q_cap) { cudaFree(f->q); f->q = nullptr; f->q_cap = 0; XQ_CUDA(cudaMalloc(&f->q, sizeof(float) * n * kOut)); f->q_cap = n; }
    XQ_CUDA(cudaMemcpyAsync(f->boards, boards_host, sizeof(xq_env_rec) * n, cudaMemcpyHostToDevice, h->stream));
    l0_forward_kernel<<<blocks(n * 32, 256), 256, 0, h->stream>>>(reinterpret_cast<const uint8_t*>(f->boards), sizeof(xq_env_rec), n, f->W0T, f->b0, f->Hbf, nullptr);
    XQ_LAUNCH_CHECK();
    if (n != f->tm_rows) { if (int rc = make_tmap(&f->tmH, f->Hbf, n, BM)) return rc; if (int rc = make_tmap(&f->tmH2, f->H2bf, n, BM)) return rc; f->tm_rows = n; }
    if (int rc = launch_gemm(h, EPI_STORE_TANH, f->tmH, f->tmW1, f->b1, n, f->q)) return rc;
    XQ_CUDA(cudaMemcpyAsync(q_host, f->q, sizeof(float) * n * kOut, cudaMemcpyDeviceToHost, h->stream));
    XQ_CUDA(cudaStreamSynchronize(h->stream));
    return XQ_OK;
}

// device-resident TD update on n transitions already in device memory
int xq_dqn_td_update_device(xq_dqn_t h, const void* batch_dev, int64_t n, int use_target_net, double lr, int apply) {
    XQ_DQN_ENTER(h);
    if (!batch_dev || n <= 0) return fail(XQ_ERR_INVALID, "xq_dqn_td_update_device: bad arguments");
    if (int rc = ensure_fast(h)) return rc;
    if (int rc = fast_reserve(h, n)) return rc;
    Fast* f = h->fast;
    if (n != f->tm_rows) { if (int rc = make_tmap(&f->tmH, f->Hbf, n, BM)) return rc; if (int rc = make_tmap(&f->tmH2, f->H2bf, n, BM)) return rc; f->tm_rows = n; }
    const Transition* batch = reinterpret_cast<const Transition*>(batch_dev);
    const uint8_t* base = reinterpret_cast<const uint8_t*>(batch);
    if (lr <= 0) lr = h->lr;
    // h(s) with the online net; h(s') with the online (ChessAI::train) or target (DQN::train) net
    l0_forward_kernel<<<blocks(n * 32, 256), 256, 0, h->stream>>>(base, sizeof(Transition), n, f->W0T, f->b0, f->Hbf, f->Hf);
    XQ_LAUNCH_CHECK();
    l0_forward_kernel<<<blocks(n * 32, 256), 256, 0, h->stream>>>(base + 48, sizeof(Transition), n, use_target_net ? f->tW0T : f->W0T,
                                                                  use_target_net ? f->tb0 : f->b0, f->H2bf, nullptr);
    XQ_LAUNCH_CHECK();
    if (int rc = launch_gemm(h, EPI_ROWMAX, f->tmH2, use_target_net ? f->tmTW1 : f->tmW1, use_target_net ? f->tb1 : f->b1, n, nullptr)) return rc;
    XQ_CUDA(cudaMemsetAsync(f->info, 0, sizeof(float) * 4, h->stream));
    const int64_t zstride = (f->cap + BM - 1) / BM * BM;
    td_delta_kernel<<<blocks(n * 32, 256), 256, 0, h->stream>>>(batch, n, f->Hf, f->W1, f->b1, f->zpart, zstride, kNTiles, (float)h->gamma, h->mode,
                                                                f->delta0, f->grad, f->info);
    XQ_LAUNCH_CHECK();
    dw0_kernel<<<dim3(91, blocks(n, kDw0Chunk)), kHid, 0, h->stream>>>(batch, n, f->delta0, f->grad);
    XQ_LAUNCH_CHECK();
    if (apply) {
        apply_kernel<<<blocks(kGradSize, 256), 256, 0, h->stream>>>(f->W0T, f->b0, f->W1, f->b1, f->W1bf, f->grad, (float)lr);
        XQ_LAUNCH_CHECK();
        h->f64_current = false;
    }
    return XQ_OK;
}

int xq_dqn_td_update(xq_dqn_t h, const xq_transition* batch_host, int64_t n, int use_target_net, double lr, float* info_host) {
    XQ_DQN_ENTER(h);
    if (!batch_host || n <= 0) return fail(XQ_ERR_INVALID, "xq_dqn_td_update: bad arguments");
    if (int rc = ensure_fast(h)) return rc;
    if (int rc = fast_reserve(h, n)) return rc;
    Fast* f = h->fast;
    XQ_CUDA(cudaMemcpyAsync(f->boards, batch_host, sizeof(Transition) * n, cudaMemcpyHostToDevice, h->stream));
    if (int rc = xq_dqn_td_update_device(h, f->boards, n, use_target_net, lr, 1)) return rc;
    if (info_host) XQ_CUDA(cudaMemcpyAsync(info_host, f->info, sizeof(float) * 4, cudaMemcpyDeviceToHost, h->stream));
    XQ_CUDA(cudaStreamSynchronize(h->stream));
    return XQ_OK;
}

int xq_dqn_grad_buffer(xq_dqn_t h, void** dev_ptr, int64_t* n_floats) {
    XQ_DQN_ENTER(h);
    if (int rc = ensure_fast(h)) return rc;
    if (dev_ptr) *dev_ptr = h->fast->grad;
    if (n_floats) *n_floats = kGradSize;
    return XQ_OK;
}

int xq_dqn_apply_grads(xq_dqn_t h, double lr) {
    XQ_DQN_ENTER(h);
    if (int rc = ensure_fast(h)) return rc;
    Fast* f = h->fast;
    if (lr <= 0) lr = h->lr;
    apply_kernel<<<blocks(kGradSize, 256), 256, 0, h->stream>>>(f->W0T, f->b0, f->W1, f->b1, f->W1bf, f->grad, (float)lr);
    XQ_LAUNCH_CHECK();
    h->f64_current = false;
    return XQ_OK;
}

}  // extern "C"
