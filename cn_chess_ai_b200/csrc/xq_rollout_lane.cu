// xq_rollout_lane.cu -- fused random-policy rollout, ONE THREAD PER BOARD, the whole board in registers (xq_rollout_lane.cuh).
//
// Replaces the loop body of ChessAI::train without the network (src/chessai.cpp:96-119).  Mapping: lane = board; the record is
// loaded once (64 B per thread, a warp reads 2 KB contiguous), converted to the position-ordered square words + bitboards through a
// small shared-memory staging area, then n_plies plies run with no barrier, no exchange and no shared-memory traffic except the
// one read of the modulo table; traces go out as one coalesced 8-byte store per ply; the record is written back once.
// CTA size: 32 threads while the grid would otherwise not cover the 148 SMs (4096 envs -> 128 one-warp CTAs on 128 SMs, every warp
// alone on its scheduler: that case is bound by the latency of one warp's ply), 128 threads for large env counts.
#include "xq_common.cuh"
#include "xq_rollout_lane.cuh"

namespace xq {

constexpr int kLaneMaxThreads = 128;

__global__ void __launch_bounds__(kLaneMaxThreads) rollout_lane_kernel(xq_env_rec* __restrict__ envs, int64_t n, uint64_t env_id0, uint64_t seed, int n_plies,
                                                                      xq_trace_rec* __restrict__ trace, xq_env_stats* __restrict__ stats,
                                                                      uint8_t* __restrict__ nonstd) {
    __shared__ uint8_t s_slot[32 * kLaneMaxThreads];      // [slot][thread]
    __shared__ uint32_t s_magic[XQ_MAX_ACTIONS + 1];
    const int tid = threadIdx.x, bs = blockDim.x;
    const int64_t env = (int64_t)blockIdx.x * bs + tid;
    for (int d = tid + 1; d <= XQ_MAX_ACTIONS; d += bs) s_magic[d] = team_mod_magic((uint32_t)d);
    LaneState st;
    LaneStats a{0, 0, 0, 0, 0, 0, 0, 0};
    bool active = env < n;
    uint32_t flags = 0;
    if (active) {
        const uint4* rec = reinterpret_cast<const uint4*>(envs + env);
        uint32_t w[12];
#pragma unroll
        for (int i = 0; i < 3; ++i) { const uint4 v = rec[i]; w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w; }
        const uint4 m = rec[3];
        for (int i = 0; i < 32; ++i) s_slot[i * bs + tid] = kDeadSq;
        Bits90 red, black, occT;
        active = team_unpack_record(w, red, black, occT, [&](int s, int q) { s_slot[s * bs + tid] = (uint8_t)q; });
        if (nonstd) nonstd[env] = active ? 0 : 1;      // a non-standard piece set is left to the generic kernel
        flags = m.x & 0xFF000000u;
        lane_load(st, [&](int s) { return (int)s_slot[s * bs + tid]; }, red, black, occT, (int)(m.x & 0xFFFFu), (int)((m.x >> 16) & 0xFFu), (int)m.y, (int)m.z, m.w);
    }
    __syncthreads();                                      // the modulo table
    if (active) {
        const uint64_t rng_base = seed + (env_id0 + (uint64_t)env) * 0x9E3779B97F4A7C15ull;
        xq_trace_rec* t = trace ? trace + env : nullptr;
#pragma unroll 1
        for (int p = 0; p < n_plies; ++p) {
            lane_ply(st, a, rng_base, s_magic, t);
            if (t) t += n;
        }
        uint32_t words[12];
        lane_store_words(st, words);
        uint4* rec = reinterpret_cast<uint4*>(envs + env);
#pragma unroll
        for (int i = 0; i < 3; ++i) rec[i] = make_uint4(words[4 * i], words[4 * i + 1], words[4 * i + 2], words[4 * i + 3]);
        rec[3] = make_uint4((uint32_t)(st.move_count & 0xFFFF) | ((uint32_t)st.player << 16) | flags, (uint32_t)st.red, (uint32_t)st.black, st.ctr);
    }
    if (stats) {
        unsigned long long v[8] = {a.steps, a.games, a.red, a.black, a.capg, a.caps, (unsigned long long)a.reward, a.legal};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            unsigned long long r = v[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xFFFFFFFFu, r, o);
            if ((tid & 31) == 0 && r != 0) atomicAdd(reinterpret_cast<unsigned long long*>(stats) + i, r);
        }
    }
}

cudaError_t launch_rollout_lane(xq_env_rec* envs, int64_t n, uint64_t env_id0, uint64_t seed, int n_plies, xq_trace_rec* trace, xq_env_stats* stats,
                                uint8_t* nonstd, cudaStream_t stream) {
    const int bs = n <= 148 * 4 * 32 ? 32 : kLaneMaxThreads;
    rollout_lane_kernel<<<(unsigned)((n + bs - 1) / bs), bs, 0, stream>>>(envs, n, env_id0, seed, n_plies, trace, stats, nonstd);
    ++g_launches;
    return cudaGetLastError();
}

}  // namespace xq
