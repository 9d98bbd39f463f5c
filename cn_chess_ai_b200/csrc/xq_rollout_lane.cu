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
#ifndef XQ_LANE_MINBLOCKS
#define XQ_LANE_MINBLOCKS 1      // A/B builds: resident CTAs per SM the register allocation must allow
#endif

// BS = threads (boards) per CTA, a compile-time constant: the [word][thread] strides of the shared-memory slices become immediates
template <int BS>
__global__ void __launch_bounds__(BS, XQ_LANE_MINBLOCKS) rollout_lane_kernel(xq_env_rec* __restrict__ envs, int64_t n, uint64_t env_id0, uint64_t seed, int n_plies,
                                                                      xq_trace_rec* __restrict__ trace, xq_env_stats* __restrict__ stats,
                                                                      uint8_t* __restrict__ nonstd, const xq_env_rec* __restrict__ src,
                                                                      xq_env_rec* __restrict__ mirror) {
    __shared__ uint8_t s_slot[32 * BS];                   // [slot][thread]
    __shared__ uint32_t s_magic[XQ_MAX_ACTIONS + 1];
    __shared__ uint32_t s_geo[kGeoWords];                 // geometry table of the leapers (xq_bitboard.cuh)
    __shared__ uint32_t s_view[kViewWords * BS];          // [word][thread]: the bitboards for run-time word indices (xq_bitboard.cuh: MemView)
    constexpr int bs = BS;
    const int tid = threadIdx.x;
    const int64_t env = (int64_t)blockIdx.x * bs + tid;
    for (int d = tid + 1; d <= XQ_MAX_ACTIONS; d += bs) s_magic[d] = team_mod_magic((uint32_t)d);
    for (int i = tid; i < kGeoWords; i += bs) s_geo[i] = geo_word(i);
    LaneState st;
    LaneStats a{0, 0, 0, 0, 0, 0, 0, 0};
    bool active = env < n;
    uint32_t flags = 0;
    if (active) {
        const uint4* rec = reinterpret_cast<const uint4*>((src ? src : envs) + env);      // src: mapped host boards (xq_env_rollout_random_io)
        uint32_t w[12];
#pragma unroll
        for (int i = 0; i < 3; ++i) { const uint4 v = rec[i]; w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w; }
        const uint4 m = rec[3];
        if (src) {      // the device array follows the host boards (a board left to the generic kernel is read from there)
            uint4* d = reinterpret_cast<uint4*>(envs + env);
#pragma unroll
            for (int i = 0; i < 3; ++i) d[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
            d[3] = m;
        }
        for (int i = 0; i < 32; ++i) s_slot[i * bs + tid] = kDeadSq;
        Bits90 red, black, occT;
        active = team_unpack_record_m(w, s_view + tid, bs, red, black, occT, [&](int s, int q) { s_slot[s * bs + tid] = (uint8_t)q; });      // (the view memory is initialised below)
        if (nonstd) nonstd[env] = active ? 0 : 1;      // a non-standard piece set is left to the generic kernel
        flags = m.x & 0xFF000000u;
        lane_load(st, [&](int s) { return (int)s_slot[s * bs + tid]; }, red, black, occT, (int)(m.x & 0xFFFFu), (int)((m.x >> 16) & 0xFFu), (int)m.y, (int)m.z, m.w);
        view_init(s_view + tid, bs);
        view_store(s_view + tid, bs, st.own, st.opp, st.occT);
    }
    __syncthreads();                                      // the modulo and geometry tables
    if (active) {
        const uint64_t rng_base = seed + (env_id0 + (uint64_t)env) * 0x9E3779B97F4A7C15ull;
        xq_trace_rec* t = trace ? trace + env : nullptr;
#pragma unroll 1
        for (int p = 0; p < n_plies; ++p) {
            lane_ply(st, a, rng_base, s_magic, s_geo, s_view + tid, bs, t);
            if (t) t += n;
        }
        uint32_t words[12];
        lane_store_words_mem(st, s_view + tid, bs, words);      // kViewWords >= 16 words per thread
        uint4* rec = reinterpret_cast<uint4*>(envs + env);
#pragma unroll
        for (int i = 0; i < 3; ++i) rec[i] = make_uint4(words[4 * i], words[4 * i + 1], words[4 * i + 2], words[4 * i + 3]);
        rec[3] = make_uint4((uint32_t)(st.move_count & 0xFFFF) | ((uint32_t)st.player << 16) | flags, (uint32_t)st.red, (uint32_t)st.black, st.ctr);
        if (mirror) {   // the same record into the caller's mapped host buffer
            uint4* mr = reinterpret_cast<uint4*>(mirror + env);
#pragma unroll
            for (int i = 0; i < 3; ++i) mr[i] = make_uint4(words[4 * i], words[4 * i + 1], words[4 * i + 2], words[4 * i + 3]);
            mr[3] = make_uint4((uint32_t)(st.move_count & 0xFFFF) | ((uint32_t)st.player << 16) | flags, (uint32_t)st.red, (uint32_t)st.black, st.ctr);
        }
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

// ---------------------------------------------------------------------------------------------------------------------------------
// API mode: ChessAI::getAllValidActions(currentPlayer) for every env (src/chessai.cpp:347-368) as ORDERED LISTS in HBM, with the same
// register-resident generator.  One thread per board: counts + descriptors of the 16 pieces (lane_movegen), every piece's place in the
// reference order from the byte-SIMD prefix (four dp4a on the (squares, counts) words), its actions written into a per-thread row of
// shared memory ([u16 pair j][thread], row stride = threads + 1 words: conflict-free for the thread-major writes and for the
// board-major reads), then the CTA streams the rows out as 16-byte stores, 256 contiguous bytes per board.  No board-in-shared-memory
// walk, no per-square loop: the loops run over (position, direction) with the same trip structure in every lane.
// Boards with a non-standard piece set are flagged and left to the generic kernel (legal_moves_kernel, xq_env.cu).
constexpr int kLmThreads = 64;      // (33 KB of list rows + 12 KB of view memory per 128 boards would exceed the 48 KB of static shared memory)
constexpr int kLmRow = 65;                                   // words per thread row: 64 pairs of actions + 1 (odd stride: bank = (thread + word) % 32)
__global__ void __launch_bounds__(kLmThreads) legal_moves_lane_kernel(const xq_env_rec* __restrict__ envs, int64_t n, uint8_t* __restrict__ counts,
                                                                     uint4* __restrict__ actions, uint8_t* __restrict__ nonstd) {
    __shared__ uint8_t s_slot[32 * kLmThreads];
    __shared__ uint32_t s_list[kLmThreads * kLmRow];
    __shared__ uint8_t s_cnt[kLmThreads];                    // list size; 0xFF = not produced here (tail of the grid, non-standard piece set)
    __shared__ uint32_t s_geo[kGeoWords];
    __shared__ uint32_t s_view[kViewWords * kLmThreads];
    const int tid = threadIdx.x;
    const int64_t env0 = (int64_t)blockIdx.x * kLmThreads, env = env0 + tid;
    for (int i = tid; i < kGeoWords; i += kLmThreads) s_geo[i] = geo_word(i);
    __syncthreads();
    int cnt = 0xFF;
    if (env < n) {
        const uint4* rec = reinterpret_cast<const uint4*>(envs + env);
        uint32_t w[12];
#pragma unroll
        for (int i = 0; i < 3; ++i) { const uint4 v = rec[i]; w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w; }
        const int player = (int)((rec[3].x >> 16) & 0xFFu);
        for (int i = 0; i < 32; ++i) s_slot[i * kLmThreads + tid] = kDeadSq;
        Bits90 red, black, occT;
        const bool ok = team_unpack_record_m(w, s_view + tid, kLmThreads, red, black, occT, [&](int s, int q) { s_slot[s * kLmThreads + tid] = (uint8_t)q; });
        if (nonstd) nonstd[env] = ok ? 0 : 1;
        if (ok) {
            uint32_t own_sq[4] = {0, 0, 0, 0};
#pragma unroll
            for (int pos = 0; pos < 16; ++pos)
                own_sq[pos >> 2] |= (uint32_t)s_slot[((player ? 16 : 0) + lane_pos_slot(pos)) * kLmThreads + tid] << (8 * (pos & 3));
            uint32_t sdesc[4], cw[4], dw[4];
            view_init(s_view + tid, kLmThreads);
            view_store(s_view + tid, kLmThreads, player ? black : red, player ? red : black, occT);
            lane_movegen(own_sq, MemView{s_view + tid, kLmThreads}, player, s_geo, sdesc, cw, dw);
            uint16_t* row = reinterpret_cast<uint16_t*>(s_list + tid * kLmRow);
            cnt = lane_emit_actions(own_sq, player, sdesc, cw, dw, [&](int idx, int a, bool live) { if (live) row[idx] = (uint16_t)a; });
            counts[env] = (uint8_t)cnt;
        }
    }
    s_cnt[tid] = (uint8_t)cnt;
    __syncthreads();
    const int live = (int)min((int64_t)kLmThreads, n - env0);
    for (int c = tid; c < live * 16; c += kLmThreads) {      // 16-byte chunk c & 15 of board c >> 4: actions 8 (c & 15) .. + 7
        const int b = c >> 4, j = (c & 15) * 4, cb = s_cnt[b];
        if (cb == 0xFF) continue;
        uint32_t v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {                        // entries past the count are XQ_ACTION_NONE (the rows are not pre-filled)
            const int a0 = 2 * (j + u);
            const uint32_t x = s_list[b * kLmRow + j + u];
            v[u] = a0 + 1 < cb ? x : (a0 < cb ? (x | 0xFFFF0000u) : 0xFFFFFFFFu);
        }
        actions[(env0 + b) * 16 + (c & 15)] = make_uint4(v[0], v[1], v[2], v[3]);
    }
}
cudaError_t launch_legal_moves_lane(const xq_env_rec* envs, int64_t n, uint8_t* counts, uint32_t* actions, uint8_t* nonstd, cudaStream_t stream) {
    legal_moves_lane_kernel<<<(unsigned)((n + kLmThreads - 1) / kLmThreads), kLmThreads, 0, stream>>>(envs, n, counts, reinterpret_cast<uint4*>(actions), nonstd);
    ++g_launches;
    return cudaGetLastError();
}

// DQN::selectAction's exploring branch for every env (src/dqn.cpp:30-34) = the random policy of the fused rollout: list[idx31 % count]
// with idx31 from xq_rng(seed, env id, the env's ply counter); XQ_ACTION_NONE for an empty list
__global__ void __launch_bounds__(256) pick_random_kernel(const xq_env_rec* __restrict__ envs, int64_t n, uint64_t env_id0, uint64_t seed,
                                                         const uint8_t* __restrict__ counts, const uint16_t* __restrict__ lists, uint16_t* __restrict__ out) {
    const int64_t env = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= n) return;
    const uint32_t cnt = counts[env];
    const uint32_t idx31 = (uint32_t)(rng(seed, env_id0 + (uint64_t)env, envs[env].ctr) >> 33);
    out[env] = cnt ? lists[env * XQ_MAX_ACTIONS + idx31 % cnt] : (uint16_t)XQ_ACTION_NONE;
}
cudaError_t launch_pick_random(const xq_env_rec* envs, int64_t n, uint64_t env_id0, uint64_t seed, const uint8_t* counts, const uint16_t* lists, uint16_t* out,
                               cudaStream_t stream) {
    pick_random_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(envs, n, env_id0, seed, counts, lists, out);
    ++g_launches;
    return cudaGetLastError();
}

cudaError_t launch_rollout_lane(xq_env_rec* envs, int64_t n, uint64_t env_id0, uint64_t seed, int n_plies, xq_trace_rec* trace, xq_env_stats* stats,
                                uint8_t* nonstd, cudaStream_t stream, const xq_env_rec* src, xq_env_rec* mirror) {
    if (n <= 148 * 4 * 32) rollout_lane_kernel<32><<<(unsigned)((n + 31) / 32), 32, 0, stream>>>(envs, n, env_id0, seed, n_plies, trace, stats, nonstd, src, mirror);
    else rollout_lane_kernel<kLaneMaxThreads><<<(unsigned)((n + kLaneMaxThreads - 1) / kLaneMaxThreads), kLaneMaxThreads, 0, stream>>>(envs, n, env_id0, seed, n_plies, trace, stats, nonstd, src, mirror);
    ++g_launches;
    return cudaGetLastError();
}

}  // namespace xq
