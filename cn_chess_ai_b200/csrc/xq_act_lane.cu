// xq_act_lane.cu -- one ply of epsilon-greedy self-play for every env, ONE THREAD PER BOARD: DQN::selectAction over
// ChessAI::getAllValidActions (src/dqn.cpp:24-56, src/chessai.cpp:347-368), then -- APPLY -- the rest of the loop body of ChessAI::train
// (src/chessai.cpp:96-119): movePiece, evaluateBoard, checkGameOver / getWinner, the transition into the replay ring, the finished-game
// event, reset; and (collector) the carried fixed-point layer-0 sums + h(s') for the next ply's contraction (xq_act_l0.cuh).
//
// Why: act_team_kernel (xq_act_team.cu, 4 threads per board, 2 CTA barriers, phase data through shared memory) executes 283 warp-
// instructions per env for selection + apply and is bound by instruction issue.  With the register-resident generator of the rollout
// (xq_rollout_lane.cuh: all 16 pieces of the mover in one thread, piece type per position a compile-time constant, no divergence on type)
// nothing is replicated and nothing is exchanged: movegen ~900 + the 198-slot walk for the first Q maximum ~1,700 thread-instructions.
// Mapping: CTA = 64 boards = 64 threads; the Q tile [64 envs][90] is transposed into shared memory [to][board] (coalesced reads,
// conflict-free stores; a thread then reads its own column), the move is applied on the packed nibble board kept in shared memory
// (one thread, one board: no replicated apply), the tail runs on both warps, lane = 4 hidden units.
// Boards with a non-standard piece set are flagged and left to the generic act_kernel (xq_selfplay.cu), as in the team kernel.
//
// STATUS: an alternative, not the default (XQ_ACT_LANE=1 selects it; bit-identical results, tests/test_selfplay_gpu.py runs both).
// Measured on a B200 at 65,536 envs (two collector streams): 74.7 us per ply against 60.5 us with act_team_kernel.  ncu
// (profiles/r2_ncu_act_lane_by_line.txt): 368 warp-instructions per env against 431 -- the selection did shrink (movegen ~47, the Q walk ~85
// per env) but a ONE-ply kernel pays per board what the fused rollout amortises over a launch: the record -> piece-slot conversion (~75 per
// env: a data-dependent loop, a warp runs its longest board) and the carried layer-0 tail (~90 per env: 32 lanes per board however the
// selection is mapped); and it is 7,200 SASS instructions of straight-line code that every warp streams through ONCE: the top stall is
// instruction fetch (1.8 stalled warps per issued instruction), 25 % issue-active at 7 warps per SM.  What would make it win: the
// piece-slot form of every env kept in HBM next to the carried sums (no conversion), and the slider walk as a loop.
#include "xq_act_l0.cuh"
#include "xq_common.cuh"
#include "xq_rollout_lane.cuh"

namespace xq {

constexpr int kLB = 64;            // boards per CTA
constexpr int kLQ = kLB + 1;       // row stride of the Q tile
constexpr int kLQRows = 90;        // Q is indexed by action.to < 90 (src/dqn.cpp:47)
constexpr int kLQStride = 96;      // q90 rows in HBM (dqn_q90_device)
constexpr uint32_t kLUpdRestart = 1u << 22, kLUpdValid = 1u << 23, kLUpdFresh = 1u << 24;

struct LaneTransition {            // == xq_transition
    uint32_t s[12], s2[12];
    uint16_t action; uint8_t mover, done;
    int32_t reward;
    uint32_t pad[6];
};
static_assert(sizeof(LaneTransition) == sizeof(xq_transition), "transition layout");

template <bool APPLY>
__global__ void __launch_bounds__(kLB) act_lane_kernel(xq_env_rec* __restrict__ envs, int64_t n, uint64_t env_id0, uint64_t seed,
                                                      const float* __restrict__ q90, uint32_t eps_thr, int train_done,
                                                      uint16_t* __restrict__ actions_out, LaneTransition* __restrict__ ring, int64_t ring_cap,
                                                      int64_t ring_pos, xq_env_stats* __restrict__ stats, xq_game_event* __restrict__ events,
                                                      unsigned long long* __restrict__ event_count, int64_t event_cap, uint32_t event_ply,
                                                      uint8_t* __restrict__ nonstd, const ActCarry carry, uint32_t event_env0) {
    __shared__ uint8_t s_slot[32 * kLB];          // [slot][board]
    __shared__ uint32_t s_words[12 * kLB];        // [word][board]: the packed nibble board, the move is applied here
    __shared__ float s_q[kLQRows * kLQ];          // [to][board]
    __shared__ uint32_t s_upd[kLB];               // what the ply did to the board, for the carried sums: from | to << 7 | code << 14 | cap << 18 | flags
    __shared__ uint32_t s_magic[XQ_MAX_ACTIONS + 1];
    __shared__ uint32_t s_geo[kGeoWords];
    __shared__ uint32_t s_view[kViewWords * kLB];
    const int tid = threadIdx.x, lane = tid & 31;
    const int64_t env0 = (int64_t)blockIdx.x * kLB, env = env0 + tid;
    for (int d = tid + 1; d <= XQ_MAX_ACTIONS; d += kLB) s_magic[d] = team_mod_magic((uint32_t)d);
    for (int i = tid; i < kGeoWords; i += kLB) s_geo[i] = geo_word(i);
    __syncthreads();                                               // the geometry table is read by the generator below
    const bool carrying = APPLY && carry.Z != nullptr;             // kernel-uniform
    if (carrying && env < n) {                                     // this env's sum (512 B) towards L2 now: the tail reads it after the selection
#pragma unroll
        for (int i = 0; i < 4; ++i) asm volatile("prefetch.global.L2 [%0];" ::"l"(carry.Z + env * 128 + i * 32));
    }
    // ---- Q tile: q90[env0 .. env0 + 64)[0..89] -> s_q[to][board] ----
    {   // 24 independent 16-byte loads per thread, all in flight before the first store (a row of 96 floats is 24 aligned float4)
        float4 v[kLQStride / 4];
#pragma unroll
        for (int j = 0; j < kLQStride / 4; ++j) {
            const int e = 4 * (tid + kLB * j), b = e / kLQStride;
            v[j] = env0 + b < n ? __ldg(reinterpret_cast<const float4*>(q90 + env0 * kLQStride + e)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < kLQStride / 4; ++j) {
            const int e = 4 * (tid + kLB * j), b = e / kLQStride, to = e - b * kLQStride;
            const float x[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (to + k < kLQRows) s_q[(to + k) * kLQ + b] = x[k];
        }
    }
    // ---- record -> nibble words (shared memory), piece slots, bitboards ----
    bool active = env < n;
    uint4 meta = make_uint4(0u, 0u, 0u, 0u);
    Bits90 red{0, 0, 0}, black{0, 0, 0}, occT{0, 0, 0};
    if (active) {
        const uint4* rec = reinterpret_cast<const uint4*>(envs + env);
        uint32_t w[12];
#pragma unroll
        for (int i = 0; i < 3; ++i) { const uint4 v = rec[i]; w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w; }
        meta = rec[3];
#pragma unroll
        for (int i = 0; i < 12; ++i) s_words[i * kLB + tid] = w[i];
        for (int i = 0; i < 32; ++i) s_slot[i * kLB + tid] = kDeadSq;
        active = team_unpack_record(w, red, black, occT, [&](int s, int q) { s_slot[s * kLB + tid] = (uint8_t)q; });
        if (nonstd) nonstd[env] = active ? 0 : 1;
    }
    int mat_red = 1480, mat_black = 1480, gen_red = 4, gen_black = 85;
    int move_count = 0, player = RED;
    uint32_t ctr = 0;
    bool fresh = false;
    uint32_t own_sq[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) own_sq[w] = lane_word<2>(w);
    Bits90 own = team_open_red(), opp = team_open_black();
    Bits90 oT = team_open_occT();
    if (active) {
        ctr = meta.w;
        gen_red = s_slot[8 * kLB + tid]; gen_black = s_slot[24 * kLB + tid];
        // a finished board is never stepped (chessai.cpp:90,96): with APPLY it restarts from the opening (ChessBoard::reset), counter kept
        fresh = APPLY && ((meta.x & 0xFFFFu) >= XQ_MAX_MOVES || gen_red == kDeadSq || gen_black == kDeadSq);
        if (!fresh) {
            uint32_t wr[4] = {0, 0, 0, 0}, wb[4] = {0, 0, 0, 0};
            mat_red = mat_black = 0;
#pragma unroll
            for (int pos = 0; pos < 16; ++pos) {
                const int s = lane_pos_slot(pos);
                const int qr = s_slot[s * kLB + tid], qb = s_slot[(16 + s) * kLB + tid];
                wr[pos >> 2] |= (uint32_t)qr << (8 * (pos & 3));
                wb[pos >> 2] |= (uint32_t)qb << (8 * (pos & 3));
                const int sc = piece_score(slot_type(s));            // material per side (ChessAI::evaluateBoard :313-341)
                mat_red += qr != kDeadSq ? sc : 0; mat_black += qb != kDeadSq ? sc : 0;
            }
            move_count = (int)(meta.x & 0xFFFFu); player = (int)((meta.x >> 16) & 0xFFu);
            const bool redp = player == RED;
#pragma unroll
            for (int w = 0; w < 4; ++w) own_sq[w] = redp ? wr[w] : wb[w];
            own = redp ? red : black; opp = redp ? black : red; oT = occT;
        } else {
            gen_red = 4; gen_black = 85;
        }
    }
    // lanes without a board (tail of the last CTA, non-standard piece sets) act on the opening position and are never stored
    uint32_t sdesc[4], cw[4], dw[4], tot = 0;
    view_init(s_view + tid, kLB);
    view_store(s_view + tid, kLB, own, opp, oT);
    lane_movegen(own_sq, MemView{s_view + tid, kLB}, player, s_geo, sdesc, cw, dw);
#pragma unroll
    for (int w = 0; w < 4; ++w) tot = dp4a_u(cw[w], 0x01010101u, tot);
    __syncthreads();                                               // the Q tile and the modulo table
    uint32_t mv = 0;
    if (tot > 0) {
        const uint64_t x = rng(seed, env_id0 + (uint64_t)env, ctr);
        const uint32_t coin31 = (uint32_t)(x & 0x7FFFFFFFu), idx31 = (uint32_t)(x >> 33);
        if (coin31 < eps_thr) mv = lane_select_kth(own_sq, player, sdesc, cw, dw, team_mod(idx31, tot, s_magic[tot]), tot);      // rand()/RAND_MAX < epsilon (src/dqn.cpp:30-34)
        else mv = lane_select_greedy(own_sq, player, sdesc, cw, dw, [&](int to) { return s_q[to * kLQ + tid]; });                // first maximum of Q[action.to] (:39-52)
    }
    unsigned a_steps = 0, a_games = 0, a_red = 0, a_black = 0, a_capg = 0, a_caps = 0, a_legal = 0;
    int a_reward = 0;
    uint32_t upd = 0;                                              // bit 23: the env has a carried sum to update
    if (active) {
        const int from = (int)(mv & 0xFFu), to = (int)((mv >> 8) & 0xFFu);
        const int a = tot ? (int)XQ_ACTION(from, to) : (int)XQ_ACTION_NONE;
        if (actions_out) actions_out[env] = (uint16_t)a;
        if (APPLY) {
            if (fresh) {      // ChessBoard::reset before the move: opening board, moveCount 0, Red to move, scores 0; flags and ctr stay
#pragma unroll
                for (int i = 0; i < 12; ++i) s_words[i * kLB + tid] = kOpening[i];
            }
            const uint32_t m0 = fresh ? (meta.x & 0xFF000000u) : meta.x;
            int red_score = fresh ? 0 : (int)meta.y, black_score = fresh ? 0 : (int)meta.z;
            const int mover = player;
            LaneTransition t;
#pragma unroll
            for (int i = 0; i < 12; ++i) t.s[i] = s_words[i * kLB + tid];
            t.mover = (uint8_t)mover;
            bool restart;
            if (tot == 0) {   // no action: the episode loop ends (chessai.cpp:100-103); recorded as a terminal null transition
                t.action = 0; t.done = 1; t.reward = 0;
#pragma unroll
                for (int i = 0; i < 12; ++i) t.s2[i] = t.s[i];
                if (events) {   // gameCompleted still fires (:161); winner = first General in square order
                    const int win = (gen_red == kDeadSq && gen_black == kDeadSq) ? NOCOLOR : (gen_red < gen_black ? RED : BLACK);
                    const unsigned long long slot = atomicAdd(event_count, 1ull);
                    if ((int64_t)slot < event_cap)
                        events[slot] = xq_game_event{event_ply, event_env0 + (uint32_t)env, red_score, black_score, (uint16_t)move_count, (uint8_t)win, 2, 0u};
                }
                ctr++; a_games++;
                restart = true;
                upd = kLUpdValid | kLUpdRestart;
            } else {
                // ChessBoard::movePiece on the nibble board (src/chessboard.cpp:43-63), through shared memory: run-time word index
                uint32_t* wf = &s_words[(from >> 3) * kLB + tid];
                const int code = (int)((*wf >> (4 * (from & 7))) & 15u);
                *wf &= ~(15u << (4 * (from & 7)));
                uint32_t* wt = &s_words[(to >> 3) * kLB + tid];
                const int cap = (int)((*wt >> (4 * (to & 7))) & 15u);
                *wt = (*wt & ~(15u << (4 * (to & 7)))) | ((uint32_t)code << (4 * (to & 7)));
                if (cap != 0) {
                    const int sc = piece_score(type_of(cap));
                    if (cap >= 8) { red_score += sc; mat_black -= sc; } else { black_score += sc; mat_red -= sc; }   // captured Black => Red scores (:53-57)
                    a_caps++;
                }
                move_count++; player ^= 1; ctr++;
                const int reward = reward_from_material(mover == RED ? mat_red - mat_black : mat_black - mat_red, move_count);
                const bool took_general = type_of(cap) == GENERAL;
                const bool over = took_general || move_count >= XQ_MAX_MOVES;
                // getWinner: colour of the first General in square order
                int gr = gen_red, gb = gen_black;
                if (from == gr) gr = to; else if (from == gb) gb = to;
                const int win = took_general ? mover : (gr < gb ? RED : BLACK);
#pragma unroll
                for (int i = 0; i < 12; ++i) t.s2[i] = s_words[i * kLB + tid];
                t.action = (uint16_t)a; t.reward = reward;
                t.done = (uint8_t)((over || (train_done && move_count + 1 >= XQ_MAX_MOVES)) ? 1 : 0);     // chessai.cpp:119 / :227
                a_steps++; a_legal += tot; a_reward += reward;
                if (over) {
                    a_games++;
                    if (win == RED) a_red++; else a_black++;
                    if (move_count < XQ_MAX_MOVES) a_capg++;
                    if (events) {   // gameCompleted(game, board->getRedScore(), board->getBlackScore()), chessai.cpp:161
                        const unsigned long long slot = atomicAdd(event_count, 1ull);
                        if ((int64_t)slot < event_cap)
                            events[slot] = xq_game_event{event_ply, event_env0 + (uint32_t)env, red_score, black_score, (uint16_t)move_count, (uint8_t)win,
                                                         (uint8_t)(took_general ? 0 : 1), 0u};
                    }
                }
                restart = over;
                upd = kLUpdValid | (uint32_t)from | ((uint32_t)to << 7) | ((uint32_t)code << 14) | ((uint32_t)cap << 18) | (over ? kLUpdRestart : 0u) | (fresh ? kLUpdFresh : 0u);
            }
            if (ring) {
#pragma unroll
                for (int i = 0; i < 6; ++i) t.pad[i] = 0;
                uint4* dst = reinterpret_cast<uint4*>(ring + (ring_pos + env) % ring_cap);
                const uint4* src = reinterpret_cast<const uint4*>(&t);
#pragma unroll
                for (int i = 0; i < 8; ++i) dst[i] = src[i];
            }
            uint4* rec = reinterpret_cast<uint4*>(envs + env);
            if (restart) {      // ChessBoard::reset (src/chessboard.cpp:95-102)
#pragma unroll
                for (int i = 0; i < 12; ++i) t.s2[i] = kOpening[i];
                rec[3] = make_uint4(m0 & 0xFF000000u, 0u, 0u, ctr);
            } else {
                rec[3] = make_uint4((uint32_t)(move_count & 0xFFFF) | ((uint32_t)player << 16) | (m0 & 0xFF000000u), (uint32_t)red_score, (uint32_t)black_score, ctr);
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) rec[i] = make_uint4(t.s2[4 * i], t.s2[4 * i + 1], t.s2[4 * i + 2], t.s2[4 * i + 3]);
            if (carrying) {     // the board the carried sum will belong to (sanitised the way l0_act_kernel remembers boards)
                uint4* pv = reinterpret_cast<uint4*>(carry.Prev + env * 12);
#pragma unroll
                for (int i = 0; i < 3; ++i)
                    pv[i] = make_uint4(act_sanitize(t.s2[4 * i]), act_sanitize(t.s2[4 * i + 1]), act_sanitize(t.s2[4 * i + 2]),
                                       act_sanitize(t.s2[4 * i + 3]) & (i == 2 ? 0xFFu : 0xFFFFFFFFu));
            }
        }
    }
    if (carrying) s_upd[tid] = upd;
    if (APPLY && stats) {      // every counter of a warp fits 32 bits: one REDUX each
        const unsigned v[8] = {a_steps, a_games, a_red, a_black, a_capg, a_caps, 0u, a_legal};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i == 6) continue;
            const unsigned r = __reduce_add_sync(0xFFFFFFFFu, v[i]);
            if (lane == 0 && r != 0) atomicAdd(reinterpret_cast<unsigned long long*>(stats) + i, (unsigned long long)r);
        }
        const int rs = __reduce_add_sync(0xFFFFFFFFu, a_reward);
        if (lane == 0 && rs != 0) atomicAdd(reinterpret_cast<unsigned long long*>(stats) + 6, (unsigned long long)(long long)rs);
    }
    if (!carrying) return;
    __syncthreads();
    // ---- both warps: carry the fixed-point layer-0 sums of the CTA's 64 envs across the ply (xq_act_l0.cuh) -- the moved piece's row leaves
    // at `from` and enters at `to`, a captured piece's row leaves; a restarted env takes the sum of the opening position -- and emit h(s) of the
    // new position for the next ply's contraction.  Warp w handles envs 32 w .. 32 w + 31, eight at a time; lane = 4 hidden units.
    {
        const int w = tid >> 5;
        const float is = *carry.inv_scale;
        const uint4* W = reinterpret_cast<const uint4*>(carry.W0Q) + lane;
        const uint4 zo = reinterpret_cast<const uint4*>(carry.zOpen)[lane];
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
            uint32_t u[8];
            uint4 z[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {                          // all 8 sums first (one round trip instead of eight), then env by env
                const int b = 32 * w + 8 * c + i;
                u[i] = s_upd[b];
                const bool carried_sum = (u[i] & (kLUpdValid | kLUpdRestart | kLUpdFresh)) == kLUpdValid;
                z[i] = carried_sum ? reinterpret_cast<const uint4*>(carry.Z + (env0 + b) * 128)[lane] : zo;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (!(u[i] & kLUpdValid)) continue;                 // warp-uniform
                const int64_t e = env0 + 32 * w + 8 * c + i;
                if (!(u[i] & kLUpdRestart)) {
                    const int from = (int)(u[i] & 127u), to = (int)((u[i] >> 7) & 127u), code = (int)((u[i] >> 14) & 15u), cap = (int)((u[i] >> 18) & 15u);
                    const uint4 r0 = W[(size_t)(from * 14 + code - 1) * 32], r1 = W[(size_t)(to * 14 + code - 1) * 32];
                    const uint4 r2 = W[(size_t)(cap ? to * 14 + cap - 1 : XQ_STATE_SIZE) * 32];      // row 1260 = zeros
                    z[i] = make_uint4(z[i].x - r0.x + r1.x - r2.x, z[i].y - r0.y + r1.y - r2.y, z[i].z - r0.z + r1.z - r2.z, z[i].w - r0.w + r1.w - r2.w);
                }
                reinterpret_cast<uint4*>(carry.Z + e * 128)[lane] = z[i];
                act_emit_h(z[i], is, carry.Hhi, carry.Hlo, e, lane);
            }
        }
    }
}

cudaError_t launch_act_lane(bool apply, xq_env_rec* envs, int64_t n, uint64_t env_id0, uint64_t seed, const float* q90, uint32_t eps_thr, int train_done,
                            uint16_t* actions_out, void* ring, int64_t ring_cap, int64_t ring_pos, xq_env_stats* stats, xq_game_event* events,
                            unsigned long long* event_count, int64_t event_cap, uint32_t event_ply, uint8_t* nonstd, const ActCarry* carry, uint32_t event_env0, cudaStream_t stream) {
    const ActCarry cy = carry ? *carry : ActCarry{};
    const unsigned grid = (unsigned)((n + kLB - 1) / kLB);
    if (apply)
        act_lane_kernel<true><<<grid, kLB, 0, stream>>>(envs, n, env_id0, seed, q90, eps_thr, train_done, actions_out, (LaneTransition*)ring, ring_cap,
                                                       ring_pos, stats, events, event_count, event_cap, event_ply, nonstd, cy, event_env0);
    else
        act_lane_kernel<false><<<grid, kLB, 0, stream>>>(envs, n, env_id0, seed, q90, eps_thr, train_done, actions_out, (LaneTransition*)ring, ring_cap,
                                                        ring_pos, stats, events, event_count, event_cap, event_ply, nonstd, cy, event_env0);
    ++g_launches;
    return cudaGetLastError();
}

}  // namespace xq
