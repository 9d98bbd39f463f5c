// xq_act_team.cu -- one ply of epsilon-greedy self-play for every env with a team of 4 threads per board: DQN::selectAction over
// ChessAI::getAllValidActions (xq_act_team.cuh), then -- APPLY -- the rest of the loop body of ChessAI::train
// (src/chessai.cpp:96-119): movePiece, evaluateBoard, checkGameOver / getWinner, the transition into the replay ring, the
// finished-game event, reset.  Replaces the thread-per-board act_kernel (xq_selfplay.cu: one thread walked all 16 pieces of a board
// through type-divergent generic code and staged the whole action list in shared memory), which stays as the fallback for boards
// with non-standard piece sets (only reachable through xq_env_set_boards).
//
// Mapping: CTA = 32 boards x 4 warps, warp = role, lane = board (as rollout_team_kernel).  Warp 0 converts the 32 records to piece
// slots + bitboards; all 128 threads transpose the Q tile [32 envs][96] into shared memory [to][board] (coalesced reads,
// conflict-free stores and loads); phase A counts moves and finds each piece's first Q maximum; phase B selects; warp 0 then
// applies the move on the packed nibble board it kept in shared memory -- a one-ply kernel needs no replicated apply.
#include "xq_act_l0.cuh"
#include "xq_act_team.cuh"
#include "xq_common.cuh"

namespace xq {

constexpr int kAB = 32;   // boards per CTA
constexpr uint32_t kUpdRestart = 1u << 22, kUpdValid = 1u << 23, kUpdFresh = 1u << 24;

struct ActTransition {    // == xq_transition
    uint32_t s[12], s2[12];
    uint16_t action; uint8_t mover, done;
    int32_t reward;
    uint32_t pad[6];
};
static_assert(sizeof(ActTransition) == sizeof(xq_transition), "transition layout");

struct ActIo {            // [item][board]
    uint8_t slot[32 * kAB];
    uint32_t bb[6 * kAB];         // red, black (row-major)
    uint32_t occT[6 * kAB];       // column-major occupancy: the Red part, the Black part
    uint32_t meta[4 * kAB];
    uint32_t words[12 * kAB];     // the packed nibble board: warp 0 applies the move here
    uint32_t mat[2 * kAB];        // material per side (ChessAI::evaluateBoard :313-341)
    uint8_t ok[2 * kAB];          // per colour: the piece set fits the 16 slots
    uint32_t upd[kAB];            // what the ply did to the board, for the carried layer-0 sums: from | to << 7 | code << 14 | cap << 18 | flags
};

template <bool APPLY>
__global__ void __launch_bounds__(kAB * 4) act_team_kernel(xq_env_rec* __restrict__ envs, int64_t n, uint64_t env_id0, uint64_t seed,
                                                          const float* __restrict__ q90, uint32_t eps_thr, int train_done,
                                                          uint16_t* __restrict__ actions_out, ActTransition* __restrict__ ring, int64_t ring_cap,
                                                          int64_t ring_pos, xq_env_stats* __restrict__ stats, xq_game_event* __restrict__ events,
                                                          unsigned long long* __restrict__ event_count, int64_t event_cap, uint32_t event_ply,
                                                          uint8_t* __restrict__ nonstd, const ActCarry carry, uint32_t event_env0) {
    __shared__ TeamShared<kAB> sh;
    __shared__ ActShared<kAB> as;
    __shared__ ActIo io;
    const int tid = threadIdx.x, lane = tid & 31;
    const TeamRole R = team_role<4>(tid >> 5);
    const int64_t env0 = (int64_t)blockIdx.x * kAB, env = env0 + lane;
    team_tables_init(sh, tid, kAB * 4);
    if (tid < kAB) sh.move[tid] = 0;
    if (APPLY && carry.Z != nullptr && env0 + (tid >> 2) < n)      // the CTA's 32 sums (16 KB) towards L2 now: the tail reads them ~20 us later
        asm volatile("prefetch.global.L2 [%0];" ::"l"(carry.Z + env0 * 128 + tid * 32));

    // ---- Q tile: q90[env0 .. env0+32)[96] -> qt[to][board], by warps 2 and 3 while warps 0 and 1 unpack the boards ------------
    if (R.role >= 2) {
        for (int b = (R.role - 2) * 16; b < (R.role - 1) * 16; ++b) {
            const bool in = env0 + b < n;
            const float* src = q90 + (env0 + b) * kQStride;
#pragma unroll
            for (int j = 0; j < kQStride / 32; ++j) as.qt[(lane + 32 * j) * (kAB + 1) + b] = in ? src[lane + 32 * j] : 0.f;
        }
    } else {
        // ---- load: warp 0 converts the Red half of 32 records to slots + bitboards, warp 1 the Black half ----------------------
        const int side = R.role;
        bool ok = env < n;
        if (ok) {
            const uint4* rec = reinterpret_cast<const uint4*>(envs + env);
            uint32_t w[12];
#pragma unroll
            for (int i = 0; i < 3; ++i) { const uint4 v = rec[i]; w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w; }
            for (int i = 0; i < 16; ++i) io.slot[(side * 16 + i) * kAB + lane] = kDeadSq;
            Bits90 bb, occT;
            ok = team_unpack_side(w, side, bb, occT, [&](int s, int q) { io.slot[(side * 16 + s) * kAB + lane] = (uint8_t)q; });
            int mat = 0;      // material of this side (ChessAI::evaluateBoard :313-341)
            for (int i = 0; i < 16; ++i) if (io.slot[(side * 16 + i) * kAB + lane] != kDeadSq) mat += piece_score(slot_type(i));
            io.mat[side * kAB + lane] = (uint32_t)mat;
            io.bb[(3 * side + 0) * kAB + lane] = bb.w0; io.bb[(3 * side + 1) * kAB + lane] = bb.w1; io.bb[(3 * side + 2) * kAB + lane] = bb.w2;
            io.occT[(3 * side + 0) * kAB + lane] = occT.w0; io.occT[(3 * side + 1) * kAB + lane] = occT.w1; io.occT[(3 * side + 2) * kAB + lane] = occT.w2;
            if (side == 0) {
                const uint4 m = rec[3];
#pragma unroll
                for (int i = 0; i < 12; ++i) io.words[i * kAB + lane] = w[i];
                io.meta[0 * kAB + lane] = m.x; io.meta[1 * kAB + lane] = m.y; io.meta[2 * kAB + lane] = m.z; io.meta[3 * kAB + lane] = m.w;
            }
        }
        io.ok[side * kAB + lane] = ok ? 1 : 0;
    }
    __syncthreads();

    // lanes without a board (tail of the last CTA, non-standard piece sets) act on the opening position and are never stored
    const bool active = (io.ok[lane] & io.ok[kAB + lane]) != 0;
    if (R.role == 0 && env < n && nonstd) nonstd[env] = active ? 0 : 1;
    TeamState st;
    team_reset(R, st);
    st.ctr = 0;
    // a finished board is never stepped (chessai.cpp:90,96): with APPLY it restarts from the opening (ChessBoard::reset), counter kept
    bool fresh = false;
    if (active) {
        const uint32_t m0 = io.meta[0 * kAB + lane];
        st.ctr = io.meta[3 * kAB + lane];
        fresh = APPLY && ((m0 & 0xFFFFu) >= XQ_MAX_MOVES || io.slot[8 * kAB + lane] == kDeadSq || io.slot[24 * kAB + lane] == kDeadSq);
        if (!fresh) {
            uint32_t wr = 0, wb = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int s = (int)((R.slots >> (8 * i)) & 0xFFu);
                wr |= (uint32_t)io.slot[s * kAB + lane] << (8 * i);
                wb |= (uint32_t)io.slot[(16 + s) * kAB + lane] << (8 * i);
            }
            const Bits90 red{io.bb[0 * kAB + lane], io.bb[1 * kAB + lane], io.bb[2 * kAB + lane]};
            const Bits90 black{io.bb[3 * kAB + lane], io.bb[4 * kAB + lane], io.bb[5 * kAB + lane]};
            st.occT = Bits90{io.occT[0 * kAB + lane] | io.occT[3 * kAB + lane], io.occT[1 * kAB + lane] | io.occT[4 * kAB + lane],
                             io.occT[2 * kAB + lane] | io.occT[5 * kAB + lane]};
            st.move_count = (int)(m0 & 0xFFFFu); st.player = (int)((m0 >> 16) & 0xFFu);
            const bool redp = st.player == RED;
            st.sq_own = redp ? wr : wb; st.sq_opp = redp ? wb : wr;
            st.own = redp ? red : black; st.opp = redp ? black : red;
        }
    }
    TeamPly pl;
    uint32_t kbest[4];
    team_phase_a<4, kAB>(R, st, pl, sh, lane, 0);
    act_best<kAB>(R, st, pl, as, lane, kbest);
    __syncthreads();
    const uint64_t x = rng(seed, env_id0 + (uint64_t)env, st.ctr);
    const uint32_t tot = act_select<kAB>(R, st, pl, sh, as, lane, kbest, x, eps_thr);
    __syncthreads();

    // ---- warp 0: the chosen action, then the rest of the ply on the nibble board ---------------------------------------------
    const bool carrying = APPLY && carry.Z != nullptr;             // kernel-uniform
    if (R.role != 0 && !carrying) return;
    if (R.role == 0) {
    unsigned long long a_steps = 0, a_games = 0, a_red = 0, a_black = 0, a_capg = 0, a_caps = 0, a_legal = 0;
    long long a_reward = 0;
    uint32_t upd = 0;                                              // bit 23: the env has a carried sum to update
    if (active) {
        const uint32_t mv = sh.move[lane];
        const int from = (int)(mv & 0xFFu), to = (int)((mv >> 8) & 0xFFu);
        const int a = tot ? (int)XQ_ACTION(from, to) : (int)XQ_ACTION_NONE;
        if (actions_out) actions_out[env] = (uint16_t)a;
        if (APPLY) {
            if (fresh) {      // ChessBoard::reset before the move: opening board, moveCount 0, Red to move, scores 0; flags and ctr stay
#pragma unroll
                for (int i = 0; i < 12; ++i) io.words[i * kAB + lane] = kOpening[i];
                io.slot[8 * kAB + lane] = 4; io.slot[24 * kAB + lane] = 85;
                io.mat[lane] = 1480; io.mat[kAB + lane] = 1480;
            }
            uint32_t w[12];
#pragma unroll
            for (int i = 0; i < 12; ++i) w[i] = io.words[i * kAB + lane];
            const uint32_t m0 = fresh ? (io.meta[0 * kAB + lane] & 0xFF000000u) : io.meta[0 * kAB + lane];
            int move_count = (int)(m0 & 0xFFFFu), player = (int)((m0 >> 16) & 0xFFu);
            int red_score = fresh ? 0 : (int)io.meta[1 * kAB + lane], black_score = fresh ? 0 : (int)io.meta[2 * kAB + lane];
            uint32_t ctr = io.meta[3 * kAB + lane];
            const int mover = player;
            ActTransition t;
#pragma unroll
            for (int i = 0; i < 12; ++i) t.s[i] = w[i];
            t.mover = (uint8_t)mover;
            bool restart;
            if (tot == 0) {   // no action: the episode loop ends (chessai.cpp:100-103); recorded as a terminal null transition
                t.action = 0; t.done = 1; t.reward = 0;
#pragma unroll
                for (int i = 0; i < 12; ++i) t.s2[i] = w[i];
                if (events) {   // gameCompleted still fires (:161); winner = first General in square order
                    const int gr = io.slot[8 * kAB + lane], gb = io.slot[24 * kAB + lane];
                    const int win = (gr == kDeadSq && gb == kDeadSq) ? NOCOLOR : (gr < gb ? RED : BLACK);
                    const unsigned long long slot = atomicAdd(event_count, 1ull);
                    if ((int64_t)slot < event_cap)
                        events[slot] = xq_game_event{event_ply, event_env0 + (uint32_t)env, red_score, black_score, (uint16_t)move_count, (uint8_t)win, 2, 0u};
                }
                ctr++; a_games++;
                restart = true;
                upd = kUpdValid | kUpdRestart;
            } else {
                // ChessBoard::movePiece on the nibble board (src/chessboard.cpp:43-63), through shared memory: run-time word index
                uint32_t* wf = &io.words[(from >> 3) * kAB + lane];
                const int code = (int)((*wf >> (4 * (from & 7))) & 15u);
                *wf &= ~(15u << (4 * (from & 7)));
                uint32_t* wt = &io.words[(to >> 3) * kAB + lane];
                const int cap = (int)((*wt >> (4 * (to & 7))) & 15u);
                *wt = (*wt & ~(15u << (4 * (to & 7)))) | ((uint32_t)code << (4 * (to & 7)));
                int mat_red = (int)io.mat[lane], mat_black = (int)io.mat[kAB + lane];
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
                int gr = io.slot[8 * kAB + lane], gb = io.slot[24 * kAB + lane];
                if (from == gr) gr = to; else if (from == gb) gb = to;
                const int win = took_general ? mover : (gr < gb ? RED : BLACK);
#pragma unroll
                for (int i = 0; i < 12; ++i) t.s2[i] = io.words[i * kAB + lane];
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
                upd = kUpdValid | (uint32_t)from | ((uint32_t)to << 7) | ((uint32_t)code << 14) | ((uint32_t)cap << 18) | (over ? kUpdRestart : 0u) | (fresh ? kUpdFresh : 0u);
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
    if (carrying) io.upd[lane] = upd;
    if (APPLY && stats) {      // every counter of a warp fits 32 bits: one REDUX each
        const unsigned v[8] = {(unsigned)a_steps, (unsigned)a_games, (unsigned)a_red, (unsigned)a_black, (unsigned)a_capg, (unsigned)a_caps, 0u, (unsigned)a_legal};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i == 6) continue;
            const unsigned r = __reduce_add_sync(0xFFFFFFFFu, v[i]);
            if (lane == 0 && r != 0) atomicAdd(reinterpret_cast<unsigned long long*>(stats) + i, (unsigned long long)r);
        }
        const int rs = __reduce_add_sync(0xFFFFFFFFu, (int)a_reward);
        if (lane == 0 && rs != 0) atomicAdd(reinterpret_cast<unsigned long long*>(stats) + 6, (unsigned long long)(long long)rs);
    }
    }   // warp 0
    if (!carrying) return;
    __syncthreads();
    // ---- all 4 warps: carry the fixed-point layer-0 sums of the CTA's 32 envs across the ply (xq_act_l0.cuh) -- the moved piece's row
    // leaves at `from` and enters at `to`, a captured piece's row leaves; a restarted env takes the sum of the opening position --
    // and emit h(s) of the new position for the next ply's contraction.  Warp w handles envs w, w + 4, ...; lane = 4 hidden units.
    {
        const int w = tid >> 5;
        const float is = *carry.inv_scale;
        const uint4* W = reinterpret_cast<const uint4*>(carry.W0Q) + lane;
        const uint4 zo = reinterpret_cast<const uint4*>(carry.zOpen)[lane];
        // all 8 sums of the warp first (one round trip to L2 / HBM instead of eight), then env by env: rows, tanh, stores
        uint32_t u[kAB / 4];
        uint4 z[kAB / 4];
#pragma unroll
        for (int i = 0; i < kAB / 4; ++i) {
            u[i] = io.upd[w + 4 * i];
            const bool carried_sum = (u[i] & (kUpdValid | kUpdRestart | kUpdFresh)) == kUpdValid;
            z[i] = carried_sum ? reinterpret_cast<const uint4*>(carry.Z + (env0 + w + 4 * i) * 128)[lane] : zo;
        }
#pragma unroll
        for (int i = 0; i < kAB / 4; ++i) {
            if (!(u[i] & kUpdValid)) continue;                      // warp-uniform
            const int64_t e = env0 + w + 4 * i;
            if (!(u[i] & kUpdRestart)) {
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

// ---------------------------------------------------------------------------------------------------------------------------------
// API mode: ChessAI::getAllValidActions(currentPlayer) for every env as ORDERED LISTS in HBM with the team of 4 threads per board
// (src/chessai.cpp:347-368).  The board-per-thread list kernel (legal_moves_lane_kernel) walks ~7,300 instructions per thread: with few
// envs (one warp per scheduler) a launch lasts as long as that one chain.  Here the chain is ~1/4: warps 0 / 1 unpack the Red / Black half
// of 32 records, every thread generates the moves of its 4 pieces (team_phase_a), finds their place in the reference order with the
// byte-SIMD prefix and writes their actions into the board's row of shared memory (team_emit_actions); the CTA then streams the rows
// out as 16-byte stores, 256 contiguous bytes per board, entries past the count filled on the way out.
constexpr int kLtRow = 65;                                   // words per board row: 64 pairs of actions + 1 (odd stride)
__global__ void __launch_bounds__(kAB * 4) legal_moves_team_kernel(const xq_env_rec* __restrict__ envs, int64_t n, uint8_t* __restrict__ counts,
                                                                  uint4* __restrict__ actions, uint8_t* __restrict__ nonstd) {
    __shared__ TeamShared<kAB> sh;
    __shared__ uint8_t s_slot[32 * kAB];
    __shared__ uint32_t s_bb[6 * kAB], s_occT[6 * kAB], s_player[kAB];
    __shared__ uint8_t s_ok[2 * kAB], s_cnt[kAB];
    __shared__ uint32_t s_list[kAB * kLtRow];
    const int tid = threadIdx.x, lane = tid & 31;
    const TeamRole R = team_role<4>(tid >> 5);
    const int64_t env0 = (int64_t)blockIdx.x * kAB, env = env0 + lane;
    team_tables_init(sh, tid, kAB * 4);
    if (R.role < 2) {
        const int side = R.role;
        bool ok = env < n;
        if (ok) {
            const uint4* rec = reinterpret_cast<const uint4*>(envs + env);
            uint32_t w[12];
#pragma unroll
            for (int i = 0; i < 3; ++i) { const uint4 v = rec[i]; w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w; }
            for (int i = 0; i < 16; ++i) s_slot[(side * 16 + i) * kAB + lane] = kDeadSq;
            Bits90 bb, occT;
            ok = team_unpack_side(w, side, bb, occT, [&](int s, int q) { s_slot[(side * 16 + s) * kAB + lane] = (uint8_t)q; });
            s_bb[(3 * side + 0) * kAB + lane] = bb.w0; s_bb[(3 * side + 1) * kAB + lane] = bb.w1; s_bb[(3 * side + 2) * kAB + lane] = bb.w2;
            s_occT[(3 * side + 0) * kAB + lane] = occT.w0; s_occT[(3 * side + 1) * kAB + lane] = occT.w1; s_occT[(3 * side + 2) * kAB + lane] = occT.w2;
            if (side == 0) s_player[lane] = (rec[3].x >> 16) & 0xFFu;
        }
        s_ok[side * kAB + lane] = ok ? 1 : 0;
    }
    __syncthreads();
    // lanes without a board (tail of the last CTA, non-standard piece sets: left to the generic kernel) work on the opening position, unused
    const bool active = (s_ok[lane] & s_ok[kAB + lane]) != 0;
    if (R.role == 0 && env < n && nonstd) nonstd[env] = active ? 0 : 1;
    TeamState st;
    team_reset(R, st);
    st.ctr = 0;
    if (active) {
        uint32_t wr = 0, wb = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int s = (int)((R.slots >> (8 * i)) & 0xFFu);
            wr |= (uint32_t)s_slot[s * kAB + lane] << (8 * i);
            wb |= (uint32_t)s_slot[(16 + s) * kAB + lane] << (8 * i);
        }
        const Bits90 red{s_bb[0 * kAB + lane], s_bb[1 * kAB + lane], s_bb[2 * kAB + lane]};
        const Bits90 black{s_bb[3 * kAB + lane], s_bb[4 * kAB + lane], s_bb[5 * kAB + lane]};
        st.occT = Bits90{s_occT[0 * kAB + lane] | s_occT[3 * kAB + lane], s_occT[1 * kAB + lane] | s_occT[4 * kAB + lane],
                         s_occT[2 * kAB + lane] | s_occT[5 * kAB + lane]};
        st.player = (int)s_player[lane];
        const bool redp = st.player == RED;
        st.sq_own = redp ? wr : wb; st.sq_opp = redp ? wb : wr;
        st.own = redp ? red : black; st.opp = redp ? black : red;
    }
    TeamPly pl;
    team_phase_a<4, kAB>(R, st, pl, sh, lane, 0);
    __syncthreads();
    uint16_t* row = reinterpret_cast<uint16_t*>(s_list + lane * kLtRow);
    const uint32_t tot = team_emit_actions<kAB>(R, st, pl, sh, lane, [&](int idx, int a) { row[idx] = (uint16_t)a; });
    if (R.role == 0) {
        s_cnt[lane] = active ? (uint8_t)tot : (uint8_t)0xFF;
        if (active) counts[env] = (uint8_t)tot;
    }
    __syncthreads();
    const int live = (int)min((int64_t)kAB, n - env0);
    for (int c = tid; c < live * 16; c += kAB * 4) {         // 16-byte chunk c & 15 of board c >> 4: actions 8 (c & 15) .. + 7
        const int b = c >> 4, j = (c & 15) * 4, cb = s_cnt[b];
        if (cb == 0xFF) continue;
        uint32_t v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {                        // entries past the count are XQ_ACTION_NONE (the rows are not pre-filled)
            const int a0 = 2 * (j + u);
            const uint32_t x = s_list[b * kLtRow + j + u];
            v[u] = a0 + 1 < cb ? x : (a0 < cb ? (x | 0xFFFF0000u) : 0xFFFFFFFFu);
        }
        actions[(env0 + b) * 16 + (c & 15)] = make_uint4(v[0], v[1], v[2], v[3]);
    }
}
cudaError_t launch_legal_moves_team(const xq_env_rec* envs, int64_t n, uint8_t* counts, uint32_t* actions, uint8_t* nonstd, cudaStream_t stream) {
    legal_moves_team_kernel<<<(unsigned)((n + kAB - 1) / kAB), kAB * 4, 0, stream>>>(envs, n, counts, reinterpret_cast<uint4*>(actions), nonstd);
    ++g_launches;
    return cudaGetLastError();
}

cudaError_t launch_act_team(bool apply, xq_env_rec* envs, int64_t n, uint64_t env_id0, uint64_t seed, const float* q90, uint32_t eps_thr, int train_done,
                            uint16_t* actions_out, void* ring, int64_t ring_cap, int64_t ring_pos, xq_env_stats* stats, xq_game_event* events,
                            unsigned long long* event_count, int64_t event_cap, uint32_t event_ply, uint8_t* nonstd, const ActCarry* carry, uint32_t event_env0, cudaStream_t stream) {
    const ActCarry cy = carry ? *carry : ActCarry{};
    const unsigned grid = (unsigned)((n + kAB - 1) / kAB);
    if (apply)
        act_team_kernel<true><<<grid, kAB * 4, 0, stream>>>(envs, n, env_id0, seed, q90, eps_thr, train_done, actions_out, (ActTransition*)ring, ring_cap,
                                                           ring_pos, stats, events, event_count, event_cap, event_ply, nonstd, cy, event_env0);
    else
        act_team_kernel<false><<<grid, kAB * 4, 0, stream>>>(envs, n, env_id0, seed, q90, eps_thr, train_done, actions_out, (ActTransition*)ring, ring_cap,
                                                            ring_pos, stats, events, event_count, event_cap, event_ply, nonstd, cy, event_env0);
    ++g_launches;
    return cudaGetLastError();
}

}  // namespace xq
