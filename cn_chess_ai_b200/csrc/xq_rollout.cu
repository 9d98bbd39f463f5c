// xq_rollout.cu -- slot-parallel fused random-policy rollout (the headline env kernel).
//
// Mapping: one CTA = 32 boards x 16 piece slots = 512 threads; thread (slot s, board b) =
// tid s*32+b, so WARP s owns slot s of 32 boards and LANE b owns board b.  Slot -> piece type is
// static (no promotion in Xiangqi), hence every warp runs ONE piece type's move generator:
// no SIMT divergence on piece type (the v1 thread-per-board kernel ran at ~3/32 active lanes).
// Each thread keeps the squares of its Red and Black slot, and a replicated copy of the board's
// 90-bit occupancy bitboards (red, black row-major; all col-major) in registers; sliders are O(1)
// (xq_bitboard.cuh).  Per ply the 16 threads of a board exchange 3 tiny messages through shared
// memory, laid out [item][board] so that lane == bank (conflict-free):
//   A  every slot publishes (square, #moves)                         -> __syncthreads
//   B  every thread derives the reference-order prefix of its slot (square-ordered, byte-SIMD
//      compare + dp4a), the list size n and k = idx31 % n; the owning slot decodes the k-th
//      destination and publishes (from,to)                           -> __syncthreads
//   C  every thread applies the move to its replicated state; the captured slot publishes its
//      piece value                                                   -> __syncthreads
//   D  scores / material / reward / terminal / winner; slot 0 (warp 0) writes trace + stats.
// Records are converted nibble-board <-> slots once per launch; HBM traffic per ply is only the
// optional 8-byte trace record.  Boards whose piece counts exceed a standard set (possible only
// through xq_env_set_boards) are flagged and left to the generic v1 kernel.
//
// Replaces the loop body of ChessAI::train without the network (src/chessai.cpp:96-119):
// getAllValidActions (:347-368) -> list[idx31 % n] -> ChessBoard::movePiece (chessboard.cpp:38-64)
// -> evaluateBoard (:311-345) -> checkGameOver/getWinner (:286-320) -> reset (:95-102).
#include "xq_bitboard.cuh"
#include "xq_common.cuh"

namespace xq {

constexpr int kB = 32;   // boards per CTA
constexpr int kS = 16;   // slots per side = threads per board

__constant__ uint8_t kOpenSq[32] = {0, 8, 1, 7, 2, 6, 3, 5, 4, 19, 25, 27, 29, 31, 33, 35,
                                    81, 89, 82, 88, 83, 87, 84, 86, 85, 64, 70, 54, 56, 58, 60, 62};
__device__ __forceinline__ Bits90 open_red() { return Bits90{0xAA0801FFu, 0x0000000Au, 0x00000000u}; }
__device__ __forceinline__ Bits90 open_black() { return Bits90{0x00000000u, 0x55400000u, 0x03FE0041u}; }
__device__ __forceinline__ Bits90 open_occT() { return Bits90{0x649A1649u, 0x98064980u, 0x0249A164u}; }

// n % d for n < 2^31, 1 <= d <= 128 without a hardware divide: q = umulhi(n, ceil(2^32/d)) is floor(n/d) or one above
// (error < n*d/2^32 < 1), so one conditional fix-up makes the remainder exact.
__device__ __forceinline__ uint32_t mod_magic(uint32_t d) { return (0xFFFFFFFFu / d) + 1u; }   // ceil(2^32/d) (2^32/d for powers of two, d >= 2)
__device__ __forceinline__ uint32_t mod_small(uint32_t n, uint32_t d, uint32_t magic) {          // magic = mod_magic(d), tabulated by the kernel
    const uint32_t q = __umulhi(n, magic);
    const int32_t r = (int32_t)(n - q * d);
    return d == 1 ? 0u : (uint32_t)(r < 0 ? r + (int32_t)d : r);
}

// bytes of qw that are < q (all values <= 127) -> 0xFF, else 0x00
__device__ __forceinline__ uint32_t bytes_lt(uint32_t qw, uint32_t q) {
    const uint32_t t = (q * 0x01010101u + 0x7F7F7F7Fu) - qw;
    return ((t & 0x80808080u) >> 7) * 0xFFu;
}

// -DXQ_TIMELINE: cycles every warp (= piece slot) of block 0 spends in the four phases of a ply and at the three barriers,
// summed over the plies of the launch (profiling builds only; scripts/tl_rollout.py reads them through xq_debug_rollout_phases)
#ifdef XQ_TIMELINE
__device__ long long g_phase[16][8];
#define XQ_PH(k) do { if (blockIdx.x == 0 && lane == 0) { const long long t_ = clock64(); g_phase[slot][k] += t_ - ph_t; ph_t = t_; } } while (0)
#else
#define XQ_PH(k) ((void)0)
#endif

struct SlotState {
    int sq_red, sq_black;
    Bits90 red, black, occT;
    int gen_red, gen_black;
    int move_count, player;
    uint32_t ctr;
    __device__ __forceinline__ void reset(int slot) {
        sq_red = kOpenSq[slot]; sq_black = kOpenSq[16 + slot];
        red = open_red(); black = open_black(); occT = open_occT();
        gen_red = 4; gen_black = 85; move_count = 0; player = RED;
    }
};

__global__ void __launch_bounds__(kB * kS, 2) rollout_slots_kernel(xq_env_rec* __restrict__ envs, int64_t n, uint64_t env_id0, uint64_t seed,
                                                                int n_plies, xq_trace_rec* __restrict__ trace,
                                                                xq_env_stats* __restrict__ stats, uint8_t* __restrict__ nonstd) {
    __shared__ uint8_t s_slot[32 * kB];        // [slot 0..31][board]   load/store conversion
    __shared__ uint32_t s_bb[9 * kB];          // [word][board]
    __shared__ uint32_t s_meta[4 * kB];
    __shared__ uint32_t s_words[12 * kB];      // nibble board words for the final store
    __shared__ uint32_t s_pubq[4 * kB];        // [slot>>2][board] bytes = squares of the mover's slots
    __shared__ uint32_t s_pubc[4 * kB];        // same layout, bytes = move counts
    __shared__ uint32_t s_move[kB];
    __shared__ uint32_t s_cap[2 * kB];
    __shared__ uint8_t s_active[kB];
    __shared__ uint32_t s_magic[XQ_MAX_ACTIONS + 1];   // mod_magic(d) for every possible list size: one shared load instead of a division per ply

    const int tid = threadIdx.x, lane = tid & 31, slot = tid >> 5;
    const int64_t env = (int64_t)blockIdx.x * kB + lane;
    const int type = slot_type(slot);
    if (tid >= 1 && tid <= XQ_MAX_ACTIONS) s_magic[tid] = mod_magic((uint32_t)tid);

    // ---- load: warp 0 converts 32 records to slots + bitboards -------------------------------
    if (slot == 0) {
        bool ok = env < n;
        if (ok) {
            const uint4* rec = reinterpret_cast<const uint4*>(envs + env);
            uint32_t w[12];
#pragma unroll
            for (int i = 0; i < 3; ++i) { const uint4 v = rec[i]; w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w; }
            const uint4 m = rec[3];
            for (int i = 0; i < 32; ++i) s_slot[i * kB + lane] = kDeadSq;
            Bits90 red{0, 0, 0}, black{0, 0, 0}, occT{0, 0, 0};
            uint64_t cnt = 0;   // 4-bit counter per piece code
#pragma unroll
            for (int wi = 0; wi < 12; ++wi) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int s = wi * 8 + i;
                    const int code = (w[wi] >> (4 * i)) & 15;
                    if (s < 90 && code != 0) {
                        const int t = type_of(code);
                        const int ord = (int)((cnt >> (4 * code)) & 15);
                        if (code == 15 || ord >= slot_cap(t)) { ok = false; }
                        else {
                            s_slot[((code >= 8 ? 16 : 0) + slot_base(t) + ord) * kB + lane] = (uint8_t)s;
                            cnt += 1ull << (4 * code);
                            if (code >= 8) black.set(s); else red.set(s);
                            const int r = row_of(s);
                            occT.set(cm_index(r, s - 9 * r));
                        }
                    }
                }
            }
            s_bb[0 * kB + lane] = red.w0; s_bb[1 * kB + lane] = red.w1; s_bb[2 * kB + lane] = red.w2;
            s_bb[3 * kB + lane] = black.w0; s_bb[4 * kB + lane] = black.w1; s_bb[5 * kB + lane] = black.w2;
            s_bb[6 * kB + lane] = occT.w0; s_bb[7 * kB + lane] = occT.w1; s_bb[8 * kB + lane] = occT.w2;
            s_meta[0 * kB + lane] = m.x; s_meta[1 * kB + lane] = m.y; s_meta[2 * kB + lane] = m.z; s_meta[3 * kB + lane] = m.w;
            if (nonstd) nonstd[env] = ok ? 0 : 1;
        }
        s_active[lane] = ok ? 1 : 0;
    }
    __syncthreads();

    const bool active = s_active[lane] != 0;
    SlotState st;
    st.reset(slot);
    int bk_red = 0, bk_black = 0, bk_mat_red = 1480, bk_mat_black = 1480;      // scores / material: used by slot 0 only (see below)
    uint64_t rng_base = 0;
    if (active) {
        st.sq_red = s_slot[slot * kB + lane]; st.sq_black = s_slot[(16 + slot) * kB + lane];
        st.red = Bits90{s_bb[0 * kB + lane], s_bb[1 * kB + lane], s_bb[2 * kB + lane]};
        st.black = Bits90{s_bb[3 * kB + lane], s_bb[4 * kB + lane], s_bb[5 * kB + lane]};
        st.occT = Bits90{s_bb[6 * kB + lane], s_bb[7 * kB + lane], s_bb[8 * kB + lane]};
        st.gen_red = s_slot[8 * kB + lane]; st.gen_black = s_slot[24 * kB + lane];
        const uint32_t m0 = s_meta[0 * kB + lane];
        st.move_count = m0 & 0xFFFF; st.player = (m0 >> 16) & 0xFF;
        bk_red = (int)s_meta[1 * kB + lane]; bk_black = (int)s_meta[2 * kB + lane]; st.ctr = s_meta[3 * kB + lane];
        // material per side from the slots (ChessAI::evaluateBoard :313-341)
        int mr = 0, mb = 0;
        for (int i = 0; i < 16; ++i) {
            const int sc = piece_score(slot_type(i));
            if (s_slot[i * kB + lane] != kDeadSq) mr += sc;
            if (s_slot[(16 + i) * kB + lane] != kDeadSq) mb += sc;
        }
        bk_mat_red = mr; bk_mat_black = mb;
        // a finished board is never stepped (chessai.cpp:90,96): restart it
        if (st.move_count >= XQ_MAX_MOVES || st.gen_red == kDeadSq || st.gen_black == kDeadSq) {
            const uint32_t c = st.ctr; st.reset(slot); st.ctr = c;
            bk_red = bk_black = 0; bk_mat_red = bk_mat_black = 1480;
        }
        rng_base = seed + (env_id0 + (uint64_t)env) * 0x9E3779B97F4A7C15ull;
    }
    uint32_t a_steps = 0, a_games = 0, a_red = 0, a_black = 0, a_capg = 0, a_caps = 0, a_legal = 0;   // per launch: n_plies < 2^24
    long long a_reward = 0;
    uint8_t* pubq8 = reinterpret_cast<uint8_t*>(s_pubq);
    uint8_t* pubc8 = reinterpret_cast<uint8_t*>(s_pubc);
    const int pub_off = (slot >> 2) * (kB * 4) + lane * 4 + (slot & 3);

    // Bookkeeping of a ply (scores, material, reward, statistics, trace record) belongs to slot 0 alone and runs ONE BARRIER LATE:
    // the captured slot publishes its piece value in phase C of ply p, slot 0 picks it up after the first barrier of ply p+1 (or after
    // the closing barrier).  Terminal test and winner need no captured value, so no thread waits for the scores: 2 barriers per ply.
    // pending ply, packed: kind (0 none, 1 a move, 2 no legal action) | mover << 2 | over << 3 | win << 4 | move_count << 8 | total << 16
    uint32_t pend = 0, pend_tr0 = 0;
    int pend_p = 0;
    auto finalize = [&]() {
        if (pend == 0) return;
        const int pend_mover = (pend >> 2) & 1, pend_win = (pend >> 4) & 3, pend_mc = (pend >> 8) & 0xFF, pend_total = pend >> 16;
        const bool pend_over = (pend >> 3) & 1;
        uint32_t capcode = 0; int reward = 0;
        if ((pend & 3) == 1) {
            const uint32_t capw = s_cap[(pend_p & 1) * kB + lane];
            const int capscore = (int)(capw & 0xFFFFu);
            capcode = capw >> 16;
            if (capscore) {
                if (pend_mover == RED) { bk_red += capscore; bk_mat_black -= capscore; } else { bk_black += capscore; bk_mat_red -= capscore; }
            }
            reward = reward_from_material(pend_mover == RED ? bk_mat_red - bk_mat_black : bk_mat_black - bk_mat_red, pend_mc);
            a_steps++; a_legal += pend_total; a_reward += reward; if (capscore) a_caps++;
            if (pend_over) { a_games++; if (pend_win == RED) a_red++; else a_black++; if (pend_mc < XQ_MAX_MOVES) a_capg++; }
        } else {
            a_games++;
        }
        if (trace)   // flags bits 4-7 = captured piece code (published by the captured slot)
            reinterpret_cast<uint2*>(trace)[(int64_t)pend_p * n + env] = make_uint2(pend_tr0 | (capcode << 28), (uint32_t)reward);
        if ((pend & 3) == 2 || pend_over) { bk_red = bk_black = 0; bk_mat_red = bk_mat_black = 1480; }     // ChessBoard::reset
        pend = 0;
    };
#ifdef XQ_TIMELINE
    long long ph_t = clock64();
    if (blockIdx.x == 0 && lane == 0) for (int k = 0; k < 8; ++k) g_phase[slot][k] = 0;
#endif
    for (int p = 0; p < n_plies; ++p) {
        // ---- A: count my moves, publish (square, count) ------------------------------------
        int myq = kDeadSq, cnt = 0;
        uint32_t desc = 0;             // everything phase B needs to decode my k-th move (xq_bitboard.cuh): the generator runs once per ply
        if (active) {
            const int color = st.player;
            myq = color ? st.sq_black : st.sq_red;
            Pos P;
            P.own = color ? st.black : st.red;
            P.occ = Bits90{st.red.w0 | st.black.w0, st.red.w1 | st.black.w1, st.red.w2 | st.black.w2};
            P.occT = st.occT;
            if (myq != kDeadSq) cnt = piece_count_dyn(type, P, myq, color, &desc);         // warp-uniform type
            pubq8[pub_off] = (uint8_t)myq;
            pubc8[pub_off] = (uint8_t)cnt;
            if (slot == 0) s_cap[(p & 1) * kB + lane] = 0;
        }
        XQ_PH(0);
        __syncthreads();
        XQ_PH(1);
        // ---- B: reference-order prefix, list size, draw, owner decodes ------------------------
        int total = 0;
        if (slot == 0 && active) finalize();       // the previous ply's scores / reward / trace (its captured value is visible now)
        if (active) {
            uint32_t prefix = 0, tot = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t qw = s_pubq[j * kB + lane], cw = s_pubc[j * kB + lane];
                prefix = __dp4a(cw & bytes_lt(qw, (uint32_t)myq), 0x01010101u, prefix);
                tot = __dp4a(cw, 0x01010101u, tot);
            }
            total = (int)tot;
            if (total > 0) {
                uint64_t z = rng_base + (uint64_t)st.ctr * 0xD1B54A32D192ED03ull;
                z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
                z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
                z ^= z >> 31;
                const uint32_t k = mod_small((uint32_t)(z >> 33), tot, s_magic[tot]);
                if (cnt > 0 && k >= prefix && k < prefix + (uint32_t)cnt) {
                    const int to = piece_decode_dyn(type, desc, myq, st.player, (int)(k - prefix));
                    s_move[lane] = (uint32_t)myq | ((uint32_t)to << 8);
                }
            }
        }
        XQ_PH(2);
        __syncthreads();
        XQ_PH(3);
        // ---- C: apply the move to the replicated state; the captured slot publishes its piece ---------
        if (active) {
            if (total > 0) {
                const uint32_t mv = s_move[lane];
                const int from = mv & 0xFF, to = (mv >> 8) & 0xFF;
                const int mover = st.player;
                const int oq = mover ? st.sq_red : st.sq_black;          // my piece of the side NOT moving
                if (oq == to) {                                          // it is captured
                    s_cap[(p & 1) * kB + lane] = (uint32_t)piece_score(type) | ((uint32_t)(type + (mover ? 0 : 7)) << 16);   // value | code
                    if (mover) st.sq_red = kDeadSq; else st.sq_black = kDeadSq;
                }
                if (myq == from) { if (mover) st.sq_black = to; else st.sq_red = to; }
                const int fr = row_of(from), tr = row_of(to);
                const Bits90 fm = Bits90::bit(from), tm = Bits90::bit(to);
                st.occT.andnot(Bits90::bit(cm_index(fr, from - 9 * fr)));
                st.occT.or_with(Bits90::bit(cm_index(tr, to - 9 * tr)));
                bool took_general;
                if (mover) { st.black.andnot(fm); st.black.or_with(tm); st.red.andnot(tm); took_general = to == st.gen_red; if (from == st.gen_black) st.gen_black = to; }
                else { st.red.andnot(fm); st.red.or_with(tm); st.black.andnot(tm); took_general = to == st.gen_black; if (from == st.gen_red) st.gen_red = to; }
                st.move_count++; st.player ^= 1; st.ctr++;
                // terminal test and winner need no captured VALUE: every thread decides locally, the scores follow one barrier later
                const bool over = took_general || st.move_count >= XQ_MAX_MOVES;
                if (slot == 0) {
                    // getWinner: colour of the first General in square order
                    const int win = took_general ? mover : (st.gen_red < st.gen_black ? RED : BLACK);
                    pend = 1u | ((uint32_t)mover << 2) | ((uint32_t)over << 3) | ((uint32_t)win << 4) | ((uint32_t)st.move_count << 8) | ((uint32_t)total << 16);
                    pend_p = p;
                    pend_tr0 = (uint32_t)XQ_ACTION(from, to) | ((uint32_t)total << 16) | ((uint32_t)((over ? 1 : 0) | ((over ? win : NOCOLOR) << 1)) << 24);
                }
                if (over) { const uint32_t c = st.ctr; st.reset(slot); st.ctr = c; }
            } else {   // no legal action: the episode loop ends (chessai.cpp:100-103); the slot restarts
                const uint32_t c = st.ctr + 1; st.reset(slot); st.ctr = c;
                if (slot == 0) { pend = 2; pend_p = p; pend_tr0 = (uint32_t)XQ_ACTION_NONE | ((uint32_t)(1 | (NOCOLOR << 1)) << 24); }
            }
        }
        XQ_PH(4);
    }

    // ---- store: slots -> nibble board -----------------------------------------------------------
    if (active) { s_slot[slot * kB + lane] = (uint8_t)st.sq_red; s_slot[(16 + slot) * kB + lane] = (uint8_t)st.sq_black; }
    __syncthreads();
    if (slot == 0) {
        if (active) {
            finalize();                           // the last ply
            for (int i = 0; i < 12; ++i) s_words[i * kB + lane] = 0;
            for (int i = 0; i < 32; ++i) {
                const int q = s_slot[i * kB + lane];
                if (q != kDeadSq) s_words[(q >> 3) * kB + lane] |= (uint32_t)(slot_type(i & 15) + (i >= 16 ? 7 : 0)) << (4 * (q & 7));
            }
            uint4* rec = reinterpret_cast<uint4*>(envs + env);
#pragma unroll
            for (int i = 0; i < 3; ++i)
                rec[i] = make_uint4(s_words[(4 * i) * kB + lane], s_words[(4 * i + 1) * kB + lane], s_words[(4 * i + 2) * kB + lane],
                                    s_words[(4 * i + 3) * kB + lane]);
            const uint32_t flags = s_meta[0 * kB + lane] & 0xFF000000u;
            rec[3] = make_uint4((uint32_t)(st.move_count & 0xFFFF) | ((uint32_t)st.player << 16) | flags, (uint32_t)bk_red, (uint32_t)bk_black, st.ctr);
        }
        if (stats) {
            unsigned long long v[8] = {a_steps, a_games, a_red, a_black, a_capg, a_caps, (unsigned long long)a_reward, a_legal};   // zero-extended counters
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                unsigned long long r = v[i];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xFFFFFFFFu, r, o);
                if (lane == 0 && r != 0) atomicAdd(reinterpret_cast<unsigned long long*>(stats) + i, r);
            }
        }
    }
}

cudaError_t launch_rollout_slots(xq_env_rec* envs, int64_t n, uint64_t env_id0, uint64_t seed, int n_plies, xq_trace_rec* trace,
                                 xq_env_stats* stats, uint8_t* nonstd, cudaStream_t stream) {
    const unsigned grid = (unsigned)((n + kB - 1) / kB);
    rollout_slots_kernel<<<grid, kB * kS, 0, stream>>>(envs, n, env_id0, seed, n_plies, trace, stats, nonstd);
    ++g_launches;
    return cudaGetLastError();
}

}  // namespace xq

#ifdef XQ_TIMELINE
extern "C" int xq_debug_rollout_phases(long long* out_host) {   // profiling builds only: [16 slots][8]
    if (cudaDeviceSynchronize() != cudaSuccess) return XQ_ERR_CUDA;
    return cudaMemcpyFromSymbol(out_host, xq::g_phase, sizeof(long long) * 16 * 8) == cudaSuccess ? XQ_OK : XQ_ERR_CUDA;
}
#endif
