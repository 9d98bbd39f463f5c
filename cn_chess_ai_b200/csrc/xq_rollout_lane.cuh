// xq_rollout_lane.cuh -- one ply of the fused random-policy rollout with the WHOLE board in ONE thread's registers.
//
// Why: the team kernel (xq_rollout_team.cuh, 4 threads per board) repeats the selection, the apply step and the bookkeeping in each
// of its 4 threads and needs two CTA barriers per ply: 87 warp-instructions per env step, ~2/3 of them replicated work, and at 1M
// envs -- where it issues 70 % of the slots -- the barrier is its largest stall.  Here a thread owns all 32 piece slots of a board:
// nothing is replicated, nothing is exchanged, there is no barrier and no shared-memory traffic in the ply (one table read), and
// because slot -> piece type is static (no promotion in Xiangqi) position i holds the same piece type in every lane of a warp:
// all 16 generators are specialised at compile time and a warp never diverges on type.  The 16 generators of a ply are independent
// instruction streams, which is what a lone warp on a scheduler needs to issue back to back.
//
// State is relative to the side to move (own / opp, swapped after each ply) as in the team kernel.  The 16 squares of a side sit
// one per byte in 4 words in POSITION order (127 = captured):
//   word 0: Chariot Chariot Cannon Cannon | word 1: Horse Horse Elephant Elephant | word 2: Advisor Advisor General Soldier |
//   word 3: Soldier x 4
// with the move counts of a ply and the leapers' direction masks in words of the same shape, so "the piece on square s" is a
// byte compare and every per-piece quantity comes out with one dp4a per word.
//
// The k-th action of ChessAI::getAllValidActions' order (squares row-major, src/chessai.cpp:347-368) is found without sorting:
// g(s) = number of actions of pieces on squares <= s is four dp4a on the (squares, counts) words, and the owner of action k sits
// on the smallest s with g(s) > k -- a 7-step bisection over the 90 squares.
//
// Host-compilable: tests/hostsim runs it board by board and diffs every ply against the oracle before any GPU time.
#pragma once
#include <stdint.h>
#include <string.h>

#include "xq_rollout_team.cuh"

namespace xq {

// position -> slot of xq_bitboard.cuh (0,1 Chariot | 2,3 Horse | 4,5 Elephant | 6,7 Advisor | 8 General | 9,10 Cannon | 11..15 Soldier)
constexpr int lane_pos_slot_c(int pos) {
    constexpr int8_t t[16] = {0, 1, 9, 10, 2, 3, 4, 5, 6, 7, 8, 11, 12, 13, 14, 15};
    return t[pos];
}
XQ_HD int lane_pos_slot(int pos) { return (int)((0xFEDCB8765432A910ull >> (4 * pos)) & 15u); }      // the same table, one nibble per position
constexpr uint32_t lane_word_c(int what, int w) {      // byte i = property of position 4 w + i: 0 value / 5 (getPieceScore, src/chessboard.cpp:443-454),
    uint32_t r = 0;                                    // 1 PieceType, 2 / 3 opening square of Red / Black (initializeBoard, :13-28)
    for (int i = 0; i < 4; ++i) {
        const int slot = lane_pos_slot_c(4 * w + i), type = slot_type_c(slot);
        const int v = what == 0 ? piece_score_c(type) / 5 : (what == 1 ? type : open_sq_c((what == 3 ? 16 : 0) + slot));
        r |= (uint32_t)v << (8 * i);
    }
    return r;
}
template <int WHAT>
XQ_HD uint32_t lane_word(int w) {
    constexpr uint32_t c0 = lane_word_c(WHAT, 0), c1 = lane_word_c(WHAT, 1), c2 = lane_word_c(WHAT, 2), c3 = lane_word_c(WHAT, 3);
    return w == 0 ? c0 : (w == 1 ? c1 : (w == 2 ? c2 : c3));
}
static_assert(lane_word_c(0, 0) == 0x09091212u && lane_word_c(0, 2) == 0x02C80404u && lane_word_c(1, 2) == 0x07010202u, "position tables");
static_assert(lane_word_c(2, 0) == 0x19130800u && lane_word_c(3, 1) == 0x57535852u && lane_word_c(3, 3) == 0x3E3C3A38u, "opening squares");

struct LaneState {
    uint32_t own_sq[4], opp_sq[4];
    Bits90 own, opp, occT;
    int move_count, player;
    uint32_t ctr;
    int red, black, mat_red, mat_black;      // ChessBoard::redScore / blackScore; material per side (ChessAI::evaluateBoard :313-341)
};
struct LaneStats {
    uint32_t steps, games, red, black, capg, caps, legal;
    long long reward;
};

XQ_HD void lane_reset_board(LaneState& st) {                 // ChessBoard::reset (src/chessboard.cpp:95-102); the RNG counter is kept
#pragma unroll
    for (int w = 0; w < 4; ++w) { st.own_sq[w] = lane_word<2>(w); st.opp_sq[w] = lane_word<3>(w); }
    st.own = team_open_red(); st.opp = team_open_black(); st.occT = team_open_occT();
    st.move_count = 0; st.player = RED;
    st.red = st.black = 0; st.mat_red = st.mat_black = 1480;
}

// one trace record: an 8-byte store, coalesced across the lanes of a warp ([ply][env] layout)
XQ_HD void lane_trace_store(xq_trace_rec* t, uint32_t w0, uint32_t w1) {
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<uint2*>(t) = make_uint2(w0, w1);
#else
    uint32_t* u = reinterpret_cast<uint32_t*>(t); u[0] = w0; u[1] = w1;
#endif
}

// g(t) = number of actions of the pieces on squares < t
XQ_HD uint32_t lane_actions_below(const uint32_t (&sq)[4], const uint32_t (&cw)[4], uint32_t t) {
    const uint32_t base = t * 0x01010101u + 0x7F7F7F7Fu;      // byte j of (base - sq) has bit 7 set iff square_j < t (bytes <= 127: no borrow crosses a byte)
    uint32_t acc = 0;
#pragma unroll
    for (int w = 0; w < 4; ++w) acc = dp4a_u(cw[w], (base - sq[w]) & 0x80808080u, acc);
    return acc >> 7;
}

// Every piece of the side to move: move counts (cw, one byte per position), the sliders' descriptors (sdesc, xq_bitboard.cuh) and the
// leapers' direction masks (dw, one byte per position; dw[0] = 0).  A captured piece (square 127) reads garbage bits, its count is discarded.
// P: a view of the bitboards of the position seen by the side to move (xq_bitboard.cuh: MemView in the kernels -- the thread's slice of
// shared memory, written by view_store after every change of the board)
template <class V>
XQ_HD void lane_movegen(const uint32_t (&own_sq)[4], const V& P, int color, const uint32_t* geo,
                        uint32_t (&sdesc)[4], uint32_t (&cw)[4], uint32_t (&dw)[4]) {
    const uint32_t* gq = geo + color * 128;      // geometry table of the side to move (xq_bitboard.cuh: geo_entry), entries 90..127 = 0
    {
        int c[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int q = (int)((own_sq[0] >> (8 * i)) & 0xFFu);
            const int n = slider_desc_v(P, q, i >= 2, &sdesc[i]);      // Chariots, then Cannons :198-246
            c[i] = q == kDeadSq ? 0 : n;
        }
        cw[0] = (uint32_t)c[0] | ((uint32_t)c[1] << 8) | ((uint32_t)c[2] << 16) | ((uint32_t)c[3] << 24);
        dw[0] = 0;
    }
#pragma unroll
    for (int w = 1; w < 4; ++w) {
        uint32_t m[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int q = (int)((own_sq[w] >> (8 * i)) & 0xFFu);
            const int pos = 4 * w + i;
            const uint32_t g = gq[q];                                // 0 for a captured piece (square 127): its mask is empty
            uint32_t v;
            if (pos < 6) v = horse_mask_g(P, q, g);                  // :248-263
            else if (pos < 8) v = elephant_mask_g(P, q, g);          // :179-196
            else if (pos < 10) v = advisor_mask_g(P, q, g);          // :162-177
            else if (pos == 10) v = general_mask_g(P, q, g);         // :149-160
            else v = soldier_mask_g(P, q, color, g);                 // :265-283
            m[i] = v;
        }
        dw[w] = m[0] | (m[1] << 8) | (m[2] << 16) | (m[3] << 24);
        cw[w] = (uint32_t)popc32(m[0]) | ((uint32_t)popc32(m[1]) << 8) | ((uint32_t)popc32(m[2]) << 16) | ((uint32_t)popc32(m[3]) << 24);
    }
}
// destination offsets of a leaper at position `pos`, one signed byte per direction in the order of generate*Moves
// (src/chessboard.cpp:150,163,180,249,267-281); `hi` = directions 4..7 of a Horse
XQ_HD uint32_t lane_dir_table(int pos, int color, uint32_t* hi) {
    *hi = (pos == 4 || pos == 5) ? 0xEDEF1113u : 0u;               // Horse: 11,7,-7,-11 | 19,17,-17,-19
    return (pos == 4 || pos == 5) ? 0xF5F9070Bu
         : ((pos == 6 || pos == 7) ? 0xECF01014u                    // Elephant: 20,16,-16,-20
         : ((pos == 8 || pos == 9) ? 0xF6F8080Au                    // Advisor: 10,8,-8,-10
         : (pos == 10 ? 0xFF01F709u                                 // General: 9,-9,1,-1
         : (0x0001FF09u ^ (color ? 0xFEu : 0u)))));                 // Soldier: 9,-1,1; a Black Soldier moves towards row 0
}

// ChessAI::getAllValidActions(side to move) (src/chessai.cpp:347-368) from the movegen words: emit(index, action, live) for every SLOT
// (piece, direction, distance) a piece could use; `live` slots are the actions, `index` their place in the reference-ordered list (squares
// row-major, then the direction order of generate*Moves, distance ascending along a ray); returns the list size.
// Straight-line: every lane of a warp walks the same 198 slots with predicated stores -- no data-dependent loop, nothing diverges
// (a first version looped `for d = 1..empties` per ray: a warp then ran the longest slide of its 32 boards around a 13-instruction body).
template <class EMIT>
XQ_HD int lane_emit_actions(const uint32_t (&own_sq)[4], int color, const uint32_t (&sdesc)[4], const uint32_t (&cw)[4], const uint32_t (&dw)[4], EMIT&& emit) {
    uint32_t tot = 0;
#pragma unroll
    for (int w = 0; w < 4; ++w) tot = dp4a_u(cw[w], 0x01010101u, tot);
#pragma unroll
    for (int i = 0; i < 4; ++i) {              // Chariots and Cannons: per ray the empty squares, then the capture
        const int sq = (int)((own_sq[0] >> (8 * i)) & 0xFFu);
        const bool alive = ((cw[0] >> (8 * i)) & 0xFFu) != 0;      // a captured slider's descriptor is garbage
        int off = (int)lane_actions_below(own_sq, cw, (uint32_t)sq);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int e = alive ? (int)((sdesc[i] >> (8 * k)) & 15u) : 0, capdist = alive ? (int)((sdesc[i] >> (8 * k + 4)) & 15u) : 0;
            const int step = k == 0 ? 1 : (k == 1 ? -1 : (k == 2 ? 9 : -9));
            constexpr int kMaxSlide[4] = {8, 8, 9, 9};
#pragma unroll
            for (int d = 1; d <= kMaxSlide[k]; ++d) emit(off + d - 1, (int)XQ_ACTION(sq, sq + step * d), d <= e);
            emit(off + e, (int)XQ_ACTION(sq, sq + step * capdist), capdist != 0);
            off += e + (capdist ? 1 : 0);
        }
    }
#pragma unroll
    for (int pos = 4; pos < 16; ++pos) {
        const int sq = (int)((own_sq[pos >> 2] >> (8 * (pos & 3))) & 0xFFu);
        const uint32_t m = (dw[pos >> 2] >> (8 * (pos & 3))) & 0xFFu;      // 0 for a captured piece
        int off = (int)lane_actions_below(own_sq, cw, (uint32_t)sq);
        uint32_t hi;
        const uint32_t lo = lane_dir_table(pos, color, &hi);
        const int ndir = pos < 6 ? 8 : (pos < 11 ? 4 : 3);
#pragma unroll
        for (int k = 0; k < ndir; ++k) {
            const bool live = (m >> k) & 1u;
            emit(off, (int)XQ_ACTION(sq, sq + (int)(int8_t)(uint8_t)((k < 4 ? lo : hi) >> (8 * (k & 3)))), live);
            off += live ? 1 : 0;
        }
    }
    return (int)tot;
}

// a > b as floats  <=>  lane_ordered_key(a) > lane_ordered_key(b) as unsigned integers (finite values; -0 == +0); 0 is below every float
XQ_HD uint32_t lane_ordered_key(float v) {
#if defined(__CUDA_ARCH__)
    const uint32_t u = __float_as_uint(v);
#else
    uint32_t u; memcpy(&u, &v, 4);
#endif
    const uint32_t z = (u << 1) == 0u ? 0u : u;
    return (z & 0x80000000u) ? ~z : (z | 0x80000000u);
}

// DQN::selectAction's greedy branch (src/dqn.cpp:39-52) over ChessAI::getAllValidActions' order without the list: the FIRST action maximising
// Q[action.to] (strict >).  Every piece keeps the first maximum over its own actions in generator order (the 198 (piece, direction, distance)
// slots walked straight-line with predicated Q reads, as in lane_emit_actions); the list orders pieces by square, so the winner is the piece with
// the largest Q and, among equals, the lowest square.  q(to) reads Q(s)[to].  Requires a non-empty list; returns from | to << 8.
template <class QGET>
XQ_HD uint32_t lane_select_greedy(const uint32_t (&own_sq)[4], int color, const uint32_t (&sdesc)[4], const uint32_t (&cw)[4], const uint32_t (&dw)[4], QGET&& q) {
    uint32_t best_key = 0, best_sq = 127, best_to = 0;
    auto offer = [&](bool any, float best, uint32_t sq, uint32_t to) {      // a piece's best against the best so far: larger Q, or equal Q on a lower square
        const uint32_t key = any ? lane_ordered_key(best) : 0u;
        const bool better = key > best_key || (key == best_key && any && sq < best_sq);
        best_key = better ? key : best_key; best_sq = better ? sq : best_sq; best_to = better ? to : best_to;
    };
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int sq = (int)((own_sq[0] >> (8 * i)) & 0xFFu);
        const bool alive = ((cw[0] >> (8 * i)) & 0xFFu) != 0;
        bool any = false; float best = 0.f; uint32_t bto = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int e = alive ? (int)((sdesc[i] >> (8 * k)) & 15u) : 0, capdist = alive ? (int)((sdesc[i] >> (8 * k + 4)) & 15u) : 0;
            const int step = k == 0 ? 1 : (k == 1 ? -1 : (k == 2 ? 9 : -9));
            constexpr int kMaxSlide[4] = {8, 8, 9, 9};
#pragma unroll
            for (int d = 1; d <= kMaxSlide[k]; ++d) {
                const bool live = d <= e;
                const int to = sq + step * d;
                const float v = live ? q(to) : 0.f;
                const bool take = live && (!any || v > best);
                best = take ? v : best; bto = take ? (uint32_t)to : bto; any |= live;
            }
            const bool live = capdist != 0;
            const int to = sq + step * capdist;
            const float v = live ? q(to) : 0.f;
            const bool take = live && (!any || v > best);
            best = take ? v : best; bto = take ? (uint32_t)to : bto; any |= live;
        }
        offer(any, best, (uint32_t)sq, bto);
    }
#pragma unroll
    for (int pos = 4; pos < 16; ++pos) {
        const int sq = (int)((own_sq[pos >> 2] >> (8 * (pos & 3))) & 0xFFu);
        const uint32_t m = (dw[pos >> 2] >> (8 * (pos & 3))) & 0xFFu;
        uint32_t hi;
        const uint32_t lo = lane_dir_table(pos, color, &hi);
        const int ndir = pos < 6 ? 8 : (pos < 11 ? 4 : 3);
        bool any = false; float best = 0.f; uint32_t bto = 0;
#pragma unroll
        for (int k = 0; k < ndir; ++k) {
            const bool live = (m >> k) & 1u;
            const int to = sq + (int)(int8_t)(uint8_t)((k < 4 ? lo : hi) >> (8 * (k & 3)));
            const float v = live ? q(to) : 0.f;
            const bool take = live && (!any || v > best);
            best = take ? v : best; bto = take ? (uint32_t)to : bto; any |= live;
        }
        offer(any, best, (uint32_t)sq, bto);
    }
    return best_sq | (best_to << 8);
}

// the k-th action of the reference-ordered list (DQN::selectAction's exploring branch, src/dqn.cpp:30-34; the random policy of the rollout):
// bisection over the squares on g(s) = actions of pieces on squares <= s, then the decode from the owner's descriptor.  k < tot.  from | to << 8.
XQ_HD uint32_t lane_select_kth(const uint32_t (&own_sq)[4], int color, const uint32_t (&sdesc)[4], const uint32_t (&cw)[4], const uint32_t (&dw)[4], uint32_t k, uint32_t tot) {
    uint32_t lo = 0, hi = XQ_SQUARES - 1, g_hi = tot;
#pragma unroll
    for (int it = 0; it < 7; ++it) {
        const uint32_t mid = (lo + hi) >> 1;
        const uint32_t g = lane_actions_below(own_sq, cw, mid + 1);
        const bool up = g > k;
        hi = up ? mid : hi; g_hi = up ? g : g_hi; lo = up ? lo : mid + 1;
    }
    const int from = (int)hi;
    uint32_t zb[4], cnt_hit = 0, lmask = 0;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        zb[w] = ((0x80808080u - (own_sq[w] ^ ((uint32_t)from * 0x01010101u))) & 0x80808080u) >> 7;
        cnt_hit = dp4a_u(cw[w], zb[w], cnt_hit);
        if (w) lmask = dp4a_u(dw[w], zb[w], lmask);
    }
    const uint32_t want = k - (g_hi - cnt_hit);
    const uint32_t hdesc = (zb[0] & 0x00000001u) ? sdesc[0] : ((zb[0] & 0x00000100u) ? sdesc[1] : ((zb[0] & 0x00010000u) ? sdesc[2] : sdesc[3]));
    const int to_s = slider_decode(hdesc, from, (int)want);
    const uint32_t flip = color ? 0xFEu : 0u;
    uint32_t tlo = 0x0001FF09u ^ flip, thi = 0u;
    tlo = (zb[1] & 0x00000101u) ? 0xF5F9070Bu : tlo; thi = (zb[1] & 0x00000101u) ? 0xEDEF1113u : 0u;
    tlo = (zb[1] & 0x01010000u) ? 0xECF01014u : tlo;
    tlo = (zb[2] & 0x00000101u) ? 0xF6F8080Au : tlo;
    tlo = (zb[2] & 0x00010000u) ? 0xFF01F709u : tlo;
    const int dk = nth_set_bit8(lmask, (int)want & 7);
    const int to_l = from + (int)(int8_t)(uint8_t)((((uint64_t)thi << 32) | tlo) >> (8 * dk));
    return (uint32_t)from | ((uint32_t)(zb[0] ? to_s : to_l) << 8);
}

// One ply of ChessAI::train's loop body without the network (src/chessai.cpp:96-119) on one board.
// magic[d] = team_mod_magic(d) for d = 1..128; geo = the geometry table (kGeoWords words, geo_entry).  trace (may be null) -> the record of this ply.
// vm / vstride: this board's slice of the view memory (kViewWords words, xq_bitboard.cuh), holding the bitboards of st on entry and on return.
XQ_HD void lane_ply(LaneState& st, LaneStats& a, uint64_t rng_base, const uint32_t* magic, const uint32_t* geo, uint32_t* vm, int vstride,
                    xq_trace_rec* trace) {
    const int color = st.player;
    uint32_t sdesc[4], cw[4], dw[4];
    lane_movegen(st.own_sq, MemView{vm, vstride}, color, geo, sdesc, cw, dw);
    uint32_t tot = 0;
#pragma unroll
    for (int w = 0; w < 4; ++w) tot = dp4a_u(cw[w], 0x01010101u, tot);
    const uint32_t draw = team_draw(rng_base, st.ctr);
    st.ctr++;
    if (tot == 0) {     // no legal action: the episode loop ends (chessai.cpp:100-103) and the slot restarts
        a.games++;
        if (trace) lane_trace_store(trace, (uint32_t)XQ_ACTION_NONE | ((uint32_t)(1 | (NOCOLOR << 1)) << 24), 0u);
        lane_reset_board(st);
        view_store(vm, vstride, st.own, st.opp, st.occT);
        return;
    }
    // ---- the k-th action in reference order: the smallest square s with g(s + 1) > k ----
    const uint32_t k = team_mod(draw, tot, magic[tot]);
    uint32_t lo = 0, hi = XQ_SQUARES - 1, g_hi = tot;
#pragma unroll
    for (int it = 0; it < 7; ++it) {
        const uint32_t mid = (lo + hi) >> 1;
        const uint32_t g = lane_actions_below(st.own_sq, cw, mid + 1);
        const bool up = g > k;
        hi = up ? mid : hi; g_hi = up ? g : g_hi; lo = up ? lo : mid + 1;
    }
    const int from = (int)hi;
    // the piece on `from`: one byte of own_sq matches; its count, slider descriptor / leaper mask and direction table
    uint32_t zb[4], cnt_hit = 0, lmask = 0;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        zb[w] = ((0x80808080u - (st.own_sq[w] ^ ((uint32_t)from * 0x01010101u))) & 0x80808080u) >> 7;      // 1 in the matching byte
        cnt_hit = dp4a_u(cw[w], zb[w], cnt_hit);
        if (w) lmask = dp4a_u(dw[w], zb[w], lmask);
    }
    const uint32_t want = k - (g_hi - cnt_hit);
    const uint32_t hdesc = (zb[0] & 0x00000001u) ? sdesc[0] : ((zb[0] & 0x00000100u) ? sdesc[1] : ((zb[0] & 0x00010000u) ? sdesc[2] : sdesc[3]));
    const int to_s = slider_decode(hdesc, from, (int)want);
    // destination offsets per direction, one signed byte each, in the order of generate*Moves (src/chessboard.cpp:150,163,180,249,267-281)
    const uint32_t flip = color ? 0xFEu : 0u;                       // a Black Soldier moves towards row 0: 9 -> -9
    uint32_t tlo = 0x0001FF09u ^ flip, thi = 0u;                    // Soldier: 9, -1, 1
    tlo = (zb[1] & 0x00000101u) ? 0xF5F9070Bu : tlo; thi = (zb[1] & 0x00000101u) ? 0xEDEF1113u : 0u;       // Horse: 11,7,-7,-11 | 19,17,-17,-19
    tlo = (zb[1] & 0x01010000u) ? 0xECF01014u : tlo;                // Elephant: 20,16,-16,-20
    tlo = (zb[2] & 0x00000101u) ? 0xF6F8080Au : tlo;                // Advisor: 10,8,-8,-10
    tlo = (zb[2] & 0x00010000u) ? 0xFF01F709u : tlo;                // General: 9,-9,1,-1
    const int dk = nth_set_bit8(lmask, (int)want & 7);
    const int to_l = from + (int)(int8_t)(uint8_t)((((uint64_t)thi << 32) | tlo) >> (8 * dk));
    const int to = zb[0] ? to_s : to_l;

    // ---- ChessBoard::movePiece (src/chessboard.cpp:38-64) ----
    const int mover = color;
    uint32_t cap5 = 0, captype = 0, new_opp[4], new_own[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        const uint32_t z = (0x80808080u - (st.opp_sq[w] ^ ((uint32_t)to * 0x01010101u))) & 0x80808080u;      // an enemy piece on `to`
        cap5 = dp4a_u(lane_word<0>(w), z >> 7, cap5);
        captype = dp4a_u(lane_word<1>(w), z >> 7, captype);
        new_opp[w] = st.opp_sq[w] | (z >> 7) * 0x7Fu;               // captured: square 127
        const uint32_t m8 = zb[w] * 0xFFu;
        new_own[w] = (st.own_sq[w] & ~m8) | (((uint32_t)to * 0x01010101u) & m8);
    }
    const int capscore = (int)cap5 * 5;
    const uint32_t capcode = captype ? captype + (mover ? 0u : 7u) : 0u;
    const int fr = row_of(from), tr = row_of(to);
    const Bits90 fm = bit_mask(from), tm = bit_mask(to);
    const Bits90 cf = bit_mask(cm_index(fr, from - 9 * fr)), ct = bit_mask(cm_index(tr, to - 9 * tr));
    const Bits90 own{(st.own.w0 & ~fm.w0) | tm.w0, (st.own.w1 & ~fm.w1) | tm.w1, (st.own.w2 & ~fm.w2) | tm.w2};
    const Bits90 opp{st.opp.w0 & ~tm.w0, st.opp.w1 & ~tm.w1, st.opp.w2 & ~tm.w2};
    const Bits90 occT{(st.occT.w0 & ~cf.w0) | ct.w0, (st.occT.w1 & ~cf.w1) | ct.w1, (st.occT.w2 & ~cf.w2) | ct.w2};
    if (capscore) {                                                 // :51-58
        if (mover == RED) { st.red += capscore; st.mat_black -= capscore; } else { st.black += capscore; st.mat_red -= capscore; }
        a.caps++;
    }
    const int mc = st.move_count + 1;
    const bool took_general = captype == GENERAL;
    const bool over = took_general | (mc >= XQ_MAX_MOVES);          // checkGameOver, :286-309
    // getWinner: colour of the first General in square order
    const int gen_own = (int)((new_own[2] >> 16) & 0xFFu), gen_opp = (int)((new_opp[2] >> 16) & 0xFFu);
    const int gen_red = mover ? gen_opp : gen_own, gen_black = mover ? gen_own : gen_opp;
    const int win = took_general ? mover : (gen_red < gen_black ? RED : BLACK);
    const int reward = reward_from_material(mover == RED ? st.mat_red - st.mat_black : st.mat_black - st.mat_red, mc);      // ChessAI::evaluateBoard
    a.steps++; a.legal += tot; a.reward += reward;
    if (over) { a.games++; if (win == RED) a.red++; else a.black++; if (mc < XQ_MAX_MOVES) a.capg++; }
    if (trace) {   // action | n_legal << 16 | flags << 24 (bit 0 done, bits 1-2 winner, bits 4-7 captured piece code), reward
        lane_trace_store(trace, (uint32_t)XQ_ACTION(from, to) | (tot << 16) | ((uint32_t)((over ? 1 : 0) | ((over ? win : NOCOLOR) << 1)) << 24) | (capcode << 28),
                         (uint32_t)reward);
    }
    // the other side is to move -- or, after the last move of a game, Red from the opening (ChessBoard::reset): selects, no branch
    const Bits90 o_red = team_open_red(), o_black = team_open_black(), o_occT = team_open_occT();
#pragma unroll
    for (int w = 0; w < 4; ++w) { st.own_sq[w] = over ? lane_word<2>(w) : new_opp[w]; st.opp_sq[w] = over ? lane_word<3>(w) : new_own[w]; }
    st.own = Bits90{over ? o_red.w0 : opp.w0, over ? o_red.w1 : opp.w1, over ? o_red.w2 : opp.w2};
    st.opp = Bits90{over ? o_black.w0 : own.w0, over ? o_black.w1 : own.w1, over ? o_black.w2 : own.w2};
    st.occT = Bits90{over ? o_occT.w0 : occT.w0, over ? o_occT.w1 : occT.w1, over ? o_occT.w2 : occT.w2};
    st.move_count = over ? 0 : mc; st.player = over ? RED : (mover ^ 1);
    st.red = over ? 0 : st.red; st.black = over ? 0 : st.black;
    st.mat_red = over ? 1480 : st.mat_red; st.mat_black = over ? 1480 : st.mat_black;
    view_store(vm, vstride, st.own, st.opp, st.occT);
}

// ---- record <-> state (once per launch) --------------------------------------------------------------------------------------------
// slot[32]: squares of the 32 piece slots (team_unpack_record), kDeadSq = captured.  A finished board is never stepped
// (chessai.cpp:90,96): it restarts from the opening.
template <class GET>
XQ_HD void lane_load(LaneState& st, GET&& slot, const Bits90& red, const Bits90& black, const Bits90& occT, int move_count, int player,
                     int red_score, int black_score, uint32_t ctr) {
    uint32_t wr[4] = {0, 0, 0, 0}, wb[4] = {0, 0, 0, 0};
    int mr = 0, mb = 0;
#pragma unroll
    for (int pos = 0; pos < 16; ++pos) {
        const int s = lane_pos_slot(pos);
        const int qr = slot(s), qb = slot(16 + s);
        wr[pos >> 2] |= (uint32_t)qr << (8 * (pos & 3));
        wb[pos >> 2] |= (uint32_t)qb << (8 * (pos & 3));
        const int sc = piece_score(slot_type(s));
        mr += qr != kDeadSq ? sc : 0; mb += qb != kDeadSq ? sc : 0;
    }
    st.ctr = ctr;
    if (move_count >= XQ_MAX_MOVES || slot(8) == kDeadSq || slot(24) == kDeadSq) { lane_reset_board(st); return; }
    const bool redp = player == RED;
#pragma unroll
    for (int w = 0; w < 4; ++w) { st.own_sq[w] = redp ? wr[w] : wb[w]; st.opp_sq[w] = redp ? wb[w] : wr[w]; }
    st.own = redp ? red : black; st.opp = redp ? black : red; st.occT = occT;
    st.move_count = move_count; st.player = player;
    st.red = red_score; st.black = black_score; st.mat_red = mr; st.mat_black = mb;
}
// the 12 nibble words of the record from the state
XQ_HD void lane_store_words(const LaneState& st, uint32_t (&words)[12]) {
#pragma unroll
    for (int i = 0; i < 12; ++i) words[i] = 0;
    const bool redp = st.player == RED;
#pragma unroll
    for (int pos = 0; pos < 16; ++pos) {
        const int type = slot_type(lane_pos_slot(pos));
        const int qr = (int)(((redp ? st.own_sq : st.opp_sq)[pos >> 2] >> (8 * (pos & 3))) & 0xFFu);
        const int qb = (int)(((redp ? st.opp_sq : st.own_sq)[pos >> 2] >> (8 * (pos & 3))) & 0xFFu);
#pragma unroll
        for (int i = 0; i < 12; ++i) {      // no dynamically indexed local array: 12 predicated ORs per piece, once per launch
            words[i] |= (qr != kDeadSq && (qr >> 3) == i) ? (uint32_t)type << (4 * (qr & 7)) : 0u;
            words[i] |= (qb != kDeadSq && (qb >> 3) == i) ? (uint32_t)(type + 7) << (4 * (qb & 7)) : 0u;
        }
    }
}

// the same through a scratch slice of memory laid out [word][thread] (the view memory, free after the last ply): one read-modify-write
// per piece on a run-time word index instead of 12 predicated ORs per piece (1,500 -> ~250 instructions per launch)
XQ_HD void lane_store_words_mem(const LaneState& st, uint32_t* m, int stride, uint32_t (&words)[12]) {
#pragma unroll
    for (int i = 0; i < 12; ++i) m[i * stride] = 0u;
    const bool redp = st.player == RED;
#pragma unroll
    for (int pos = 0; pos < 16; ++pos) {
        const uint32_t type = (uint32_t)slot_type(lane_pos_slot(pos));
        const uint32_t qr = ((redp ? st.own_sq : st.opp_sq)[pos >> 2] >> (8 * (pos & 3))) & 0xFFu;
        const uint32_t qb = ((redp ? st.opp_sq : st.own_sq)[pos >> 2] >> (8 * (pos & 3))) & 0xFFu;
        // a captured piece (square 127) lands in word 15 of the slice, beyond the 12 words of the record
        m[(qr >> 3) * stride] |= type << (4 * (qr & 7));
        m[(qb >> 3) * stride] |= (type + 7u) << (4 * (qb & 7));
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) words[i] = m[i * stride];
}

}  // namespace xq
