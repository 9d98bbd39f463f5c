// xq_act_team.cuh -- DQN::selectAction for a TEAM of 4 threads per board (the team of xq_rollout_team.cuh): the selection half of
// one self-play ply.  Host-compilable: tests/hostsim runs the 4 threads of a board phase by phase and diffs the chosen action
// against the oracle's selectAction restatement on the same Q values and draws.
//
// Reference (src/dqn.cpp:24-56 over ChessAI::getAllValidActions, src/chessai.cpp:347-368):
//   coin = rand()/RAND_MAX < eps  -> validActions[rand() % n]
//   else                          -> the FIRST action of the reference-ordered list maximising Q[action.to] (strict >)
// The list is never materialised.  Explore = the k-th action, found like in the rollout kernel (reference-order prefix of the
// per-piece move counts).  Exploit: every thread walks the destinations of its 4 pieces in generator order (decoded from the
// slider descriptor / leaper mask of phase A), keeps the first maximum of Q[to] per piece, and publishes it as an order-preserving
// integer key; the list orders pieces by square, so the winner is the piece with the largest key and, among equals, the lowest
// square -- its owner decodes the move.
#pragma once
#include <string.h>

#include "xq_rollout_team.cuh"

namespace xq {

constexpr int kQStride = 96;          // Q(s)[0..95] per env, row-major (dqn_q90_device)

template <int KB>
struct ActShared {
    uint32_t key[16 * KB];            // [role * 4 + position][board]: ordered-integer image of the piece's best Q (0: no move)
    float qt[kQStride * (KB + 1)];    // Q tile [to][board], one word of padding per row: the transposing store is conflict-free
};

#if defined(__CUDA_ARCH__)
XQ_HD uint32_t float_bits(float v) { return __float_as_uint(v); }
#else
XQ_HD uint32_t float_bits(float v) { uint32_t u; memcpy(&u, &v, 4); return u; }
#endif
// a > b as floats  <=>  ordered_key(a) > ordered_key(b) as unsigned integers (finite values; -0 < +0, which never decides a strict >
// between the two because tanh outputs of distinct sums are compared by value: see the test) ; 0 is below every float
XQ_HD uint32_t ordered_key(float v) {
    const uint32_t u = float_bits(v) + 0u;
    const uint32_t z = (u << 1) == 0u ? 0u : u;                 // -0.0 -> +0.0: equal as floats, must be equal as keys
    return (z & 0x80000000u) ? ~z : (z | 0x80000000u);
}

// phase A2 (after team_phase_a): first maximum of Q[to] over the moves of each of my pieces, in generator order
template <int KB>
XQ_HD void act_best(const TeamRole& R, const TeamState& st, const TeamPly& pl, ActShared<KB>& as, int lane, uint32_t (&kbest)[4]) {
    const float* qt = as.qt + lane;
    const uint32_t flip = st.player ? 0xFFFFFFFFu : 0u;
    // position 0: a slider; ray k holds e empty squares, then (capdist != 0) the capture square (slider_desc, xq_bitboard.cuh)
    {
        const int sq = (int)(st.sq_own & 0xFFu);
        const uint32_t desc = (pl.cntw & 0xFFu) ? pl.desc[0] : 0u;      // a captured slider's descriptor is garbage
        float best = 0.f;
        uint32_t key = 0, idx = 0, kb = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int e = (int)((desc >> (8 * k)) & 15u), capdist = (int)((desc >> (8 * k + 4)) & 15u);
            const int step = k == 0 ? 1 : (k == 1 ? -1 : (k == 2 ? 9 : -9));
            const int cnt = e + (capdist ? 1 : 0);
            for (int j = 0; j < cnt; ++j) {
                const int to = sq + step * (j < e ? j + 1 : capdist);
                const float v = qt[to * (KB + 1)];
                if (key == 0 || v > best) { best = v; key = 1; kb = idx; }
                ++idx;
            }
        }
        const bool any = (pl.cntw & 0xFFu) != 0;
        as.key[(R.role * 4 + 0) * KB + lane] = any ? ordered_key(best) : 0u;
        kbest[0] = kb;
    }
#pragma unroll
    for (int i = 1; i < 4; ++i) {
        const int sq = (int)((st.sq_own >> (8 * i)) & 0xFFu);
        const uint32_t cnt = (pl.cntw >> (8 * i)) & 0xFFu;
        uint32_t mask = cnt ? (pl.desc[i] & 0xFFu) : 0u;
        const uint64_t tab = ((uint64_t)(i == 1 ? R.tab_hi : 0u) << 32) | (R.tab_lo[i] ^ (R.sold[i] & flip));
        float best = 0.f;
        uint32_t key = 0, idx = 0, kb = 0;
        while (mask) {
            const int k = ffs32(mask) - 1;
            mask &= mask - 1;
            const int to = sq + (int)(int8_t)(uint8_t)(tab >> (8 * k));
            const float v = qt[to * (KB + 1)];
            if (key == 0 || v > best) { best = v; key = 1; kb = idx; }
            ++idx;
        }
        as.key[(R.role * 4 + i) * KB + lane] = key ? ordered_key(best) : 0u;
        kbest[i] = kb;
    }
}

// API mode (legal_moves_team_kernel): ChessAI::getAllValidActions(side to move) as the ORDERED LIST (src/chessai.cpp:347-368), after team_phase_a.
// Every thread finds the place of each of its 4 pieces in the reference order (byte-SIMD prefix over the published (squares, counts) words, as
// in team_phase_b) and emits the piece's actions in generator order: emit(index in the list, action).  Returns the list size.
template <int KB, class EMIT>
XQ_HD uint32_t team_emit_actions(const TeamRole& R, const TeamState& st, const TeamPly& pl, const TeamShared<KB>& sh, int lane, EMIT&& emit) {
    uint32_t qw[4], cw[4], tot = 0;
#pragma unroll
    for (int w = 0; w < 4; ++w) { qw[w] = sh.q[w * KB + lane]; cw[w] = sh.c[w * KB + lane]; tot = dp4a_u(cw[w], 0x01010101u, tot); }
    const uint32_t flip = st.player ? 0xFFFFFFFFu : 0u;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int sq = (int)((st.sq_own >> (8 * i)) & 0xFFu);
        const uint32_t cnt = (pl.cntw >> (8 * i)) & 0xFFu;             // 0 for a captured piece (its descriptor is garbage)
        const uint32_t base = (uint32_t)sq * 0x01010101u + 0x7F7F7F7Fu;
        uint32_t acc = 0;
#pragma unroll
        for (int w = 0; w < 4; ++w) acc = dp4a_u(cw[w], (base - qw[w]) & 0x80808080u, acc);
        int idx = (int)(acc >> 7);
        if (i == 0) {      // the slider: per ray the empty squares, then the capture (slider_desc, xq_bitboard.cuh)
            const uint32_t desc = cnt ? pl.desc[0] : 0u;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int e = (int)((desc >> (8 * k)) & 15u), capdist = (int)((desc >> (8 * k + 4)) & 15u);
                const int step = k == 0 ? 1 : (k == 1 ? -1 : (k == 2 ? 9 : -9));
                const int c = e + (capdist ? 1 : 0);
                for (int j = 0; j < c; ++j) emit(idx++, (int)XQ_ACTION(sq, sq + step * (j < e ? j + 1 : capdist)));
            }
        } else {           // a leaper: the playable directions in the order of generate*Moves
            uint32_t mask = cnt ? (pl.desc[i] & 0xFFu) : 0u;
            const uint64_t tab = ((uint64_t)(i == 1 ? R.tab_hi : 0u) << 32) | (R.tab_lo[i] ^ (R.sold[i] & flip));
            while (mask) {
                const int k = ffs32(mask) - 1;
                mask &= mask - 1;
                emit(idx++, (int)XQ_ACTION(sq, sq + (int)(int8_t)(uint8_t)(tab >> (8 * k))));
            }
        }
    }
    return tot;
}

// phase B: the chosen action; the owning thread decodes it and publishes from | to << 8 in sh.move.  Returns the list size
// (0: no legal action, nothing published).  x = xq_rng(seed, env id, ctr) of this ply.
template <int KB>
XQ_HD uint32_t act_select(const TeamRole& R, const TeamState& st, const TeamPly& pl, TeamShared<KB>& sh, const ActShared<KB>& as, int lane,
                          const uint32_t (&kbest)[4], uint64_t x, uint32_t eps_thr) {
    uint32_t qw[4], cw[4], tot = 0;
#pragma unroll
    for (int w = 0; w < 4; ++w) { qw[w] = sh.q[w * KB + lane]; cw[w] = sh.c[w * KB + lane]; tot = dp4a_u(cw[w], 0x01010101u, tot); }
    if (tot == 0) return 0;
    const uint32_t coin31 = (uint32_t)(x & 0x7FFFFFFFu), idx31 = (uint32_t)(x >> 33);
    const bool explore = coin31 < eps_thr;                          // rand()/RAND_MAX < epsilon (src/dqn.cpp:30-34)
    const uint32_t k = team_mod(idx31, tot, sh.magic[tot]);
    // exploit: the largest key, then the lowest square among its holders (keys of pieces without a move are 0 < every real key)
    uint32_t kmax = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) { const uint32_t kj = as.key[j * KB + lane]; kmax = kj > kmax ? kj : kmax; }
    uint32_t wsq = 127;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const uint32_t kj = as.key[j * KB + lane], sj = (qw[j >> 2] >> (8 * (j & 3))) & 0xFFu;
        wsq = (kj == kmax && sj < wsq) ? sj : wsq;
    }
    uint32_t hsq = 0, hdesc = 0, hwant = 0, hlo = 0, hhi = 0;
    bool hit = false, hslider = false;
    const uint32_t flip = st.player ? 0xFFFFFFFFu : 0u;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t sq = (st.sq_own >> (8 * i)) & 0xFFu;
        const uint32_t base = sq * 0x01010101u + 0x7F7F7F7Fu;      // reference-order prefix, as in team_phase_b
        uint32_t acc = 0;
#pragma unroll
        for (int w = 0; w < 4; ++w) acc = dp4a_u(cw[w], (base - qw[w]) & 0x80808080u, acc);
        const uint32_t want_x = k - (acc >> 7), cnt = (pl.cntw >> (8 * i)) & 0xFFu;
        const bool h = explore ? want_x < cnt : (cnt != 0 && sq == wsq && as.key[(R.role * 4 + i) * KB + lane] == kmax);
        const uint32_t want = explore ? want_x : kbest[i];
        hit |= h;
        hsq = h ? sq : hsq; hdesc = h ? pl.desc[i] : hdesc; hwant = h ? want : hwant;
        hlo = h ? (R.tab_lo[i] ^ (R.sold[i] & flip)) : hlo;
        if (i == 0) hslider = h;
        if (i == 1) hhi = h ? R.tab_hi : 0u;
    }
    if (hit) {
        const int to_s = slider_decode(hdesc, (int)hsq, (int)hwant);
        const int dk = nth_set_bit8(hdesc & 0xFFu, (int)hwant & 7);
        const int to_l = (int)hsq + (int)(int8_t)(uint8_t)((((uint64_t)hhi << 32) | hlo) >> (8 * dk));
        sh.move[lane] = hsq | ((uint32_t)(hslider ? to_s : to_l) << 8);
    }
    return tot;
}

}  // namespace xq
