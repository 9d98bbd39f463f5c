// TEST-ONLY host build of the device rules header (cn_chess_ai_b200/csrc/xq_rules.cuh).
// It lets the CPU-only test suite diff the exact source the CUDA kernels compile against
// the oracle on millions of positions before any GPU time is spent.  This library lives
// under tests/ and is never loaded by the product package: the product has no CPU path.
#include <cstdint>
#include <cstring>
#include "../../cn_chess_ai_b200/csrc/xq_rules.cuh"
#include "../../cn_chess_ai_b200/csrc/xq_bitboard.cuh"
#include "../../include/xq.h"

namespace {
struct RecBoard {
    const uint32_t* w;
    int get(int s) const { return (w[s >> 3] >> ((s & 7) * 4)) & 15; }
};
}

extern "C" {
void hs_all_actions(const xq_env_rec* recs, long n, uint8_t* counts, uint16_t* actions) {
    for (long i = 0; i < n; ++i) {
        RecBoard b{recs[i].sq};
        uint16_t* out = actions + i * XQ_MAX_ACTIONS;
        for (int k = 0; k < XQ_MAX_ACTIONS; ++k) out[k] = XQ_ACTION_NONE;
        int cnt = 0;
        xq::all_actions(b, recs[i].player, [&](int from, int to) { if (cnt < XQ_MAX_ACTIONS) out[cnt++] = XQ_ACTION(from, to); });
        counts[i] = (uint8_t)cnt;
    }
}
int hs_valid_moves(const xq_env_rec* rec, int row, int col, uint8_t* to) {
    RecBoard b{rec->sq};
    int n = 0;
    if (!xq::inside(row, col)) return 0;
    int code = b.get(row * 9 + col);
    if (code) xq::gen_piece(b, row, col, code, [&](int t) { to[n++] = (uint8_t)t; });
    return n;
}
int hs_is_valid_move(const xq_env_rec* rec, int fr, int fc, int tr, int tc) {
    RecBoard b{rec->sq};
    return xq::is_valid_move(b, fr, fc, tr, tc) ? 1 : 0;
}
// bitboard count + k-th decode (xq_bitboard.cuh), enumerated piece by piece in square order
void hs_bb_all_actions(const xq_env_rec* recs, long n, uint8_t* counts, uint16_t* actions) {
    for (long i = 0; i < n; ++i) {
        RecBoard b{recs[i].sq};
        const int player = recs[i].player;
        xq::Pos P{{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        for (int s = 0; s < 90; ++s) {
            const int code = b.get(s);
            if (!code) continue;
            P.occ.set(s); P.occT.set(xq::cm_index(s / 9, s % 9));
            if ((code >= 8) == (player == 1)) P.own.set(s);
        }
        uint16_t* out = actions + i * XQ_MAX_ACTIONS;
        for (int k = 0; k < XQ_MAX_ACTIONS; ++k) out[k] = XQ_ACTION_NONE;
        int cnt = 0;
        for (int s = 0; s < 90; ++s) {
            const int code = b.get(s);
            if (!code || (code >= 8) != (player == 1)) continue;
            int dummy = -1;
            const int c = xq::piece_moves_dyn(xq::type_of(code), P, s, player, -1, &dummy);
            for (int j = 0; j < c; ++j) {
                int to = -1;
                xq::piece_moves_dyn(xq::type_of(code), P, s, player, j, &to);
                if (cnt < XQ_MAX_ACTIONS) out[cnt++] = XQ_ACTION(s, to);
            }
        }
        counts[i] = (uint8_t)cnt;
    }
}
int hs_reward(int material_diff, int move_count) { return xq::reward_from_material(material_diff, move_count); }
int hs_piece_score(int type) { return xq::piece_score(type); }
uint64_t hs_rng(uint64_t seed, uint64_t env, uint32_t ctr) { return xq::rng(seed, env, ctr); }
}
