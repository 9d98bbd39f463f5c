// xq_rules.cuh -- Xiangqi rules of the reference ChessBoard as branch-light device functions.
//
// Everything is templated on a board accessor B with `int get(int sq) const` (4-bit piece code)
// so that the same source runs (a) in the CUDA kernels on a bank-conflict-free shared-memory
// board and (b), compiled by g++ with XQ_HOSTSIM, inside the CPU-only unit tests that diff it
// against the oracle (tests/hostsim; never part of the product library).
//
// The reference generators call isValidMove on every candidate (src/chessboard.cpp:149-283 ->
// :66-93 -> :328-440); the functions below implement the CONJUNCTION of generator test and
// predicate directly, without re-scanning slider paths, and keep the reference's emission
// order.  Rules are the reference's "capture the general" pseudo-legal
// rules: no turn test, no king-safety test, no flying-general rule.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define XQ_HD __host__ __device__ __forceinline__
#else
#define XQ_HD inline
#endif

namespace xq {

enum : int { EMPTY = 0, GENERAL = 1, ADVISOR = 2, ELEPHANT = 3, HORSE = 4, CHARIOT = 5, CANNON = 6, SOLDIER = 7 };
enum : int { RED = 0, BLACK = 1, NOCOLOR = 2 };

XQ_HD int type_of(int code) { return code >= 8 ? code - 7 : code; }
XQ_HD int color_of(int code) { return code == 0 ? NOCOLOR : (code >= 8 ? BLACK : RED); }
// (bitwise &, | on purpose: short-circuit operators became branch regions in the generators)
XQ_HD constexpr bool inside(int r, int c) { return ((unsigned)r < 10u) & ((unsigned)c < 9u); }  // :323-325
XQ_HD constexpr bool in_palace_of(int color, int r, int c) {                                  // chessboard.h:65-71
    return ((unsigned)(c - 3) < 3u) & ((unsigned)(r - (color == RED ? 0 : 7)) < 3u);
}
XQ_HD constexpr bool in_any_palace(int r, int c) { return ((unsigned)(c - 3) < 3u) & (((unsigned)r < 3u) | ((unsigned)(r - 7) < 3u)); }

// getPieceScore / PieceScore, src/chessboard.cpp:443-454, include/chessboard.h:23-31.
// Packed table: scores are multiples of 5 -> score/5 fits a byte (200,4,4,8,18,9,2).
XQ_HD int piece_score(int type) {
    const uint64_t tbl = 0x020912080404C800ull;  // byte t = score(t)/5
    return (int)((tbl >> (type * 8)) & 0xFF) * 5;
}

// own-or-empty test used by every generator through isValidMove (:78-80)
XQ_HD bool not_own(int code_to, int color) { return code_to == 0 || (code_to >= 8) != (color == BLACK); }

// ---- ChessBoard::isValidMove as a stand-alone predicate (:66-93, :328-440) -------------------
template <class B>
XQ_HD int count_between(const B& b, int fr, int fc, int tr, int tc) {
    const bool same_row = fr == tr;
    const int step = same_row ? (tc > fc ? 1 : -1) : (tr > fr ? 1 : -1);
    const int start = same_row ? fc : fr, end = same_row ? tc : tr;
    int n = 0;
    for (int i = start + step; i != end; i += step) n += b.get(same_row ? fr * 9 + i : i * 9 + fc) != 0;
    return n;
}

template <class B>
XQ_HD bool is_valid_move(const B& b, int fr, int fc, int tr, int tc) {
    if (!inside(fr, fc) || !inside(tr, tc)) return false;
    const int from = b.get(fr * 9 + fc), to = b.get(tr * 9 + tc);
    if (from == 0) return false;
    const int color = color_of(from);
    if (!not_own(to, color)) return false;
    const int rd = tr - fr, cd = tc - fc;
    const int ard = rd < 0 ? -rd : rd, acd = cd < 0 ? -cd : cd;
    switch (type_of(from)) {
        case GENERAL: return in_any_palace(fr, fc) && in_any_palace(tr, tc) && ard + acd == 1;          // :328-343
        case ADVISOR: return in_any_palace(tr, tc) && ard == 1 && acd == 1;                             // :346-353
        case ELEPHANT:                                                                                  // :355-367
            return ard == 2 && acd == 2 && ((fr < 5) == (tr < 5)) && b.get(((fr + tr) / 2) * 9 + (fc + tc) / 2) == 0;
        case HORSE:                                                                                     // :369-380
            if ((ard == 2 && acd == 1) || (ard == 1 && acd == 2)) return b.get((fr + rd / 2) * 9 + fc + cd / 2) == 0;
            return false;
        case CHARIOT:                                                                                   // :382-397
            if (fr != tr && fc != tc) return false;
            return count_between(b, fr, fc, tr, tc) == 0;
        case CANNON: {                                                                                  // :399-421
            if (fr != tr && fc != tc) return false;
            const int n = count_between(b, fr, fc, tr, tc);
            return to == 0 ? n == 0 : n == 1;
        }
        case SOLDIER:                                                                                   // :423-440
            if (color == RED) return (rd == 1 && acd == 0) || (fr >= 5 && rd == 0 && acd == 1);
            return (rd == -1 && acd == 0) || (fr < 5 && rd == 0 && acd == 1);
        default: return false;
    }
}

// ---- ChessBoard::getValidMoves (:112-147) for the piece `code` standing on (r,c) -------------
// emit(to_sq) is called once per destination, in the reference's order.
template <class B, class F>
XQ_HD void gen_piece(const B& b, int r, int c, int code, F&& emit) {
    const int color = color_of(code);
    const int sq = r * 9 + c;
    switch (type_of(code)) {
        case GENERAL: {  // :149-160, dirs (1,0),(-1,0),(0,1),(0,-1); from and to inside a palace (:330-337)
            if (!in_any_palace(r, c)) break;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int dr = k == 0 ? 1 : (k == 1 ? -1 : 0), dc = k == 2 ? 1 : (k == 3 ? -1 : 0);
                const int nr = r + dr, nc = c + dc;
                if (inside(nr, nc) && in_any_palace(nr, nc) && not_own(b.get(nr * 9 + nc), color)) emit(nr * 9 + nc);
            }
            break;
        }
        case ADVISOR: {  // :162-177, dirs (1,1),(1,-1),(-1,1),(-1,-1); destination in the OWN palace (:170-172)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int nr = r + (k < 2 ? 1 : -1), nc = c + ((k & 1) ? -1 : 1);
                if (in_palace_of(color, nr, nc) && not_own(b.get(nr * 9 + nc), color)) emit(nr * 9 + nc);
            }
            break;
        }
        case ELEPHANT: {  // :179-196, dirs (2,2),(2,-2),(-2,2),(-2,-2); own side (:190) and no river crossing (:359)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int dr = k < 2 ? 2 : -2, dc = (k & 1) ? -2 : 2;
                const int nr = r + dr, nc = c + dc;
                if (!inside(nr, nc)) continue;
                const bool side_ok = color == RED ? (nr <= 4 && r < 5) : (nr >= 5 && r >= 5);
                if (side_ok && b.get((r + dr / 2) * 9 + c + dc / 2) == 0 && not_own(b.get(nr * 9 + nc), color)) emit(nr * 9 + nc);
            }
            break;
        }
        case HORSE: {  // :248-263, dirs (1,2),(1,-2),(-1,2),(-1,-2),(2,1),(2,-1),(-2,1),(-2,-1); leg = dir/2 truncated
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int a = (k & 2) ? -1 : 1, bb = (k & 1) ? -1 : 1;
                const int dr = k < 4 ? a : 2 * a, dc = k < 4 ? 2 * bb : bb;
                const int nr = r + dr, nc = c + dc;
                if (!inside(nr, nc)) continue;
                const int leg = k < 4 ? sq + bb : sq + 9 * a;
                if (b.get(leg) == 0 && not_own(b.get(nr * 9 + nc), color)) emit(nr * 9 + nc);
            }
            break;
        }
        case CHARIOT: {  // :198-218, dirs (0,1),(0,-1),(1,0),(-1,0)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int dr = k == 2 ? 1 : (k == 3 ? -1 : 0), dc = k == 0 ? 1 : (k == 1 ? -1 : 0);
                int nr = r + dr, nc = c + dc;
                while (inside(nr, nc)) {
                    const int t = b.get(nr * 9 + nc);
                    if (t == 0) { emit(nr * 9 + nc); }
                    else { if (not_own(t, color)) emit(nr * 9 + nc); break; }
                    nr += dr; nc += dc;
                }
            }
            break;
        }
        case CANNON: {  // :220-246: empties up to the screen, then the first piece behind it iff it is an enemy
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int dr = k == 2 ? 1 : (k == 3 ? -1 : 0), dc = k == 0 ? 1 : (k == 1 ? -1 : 0);
                int nr = r + dr, nc = c + dc;
                while (inside(nr, nc) && b.get(nr * 9 + nc) == 0) { emit(nr * 9 + nc); nr += dr; nc += dc; }
                nr += dr; nc += dc;  // step over the screen
                while (inside(nr, nc)) {
                    const int t = b.get(nr * 9 + nc);
                    if (t != 0) { if (not_own(t, color)) emit(nr * 9 + nc); break; }
                    nr += dr; nc += dc;
                }
            }
            break;
        }
        case SOLDIER: {  // :265-283: forward, then (col-1),(col+1) once across the river
            const int nr = r + (color == RED ? 1 : -1);
            if (inside(nr, c) && not_own(b.get(nr * 9 + c), color)) emit(nr * 9 + c);
            if (color == RED ? r > 4 : r < 5) {
                if (c > 0 && not_own(b.get(sq - 1), color)) emit(sq - 1);
                if (c < 8 && not_own(b.get(sq + 1), color)) emit(sq + 1);
            }
            break;
        }
        default: break;
    }
}

// ---- ChessAI::getAllValidActions (src/chessai.cpp:347-368): row-major scan, colour == player ----
// emit(from_sq, to_sq)
template <class B, class F>
XQ_HD void all_actions(const B& b, int player, F&& emit) {
    for (int r = 0; r < 10; ++r)
        for (int c = 0; c < 9; ++c) {
            const int code = b.get(r * 9 + c);
            if (code != 0 && (code >= 8) == (player == BLACK)) {
                const int from = r * 9 + c;
                gen_piece(b, r, c, code, [&](int to) { emit(from, to); });
            }
        }
}

// ---- opt-in STRICT legality (xq_env_legal_moves_strict) -- NOT the reference's rules --------------------------------------------
// The reference plays "capture the General": it has no king-safety test and no flying-general rule, and every
// parity path of this library follows it.  Standard Xiangqi additionally forbids a move after which the mover's own General could
// be taken.  This predicate states that on top of the reference's own generators: the pseudo-legal action (from, to) of `player`
// is kept iff, on the board after it, (a) no enemy piece has `g` among the destinations ChessBoard::getValidMoves generates for it,
// g = the mover's first General in square order ("self-check"), and (b) the two first Generals do not stand on one file with
// nothing between them ("flying general").  A side without a General has nothing to protect: all its actions are kept.
// The board is modified and restored.
template <class B>
XQ_HD bool leaves_general_safe(B& b, int player, int from, int to) {
    const int code = b.get(from), cap = b.get(to);
    b.set(to, code);
    b.set(from, 0);
    const int own_gen = player == RED ? GENERAL : GENERAL + 7, opp_gen = player == RED ? GENERAL + 7 : GENERAL;
    int g = -1, eg = -1;
    for (int s = 0; s < 90; ++s) {
        const int c = b.get(s);
        if (c == own_gen && g < 0) g = s;
        if (c == opp_gen && eg < 0) eg = s;
    }
    bool safe = true;
    if (g >= 0) {
        for (int s = 0; s < 90 && safe; ++s) {
            const int c = b.get(s);
            if (c != 0 && (c >= 8) != (player == BLACK)) gen_piece(b, s / 9, s % 9, c, [&](int t) { if (t == g) safe = false; });
        }
        if (safe && eg >= 0 && g % 9 == eg % 9) safe = count_between(b, g / 9, g % 9, eg / 9, eg % 9) != 0;
    }
    b.set(from, code);
    b.set(to, cap);
    return safe;
}

// ChessAI::evaluateBoard's last two lines (src/chessai.cpp:343-344): (int)(score - moveCount*0.1) with
// IEEE double mul-then-sub equals this truncating integer division for every reachable
// (score, moveCount) (re-proved by tests/test_oracle.py).  No floating point on device.
XQ_HD int reward_from_material(int material_diff, int move_count) { return (10 * material_diff - move_count) / 10; }

// Counter RNG (include/xq.h: xq_rng)
XQ_HD uint64_t rng(uint64_t seed, uint64_t env_id, uint32_t ctr) {
    uint64_t z = seed + env_id * 0x9E3779B97F4A7C15ull + (uint64_t)ctr * 0xD1B54A32D192ED03ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

}  // namespace xq
