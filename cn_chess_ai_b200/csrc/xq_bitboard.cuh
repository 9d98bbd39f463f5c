// xq_bitboard.cuh -- per-piece move COUNT (+ a 32-bit descriptor) and k-th-move DECODE (from the descriptor) on 90-bit occupancy bitboards.
//
// Used by the slot-parallel rollout kernel (xq_rollout.cu): one thread owns one piece, warps are
// piece-type uniform, so every function here runs without divergence across a warp.  Sliders are
// O(1): the rank / file occupancy is pulled out of the row-major / column-major bitboards with a
// funnel shift and the first and second blocker on each ray come from ffs/clz -- no per-square
// loops.  Emission order is the reference's: direction order as written in
// generate*Moves (src/chessboard.cpp:149-283), distance ascending along a ray.
//
// Host-compilable (tests/hostsim) so the CPU suite can diff count+decode against the oracle.
#pragma once
#include <stdint.h>

#include "xq_rules.cuh"

namespace xq {

#if defined(__CUDA_ARCH__)
XQ_HD int ffs32(uint32_t x) { return __ffs((int)x); }          // 1-based, 0 if x == 0
XQ_HD int clz32(uint32_t x) { return __clz((int)x); }
XQ_HD int popc32(uint32_t x) { return __popc(x); }
XQ_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, int sh) { return __funnelshift_r(lo, hi, sh); }
#else
XQ_HD int ffs32(uint32_t x) { return __builtin_ffs((int)x); }
XQ_HD int clz32(uint32_t x) { return x ? __builtin_clz(x) : 32; }
XQ_HD int popc32(uint32_t x) { return __builtin_popcount(x); }
XQ_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, int sh) { return (uint32_t)((((uint64_t)hi << 32) | lo) >> (sh & 31)); }
#endif

// 90-bit set in three words.  Row-major index = row*9+col; column-major index = col*10+row.
struct Bits90 {
    uint32_t w0, w1, w2;
    XQ_HD uint32_t word(int i) const { return i == 0 ? w0 : (i == 1 ? w1 : w2); }
    XQ_HD bool test(int i) const { return (word(i >> 5) >> (i & 31)) & 1u; }
    XQ_HD void set(int i) { const uint32_t b = 1u << (i & 31); const int w = i >> 5; w0 |= w == 0 ? b : 0u; w1 |= w == 1 ? b : 0u; w2 |= w == 2 ? b : 0u; }
    XQ_HD void clear(int i) { const uint32_t b = ~(1u << (i & 31)); const int w = i >> 5; w0 &= w == 0 ? b : ~0u; w1 &= w == 1 ? b : ~0u; w2 &= w == 2 ? b : ~0u; }
    // single-bit mask of index i spread over the three words (computed once, applied with plain logic ops)
    static XQ_HD Bits90 bit(int i) { const uint32_t b = 1u << (i & 31); const int w = i >> 5; return Bits90{w == 0 ? b : 0u, w == 1 ? b : 0u, w == 2 ? b : 0u}; }
    XQ_HD void or_with(const Bits90& m) { w0 |= m.w0; w1 |= m.w1; w2 |= m.w2; }
    XQ_HD void andnot(const Bits90& m) { w0 &= ~m.w0; w1 &= ~m.w1; w2 &= ~m.w2; }
    // nbits (<= 10) starting at bit pos
    XQ_HD uint32_t field(int pos, int nbits) const {
        const int w = pos >> 5;
        const uint32_t lo = word(w), hi = w == 0 ? w1 : (w == 1 ? w2 : 0u);
        return funnel_r(lo, hi, pos & 31) & ((1u << nbits) - 1u);
    }
};
XQ_HD int rm_index(int r, int c) { return r * 9 + c; }
XQ_HD int cm_index(int r, int c) { return c * 10 + r; }
XQ_HD int row_of(int sq) { return (sq * 57) >> 9; }      // sq/9 for 0 <= sq < 128

// Position seen by the side to move: `own`/`occ` row-major, `occT` column-major (all pieces).
struct Pos {
    Bits90 own, occ, occT;
};

// One slider ray on a line occupancy L (nb bits) from index p: number of empty squares before the
// first blocker, index of the first blocker (-1: none) and of the second blocker (-1: none).
struct Ray { int empties, first, second; };
XQ_HD Ray ray_up(uint32_t L, int p, int nb) {      // towards higher index
    const uint32_t m = L >> (p + 1);
    Ray r;
    if (m == 0) { r.empties = nb - 1 - p; r.first = -1; r.second = -1; return r; }
    const int d = ffs32(m);
    r.empties = d - 1; r.first = p + d;
    const uint32_t m2 = m & (m - 1);
    r.second = m2 ? p + ffs32(m2) : -1;
    return r;
}
XQ_HD Ray ray_down(uint32_t L, int p, int) {       // towards lower index
    const uint32_t m = L & ((1u << p) - 1u);
    Ray r;
    if (m == 0) { r.empties = p; r.first = -1; r.second = -1; return r; }
    const int top = 31 - clz32(m);
    r.empties = p - top - 1; r.first = top;
    const uint32_t m2 = m ^ (1u << top);
    r.second = m2 ? 31 - clz32(m2) : -1;
    return r;
}

// Sliders.  IS_CANNON: capture target is the SECOND blocker (src/chessboard.cpp:220-246, :399-421),
// else the first (:198-218, :382-397).  Ray order E, W, S(row+1), N = (0,1),(0,-1),(1,0),(-1,0).
// The generator runs ONCE per ply: it returns the count and a 32-bit DESCRIPTOR -- per ray k a byte
// (number of empty squares before the first blocker) | (distance of the capture square, 0 = none) << 4 --
// from which the k-th destination (reference order) is decoded without touching the board again.
template <bool IS_CANNON>
XQ_HD int slider_desc(const Pos& P, int sq, uint32_t* desc) {
    const int r = row_of(sq), c = sq - 9 * r;
    const uint32_t rank = P.occ.field(9 * r, 9), file = P.occT.field(10 * c, 10);
    int total = 0;
    uint32_t d = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const bool horiz = k < 2;
        const int p = horiz ? c : r;
        const Ray ray = (k & 1) ? ray_down(horiz ? rank : file, p, horiz ? 9 : 10)
                                : ray_up(horiz ? rank : file, p, horiz ? 9 : 10);
        const int tgt = IS_CANNON ? ray.second : ray.first;
        int capdist = 0;
        if (tgt >= 0) { const int s = horiz ? 9 * r + tgt : 9 * tgt + c; if (!P.own.test(s)) capdist = (k & 1) ? p - tgt : tgt - p; }
        d |= ((uint32_t)ray.empties | ((uint32_t)capdist << 4)) << (8 * k);
        total += ray.empties + (capdist ? 1 : 0);
    }
    *desc = d;
    return total;
}
XQ_HD int slider_decode(uint32_t desc, int sq, int want) {
    int to = sq;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int e = (desc >> (8 * k)) & 15, capdist = (desc >> (8 * k + 4)) & 15;
        const int cnt = e + (capdist ? 1 : 0);
        const int step = k == 0 ? 1 : (k == 1 ? -1 : (k == 2 ? 9 : -9));
        if (want >= 0 && want < cnt) to = sq + step * (want < e ? want + 1 : capdist);
        want -= cnt;                      // once negative it stays negative: exactly one ray matches
    }
    return to;
}

// Leapers: bit k of the returned mask = direction k (reference order) is playable; the destinations are recomputed by
// *_dir().  color: RED/BLACK of the piece.
// Every square a leaper looks at lies within +-20 of its own square, so instead of a dynamic 3-word bit test per
// square (two selects + shift + and, 16 of them for a horse) the 41 bits around the piece are pulled out of a
// bitboard ONCE with two funnel shifts; after that every test is a compile-time bit position.  Squares beyond the
// board read as empty -- the row / column bounds are tested separately, as the reference does (isInsideBoard).
struct Win41 {
    uint32_t lo, hi;                     // bit (off + 20) = board bit (sq + off), off in [-20, 20]
    XQ_HD uint32_t at(int off) const { const int p = off + 20; return (p < 32 ? lo >> p : hi >> (p - 32)) & 1u; }
};
XQ_HD Win41 window(const Bits90& b, int sq) {
    const int q = sq + 12;               // (sq - 20) + 32: bit index into the padded word array {0, w0, w1, w2, 0, 0}
    const int wi = q >> 5, sh = q & 31;
    const uint32_t x0 = wi == 0 ? 0u : (wi == 1 ? b.w0 : (wi == 2 ? b.w1 : b.w2));
    const uint32_t x1 = wi == 0 ? b.w0 : (wi == 1 ? b.w1 : (wi == 2 ? b.w2 : 0u));
    const uint32_t x2 = wi == 0 ? b.w1 : (wi == 1 ? b.w2 : 0u);
    return Win41{funnel_r(x0, x1, sh), funnel_r(x1, x2, sh)};
}
XQ_HD int general_dir(int k) { return k == 0 ? 9 : (k == 1 ? -9 : (k == 2 ? 1 : -1)); }           // :150
XQ_HD uint32_t general_mask(const Pos& P, int sq) {                                               // :149-160, :328-343
    const int r = row_of(sq), c = sq - 9 * r;
    if (!in_any_palace(r, c)) return 0;
    const Win41 own = window(P.own, sq);
    uint32_t m = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int nr = r + (k == 0 ? 1 : (k == 1 ? -1 : 0)), nc = c + (k == 2 ? 1 : (k == 3 ? -1 : 0));
        if (in_any_palace(nr, nc) && !own.at(k == 0 ? 9 : (k == 1 ? -9 : (k == 2 ? 1 : -1)))) m |= 1u << k;
    }
    return m;
}
XQ_HD int advisor_dir(int k) { return k == 0 ? 10 : (k == 1 ? 8 : (k == 2 ? -8 : -10)); }         // :163
XQ_HD uint32_t advisor_mask(const Pos& P, int sq, int color) {                                    // :162-177
    const int r = row_of(sq), c = sq - 9 * r;
    const Win41 own = window(P.own, sq);
    uint32_t m = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int nr = r + (k < 2 ? 1 : -1), nc = c + ((k & 1) ? -1 : 1);
        if (in_palace_of(color, nr, nc) && !own.at(k == 0 ? 10 : (k == 1 ? 8 : (k == 2 ? -8 : -10)))) m |= 1u << k;
    }
    return m;
}
XQ_HD int elephant_dir(int k) { return k == 0 ? 20 : (k == 1 ? 16 : (k == 2 ? -16 : -20)); }      // :180
XQ_HD uint32_t elephant_mask(const Pos& P, int sq, int color) {                                   // :179-196, :355-367
    const int r = row_of(sq), c = sq - 9 * r;
    const Win41 own = window(P.own, sq), occ = window(P.occ, sq);
    uint32_t m = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int dr = k < 2 ? 2 : -2, dc = (k & 1) ? -2 : 2;
        const int nr = r + dr, nc = c + dc;
        const bool side_ok = color == RED ? (nr <= 4 && r < 5) : (nr >= 5 && r >= 5);
        const int d = k == 0 ? 20 : (k == 1 ? 16 : (k == 2 ? -16 : -20));
        if (inside(nr, nc) && side_ok && !occ.at(d / 2) && !own.at(d)) m |= 1u << k;
    }
    return m;
}
XQ_HD int horse_dir(int k) {                                                                       // :249
    const int a = (k & 2) ? -1 : 1, b = (k & 1) ? -1 : 1;
    return k < 4 ? 9 * a + 2 * b : 18 * a + b;
}
XQ_HD uint32_t horse_mask(const Pos& P, int sq) {                                                 // :248-263, :369-380
    const int r = row_of(sq), c = sq - 9 * r;
    const Win41 own = window(P.own, sq), occ = window(P.occ, sq);
    uint32_t m = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int a = (k & 2) ? -1 : 1, b = (k & 1) ? -1 : 1;
        const int nr = r + (k < 4 ? a : 2 * a), nc = c + (k < 4 ? 2 * b : b);
        const int leg = k < 4 ? b : 9 * a, dest = k < 4 ? 9 * a + 2 * b : 18 * a + b;       // offsets from sq: compile-time after unrolling
        if (inside(nr, nc) && !occ.at(leg) && !own.at(dest)) m |= 1u << k;
    }
    return m;
}
XQ_HD int soldier_dir(int k, int color) { return k == 0 ? (color == RED ? 9 : -9) : (k == 1 ? -1 : 1); }   // :267-281
XQ_HD uint32_t soldier_mask(const Pos& P, int sq, int color) {                                    // :265-283
    const int r = row_of(sq), c = sq - 9 * r;
    const Win41 own = window(P.own, sq);
    uint32_t m = 0;
    const int nr = r + (color == RED ? 1 : -1);
    if ((unsigned)nr < 10u && !(color == RED ? own.at(9) : own.at(-9))) m |= 1u;
    if (color == RED ? r > 4 : r < 5) {
        if (c > 0 && !own.at(-1)) m |= 2u;
        if (c < 8 && !own.at(1)) m |= 4u;
    }
    return m;
}
// index of the j-th (0-based) set bit of an 8-bit mask
XQ_HD int nth_set_bit(uint32_t m, int j) {
#pragma unroll
    for (int i = 0; i < 7; ++i) { if (j > 0) { m &= m - 1; --j; } }
    return ffs32(m) - 1;
}

// Uniform entry points.  piece_count<TYPE>: number of moves of the piece of `type` / `color` on `sq` plus the descriptor
// (sliders: see slider_desc; leapers: the direction mask).  piece_decode<TYPE>: the want-th destination in reference
// order from the descriptor alone.
template <int TYPE>
XQ_HD int piece_count(const Pos& P, int sq, int color, uint32_t* desc) {
    if (TYPE == CHARIOT) return slider_desc<false>(P, sq, desc);
    if (TYPE == CANNON) return slider_desc<true>(P, sq, desc);
    uint32_t m;
    if (TYPE == GENERAL) m = general_mask(P, sq);
    else if (TYPE == ADVISOR) m = advisor_mask(P, sq, color);
    else if (TYPE == ELEPHANT) m = elephant_mask(P, sq, color);
    else if (TYPE == HORSE) m = horse_mask(P, sq);
    else m = soldier_mask(P, sq, color);
    *desc = m;
    return popc32(m);
}
template <int TYPE>
XQ_HD int piece_decode(uint32_t desc, int sq, int color, int want) {
    if (TYPE == CHARIOT || TYPE == CANNON) return slider_decode(desc, sq, want);
    const int k = nth_set_bit(desc, want);
    int d;
    if (TYPE == GENERAL) d = general_dir(k);
    else if (TYPE == ADVISOR) d = advisor_dir(k);
    else if (TYPE == ELEPHANT) d = elephant_dir(k);
    else if (TYPE == HORSE) d = horse_dir(k);
    else d = soldier_dir(k, color);
    return sq + d;
}

XQ_HD int piece_count_dyn(int type, const Pos& P, int sq, int color, uint32_t* desc) {
    switch (type) {
        case GENERAL: return piece_count<GENERAL>(P, sq, color, desc);
        case ADVISOR: return piece_count<ADVISOR>(P, sq, color, desc);
        case ELEPHANT: return piece_count<ELEPHANT>(P, sq, color, desc);
        case HORSE: return piece_count<HORSE>(P, sq, color, desc);
        case CHARIOT: return piece_count<CHARIOT>(P, sq, color, desc);
        case CANNON: return piece_count<CANNON>(P, sq, color, desc);
        case SOLDIER: return piece_count<SOLDIER>(P, sq, color, desc);
        default: *desc = 0; return 0;
    }
}
XQ_HD int piece_decode_dyn(int type, uint32_t desc, int sq, int color, int want) {
    switch (type) {
        case GENERAL: return piece_decode<GENERAL>(desc, sq, color, want);
        case ADVISOR: return piece_decode<ADVISOR>(desc, sq, color, want);
        case ELEPHANT: return piece_decode<ELEPHANT>(desc, sq, color, want);
        case HORSE: return piece_decode<HORSE>(desc, sq, color, want);
        case CHARIOT: case CANNON: return slider_decode(desc, sq, want);
        case SOLDIER: return piece_decode<SOLDIER>(desc, sq, color, want);
        default: return sq;
    }
}
// count, and (want >= 0) the want-th destination: the two steps above in one call (host differential tests)
XQ_HD int piece_moves_dyn(int type, const Pos& P, int sq, int color, int want, int* to) {
    uint32_t desc;
    const int n = piece_count_dyn(type, P, sq, color, &desc);
    if (want >= 0) *to = piece_decode_dyn(type, desc, sq, color, want);
    return n;
}

// Piece slots: every side has 16 fixed slots (no promotion in Xiangqi); slot -> type is static,
// which is what makes warps type-uniform.  0,1 Chariot | 2,3 Horse | 4,5 Elephant | 6,7 Advisor |
// 8 General | 9,10 Cannon | 11..15 Soldier.
XQ_HD int slot_type(int s) {
    return s < 2 ? CHARIOT : (s < 4 ? HORSE : (s < 6 ? ELEPHANT : (s < 8 ? ADVISOR : (s == 8 ? GENERAL : (s < 11 ? CANNON : SOLDIER)))));
}
XQ_HD int slot_base(int type) { return (int)((0xB9024680u >> (4 * type)) & 15u); }   // first slot of a type
XQ_HD int slot_cap(int type) { return (int)((0x52222210u >> (4 * type)) & 15u); }    // pieces of a type per side
constexpr int kDeadSq = 127;

}  // namespace xq
