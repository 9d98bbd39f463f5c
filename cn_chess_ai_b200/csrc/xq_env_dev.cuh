// xq_env_dev.cuh -- device helpers shared by the thread-per-board kernels (xq_env.cu, xq_selfplay.cu).
#pragma once
#include "xq_common.cuh"

namespace xq {

constexpr int kThreads = 128;          // thread-per-board kernels: boards per CTA
constexpr int kListStride = 65;        // words per thread of the staged action list (64 + 1 pad)

// ------------------------------------------------------------------------------------------
// Board summary: material per colour, general presence, first general in index order
// (ChessBoard::checkGameOver :286-309, getWinner :312-320, ChessAI::evaluateBoard :311-342).
struct Summary {
    int mat_red, mat_black;
    int winner;          // colour of the first General in square order, NOCOLOR if none
    bool red_alive, black_alive;
};

template <class B>
__device__ __forceinline__ Summary summarize(const B& b) {
    Summary s{0, 0, NOCOLOR, false, false};
    for (int w = 0; w < 12; ++w) {
        uint32_t word = b.base[w * b.stride];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int code = (word >> (4 * i)) & 15;
            if (code == 0) continue;
            const int sc = piece_score(type_of(code));
            if (code >= 8) s.mat_black += sc; else s.mat_red += sc;
            if (code == GENERAL) { s.red_alive = true; if (s.winner == NOCOLOR) s.winner = RED; }
            if (code == GENERAL + 7) { s.black_alive = true; if (s.winner == NOCOLOR) s.winner = BLACK; }
        }
    }
    return s;
}

__device__ __forceinline__ void reset_board(SmemBoard& b) {
#pragma unroll
    for (int w = 0; w < 12; ++w) b.base[w * b.stride] = kOpening[w];
}

struct Meta {
    int move_count, player, red_score, black_score;
    uint32_t ctr;
    uint8_t flags;
    __device__ __forceinline__ void load(const xq_env_rec* r) {
        const uint4 m = reinterpret_cast<const uint4*>(r)[3];
        move_count = m.x & 0xFFFF; player = (m.x >> 16) & 0xFF; flags = (uint8_t)(m.x >> 24);
        red_score = (int)m.y; black_score = (int)m.z; ctr = m.w;
    }
    __device__ __forceinline__ void store(xq_env_rec* r) const {
        reinterpret_cast<uint4*>(r)[3] = make_uint4((uint32_t)(move_count & 0xFFFF) | ((uint32_t)player << 16) | ((uint32_t)flags << 24),
                                                    (uint32_t)red_score, (uint32_t)black_score, ctr);
    }
    __device__ __forceinline__ void reset() { move_count = 0; player = RED; red_score = 0; black_score = 0; }
};

// ChessBoard::movePiece after validation (src/chessboard.cpp:43-63); returns the captured code
__device__ __forceinline__ int apply_move(SmemBoard& b, Meta& m, int from, int to) {
    const int cap = b.get(to);
    b.set(to, b.get(from));
    b.set(from, 0);
    if (cap != 0) {
        const int sc = piece_score(type_of(cap));
        if (cap >= 8) m.red_score += sc; else m.black_score += sc;   // captured Black => Red scores (:53-57)
    }
    m.move_count++;
    m.player ^= 1;
    return cap;
}


}  // namespace xq
