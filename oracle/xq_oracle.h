/* xq_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the reference's self-play hot path (Qervas/cn_chess_ai):
 * the ChessBoard rules, ChessAI's action enumeration / state encoding / reward, and
 * the DQN math of src/dqn.cu in FP64.  It is the CHECKER for the CUDA path: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.  The product (cn_chess_ai_b200/) never links, imports or calls it.
 *
 * Parity pinning: the reference ships no tests or golden vectors, so this
 * restatement is pinned against the reference's OWN sources compiled unmodified
 * (oracle/_ref/libxq_ref.so, see oracle/Makefile) by tests/test_oracle_vs_ref.py, and
 * against fixtures generated from that build (tests/golden/, tests/golden/make_golden.py).
 *
 * Every function cites the reference file:line it follows (paths under /root/reference).
 */
#ifndef XQ_ORACLE_H
#define XQ_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Packed environment record, identical to the HBM record of the CUDA path
 * (include/xq.h: xq_env_rec).  Square s = row*9+col lives in nibble (s&7) of sq[s>>3];
 * code 0 empty, 1..7 Red {General,Advisor,Elephant,Horse,Chariot,Cannon,Soldier}
 * (= PieceType, include/chessboard.h:8-10), 8..14 Black (= PieceType+7), which is the
 * one-hot channel+1 of ChessAI::getStateRepresentation (src/chessai.cpp:278-282). */
typedef struct {
    uint32_t sq[12];      /* 90 nibbles + 6 zero nibbles                      */
    uint16_t move_count;  /* ChessBoard::moveCount                            */
    uint8_t player;       /* ChessBoard::currentPlayer: 0 Red, 1 Black        */
    uint8_t flags;        /* reserved, 0                                      */
    int32_t red_score;    /* ChessBoard::redScore                             */
    int32_t black_score;  /* ChessBoard::blackScore                           */
    uint32_t ctr;         /* plies applied to this slot so far (RNG counter)  */
} xqo_env;

#define XQO_MAX_ACTIONS 128
#define XQO_ACTION(from, to) ((uint16_t)(((from) << 7) | (to)))
#define XQO_FROM(a) ((a) >> 7)
#define XQO_TO(a) ((a) & 127)

void xqo_reset(xqo_env* e);
int xqo_piece_at(const xqo_env* e, int row, int col);
int xqo_is_valid_move(const xqo_env* e, int fr, int fc, int tr, int tc);
int xqo_valid_moves(const xqo_env* e, int row, int col, uint8_t* to_sq);
int xqo_all_actions(const xqo_env* e, int player, uint16_t* actions);
int xqo_move(xqo_env* e, int fr, int fc, int tr, int tc);
int xqo_game_over(const xqo_env* e);
int xqo_winner(const xqo_env* e);
int xqo_evaluate(const xqo_env* e, int player, int move_count);
int xqo_evaluate_int(const xqo_env* e, int player, int move_count);
void xqo_state(const xqo_env* e, double* out1260);
int xqo_piece_score(int type);

/* counter RNG shared (by specification) with the CUDA path */
uint64_t xqo_rng(uint64_t seed, uint64_t env_id, uint32_t ctr);

/* per-ply trace record, identical to include/xq.h: xq_trace_rec */
typedef struct {
    uint16_t action;   /* (from<<7)|to                               */
    uint8_t n_legal;   /* size of the ordered action list            */
    uint8_t flags;     /* bit0 done, bits1-2 winner (0 R,1 B,2 none), bits 4-7 captured code */
    int32_t reward;    /* ChessAI::evaluateBoard for the mover       */
} xqo_trace;

typedef struct {
    uint64_t steps, games, red_wins, black_wins, cap_games, captures;
    int64_t reward_sum;
    uint64_t legal_sum;
} xqo_stats;

void xqo_rollout_random(xqo_env* envs, long n_envs, uint64_t env_id0, uint64_t seed, int n_plies,
                        xqo_trace* trace /* [n_plies][n_envs] or NULL */, xqo_stats* stats);
double xqo_bench_rollout_random(int n_threads, long envs_per_thread, int n_plies, uint64_t seed, long* total);

/* batched per-position queries (for differential tests) */
void xqo_batch_all_actions(const xqo_env* envs, long n, uint8_t* counts, uint16_t* actions /* [n][128] */);
/* opt-in strict legality (not a rule of the reference): self-check and flying-general rejection on top of its generators */
int xqo_all_actions_strict(const xqo_env* e, int player, uint16_t* actions);
void xqo_batch_all_actions_strict(const xqo_env* envs, long n, uint8_t* counts, uint16_t* actions /* [n][128], 0xFFFF past the count */);
void xqo_batch_step(xqo_env* envs, long n, const uint16_t* actions, int32_t* reward, uint8_t* done, uint8_t* winner,
                    uint8_t* captured, uint8_t* valid);

/* ---- DQN math (FP64), layers[] = sizes, weights [layer][out][in] ‖, biases ‖ (src/dqn.cu:112-140) ---- */
void xqo_nn_forward(const int* layers, int n_layers, const double* w, const double* b, const double* x, double* out);
void xqo_nn_backprop(const int* layers, int n_layers, double* w, double* b, const double* x, const double* target,
                     double lr, int corrected);
/* gradient of the implicit loss 1/2||a-t||^2 at frozen weights (what backprop subtracts, / lr) */
void xqo_nn_grad(const int* layers, int n_layers, const double* w, const double* b, const double* x, const double* target,
                 int corrected, double* gw, double* gb);
int xqo_select_action(const double* q, const uint16_t* actions, int n, uint32_t coin31, uint32_t idx31, double eps);
uint32_t xqo_eps_threshold(double eps);
void xqo_td_target(const double* q_s, const double* q_next, int n_out, int a_to, double reward, int done, double gamma,
                   double* target);

/* the TD gradient of a batch: per-sample forward x2 + xqo_td_target + xqo_nn_grad at frozen weights, summed, on n_threads host threads */
void xqo_td_batch_grad(const int* layers, int n_layers, const double* w, const double* b, const double* tw, const double* tb, const double* x,
                       const double* x2, const int32_t* to, const int32_t* reward, const uint8_t* done, double gamma, int corrected, long n,
                       int n_threads, double* gw, double* gb, double* loss);

#ifdef __cplusplus
}
#endif
#endif
