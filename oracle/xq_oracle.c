/* xq_oracle.c -- TEST INFRASTRUCTURE ONLY (see xq_oracle.h).
 * CPU restatement of the reference hot path; every function cites the reference
 * file:line it follows.  Built with -ffp-contract=off. */
#include "xq_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

enum { EMPTY = 0, GENERAL = 1, ADVISOR = 2, ELEPHANT = 3, HORSE = 4, CHARIOT = 5, CANNON = 6, SOLDIER = 7 };
enum { RED = 0, BLACK = 1, NONE = 2 };
#define ROWS 10
#define COLS 9

static inline int get_code(const xqo_env* e, int s) { return (e->sq[s >> 3] >> ((s & 7) * 4)) & 15; }
static inline void set_code(xqo_env* e, int s, int code) {
    e->sq[s >> 3] = (e->sq[s >> 3] & ~(15u << ((s & 7) * 4))) | ((uint32_t)code << ((s & 7) * 4));
}
static inline int type_of(int code) { return code == 0 ? EMPTY : (code <= 7 ? code : code - 7); }
static inline int color_of(int code) { return code == 0 ? NONE : (code <= 7 ? RED : BLACK); }

/* ChessBoard::isInsideBoard, src/chessboard.cpp:323-325 */
static inline int inside(int r, int c) { return r >= 0 && r < ROWS && c >= 0 && c < COLS; }
/* ChessBoard::getPieceAt, src/chessboard.cpp:31-36 (off-board => Empty/None) */
int xqo_piece_at(const xqo_env* e, int r, int c) { return inside(r, c) ? get_code(e, r * COLS + c) : 0; }
/* include/chessboard.h:65-75 */
static inline int in_red_palace(int r, int c) { return r >= 0 && r <= 2 && c >= 3 && c <= 5; }
static inline int in_black_palace(int r, int c) { return r >= 7 && r <= 9 && c >= 3 && c <= 5; }
static inline int in_own_side(int color, int r) { return color == RED ? (r >= 0 && r <= 4) : (r >= 5 && r <= 9); }

/* getPieceScore, src/chessboard.cpp:443-454 (PieceScore, include/chessboard.h:23-31) */
int xqo_piece_score(int type) {
    switch (type) {
        case GENERAL: return 1000;
        case ADVISOR: return 20;
        case ELEPHANT: return 20;
        case HORSE: return 40;
        case CHARIOT: return 90;
        case CANNON: return 45;
        case SOLDIER: return 10;
        default: return 0;
    }
}

/* ChessBoard::initializeBoard / reset, src/chessboard.cpp:8-29,95-102 */
void xqo_reset(xqo_env* e) {
    static const int back[9] = {CHARIOT, HORSE, ELEPHANT, ADVISOR, GENERAL, ADVISOR, ELEPHANT, HORSE, CHARIOT};
    uint32_t ctr = e->ctr;
    memset(e, 0, sizeof(*e));
    e->ctr = ctr;
    for (int c = 0; c < 9; ++c) {
        set_code(e, 0 * COLS + c, back[c]);
        set_code(e, 9 * COLS + c, back[c] + 7);
    }
    set_code(e, 2 * COLS + 1, CANNON); set_code(e, 2 * COLS + 7, CANNON);
    set_code(e, 7 * COLS + 1, CANNON + 7); set_code(e, 7 * COLS + 7, CANNON + 7);
    for (int c = 0; c < 9; c += 2) { set_code(e, 3 * COLS + c, SOLDIER); set_code(e, 6 * COLS + c, SOLDIER + 7); }
}

/* per-piece predicates, src/chessboard.cpp:328-440 */
static int valid_general(int fr, int fc, int tr, int tc) {                       /* :328-343 */
    int from_p = in_red_palace(fr, fc) || in_black_palace(fr, fc);
    int to_p = in_red_palace(tr, tc) || in_black_palace(tr, tc);
    if (!from_p || !to_p) return 0;
    return abs(tr - fr) + abs(tc - fc) == 1;
}
static int valid_advisor(int fr, int fc, int tr, int tc) {                       /* :346-353 */
    int to_p = in_red_palace(tr, tc) || in_black_palace(tr, tc);
    return to_p && abs(tr - fr) == 1 && abs(tc - fc) == 1;
}
static int valid_elephant(const xqo_env* e, int fr, int fc, int tr, int tc) {    /* :355-367 */
    int no_cross = (fr < 5 && tr < 5) || (fr >= 5 && tr >= 5);
    int mr = (fr + tr) / 2, mc = (fc + tc) / 2;
    int clear = type_of(xqo_piece_at(e, mr, mc)) == EMPTY;
    return abs(tr - fr) == 2 && abs(tc - fc) == 2 && no_cross && clear;
}
static int valid_horse(const xqo_env* e, int fr, int fc, int tr, int tc) {       /* :369-380 */
    int rd = abs(tr - fr), cd = abs(tc - fc);
    if ((rd == 2 && cd == 1) || (rd == 1 && cd == 2)) {
        int br = fr + (tr - fr) / 2, bc = fc + (tc - fc) / 2;   /* C truncating division */
        return type_of(xqo_piece_at(e, br, bc)) == EMPTY;
    }
    return 0;
}
static int count_between(const xqo_env* e, int fr, int fc, int tr, int tc) {
    int step = (fr == tr) ? (tc > fc ? 1 : -1) : (tr > fr ? 1 : -1);
    int start = (fr == tr) ? fc : fr, end = (fr == tr) ? tc : tr, n = 0;
    for (int i = start + step; i != end; i += step)
        if (type_of(xqo_piece_at(e, fr == tr ? fr : i, fr == tr ? i : fc)) != EMPTY) ++n;
    return n;
}
static int valid_chariot(const xqo_env* e, int fr, int fc, int tr, int tc) {     /* :382-397 */
    if (fr != tr && fc != tc) return 0;
    return count_between(e, fr, fc, tr, tc) == 0;
}
static int valid_cannon(const xqo_env* e, int fr, int fc, int tr, int tc) {      /* :399-421 */
    if (fr != tr && fc != tc) return 0;
    int n = count_between(e, fr, fc, tr, tc);
    return type_of(xqo_piece_at(e, tr, tc)) == EMPTY ? n == 0 : n == 1;
}
static int valid_soldier(const xqo_env* e, int fr, int fc, int tr, int tc) {     /* :423-440 */
    int rd = tr - fr, cd = abs(tc - fc);
    if (color_of(xqo_piece_at(e, fr, fc)) == RED) {
        if (fr < 5) return rd == 1 && cd == 0;
        return (rd == 1 && cd == 0) || (rd == 0 && cd == 1);
    }
    if (fr >= 5) return rd == -1 && cd == 0;
    return (rd == -1 && cd == 0) || (rd == 0 && cd == 1);
}

/* ChessBoard::isValidMove, src/chessboard.cpp:66-93: no turn test, no king-safety test (SURVEY F1,F2) */
int xqo_is_valid_move(const xqo_env* e, int fr, int fc, int tr, int tc) {
    if (!inside(fr, fc) || !inside(tr, tc)) return 0;
    int from = get_code(e, fr * COLS + fc), to = get_code(e, tr * COLS + tc);
    if (type_of(from) == EMPTY) return 0;
    if (color_of(from) == color_of(to) && type_of(to) != EMPTY) return 0;
    switch (type_of(from)) {
        case GENERAL: return valid_general(fr, fc, tr, tc);
        case ADVISOR: return valid_advisor(fr, fc, tr, tc);
        case ELEPHANT: return valid_elephant(e, fr, fc, tr, tc);
        case HORSE: return valid_horse(e, fr, fc, tr, tc);
        case CHARIOT: return valid_chariot(e, fr, fc, tr, tc);
        case CANNON: return valid_cannon(e, fr, fc, tr, tc);
        case SOLDIER: return valid_soldier(e, fr, fc, tr, tc);
        default: return 0;
    }
}

/* ChessBoard::getValidMoves + generate*Moves, src/chessboard.cpp:112-283.  Destination
 * order = direction order as written in each generator. */
int xqo_valid_moves(const xqo_env* e, int row, int col, uint8_t* out) {
    int code = xqo_piece_at(e, row, col), n = 0;
    if (type_of(code) == EMPTY) return 0;
    int color = color_of(code);
    switch (type_of(code)) {
        case GENERAL: {                                                           /* :149-160 */
            static const int d[4][2] = {{1, 0}, {-1, 0}, {0, 1}, {0, -1}};
            for (int k = 0; k < 4; ++k) {
                int nr = row + d[k][0], nc = col + d[k][1];
                if (inside(nr, nc) && xqo_is_valid_move(e, row, col, nr, nc)) out[n++] = (uint8_t)(nr * COLS + nc);
            }
            break;
        }
        case ADVISOR: {                                                           /* :162-177 */
            static const int d[4][2] = {{1, 1}, {1, -1}, {-1, 1}, {-1, -1}};
            for (int k = 0; k < 4; ++k) {
                int nr = row + d[k][0], nc = col + d[k][1];
                if (inside(nr, nc) &&
                    ((color == RED && in_red_palace(nr, nc)) || (color == BLACK && in_black_palace(nr, nc))) &&
                    xqo_is_valid_move(e, row, col, nr, nc))
                    out[n++] = (uint8_t)(nr * COLS + nc);
            }
            break;
        }
        case ELEPHANT: {                                                          /* :179-196 */
            static const int d[4][2] = {{2, 2}, {2, -2}, {-2, 2}, {-2, -2}};
            for (int k = 0; k < 4; ++k) {
                int nr = row + d[k][0], nc = col + d[k][1];
                int mr = row + d[k][0] / 2, mc = col + d[k][1] / 2;
                if (inside(nr, nc) && in_own_side(color, nr) && type_of(xqo_piece_at(e, mr, mc)) == EMPTY &&
                    xqo_is_valid_move(e, row, col, nr, nc))
                    out[n++] = (uint8_t)(nr * COLS + nc);
            }
            break;
        }
        case CHARIOT: {                                                           /* :198-218 */
            static const int d[4][2] = {{0, 1}, {0, -1}, {1, 0}, {-1, 0}};
            for (int k = 0; k < 4; ++k) {
                int nr = row + d[k][0], nc = col + d[k][1];
                while (inside(nr, nc)) {
                    if (xqo_is_valid_move(e, row, col, nr, nc)) {
                        out[n++] = (uint8_t)(nr * COLS + nc);
                        if (type_of(xqo_piece_at(e, nr, nc)) != EMPTY) break;
                    } else {
                        break;
                    }
                    nr += d[k][0]; nc += d[k][1];
                }
            }
            break;
        }
        case CANNON: {                                                            /* :220-246 */
            static const int d[4][2] = {{0, 1}, {0, -1}, {1, 0}, {-1, 0}};
            for (int k = 0; k < 4; ++k) {
                int nr = row + d[k][0], nc = col + d[k][1], screen = 0;
                while (inside(nr, nc)) {
                    if (!screen) {
                        if (type_of(xqo_piece_at(e, nr, nc)) == EMPTY) out[n++] = (uint8_t)(nr * COLS + nc);
                        else screen = 1;
                    } else if (type_of(xqo_piece_at(e, nr, nc)) != EMPTY && xqo_is_valid_move(e, row, col, nr, nc)) {
                        out[n++] = (uint8_t)(nr * COLS + nc);
                        break;
                    }
                    nr += d[k][0]; nc += d[k][1];
                }
            }
            break;
        }
        case HORSE: {                                                             /* :248-263 */
            static const int d[8][2] = {{1, 2}, {1, -2}, {-1, 2}, {-1, -2}, {2, 1}, {2, -1}, {-2, 1}, {-2, -1}};
            for (int k = 0; k < 8; ++k) {
                int nr = row + d[k][0], nc = col + d[k][1];
                int lr = row + d[k][0] / 2, lc = col + d[k][1] / 2;
                if (inside(nr, nc) && type_of(xqo_piece_at(e, lr, lc)) == EMPTY && xqo_is_valid_move(e, row, col, nr, nc))
                    out[n++] = (uint8_t)(nr * COLS + nc);
            }
            break;
        }
        case SOLDIER: {                                                           /* :265-283 */
            int fwd = color == RED ? 1 : -1, nr = row + fwd;
            if (inside(nr, col) && xqo_is_valid_move(e, row, col, nr, col)) out[n++] = (uint8_t)(nr * COLS + col);
            if ((color == RED && row > 4) || (color == BLACK && row < 5)) {
                int cs[2] = {col - 1, col + 1};
                for (int k = 0; k < 2; ++k)
                    if (inside(row, cs[k]) && xqo_is_valid_move(e, row, col, row, cs[k])) out[n++] = (uint8_t)(row * COLS + cs[k]);
            }
            break;
        }
        default: break;
    }
    return n;
}

/* ChessAI::getAllValidActions, src/chessai.cpp:347-368: row-major scan, piece.color == player */
int xqo_all_actions(const xqo_env* e, int player, uint16_t* actions) {
    int n = 0;
    uint8_t to[32];
    for (int r = 0; r < ROWS; ++r)
        for (int c = 0; c < COLS; ++c)
            if (color_of(get_code(e, r * COLS + c)) == player) {
                int m = xqo_valid_moves(e, r, c, to);
                for (int k = 0; k < m && n < XQO_MAX_ACTIONS; ++k) actions[n++] = XQO_ACTION(r * COLS + c, to[k]);
            }
    return n;
}

/* ChessBoard::movePiece, src/chessboard.cpp:38-64; returns captured code, 0 for quiet or rejected */
int xqo_move(xqo_env* e, int fr, int fc, int tr, int tc) {
    if (!xqo_is_valid_move(e, fr, fc, tr, tc)) return 0;
    int f = fr * COLS + fc, t = tr * COLS + tc;
    int cap = get_code(e, t);
    set_code(e, t, get_code(e, f));
    set_code(e, f, 0);
    if (type_of(cap) != EMPTY) {
        int s = xqo_piece_score(type_of(cap));
        if (color_of(cap) == RED) e->black_score += s; else e->red_score += s;
    }
    e->move_count++;
    e->player = e->player == RED ? BLACK : RED;
    return cap;
}

/* ChessBoard::checkGameOver, src/chessboard.cpp:286-309 (maxMovePerGame 200, include/chessboard.h:63) */
int xqo_game_over(const xqo_env* e) {
    int ra = 0, ba = 0;
    if (e->move_count >= 200) return 1;
    for (int i = 0; i < 90; ++i) {
        int c = get_code(e, i);
        if (c == GENERAL) ra = 1; else if (c == GENERAL + 7) ba = 1;
        if (ra && ba) return 0;
    }
    return 1;
}
/* ChessBoard::getWinner, src/chessboard.cpp:312-320: colour of the first General in index order */
int xqo_winner(const xqo_env* e) {
    for (int i = 0; i < 90; ++i) {
        int c = get_code(e, i);
        if (type_of(c) == GENERAL) return color_of(c);
    }
    return NONE;
}

/* ChessAI::evaluateBoard, src/chessai.cpp:311-345, restated literally:
 * `int score; score -= moveCount * 0.1;` == (int)((double)score - (double)moveCount*0.1) */
int xqo_evaluate(const xqo_env* e, int player, int move_count) {
    int score = 0;
    for (int i = 0; i < 90; ++i) {
        int c = get_code(e, i);
        if (color_of(c) == player) score += xqo_piece_score(type_of(c));
        else if (color_of(c) != NONE) score -= xqo_piece_score(type_of(c));
    }
    volatile double prod = (double)move_count * 0.1; /* one rounding, then the subtraction: no FMA */
    double v = (double)score - prod;
    return (int)v;
}
/* the integer-only form the CUDA path uses; tests prove it equals xqo_evaluate */
int xqo_evaluate_int(const xqo_env* e, int player, int move_count) {
    int score = 0;
    for (int i = 0; i < 90; ++i) {
        int c = get_code(e, i);
        if (color_of(c) == player) score += xqo_piece_score(type_of(c));
        else if (color_of(c) != NONE) score -= xqo_piece_score(type_of(c));
    }
    return (10 * score - move_count) / 10;
}

/* ChessAI::getStateRepresentation, src/chessai.cpp:268-289 */
void xqo_state(const xqo_env* e, double* out) {
    memset(out, 0, 1260 * sizeof(double));
    for (int s = 0; s < 90; ++s) {
        int c = get_code(e, s);
        if (c) out[s * 14 + (c - 1)] = 1.0;
    }
}

/* Counter RNG (framework-defined; the reference is unseedable): splitmix64
 * finaliser over seed + env_id*phi + ctr*C.  idx31 = x>>33, coin31 = x & 0x7fffffff. */
uint64_t xqo_rng(uint64_t seed, uint64_t env_id, uint32_t ctr) {
    uint64_t z = seed + env_id * 0x9E3779B97F4A7C15ull + (uint64_t)ctr * 0xD1B54A32D192ED03ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* One random-policy ply = loop body of ChessAI::train without the network
 * (src/chessai.cpp:96-119): list -> pick -> movePiece -> evaluateBoard -> checkGameOver,
 * reset on terminal (chessai.cpp:90).  Terminal here is checkGameOver() (startSelfPlay
 * semantics, :227); train()'s one-ply-early `done` is exposed by the episode driver. */
static void step_random(xqo_env* e, uint64_t env_id, uint64_t seed, xqo_trace* tr, xqo_stats* st) {
    uint16_t acts[XQO_MAX_ACTIONS];
    if (xqo_game_over(e)) { uint32_t c = e->ctr; xqo_reset(e); e->ctr = c; }  /* chessai.cpp:90,96: a finished board is never stepped */
    int mover = e->player;
    int n = xqo_all_actions(e, mover, acts);
    if (n == 0) {  /* chessai.cpp:100-103: no action => the episode loop ends; we restart the game */
        if (tr) { tr->action = 0xFFFF; tr->n_legal = 0; tr->flags = 1 | (2 << 1); tr->reward = 0; }
        uint32_t c = e->ctr + 1; xqo_reset(e); e->ctr = c;
        if (st) st->games++;
        return;
    }
    uint64_t x = xqo_rng(seed, env_id, e->ctr);
    uint16_t a = acts[(uint32_t)(x >> 33) % (uint32_t)n];
    int f = XQO_FROM(a), t = XQO_TO(a);
    int cap = xqo_move(e, f / 9, f % 9, t / 9, t % 9);
    int reward = xqo_evaluate(e, mover, e->move_count);
    int done = xqo_game_over(e);
    int win = done ? xqo_winner(e) : NONE;
    e->ctr++;
    if (tr) { tr->action = a; tr->n_legal = (uint8_t)n; tr->flags = (uint8_t)((done ? 1 : 0) | (win << 1) | (cap << 4)); tr->reward = reward; }
    if (st) {
        st->steps++; st->legal_sum += (uint64_t)n; st->reward_sum += reward; if (cap) st->captures++;
        if (done) { st->games++; if (win == RED) st->red_wins++; else if (win == BLACK) st->black_wins++; if (e->move_count < 200) st->cap_games++; }
    }
    if (done) { uint32_t c = e->ctr; xqo_reset(e); e->ctr = c; }
}

void xqo_rollout_random(xqo_env* envs, long n_envs, uint64_t env_id0, uint64_t seed, int n_plies, xqo_trace* trace,
                        xqo_stats* stats) {
    if (stats) memset(stats, 0, sizeof(*stats));
    for (long i = 0; i < n_envs; ++i)
        for (int p = 0; p < n_plies; ++p)
            step_random(&envs[i], env_id0 + (uint64_t)i, seed, trace ? &trace[(long)p * n_envs + i] : NULL, stats);
}

typedef struct { long n; int plies; uint64_t seed; uint64_t id0; long steps; } bench_arg;
static void* bench_thread(void* p) {
    bench_arg* a = (bench_arg*)p;
    xqo_env* envs = (xqo_env*)calloc((size_t)a->n, sizeof(xqo_env));
    for (long i = 0; i < a->n; ++i) xqo_reset(&envs[i]);
    xqo_stats st;
    xqo_rollout_random(envs, a->n, a->id0, a->seed, a->plies, NULL, &st);
    a->steps = (long)st.steps;
    free(envs);
    return NULL;
}
double xqo_bench_rollout_random(int n_threads, long envs_per_thread, int n_plies, uint64_t seed, long* total) {
    pthread_t th[256];
    bench_arg args[256];
    if (n_threads > 256) n_threads = 256;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < n_threads; ++t) {
        args[t].n = envs_per_thread; args[t].plies = n_plies; args[t].seed = seed; args[t].id0 = (uint64_t)t * (uint64_t)envs_per_thread; args[t].steps = 0;
        pthread_create(&th[t], NULL, bench_thread, &args[t]);
    }
    long tot = 0;
    for (int t = 0; t < n_threads; ++t) { pthread_join(th[t], NULL); tot += args[t].steps; }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (total) *total = tot;
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* Opt-in STRICT legality (include/xq.h: xq_env_legal_moves_strict) -- not a rule of the reference, stated on top of its generators:
 * the action is kept iff after it (a) no enemy piece has the mover's first General (square order) among the destinations
 * ChessBoard::getValidMoves generates for it, and (b) the two first Generals do not face each other on a file with nothing between.
 * Pinned against the reference's own classes composed the same way (ref_wrap.cpp: ref_env_all_actions_strict). */
int xqo_all_actions_strict(const xqo_env* e0, int player, uint16_t* actions) {
    uint16_t all[XQO_MAX_ACTIONS];
    uint8_t to[32];
    int n = xqo_all_actions(e0, player, all), kept = 0;
    for (int k = 0; k < n; ++k) {
        xqo_env e = *e0;
        int from = all[k] >> 7, t = all[k] & 127;
        set_code(&e, t, get_code(&e, from));
        set_code(&e, from, 0);
        int g = -1, eg = -1, safe = 1;
        for (int s = 0; s < 90; ++s) {
            int c = get_code(&e, s);
            if (type_of(c) == GENERAL && color_of(c) == player && g < 0) g = s;
            if (type_of(c) == GENERAL && color_of(c) != player && eg < 0) eg = s;
        }
        if (g >= 0) {
            for (int s = 0; s < 90 && safe; ++s) {
                int c = get_code(&e, s);
                if (c != 0 && color_of(c) != player) {
                    int m = xqo_valid_moves(&e, s / COLS, s % COLS, to);
                    for (int j = 0; j < m; ++j) if (to[j] == g) safe = 0;
                }
            }
            if (safe && eg >= 0 && g % COLS == eg % COLS) {
                int lo = g < eg ? g : eg, hi = g < eg ? eg : g, between = 0;
                for (int s = lo + COLS; s < hi; s += COLS) between += get_code(&e, s) != 0;
                if (between == 0) safe = 0;
            }
        }
        if (safe) actions[kept++] = all[k];
    }
    return kept;
}
void xqo_batch_all_actions_strict(const xqo_env* envs, long n, uint8_t* counts, uint16_t* actions) {
    for (long i = 0; i < n; ++i) {
        uint16_t* a = actions + i * XQO_MAX_ACTIONS;
        memset(a, 0xFF, XQO_MAX_ACTIONS * sizeof(uint16_t));
        counts[i] = (uint8_t)xqo_all_actions_strict(&envs[i], envs[i].player, a);
    }
}

void xqo_batch_all_actions(const xqo_env* envs, long n, uint8_t* counts, uint16_t* actions) {
    for (long i = 0; i < n; ++i) {
        uint16_t* a = actions + i * XQO_MAX_ACTIONS;
        memset(a, 0, XQO_MAX_ACTIONS * sizeof(uint16_t));
        counts[i] = (uint8_t)xqo_all_actions(&envs[i], envs[i].player, a);
    }
}

/* movePiece + evaluateBoard(mover, post-move moveCount) + checkGameOver + getWinner, no reset.
 * A rejected move leaves the record untouched (chessboard.cpp:39-41) and reports valid=0. */
void xqo_batch_step(xqo_env* envs, long n, const uint16_t* actions, int32_t* reward, uint8_t* done, uint8_t* winner,
                    uint8_t* captured, uint8_t* valid) {
    for (long i = 0; i < n; ++i) {
        xqo_env* e = &envs[i];
        int f = XQO_FROM(actions[i]), t = XQO_TO(actions[i]), mover = e->player;
        int ok = f < 90 && t < 90 && xqo_is_valid_move(e, f / 9, f % 9, t / 9, t % 9);
        int cap = ok ? xqo_move(e, f / 9, f % 9, t / 9, t % 9) : 0;
        if (ok) e->ctr++;
        valid[i] = (uint8_t)ok; captured[i] = (uint8_t)cap;
        reward[i] = xqo_evaluate(e, mover, e->move_count);
        done[i] = (uint8_t)xqo_game_over(e);
        winner[i] = (uint8_t)xqo_winner(e);
    }
}

/* ---------------------------- DQN math, FP64 ---------------------------- */
/* NeuralNetwork::forward + forwardKernel(6-arg), src/dqn.cu:184-195,199-260: sum from 0, bias last */
void xqo_nn_forward(const int* layers, int n_layers, const double* w, const double* b, const double* x, double* out) {
    int L = n_layers - 1, maxw = 0;
    for (int l = 0; l < n_layers; ++l) if (layers[l] > maxw) maxw = layers[l];
    double* cur = (double*)malloc(sizeof(double) * (size_t)maxw);
    double* nxt = (double*)malloc(sizeof(double) * (size_t)maxw);
    memcpy(cur, x, sizeof(double) * (size_t)layers[0]);
    size_t wo = 0, bo = 0;
    for (int l = 0; l < L; ++l) {
        int in = layers[l], on = layers[l + 1];
        for (int o = 0; o < on; ++o) {
            double sum = 0.0;
            for (int i = 0; i < in; ++i) sum += cur[i] * w[wo + (size_t)o * in + i];
            sum += b[bo + o];
            nxt[o] = tanh(sum);
        }
        wo += (size_t)in * on; bo += (size_t)on;
        double* t = cur; cur = nxt; nxt = t;
    }
    memcpy(out, cur, sizeof(double) * (size_t)layers[L]);
    free(cur); free(nxt);
}

/* Shared by xqo_nn_backprop / xqo_nn_grad: forward with z (forwardKernel 7-arg, :275-286, sum starts
 * at the bias), output delta (:288-295), hidden delta as written (:297-308 with the call-site sizes of
 * :406-423) or corrected.  deltas[l] has layers[l+1] entries. */
static void nn_deltas(const int* layers, int n_layers, const double* w, const double* b, const double* x,
                      const double* target, int corrected, double** act, double** delta) {
    int L = n_layers - 1;
    size_t* wofs = (size_t*)malloc(sizeof(size_t) * (size_t)L);
    size_t* bofs = (size_t*)malloc(sizeof(size_t) * (size_t)L);
    double** z = (double**)malloc(sizeof(double*) * (size_t)L);
    size_t wo = 0, bo = 0;
    act[0] = (double*)malloc(sizeof(double) * (size_t)layers[0]);
    memcpy(act[0], x, sizeof(double) * (size_t)layers[0]);
    for (int l = 0; l < L; ++l) {
        int in = layers[l], on = layers[l + 1];
        wofs[l] = wo; bofs[l] = bo;
        act[l + 1] = (double*)malloc(sizeof(double) * (size_t)on);
        z[l] = (double*)malloc(sizeof(double) * (size_t)on);
        delta[l] = (double*)calloc((size_t)on, sizeof(double));
        for (int o = 0; o < on; ++o) {
            double sum = b[bo + o];
            for (int i = 0; i < in; ++i) sum += act[l][i] * w[wo + (size_t)o * in + i];
            z[l][o] = sum; act[l + 1][o] = tanh(sum);
        }
        wo += (size_t)in * on; bo += (size_t)on;
    }
    for (int o = 0; o < layers[L]; ++o) {
        double err = act[L][o] - target[o];
        double der = 1 - tanh(z[L - 1][o]) * tanh(z[L - 1][o]);
        delta[L - 1][o] = err * der;
    }
    for (int l = L - 2; l >= 0; --l) {
        const double* Wn = w + wofs[l + 1];
        int width = layers[l + 1];
        if (!corrected) {
            int inputSize = layers[l + 1], outputSize = layers[l];
            for (int idx = 0; idx < width && idx < outputSize; ++idx) {
                double sum = 0.0;
                for (int i = 0; i < inputSize; ++i) sum += Wn[(size_t)i * outputSize + idx] * delta[l + 1][i];
                double der = 1 - tanh(z[l][idx]) * tanh(z[l][idx]);
                delta[l][idx] = sum * der;
            }
        } else {
            int nextw = layers[l + 2];
            for (int j = 0; j < width; ++j) {
                double sum = 0.0;
                for (int o = 0; o < nextw; ++o) sum += Wn[(size_t)o * width + j] * delta[l + 1][o];
                double der = 1 - tanh(z[l][j]) * tanh(z[l][j]);
                delta[l][j] = sum * der;
            }
        }
    }
    for (int l = 0; l < L; ++l) free(z[l]);
    free(z); free(wofs); free(bofs);
}

/* NeuralNetwork::backpropagate, src/dqn.cu:323-467; update = updateWeightsBiasesKernel :310-319 */
void xqo_nn_backprop(const int* layers, int n_layers, double* w, double* b, const double* x, const double* target,
                     double lr, int corrected) {
    int L = n_layers - 1;
    double** act = (double**)malloc(sizeof(double*) * (size_t)(L + 1));
    double** delta = (double**)malloc(sizeof(double*) * (size_t)L);
    nn_deltas(layers, n_layers, w, b, x, target, corrected, act, delta);
    size_t wo = 0, bo = 0;
    for (int l = 0; l < L; ++l) {
        int in = layers[l], on = layers[l + 1];
        for (int o = 0; o < on; ++o) {
            b[bo + o] -= lr * delta[l][o];
            for (int i = 0; i < in; ++i) w[wo + (size_t)o * in + i] -= lr * delta[l][o] * act[l][i];
        }
        wo += (size_t)in * on; bo += (size_t)on;
    }
    for (int l = 0; l < L; ++l) { free(act[l]); free(delta[l]); }
    free(act[L]); free(act); free(delta);
}

void xqo_nn_grad(const int* layers, int n_layers, const double* w, const double* b, const double* x, const double* target,
                 int corrected, double* gw, double* gb) {
    int L = n_layers - 1;
    double** act = (double**)malloc(sizeof(double*) * (size_t)(L + 1));
    double** delta = (double**)malloc(sizeof(double*) * (size_t)L);
    nn_deltas(layers, n_layers, w, b, x, target, corrected, act, delta);
    size_t wo = 0, bo = 0;
    for (int l = 0; l < L; ++l) {
        int in = layers[l], on = layers[l + 1];
        for (int o = 0; o < on; ++o) {
            gb[bo + o] = delta[l][o];
            for (int i = 0; i < in; ++i) gw[wo + (size_t)o * in + i] = delta[l][o] * act[l][i];
        }
        wo += (size_t)in * on; bo += (size_t)on;
    }
    for (int l = 0; l < L; ++l) { free(act[l]); free(delta[l]); }
    free(act[L]); free(act); free(delta);
}

/* smallest T with (double)c / RAND_MAX < eps  <=>  c < T, c in [0, RAND_MAX]  (src/dqn.cpp:30-31) */
uint32_t xqo_eps_threshold(double eps) {
    const double rm = 2147483647.0;
    if (!(eps > 0.0)) return 0;
    if (eps > 1.0) return 0x80000000u;
    double g = ceil(eps * rm);
    int64_t t = (int64_t)g;
    while (t > 0 && !((double)(t - 1) / rm < eps)) --t;
    while (t <= 2147483647LL && ((double)t / rm < eps)) ++t;
    return (uint32_t)t;
}

/* DQN::selectAction, src/dqn.cpp:24-56, with the two rand() results supplied:
 * coin31 = first rand(), idx31 = second rand() (only consumed when exploring).
 * Greedy = FIRST action maximising q[action.to] (strict >, from -inf).  Returns the list index. */
int xqo_select_action(const double* q, const uint16_t* actions, int n, uint32_t coin31, uint32_t idx31, double eps) {
    if (n <= 0) return -1;
    double rv = (double)coin31 / 2147483647.0;
    if (rv < eps) return (int)(idx31 % (uint32_t)n);
    double best = -INFINITY;
    int bi = 0;
    for (int i = 0; i < n; ++i) {
        double v = q[XQO_TO(actions[i])];
        if (v > best) { best = v; bi = i; }
    }
    return bi;
}

/* TD target inlined in ChessAI::train, src/chessai.cpp:121-128 (online network; dead DQN::train
 * src/dqn.cpp:157-172 has the same arithmetic with the target network's q_next) */
void xqo_td_target(const double* q_s, const double* q_next, int n_out, int a_to, double reward, int done, double gamma,
                   double* target) {
    memcpy(target, q_s, sizeof(double) * (size_t)n_out);
    if (done) { target[a_to] = reward; return; }
    double m = q_next[0];
    for (int i = 1; i < n_out; ++i) if (q_next[i] > m) m = q_next[i];
    target[a_to] = reward + gamma * m;
}

/* TD gradient of a BATCH: the per-sample steps of ChessAI::train's body (src/chessai.cpp:121-131) / DQN::train (src/dqn.cpp:157-172)
 * at frozen weights, summed over the samples -- getQValues(s) with the online net, getQValues(s') with the bootstrap net (tw, tb; pass
 * w, b for the live loop), xqo_td_target, xqo_nn_grad -- on n_threads host threads (samples split in contiguous chunks, partial sums
 * added in chunk order).  x, x2: [n][layers[0]] states; to / reward / done: [n].  gw / gb: full-size sums; *loss = sum of 1/2 (q[to] - t)^2. */
typedef struct {
    const int* layers; int n_layers; const double *w, *b, *tw, *tb, *x, *x2; const int32_t *to, *reward; const uint8_t* done;
    double gamma; int corrected; long first, count; double *gw, *gb; double loss;
} xqo_td_job;
static void* td_batch_thread(void* p) {
    xqo_td_job* j = (xqo_td_job*)p;
    const int nin = j->layers[0], nout = j->layers[j->n_layers - 1];
    size_t nw = 0, nb = 0;
    for (int l = 0; l + 1 < j->n_layers; ++l) { nw += (size_t)j->layers[l] * j->layers[l + 1]; nb += (size_t)j->layers[l + 1]; }
    double *qs = (double*)malloc(sizeof(double) * nout), *qn = (double*)malloc(sizeof(double) * nout), *tg = (double*)malloc(sizeof(double) * nout);
    double *g1 = (double*)malloc(sizeof(double) * nw), *g2 = (double*)malloc(sizeof(double) * nb);
    for (long i = j->first; i < j->first + j->count; ++i) {
        xqo_nn_forward(j->layers, j->n_layers, j->w, j->b, j->x + (size_t)i * nin, qs);
        xqo_nn_forward(j->layers, j->n_layers, j->tw, j->tb, j->x2 + (size_t)i * nin, qn);
        xqo_td_target(qs, qn, nout, j->to[i], (double)j->reward[i], j->done[i], j->gamma, tg);
        xqo_nn_grad(j->layers, j->n_layers, j->w, j->b, j->x + (size_t)i * nin, tg, j->corrected, g1, g2);
        for (size_t k = 0; k < nw; ++k) j->gw[k] += g1[k];
        for (size_t k = 0; k < nb; ++k) j->gb[k] += g2[k];
        j->loss += 0.5 * (qs[j->to[i]] - tg[j->to[i]]) * (qs[j->to[i]] - tg[j->to[i]]);
    }
    free(qs); free(qn); free(tg); free(g1); free(g2);
    return NULL;
}
void xqo_td_batch_grad(const int* layers, int n_layers, const double* w, const double* b, const double* tw, const double* tb, const double* x,
                       const double* x2, const int32_t* to, const int32_t* reward, const uint8_t* done, double gamma, int corrected, long n,
                       int n_threads, double* gw, double* gb, double* loss) {
    size_t nw = 0, nb = 0;
    for (int l = 0; l + 1 < n_layers; ++l) { nw += (size_t)layers[l] * layers[l + 1]; nb += (size_t)layers[l + 1]; }
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    if ((long)n_threads > n) n_threads = (int)(n > 0 ? n : 1);
    pthread_t th[256];
    xqo_td_job* jobs = (xqo_td_job*)calloc((size_t)n_threads, sizeof(xqo_td_job));
    for (int t = 0; t < n_threads; ++t) {
        const long first = n * t / n_threads, last = n * (t + 1) / n_threads;
        jobs[t] = (xqo_td_job){layers, n_layers, w, b, tw, tb, x, x2, to, reward, done, gamma, corrected, first, last - first,
                               (double*)calloc(nw, sizeof(double)), (double*)calloc(nb, sizeof(double)), 0.0};
        pthread_create(&th[t], NULL, td_batch_thread, &jobs[t]);
    }
    memset(gw, 0, sizeof(double) * nw); memset(gb, 0, sizeof(double) * nb);
    *loss = 0.0;
    for (int t = 0; t < n_threads; ++t) {
        pthread_join(th[t], NULL);
        for (size_t k = 0; k < nw; ++k) gw[k] += jobs[t].gw[k];
        for (size_t k = 0; k < nb; ++k) gb[k] += jobs[t].gb[k];
        *loss += jobs[t].loss;
        free(jobs[t].gw); free(jobs[t].gb);
    }
    free(jobs);
}
