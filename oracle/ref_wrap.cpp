// TEST INFRASTRUCTURE ONLY (oracle).  C-ABI shell around the reference's own
// classes, which are compiled UNMODIFIED from /root/reference/src/{chessboard,
// chessai,dqn}.cpp (see oracle/Makefile) into oracle/_ref/libxq_ref.so.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load that library.  No reference source is copied here.
//
// Access to ChessAI's private helpers (getStateRepresentation, evaluateBoard,
// getAllValidActions; include/chessai.h:42-52) and ChessBoard's private fields
// (for position injection) uses `#define private public` around the reference
// headers in THIS translation unit only; the reference .cpp files are compiled
// with their own, unmodified view of the headers.
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>
#include <cuda_runtime.h>
#include "xq_qtshim.h"
#define private public
#include "chessboard.h"
#include "action.h"
#include "dqn.h"
#include "chessai.h"
#undef private

// ---- Qt signal bodies (moc would generate these); they record events --------
namespace { struct GameEvent { int game, red, black; }; std::vector<GameEvent> g_events; int g_training_finished = 0, g_selfplay_finished = 0; }
void ChessAI::gameCompleted(int g, int r, int b) { g_events.push_back({g, r, b}); }
void ChessAI::trainingFinished() { ++g_training_finished; }
void ChessAI::selfPlayFinished() { ++g_selfplay_finished; }

// ---- injected rand()/srand(): the reference's ε-greedy (src/dqn.cpp:30-33) is
// driven by an externally supplied draw stream.  The library is
// linked -Bsymbolic so the reference objects bind to these definitions.
namespace { std::vector<int> g_rand; size_t g_rand_pos = 0; }
extern "C" int rand(void) {
    if (g_rand_pos < g_rand.size()) return g_rand[g_rand_pos++];
    ++g_rand_pos;
    return 0;
}
extern "C" void srand(unsigned) {}

namespace {
inline int code_of(const ChessPiece& p) {
    if (p.type == PieceType::Empty) return 0;
    return static_cast<int>(p.type) + (p.color == PieceColor::Black ? 7 : 0);
}
inline ChessPiece piece_of(int code) {
    if (code <= 0 || code > 14) return ChessPiece();
    return code <= 7 ? ChessPiece(static_cast<PieceType>(code), PieceColor::Red)
                     : ChessPiece(static_cast<PieceType>(code - 7), PieceColor::Black);
}
struct Env { ChessBoard board; std::unique_ptr<ChessAI> ai; Env() : ai(new ChessAI(&board)) {} };
thread_local std::string g_err;
}

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

void ref_rand_load(const int* vals, long n) { g_rand.assign(vals, vals + n); g_rand_pos = 0; }
long ref_rand_consumed() { return static_cast<long>(g_rand_pos); }

void* ref_env_new() { return new Env(); }
void ref_env_free(void* h) { delete static_cast<Env*>(h); }
void ref_env_reset(void* h) { static_cast<Env*>(h)->board.reset(); }

// codes[90]: 0 empty, 1..7 Red G,A,E,H,R,C,S, 8..14 Black; meta = {moveCount, player, redScore, blackScore}
void ref_env_get(void* h, uint8_t* codes, int* meta) {
    ChessBoard& b = static_cast<Env*>(h)->board;
    for (int r = 0; r < 10; ++r) for (int c = 0; c < 9; ++c) codes[r * 9 + c] = static_cast<uint8_t>(code_of(b.getPieceAt(r, c)));
    meta[0] = b.getMoveCount(); meta[1] = b.getCurrentPlayer() == PieceColor::Red ? 0 : 1;
    meta[2] = b.getRedScore(); meta[3] = b.getBlackScore();
}
void ref_env_set(void* h, const uint8_t* codes, const int* meta) {
    ChessBoard& b = static_cast<Env*>(h)->board;
    for (int i = 0; i < 90; ++i) b.board[i] = piece_of(codes[i]);
    b.moveCount = meta[0]; b.currentPlayer = meta[1] == 0 ? PieceColor::Red : PieceColor::Black;
    b.redScore = meta[2]; b.blackScore = meta[3];
}
int ref_env_is_valid_move(void* h, int fr, int fc, int tr, int tc) { return static_cast<Env*>(h)->board.isValidMove(fr, fc, tr, tc) ? 1 : 0; }
int ref_env_valid_moves(void* h, int row, int col, int* to_sq) {
    auto mv = static_cast<Env*>(h)->board.getValidMoves(row, col);
    int n = 0; for (const auto& m : mv) to_sq[n++] = m.first * 9 + m.second; return n;
}
// ChessAI::getAllValidActions (src/chessai.cpp:347-368); out = from*128+to... stored as from,to pairs
int ref_env_all_actions(void* h, int player, int* from_to) {
    Env* e = static_cast<Env*>(h);
    auto v = e->ai->getAllValidActions(player == 0 ? PieceColor::Red : PieceColor::Black);
    int n = 0; for (const auto& a : v) { from_to[2 * n] = a.from; from_to[2 * n + 1] = a.to; ++n; } return n;
}
// Strict legality stated with the reference's OWN classes (the pin of xqo_all_actions_strict): every action of getAllValidActions is
// played on a copy of the board with ChessBoard::movePiece; it is kept iff no action of the other side's getAllValidActions lands on
// the mover's first General and the two first Generals do not face each other on an empty file.
int ref_env_all_actions_strict(void* h, int player, int* from_to) {
    Env* e = static_cast<Env*>(h);
    const PieceColor me = player == 0 ? PieceColor::Red : PieceColor::Black, other = player == 0 ? PieceColor::Black : PieceColor::Red;
    auto v = e->ai->getAllValidActions(me);
    int n = 0;
    for (const auto& a : v) {
        ChessBoard c = e->board;
        c.movePiece(a.from / 9, a.from % 9, a.to / 9, a.to % 9);
        int g = -1, eg = -1;
        for (int s = 0; s < 90; ++s) {
            const ChessPiece p = c.getPieceAt(s / 9, s % 9);
            if (p.type == PieceType::General && p.color == me && g < 0) g = s;
            if (p.type == PieceType::General && p.color == other && eg < 0) eg = s;
        }
        bool safe = true;
        if (g >= 0) {
            ChessAI ai2(&c);
            for (const auto& r : ai2.getAllValidActions(other)) if (r.to == g) safe = false;
            if (safe && eg >= 0 && g % 9 == eg % 9) {
                int between = 0;
                for (int s = std::min(g, eg) + 9; s < std::max(g, eg); s += 9) between += c.getPieceAt(s / 9, s % 9).type != PieceType::Empty;
                if (between == 0) safe = false;
            }
        }
        if (safe) { from_to[2 * n] = a.from; from_to[2 * n + 1] = a.to; ++n; }
    }
    return n;
}
// ChessBoard::movePiece (src/chessboard.cpp:38-64); returns code of captured piece (0 also for a rejected move)
int ref_env_move(void* h, int from, int to) {
    return code_of(static_cast<Env*>(h)->board.movePiece(from / 9, from % 9, to / 9, to % 9));
}
int ref_env_game_over(void* h) { return static_cast<Env*>(h)->board.checkGameOver() ? 1 : 0; }
int ref_env_winner(void* h) { PieceColor w = static_cast<Env*>(h)->board.getWinner(); return w == PieceColor::Red ? 0 : (w == PieceColor::Black ? 1 : 2); }
int ref_env_evaluate(void* h, int player, int moveCount) {
    return static_cast<Env*>(h)->ai->evaluateBoard(player == 0 ? PieceColor::Red : PieceColor::Black, moveCount);
}
void ref_env_state(void* h, double* out1260) {
    auto s = static_cast<Env*>(h)->ai->getStateRepresentation();
    std::memcpy(out1260, s.data(), s.size() * sizeof(double));
}
int ref_piece_score(int type) { return getPieceScore(static_cast<PieceType>(type)); }

// Random-policy rollout of one env through the reference classes, the loop body
// of ChessAI::train (src/chessai.cpp:96-143) without the network:
// getAllValidActions -> pick draws[i] % n -> movePiece -> evaluateBoard -> checkGameOver,
// reset() on terminal.  trace (optional) gets 6 ints per ply:
// {n_legal, from, to, reward, done, winner}.  Returns plies applied.
long ref_env_rollout_random(void* h, const uint32_t* draws, long n_plies, int* trace, long* games_out) {
    Env* e = static_cast<Env*>(h);
    ChessBoard& b = e->board;
    long games = 0, p = 0;
    for (; p < n_plies; ++p) {
        PieceColor mover = b.getCurrentPlayer();
        auto acts = e->ai->getAllValidActions(mover);
        if (acts.empty()) break;
        const Action a = acts[draws[p] % acts.size()];
        b.movePiece(a.from / 9, a.from % 9, a.to / 9, a.to % 9);
        const int reward = e->ai->evaluateBoard(mover, b.getMoveCount());
        const bool done = b.checkGameOver();
        if (trace) {
            int* t = trace + 6 * p;
            t[0] = static_cast<int>(acts.size()); t[1] = a.from; t[2] = a.to; t[3] = reward; t[4] = done ? 1 : 0;
            PieceColor w = b.getWinner();
            t[5] = done ? (w == PieceColor::Red ? 0 : (w == PieceColor::Black ? 1 : 2)) : 2;
        }
        if (done) { ++games; b.reset(); }
    }
    if (games_out) *games_out = games;
    return p;
}

// Multi-threaded timing leg for bench.py (--impl reference / cpu_baseline):
// n_threads independent reference boards, each n_plies random-policy plies.
double ref_bench_rollout_random(int n_threads, long n_plies, uint64_t seed, long* total_plies) {
    std::vector<long> done(n_threads, 0);
    auto work = [&](int t) {
        Env e; uint64_t x = seed + 0x9E3779B97F4A7C15ull * (t + 1);
        std::vector<uint32_t> dr(4096);
        long left = n_plies;
        while (left > 0) {
            long m = std::min<long>(left, 4096);
            for (long i = 0; i < m; ++i) { x += 0x9E3779B97F4A7C15ull; uint64_t z = x; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31; dr[i] = static_cast<uint32_t>(z >> 33); }
            done[t] += ref_env_rollout_random(&e, dr.data(), m, nullptr, nullptr);
            left -= m;
        }
    };
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) th.emplace_back(work, t);
    for (auto& t : th) t.join();
    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    long tot = 0; for (long d : done) tot += d;
    if (total_plies) *total_plies = tot;
    return s;
}

// CPU legs of BASELINE.md section 4 that include the network, through the reference's own classes (DQN over the CPU NeuralNetwork of
// oracle/nn_cpu.cpp in libxq_ref.so), each thread an independent board + network, for `seconds` of wall time:
//   mode 0 = config (i): random policy + ONE forward per ply (getAllValidActions -> getStateRepresentation -> getQValues -> movePiece ->
//            evaluateBoard -> checkGameOver), any number of threads;
//   mode 1 = config (iii): DQN::selectAction(state, 0.1, valid) -- epsilon-greedy, a forward on 90 % of the plies -- ONE thread only
//            (the injected rand() stream is a process global).
// Returns the wall seconds; *plies_out / *games_out = totals over the threads.
double ref_bench_policy(int mode, int n_threads, double seconds, uint64_t seed, long* plies_out, long* games_out) {
    if (mode == 1) n_threads = 1;
    std::vector<long> plies(n_threads, 0), games(n_threads, 0);
    const auto t0 = std::chrono::steady_clock::now();
    auto work = [&](int t) {
        Env e;
        DQN net(std::vector<int>{90 * 14, 128, 90 * 90});
        uint64_t x = seed + 0x9E3779B97F4A7C15ull * (t + 1);
        auto next = [&]() { x += 0x9E3779B97F4A7C15ull; uint64_t z = x; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31; return static_cast<uint32_t>(z >> 33); };
        ChessBoard& b = e.board;
        while (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() < seconds) {
            const PieceColor mover = b.getCurrentPlayer();
            auto acts = e.ai->getAllValidActions(mover);
            if (acts.empty()) { b.reset(); continue; }
            const std::vector<double> state = e.ai->getStateRepresentation();
            Action a;
            if (mode == 0) { const std::vector<double> q = net.getQValues(state); a = acts[next() % acts.size()]; (void)q; }
            else { g_rand.assign({static_cast<int>(next() & 0x7FFFFFFF), static_cast<int>(next() & 0x7FFFFFFF)}); g_rand_pos = 0; a = net.selectAction(state, 0.1, acts); }
            b.movePiece(a.from / 9, a.from % 9, a.to / 9, a.to % 9);
            (void)e.ai->evaluateBoard(mover, b.getMoveCount());
            ++plies[t];
            if (b.checkGameOver()) { ++games[t]; b.reset(); }
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) th.emplace_back(work, t);
    for (auto& t : th) t.join();
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    long tp = 0, tg = 0;
    for (int t = 0; t < n_threads; ++t) { tp += plies[t]; tg += games[t]; }
    if (plies_out) *plies_out = tp;
    if (games_out) *games_out = tg;
    return s;
}

// ---- DQN (reference src/dqn.cpp verbatim over oracle/nn_cpu.cpp or, in the
// CUDA flavour of this library, over the reference's own src/dqn.cu) ----------
void* ref_dqn_new(const int* layers, int n) {
    try { return new DQN(std::vector<int>(layers, layers + n)); } catch (const std::exception& ex) { g_err = ex.what(); return nullptr; }
}
void ref_dqn_free(void* h) { delete static_cast<DQN*>(h); }
long ref_dqn_num_weights(void* h) { return static_cast<long>(static_cast<DQN*>(h)->qNetwork->host_weights.size()); }
long ref_dqn_num_biases(void* h) { return static_cast<long>(static_cast<DQN*>(h)->qNetwork->host_biases.size()); }
void ref_dqn_set_params(void* h, const double* w, const double* b) {
    DQN* d = static_cast<DQN*>(h);
    std::copy(w, w + d->qNetwork->host_weights.size(), d->qNetwork->host_weights.begin());
    std::copy(b, b + d->qNetwork->host_biases.size(), d->qNetwork->host_biases.begin());
    d->qNetwork->copyToDevice();
    d->updateTargetNetwork();
}
// reads back the trained ("device") parameters, which the reference itself never does
void ref_dqn_get_params(void* h, double* w, double* b) {
    DQN* d = static_cast<DQN*>(h);
    d->qNetwork->copyFromDevice();
    std::copy(d->qNetwork->host_weights.begin(), d->qNetwork->host_weights.end(), w);
    std::copy(d->qNetwork->host_biases.begin(), d->qNetwork->host_biases.end(), b);
}
int ref_dqn_forward(void* h, const double* state, int n_in, double* q, int n_out) {
    try { auto v = static_cast<DQN*>(h)->getQValues(std::vector<double>(state, state + n_in)); std::copy(v.begin(), v.begin() + std::min<size_t>(v.size(), n_out), q); return static_cast<int>(v.size()); }
    catch (const std::exception& ex) { g_err = ex.what(); return -1; }
}
int ref_dqn_backprop(void* h, const double* state, int n_in, const double* target, int n_out, double lr) {
    try { static_cast<DQN*>(h)->backpropagate(std::vector<double>(state, state + n_in), std::vector<double>(target, target + n_out), lr); return 0; }
    catch (const std::exception& ex) { g_err = ex.what(); return -1; }
}
// DQN::selectAction (src/dqn.cpp:24-56) with rand() injected via ref_rand_load
int ref_dqn_select(void* h, const double* state, int n_in, double eps, const int* from_to, int n_actions, int* out_from_to) {
    try {
        std::vector<Action> va(n_actions);
        for (int i = 0; i < n_actions; ++i) va[i] = Action{from_to[2 * i], from_to[2 * i + 1]};
        Action a = static_cast<DQN*>(h)->selectAction(std::vector<double>(state, state + n_in), eps, va);
        out_from_to[0] = a.from; out_from_to[1] = a.to; return 0;
    } catch (const std::exception& ex) { g_err = ex.what(); return -1; }
}
// the never-called target-network TD step, DQN::train (src/dqn.cpp:157-172)
int ref_dqn_train(void* h, const double* s, int n_in, int action, double reward, const double* s2, int done) {
    try { static_cast<DQN*>(h)->train(std::vector<double>(s, s + n_in), action, reward, std::vector<double>(s2, s2 + n_in), done != 0); return 0; }
    catch (const std::exception& ex) { g_err = ex.what(); return -1; }
}
void ref_dqn_update_target(void* h) { static_cast<DQN*>(h)->updateTargetNetwork(); }
int ref_dqn_save(void* h, const char* path) { try { static_cast<DQN*>(h)->saveModel(path); return 0; } catch (const std::exception& ex) { g_err = ex.what(); return -1; } }
int ref_dqn_load(void* h, const char* path) { try { static_cast<DQN*>(h)->loadModel(path); return 0; } catch (const std::exception& ex) { g_err = ex.what(); return -1; } }

// ---- ChessAI::train / startSelfPlay verbatim (src/chessai.cpp:85-170,191-266) ----
// weights (optional) replace the random_device init so the run is reproducible.
int ref_ai_train(void* h, int episodes, const double* w, const double* b, double* w_out, double* b_out) {
    Env* e = static_cast<Env*>(h);
    try {
        g_events.clear();
        e->ai->initializeDQN();
        if (w && b) ref_dqn_set_params(e->ai->dqn.get(), w, b);
        e->ai->train(episodes);
        if (w_out && b_out) ref_dqn_get_params(e->ai->dqn.get(), w_out, b_out);
        return 0;
    } catch (const std::exception& ex) { g_err = ex.what(); return -1; }
}
int ref_ai_selfplay(void* h, int games, const double* w, const double* b, double* w_out, double* b_out) {
    Env* e = static_cast<Env*>(h);
    try {
        g_events.clear();
        e->ai->initializeDQN();
        if (w && b) ref_dqn_set_params(e->ai->dqn.get(), w, b);
        e->ai->startSelfPlay(games);
        if (w_out && b_out) ref_dqn_get_params(e->ai->dqn.get(), w_out, b_out);
        return 0;
    } catch (const std::exception& ex) { g_err = ex.what(); return -1; }
}
// ChessAI::getAIMove (src/chessai.cpp:29-83); out = {fr,fc,tr,tc}
int ref_ai_get_move(void* h, int player, const double* w, const double* b, int* out4) {
    Env* e = static_cast<Env*>(h);
    try {
        e->ai->initializeDQN();
        if (w && b) ref_dqn_set_params(e->ai->dqn.get(), w, b);
        auto m = e->ai->getAIMove(player == 0 ? PieceColor::Red : PieceColor::Black);
        out4[0] = m.first.first; out4[1] = m.first.second; out4[2] = m.second.first; out4[3] = m.second.second; return 0;
    } catch (const std::exception& ex) { g_err = ex.what(); return -1; }
}
void ref_ai_log_game(void* h, int game, int red, int black, int num_games) {
    Env* e = static_cast<Env*>(h); e->ai->numGames = num_games; e->ai->onGameCompleted(game, red, black);
}
int ref_events_count() { return static_cast<int>(g_events.size()); }
void ref_events_get(int* out3n) { for (size_t i = 0; i < g_events.size(); ++i) { out3n[3 * i] = g_events[i].game; out3n[3 * i + 1] = g_events[i].red; out3n[3 * i + 2] = g_events[i].black; } }
int ref_training_finished_count() { return g_training_finished; }

}  // extern "C"
