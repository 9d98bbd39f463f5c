// xq_adapter.hpp -- Qt-free, source-compatible stand-ins for the reference's hot-path classes, implemented
// on the C ABI of include/xq.h (header-only; link libxq_b200.so).
//
//   reference header (under /root/reference/include)      class here (same names, same signatures)
//   chessboard.h:8-31   PieceType / PieceColor / ChessPiece / PieceScore
//   chessboard.h:33-57  ChessBoard        -> one-env view of xq_env_* (N = 1)
//   action.h:4-11       Action
//   dqn.h:43-74         NeuralNetwork     -> the online network of an xq_dqn handle (FP64 path), public host_weights / host_biases / offsets
//   dqn.h:97-110        DQN               -> xq_dqn_* (FP64 path; the batched tensor-core path is xq_dqn_td_update ...)
//   chessai.h:21-37     ChessAI           -> getAIMove / train / startSelfPlay over the batched engine
//
// QVector<QPair<int,int>> becomes std::vector<std::pair<int,int>>, QString becomes std::string; Qt signals become
// std::function callbacks (on_game_completed = gameCompleted(int,int,int), chessai.h:35).  Errors keep the
// reference's exception types: std::runtime_error from the CUDA layer / file I/O / empty action list
// (dqn.h:16-24, dqn.cpp:26-28,79-81,114-116), std::invalid_argument on size mismatches (dqn.cu:17-19,200-202,324-329).
// The rules engine never throws (chessboard.cpp); invalid input yields Empty pieces / false, as in the reference.
#pragma once
#include <cstdint>
#include <functional>
#include <memory>
#include <random>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/xq.h"

enum class PieceType { Empty, General, Advisor, Elephant, Horse, Chariot, Cannon, Soldier };
enum class PieceColor { Red, Black, None };
struct ChessPiece {
    PieceType type;
    PieceColor color;
    ChessPiece() : type(PieceType::Empty), color(PieceColor::None) {}
    ChessPiece(PieceType t, PieceColor c) : type(t), color(c) {}
};
enum class PieceScore { General = 1000, Advisor = 20, Elephant = 20, Horse = 40, Chariot = 90, Cannon = 45, Soldier = 10 };
inline int getPieceScore(PieceType type) {   // chessboard.cpp:443-454
    switch (type) {
        case PieceType::General: return 1000; case PieceType::Advisor: return 20; case PieceType::Elephant: return 20;
        case PieceType::Horse: return 40; case PieceType::Chariot: return 90; case PieceType::Cannon: return 45;
        case PieceType::Soldier: return 10; default: return 0;
    }
}
struct Action {
    int from, to;
    bool operator==(const Action& o) const { return from == o.from && to == o.to; }
};

namespace xq_adapter {
inline void check(int rc) { if (rc != XQ_OK) throw std::runtime_error(std::string("xq: ") + xq_last_error()); }
inline ChessPiece piece_of(int code) {
    if (code <= 0 || code > 14) return ChessPiece();
    return code <= 7 ? ChessPiece(static_cast<PieceType>(code), PieceColor::Red) : ChessPiece(static_cast<PieceType>(code - 7), PieceColor::Black);
}
}  // namespace xq_adapter

// ChessBoard: value-semantic like the reference (copy = new device env holding the same record).
class ChessBoard {
public:
    ChessBoard() { xq_adapter::check(xq_env_create(1, 0, 0, 0, &env_)); pull(); }
    ChessBoard(const ChessBoard& o) { xq_adapter::check(xq_env_create(1, 0, 0, 0, &env_)); rec_ = o.rec_; push(); }
    ChessBoard& operator=(const ChessBoard& o) { if (this != &o) { rec_ = o.rec_; push(); } return *this; }
    ~ChessBoard() { xq_env_destroy(env_); }

    void initializeBoard() { int mc = rec_.move_count, pl = rec_.player, r = rec_.red_score, b = rec_.black_score;   // :8-29 touches squares only
                             xq_adapter::check(xq_env_reset(env_, nullptr)); pull(); rec_.move_count = (uint16_t)mc; rec_.player = (uint8_t)pl; rec_.red_score = r; rec_.black_score = b; push(); }
    void reset() { xq_adapter::check(xq_env_reset(env_, nullptr)); pull(); }                                          // :95-102
    bool isInsideBoard(int row, int col) const { return row >= 0 && row < XQ_ROWS && col >= 0 && col < XQ_COLS; }     // :323-325
    ChessPiece getPieceAt(int row, int col) const {                                                                   // :31-36
        if (!isInsideBoard(row, col)) return ChessPiece();
        const int s = row * 9 + col;
        return xq_adapter::piece_of((rec_.sq[s >> 3] >> ((s & 7) * 4)) & 15);
    }
    bool isValidMove(int fr, int fc, int tr, int tc) const {                                                          // :66-93
        int32_t q[4] = {fr, fc, tr, tc}; uint8_t v = 0;
        xq_adapter::check(xq_env_is_valid_move(env_, q, &v));
        return v != 0;
    }
    ChessPiece movePiece(int fr, int fc, int tr, int tc) {                                                            // :38-64
        if (!isInsideBoard(fr, fc) || !isInsideBoard(tr, tc)) return ChessPiece();
        xq_action a = XQ_ACTION(fr * 9 + fc, tr * 9 + tc); uint8_t cap = 0, valid = 0;
        xq_adapter::check(xq_env_step(env_, &a, nullptr, nullptr, nullptr, &cap, &valid, 0));
        if (valid) pull();
        return xq_adapter::piece_of(cap);
    }
    std::vector<std::pair<int, int>> getValidMoves(int row, int col) const {                                          // :112-147
        uint8_t cnt = 0, to[20];
        xq_adapter::check(xq_env_valid_moves(env_, row, col, &cnt, to));
        std::vector<std::pair<int, int>> out;
        for (int i = 0; i < cnt; ++i) out.emplace_back(to[i] / 9, to[i] % 9);
        return out;
    }
    int getRedScore() const { return rec_.red_score; }
    int getBlackScore() const { return rec_.black_score; }
    int getMoveCount() const { return rec_.move_count; }
    PieceColor getCurrentPlayer() const { return rec_.player == 0 ? PieceColor::Red : PieceColor::Black; }
    bool checkGameOver() const {                                                                                      // :286-309
        if (rec_.move_count >= XQ_MAX_MOVES) return true;
        bool r = false, b = false;
        for (int s = 0; s < 90; ++s) { const int c = (rec_.sq[s >> 3] >> ((s & 7) * 4)) & 15; r |= c == 1; b |= c == 8; }
        return !(r && b);
    }
    PieceColor getWinner() const {                                                                                    // :312-320
        for (int s = 0; s < 90; ++s) { const int c = (rec_.sq[s >> 3] >> ((s & 7) * 4)) & 15; if (c == 1) return PieceColor::Red; if (c == 8) return PieceColor::Black; }
        return PieceColor::None;
    }
    // per-piece predicates (:328-440) are reachable through isValidMove; kept for source compatibility
    bool isValidGeneralMove(int a, int b, int c, int d) const { return typed(PieceType::General, a, b, c, d); }
    bool isValidAdvisorMove(int a, int b, int c, int d) const { return typed(PieceType::Advisor, a, b, c, d); }
    bool isValidElephantMove(int a, int b, int c, int d) const { return typed(PieceType::Elephant, a, b, c, d); }
    bool isValidHorseMove(int a, int b, int c, int d) const { return typed(PieceType::Horse, a, b, c, d); }
    bool isValidChariotMove(int a, int b, int c, int d) const { return typed(PieceType::Chariot, a, b, c, d); }
    bool isValidCannonMove(int a, int b, int c, int d) const { return typed(PieceType::Cannon, a, b, c, d); }
    bool isValidSoldierMove(int a, int b, int c, int d) const { return typed(PieceType::Soldier, a, b, c, d); }

    // batched-engine access for ChessAI
    xq_env_t handle() const { return env_; }
    const xq_env_rec& record() const { return rec_; }
    void refresh() { pull(); }

private:
    bool typed(PieceType t, int fr, int fc, int tr, int tc) const { return getPieceAt(fr, fc).type == t && isValidMove(fr, fc, tr, tc); }
    void pull() { xq_adapter::check(xq_env_get_boards(env_, &rec_, 0, 1)); }
    void push() { xq_adapter::check(xq_env_set_boards(env_, &rec_, 0, 1)); }
    xq_env_t env_ = nullptr;
    xq_env_rec rec_{};
};

// NeuralNetwork (include/dqn.h:43-74, src/dqn.cu): flat FP64 host copies `host_weights` ([layer][out][in] row-major, layers concatenated) and
// `host_biases` with their per-layer offsets are PUBLIC in the reference -- DQN::saveModel reads them directly (src/dqn.cpp:87-95) -- and the device
// copy is a separate object: forward / backpropagate work on the DEVICE parameters and never touch the host vectors (src/dqn.cu:199-260, 323-467),
// copyToDevice / copyFromDevice move between the two (:480-492), copyWeightsAndBiasesFrom takes the other network's HOST copy (:507-515).
// Initialisation: U(-0.05, 0.05) weights from mt19937, zero biases, fill order layer -> out -> in (:96-123); the reference seeds from
// random_device, here a seed may be given (default: random_device, like the reference).
class NeuralNetwork {
public:
    std::vector<double> host_weights, host_biases;
    std::vector<size_t> weightOffsets, biasOffsets;
    std::vector<int> layerSizes;

    NeuralNetwork() = delete;
    explicit NeuralNetwork(const std::vector<int>& layerSizes_, uint64_t seed = 0, bool seeded = false) : layerSizes(layerSizes_), seed_(seed), seeded_(seeded) {
        if (layerSizes.size() < 2) throw std::invalid_argument("NeuralNetwork must have at least two layers (input and output).");      // :17-19
        create();
        initializeHostWeightsAndBiases();
        copyToDevice();
    }
    NeuralNetwork(const NeuralNetwork& o) : host_weights(o.host_weights), host_biases(o.host_biases), weightOffsets(o.weightOffsets), biasOffsets(o.biasOffsets),
                                           layerSizes(o.layerSizes), seed_(o.seed_), seeded_(o.seeded_) { create(); copyToDevice(); }
    NeuralNetwork(NeuralNetwork&& o) noexcept : host_weights(std::move(o.host_weights)), host_biases(std::move(o.host_biases)), weightOffsets(std::move(o.weightOffsets)),
                                                 biasOffsets(std::move(o.biasOffsets)), layerSizes(std::move(o.layerSizes)), seed_(o.seed_), seeded_(o.seeded_), h_(o.h_) { o.h_ = nullptr; }
    ~NeuralNetwork() { if (h_) xq_dqn_destroy(h_); }
    NeuralNetwork& operator=(const NeuralNetwork& o) {
        if (this != &o) {
            if (layerSizes != o.layerSizes) { if (h_) xq_dqn_destroy(h_); h_ = nullptr; layerSizes = o.layerSizes; create(); }
            host_weights = o.host_weights; host_biases = o.host_biases; weightOffsets = o.weightOffsets; biasOffsets = o.biasOffsets;
            copyToDevice();
        }
        return *this;
    }
    NeuralNetwork& operator=(NeuralNetwork&& o) noexcept {
        if (this != &o) {
            if (h_) xq_dqn_destroy(h_);
            host_weights = std::move(o.host_weights); host_biases = std::move(o.host_biases); weightOffsets = std::move(o.weightOffsets);
            biasOffsets = std::move(o.biasOffsets); layerSizes = std::move(o.layerSizes); h_ = o.h_; o.h_ = nullptr;
        }
        return *this;
    }

    std::vector<double> forward(const std::vector<double>& input) {                                                  // :199-260
        if (input.size() != (size_t)layerSizes.front()) throw std::invalid_argument("Input size does not match network input layer size.");
        std::vector<double> q(layerSizes.back());
        xq_adapter::check(xq_dqn_forward(h_, input.data(), 1, q.data()));
        return q;
    }
    void backpropagate(const std::vector<double>& input, const std::vector<double>& target, double learningRate) {   // :323-467
        if (input.size() != (size_t)layerSizes.front()) throw std::invalid_argument("Input size does not match network input layer size.");
        if (target.size() != (size_t)layerSizes.back()) throw std::invalid_argument("Target size does not match network output layer size.");
        xq_adapter::check(xq_dqn_backprop(h_, input.data(), target.data(), 1, learningRate));
    }
    void copyToDevice() { xq_adapter::check(xq_dqn_set_params(h_, host_weights.data(), host_biases.data())); }        // :480-485
    void copyFromDevice() { xq_adapter::check(xq_dqn_get_params(h_, host_weights.data(), host_biases.data())); }      // :487-492
    void initializeHostWeightsAndBiases() {                                                                           // :96-146
        std::mt19937 gen(seeded_ ? (std::mt19937::result_type)seed_ : std::random_device{}());
        std::uniform_real_distribution<> dis(-0.05, 0.05);
        host_weights.clear(); host_biases.clear();
        weightOffsets.assign(layerSizes.size() - 1, 0); biasOffsets.assign(layerSizes.size() - 1, 0);
        for (size_t l = 0; l + 1 < layerSizes.size(); ++l) {
            weightOffsets[l] = host_weights.size(); biasOffsets[l] = host_biases.size();
            for (int o = 0; o < layerSizes[l + 1]; ++o) for (int i = 0; i < layerSizes[l]; ++i) host_weights.push_back(dis(gen));
            for (int o = 0; o < layerSizes[l + 1]; ++o) host_biases.push_back(0.0);
        }
    }
    void copyWeightsAndBiasesFrom(const NeuralNetwork& other) {                                                       // :507-515: the other's HOST copy
        if (other.host_weights.size() != host_weights.size() || other.host_biases.size() != host_biases.size())
            throw std::invalid_argument("copyWeightsAndBiasesFrom: layer sizes differ");
        host_weights = other.host_weights; host_biases = other.host_biases;
        copyToDevice();
    }
    xq_dqn_t handle() const { return h_; }

private:
    void create() {
        std::vector<int32_t> l(layerSizes.begin(), layerSizes.end());
        xq_adapter::check(xq_dqn_create(l.data(), (int)l.size(), 0.001, 0.99, 0, seed_, XQ_DQN_AS_WRITTEN, &h_));
    }
    uint64_t seed_ = 0;
    bool seeded_ = false;
    xq_dqn_t h_ = nullptr;
};

class DQN {
public:
    DQN(const std::vector<int>& layerSizes, double learningRate = 0.001, double gamma = 0.99, uint64_t seed = 0x5eed)
        : layers_(layerSizes), learningRate_(learningRate), gamma_(gamma) {
        if (layerSizes.size() < 2) throw std::invalid_argument("NeuralNetwork must have at least two layers (input and output).");   // dqn.cu:17-19
        std::vector<int32_t> l(layerSizes.begin(), layerSizes.end());
        xq_adapter::check(xq_dqn_create(l.data(), (int)l.size(), learningRate, gamma, 0, seed, XQ_DQN_AS_WRITTEN, &h_));
    }
    virtual ~DQN() { xq_dqn_destroy(h_); }
    DQN(const DQN&) = delete;
    DQN& operator=(const DQN&) = delete;

    // rand() of dqn.cpp:30-33 is replaced by the framework's counter RNG; (coin31, idx31) may be injected
    Action selectAction(const std::vector<double>& state, double epsilon, const std::vector<Action> validActions) {
        const uint64_t x = xq_rng(0xD09Aull, 0, draws_++);
        return selectAction(state, epsilon, validActions, (uint32_t)(x & 0x7FFFFFFFu), (uint32_t)(x >> 33));
    }
    Action selectAction(const std::vector<double>& state, double epsilon, const std::vector<Action>& validActions, uint32_t coin31, uint32_t idx31) {
        if (validActions.empty()) throw std::runtime_error("No valid actions available.");                            // dqn.cpp:26-28
        if (state.size() != (size_t)layers_.front()) throw std::invalid_argument("Input size does not match network input layer size.");
        std::vector<xq_action> a(validActions.size());
        for (size_t i = 0; i < a.size(); ++i) a[i] = XQ_ACTION(validActions[i].from, validActions[i].to);
        int idx = 0;
        xq_adapter::check(xq_dqn_select_action(h_, state.data(), epsilon, a.data(), (int)a.size(), coin31, idx31, &idx));
        return validActions[idx];
    }
    void backpropagate(const std::vector<double>& state, const std::vector<double>& target, double learningRate) {    // dqn.cpp:59-62
        if (state.size() != (size_t)layers_.front()) throw std::invalid_argument("Input size does not match network input layer size.");
        if (target.size() != (size_t)layers_.back()) throw std::invalid_argument("Target size does not match network output layer size.");
        xq_adapter::check(xq_dqn_backprop(h_, state.data(), target.data(), 1, learningRate));
    }
    std::vector<double> getQValues(const std::vector<double>& state) {                                                // dqn.cpp:65-68
        if (state.size() != (size_t)layers_.front()) throw std::invalid_argument("Input size does not match network input layer size.");
        std::vector<double> q(layers_.back());
        xq_adapter::check(xq_dqn_forward(h_, state.data(), 1, q.data()));
        return q;
    }
    void updateTargetNetwork() { xq_adapter::check(xq_dqn_sync_target(h_)); }                                         // dqn.cpp:71-73
    void saveModel(const std::string& filename) { xq_adapter::check(xq_dqn_save(h_, filename.c_str())); }             // dqn.cpp:76-108
    void loadModel(const std::string& filename) { xq_adapter::check(xq_dqn_load(h_, filename.c_str())); }             // dqn.cpp:111-154
    void train(const std::vector<double>& state, int action, double reward, const std::vector<double>& nextState, bool done) {   // dqn.cpp:157-172
        xq_adapter::check(xq_dqn_train(h_, state.data(), action, reward, nextState.data(), done ? 1 : 0, 1, learningRate_));
    }
    xq_dqn_t handle() const { return h_; }

private:
    std::vector<int> layers_;
    double learningRate_, gamma_;
    xq_dqn_t h_ = nullptr;
    uint32_t draws_ = 0;
};

class ChessAI {
public:
    explicit ChessAI(ChessBoard* board) : board(board) {}
    std::function<void(int, int, int)> on_game_completed;   // signal gameCompleted(gameNumber, redScore, blackScore), chessai.h:35
    std::function<void()> on_training_finished, on_self_play_finished;

    void initializeDQN() { if (!dqn) dqn.reset(new DQN(std::vector<int>{90 * 14, 128, 90 * 90})); }                  // chessai.cpp:395-404
    bool isDQNInitialized() const { return dqn != nullptr; }
    void saveModel(const std::string& f) { if (dqn) dqn->saveModel(f); }
    void loadModel(const std::string& f) { if (dqn) dqn->loadModel(f); }
    DQN* network() { return dqn.get(); }

    std::vector<double> getStateRepresentation() {                                                                    // chessai.cpp:268-289
        std::vector<double> s(XQ_STATE_SIZE);
        xq_adapter::check(xq_env_state_onehot(board->handle(), s.data()));
        return s;
    }
    int evaluateBoard(PieceColor color, int moveCount) {                                                              // chessai.cpp:311-345
        int score = 0;
        for (int r = 0; r < 10; ++r) for (int c = 0; c < 9; ++c) {
            const ChessPiece p = board->getPieceAt(r, c);
            if (p.color == color) score += getPieceScore(p.type); else if (p.color != PieceColor::None) score -= getPieceScore(p.type);
        }
        return (10 * score - moveCount) / 10;   // == (int)(score - moveCount*0.1) for every reachable pair
    }
    std::vector<Action> getAllValidActions(PieceColor player) const {                                                 // chessai.cpp:347-368
        std::vector<Action> out;
        if (player == board->getCurrentPlayer()) {
            uint8_t cnt = 0; xq_action a[XQ_MAX_ACTIONS];
            xq_adapter::check(xq_env_legal_moves(board->handle(), &cnt, a));
            for (int i = 0; i < cnt; ++i) out.push_back(Action{XQ_ACTION_FROM(a[i]), XQ_ACTION_TO(a[i])});
        } else {   // the reference enumerates either colour regardless of turn
            for (int r = 0; r < 10; ++r) for (int c = 0; c < 9; ++c)
                if (board->getPieceAt(r, c).color == player)
                    for (const auto& m : board->getValidMoves(r, c)) out.push_back(Action{r * 9 + c, m.first * 9 + m.second});
        }
        return out;
    }
    // chessai.cpp:29-83: up to 10 eps-greedy attempts re-validated against getValidMoves, then a uniform fallback, (-1,-1) sentinel
    std::pair<std::pair<int, int>, std::pair<int, int>> getAIMove(PieceColor color) {
        std::vector<double> state = getStateRepresentation();
        for (int attempt = 0; attempt < 10; ++attempt) {
            std::vector<Action> valid = getAllValidActions(color);
            if (valid.empty()) continue;
            const Action a = dqn->selectAction(state, 0.1, valid);
            const int fr = a.from / 9, fc = a.from % 9, tr = a.to / 9, tc = a.to % 9;
            const auto moves = board->getValidMoves(fr, fc);
            if (board->getPieceAt(fr, fc).color == color && !moves.empty())
                for (const auto& m : moves) if (m.first == tr && m.second == tc) return {{fr, fc}, {tr, tc}};
        }
        std::vector<Action> valid = getAllValidActions(color);
        if (valid.empty()) return {{-1, -1}, {-1, -1}};
        const Action a = valid[xq_rng(0xA1ull, 1, fallback_++) % valid.size()];
        return {{a.from / 9, a.from % 9}, {a.to / 9, a.to % 9}};
    }
    // chessai.cpp:85-170 (train) and :191-266 (startSelfPlay): the single-board online loop, one TD step per ply
    void train(int numEpisodes) { initializeDQN(); run(numEpisodes, true); if (on_training_finished) on_training_finished(); }
    void startSelfPlay(int numGames) { run(numGames, false); if (on_self_play_finished) on_self_play_finished(); }

private:
    void run(int games, bool is_train) {
        for (int g = 0; g < games; ++g) {
            board->reset();
            PieceColor player = PieceColor::Red;
            std::vector<double> state = getStateRepresentation();
            int moveCount = 0;
            while (!board->checkGameOver() && (!is_train || moveCount < 200)) {
                if (!is_train) player = board->getCurrentPlayer();
                std::vector<Action> valid = getAllValidActions(player);
                if (valid.empty()) break;
                const Action a = dqn->selectAction(state, 0.1, valid);
                board->movePiece(a.from / 9, a.from % 9, a.to / 9, a.to % 9);
                moveCount = board->getMoveCount();
                const double reward = evaluateBoard(player, moveCount);
                std::vector<double> next = getStateRepresentation();
                const bool done = board->checkGameOver() || (is_train && moveCount + 1 >= 200);          // :119 vs :227
                xq_adapter::check(xq_dqn_train(dqn->handle(), state.data(), a.to, reward, next.data(), done ? 1 : 0, 0, learningRate));   // :121-131
                state = next;
                if (is_train) player = player == PieceColor::Red ? PieceColor::Black : PieceColor::Red;
                if (moveCount % 100 == 0) dqn->updateTargetNetwork();                                     // :140, :245
            }
            if (on_game_completed) on_game_completed(g + 1, board->getRedScore(), board->getBlackScore());
            if ((g + 1) % 100 == 0) saveModel("model_after_" + std::to_string(g + 1) + "_games.bin");    // :165-167
        }
    }
    ChessBoard* board;
    std::unique_ptr<DQN> dqn;
    double learningRate = 0.001, gamma = 0.99;
    uint32_t fallback_ = 0;
};

// ---- ChessAI for N boards at once: what a Worker thread (mainwindow.h:130-147) calls when it wants throughput --------------
// Same slots / signals as ChessAI (train, startSelfPlay, gameCompleted, trainingFinished, selfPlayFinished, saveModel / loadModel),
// but the games run side by side on the device: the epsilon-greedy collector with tensor-core inference fills a GPU replay ring,
// batched TD updates train from it, and every finished game is reported in (ply, board) order through the same callback, the
// same game_log.txt line (chessai.cpp:370-393) and the same autosave cadence (:165-167) -- xq_train_run (include/xq.h).
class BatchedChessAI {
public:
    explicit BatchedChessAI(int64_t n_boards, uint64_t seed = 0, int device = 0, int64_t replay_capacity = 1 << 20) {
        xq_adapter::check(xq_env_create(n_boards, device, seed, 0, &env_));
        xq_adapter::check(xq_replay_create(replay_capacity, device, &replay_));
    }
    ~BatchedChessAI() { if (replay_) xq_replay_destroy(replay_); if (env_) xq_env_destroy(env_); }
    BatchedChessAI(const BatchedChessAI&) = delete;
    BatchedChessAI& operator=(const BatchedChessAI&) = delete;

    std::function<void(int, int, int)> on_game_completed;   // gameCompleted(gameNumber, redScore, blackScore)
    std::function<void()> on_training_finished, on_self_play_finished;
    std::string log_path;                                   // "game_log.txt" in the reference (chessai.cpp:196); empty = no log
    int plies_per_round = 16, updates_per_round = 4, target_sync_plies = 100, autosave_games = 100;
    int64_t batch = 4096;

    void initializeDQN() { if (!dqn) dqn.reset(new DQN(std::vector<int>{90 * 14, 128, 90 * 90})); }
    bool isDQNInitialized() const { return dqn != nullptr; }
    void saveModel(const std::string& f) { if (dqn) dqn->saveModel(f); }
    void loadModel(const std::string& f) { if (dqn) dqn->loadModel(f); }
    DQN* network() { return dqn.get(); }

    xq_train_report train(int numEpisodes) {                // ChessAI::train, chessai.cpp:85-170
        initializeDQN();
        const xq_train_report r = run(numEpisodes, updates_per_round, 1);
        if (on_training_finished) on_training_finished();
        return r;
    }
    xq_train_report startSelfPlay(int numGames) {           // ChessAI::startSelfPlay, chessai.cpp:191-266 (it trains online too, :229-239)
        initializeDQN();
        const xq_train_report r = run(numGames, updates_per_round, 0);
        if (on_self_play_finished) on_self_play_finished();
        return r;
    }

private:
    static void trampoline(void* user, int64_t g, int32_t r, int32_t b) {
        auto* self = static_cast<BatchedChessAI*>(user);
        if (self->on_game_completed) self->on_game_completed((int)g, r, b);
    }
    xq_train_report run(int games, int updates, int train_done) {
        xq_train_config cfg = {};
        cfg.n_games = games; cfg.plies_per_round = plies_per_round; cfg.updates_per_round = updates; cfg.batch = batch;
        cfg.eps = 0.1; cfg.lr = learningRate; cfg.use_target_net = 1; cfg.target_sync_plies = target_sync_plies; cfg.train_done = train_done;
        cfg.autosave_games = autosave_games; cfg.autosave_prefix = nullptr; cfg.log_path = log_path.empty() ? nullptr : log_path.c_str();
        cfg.sample_seed = 0x5eed;
        xq_train_report rep = {};
        xq_adapter::check(xq_train_run(dqn->handle(), env_, replay_, &cfg, &BatchedChessAI::trampoline, this, &rep));
        return rep;
    }
    xq_env_t env_ = nullptr;
    xq_replay_t replay_ = nullptr;
    std::unique_ptr<DQN> dqn;
    double learningRate = 0.001;
};
