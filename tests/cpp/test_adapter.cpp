// Drives the C++ adapter classes (cn_chess_ai_b200/adapter/xq_adapter.hpp) the way MainWindow / Worker drive the
// reference classes; prints a transcript that tests/test_adapter_gpu.py compares with the reference build.
#include <string>
#include <cstdlib>
#include <cstdio>
#include <cmath>
#include "../../cn_chess_ai_b200/adapter/xq_adapter.hpp"

// `test_adapter trainparity|selfplayparity <model in> <episodes> <model out>`: ChessAI::train / startSelfPlay from a given model file; prints one `event g red black` line per
// finished game -- tests/test_adapter_gpu.py runs the reference's own ChessAI::train on the same weights and the same rand() stream
static int train_parity(bool self_play, const char* in, int episodes, const char* out) {
    ChessBoard board;
    ChessAI ai(&board);
    ai.initializeDQN();
    ai.loadModel(in);
    ai.on_game_completed = [](int g, int r, int b) { std::printf("event %d %d %d\n", g, r, b); };
    if (self_play) ai.startSelfPlay(episodes); else ai.train(episodes);
    ai.saveModel(out);
    return 0;
}

int main(int argc, char** argv) {
    if (argc == 5 && (std::string(argv[1]) == "trainparity" || std::string(argv[1]) == "selfplayparity"))
        return train_parity(std::string(argv[1]) == "selfplayparity", argv[2], std::atoi(argv[3]), argv[4]);
    ChessBoard board;
    ChessAI ai(&board);
    auto acts = ai.getAllValidActions(PieceColor::Red);
    std::printf("opening_actions %zu first %d %d\n", acts.size(), acts[0].from, acts[0].to);
    std::printf("valid_2_1_9_1 %d valid_0_0_5_5 %d\n", (int)board.isValidMove(2, 1, 9, 1), (int)board.isValidMove(0, 0, 5, 5));
    ChessPiece cap = board.movePiece(2, 1, 9, 1);
    std::printf("captured %d %d red_score %d eval %d %d %d\n", (int)cap.type, (int)cap.color, board.getRedScore(), ai.evaluateBoard(PieceColor::Red, 1),
                ai.evaluateBoard(PieceColor::Black, 1), ai.evaluateBoard(PieceColor::Red, 30));
    auto moves = board.getValidMoves(9, 0);
    std::printf("moves_9_0 %zu player %d movecount %d over %d\n", moves.size(), (int)board.getCurrentPlayer(), board.getMoveCount(), (int)board.checkGameOver());
    ChessBoard copy = board;
    copy.reset();
    std::printf("copy_movecount %d orig_movecount %d\n", copy.getMoveCount(), board.getMoveCount());
    auto st = ai.getStateRepresentation();
    double ones = 0; for (double v : st) ones += v;
    std::printf("state %zu ones %.0f\n", st.size(), ones);
    DQN net(std::vector<int>{1260, 128, 8100}, 0.001, 0.99, 7);
    auto q = net.getQValues(st);
    std::printf("q_size %zu q_in_range %d\n", q.size(), (int)(q[0] > -1 && q[0] < 1));
    try { net.getQValues(std::vector<double>(5)); std::printf("no_throw\n"); } catch (const std::invalid_argument&) { std::printf("invalid_argument\n"); }
    try { net.selectAction(st, 0.1, std::vector<Action>{}); std::printf("no_throw\n"); } catch (const std::runtime_error&) { std::printf("runtime_error\n"); }
    int games = 0, last_red = -1;
    ai.on_game_completed = [&](int g, int r, int) { games = g; last_red = r; };
    ai.train(1);
    std::printf("trained_games %d red_score_nonneg %d dqn %d\n", games, (int)(last_red >= 0), (int)ai.isDQNInitialized());
    auto mv = ai.getAIMove(board.getCurrentPlayer());
    std::printf("ai_move_valid %d\n", (int)board.isValidMove(mv.first.first, mv.first.second, mv.second.first, mv.second.second));
    {   // getAIMove's three exits (src/chessai.cpp:29-83): a selected move re-validated against getValidMoves; the uniform fallback after 10 failed
        // attempts; the (-1,-1) sentinel without any action.  A colour with no piece on the board has no action; asking for the colour that is NOT to
        // move exercises the generic per-square enumeration of getAllValidActions (the reference never tests the turn)
        ChessBoard b2;
        ChessAI ai2(&b2);
        ai2.initializeDQN();
        int ok = 0;
        for (int k = 0; k < 12; ++k) {
            auto m = ai2.getAIMove(PieceColor::Black);           // Red is to move: Black's moves are still enumerated and valid per isValidMove
            ok += b2.getPieceAt(m.first.first, m.first.second).color == PieceColor::Black && b2.isValidMove(m.first.first, m.first.second, m.second.first, m.second.second);
        }
        std::printf("ai_move_other_colour_valid %d\n", ok);
        auto none = ai2.getAIMove(PieceColor::None);             // no piece has colour None: no action -> sentinel
        std::printf("ai_move_sentinel %d %d %d %d\n", none.first.first, none.first.second, none.second.first, none.second.second);
    }
    {   // NeuralNetwork (include/dqn.h:43-74): public host vectors + offsets, device copy separate from the host copy
        NeuralNetwork nn(std::vector<int>{6, 5, 4}, 99, true);
        std::printf("nn_sizes %zu %zu offsets %zu %zu %zu %zu\n", nn.host_weights.size(), nn.host_biases.size(), nn.weightOffsets[0], nn.weightOffsets[1], nn.biasOffsets[0], nn.biasOffsets[1]);
        bool in_range = true, bias_zero = true;
        for (double w : nn.host_weights) in_range &= w >= -0.05 && w <= 0.05;
        for (double b : nn.host_biases) bias_zero &= b == 0.0;
        std::vector<double> x{1, 0, 0, 1, 0, 1}, t{0.5, -0.5, 0.25, 0.0};
        const auto q0 = nn.forward(x);
        const std::vector<double> w_before = nn.host_weights;
        nn.backpropagate(x, t, 0.1);
        const bool host_untouched = nn.host_weights == w_before;      // the reference never copies back by itself
        const auto q1 = nn.forward(x);
        nn.copyFromDevice();
        const bool host_changed = nn.host_weights != w_before;
        NeuralNetwork other(std::vector<int>{6, 5, 4}, 7, true);
        other.copyWeightsAndBiasesFrom(nn);
        const auto q2 = other.forward(x);
        NeuralNetwork cp(nn);
        const auto q3 = cp.forward(x);
        double d01 = 0, d12 = 0, d13 = 0;
        for (size_t i = 0; i < q0.size(); ++i) { d01 += std::abs(q0[i] - q1[i]); d12 += std::abs(q1[i] - q2[i]); d13 += std::abs(q1[i] - q3[i]); }
        std::printf("nn_init_ok %d %d trained %d host_untouched %d host_changed %d copy_from_equal %d copy_ctor_equal %d\n", (int)in_range, (int)bias_zero, (int)(d01 > 0),
                    (int)host_untouched, (int)host_changed, (int)(d12 == 0), (int)(d13 == 0));
        try { NeuralNetwork bad(std::vector<int>{3}); std::printf("no_throw\n"); } catch (const std::invalid_argument&) { std::printf("invalid_argument\n"); }
        try { nn.forward(std::vector<double>(2)); std::printf("no_throw\n"); } catch (const std::invalid_argument&) { std::printf("invalid_argument\n"); }
    }
    {   // the batched episode driver behind the same slots / signals
        BatchedChessAI many(256, 11, 0, 1 << 15);
        many.batch = 512; many.autosave_games = 0; many.log_path = "game_log.txt";
        int n = 0, last = 0;
        many.on_game_completed = [&](int g, int, int) { ++n; last = g; };
        bool finished = false;
        many.on_training_finished = [&] { finished = true; };
        const xq_train_report rep = many.train(300);
        std::printf("batched_games %d last %d finished %d report %lld updates_positive %d\n", n, last, (int)finished, (long long)rep.games, (int)(rep.updates > 0));
    }
    return 0;
}
