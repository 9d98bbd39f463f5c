// Drives the C++ adapter classes (cn_chess_ai_b200/adapter/xq_adapter.hpp) the way MainWindow / Worker drive the
// reference classes; prints a transcript that tests/test_adapter_gpu.py compares with the reference build.
#include <string>
#include <cstdlib>
#include <cstdio>
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
