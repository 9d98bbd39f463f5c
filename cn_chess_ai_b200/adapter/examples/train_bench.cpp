// Times the drop-in single-board path: the adapter's ChessAI::train(N) (xq_adapter.hpp: the reference's loop of src/chessai.cpp:85-170, one
// board, one TD step per ply, every device operation through the C ABI's FP64 entry points) -- what a maintainer gets after swapping the
// headers per INTEGRATION.md section 1, next to the reference's own loop (bench.py: dqn.reference_train).
//   g++ -std=c++17 -O2 train_bench.cpp -L../.. -lxq_b200 -Wl,-rpath,../..  &&  ./a.out 3
// prints one JSON line: {"transitions_per_s": ..., "plies": ..., "seconds": ..., "games": ...}
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "../xq_adapter.hpp"

int main(int argc, char** argv) {
    const int games = argc > 1 ? std::atoi(argv[1]) : 3;
    ChessBoard board;
    ChessAI ai(&board);
    long plies = 0;
    ai.on_game_completed = [&](int, int, int) { plies += board.getMoveCount(); };
    ai.train(1);                                   // warm-up game: CUDA context, first allocations
    plies = 0;
    const auto t0 = std::chrono::steady_clock::now();
    ai.train(games);
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::printf("{\"transitions_per_s\": %.3f, \"plies\": %ld, \"seconds\": %.6f, \"games\": %d, \"batch\": 1, "
                "\"kind\": \"C++ adapter ChessAI::train over the C ABI (xq_env_* + xq_dqn_train, FP64 kernels), one board\"}\n",
                plies / s, plies, s, games);
    return 0;
}
