// xq_trainer.cu -- the episode driver: batched equivalent of ChessAI::train / ChessAI::startSelfPlay
// (src/chessai.cpp:85-170, :191-266) over the device-resident collector, replay ring and TD update.
//
// The reference plays ONE game at a time and trains on every ply; its observable protocol per finished game is
//   emit gameCompleted(episode + 1, board->getRedScore(), board->getBlackScore())          (:161, :257)
//   ChessAI::onGameCompleted -> one line in game_log.txt                                     (:370-393)
//   saveModel("model_after_%1_games.bin") every 100 games                                    (:165-167)
//   updateTargetNetwork() every 100 plies                                                    (:140, :245)
// Here thousands of games run side by side: a round is [plies_per_round collector plies over all envs ->
// updates_per_round batched TD updates]; the games that finished during the round are reported in the
// deterministic order (ply, env) with consecutive game numbers, each through the same callback / log line /
// autosave cadence.  Host logic only: every device operation goes through the library's own C ABI.
#include <chrono>
#include <string>
#include <vector>

#include "xq_common.cuh"

using namespace xq;

extern "C" int xq_train_run(xq_dqn_t h, xq_env_t env, xq_replay_t r, const xq_train_config* cfg, xq_game_completed_fn cb, void* user,
                            xq_train_report* report) {
    if (!h || !env || !cfg) return fail(XQ_ERR_INVALID, "xq_train_run: null argument");
    if (cfg->n_games <= 0 || cfg->plies_per_round <= 0 || cfg->updates_per_round < 0)
        return fail(XQ_ERR_INVALID, "xq_train_run: n_games and plies_per_round must be > 0, updates_per_round >= 0");
    if (cfg->updates_per_round > 0 && (!r || cfg->batch <= 0)) return fail(XQ_ERR_INVALID, "xq_train_run: training needs a replay buffer and batch > 0");
    int64_t n_envs = 0;
    if (int rc = xq_env_count(env, &n_envs)) return rc;
    const int64_t cap = n_envs * (int64_t)cfg->plies_per_round;      // at most one finished game per env per ply
    if (int rc = xq_env_enable_game_events(env, cap)) return rc;
    std::vector<xq_game_event> ev((size_t)cap);
    FILE* log = nullptr;
    if (cfg->log_path) {
        log = fopen(cfg->log_path, "a");                             // QIODevice::Append (src/chessai.cpp:197-201)
        if (!log) return fail(XQ_ERR_IO, "xq_train_run: cannot open log file %s", cfg->log_path);
    }
    const std::string prefix = cfg->autosave_prefix ? cfg->autosave_prefix : "model_after_";
    xq_train_report rep = {};
    const auto t0 = std::chrono::steady_clock::now();
    int64_t next_sync = cfg->target_sync_plies > 0 ? cfg->target_sync_plies : -1;
    int rc = XQ_OK;
    while (rc == XQ_OK && rep.games < cfg->n_games) {
        if ((rc = xq_selfplay_collect(h, env, r, cfg->plies_per_round, cfg->eps, cfg->train_done))) break;
        rep.plies += cfg->plies_per_round;
        rep.transitions += n_envs * (int64_t)cfg->plies_per_round;
        if (cfg->updates_per_round > 0) {
            int64_t size = 0;
            if ((rc = xq_replay_info(r, &size, nullptr, nullptr))) break;
            if (size > 0) {      // counters rep.updates, +1, ...; pipelined over two streams when the bootstrap net is the target net
                if ((rc = xq_dqn_td_update_replay_n(h, r, cfg->batch, cfg->sample_seed, (uint32_t)rep.updates, cfg->updates_per_round, cfg->use_target_net,
                                                    cfg->lr))) break;
                rep.updates += cfg->updates_per_round;
            }
        }
        while (next_sync > 0 && rep.plies >= next_sync) {            // dqn->updateTargetNetwork() every target_sync_plies plies
            if ((rc = xq_dqn_sync_target(h))) break;
            ++rep.target_syncs;
            next_sync += cfg->target_sync_plies;
        }
        if (rc) break;
        int64_t n = 0, dropped = 0;
        if ((rc = xq_env_drain_game_events(env, ev.data(), cap, &n, &dropped))) break;
        rep.events_dropped += dropped;
        for (int64_t i = 0; i < n && rep.games < cfg->n_games; ++i) {
            const xq_game_event& e = ev[(size_t)i];
            ++rep.games;
            if (e.red_score > e.black_score) ++rep.red_wins; else if (e.black_score > e.red_score) ++rep.black_wins;
            if (cb) cb(user, rep.games, e.red_score, e.black_score);
            if (log) {                                               // ChessAI::onGameCompleted, src/chessai.cpp:376-386
                const char* result = e.red_score > e.black_score ? "Red wins!" : (e.black_score > e.red_score ? "Black wins!" : "It's a draw!");
                fprintf(log, "Game %lld completed. Red Score: %d, Black Score: %d. %s\n", (long long)rep.games, e.red_score, e.black_score, result);
                if (rep.games == cfg->n_games) fprintf(log, "AI self-play session completed. Total games: %lld\n\n", (long long)cfg->n_games);
                fflush(log);
            }
            if (cfg->autosave_games > 0 && rep.games % cfg->autosave_games == 0) {     // :165-167
                const std::string path = prefix + std::to_string(rep.games) + "_games.bin";
                if ((rc = xq_dqn_save(h, path.c_str()))) break;
                ++rep.autosaves;
            }
        }
    }
    if (log) fclose(log);
    if (rc == XQ_OK) rc = xq_env_sync(env);
    rep.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (report) *report = rep;
    return rc;
}
