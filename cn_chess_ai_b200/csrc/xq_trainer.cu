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
//
// Multi-GPU (the network handle is connected to its peers, xq_dqn_dist_connect; one process per GPU, every rank makes the same call
// with its own env / replay shard): the collector and the replay ring stay rank-local, every TD update exchanges its gradient inside
// the contraction kernel, and once per round the ranks all-gather their finished games through the peer-mapped exchange buffer
// (xq_dqn_dist_allgather).  Every rank merges them in (ply, GLOBAL env) order -- the same stream on every rank and for every GPU
// count -- so all ranks count the same games and stop after the same round; the log file and the autosaves are written by rank 0.
#include <algorithm>
#include <chrono>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

#include "xq_common.cuh"

using namespace xq;

extern "C" int xq_dqn_dist_info(xq_dqn_t h, int* rank, int* world);
extern "C" int xq_dqn_dist_allgather(xq_dqn_t h, const void* send_host, int64_t bytes, void* recv_host);

namespace {
struct RoundHeader { int64_t n_events, n_envs; uint64_t env_id0; int64_t dropped; };      // 32 B per rank per round
struct GlobalEvent { uint64_t genv; xq_game_event e; };
constexpr int64_t kMsgBytes = 64 << 10;                          // the message of xq_dqn_dist_allgather
constexpr int64_t kEventsFirst = (kMsgBytes - (int64_t)sizeof(RoundHeader)) / (int64_t)sizeof(xq_game_event);      // events next to the header
constexpr int64_t kEventsPerChunk = kMsgBytes / (int64_t)sizeof(xq_game_event);

// the finished games of all ranks of this round, in (ply, global env) order.  ONE all-gather in the common case: the message is the
// round header followed by as many events as fit (2,729); only a round with more finished games on some rank needs further calls.
int gather_round(xq_dqn_t h, int world, const std::vector<xq_game_event>& ev, int64_t n, uint64_t env_id0, int64_t n_envs, int64_t dropped,
                 std::vector<GlobalEvent>* merged, int64_t* envs_total, int64_t* dropped_total, std::vector<uint8_t>* send, std::vector<uint8_t>* recv) {
    send->assign((size_t)kMsgBytes, 0);
    recv->resize((size_t)(kMsgBytes * world));
    const RoundHeader mine{n, n_envs, env_id0, dropped};
    memcpy(send->data(), &mine, sizeof(mine));
    const int64_t m0 = std::min(n, kEventsFirst);
    if (m0 > 0) memcpy(send->data() + sizeof(RoundHeader), ev.data(), (size_t)m0 * sizeof(xq_game_event));
    // fixed message size (the ranks cannot know each other's counts beforehand): 64 KB is ~0.1 us of NVLink bandwidth
    if (int rc = xq_dqn_dist_allgather(h, send->data(), kMsgBytes, recv->data())) return rc;
    std::vector<RoundHeader> hd((size_t)world);
    int64_t max_n = 0;
    *envs_total = 0; *dropped_total = 0;
    merged->clear();
    for (int r = 0; r < world; ++r) {
        memcpy(&hd[(size_t)r], recv->data() + (size_t)r * kMsgBytes, sizeof(RoundHeader));
        max_n = std::max(max_n, hd[(size_t)r].n_events); *envs_total += hd[(size_t)r].n_envs; *dropped_total += hd[(size_t)r].dropped;
        const int64_t mr = std::min(hd[(size_t)r].n_events, kEventsFirst);
        for (int64_t i = 0; i < mr; ++i) {
            xq_game_event e;
            memcpy(&e, recv->data() + (size_t)r * kMsgBytes + sizeof(RoundHeader) + (size_t)i * sizeof(xq_game_event), sizeof(e));
            merged->push_back(GlobalEvent{hd[(size_t)r].env_id0 + e.env, e});
        }
    }
    for (int64_t c0 = kEventsFirst; c0 < max_n; c0 += kEventsPerChunk) {
        const int64_t m = std::max<int64_t>(0, std::min(kEventsPerChunk, n - c0));
        send->assign((size_t)kMsgBytes, 0);
        if (m > 0) memcpy(send->data(), ev.data() + c0, (size_t)m * sizeof(xq_game_event));
        if (int rc = xq_dqn_dist_allgather(h, send->data(), kMsgBytes, recv->data())) return rc;
        for (int r = 0; r < world; ++r) {
            const int64_t mr = std::max<int64_t>(0, std::min(kEventsPerChunk, hd[(size_t)r].n_events - c0));
            for (int64_t i = 0; i < mr; ++i) {
                xq_game_event e;
                memcpy(&e, recv->data() + (size_t)r * kMsgBytes + (size_t)i * sizeof(xq_game_event), sizeof(e));
                merged->push_back(GlobalEvent{hd[(size_t)r].env_id0 + e.env, e});
            }
        }
    }
    // every rank's events arrive sorted by (ply, env) (xq_env_drain_game_events): group them by rank (chunks of one rank may be interleaved with other
    // ranks' chunks), then merge the `world` sorted runs pairwise -- n log(world) comparisons instead of a full sort of thousands of records per round
    auto less = [](const GlobalEvent& a, const GlobalEvent& b) { return a.e.ply != b.e.ply ? a.e.ply < b.e.ply : a.genv < b.genv; };
    if (max_n > kEventsFirst) {
        std::stable_sort(merged->begin(), merged->end(), less);      // rare: some rank finished more than 2,729 games in one round
    } else {
        std::vector<size_t> bound((size_t)world + 1, 0);
        for (int r = 0; r < world; ++r) bound[(size_t)r + 1] = bound[(size_t)r] + (size_t)std::min(hd[(size_t)r].n_events, kEventsFirst);
        for (int width = 1; width < world; width *= 2)
            for (int r = 0; r + width < world; r += 2 * width)
                std::inplace_merge(merged->begin() + (ptrdiff_t)bound[(size_t)r], merged->begin() + (ptrdiff_t)bound[(size_t)(r + width)],
                                   merged->begin() + (ptrdiff_t)bound[(size_t)std::min(r + 2 * width, world)], less);
    }
    return XQ_OK;
}
double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
}  // namespace

extern "C" int xq_train_run(xq_dqn_t h, xq_env_t env, xq_replay_t r, const xq_train_config* cfg, xq_game_completed_fn cb, void* user,
                            xq_train_report* report) {
    if (!h || !env || !cfg) return fail(XQ_ERR_INVALID, "xq_train_run: null argument");
    if (cfg->n_games <= 0 || cfg->plies_per_round <= 0 || cfg->updates_per_round < 0)
        return fail(XQ_ERR_INVALID, "xq_train_run: n_games and plies_per_round must be > 0, updates_per_round >= 0");
    if (cfg->updates_per_round > 0 && (!r || cfg->batch <= 0)) return fail(XQ_ERR_INVALID, "xq_train_run: training needs a replay buffer and batch > 0");
    int64_t n_envs = 0;
    if (int rc = xq_env_count(env, &n_envs)) return rc;
    int rank = 0, world = 1;
    if (int rc = xq_dqn_dist_info(h, &rank, &world)) return rc;
    EnvInfo ei;
    if (int rc = env_info(env, &ei)) return rc;
    const int64_t cap = n_envs * (int64_t)cfg->plies_per_round;      // at most one finished game per env per ply
    if (int rc = xq_env_enable_game_events(env, cap)) return rc;
    std::vector<xq_game_event> ev((size_t)cap);
    std::vector<GlobalEvent> merged;
    std::vector<uint8_t> msg_send, msg_recv;
    const bool profile = getenv("XQ_TRAIN_PROFILE") != nullptr;      // where a round's wall time goes (stderr)
    double t_collect = 0, t_update = 0, t_drain = 0, t_gather = 0, t_host = 0;
    FILE* log = nullptr;
    if (cfg->log_path && rank == 0) {
        log = fopen(cfg->log_path, "a");                             // QIODevice::Append (src/chessai.cpp:197-201)
        if (!log) return fail(XQ_ERR_IO, "xq_train_run: cannot open log file %s", cfg->log_path);
    }
    const std::string prefix = cfg->autosave_prefix ? cfg->autosave_prefix : "model_after_";
    // every rank draws from its own ring: decorrelate the draws of the ranks
    const uint64_t sample_seed = cfg->sample_seed + 0x9E3779B97F4A7C15ull * (uint64_t)rank;
    xq_train_report rep = {};
    const auto t0 = std::chrono::steady_clock::now();
    int64_t next_sync = cfg->target_sync_plies > 0 ? cfg->target_sync_plies : -1;
    int rc = XQ_OK;
    while (rc == XQ_OK && rep.games < cfg->n_games) {
        double tp = now_s();
        if ((rc = xq_selfplay_collect(h, env, r, cfg->plies_per_round, cfg->eps, cfg->train_done))) break;
        t_collect += now_s() - tp; tp = now_s();
        rep.plies += cfg->plies_per_round;
        if (cfg->updates_per_round > 0) {
            int64_t size = 0;
            if ((rc = xq_replay_info(r, &size, nullptr, nullptr))) break;
            if (size > 0) {      // counters rep.updates, +1, ...; pipelined over two streams when the bootstrap net is the target net
                if ((rc = xq_dqn_td_update_replay_n(h, r, cfg->batch, sample_seed, (uint32_t)rep.updates, cfg->updates_per_round, cfg->use_target_net,
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
        t_update += now_s() - tp; tp = now_s();
        int64_t n = 0, dropped = 0, envs_round = n_envs;
        if ((rc = xq_env_drain_game_events(env, ev.data(), cap, &n, &dropped))) break;
        t_drain += now_s() - tp; tp = now_s();
        if (world > 1) {         // the round's games of ALL ranks, merged; also where a timed-out exchange surfaces (sticky status)
            if ((rc = gather_round(h, world, ev, n, ei.env_id0, n_envs, dropped, &merged, &envs_round, &dropped, &msg_send, &msg_recv))) break;
            n = (int64_t)merged.size();
        }
        t_gather += now_s() - tp; tp = now_s();
        rep.transitions += envs_round * (int64_t)cfg->plies_per_round;
        rep.events_dropped += dropped;
        const int64_t games_before = rep.games;
        for (int64_t i = 0; i < n && rep.games < cfg->n_games; ++i) {
            const xq_game_event& e = world > 1 ? merged[(size_t)i].e : ev[(size_t)i];
            ++rep.games;
            if (e.red_score > e.black_score) ++rep.red_wins; else if (e.black_score > e.red_score) ++rep.black_wins;
            if (cb) cb(user, rep.games, e.red_score, e.black_score);
            if (log) {                                               // ChessAI::onGameCompleted, src/chessai.cpp:376-386
                const char* result = e.red_score > e.black_score ? "Red wins!" : (e.black_score > e.red_score ? "Black wins!" : "It's a draw!");
                fprintf(log, "Game %lld completed. Red Score: %d, Black Score: %d. %s\n", (long long)rep.games, e.red_score, e.black_score, result);
                if (rep.games == cfg->n_games) fprintf(log, "AI self-play session completed. Total games: %lld\n\n", (long long)cfg->n_games);
            }
        }
        if (log) fflush(log);
        // saveModel every autosave_games games (:165-167).  Thousands of games finish per round and the weights only change between
        // rounds, so ONE snapshot per round in which a multiple was crossed, named after the last crossed multiple: it holds the weights
        // at the end of that round (the reference's file holds them after exactly N games of a one-game-at-a-time loop).
        if (cfg->autosave_games > 0 && rep.games / cfg->autosave_games > games_before / cfg->autosave_games) {
            if (rank == 0) {
                const std::string path = prefix + std::to_string(rep.games / cfg->autosave_games * cfg->autosave_games) + "_games.bin";
                if ((rc = xq_dqn_save(h, path.c_str()))) break;
            }
            ++rep.autosaves;
        }
        t_host += now_s() - tp;
    }
    if (profile) fprintf(stderr, "xq_train_run rank %d: collect (enqueue) %.3f s, updates (enqueue) %.3f s, drain (sync) %.3f s, gather %.3f s, host %.3f s over %lld rounds\n", rank, t_collect, t_update, t_drain, t_gather, t_host, (long long)(rep.plies / cfg->plies_per_round));
    if (log) fclose(log);
    if (rc == XQ_OK) rc = xq_env_sync(env);
    rep.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (report) *report = rep;
    return rc;
}
