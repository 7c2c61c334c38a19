// ccx_datagen.cu — batched GreedyDataGenerator.generate_play (data_generators.py:14-80): greedy-vs-greedy games
// whose every position is recorded with pi = uniform over GreedyPlayer.decide_move(training=True)'s
// filtered_best_moves (player.py:99-118), for supervised pre-training (train_on_greedy.py:17-42).
//
// Games are deterministic functions of (seed, global game id), so generation is two passes over the same
// kernel: pass 1 plays every game to the end and reports how many records it keeps and who won; the host
// prefix-sums the counts; pass 2 replays the games and writes the records at their offsets — the output is
// exactly sized and contiguous, no per-game cap.
#include "ccx_device.cuh"
#include "ccx_internal.h"

#define DG_THREADS 64

__device__ __forceinline__ Game dg_load(const u64 *__restrict__ st, int64_t n, int64_t i)
{
    u64 occ1 = st[0 * n + i], occ2 = st[1 * n + i], c1 = st[2 * n + i], c2 = st[3 * n + i];
    Game g;
    g.meta = st[4 * n + i] & 0x00FFFFFFFFFFFFFFULL;
    bool p2 = (g.meta >> 48) & 1;
    g.occ_me = p2 ? occ2 : occ1; g.occ_op = p2 ? occ1 : occ2;
    g.cells_me = p2 ? c2 : c1;   g.cells_op = p2 ? c1 : c2;
    return g;
}

// One thread per game.  offsets == nullptr: counting pass (n_records, winner).  Otherwise records are written to
// rec_state[5][M] (absolute-player words 0-4, the layout ccx_encode reads), rec_cand[6][M] (candidate masks) and
// rec_v[M] (+1 / -1 / 0 from the point of view of the side to move, given the winners of the counting pass).
//   random_plies : data_generators.py:31-41 (random_start): that many uniformly random legal plies first, unrecorded
//   drop_first   : data_generators.py:74-75 (randomised boards): the first BOARD_HIST_MOVES records are dropped
//   stuck_plies  : data_generators.py:65-67 replaces the reference's 0.1 s wall-clock limit with a ply count; a stuck
//                  game keeps its first `stuck_keep` (AVERAGE_TOTAL_MOVE) records with reward 0
__global__ void __launch_bounds__(DG_THREADS)
k_greedy_generate(const u64 *__restrict__ st, int64_t n, int64_t gid0, u32 k0, u32 k1, int random_plies, int drop_first,
                  int stuck_plies, int stuck_keep, int32_t *__restrict__ n_records, uint8_t *__restrict__ winner_io,
                  const int64_t *__restrict__ offsets, int64_t M, u64 *__restrict__ rec_state, u64 *__restrict__ rec_cand,
                  int8_t *__restrict__ rec_v, int32_t *__restrict__ rec_game, const uint8_t *__restrict__ jt)
{
    __shared__ __align__(16) uint8_t sT[CCX_JT_BYTES];
    for (int q = threadIdx.x; q < CCX_JT_BYTES / 16; q += blockDim.x) reinterpret_cast<uint4 *>(sT)[q] = reinterpret_cast<const uint4 *>(jt)[q];
    __syncthreads();
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Game g = dg_load(st, n, i);
    const u64 gid = (u64)(gid0 + i);
    const bool write = offsets != nullptr;
    const int64_t base = write ? offsets[i] : 0;
    const int final_winner = write ? (int)winner_io[i] : 0;         // 0 = stuck / draw
    const int keep_n = write ? n_records[i] : 0;
    u32 ply = 0;
    for (int t = 0; t < random_plies; t++, ply++) {                 // selfplay-style random opening, no win check (data_generators.py:38)
        u64 dest[6];
        movegen_rays(g.occ_me | g.occ_op, g.cells_me, dest, sT);
        u32 nonempty = 0;
#pragma unroll
        for (int k = 0; k < 6; k++) nonempty += dest[k] != 0;
        if (!nonempty) break;
        Philox4 r = philox4x32_10(k0, k1, ply, 0u, (u32)gid, (u32)(gid >> 32));
        int from, to;
        int id = pick_random(g, dest, nonempty, r.x, r.y, from, to);
        apply_move(g, id, from, to);
    }
    int recorded = 0, winner = 0;
    bool stuck = false;
    for (;;) {
        u64 dest[6], cand[6];
        movegen_rays(g.occ_me | g.occ_op, g.cells_me, dest, sT);
        int total = greedy_candidates(g, dest, cand);
        if (total == 0 || recorded >= stuck_plies) { stuck = true; break; }      // the reference raises on no moves (player.py:113)
        if (write) {
            // which output slot, if any, does record number `recorded` of this game get?
            const bool was_stuck = final_winner == 0;
            const int slot = was_stuck ? recorded : recorded - drop_first;
            if (slot >= 0 && slot < keep_n) {
                const int64_t r = base + slot;
                const bool p2 = (g.meta >> 48) & 1;
                rec_state[0 * M + r] = p2 ? g.occ_op : g.occ_me; rec_state[1 * M + r] = p2 ? g.occ_me : g.occ_op;
                rec_state[2 * M + r] = p2 ? g.cells_op : g.cells_me; rec_state[3 * M + r] = p2 ? g.cells_me : g.cells_op;
                rec_state[4 * M + r] = g.meta;
#pragma unroll
                for (int k = 0; k < 6; k++) rec_cand[k * M + r] = cand[k];
                const int mover = p2 ? 2 : 1;
                rec_v[r] = was_stuck ? 0 : (mover == final_winner ? 1 : -1);       // utils.py:34-44, 65-71
                if (rec_game) rec_game[r] = (int32_t)i;
            }
        }
        recorded++;
        Philox4 r = philox4x32_10(k0, k1, ply, 1u, (u32)gid, (u32)(gid >> 32));
        int from, to;
        int id = pick_candidate(g, cand, total, r.x, from, to);                   // random.choice(best_moves), data_generators.py:55
        apply_move(g, id, from, to);
        ply++;
        winner = winner_of(g);
        if (winner) break;
    }
    if (!write) {
        int keep = stuck ? (recorded < stuck_keep ? recorded : stuck_keep) : (recorded - drop_first > 0 ? recorded - drop_first : 0);
        n_records[i] = keep;
        winner_io[i] = (uint8_t)(stuck ? 0 : winner);
    }
}

// pi_y[r][idx] = 1 / #candidates on every candidate's policy index (data_generators.py:45-51), zero elsewhere
__global__ void __launch_bounds__(128)
k_cand_to_pi(const u64 *__restrict__ rec_cand, int64_t M, float *__restrict__ pi)
{
    int64_t r = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (r >= M) return;
    u64 cand[6]; int total = 0;
#pragma unroll
    for (int k = 0; k < 6; k++) { cand[k] = rec_cand[k * M + r]; total += __popcll(cand[k]); }
    const float w = total ? 1.0f / (float)total : 0.f;
    for (int a = lane; a < CCX_NUM_ACTIONS; a += 32) {
        int id = a / 49, off = a % 49, cell = (off / 7) * 8 + (off % 7);
        pi[r * CCX_NUM_ACTIONS + a] = ((cand[id] >> cell) & 1ULL) ? w : 0.f;
    }
}

extern "C" {

int ccx_greedy_generate(ccx_handle *h, int64_t n, const uint64_t *state, int64_t game_id0, uint64_t seed, int32_t random_plies,
                        int32_t drop_first, int32_t stuck_plies, int32_t stuck_keep, int32_t *n_records, uint8_t *winner,
                        const int64_t *offsets, int64_t total_records, uint64_t *rec_state, uint64_t *rec_cand, int8_t *rec_v,
                        int32_t *rec_game)
{
    if (!h || n < 0 || random_plies < 0 || drop_first < 0 || stuck_plies < 1 || stuck_keep < 0 || (n && (!state || !n_records || !winner)))
        return CCX_ERR_ARG;
    if (offsets && (total_records < 0 || (total_records && (!rec_state || !rec_cand || !rec_v)))) return CCX_ERR_ARG;
    if (n == 0) return CCX_OK;
    unsigned grid = (unsigned)((n + DG_THREADS - 1) / DG_THREADS);
    k_greedy_generate<<<grid, DG_THREADS, 0, h->stream>>>((const u64 *)state, n, game_id0, (u32)seed, (u32)(seed >> 32), random_plies,
                                                          drop_first, stuck_plies, stuck_keep, n_records, winner, offsets, total_records,
                                                          (u64 *)rec_state, (u64 *)rec_cand, rec_v, rec_game, h->jump_table);
    CCX_LAUNCHED(h);
    return CCX_OK;
}

int ccx_cand_to_pi(ccx_handle *h, int64_t m, const uint64_t *rec_cand, float *pi_y)
{
    if (!h || m < 0 || (m && (!rec_cand || !pi_y))) return CCX_ERR_ARG;
    if (m == 0) return CCX_OK;
    k_cand_to_pi<<<(unsigned)((m + 3) / 4), 128, 0, h->stream>>>((const u64 *)rec_cand, m, pi_y);
    CCX_LAUNCHED(h);
    return CCX_OK;
}

}  // extern "C"
