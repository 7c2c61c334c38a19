/*
 * ccx_oracle.h — CPU restatement of the reference's hot path (TEST INFRASTRUCTURE, NOT PRODUCT).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load
 * this library.  The product package (chinesecheckersagent_b200/) never links or calls it.
 *
 * Every function restates one reference function with the reference's own data structures
 * (7x7 uint8 planes, id->(r,c) tables, a 16-entry move deque) and the reference's own control flow
 * (recursive DFS for jump chains, diagonal scans for wins), NOT the bitboard formulation the CUDA
 * kernels use, so that the two implementations are independent.  file:line citations are into
 * /root/reference.
 *
 * Parity pin: tests/test_oracle_golden.py checks this library against fixtures produced by running
 * the unmodified Python reference (tests/golden/gen_golden.py, run in the build container).
 */
#ifndef CCX_ORACLE_H
#define CCX_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_W 7              /* config.py:10 BOARD_WIDTH = BOARD_HEIGHT */
#define ORC_NCHK 6           /* config.py:8  NUM_CHECKERS */
#define ORC_HIST_PLANES 3    /* config.py:11 BOARD_HIST_MOVES */
#define ORC_TOTAL_HIST 16    /* config.py:15 TOTAL_HIST_MOVES */
#define ORC_UNIQUE_DEST 3    /* config.py:16 UNIQUE_DEST_LIMIT */
#define ORC_MAX_DESTS 48     /* generous bound on destinations of one checker */
#define ORC_NACT 294         /* 6*7*7 policy indices, utils.py:164-171 */

typedef struct {
    uint8_t board[ORC_HIST_PLANES][ORC_W][ORC_W];   /* board.py:19  (plane-major here) */
    int8_t  pos[2][ORC_NCHK][2];                    /* board.py:42-46 checkers_pos[player][id] */
    uint8_t hist[ORC_TOTAL_HIST][4];                /* board.py:54 hist_moves, oldest first: fr,fc,tr,tc */
    int32_t nhist;                                  /* len(hist_moves) */
    int32_t plies;                                  /* number of place() calls so far */
} orc_board;

/* --- board.py ------------------------------------------------------------------------------ */
void orc_init(orc_board *b);                                                  /* board.py:10-57  */
void orc_init_cells(orc_board *b, const int8_t p1[6][2], const int8_t p2[6][2]); /* board.py:61-85 */
int  orc_check_win(const orc_board *b);                                       /* board.py:89-111 */
int  orc_valid_checker_moves(orc_board *b, int player, int r, int c, int8_t out[][2]); /* :139-162 */
void orc_get_valid_moves(orc_board *b, int player, int8_t out[6][ORC_MAX_DESTS][2], int32_t n[6]); /* :215-222 */
int  orc_place(orc_board *b, int player, int fr, int fc, int tr, int tc);     /* board.py:226-250 */
int  orc_player_progress(const orc_board *b, int player);                     /* board.py:254-266 */
int  orc_player_forward_distance(const orc_board *b, int player);             /* board.py:270-288 */
/* --- utils.py ------------------------------------------------------------------------------ */
void orc_to_model_input(const orc_board *b, int cur_player, uint8_t out[7][7][7]); /* utils.py:101-160 */
/* --- player.py ----------------------------------------------------------------------------- */
int  orc_greedy_candidates(orc_board *b, int player, int8_t out[][4]);        /* player.py:99-118 */

/* --- packed (SoA word) state <-> orc_board; layout documented in include/ccx.h -------------- */
void orc_unpack(const uint64_t w[8], orc_board *b, int *to_move, int *status);
void orc_pack(const orc_board *b, int to_move, int status, uint64_t w[8]);

/* Batched drivers over plane-major packed state  st[k*n + i], k < 8.  These mirror the ccx_* device
 * entry points one for one and are what the parity tests compare against. */
void orc_movegen_batch(const uint64_t *st, int64_t n, uint64_t *masks /* [6][n] */);
void orc_encode_batch(const uint64_t *st, int64_t n, uint8_t *out /* [n][7][7][7] */);
void orc_greedy_batch(const uint64_t *st, int64_t n, uint64_t *cand_masks /* [6][n] */);
void orc_movelist_batch(const uint64_t *st, int64_t n, int8_t *out /* [n][6][24] */, int8_t *cnt /* [n][6] */);
void orc_apply_batch(uint64_t *st, int64_t n, const uint8_t *from, const uint8_t *to, uint8_t *winner);
void orc_info_batch(const uint64_t *st, int64_t n, int16_t *out /* [n][5] */);
void orc_greedy_list_batch(const uint64_t *st, int64_t n, int16_t *out /* [n][32][2] */, int16_t *cnt);
/* random-legal stepping (selfplay.py:83-104 move choice; Philox keyed by (seed, game id)) */
void orc_step_random(uint64_t *st, int64_t n, int64_t game_id0, uint64_t seed, uint32_t step0,
                     int32_t plies, uint64_t *wins /* [2] += */, uint64_t *trace, int64_t trace_games,
                     int32_t nthreads);
/* greedy-vs-greedy games with Game.start termination (game.py:58-100) */
void orc_play_greedy(uint64_t *st, int64_t n, int64_t game_id0, uint64_t seed, int32_t max_plies,
                     int32_t nthreads);

/* --- MCTS.py (deterministic tie-break = first maximal edge; evaluator = callback) ------------ */
typedef void (*orc_eval_fn)(void *ctx, const uint8_t planes[7][7][7], double p[ORC_NACT], double *v);
/* Runs MCTS.search's simulation loop (MCTS.py:121-125) on one root; returns per-action visit counts
 * and pi (MCTS.py:131-137).  pre_expand=1 reproduces selfplay.make_move (root expanded first,
 * selfplay.py:117); root_noise (may be NULL) has one entry per root edge in edge order
 * (selfplay.py:121-124).  canonical=1 orders each checker's destinations by r*7+c. */
int orc_mcts_search(const uint64_t root[8], int32_t num_itr, double cpuct, double tau, int pre_expand,
                    int canonical, const double *root_noise, orc_eval_fn eval, void *ctx,
                    uint32_t visits[ORC_NACT], double pi[ORC_NACT], int32_t *n_nodes,
                    double *q_out /* [ORC_NACT] or NULL */);
/* convenience: uniform prior 1/294 and v = 0.0 (SURVEY.md §8d cfg 4) */
void orc_mcts_stub_batch(const uint64_t *st, int64_t n, int32_t num_itr, double cpuct, double tau,
                         int pre_expand, uint32_t *visits /* [n][294] */, double *pi /* [n][294] */,
                         int32_t *n_nodes, int32_t nthreads);

/* batched searches over packed roots; orc_mcts_batch_ties adds the reference's epsilon-tie list (MCTS.py:65-72) with the engine's
 * Philox draw in place of random.choice (ccx_mcts_set_tiebreak mode 1) */
void orc_mcts_batch(const uint64_t *st, int64_t n, int32_t num_itr, double cpuct, double tau, int pre_expand,
                    int evaluator, const double *noise, int32_t noise_stride, uint32_t *visits, double *pi,
                    int32_t *n_nodes, double *q, int32_t nthreads);
void orc_mcts_batch_ties(const uint64_t *st, int64_t n, int32_t num_itr, double cpuct, double tau, int pre_expand,
                         int evaluator, const double *noise, int32_t noise_stride, uint32_t *visits, double *pi,
                         int32_t *n_nodes, double *q, int32_t nthreads, uint64_t tie_seed, int64_t tie_uid0);

/* pthread parallel-for used by the batched drivers (nthreads <= 1 runs inline) */
typedef void (*orc_range_fn)(void *ctx, int64_t lo, int64_t hi);
void orc_parallel_for(orc_range_fn fn, void *ctx, int64_t n, int32_t nthreads);

/* Philox4x32-10 (Salmon et al. 2011), exposed for tests */
void orc_philox(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                uint32_t out[4]);

#ifdef __cplusplus
}
#endif
#endif
