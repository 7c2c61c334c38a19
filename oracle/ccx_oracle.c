/*
 * ccx_oracle.c — CPU restatement of the reference's env hot path (TEST INFRASTRUCTURE, NOT PRODUCT).
 * See ccx_oracle.h for the contract.  Citations are file:line into /root/reference.
 *
 * Deliberately written with the reference's array/recursion formulation (board.py), not with
 * bitboards: it must be an independent check on the CUDA kernels.
 */
#include "ccx_oracle.h"
#include <string.h>
#include <stdlib.h>
#include <pthread.h>

#define P1 1
#define P2 2

/* board.py:33-40 — N, E, SE, S, W, NW */
static const int DIRS[6][2] = { {-1, 0}, {0, 1}, {1, 1}, {1, 0}, {0, -1}, {-1, -1} };

/* board_utils.py:15-16 */
static int is_valid_pos(int i, int j) { return i >= 0 && i < ORC_W && j >= 0 && j < ORC_W; }

/* board_utils.py:3-7 (only the row is used by the hot path) */
static int human_row(int i, int j) { return i - j + ORC_W; }

/* board.py:10-57 */
void orc_init(orc_board *b)
{
    static const uint8_t start[7][7] = {
        {0, 0, 0, 0, 2, 2, 2},
        {0, 0, 0, 0, 0, 2, 2},
        {0, 0, 0, 0, 0, 0, 2},
        {0, 0, 0, 0, 0, 0, 0},
        {1, 0, 0, 0, 0, 0, 0},
        {1, 1, 0, 0, 0, 0, 0},
        {1, 1, 1, 0, 0, 0, 0} };
    /* board.py:42-46 */
    static const int8_t p1[6][2] = { {6, 0}, {5, 0}, {6, 1}, {4, 0}, {5, 1}, {6, 2} };
    static const int8_t p2[6][2] = { {0, 6}, {1, 6}, {0, 5}, {2, 6}, {1, 5}, {0, 4} };
    memset(b, 0, sizeof(*b));
    memcpy(b->board[0], start, sizeof(start));
    memcpy(b->pos[0], p1, sizeof(p1));
    memcpy(b->pos[1], p2, sizeof(p2));
}

/* board.py:61-85 with the 12 chosen cells supplied by the caller (first 6 -> P1 ids 0..5) */
void orc_init_cells(orc_board *b, const int8_t p1[6][2], const int8_t p2[6][2])
{
    memset(b, 0, sizeof(*b));
    for (int id = 0; id < ORC_NCHK; id++) {
        b->board[0][p1[id][0]][p1[id][1]] = P1;
        b->board[0][p2[id][0]][p2[id][1]] = P2;
        b->pos[0][id][0] = p1[id][0]; b->pos[0][id][1] = p1[id][1];
        b->pos[1][id][0] = p2[id][0]; b->pos[1][id][1] = p2[id][1];
    }
}

/* board.py:89-111: diagonals k = 4,5,6 all PLAYER_ONE -> 1; diagonals -k all PLAYER_TWO -> 2 */
int orc_check_win(const orc_board *b)
{
    int one_win = 1, two_win = 1;
    for (int k = ORC_W - 3; k < ORC_W; k++) {
        if (one_win)
            for (int i = 0; i + k < ORC_W; i++)          /* cur_board.diagonal(k): (i, i+k) */
                if (b->board[0][i][i + k] != P1) { one_win = 0; break; }
        if (two_win)
            for (int i = 0; i + k < ORC_W; i++)          /* cur_board.diagonal(-k): (i+k, i) */
                if (b->board[0][i + k][i] != P2) { two_win = 0; break; }
        if (!one_win && !two_win) return 0;
    }
    return one_win ? P1 : P2;
}

/* board.py:166-211 — recursive DFS over mirror ("long") jumps */
static void jump_moves(orc_board *b, int8_t out[][2], int *n, uint8_t check_map[7][7], int cr, int cc)
{
    for (int d = 0; d < 6; d++) {
        int step = 1;
        int ri = DIRS[d][0], ci = DIRS[d][1];
        int row = cr + ri, col = cc + ci;
        int valid = 1;
        for (;;) {                                        /* :178-186 find the first checker */
            if (!is_valid_pos(row, col)) { valid = 0; break; }
            if (b->board[0][row][col] != 0) break;
            step++; row += ri; col += ci;
        }
        if (!valid) continue;
        for (int i = 0; i < step; i++) {                  /* :192-197 mirror cells must be empty */
            row += ri; col += ci;
            if (!is_valid_pos(row, col) || b->board[0][row][col] != 0) { valid = 0; break; }
        }
        if (!valid) continue;
        if (check_map[row][col] == 1) continue;           /* :204 */
        out[*n][0] = (int8_t)row; out[*n][1] = (int8_t)col; (*n)++;   /* :208 */
        check_map[row][col] = 1;
        jump_moves(b, out, n, check_map, row, col);       /* :211 */
    }
}

/* board.py:139-162 */
int orc_valid_checker_moves(orc_board *b, int player, int r, int c, int8_t out[][2])
{
    uint8_t check_map[7][7];
    int n = 0;
    memset(check_map, 0, sizeof(check_map));
    check_map[r][c] = 1;                                  /* :147-148 (origin is result[0], removed at :161) */
    for (int d = 0; d < 6; d++) {                         /* :149-155 single steps */
        int row = r + DIRS[d][0], col = c + DIRS[d][1];
        if (!is_valid_pos(row, col)) continue;
        if (b->board[0][row][col] == 0) {
            out[n][0] = (int8_t)row; out[n][1] = (int8_t)col; n++;
            check_map[row][col] = 1;
        }
    }
    b->board[0][r][c] = 0;                                /* :158 lift the mover */
    jump_moves(b, out, &n, check_map, r, c);
    b->board[0][r][c] = (uint8_t)player;                  /* :160 */
    return n;
}

/* board.py:215-222 — checkers in id order */
void orc_get_valid_moves(orc_board *b, int player, int8_t out[6][ORC_MAX_DESTS][2], int32_t n[6])
{
    for (int id = 0; id < ORC_NCHK; id++)
        n[id] = orc_valid_checker_moves(b, player, b->pos[player - 1][id][0], b->pos[player - 1][id][1], out[id]);
}

/* board.py:226-250 */
int orc_place(orc_board *b, int player, int fr, int fc, int tr, int tc)
{
    uint8_t cur[7][7];
    memcpy(cur, b->board[0], sizeof(cur));
    uint8_t t = cur[fr][fc]; cur[fr][fc] = cur[tr][tc]; cur[tr][tc] = t;      /* :231-232 swap */
    for (int id = 0; id < ORC_NCHK; id++)                                     /* :235-238 */
        if (b->pos[player - 1][id][0] == fr && b->pos[player - 1][id][1] == fc) {
            b->pos[player - 1][id][0] = (int8_t)tr; b->pos[player - 1][id][1] = (int8_t)tc;
            break;
        }
    memcpy(b->board[2], b->board[1], sizeof(cur));                            /* :243 shift history */
    memcpy(b->board[1], b->board[0], sizeof(cur));
    memcpy(b->board[0], cur, sizeof(cur));
    if (b->nhist == ORC_TOTAL_HIST) {                                         /* :246-248 */
        memmove(b->hist[0], b->hist[1], (ORC_TOTAL_HIST - 1) * 4);
        b->nhist--;
    }
    b->hist[b->nhist][0] = (uint8_t)fr; b->hist[b->nhist][1] = (uint8_t)fc;
    b->hist[b->nhist][2] = (uint8_t)tr; b->hist[b->nhist][3] = (uint8_t)tc;
    b->nhist++;
    b->plies++;
    return orc_check_win(b);                                                  /* :250 */
}

/* board.py:254-266 */
int orc_player_progress(const orc_board *b, int player)
{
    int reached = 0;
    for (int k = ORC_W - 3; k < ORC_W; k++)
        for (int i = 0; i + k < ORC_W; i++) {
            uint8_t v = (player == P1) ? b->board[0][i][i + k] : b->board[0][i + k][i];
            if (v == player) reached++;
        }
    return reached;
}

/* board.py:270-288 */
int orc_player_forward_distance(const orc_board *b, int player)
{
    int distance = (player == P1) ? 70 : -14;            /* config.py:13-14 */
    for (int id = 0; id < ORC_NCHK; id++) {
        int row = human_row(b->pos[player - 1][id][0], b->pos[player - 1][id][1]);
        distance += (player == P1) ? -row : row;
    }
    return distance;
}

/* utils.py:101-160 */
void orc_to_model_input(const orc_board *b, int cur_player, uint8_t out[7][7][7])
{
    uint8_t cur_layer[7][7], op_layer[7][7];
    int op_player = P1 + P2 - cur_player;
    memset(out, 0, 7 * 7 * 7);
    for (int i = 0; i < 7; i++)
        for (int j = 0; j < 7; j++) {                     /* :116-121,126 putmask */
            uint8_t v = b->board[0][i][j];
            cur_layer[i][j] = (v != cur_player) ? 0 : v;
            op_layer[i][j] = (v != op_player) ? 0 : v;
        }
    for (int id = 0; id < ORC_NCHK; id++) {               /* :123-128 label with id+1 */
        cur_layer[b->pos[cur_player - 1][id][0]][b->pos[cur_player - 1][id][1]] = (uint8_t)(id + 1);
        op_layer[b->pos[op_player - 1][id][0]][b->pos[op_player - 1][id][1]] = (uint8_t)(id + 1);
    }
    for (int i = 0; i < 7; i++)
        for (int j = 0; j < 7; j++) { out[i][j][0] = cur_layer[i][j]; out[i][j][1] = op_layer[i][j]; }

    int moved_player = op_player;                         /* :134 */
    int hist_index = b->nhist - 1;
    for (int channel = 1; channel < ORC_HIST_PLANES; channel++) {
        int any = 0;                                      /* :137 np.any(board[:,:,channel]) */
        for (int i = 0; i < 7 && !any; i++)
            for (int j = 0; j < 7; j++) if (b->board[channel][i][j]) { any = 1; break; }
        if (!any) break;
        const uint8_t *mv = b->hist[hist_index];          /* :139-141 */
        uint8_t (*layer)[7] = (moved_player == cur_player) ? cur_layer : op_layer;
        uint8_t value = layer[mv[2]][mv[3]];              /* :143-150 undo the move */
        layer[mv[2]][mv[3]] = layer[mv[0]][mv[1]];
        layer[mv[0]][mv[1]] = value;
        hist_index--;
        moved_player = P1 + P2 - moved_player;
        for (int i = 0; i < 7; i++)
            for (int j = 0; j < 7; j++) {
                out[i][j][channel * 2] = cur_layer[i][j];
                out[i][j][channel * 2 + 1] = op_layer[i][j];
            }
    }
    if (cur_player == P2)                                 /* :157-158 */
        for (int i = 0; i < 7; i++) for (int j = 0; j < 7; j++) out[i][j][6] = 1;
}

/* player.py:99-118 (deterministic branch, training=True return value, np indices) */
int orc_greedy_candidates(orc_board *b, int player, int8_t out[][4])
{
    int8_t mv[6][ORC_MAX_DESTS][2]; int32_t n[6];
    int8_t best[6 * ORC_MAX_DESTS][4]; int nbest = 0;
    int max_dist = -1000;
    orc_get_valid_moves(b, player, mv, n);
    for (int id = 0; id < ORC_NCHK; id++) {
        int sr = b->pos[player - 1][id][0], sc = b->pos[player - 1][id][1];
        for (int k = 0; k < n[id]; k++) {
            int dist = human_row(mv[id][k][0], mv[id][k][1]) - human_row(sr, sc);   /* :103 */
            if (player == P1) dist = -dist;                                          /* :104-105 */
            if (dist > max_dist) { max_dist = dist; nbest = 0; }                     /* :106-108 */
            if (dist == max_dist) {
                best[nbest][0] = (int8_t)sr; best[nbest][1] = (int8_t)sc;
                best[nbest][2] = mv[id][k][0]; best[nbest][3] = mv[id][k][1]; nbest++;
            }
        }
    }
    if (nbest == 0) return 0;                             /* reference raises ValueError (max([])) */
    int last_key = -1000;                                 /* :113 first maximiser of +-start_row */
    for (int k = 0; k < nbest; k++) {
        int row = human_row(best[k][0], best[k][1]);
        int key = (player == P1) ? row : -row;
        if (key > last_key) last_key = key;
    }
    int last_row = (player == P1) ? last_key : -last_key;
    int nout = 0;
    for (int k = 0; k < nbest; k++)                       /* :115 */
        if (human_row(best[k][0], best[k][1]) == last_row) { memcpy(out[nout], best[k], 4); nout++; }
    return nout;
}

/* ------------------------------------------------------------------------------------------- */
/* packed state (include/ccx.h "State layout")                                                  */

#define CELL(r, c) ((uint8_t)(8 * (r) + (c)))
#define CR(x) ((x) >> 3)
#define CC(x) ((x) & 7)

void orc_unpack(const uint64_t w[8], orc_board *b, int *to_move, int *status)
{
    int8_t p[2][6][2];
    for (int pl = 0; pl < 2; pl++)
        for (int id = 0; id < 6; id++) {
            uint8_t cell = (uint8_t)(w[2 + pl] >> (8 * id));
            p[pl][id][0] = (int8_t)CR(cell); p[pl][id][1] = (int8_t)CC(cell);
        }
    orc_init_cells(b, p[0], p[1]);
    uint64_t meta = w[4];
    int ply = (int)((meta >> 32) & 0xFFFF);
    int tm = (int)((meta >> 48) & 0xFF);
    if (to_move) *to_move = tm;
    if (status) *status = (int)((meta >> 56) & 0xFF);
    b->plies = ply;
    /* rebuild history planes 1,2 by undoing the last two moves (mover of the last move = 1 - to_move) */
    uint8_t prev[7][7];
    memcpy(prev, b->board[0], sizeof(prev));
    for (int k = 0; k < 2 && k < ply; k++) {
        uint8_t from = (uint8_t)(meta >> (16 * k)), to = (uint8_t)(meta >> (16 * k + 8));
        uint8_t t = prev[CR(to)][CC(to)]; prev[CR(to)][CC(to)] = prev[CR(from)][CC(from)]; prev[CR(from)][CC(from)] = t;
        memcpy(b->board[k + 1], prev, sizeof(prev));
    }
    /* hist_moves: only the destinations of plies older than two are kept in the packed state */
    int nh = ply < ORC_TOTAL_HIST ? ply : ORC_TOTAL_HIST;
    b->nhist = nh;
    for (int k = 0; k < nh; k++) {                        /* k = 0 most recent */
        uint8_t dest = (uint8_t)(w[5 + (k >> 3)] >> (8 * (k & 7)));
        uint8_t *h = b->hist[nh - 1 - k];
        h[0] = h[1] = 0xFF; h[2] = (uint8_t)CR(dest); h[3] = (uint8_t)CC(dest);
        if (k < 2) {   /* the last two moves are authoritative in META (states may come without HIST words) */
            uint8_t from = (uint8_t)(meta >> (16 * k)), to = (uint8_t)(meta >> (16 * k + 8));
            h[0] = (uint8_t)CR(from); h[1] = (uint8_t)CC(from); h[2] = (uint8_t)CR(to); h[3] = (uint8_t)CC(to);
        }
    }
}

void orc_pack(const orc_board *b, int to_move, int status, uint64_t w[8])
{
    uint64_t occ[2] = {0, 0}, cells[2] = {0, 0};
    for (int pl = 0; pl < 2; pl++)
        for (int id = 0; id < 6; id++) {
            uint8_t cell = CELL(b->pos[pl][id][0], b->pos[pl][id][1]);
            occ[pl] |= 1ull << cell;
            cells[pl] |= (uint64_t)cell << (8 * id);
        }
    uint64_t meta = 0, hist[2] = {~0ull, ~0ull};
    for (int k = 0; k < 2; k++) {
        uint64_t from = 0xFF, to = 0xFF;
        if (k < b->nhist) {
            const uint8_t *h = b->hist[b->nhist - 1 - k];
            from = CELL(h[0], h[1]); to = CELL(h[2], h[3]);
        }
        meta |= from << (16 * k) | to << (16 * k + 8);
    }
    for (int k = 0; k < b->nhist; k++) {
        const uint8_t *h = b->hist[b->nhist - 1 - k];
        hist[k >> 3] &= ~(0xFFull << (8 * (k & 7)));
        hist[k >> 3] |= (uint64_t)CELL(h[2], h[3]) << (8 * (k & 7));
    }
    meta |= (uint64_t)(b->plies & 0xFFFF) << 32 | (uint64_t)(to_move & 0xFF) << 48 | (uint64_t)(status & 0xFF) << 56;
    w[0] = occ[0]; w[1] = occ[1]; w[2] = cells[0]; w[3] = cells[1]; w[4] = meta; w[5] = hist[0]; w[6] = hist[1];
    /* w[7] (self-play counters) is owned by the caller */
}

static void load_words(const uint64_t *st, int64_t n, int64_t i, uint64_t w[8])
{
    for (int k = 0; k < 8; k++) w[k] = st[k * n + i];
}
static void store_words(uint64_t *st, int64_t n, int64_t i, const uint64_t w[8])
{
    for (int k = 0; k < 7; k++) st[k * n + i] = w[k];
}

static void masks_of(orc_board *b, int player, uint64_t m[6])
{
    int8_t mv[6][ORC_MAX_DESTS][2]; int32_t cnt[6];
    orc_get_valid_moves(b, player, mv, cnt);
    for (int id = 0; id < 6; id++) {
        m[id] = 0;
        for (int k = 0; k < cnt[id]; k++) m[id] |= 1ull << CELL(mv[id][k][0], mv[id][k][1]);
    }
}

void orc_movegen_batch(const uint64_t *st, int64_t n, uint64_t *masks)
{
    for (int64_t i = 0; i < n; i++) {
        uint64_t w[8], m[6]; orc_board b; int tm;
        load_words(st, n, i, w);
        orc_unpack(w, &b, &tm, NULL);
        masks_of(&b, tm + 1, m);
        for (int id = 0; id < 6; id++) masks[id * n + i] = m[id];
    }
}

void orc_encode_batch(const uint64_t *st, int64_t n, uint8_t *out)
{
    for (int64_t i = 0; i < n; i++) {
        uint64_t w[8]; orc_board b; int tm;
        load_words(st, n, i, w);
        orc_unpack(w, &b, &tm, NULL);
        orc_to_model_input(&b, tm + 1, (uint8_t (*)[7][7])(out + i * 343));
    }
}

void orc_greedy_batch(const uint64_t *st, int64_t n, uint64_t *cand)
{
    for (int64_t i = 0; i < n; i++) {
        uint64_t w[8]; orc_board b; int tm;
        int8_t c[6 * ORC_MAX_DESTS][4];
        load_words(st, n, i, w);
        orc_unpack(w, &b, &tm, NULL);
        int nc = orc_greedy_candidates(&b, tm + 1, c);
        for (int id = 0; id < 6; id++) cand[id * n + i] = 0;
        for (int k = 0; k < nc; k++)
            for (int id = 0; id < 6; id++)
                if (b.pos[tm][id][0] == c[k][0] && b.pos[tm][id][1] == c[k][1])
                    cand[id * n + i] |= 1ull << CELL(c[k][2], c[k][3]);
    }
}


/* reference-ORDER move lists (checker id order, then board.py:149-155 walks, then DFS pre-order) */
void orc_movelist_batch(const uint64_t *st, int64_t n, int8_t *out /* [n][6][24] */, int8_t *cnt /* [n][6] */)
{
    for (int64_t i = 0; i < n; i++) {
        uint64_t w[8]; orc_board b; int tm;
        int8_t mv[6][ORC_MAX_DESTS][2]; int32_t c[6];
        load_words(st, n, i, w);
        orc_unpack(w, &b, &tm, NULL);
        orc_get_valid_moves(&b, tm + 1, mv, c);
        for (int id = 0; id < 6; id++) {
            cnt[i * 6 + id] = (int8_t)c[id];
            for (int k = 0; k < 24; k++)
                out[(i * 6 + id) * 24 + k] = k < c[id] ? (int8_t)CELL(mv[id][k][0], mv[id][k][1]) : (int8_t)-1;
        }
    }
}

/* Board.place on every state (board.py:226-250); from/to are cell indices 8r+c */
void orc_apply_batch(uint64_t *st, int64_t n, const uint8_t *from, const uint8_t *to, uint8_t *winner)
{
    for (int64_t i = 0; i < n; i++) {
        uint64_t w[8]; orc_board b; int tm, status;
        load_words(st, n, i, w);
        orc_unpack(w, &b, &tm, &status);
        int win = orc_place(&b, tm + 1, CR(from[i]), CC(from[i]), CR(to[i]), CC(to[i]));
        winner[i] = (uint8_t)win;
        orc_pack(&b, tm ^ 1, status, w);
        store_words(st, n, i, w);
    }
}

/* check_win, player_progress(1), player_progress(2), forward_distance(1), forward_distance(2) */
void orc_info_batch(const uint64_t *st, int64_t n, int16_t *out /* [n][5] */)
{
    for (int64_t i = 0; i < n; i++) {
        uint64_t w[8]; orc_board b;
        load_words(st, n, i, w);
        orc_unpack(w, &b, NULL, NULL);
        out[i * 5 + 0] = (int16_t)orc_check_win(&b);
        out[i * 5 + 1] = (int16_t)orc_player_progress(&b, P1);
        out[i * 5 + 2] = (int16_t)orc_player_progress(&b, P2);
        out[i * 5 + 3] = (int16_t)orc_player_forward_distance(&b, P1);
        out[i * 5 + 4] = (int16_t)orc_player_forward_distance(&b, P2);
    }
}

/* filtered_best_moves in reference order as (from cell, to cell) pairs */
void orc_greedy_list_batch(const uint64_t *st, int64_t n, int16_t *out /* [n][32][2] */, int16_t *cnt)
{
    for (int64_t i = 0; i < n; i++) {
        uint64_t w[8]; orc_board b; int tm;
        int8_t c[6 * ORC_MAX_DESTS][4];
        load_words(st, n, i, w);
        orc_unpack(w, &b, &tm, NULL);
        int nc = orc_greedy_candidates(&b, tm + 1, c);
        cnt[i] = (int16_t)nc;
        for (int k = 0; k < 32; k++) {
            out[(i * 32 + k) * 2 + 0] = k < nc ? CELL(c[k][0], c[k][1]) : -1;
            out[(i * 32 + k) * 2 + 1] = k < nc ? CELL(c[k][2], c[k][3]) : -1;
        }
    }
}

/* ------------------------------------------------------------------------------------------- */
/* tiny pthread parallel-for (chunks handed out through an atomic cursor)                         */

typedef struct { orc_range_fn fn; void *ctx; int64_t n, chunk; int64_t cursor; } pf_state;

static void *pf_worker(void *vs)
{
    pf_state *s = (pf_state *)vs;
    for (;;) {
        int64_t lo = __atomic_fetch_add(&s->cursor, s->chunk, __ATOMIC_RELAXED);
        if (lo >= s->n) break;
        int64_t hi = lo + s->chunk < s->n ? lo + s->chunk : s->n;
        s->fn(s->ctx, lo, hi);
    }
    return NULL;
}

void orc_parallel_for(orc_range_fn fn, void *ctx, int64_t n, int32_t nthreads)
{
    if (nthreads <= 1 || n <= 1) { fn(ctx, 0, n); return; }
    if (nthreads > 256) nthreads = 256;
    pf_state s = { fn, ctx, n, 0, 0 };
    s.chunk = n / ((int64_t)nthreads * 8); if (s.chunk < 1) s.chunk = 1;
    pthread_t th[256];
    int started = 0;
    for (int t = 0; t < nthreads - 1; t++) if (pthread_create(&th[started], NULL, pf_worker, &s) == 0) started++;
    pf_worker(&s);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
}

/* ------------------------------------------------------------------------------------------- */
/* Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11) — the engine's counter RNG specification    */

void orc_philox(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4])
{
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }

static int nth_set_bit(uint64_t m, int k)
{
    for (int bit = 0; bit < 64; bit++)
        if ((m >> bit) & 1) { if (k == 0) return bit; k--; }
    return -1;
}

/* selfplay.py:83-104 move choice (uniform over checkers that can move, then uniform over that
 * checker's destinations in canonical ascending r*7+c order), then Board.place (:101).
 * A won game is counted and restarted from Board() (SURVEY.md §8d cfg 2). */
typedef struct {
    uint64_t *st; int64_t n, game_id0; uint64_t seed; uint32_t step0; int32_t plies;
    uint64_t *trace; int64_t trace_games; uint64_t w1, w2;
} step_job;

static void step_random_range(void *vj, int64_t lo, int64_t hi)
{
    step_job *jb = (step_job *)vj;
    uint64_t *st = jb->st; int64_t n = jb->n, game_id0 = jb->game_id0; uint64_t seed = jb->seed;
    uint32_t step0 = jb->step0; int32_t plies = jb->plies; uint64_t *trace = jb->trace;
    int64_t trace_games = jb->trace_games;
    uint64_t w1 = 0, w2 = 0;
    for (int64_t i = lo; i < hi; i++) {
        uint64_t w[8]; orc_board b; int tm, status;
        load_words(st, n, i, w);
        orc_unpack(w, &b, &tm, &status);
        uint64_t gid = (uint64_t)(game_id0 + i);
        for (int t = 0; t < plies; t++) {
            uint64_t m[6]; uint32_t rnd[4];
            int player = tm + 1;
            masks_of(&b, player, m);
            int nonempty = 0;
            for (int id = 0; id < 6; id++) nonempty += (m[id] != 0);
            uint64_t tr_before[5];
            if (trace && i < trace_games) { uint64_t pw[8]; orc_pack(&b, tm, 0, pw); memcpy(tr_before, pw, 40); }
            int from = 0xFF, to = 0xFF, id_pick = 0xFF, winner = 0;
            if (nonempty) {
                orc_philox((uint32_t)seed, (uint32_t)(seed >> 32), step0 + (uint32_t)t, 0u, (uint32_t)gid, (uint32_t)(gid >> 32), rnd);
                int j = (int)mulhi32(rnd[0], (uint32_t)nonempty);
                for (int id = 0; id < 6; id++)
                    if (m[id]) { if (j == 0) { id_pick = id; break; } j--; }
                int cnt = __builtin_popcountll(m[id_pick]);
                to = nth_set_bit(m[id_pick], (int)mulhi32(rnd[1], (uint32_t)cnt));
                from = CELL(b.pos[tm][id_pick][0], b.pos[tm][id_pick][1]);
                winner = orc_place(&b, player, CR(from), CC(from), CR(to), CC(to));
                tm ^= 1;
            }
            if (trace && i < trace_games) {
                uint64_t *row = trace + ((int64_t)t * trace_games + i) * 12;
                memcpy(row, tr_before, 40);
                memcpy(row + 5, m, 48);
                row[11] = (uint64_t)from | (uint64_t)to << 8 | (uint64_t)winner << 16 | (uint64_t)id_pick << 24;
            }
            if (winner) {
                if (winner == 1) w1++; else w2++;
                orc_init(&b); tm = 0;
            }
        }
        orc_pack(&b, tm, 0, w);
        store_words(st, n, i, w);
    }
    __atomic_fetch_add(&jb->w1, w1, __ATOMIC_RELAXED);
    __atomic_fetch_add(&jb->w2, w2, __ATOMIC_RELAXED);
}

void orc_step_random(uint64_t *st, int64_t n, int64_t game_id0, uint64_t seed, uint32_t step0,
                     int32_t plies, uint64_t *wins, uint64_t *trace, int64_t trace_games, int32_t nthreads)
{
    step_job jb = { st, n, game_id0, seed, step0, plies, trace, trace_games, 0, 0 };
    orc_parallel_for(step_random_range, &jb, n, nthreads);
    if (wins) { wins[0] += jb.w1; wins[1] += jb.w2; }
}

/* game.py:58-100 with both players GreedyPlayer (player.py:99-121), uniform pick among
 * filtered_best_moves driven by Philox (purpose 1); enforce_move_limit=False unless max_plies>0 caps
 * the loop (a cap is a driver safety net, status stays 0 = running). */
typedef struct { uint64_t *st; int64_t n, game_id0; uint64_t seed; int32_t max_plies; } greedy_job;

static void play_greedy_range(void *vj, int64_t lo, int64_t hi)
{
    greedy_job *jb = (greedy_job *)vj;
    uint64_t *st = jb->st; int64_t n = jb->n, game_id0 = jb->game_id0; uint64_t seed = jb->seed;
    int32_t max_plies = jb->max_plies;
    for (int64_t i = lo; i < hi; i++) {
        uint64_t w[8]; orc_board b; int tm, status;
        load_words(st, n, i, w);
        orc_unpack(w, &b, &tm, &status);
        uint64_t gid = (uint64_t)(game_id0 + i);
        for (int t = 0; status == 0 && t < max_plies; t++) {
            int8_t c[6 * ORC_MAX_DESTS][4]; uint32_t rnd[4];
            int player = tm + 1;
            int nc = orc_greedy_candidates(&b, player, c);
            if (nc == 0) { status = 5; break; }           /* reference would raise */
            /* canonical order for the uniform pick: ascending (checker id, dest r*7+c) */
            int order[6 * ORC_MAX_DESTS]; int key[6 * ORC_MAX_DESTS];
            for (int k = 0; k < nc; k++) {
                int id = 0;
                for (; id < 6; id++) if (b.pos[tm][id][0] == c[k][0] && b.pos[tm][id][1] == c[k][1]) break;
                key[k] = id * 64 + CELL(c[k][2], c[k][3]); order[k] = k;
            }
            for (int a = 1; a < nc; a++) {                /* insertion sort */
                int o = order[a], b2 = a - 1;
                while (b2 >= 0 && key[order[b2]] > key[o]) { order[b2 + 1] = order[b2]; b2--; }
                order[b2 + 1] = o;
            }
            orc_philox((uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)b.plies, 1u, (uint32_t)gid, (uint32_t)(gid >> 32), rnd);
            const int8_t *pick = c[order[mulhi32(rnd[0], (uint32_t)nc)]];
            int winner = orc_place(&b, player, pick[0], pick[1], pick[2], pick[3]);   /* game.py:65 */
            tm ^= 1;
            if (winner) { status = winner; break; }                                   /* :70-71 */
            /* :73-82 — history_dests == the packed dest history (16 most recent) */
            if (b.nhist == ORC_TOTAL_HIST) {
                int uniq = 0; uint8_t seen[8][2];
                for (int k = b.nhist - 1; k >= 0; k -= 2) {
                    int dup = 0;
                    for (int u = 0; u < uniq; u++) if (seen[u][0] == b.hist[k][2] && seen[u][1] == b.hist[k][3]) dup = 1;
                    if (!dup) { seen[uniq][0] = b.hist[k][2]; seen[uniq][1] = b.hist[k][3]; uniq++; }
                }
                if (uniq <= ORC_UNIQUE_DEST) { status = 3; break; }
            }
        }
        orc_pack(&b, tm, status, w);
        store_words(st, n, i, w);
    }
}

void orc_play_greedy(uint64_t *st, int64_t n, int64_t game_id0, uint64_t seed, int32_t max_plies, int32_t nthreads)
{
    greedy_job jb = { st, n, game_id0, seed, max_plies };
    orc_parallel_for(play_greedy_range, &jb, n, nthreads);
}
