/*
 * ccx_oracle_mcts.c — CPU restatement of MCTS.py (TEST INFRASTRUCTURE, NOT PRODUCT).
 *
 * Follows the reference's object formulation: every expansion eagerly deep-copies the board into one
 * child node per legal move (MCTS.py:97-109), selection scans the edge list with float64 PUCT
 * (MCTS.py:56-69), backup walks the breadcrumbs (MCTS.py:83-90, 112-118).  The only deliberate
 * harness-level choices: ties are broken towards the FIRST maximal edge (what
 * `random.choice` degenerates to with the FirstChoice shim) and, with canonical=1, each checker's
 * destination list is sorted by r*7+c.  Pinned against the real MCTS.py by tests/golden/mcts_golden.npz.
 */
#include "ccx_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct onode onode;
typedef struct {
    int in_player;            /* Edge.currPlayer = inNode.currPlayer  (MCTS.py:28) */
    int8_t from[2], to[2];
    int idx;                  /* utils.encode_checker_index */
    int64_t N;                /* stats['N'] */
    double W, Q, P;           /* stats['W'], ['Q'], ['P'] */
    onode *out;
} oedge;

struct onode {
    orc_board state;
    int player;               /* Node.currPlayer */
    int n_edges;
    oedge *edges;
};

typedef struct { onode **all; int n, cap; } arena;

static onode *new_node(arena *a, const orc_board *b, int player)
{
    onode *nd = (onode *)malloc(sizeof(onode));
    nd->state = *b; nd->player = player; nd->n_edges = 0; nd->edges = NULL;
    if (a->n == a->cap) { a->cap = a->cap ? a->cap * 2 : 1024; a->all = (onode **)realloc(a->all, sizeof(onode *) * (size_t)a->cap); }
    a->all[a->n++] = nd;
    return nd;
}

static void free_arena(arena *a)
{
    for (int i = 0; i < a->n; i++) { free(a->all[i]->edges); free(a->all[i]); }
    free(a->all);
}

/* tie rule of the engine's mode 1 (include/ccx.h ccx_mcts_set_tiebreak): `random.choice(chosen_edges)` (MCTS.py:72) drawn with
 * Philox4x32-10, key = seed, counter = (simulation index, depth, uid ^ roothash*0x9E3779B9, (uid >> 32) ^ 0x7E1B), index = mulhi(x, len) */
typedef struct { int on; uint64_t seed, uid; uint32_t serial, sim; } tie_rule;

/* MCTS.py:49-76.  tie == NULL or !tie->on: first-maximum tie-break (random.choice -> seq[0]); else the reference's
 * chosen_edges list (strict '>' resets it, fabs(QU - maxQU) < EPSILON appends, :65-69) with the Philox draw above. */
static onode *move_to_leaf(onode *root, double cpuct, oedge **crumbs, int *ncrumbs, const tie_rule *tie)
{
    onode *cur = root;
    *ncrumbs = 0;
    int depth = 0;
    while (cur->n_edges != 0) {
        double maxQU = -INFINITY;
        int64_t N_sum = 0;
        oedge *chosen = NULL;
        oedge *list[128]; int nlist = 0;
        for (int i = 0; i < cur->n_edges; i++) N_sum += cur->edges[i].N;              /* :58-59 */
        for (int i = 0; i < cur->n_edges; i++) {
            oedge *e = &cur->edges[i];
            double U = cpuct * e->P * sqrt((double)N_sum) / (1. + (double)e->N);         /* :62 */
            double QU = e->Q + U;                                                        /* :63 */
            if (QU > maxQU) { maxQU = QU; chosen = e; nlist = 0; list[nlist++] = e; }   /* :65-67 */
            else if (fabs(QU - maxQU) < 1e-5 && nlist < 128) list[nlist++] = e;          /* :68-69, EPSILON config.py:36 */
        }
        if (tie && tie->on && nlist > 1) {                                               /* :72 random.choice(chosen_edges) */
            uint32_t out[4];
            orc_philox((uint32_t)tie->seed, (uint32_t)(tie->seed >> 32), tie->sim, (uint32_t)depth,
                       (uint32_t)tie->uid ^ (tie->serial * 0x9E3779B9u), (uint32_t)(tie->uid >> 32) ^ 0x7E1Bu, out);
            chosen = list[(uint32_t)(((uint64_t)out[0] * (uint64_t)nlist) >> 32)];
        }
        crumbs[(*ncrumbs)++] = chosen;                                                   /* :73 */
        cur = chosen->out;
        depth++;
    }
    return cur;
}

static int cmp_dest(const void *a, const void *b)
{
    const int8_t *x = (const int8_t *)a, *y = (const int8_t *)b;
    return (x[0] * 7 + x[1]) - (y[0] * 7 + y[1]);
}

/* MCTS.py:79-118 */
static void expand_and_backup(arena *a, onode *leaf, oedge **crumbs, int ncrumbs, int canonical,
                              orc_eval_fn eval, void *ctx)
{
    int winner = orc_check_win(&leaf->state);                                            /* :81 */
    if (winner) {
        for (int i = 0; i < ncrumbs; i++) {                                              /* :83-89 */
            oedge *e = crumbs[i];
            int direction = (e->in_player == leaf->player) ? -1 : 1;
            e->N += 1;
            e->W += 1.0 * direction;
            e->Q = e->W / (double)e->N;
        }
        return;
    }
    uint8_t planes[7][7][7];
    double p[ORC_NACT], v = 0.0;
    orc_to_model_input(&leaf->state, leaf->player, planes);                              /* :93 */
    eval(ctx, planes, p, &v);
    int8_t mv[6][ORC_MAX_DESTS][2]; int32_t cnt[6];
    orc_get_valid_moves(&leaf->state, leaf->player, mv, cnt);                            /* :95 */
    int total = 0;
    for (int id = 0; id < 6; id++) total += cnt[id];
    leaf->edges = (oedge *)malloc(sizeof(oedge) * (size_t)(total ? total : 1));
    for (int id = 0; id < 6; id++) {                                                     /* :97 dict order = id order */
        if (canonical) qsort(mv[id], (size_t)cnt[id], 2, cmp_dest);
        int fr = leaf->state.pos[leaf->player - 1][id][0], fc = leaf->state.pos[leaf->player - 1][id][1];
        for (int k = 0; k < cnt[id]; k++) {
            oedge *e = &leaf->edges[leaf->n_edges++];
            int tr = mv[id][k][0], tc = mv[id][k][1];
            orc_board next = leaf->state;                                                /* :104 deepcopy */
            orc_place(&next, leaf->player, fr, fc, tr, tc);                              /* :105 */
            e->in_player = leaf->player;
            e->from[0] = (int8_t)fr; e->from[1] = (int8_t)fc; e->to[0] = (int8_t)tr; e->to[1] = (int8_t)tc;
            e->idx = id * 49 + tr * 7 + tc;                                              /* :101 */
            e->N = 0; e->W = 0.0; e->Q = 0.0; e->P = p[e->idx];                          /* MCTS.py:32-37 */
            e->out = new_node(a, &next, 3 - leaf->player);                               /* :102,107 */
        }
    }
    for (int i = 0; i < ncrumbs; i++) {                                                  /* :112-118 */
        oedge *e = crumbs[i];
        int direction = (e->in_player == leaf->player) ? 1 : -1;
        e->N += 1;
        e->W += v * direction;
        e->Q = e->W / (double)e->N;
    }
}

static int mcts_search_impl(const uint64_t rootw[8], int32_t num_itr, double cpuct, double tau, int pre_expand,
                            int canonical, const double *root_noise, orc_eval_fn eval, void *ctx,
                            uint32_t visits[ORC_NACT], double pi[ORC_NACT], int32_t *n_nodes, double *q_out, tie_rule *tie);

int orc_mcts_search(const uint64_t rootw[8], int32_t num_itr, double cpuct, double tau, int pre_expand,
                    int canonical, const double *root_noise, orc_eval_fn eval, void *ctx,
                    uint32_t visits[ORC_NACT], double pi[ORC_NACT], int32_t *n_nodes, double *q_out)
{
    return mcts_search_impl(rootw, num_itr, cpuct, tau, pre_expand, canonical, root_noise, eval, ctx, visits, pi, n_nodes, q_out, NULL);
}

static int mcts_search_impl(const uint64_t rootw[8], int32_t num_itr, double cpuct, double tau, int pre_expand,
                            int canonical, const double *root_noise, orc_eval_fn eval, void *ctx,
                            uint32_t visits[ORC_NACT], double pi[ORC_NACT], int32_t *n_nodes, double *q_out, tie_rule *tie)
{
    arena a = {0};
    orc_board b; int tm;
    orc_unpack(rootw, &b, &tm, NULL);
    onode *root = new_node(&a, &b, tm + 1);
    oedge **crumbs = (oedge **)malloc(sizeof(oedge *) * (size_t)(num_itr + 2));
    int ncrumbs = 0;
    memset(visits, 0, sizeof(uint32_t) * ORC_NACT);
    memset(pi, 0, sizeof(double) * ORC_NACT);
    if (q_out) memset(q_out, 0, sizeof(double) * ORC_NACT);
    if (pre_expand) {                                                                    /* selfplay.py:117 */
        expand_and_backup(&a, root, crumbs, 0, canonical, eval, ctx);
        if (root_noise)
            for (int i = 0; i < root->n_edges; i++) {                                    /* selfplay.py:122-124 */
                root->edges[i].P *= (1. - 0.25);
                root->edges[i].P += 0.25 * root_noise[i];
            }
    }
    for (int it = 0; it < num_itr; it++) {                                               /* MCTS.py:123-125 */
        if (tie) tie->sim = (uint32_t)(it + (pre_expand ? 1 : 0));                       /* the engine spends selection 0 on the root expansion */
        onode *leaf = move_to_leaf(root, cpuct, crumbs, &ncrumbs, tie);
        expand_and_backup(&a, leaf, crumbs, ncrumbs, canonical, eval, ctx);
    }
    double sum = 0.0;
    for (int i = 0; i < root->n_edges; i++) {                                            /* :131-135 */
        oedge *e = &root->edges[i];
        visits[e->idx] = (uint32_t)e->N;
        pi[e->idx] = pow((double)e->N, 1. / tau);
        if (q_out) q_out[e->idx] = e->Q;
    }
    for (int i = 0; i < ORC_NACT; i++) sum += pi[i];
    if (sum > 0) for (int i = 0; i < ORC_NACT; i++) pi[i] /= sum;                        /* :137 */
    if (n_nodes) *n_nodes = a.n;
    int n_edges = root->n_edges;
    free(crumbs);
    free_arena(&a);
    return n_edges;
}

/* ---- test evaluators (engine-side specification, not reference code) ---------------------------- */

static void eval_uniform(void *ctx, const uint8_t planes[7][7][7], double p[ORC_NACT], double *v)
{
    (void)ctx; (void)planes;
    for (int i = 0; i < ORC_NACT; i++) p[i] = 1 / 294.;
    *v = 0.0;
}

static uint64_t splitmix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* deterministic pseudo-random evaluator keyed by the network input planes: exercises Q-dependent
 * selection bit-exactly.  H = sum_k splitmix64(k+1) * x_k (mod 2^64) over the 343 plane values in
 * (row, col, channel) order; P[i] = philox(H, (i,7,0,0)).x / 2^32; v = philox(H, (294,7,0,0)).x / 2^31 - 1 */
uint64_t orc_planes_hash(const uint8_t planes[7][7][7])
{
    const uint8_t *x = &planes[0][0][0];
    uint64_t h = 0;
    for (int k = 0; k < 343; k++) if (x[k]) h += splitmix64((uint64_t)k + 1) * x[k];
    return h;
}

static void eval_hash(void *ctx, const uint8_t planes[7][7][7], double p[ORC_NACT], double *v)
{
    (void)ctx;
    uint64_t h = orc_planes_hash(planes);
    uint32_t out[4];
    for (int i = 0; i < ORC_NACT; i++) {
        orc_philox((uint32_t)h, (uint32_t)(h >> 32), (uint32_t)i, 7u, 0u, 0u, out);
        p[i] = (double)out[0] / 4294967296.0;
    }
    orc_philox((uint32_t)h, (uint32_t)(h >> 32), 294u, 7u, 0u, 0u, out);
    *v = (double)out[0] / 2147483648.0 - 1.0;
}

typedef struct {
    const uint64_t *st; int64_t n; int32_t num_itr; double cpuct, tau; int pre_expand, evaluator;
    const double *noise; int32_t noise_stride;
    uint32_t *visits; double *pi; int32_t *n_nodes; double *q;
    int tie_on; uint64_t tie_seed; int64_t tie_uid0;
} mcts_job;

static void mcts_range(void *vj, int64_t lo, int64_t hi)
{
    mcts_job *jb = (mcts_job *)vj;
    for (int64_t i = lo; i < hi; i++) {
        uint64_t w[8];
        for (int k = 0; k < 8; k++) w[k] = jb->st[k * jb->n + i];
        uint64_t z = w[0] * 0x9E3779B97F4A7C15ull + w[1] * 0xC2B2AE3D27D4EB4Full + (w[4] & 0x0000FFFFFFFFFFFFull);   /* include/ccx.h: root hash */
        tie_rule tie = { jb->tie_on, jb->tie_seed, (uint64_t)(jb->tie_uid0 + i), (uint32_t)(z >> 32) ^ (uint32_t)z, 0 };
        mcts_search_impl(w, jb->num_itr, jb->cpuct, jb->tau, jb->pre_expand, 1,
                         jb->noise ? jb->noise + i * jb->noise_stride : NULL,
                         jb->evaluator == 0 ? eval_uniform : eval_hash, NULL,
                         jb->visits + i * ORC_NACT, jb->pi + i * ORC_NACT, jb->n_nodes ? jb->n_nodes + i : NULL,
                         jb->q ? jb->q + i * ORC_NACT : NULL, jb->tie_on ? &tie : NULL);
    }
}

/* evaluator: 0 = uniform prior 1/294, v = 0.0 (SURVEY.md §8d cfg 4); 1 = hash evaluator above.
 * noise (may be NULL): [n][noise_stride] Dirichlet samples, one per root edge in edge order. */
void orc_mcts_batch(const uint64_t *st, int64_t n, int32_t num_itr, double cpuct, double tau, int pre_expand,
                    int evaluator, const double *noise, int32_t noise_stride, uint32_t *visits, double *pi,
                    int32_t *n_nodes, double *q, int32_t nthreads)
{
    mcts_job jb = { st, n, num_itr, cpuct, tau, pre_expand, evaluator, noise, noise_stride, visits, pi, n_nodes, q, 0, 0, 0 };
    orc_parallel_for(mcts_range, &jb, n, nthreads);
}

/* the same with the engine's tie rule 1 (ccx_mcts_set_tiebreak(h, 1, seed, uid0)) */
void orc_mcts_batch_ties(const uint64_t *st, int64_t n, int32_t num_itr, double cpuct, double tau, int pre_expand,
                         int evaluator, const double *noise, int32_t noise_stride, uint32_t *visits, double *pi,
                         int32_t *n_nodes, double *q, int32_t nthreads, uint64_t tie_seed, int64_t tie_uid0)
{
    mcts_job jb = { st, n, num_itr, cpuct, tau, pre_expand, evaluator, noise, noise_stride, visits, pi, n_nodes, q,
                    1, tie_seed, tie_uid0 };
    orc_parallel_for(mcts_range, &jb, n, nthreads);
}

void orc_mcts_stub_batch(const uint64_t *st, int64_t n, int32_t num_itr, double cpuct, double tau,
                         int pre_expand, uint32_t *visits, double *pi, int32_t *n_nodes, int32_t nthreads)
{
    orc_mcts_batch(st, n, num_itr, cpuct, tau, pre_expand, 0, NULL, 0, visits, pi, n_nodes, NULL, nthreads);
}
