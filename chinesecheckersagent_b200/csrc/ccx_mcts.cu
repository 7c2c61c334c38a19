// ccx_mcts.cu — batched MCTS (MCTS.py:13-153) as per-tree device node pools, one warp per tree.
//
// Reference semantics kept bit for bit: one simulation at a time per tree, PUCT in
// IEEE float64 evaluated left to right without FMA contraction (MCTS.py:62-63), strict-'>' running
// maximum so the first maximal edge wins (MCTS.py:65-69 with the first-choice tie-break), un-normalised
// priors (MCTS.py:108), terminal leaves never expanded (MCTS.py:81-90), no virtual loss, no tree reuse.
// What changes is the machinery: instead of deep-copying a Board into one Node object per legal move
// (MCTS.py:104-107) a tree is a bump-allocated edge pool (SoA: N, W, P, child, move) plus a pool of the
// nodes that were actually visited; a child's 40-byte state is materialised lazily on its first visit.
//
// Edge order = checker id ascending, then destination cell ascending ("canonical", BASELINE.json).
#include "ccx_device.cuh"
#include "ccx_internal.h"
#include <new>
#include <cmath>
#include <cstdlib>

#define FULL 0xFFFFFFFFu
#define MCTS_WARPS_PER_BLOCK 2
#define NODE_WORDS 6            // 5 state words + info word

// info word of a node: edge_begin (32) | n_edges (8) | spare (16) | winner (4) | expanded (4).  The copy the descent reads lives
// on the parent's edge (eInfo) for every node but the root.
// (Tried and dropped, r01c: keeping N_sum in the spare bits, maintained by the backup, so that sqrt(N_sum) is ready before
// the edge statistics arrive — self-play +0.8 %, stub search -3 %: the extra read-modify-write per level costs what it saves.)
// (Tried and dropped, r01f: an L2 prefetch of the chosen edge's W during the descent, for the backup that reads it — stub search
// 5.49e8 -> 5.35e8 sims/s, self-play ply 18.37 -> 18.60 ms.)
__device__ __forceinline__ u64 make_info(u32 eb, u32 ne, u32 winner, u32 expanded)
{
    return (u64)eb | ((u64)(ne & 0xFF) << 32) | ((u64)(winner & 0xF) << 56) | ((u64)(expanded & 0xF) << 60);
}
__device__ __forceinline__ int info_ne(u64 info) { return (int)((info >> 32) & 0xFF); }
__device__ __forceinline__ int info_eb(u64 info) { return (int)(u32)info; }
__device__ __forceinline__ int info_winner(u64 info) { return (int)((info >> 56) & 0xF); }

struct ccx_trees {
    int64_t cap_trees = 0;
    int32_t nodes_per_tree = 0, edges_per_tree = 0, path_max = 0;
    u64 *node = nullptr;        // [T][NPT][NODE_WORDS]
    u32 *eN = nullptr;          // [T][EPT]
    double *eW = nullptr, *eP = nullptr;
    double *eQ = nullptr;       // W / N as of the last backup (MCTS.py:89,118 store Q the same way): the descent divides once per edge, not twice
    int32_t *eChild = nullptr;  // node index inside the tree or -1
    u64 *eInfo = nullptr;       // [T][EPT] copy of the child node's info word (valid once eChild >= 0): the descent reads it
                                // together with eChild, so a tree level costs two dependent memory round trips, not four
    uint16_t *eMove = nullptr;  // checker id << 8 | destination cell
    int32_t *path = nullptr;    // [T][PATH_MAX] edge indices of the current simulation
    int32_t *tree_meta = nullptr;   // [T][8]: n_nodes, n_edges, overflow, path_len, leaf_node, leaf_kind, selections done, root hash
    size_t bytes = 0;
    // PUCT tie rule (ccx_mcts_set_tiebreak): 0 = first maximal edge, 1 = the reference's epsilon-tie list with a Philox draw
    int32_t tie_mode = 0;       // by value: selects the kernel instantiation (and keys the cached round graph)
    u32 *tie_params = nullptr;  // device u32[4]: Philox key lo/hi, uid0 lo/hi — in memory so that a new seed does not invalidate the graph
    const int64_t *tie_uids = nullptr;   // device, caller-owned (ccx_set_slot_ids): uid of tree i when the batch is compacted, else uid0 + i
};

struct TreeView {
    u64 *node; u32 *eN; double *eW; double *eP; double *eQ; int32_t *eChild; u64 *eInfo; uint16_t *eMove; int32_t *path; int32_t *meta;
    int32_t npt, ept, path_max;
    const u32 *tie; const int64_t *tie_uids; u64 tree_index;
};

__device__ __forceinline__ TreeView tree_view(const ccx_trees &t, int64_t tree)
{
    TreeView v;
    v.node = t.node + tree * t.nodes_per_tree * NODE_WORDS;
    v.eN = t.eN + tree * t.edges_per_tree;
    v.eW = t.eW + tree * t.edges_per_tree;
    v.eP = t.eP + tree * t.edges_per_tree;
    v.eQ = t.eQ + tree * t.edges_per_tree;
    v.eChild = t.eChild + tree * t.edges_per_tree;
    v.eInfo = t.eInfo + tree * t.edges_per_tree;
    v.eMove = t.eMove + tree * t.edges_per_tree;
    v.path = t.path + tree * t.path_max;
    v.meta = t.tree_meta + tree * 8;
    v.npt = t.nodes_per_tree; v.ept = t.edges_per_tree; v.path_max = t.path_max;
    v.tie = t.tie_params; v.tie_uids = t.tie_uids; v.tree_index = (u64)tree;
    return v;
}

enum { META_NNODES = 0, META_NEDGES = 1, META_OVERFLOW = 2, META_PATHLEN = 3, META_LEAF = 4, META_LEAFKIND = 5, META_SELECTS = 6,
       META_SERIAL = 7 };
#define TIE_EPSILON 1e-5        // config.py:36 EPSILON
enum { LEAF_EVAL = 0, LEAF_TERMINAL = 1, LEAF_DEAD = 2 };
enum { EVAL_UNIFORM = 0, EVAL_HASH = 1, EVAL_NET = 2 };

// sqrt of small integers, correctly rounded (filled on the host with IEEE sqrt, which is what the device's sqrt(double) returns
// too): np.sqrt(N_sum) of the descent becomes one broadcast constant load instead of a ~150-cycle dependent sequence per level
#define SQRT_TABLE 1024
__constant__ double c_sqrt[SQRT_TABLE];
__device__ __forceinline__ double sqrt_count(u32 n) { return n < SQRT_TABLE ? c_sqrt[n] : sqrt((double)n); }

__device__ __forceinline__ u64 shfl64(u64 v, int src)
{
    u32 lo = __shfl_sync(FULL, (u32)v, src), hi = __shfl_sync(FULL, (u32)(v >> 32), src);
    return (u64)lo | ((u64)hi << 32);
}

__device__ __forceinline__ Game game_of_words(u64 occ1, u64 occ2, u64 c1, u64 c2, u64 meta)
{
    Game g;
    g.meta = meta;
    bool p2 = (meta >> 48) & 1;
    g.occ_me = p2 ? occ2 : occ1; g.occ_op = p2 ? occ1 : occ2;
    g.cells_me = p2 ? c2 : c1;   g.cells_op = p2 ? c1 : c2;
    return g;
}

__device__ __forceinline__ Game load_node_game(const TreeView &tv, int node)
{
    const u64 *w = tv.node + (int64_t)node * NODE_WORDS;
    return game_of_words(w[0], w[1], w[2], w[3], w[4]);
}

__device__ __forceinline__ void store_node(const TreeView &tv, int node, const Game &g, u64 info)
{
    u64 *w = tv.node + (int64_t)node * NODE_WORDS;
    bool p2 = (g.meta >> 48) & 1;
    w[0] = p2 ? g.occ_op : g.occ_me; w[1] = p2 ? g.occ_me : g.occ_op;
    w[2] = p2 ? g.cells_op : g.cells_me; w[3] = p2 ? g.cells_me : g.cells_op;
    w[4] = g.meta; w[5] = info;
}

// ---- test evaluators (specification shared with the CPU checker; not reference code) ---------------
__host__ __device__ constexpr u64 splitmix64_c(u64 z)
{
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__host__ __device__ constexpr u64 plane6_sum()
{
    u64 s = 0;
    for (int c = 0; c < 49; c++) s += splitmix64_c((u64)(c * 7 + 6) + 1);
    return s;
}

// hash of utils.to_model_input(state, side to move): sum over non-zero plane entries of
// splitmix64(k+1) * value, k = (r*7+c)*7 + channel.  Warp-cooperative: 36 labelled entries + plane 6.
__device__ __forceinline__ u64 warp_sum_u64(u64 v)
{
#pragma unroll
    for (int off = 16; off; off >>= 1) v += shfl64(v, (threadIdx.x & 31) ^ off);
    return v;
}

__device__ __forceinline__ u64 leaf_hash(const Game &g, int lane)
{
    int plies = (int)((g.meta >> 32) & 0xFFFF);
    u64 cur0 = g.cells_me, opp0 = g.cells_op;
    u64 opp1 = undo_in_cells(opp0, (int)(g.meta & 0xFF), (int)((g.meta >> 8) & 0xFF));
    u64 cur2 = undo_in_cells(cur0, (int)((g.meta >> 16) & 0xFF), (int)((g.meta >> 24) & 0xFF));
    u64 h = 0;
    for (int e = lane; e < 36; e += 32) {
        int hs = e / 12, side = (e / 6) & 1, id = e % 6;
        if (hs > plies) continue;
        u64 cells = side ? (hs >= 1 ? opp1 : opp0) : (hs >= 2 ? cur2 : cur0);
        int cell = (int)((cells >> (8 * id)) & 0xFF);
        int k = ((cell >> 3) * 7 + (cell & 7)) * 7 + 2 * hs + side;
        h += splitmix64_c((u64)k + 1) * (u64)(id + 1);
    }
    h = warp_sum_u64(h);
    if ((g.meta >> 48) & 1) h += plane6_sum();                   // utils.py:157-158
    return h;
}

// ---- select (MCTS.moveToLeaf, MCTS.py:49-76) ---------------------------------------------------------
// Returns the leaf node index (materialising it if this is its first visit) and leaves the path in
// tv.path / path_len.  leaf_kind: LEAF_EVAL (needs evaluation + expansion), LEAF_TERMINAL (winner != 0).
// PREFETCH: fetch the child index and child info word of EVERY edge together with the edge statistics, so that a tree level
// is ONE dependent memory round trip (the chosen edge's pair comes out of a shuffle).  Pays off for the deep, narrow trees a
// real net produces (round-based kernels: -1.9 % per self-play ply); costs bandwidth on the wide, shallow trees of the
// uniform-prior stub (persistent kernel: -19 %), which therefore reads the pair after the arg-max.
template <bool PREFETCH, bool TIES>
__device__ __forceinline__ int select_leaf(const TreeView &tv, int lane, double cpuct, int &path_len, int &leaf_kind)
{
    int node = 0, depth = 0;
    u64 info = tv.node[5];
    // tie rule 1: simulation index and search serial of this tree key the Philox draws (read before lane 0 advances the counter)
    const u32 sim_index = TIES ? (u32)tv.meta[META_SELECTS] : 0u;
    const u32 serial = TIES ? (u32)tv.meta[META_SERIAL] : 0u;
    for (;;) {
        int winner = info_winner(info);
        int ne = info_ne(info);
        if (winner) { leaf_kind = LEAF_TERMINAL; break; }
        if (ne == 0) { leaf_kind = LEAF_EVAL; break; }           // Node.isLeaf()
        int eb = info_eb(info);
        // one pass over the edges: N, W, P of edges lane, lane+32, ... stay in registers (<= 126 legal moves)
        // (tried, r02h: skipping the chunks past the node's last edge with warp-uniform branches — mean branching is 24, one chunk
        // of the four usually suffices — same trees; stub search 1.323 against 1.315 ms, self-play ply 29.6 against 29.0 ms: the
        // predicated-off loads cost less than the branches)
        u32 Nr[4]; double Wr[4], Pr[4]; int Cr[4]; u64 Ir[4];
        u32 nsum = 0;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            int j = lane + 32 * c;
            bool in = j < ne;
            if (PREFETCH) {
                Cr[c] = in ? tv.eChild[eb + j] : -1;
                Ir[c] = in ? tv.eInfo[eb + j] : 0ULL;
            }
            Nr[c] = in ? tv.eN[eb + j] : 0u;
            Wr[c] = in ? tv.eQ[eb + j] : 0.0;                    // Q = W / N, stored by the backup
            Pr[c] = in ? tv.eP[eb + j] : 0.0;
            nsum += Nr[c];
        }
        nsum = __reduce_add_sync(FULL, nsum);                    // N_sum (MCTS.py:58-59)
        double sq = sqrt_count(nsum);                            // np.sqrt(N_sum)
        double best = -INFINITY; int besti = 0x7FFFFFFF;
        double QUr[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            int j = lane + 32 * c;
            QUr[c] = -INFINITY;
            if (j < ne) {
                u32 N = Nr[c];
                double Q = Wr[c];                                                               // MCTS.py:89,118
                double U = __ddiv_rn(__dmul_rn(__dmul_rn(cpuct, Pr[c]), sq), __dadd_rn(1.0, (double)N));   // :62
                double QU = __dadd_rn(Q, U);                                                    // :63
                if (TIES) QUr[c] = QU;
                if (QU > best) { best = QU; besti = j; }                                        // :65-67
            }
        }
        // warp arg-max, ties to the smallest edge index (= first maximal edge in list order): order-preserving 64-bit key of
        // the double (no NaNs here; + 0.0 folds a -0.0 into +0.0 so that equal values have equal keys), three warp reductions
        {
            u64 u = (u64)__double_as_longlong(best + 0.0);
            u64 key = (u >> 63) ? ~u : (u | 0x8000000000000000ULL);
            u32 hi = (u32)(key >> 32), lo = (u32)key;
            u32 mhi = __reduce_max_sync(FULL, hi);
            u32 mlo = __reduce_max_sync(FULL, hi == mhi ? lo : 0u);
            besti = (int)__reduce_min_sync(FULL, (hi == mhi && lo == mlo) ? (u32)besti : 0x7FFFFFFFu);
            if (TIES) {
                // MCTS.py:65-72: chosen_edges = [first edge attaining the maximum] + every later edge with fabs(QU - maxQU) < EPSILON
                // (earlier near-ties were dropped by the reset at :66-67); random.choice -> Philox, index = mulhi(x, len)
                const u64 mkey = ((u64)mhi << 32) | mlo;
                const double M = __longlong_as_double((long long)((mkey >> 63) ? (mkey & 0x7FFFFFFFFFFFFFFFULL) : ~mkey));
                u32 cm[4]; int total = 0;
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    int j = lane + 32 * c;
                    bool cand = j < ne && (j == besti || (j > besti && fabs(QUr[c] - M) < TIE_EPSILON));
                    cm[c] = __ballot_sync(FULL, cand);
                    total += __popc(cm[c]);
                }
                if (total > 1) {
                    const u64 uid = tv.tie_uids ? (u64)tv.tie_uids[tv.tree_index] : (((u64)tv.tie[3] << 32) | tv.tie[2]) + tv.tree_index;
                    Philox4 rnd = philox4x32_10(tv.tie[0], tv.tie[1], sim_index, (u32)depth, (u32)uid ^ (serial * 0x9E3779B9u),
                                                (u32)(uid >> 32) ^ 0x7E1Bu);
                    int k = (int)__umulhi(rnd.x, (u32)total);
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        int pc = __popc(cm[c]);
                        if (k >= 0 && k < pc) { besti = 32 * c + (int)__fns(cm[c], 0, k + 1); k = -1; }
                        else if (k >= 0) k -= pc;
                    }
                }
            }
        }
        int e = eb + besti;
        if (lane == 0 && depth < tv.path_max) tv.path[depth] = e;
        depth++;
        int child; u64 cinfo;                                    // cinfo is meaningful iff child >= 0
        if (PREFETCH) {
            const int chunk = besti >> 5, src = besti & 31;
            child = __shfl_sync(FULL, chunk == 0 ? Cr[0] : chunk == 1 ? Cr[1] : chunk == 2 ? Cr[2] : Cr[3], src);
            cinfo = shfl64(chunk == 0 ? Ir[0] : chunk == 1 ? Ir[1] : chunk == 2 ? Ir[2] : Ir[3], src);
        } else {
            child = tv.eChild[e];
            cinfo = tv.eInfo[e];
        }
        if (child < 0) {
            // first visit: materialise the child state = Board.place on a copy (MCTS.py:104-105)
            int nn = tv.meta[META_NNODES];
            if (nn >= tv.npt) { if (lane == 0) tv.meta[META_OVERFLOW] = 1; leaf_kind = LEAF_DEAD; node = 0; break; }
            Game g = load_node_game(tv, node);
            u32 mv = tv.eMove[e];
            int id = (int)(mv >> 8), to = (int)(mv & 0xFF);
            int from = (int)((g.cells_me >> (8 * id)) & 0xFF);
            apply_move(g, id, from, to);
            int w = winner_of(g);
            __syncwarp();
            if (lane == 0) {
                store_node(tv, nn, g, make_info(0, 0, (u32)w, 0));
                tv.eChild[e] = nn;
                tv.eInfo[e] = make_info(0, 0, (u32)w, 0);
                tv.meta[META_NNODES] = nn + 1;
            }
            __syncwarp();
            node = nn;
            leaf_kind = w ? LEAF_TERMINAL : LEAF_EVAL;
            break;
        }
        node = child;
        info = cinfo;
    }
    path_len = depth;
    if (TIES && lane == 0) tv.meta[META_SELECTS] = (int32_t)(sim_index + 1u);
    return node;
}

// ---- backup (MCTS.py:83-90 terminal, 112-118 value) ---------------------------------------------------
__device__ __forceinline__ void backup(const TreeView &tv, int lane, int path_len, double v, bool terminal)
{
    for (int d = lane; d < path_len; d += 32) {
        int e = tv.path[d];
        bool same = ((path_len - d) & 1) == 0;                   // edge.currPlayer == leafNode.currPlayer
        double delta = terminal ? (same ? -1.0 : 1.0)            // REWARD['win'] * direction
                                : __dmul_rn(v, same ? 1.0 : -1.0);
        u32 N = tv.eN[e] + 1;
        double W = __dadd_rn(tv.eW[e], delta);
        tv.eN[e] = N;
        tv.eW[e] = W;
        tv.eQ[e] = __ddiv_rn(W, (double)N);                       // edge.stats['Q'] = W / N
    }
}

// ---- expand (MCTS.py:95-109) ---------------------------------------------------------------------------
// Lanes 0..5 run one checker's flood fill each; the six masks are then broadcast and every lane writes a
// strided share of the edge list.  prior(idx) is supplied by the evaluator functor.
// parent_edge = the edge that leads to `node` (last edge of the path), -1 for the root: its eInfo copy is refreshed.
template <typename PriorFn>
__device__ __forceinline__ bool expand_node(const TreeView &tv, int lane, int node, const Game &g, PriorFn prior, const uint8_t *sT,
                                            int parent_edge)
{
    u64 mine = 0;
    if (lane < 6) {
        u64 occ_all = g.occ_me | g.occ_op;
        u64 o = 1ULL << ((g.cells_me >> (8 * lane)) & 0xFF);
        u64 occ = occ_all & ~o, empty = ~occ & CCX_VALID;
        u64 todo = o, reach = 0;
        while (todo) {                                        // ray expansion of one cell per iteration (ccx_device.cuh)
            int i = 63 - __clzll((long long)todo);         // order-independent closure; top bit = one FLO
            todo ^= 1ULL << i;
            u64 nw = expand_cell(i, occ, sT) & ~(reach | o);
            reach |= nw; todo |= nw;
        }
        mine = (neighbours(o) & empty) | reach;
    }
    u64 dest[6]; int pre[7]; pre[0] = 0;
#pragma unroll
    for (int k = 0; k < 6; k++) { dest[k] = shfl64(mine, k); pre[k + 1] = pre[k] + __popcll(dest[k]); }
    int total = pre[6];
    int eb = tv.meta[META_NEDGES];
    if (eb + total > tv.ept) { if (lane == 0) tv.meta[META_OVERFLOW] = 1; return false; }
    for (int j = lane; j < total; j += 32) {
        int k = 0;
#pragma unroll
        for (int q = 1; q < 6; q++) k += (j >= pre[q]);
        u64 m = 0; int base = 0;
#pragma unroll
        for (int q = 0; q < 6; q++) if (k == q) { m = dest[q]; base = pre[q]; }
        int to = select64(m, (u32)(j - base));
        int idx = k * 49 + (to >> 3) * 7 + (to & 7);             // utils.encode_checker_index
        tv.eN[eb + j] = 0; tv.eW[eb + j] = 0.0; tv.eQ[eb + j] = 0.0; tv.eP[eb + j] = prior(idx);     // MCTS.py:32-37,108
        tv.eChild[eb + j] = -1;
        tv.eMove[eb + j] = (uint16_t)((k << 8) | to);
    }
    __syncwarp();
    if (lane == 0) {
        tv.meta[META_NEDGES] = eb + total;
        u64 *info = tv.node + (int64_t)node * NODE_WORDS + 5;
        *info = make_info((u32)eb, (u32)total, 0, 1);
        if (parent_edge >= 0) tv.eInfo[parent_edge] = make_info((u32)eb, (u32)total, 0, 1);
    }
    __syncwarp();
    return true;
}

struct UniformPrior { __device__ double operator()(int) const { return 1.0 / 294.0; } };
struct HashPrior {
    u32 k0, k1;
    __device__ double operator()(int idx) const { return (double)philox4x32_10(k0, k1, (u32)idx, 7u, 0u, 0u).x / 4294967296.0; }
};
struct TablePrior {      // priors from an evaluator's output row p[294] (float64)
    const double *p;
    __device__ double operator()(int idx) const { return p[idx]; }
};

// evaluate + expand + backup of one leaf with an in-kernel evaluator
template <int EVAL>
__device__ __forceinline__ void eval_expand_backup(const TreeView &tv, int lane, int leaf, int path_len, const uint8_t *sT)
{
    Game g = load_node_game(tv, leaf);
    const int parent_edge = path_len > 0 ? tv.path[path_len - 1] : -1;
    double v = 0.0;
    bool ok;
    if (EVAL == EVAL_HASH) {
        u64 h = leaf_hash(g, lane);
        HashPrior pr = {(u32)h, (u32)(h >> 32)};
        v = (double)philox4x32_10(pr.k0, pr.k1, 294u, 7u, 0u, 0u).x / 2147483648.0 - 1.0;
        ok = expand_node(tv, lane, leaf, g, pr, sT, parent_edge);
    } else {
        ok = expand_node(tv, lane, leaf, g, UniformPrior(), sT, parent_edge);
    }
    if (ok) backup(tv, lane, path_len, v, false);
}

// selfplay.py:121-124 — root priors mixed with caller-supplied Dirichlet noise (one value per root edge)
__device__ __forceinline__ void mix_root_noise(const TreeView &tv, int lane, const double *noise, bool normalize = false)
{
    u64 info = tv.node[5];
    int ne = info_ne(info), eb = info_eb(info);
    double scale = 1.0;
    if (normalize) {                      // raw gamma draws -> Dirichlet sample over the ne root edges
        double sum = 0.0;
        for (int j = lane; j < ne; j += 32) sum += noise[j];
#pragma unroll
        for (int off = 16; off; off >>= 1) sum += __shfl_xor_sync(FULL, sum, off);
        scale = sum > 0.0 ? 1.0 / sum : 0.0;
    }
    for (int j = lane; j < ne; j += 32) {
        double p = __dmul_rn(tv.eP[eb + j], 1. - 0.25);
        double nz = normalize ? noise[j] * scale : noise[j];
        tv.eP[eb + j] = __dadd_rn(p, __dmul_rn(0.25, nz));
    }
    __syncwarp();
}

// A root whose status is not RUNNING or whose ply count is below min_ply gets an inactive tree
// (META_OVERFLOW = 2): every later phase skips it (self-play: opening random plies, finished games).
__device__ __forceinline__ void init_tree(const TreeView &tv, int lane, const u64 *roots, int64_t n, int64_t tree, int min_ply = 0,
                                          int ply_parity = -1)
{
    if (lane == 0) {
        u64 m = roots[4 * n + tree];
        bool inactive = min_ply >= 0 && ((m >> 56) != 0 || (int)((m >> 32) & 0xFFFF) < min_ply);
        if (ply_parity >= 0 && (int)((m >> 32) & 1) != ply_parity) inactive = true;      // two-net self-play: the other net's plies
        tv.meta[META_SELECTS] = 0;
        // "search serial" of the tie rule: a hash of the root position, so that the draws of a search depend on (seed, tree uid,
        // root) only — never on what the handle searched before
        u64 z = roots[0 * n + tree] * 0x9E3779B97F4A7C15ULL + roots[1 * n + tree] * 0xC2B2AE3D27D4EB4FULL + (m & 0x0000FFFFFFFFFFFFULL);
        tv.meta[META_SERIAL] = (int32_t)((u32)(z >> 32) ^ (u32)z);
        Game g = game_of_words(roots[0 * n + tree], roots[1 * n + tree], roots[2 * n + tree], roots[3 * n + tree],
                               roots[4 * n + tree] & 0x00FFFFFFFFFFFFFFULL);
        store_node(tv, 0, g, make_info(0, 0, (u32)winner_of(g), 0));
        tv.meta[META_NNODES] = 1; tv.meta[META_NEDGES] = 0; tv.meta[META_OVERFLOW] = inactive ? 2 : 0; tv.meta[META_PATHLEN] = 0;
    }
    __syncwarp();
}

// Persistent search with an in-kernel evaluator: every warp runs all simulations of its tree.
// select_leaf<PREFETCH> in the persistent kernel: measured again in r02 under the 72-register cap (one resident wave): 1.321 ms
// per 4,096-tree search with the prefetch against 1.310 ms without (round 1 saw -19 % because 84 registers split the launch
// into 1.15 waves).  No gain either way: at 1,100 warp instructions per simulation and 59 % issue utilisation the kernel is as
// much instruction- as latency-bound.
#ifndef CCX_SEARCH_PREFETCH
#define CCX_SEARCH_PREFETCH false
#endif
template <int EVAL, bool TIES>
__global__ void __launch_bounds__(32 * MCTS_WARPS_PER_BLOCK, 14)
k_mcts_search(ccx_trees trees, const u64 *__restrict__ roots, int64_t n, int num_itr, double cpuct, int pre_expand,
              const double *__restrict__ noise, int noise_stride, const uint8_t *__restrict__ jt)
{
    __shared__ __align__(16) uint8_t sT[CCX_JT_BYTES];
    for (int q = threadIdx.x; q < CCX_JT_BYTES / 16; q += blockDim.x) reinterpret_cast<uint4 *>(sT)[q] = reinterpret_cast<const uint4 *>(jt)[q];
    __syncthreads();
    int64_t tree = (int64_t)blockIdx.x * MCTS_WARPS_PER_BLOCK + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (tree >= n) return;
    TreeView tv = tree_view(trees, tree);
    init_tree(tv, lane, roots, n, tree, -1);
    if (pre_expand && info_winner(tv.node[5]) == 0) {                    // selfplay.py:117
        eval_expand_backup<EVAL>(tv, lane, 0, 0, sT);
        if (noise) mix_root_noise(tv, lane, noise + tree * noise_stride);
    }
    if (pre_expand && TIES && lane == 0) tv.meta[META_SELECTS] = 1;      // the round-based drivers spend selection 0 on the root expansion
    __syncwarp();
    for (int it = 0; it < num_itr; it++) {                                   // MCTS.py:123-125
        int path_len, kind;
        int leaf = select_leaf<CCX_SEARCH_PREFETCH, TIES>(tv, lane, cpuct, path_len, kind);
        __syncwarp();
        if (kind == LEAF_TERMINAL) backup(tv, lane, path_len, 0.0, true);
        else if (kind == LEAF_EVAL) eval_expand_backup<EVAL>(tv, lane, leaf, path_len, sT);
        __syncwarp();
        if (tv.meta[META_OVERFLOW]) break;
    }
}

// ---- round-based pieces for an external evaluator (the policy/value net) ----------------------------
// phase A: select one leaf per tree; terminal leaves are backed up at once; for the others the leaf's
// state words are written to leaf_state[5][n] (input of ccx_encode / the net)
template <bool TIES>
__global__ void __launch_bounds__(32 * MCTS_WARPS_PER_BLOCK)
k_mcts_select(ccx_trees trees, int64_t n, double cpuct, u64 *__restrict__ leaf_state)
{
    int64_t tree = (int64_t)blockIdx.x * MCTS_WARPS_PER_BLOCK + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (tree >= n) return;
    TreeView tv = tree_view(trees, tree);
    if (tv.meta[META_OVERFLOW]) {
        // inactive / overflowed tree: hand the evaluator a well-formed dummy position (its output is ignored)
        if (lane == 0) {
            tv.meta[META_LEAFKIND] = LEAF_DEAD;
            leaf_state[0 * n + tree] = CCX_START_OCC1; leaf_state[1 * n + tree] = CCX_START_OCC2;
            leaf_state[2 * n + tree] = CCX_START_CELLS1; leaf_state[3 * n + tree] = CCX_START_CELLS2;
            leaf_state[4 * n + tree] = CCX_START_META;
        }
        return;
    }
    int path_len, kind;
    int leaf = select_leaf<true, TIES>(tv, lane, cpuct, path_len, kind);
    __syncwarp();
    if (kind == LEAF_TERMINAL) backup(tv, lane, path_len, 0.0, true);
    if (lane < 5) leaf_state[lane * n + tree] = tv.node[(int64_t)leaf * NODE_WORDS + lane];
    if (lane == 0) { tv.meta[META_PATHLEN] = path_len; tv.meta[META_LEAF] = leaf; tv.meta[META_LEAFKIND] = kind; }
}

// phase B: expand the selected leaf with priors p[n][294] (float64, already soft-maxed: model.py:21-24)
// and back up v[n]; root_only_noise != NULL mixes Dirichlet noise into the root priors (first call).
__global__ void __launch_bounds__(32 * MCTS_WARPS_PER_BLOCK)
k_mcts_expand_backup(ccx_trees trees, int64_t n, const double *__restrict__ p, const double *__restrict__ v,
                     const double *__restrict__ noise, int noise_stride, int noise_normalize, const uint8_t *__restrict__ jt)
{
    __shared__ __align__(16) uint8_t sT[CCX_JT_BYTES];
    for (int q = threadIdx.x; q < CCX_JT_BYTES / 16; q += blockDim.x) reinterpret_cast<uint4 *>(sT)[q] = reinterpret_cast<const uint4 *>(jt)[q];
    __syncthreads();
    int64_t tree = (int64_t)blockIdx.x * MCTS_WARPS_PER_BLOCK + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (tree >= n) return;
    TreeView tv = tree_view(trees, tree);
    if (tv.meta[META_LEAFKIND] != LEAF_EVAL) return;
    int leaf = tv.meta[META_LEAF], path_len = tv.meta[META_PATHLEN];
    Game g = load_node_game(tv, leaf);
    TablePrior pr = {p + tree * CCX_NUM_ACTIONS};
    if (expand_node(tv, lane, leaf, g, pr, sT, path_len > 0 ? tv.path[path_len - 1] : -1)) backup(tv, lane, path_len, v[tree], false);
    if (noise && leaf == 0) mix_root_noise(tv, lane, noise + tree * noise_stride, noise_normalize != 0);
}

// ---- fused round pieces for the built-in net evaluator (ccx_mcts_run_net) -------------------------------
// phase A': select + utils.to_model_input of the leaf (utils.py:101-160) written straight into the net's uint8
// input batch, one warp per tree (the leaf's 343 bytes are staged in shared memory: zero, scatter the 36
// labels, copy out)
template <bool TIES>
__device__ __forceinline__ void do_select_encode(const TreeView &tv, int lane, double cpuct, uint8_t *__restrict__ sp,
                                                 uint8_t *__restrict__ dst)
{
    Game g;
    if (tv.meta[META_OVERFLOW]) {
        // inactive / overflowed tree: hand the evaluator a well-formed dummy position (its output is ignored)
        if (lane == 0) tv.meta[META_LEAFKIND] = LEAF_DEAD;
        reset_start(g);
    } else {
        int path_len, kind;
        int leaf = select_leaf<true, TIES>(tv, lane, cpuct, path_len, kind);
        __syncwarp();
        if (kind == LEAF_TERMINAL) backup(tv, lane, path_len, 0.0, true);
        if (lane == 0) { tv.meta[META_PATHLEN] = path_len; tv.meta[META_LEAF] = leaf; tv.meta[META_LEAFKIND] = kind; }
        g = load_node_game(tv, leaf);
    }
    for (int i = lane; i < 88; i += 32) reinterpret_cast<uint32_t *>(sp)[i] = 0u;
    __syncwarp();
    {
        const int plies = (int)((g.meta >> 32) & 0xFFFF);
        const u64 cur0 = g.cells_me, opp0 = g.cells_op;
        const u64 opp1 = undo_in_cells(opp0, (int)(g.meta & 0xFF), (int)((g.meta >> 8) & 0xFF));
        const u64 cur2 = undo_in_cells(cur0, (int)((g.meta >> 16) & 0xFF), (int)((g.meta >> 24) & 0xFF));
        for (int e = lane; e < 36; e += 32) {
            int hs = e / 12, side = (e / 6) & 1, id = e % 6;
            if (hs > plies) continue;
            u64 cells = side ? (hs >= 1 ? opp1 : opp0) : (hs >= 2 ? cur2 : cur0);
            int c = (int)((cells >> (8 * id)) & 0xFF);
            if (c < 55 && (c & 7) < 7) sp[((c >> 3) * 7 + (c & 7)) * 7 + 2 * hs + side] = (uint8_t)(id + 1);
        }
        if ((g.meta >> 48) & 1)
            for (int c = lane; c < 49; c += 32) sp[c * 7 + 6] = 1;                // utils.py:157-158
    }
    __syncwarp();
    for (int i = lane; i < 343; i += 32) dst[i] = sp[i];
}

template <bool TIES>
__global__ void __launch_bounds__(32 * MCTS_WARPS_PER_BLOCK, 14)
k_mcts_select_encode(ccx_trees trees, int64_t tree0, int64_t n, double cpuct, uint8_t *__restrict__ planes)
{
    __shared__ __align__(16) uint8_t sP[MCTS_WARPS_PER_BLOCK][352];
    int64_t tree = tree0 + (int64_t)blockIdx.x * MCTS_WARPS_PER_BLOCK + (threadIdx.x >> 5);      // trees [tree0, tree0 + n)
    int lane = threadIdx.x & 31;
    if (tree >= tree0 + n) return;
    TreeView tv = tree_view(trees, tree);
    do_select_encode<TIES>(tv, lane, cpuct, sP[threadIdx.x >> 5], planes + tree * 343);
}

// phase B': float64 softmax over all 294 logits (model.py:21-24, utils.py:187-192; same arithmetic as
// k_softmax_f64) + expand + backup, one warp per tree; the priors never touch global memory
__device__ __forceinline__ void do_softmax_expand_backup(const TreeView &tv, int lane, const float *__restrict__ lg, double v,
                                                         const double *__restrict__ noise, int noise_normalize, double *__restrict__ pr,
                                                         const uint8_t *__restrict__ sT)
{
    if (tv.meta[META_LEAFKIND] != LEAF_EVAL) return;
    {
        double x[10], mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 10; j++) {
            int i = lane + 32 * j;
            x[j] = i < 294 ? (double)lg[i] : -INFINITY;
            mx = fmax(mx, x[j]);
        }
#pragma unroll
        for (int off = 16; off; off >>= 1) mx = fmax(mx, __shfl_xor_sync(FULL, mx, off));
        double sum = 0.0;
#pragma unroll
        for (int j = 0; j < 10; j++) { x[j] = (lane + 32 * j) < 294 ? exp(x[j] - mx) : 0.0; sum += x[j]; }
#pragma unroll
        for (int off = 16; off; off >>= 1) sum += __shfl_xor_sync(FULL, sum, off);
#pragma unroll
        for (int j = 0; j < 10; j++) if (lane + 32 * j < 294) pr[lane + 32 * j] = x[j] / sum;
    }
    __syncwarp();
    int leaf = tv.meta[META_LEAF], path_len = tv.meta[META_PATHLEN];
    Game g = load_node_game(tv, leaf);
    TablePrior tp = {pr};
    if (expand_node(tv, lane, leaf, g, tp, sT, path_len > 0 ? tv.path[path_len - 1] : -1)) backup(tv, lane, path_len, v, false);
    if (noise && leaf == 0) mix_root_noise(tv, lane, noise, noise_normalize != 0);
}

__global__ void __launch_bounds__(32 * MCTS_WARPS_PER_BLOCK)
k_mcts_softmax_expand_backup(ccx_trees trees, int64_t tree0, int64_t n, const float *__restrict__ logits, const float *__restrict__ value,
                             const double *__restrict__ noise, int noise_stride, int noise_normalize, const uint8_t *__restrict__ jt)
{
    __shared__ double sPr[MCTS_WARPS_PER_BLOCK][296];
    const uint8_t *sT = jt;          // one expansion per launch: the 6 KB jump table is read through L1 instead of staged per block
    int64_t tree = tree0 + (int64_t)blockIdx.x * MCTS_WARPS_PER_BLOCK + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (tree >= tree0 + n) return;
    TreeView tv = tree_view(trees, tree);
    do_softmax_expand_backup(tv, lane, logits + tree * 294, (double)value[tree], noise ? noise + tree * noise_stride : nullptr,
                             noise_normalize, sPr[threadIdx.x >> 5], sT);
}

// one launch per round in steady state: finish the previous round's leaf (softmax + expand + backup), then select and encode
// the next one — the same warp owns the tree in both halves, so the halves need no grid-wide ordering between them
// launch bound: 14 blocks per SM (<= 72 registers) keeps all 2,048 blocks of a 4,096-tree round resident in ONE wave (148 x 14 =
// 2,072); at 80 registers (12 blocks per SM) the last 272 blocks start only when others finish: 1.15 waves, slowest SM 88.6 K cycles
// against 59.7 K on average (profiles/r02b_kernels_ncu.md)
template <bool TIES>
__global__ void __launch_bounds__(32 * MCTS_WARPS_PER_BLOCK, 14)
k_mcts_round(ccx_trees trees, int64_t tree0, int64_t n, double cpuct, const float *__restrict__ logits, const float *__restrict__ value,
             const double *__restrict__ noise, int noise_stride, int noise_normalize, uint8_t *__restrict__ planes,
             const uint8_t *__restrict__ jt)
{
    __shared__ double sPr[MCTS_WARPS_PER_BLOCK][296];
    __shared__ __align__(16) uint8_t sP[MCTS_WARPS_PER_BLOCK][352];
    const uint8_t *sT = jt;          // one expansion per launch: the 6 KB jump table is read through L1 instead of staged per block
    int64_t tree = tree0 + (int64_t)blockIdx.x * MCTS_WARPS_PER_BLOCK + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (tree >= tree0 + n) return;
    TreeView tv = tree_view(trees, tree);
    do_softmax_expand_backup(tv, lane, logits + tree * 294, (double)value[tree], noise ? noise + tree * noise_stride : nullptr,
                             noise_normalize, sPr[threadIdx.x >> 5], sT);
    __syncwarp();
    __threadfence_block();
    do_select_encode<TIES>(tv, lane, cpuct, sP[threadIdx.x >> 5], planes + tree * 343);
}

__global__ void __launch_bounds__(32 * MCTS_WARPS_PER_BLOCK)
k_mcts_init(ccx_trees trees, const u64 *__restrict__ roots, int64_t n, int min_ply, int ply_parity)
{
    int64_t tree = (int64_t)blockIdx.x * MCTS_WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (tree >= n) return;
    TreeView tv = tree_view(trees, tree);
    init_tree(tv, threadIdx.x & 31, roots, n, tree, min_ply, ply_parity);
}

// ---- finalize (MCTS.py:131-137): visit counts, pi = N^(1/tau) / sum, root Q ---------------------------
__global__ void __launch_bounds__(32 * MCTS_WARPS_PER_BLOCK)
k_mcts_finalize(ccx_trees trees, int64_t n, double inv_tau, u32 *__restrict__ visits, double *__restrict__ pi,
                double *__restrict__ q, int32_t *__restrict__ n_nodes)
{
    int64_t tree = (int64_t)blockIdx.x * MCTS_WARPS_PER_BLOCK + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (tree >= n) return;
    TreeView tv = tree_view(trees, tree);
    for (int a = lane; a < CCX_NUM_ACTIONS; a += 32) {
        visits[tree * CCX_NUM_ACTIONS + a] = 0;
        if (pi) pi[tree * CCX_NUM_ACTIONS + a] = 0.0;
        if (q) q[tree * CCX_NUM_ACTIONS + a] = 0.0;
    }
    __syncwarp();
    u64 info = tv.node[5];
    int ne = info_ne(info), eb = info_eb(info);
    double sum = 0.0;
    for (int j = lane; j < ne; j += 32) {
        u32 N = tv.eN[eb + j];
        double pw = inv_tau == 1.0 ? (double)N : pow((double)N, inv_tau);       // MCTS.py:132
        sum += pw;
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) sum += __shfl_xor_sync(FULL, sum, off);   // exact for tau = 1 (integers)
    for (int j = lane; j < ne; j += 32) {
        u32 N = tv.eN[eb + j];
        u32 mv = tv.eMove[eb + j];
        int to = (int)(mv & 0xFF);
        int idx = (int)(mv >> 8) * 49 + (to >> 3) * 7 + (to & 7);
        visits[tree * CCX_NUM_ACTIONS + idx] = N;
        if (pi) {
            double pw = inv_tau == 1.0 ? (double)N : pow((double)N, inv_tau);
            pi[tree * CCX_NUM_ACTIONS + idx] = sum > 0.0 ? __ddiv_rn(pw, sum) : 0.0;       // MCTS.py:137
        }
        if (q) q[tree * CCX_NUM_ACTIONS + idx] = N ? __ddiv_rn(tv.eW[eb + j], (double)N) : 0.0;
    }
    if (n_nodes && lane == 0)
        n_nodes[tree] = tv.meta[META_OVERFLOW] ? -tv.meta[META_OVERFLOW] : tv.meta[META_NEDGES] + 1;   // reference node count: root + one per edge
}


// ---- root edge access for the Python Node/Edge mirror (MCTS.py:24-37; selfplay.py:121-124 mutates P) ----
__global__ void __launch_bounds__(32 * MCTS_WARPS_PER_BLOCK)
k_mcts_get_root(ccx_trees trees, int64_t n, int stride, int32_t *__restrict__ n_edges, uint16_t *__restrict__ moves,
                u32 *__restrict__ N, double *__restrict__ W, double *__restrict__ P)
{
    int64_t tree = (int64_t)blockIdx.x * MCTS_WARPS_PER_BLOCK + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (tree >= n) return;
    TreeView tv = tree_view(trees, tree);
    u64 info = tv.node[5];
    int ne = info_ne(info), eb = info_eb(info);
    if (lane == 0) n_edges[tree] = ne;
    for (int j = lane; j < ne && j < stride; j += 32) {
        moves[tree * stride + j] = tv.eMove[eb + j];
        N[tree * stride + j] = tv.eN[eb + j];
        W[tree * stride + j] = tv.eW[eb + j];
        P[tree * stride + j] = tv.eP[eb + j];
    }
}

__global__ void __launch_bounds__(32 * MCTS_WARPS_PER_BLOCK)
k_mcts_set_root_priors(ccx_trees trees, int64_t n, int stride, const double *__restrict__ P)
{
    int64_t tree = (int64_t)blockIdx.x * MCTS_WARPS_PER_BLOCK + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (tree >= n) return;
    TreeView tv = tree_view(trees, tree);
    u64 info = tv.node[5];
    int ne = info_ne(info), eb = info_eb(info);
    for (int j = lane; j < ne && j < stride; j += 32) tv.eP[eb + j] = P[tree * stride + j];
}

// ---- host side ----------------------------------------------------------------------------------------

// cached CUDA graph of the round loop (see ccx_mcts_run_net)
struct ccx_round_graph {
    cudaGraphExec_t exec = nullptr;
    bool seen = false;
    uint64_t epoch = 0;
    int64_t n = 0, launches = 0;
    int32_t rounds = 0, stride = 0, normalize = 0;
    double cpuct = 0.0;
    const double *noise = nullptr;
    ccx_trees trees;
};

void ccx_round_graph_free(ccx_handle *h)
{
    ccx_round_graph *g = (ccx_round_graph *)h->round_graph;
    if (g) {
        if (g->exec) cudaGraphExecDestroy(g->exec);
        delete g;
    }
    h->round_graph = nullptr;
    if (h->cap_stream) { cudaStreamDestroy(h->cap_stream); h->cap_stream = nullptr; }
}

void ccx_trees_set_uids(ccx_handle *h, const int64_t *uids)
{
    if (h->trees) h->trees->tie_uids = uids;          // by-value kernel argument: part of the cached round graph's key
}

void ccx_trees_free(ccx_handle *h)
{
    ccx_trees *t = h->trees;
    if (!t) return;
    ccx_round_graph_free(h);
    void *ptrs[] = {t->node, t->eN, t->eW, t->eP, t->eQ, t->eChild, t->eInfo, t->eMove, t->path, t->tree_meta, t->tie_params};
    for (void *p : ptrs) if (p) cudaFree(p);
    delete t;
    h->trees = nullptr;
}

static int trees_reserve(ccx_handle *h, int64_t n, int32_t num_itr, int32_t edges_per_tree)
{
    {   // per device: the constant table lives in this module's image on the handle's device
        static bool filled[64] = {false};
        if (h->device >= 0 && h->device < 64 && !filled[h->device]) {
            static double host_sqrt[SQRT_TABLE];
            for (int i = 0; i < SQRT_TABLE; i++) host_sqrt[i] = sqrt((double)i);
            CCX_CUDA(h, cudaMemcpyToSymbol(c_sqrt, host_sqrt, sizeof(host_sqrt)));
            filled[h->device] = true;
        }
    }
    int32_t npt = num_itr + 2;
    int32_t ept = edges_per_tree > 0 ? edges_per_tree : 64 * (num_itr + 1);
    int32_t pm = num_itr + 2;
    ccx_trees *t = h->trees;
    if (t && t->cap_trees >= n && t->nodes_per_tree >= npt && t->edges_per_tree == ept && t->path_max >= pm) return CCX_OK;
    ccx_trees_free(h);
    h->epoch++;
    t = new (std::nothrow) ccx_trees();
    if (!t) return CCX_ERR_NOMEM;
    h->trees = t;
    t->tie_mode = h->tie_mode;
    t->tie_uids = h->slot_ids;
    t->cap_trees = n; t->nodes_per_tree = npt; t->edges_per_tree = ept; t->path_max = pm;
    size_t T = (size_t)n;
    CCX_CUDA(h, cudaMalloc(&t->node, T * npt * NODE_WORDS * 8));
    CCX_CUDA(h, cudaMalloc(&t->eN, T * ept * 4));
    CCX_CUDA(h, cudaMalloc(&t->eW, T * ept * 8));
    CCX_CUDA(h, cudaMalloc(&t->eP, T * ept * 8));
    CCX_CUDA(h, cudaMalloc(&t->eQ, T * ept * 8));
    CCX_CUDA(h, cudaMalloc(&t->eChild, T * ept * 4));
    CCX_CUDA(h, cudaMalloc(&t->eInfo, T * ept * 8));
    CCX_CUDA(h, cudaMalloc(&t->eMove, T * ept * 2));
    CCX_CUDA(h, cudaMalloc(&t->path, T * pm * 4));
    CCX_CUDA(h, cudaMalloc(&t->tree_meta, T * 8 * 4));
    CCX_CUDA(h, cudaMalloc(&t->tie_params, 16));
    {
        const u32 tp[4] = {(u32)h->tie_seed, (u32)(h->tie_seed >> 32), (u32)(uint64_t)h->tie_uid0, (u32)((uint64_t)h->tie_uid0 >> 32)};
        CCX_CUDA(h, cudaMemcpyAsync(t->tie_params, tp, 16, cudaMemcpyHostToDevice, h->stream));    // pageable source: staged before the call returns
    }
    t->bytes = T * ((size_t)npt * NODE_WORDS * 8 + (size_t)ept * 42 + (size_t)pm * 4 + 32);
    return CCX_OK;
}

static inline unsigned tree_blocks(int64_t n) { return (unsigned)((n + MCTS_WARPS_PER_BLOCK - 1) / MCTS_WARPS_PER_BLOCK); }

extern "C" {

int ccx_mcts_search(ccx_handle *h, int64_t n, const uint64_t *roots, int32_t evaluator, int32_t num_itr, double cpuct,
                    double tau, int32_t pre_expand, const double *root_noise, int32_t noise_stride,
                    int32_t edges_per_tree, uint32_t *visits, double *pi, double *q, int32_t *n_nodes)
{
    if (!h || n < 0 || num_itr < 0 || !(tau > 0.0) || (n && (!roots || !visits))) return CCX_ERR_ARG;
    if (evaluator != EVAL_UNIFORM && evaluator != EVAL_HASH) return CCX_ERR_UNSUPPORTED;   // the net runs round-based
    if (root_noise && noise_stride < 1) return CCX_ERR_ARG;
    if (n == 0) return CCX_OK;
    int rc = trees_reserve(h, n, num_itr, edges_per_tree);
    if (rc) return rc;
    unsigned grid = tree_blocks(n);
#define CCX_SEARCH(EV, TI) k_mcts_search<EV, TI><<<grid, 32 * MCTS_WARPS_PER_BLOCK, 0, h->stream>>>(*h->trees, (const u64 *)roots, n, num_itr, \
                                                                                                   cpuct, pre_expand, root_noise, noise_stride, h->jump_table)
    const bool ties = h->trees->tie_mode != 0;
    if (evaluator == EVAL_UNIFORM) { if (ties) CCX_SEARCH(EVAL_UNIFORM, true); else CCX_SEARCH(EVAL_UNIFORM, false); }
    else { if (ties) CCX_SEARCH(EVAL_HASH, true); else CCX_SEARCH(EVAL_HASH, false); }
#undef CCX_SEARCH
    CCX_LAUNCHED(h);
    k_mcts_finalize<<<grid, 32 * MCTS_WARPS_PER_BLOCK, 0, h->stream>>>(*h->trees, n, 1.0 / tau, visits, pi, q, n_nodes);
    CCX_LAUNCHED(h);
    return CCX_OK;
}

int ccx_mcts_set_tiebreak(ccx_handle *h, int32_t mode, uint64_t seed, int64_t uid0)
{
    if (!h || (mode != 0 && mode != 1)) return CCX_ERR_ARG;
    h->tie_mode = mode; h->tie_seed = seed; h->tie_uid0 = uid0;
    if (h->trees) {
        ccx_trees *t = h->trees;
        t->tie_mode = mode;                           // by-value kernel argument: selects the instantiation, keys the cached round graph
        const u32 tp[4] = {(u32)seed, (u32)(seed >> 32), (u32)(uint64_t)uid0, (u32)((uint64_t)uid0 >> 32)};
        CCX_CUDA(h, cudaMemcpyAsync(t->tie_params, tp, 16, cudaMemcpyHostToDevice, h->stream));     // in stream order with the searches
    }
    return CCX_OK;
}

int ccx_mcts_begin(ccx_handle *h, int64_t n, const uint64_t *roots, int32_t num_itr, int32_t edges_per_tree, int32_t min_ply,
                   int32_t ply_parity)
{
    if (!h || n < 0 || num_itr < 0 || (n && !roots) || ply_parity < -1 || ply_parity > 1) return CCX_ERR_ARG;
    if (n == 0) return CCX_OK;
    int rc = trees_reserve(h, n, num_itr, edges_per_tree);
    if (rc) return rc;
    k_mcts_init<<<tree_blocks(n), 32 * MCTS_WARPS_PER_BLOCK, 0, h->stream>>>(*h->trees, (const u64 *)roots, n, min_ply, ply_parity);
    CCX_LAUNCHED(h);
    return CCX_OK;
}

int ccx_mcts_select(ccx_handle *h, int64_t n, double cpuct, uint64_t *leaf_state)
{
    if (!h || !h->trees || n < 0 || n > h->trees->cap_trees || (n && !leaf_state)) return CCX_ERR_ARG;
    if (n == 0) return CCX_OK;
    if (h->trees->tie_mode) k_mcts_select<true><<<tree_blocks(n), 32 * MCTS_WARPS_PER_BLOCK, 0, h->stream>>>(*h->trees, n, cpuct, (u64 *)leaf_state);
    else k_mcts_select<false><<<tree_blocks(n), 32 * MCTS_WARPS_PER_BLOCK, 0, h->stream>>>(*h->trees, n, cpuct, (u64 *)leaf_state);
    CCX_LAUNCHED(h);
    return CCX_OK;
}

int ccx_mcts_expand_backup(ccx_handle *h, int64_t n, const double *p, const double *v, const double *root_noise,
                           int32_t noise_stride, int32_t noise_normalize)
{
    if (!h || !h->trees || n < 0 || n > h->trees->cap_trees || (n && (!p || !v))) return CCX_ERR_ARG;
    if (n == 0) return CCX_OK;
    k_mcts_expand_backup<<<tree_blocks(n), 32 * MCTS_WARPS_PER_BLOCK, 0, h->stream>>>(*h->trees, n, p, v, root_noise, noise_stride, noise_normalize, h->jump_table);
    CCX_LAUNCHED(h);
    return CCX_OK;
}

// the round loop itself, issued on h->stream (and h->stream2 for the second half of a split batch)
static int run_net_rounds(ccx_handle *h, int64_t n, int32_t rounds, double cpuct, const double *root_noise, int32_t noise_stride,
                          int32_t noise_normalize)
{
    uint8_t *planes; float *logits, *value;
    int rc;
    if ((rc = ccx_net_scratch(h, n, &planes, &logits, &value))) return rc;
    // Two halves of the batch run as two independent round pipelines on two streams: trees are independent, and the
    // low-occupancy stretches of one half (the tail of its tree kernel = the deepest trees, the last tile of its trunk kernel)
    // are filled by the other half's kernels.  Same trees as the single-stream order, bit for bit.
    static const bool no_split = getenv("CCX_NO_SPLIT") != nullptr;
    // batch size from which the two-half pipeline is used: 8,192 in the 16-bit mode; 16,384 in the accurate mode, whose three-context
    // trunk loses more on a halved batch (r02h, accurate mode: 8,192 slots 52.2 ms unsplit / 53.7 split, 16,384 slots 100.8 / 98.3)
    static const int64_t split_env = getenv("CCX_SPLIT_MIN") ? atoll(getenv("CCX_SPLIT_MIN")) : -1;
    const int64_t split_min = split_env >= 0 ? split_env : h->net_mode == 2 ? 16384 : 8192;
    const bool split = !no_split && (h->net_mode == 1 || h->net_mode == 2) && n >= split_min;     // measured (16-bit mode): 16,384 slots 65.1 -> 62.2 ms per ply; no gain at 4,096
    int64_t part_n[2] = {split ? (n / 2) & ~(int64_t)127 : n, 0};          // halves start on a 128-position tile of the net's scratch
    part_n[1] = n - part_n[0];
    cudaStream_t streams[2] = {h->stream, h->stream};
    if (split) {
        if (!h->stream2) {
            CCX_CUDA(h, cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking));
            CCX_CUDA(h, cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
            CCX_CUDA(h, cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
        }
        streams[1] = h->stream2;
        CCX_CUDA(h, cudaEventRecord(h->ev_fork, h->stream));
        CCX_CUDA(h, cudaStreamWaitEvent(h->stream2, h->ev_fork, 0));
    }
    // make sure the evaluator's internal scratch covers the whole batch before two streams share it
    if (h->net_mode == 1 && (rc = ccx_net_forward_tc_on(h, h->stream, n, 0, 0, planes, logits, value))) return rc;
    if (h->net_mode == 2 && (rc = ccx_net_forward_acc_on(h, h->stream, n, 0, 0, planes, logits, value, ccx_net_pold_bias(h)))) return rc;
    const int parts = split ? 2 : 1;
    // round r: [r == 0: select+encode | r > 0: finish round r-1's leaf, then select+encode] -> net; one last finish at the end.
    // Tried and dropped (r01c): programmatic dependent launch for the three kernels of a round (prologues before
    // griddepcontrol.wait, launch_dependents at kernel start) — 23.8 ms per 4,096-slot ply against 21.3 ms without, and
    // 94 ms against 73 ms at 16,384 slots: early-scheduled dependent CTAs take SM slots from the producer's later waves.
    for (int r = 0; r <= rounds; r++) {
        for (int q = 0; q < parts; q++) {
            const int64_t t0 = q ? part_n[0] : 0, nq = part_n[q];
            if (nq == 0) continue;
            cudaStream_t st = streams[q];
            const unsigned grid = tree_blocks(nq);
            const double *nz = root_noise;
            const bool ties = h->trees->tie_mode != 0;
            if (r == 0) {
                if (ties) k_mcts_select_encode<true><<<grid, 32 * MCTS_WARPS_PER_BLOCK, 0, st>>>(*h->trees, t0, nq, cpuct, planes);
                else k_mcts_select_encode<false><<<grid, 32 * MCTS_WARPS_PER_BLOCK, 0, st>>>(*h->trees, t0, nq, cpuct, planes);
            } else if (r < rounds) {
                if (ties) k_mcts_round<true><<<grid, 32 * MCTS_WARPS_PER_BLOCK, 0, st>>>(*h->trees, t0, nq, cpuct, logits, value, r == 1 ? nz : nullptr,
                                                                                       noise_stride, noise_normalize, planes, h->jump_table);
                else k_mcts_round<false><<<grid, 32 * MCTS_WARPS_PER_BLOCK, 0, st>>>(*h->trees, t0, nq, cpuct, logits, value, r == 1 ? nz : nullptr,
                                                                                     noise_stride, noise_normalize, planes, h->jump_table);
            }
            else
                k_mcts_softmax_expand_backup<<<grid, 32 * MCTS_WARPS_PER_BLOCK, 0, st>>>(*h->trees, t0, nq, logits, value,
                                                                                         rounds == 1 ? nz : nullptr, noise_stride,
                                                                                         noise_normalize, h->jump_table);
            CCX_LAUNCHED(h);
            if (r < rounds) {
                if (h->net_mode == 1) rc = ccx_net_forward_tc_on(h, st, n, t0, nq, planes + t0 * 343, logits + t0 * 294, value + t0);
                else if (h->net_mode == 2) rc = ccx_net_forward_acc_on(h, st, n, t0, nq, planes + t0 * 343, logits + t0 * 294, value + t0, ccx_net_pold_bias(h));
                else rc = ccx_net_forward_active(h, nq, planes + t0 * 343, logits + t0 * 294, value + t0);
                if (rc) return rc;
            }
        }
    }
    if (split) {
        CCX_CUDA(h, cudaEventRecord(h->ev_join, h->stream2));
        CCX_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_join, 0));
    }
    return CCX_OK;
}

// One search = rounds x (tree kernel, trunk, policy dense): ~530 dependent launches for 175 simulations.  The loop is the same
// from ply to ply (same buffers, same arguments), so it is captured once into a CUDA graph and replayed: 19.45 -> 18.25 ms per
// 4,096-tree search in the 16-bit mode and 33.4 -> 32.2 ms in the accurate mode on B200 (the gaps between dependent kernels
// shrink).  Life cycle per argument set: first call runs directly (and performs every lazy allocation), second call is
// captured on an internal stream and instantiated, later calls replay.  `epoch` + the by-value tree descriptor + the call's
// arguments are the cache key; CCX_NO_GRAPH=1 switches the replay off.
static bool round_graph_matches(const ccx_round_graph *g, const ccx_handle *h, int64_t n, int32_t rounds, double cpuct,
                                const double *noise, int32_t stride, int32_t normalize)
{
    return g->seen && g->epoch == h->epoch && g->n == n && g->rounds == rounds && g->cpuct == cpuct && g->noise == noise &&
           g->stride == stride && g->normalize == normalize && memcmp(&g->trees, h->trees, sizeof(ccx_trees)) == 0;
}

int ccx_mcts_run_net(ccx_handle *h, int64_t n, int32_t rounds, double cpuct, const double *root_noise, int32_t noise_stride,
                     int32_t noise_normalize)
{
    if (!h || !h->trees || n < 0 || n > h->trees->cap_trees || rounds < 0) return CCX_ERR_ARG;
    if (root_noise && noise_stride < 1) return CCX_ERR_ARG;
    if (n == 0 || rounds == 0) return CCX_OK;
    static const bool no_graph = getenv("CCX_NO_GRAPH") != nullptr;
    cudaStreamCaptureStatus user_capture = cudaStreamCaptureStatusNone;
    if (no_graph || h->graph_off || rounds < 8 || cudaStreamIsCapturing(h->stream, &user_capture) != cudaSuccess ||
        user_capture != cudaStreamCaptureStatusNone) {
        cudaGetLastError();
        return run_net_rounds(h, n, rounds, cpuct, root_noise, noise_stride, noise_normalize);
    }
    ccx_round_graph *g = (ccx_round_graph *)h->round_graph;
    if (!g) {
        g = new (std::nothrow) ccx_round_graph();
        if (!g) return CCX_ERR_NOMEM;
        h->round_graph = g;
    }
    if (!round_graph_matches(g, h, n, rounds, cpuct, root_noise, noise_stride, noise_normalize)) {
        // new argument set: run directly once (lazy allocations happen here), remember the key as it stands afterwards
        if (g->exec) { cudaGraphExecDestroy(g->exec); g->exec = nullptr; }
        g->seen = false;
        int rc = run_net_rounds(h, n, rounds, cpuct, root_noise, noise_stride, noise_normalize);
        if (rc) return rc;
        g->seen = true; g->epoch = h->epoch; g->n = n; g->rounds = rounds; g->cpuct = cpuct; g->noise = root_noise;
        g->stride = noise_stride; g->normalize = noise_normalize;
        memcpy(&g->trees, h->trees, sizeof(ccx_trees));
        return CCX_OK;
    }
    if (!g->exec) {
        if (!h->cap_stream) CCX_CUDA(h, cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));
        cudaStream_t user = h->stream;
        const int64_t before = h->launches;
        cudaGraph_t graph = nullptr;
        static const bool debug = getenv("CCX_GRAPH_DEBUG") != nullptr;
        cudaError_t ce = cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeThreadLocal);
        bool ok = ce == cudaSuccess;
        if (!ok && debug) fprintf(stderr, "ccx graph: begin capture: %s\n", cudaGetErrorString(ce));
        if (ok) {
            h->stream = h->cap_stream;
            int rc = run_net_rounds(h, n, rounds, cpuct, root_noise, noise_stride, noise_normalize);
            h->stream = user;
            ce = cudaStreamEndCapture(h->cap_stream, &graph);
            ok = ce == cudaSuccess && rc == CCX_OK && graph != nullptr;
            if (!ok && debug) fprintf(stderr, "ccx graph: capture rc %d, end capture: %s (%s)\n", rc, cudaGetErrorString(ce), h->cuda_err);
        }
        if (ok) {
            ce = cudaGraphInstantiate(&g->exec, graph, 0);
            ok = ce == cudaSuccess;
            if (!ok && debug) fprintf(stderr, "ccx graph: instantiate: %s\n", cudaGetErrorString(ce));
        }
        if (graph) cudaGraphDestroy(graph);
        g->launches = h->launches - before;
        h->launches = before;
        if (debug && ok && g->epoch != h->epoch) fprintf(stderr, "ccx graph: epoch moved during capture\n");
        if (!ok || g->epoch != h->epoch) {                 // capture refused or something was reallocated under it: no graphs on this handle
            cudaGetLastError();
            if (g->exec) { cudaGraphExecDestroy(g->exec); g->exec = nullptr; }
            g->seen = false;
            h->graph_off = 1;
            return run_net_rounds(h, n, rounds, cpuct, root_noise, noise_stride, noise_normalize);
        }
    }
    CCX_CUDA(h, cudaGraphLaunch(g->exec, h->stream));
    h->launches += g->launches;
    h->graph_replays++;
    return CCX_OK;
}

int ccx_mcts_finalize(ccx_handle *h, int64_t n, double tau, uint32_t *visits, double *pi, double *q, int32_t *n_nodes)
{
    if (!h || !h->trees || n < 0 || n > h->trees->cap_trees || !(tau > 0.0) || (n && !visits)) return CCX_ERR_ARG;
    if (n == 0) return CCX_OK;
    k_mcts_finalize<<<tree_blocks(n), 32 * MCTS_WARPS_PER_BLOCK, 0, h->stream>>>(*h->trees, n, 1.0 / tau, visits, pi, q, n_nodes);
    CCX_LAUNCHED(h);
    return CCX_OK;
}

int ccx_mcts_get_root(ccx_handle *h, int64_t n, int32_t stride, int32_t *n_edges, uint16_t *moves, uint32_t *N, double *W,
                      double *P)
{
    if (!h || !h->trees || n < 0 || n > h->trees->cap_trees || stride < 1 || (n && (!n_edges || !moves || !N || !W || !P)))
        return CCX_ERR_ARG;
    if (n == 0) return CCX_OK;
    k_mcts_get_root<<<tree_blocks(n), 32 * MCTS_WARPS_PER_BLOCK, 0, h->stream>>>(*h->trees, n, stride, n_edges, moves, N, W, P);
    CCX_LAUNCHED(h);
    return CCX_OK;
}

int ccx_mcts_set_root_priors(ccx_handle *h, int64_t n, int32_t stride, const double *P)
{
    if (!h || !h->trees || n < 0 || n > h->trees->cap_trees || stride < 1 || (n && !P)) return CCX_ERR_ARG;
    if (n == 0) return CCX_OK;
    k_mcts_set_root_priors<<<tree_blocks(n), 32 * MCTS_WARPS_PER_BLOCK, 0, h->stream>>>(*h->trees, n, stride, P);
    CCX_LAUNCHED(h);
    return CCX_OK;
}

int64_t ccx_mcts_pool_bytes(const ccx_handle *h) { return (h && h->trees) ? (int64_t)h->trees->bytes : 0; }

}  // extern "C"
