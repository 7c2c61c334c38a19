// hostcheck.cu — TEST-ONLY host instantiation of the per-game device routines in
// chinesecheckersagent_b200/csrc/ccx_device.cuh.  It lets the CPU test-suite (no GPU in the build
// container) run the exact bitboard source the kernels run and compare it with the oracle.
// It is never linked into libccx.so and nothing in the product can reach it.
#include "../../chinesecheckersagent_b200/csrc/ccx_device.cuh"
#include "../../include/ccx.h"

static Game load_game_h(const u64 *st, int64_t n, int64_t i)
{
    u64 occ1 = st[0 * n + i], occ2 = st[1 * n + i], c1 = st[2 * n + i], c2 = st[3 * n + i];
    Game g;
    g.meta = st[4 * n + i];
    bool p2 = (g.meta >> 48) & 1;
    g.occ_me = p2 ? occ2 : occ1; g.occ_op = p2 ? occ1 : occ2;
    g.cells_me = p2 ? c2 : c1;   g.cells_op = p2 ? c1 : c2;
    return g;
}
static void store_game_h(u64 *st, int64_t n, int64_t i, const Game &g)
{
    bool p2 = (g.meta >> 48) & 1;
    st[0 * n + i] = p2 ? g.occ_op : g.occ_me; st[1 * n + i] = p2 ? g.occ_me : g.occ_op;
    st[2 * n + i] = p2 ? g.cells_op : g.cells_me; st[3 * n + i] = p2 ? g.cells_me : g.cells_op;
    st[4 * n + i] = g.meta;
}

extern "C" {

void hc_movegen(const u64 *st, int64_t n, u64 *masks)
{
    for (int64_t i = 0; i < n; i++) {
        Game g = load_game_h(st, n, i);
        u64 dest[6];
        movegen(g.occ_me | g.occ_op, g.cells_me, dest);
        for (int k = 0; k < 6; k++) masks[k * n + i] = dest[k];
    }
}

void hc_movegen_rays(const u64 *st, int64_t n, u64 *masks)
{
    static uint8_t T[CCX_JT_BYTES];
    build_jump_table(T, 0, 1);
    for (int64_t i = 0; i < n; i++) {
        Game g = load_game_h(st, n, i);
        u64 dest[6];
        movegen_rays(g.occ_me | g.occ_op, g.cells_me, dest, T);
        for (int k = 0; k < 6; k++) masks[k * n + i] = dest[k];
    }
}

// movegen with the three-layout expansion (expand_cell_tri / the layout-2 table), the way k_step_random_tri maintains the copies:
// occT / occD of the whole board built from the checker cells, the mover lifted in each layout
void hc_movegen_tri(const u64 *st, int64_t n, u64 *masks, int use_lut2)
{
    static uint8_t T3[CCX_JT3_BYTES] __attribute__((aligned(16)));
    static uint8_t T2[CCX_JT2_BYTES];
    static u64 CI3[64]; static u32 CI2[64];
    build_jump_table3(T3, 0, 1);
    build_jump_table2(T2, 0, 1);
    for (int c = 0; c < 64; c++) { CI3[c] = ((CCX_VALID >> c) & 1) ? tri_cell_info(c) : 0; CI2[c] = cell_diag_info2(c); }
    for (int64_t i = 0; i < n; i++) {
        Game g = load_game_h(st, n, i);
        u64 occ_all = g.occ_me | g.occ_op, occT_all = 0, occD_all = 0;
        for (int c = 0; c < 64; c++) if ((occ_all >> c) & 1) { occT_all |= 1ULL << tri_tbit(c); occD_all |= 1ULL << tri_dbit(c); }
        for (int k = 0; k < 6; k++) {
            int cell = (int)((g.cells_me >> (8 * k)) & 0xFF);
            u64 o = 1ULL << cell, occ = occ_all & ~o;
            u64 occT = occT_all & ~(1ULL << tri_tbit(cell)), occD = occD_all & ~(1ULL << tri_dbit(cell));
            u64 todo = o, reach = 0;
            while (todo) {
                int c = 63 - __builtin_clzll(todo);
                todo ^= 1ULL << c;
                u64 nw = (use_lut2 ? expand_cell_lut2(c, occ, T2, CI2) : expand_cell_tri(c, occ, occT, occD, T3, CI3)) & ~(reach | o);
                reach |= nw; todo |= nw;
            }
            masks[k * n + i] = (neighbours(o) & ~occ & CCX_VALID) | reach;
        }
    }
}

void hc_greedy(const u64 *st, int64_t n, u64 *masks)
{
    for (int64_t i = 0; i < n; i++) {
        Game g = load_game_h(st, n, i);
        u64 dest[6], cand[6];
        movegen(g.occ_me | g.occ_op, g.cells_me, dest);
        greedy_candidates(g, dest, cand);
        for (int k = 0; k < 6; k++) masks[k * n + i] = cand[k];
    }
}

void hc_apply(u64 *st, int64_t n, const uint8_t *from, const uint8_t *to, uint8_t *winner)
{
    for (int64_t i = 0; i < n; i++) {
        Game g = load_game_h(st, n, i);
        int id = 0;
        for (int k = 5; k >= 0; k--) if (((g.cells_me >> (8 * k)) & 0xFF) == (u64)from[i]) id = k;
        apply_move(g, id, from[i], to[i]);
        store_game_h(st, n, i, g);
        u64 lo = st[5 * n + i], hi = st[6 * n + i];
        push_hist(lo, hi, to[i]);
        st[5 * n + i] = lo; st[6 * n + i] = hi;
        winner[i] = (uint8_t)winner_of(g);
    }
}

void hc_step_random(u64 *st, int64_t n, int64_t gid0, uint64_t seed, uint32_t step0, int plies, u64 *wins,
                    u64 *trace, int64_t trace_games)
{
    for (int64_t i = 0; i < n; i++) {
        Game g = load_game_h(st, n, i);
        u64 gid = (u64)(gid0 + i);
        for (int t = 0; t < plies; t++) {
            u64 dest[6];
            movegen(g.occ_me | g.occ_op, g.cells_me, dest);
            u32 nonempty = 0;
            for (int k = 0; k < 6; k++) nonempty += dest[k] != 0;
            u64 *row = nullptr;
            if (trace && i < trace_games) {
                row = trace + ((int64_t)t * trace_games + i) * CCX_TRACE_WORDS;
                bool p2 = (g.meta >> 48) & 1;
                row[0] = p2 ? g.occ_op : g.occ_me; row[1] = p2 ? g.occ_me : g.occ_op;
                row[2] = p2 ? g.cells_op : g.cells_me; row[3] = p2 ? g.cells_me : g.cells_op;
                row[4] = g.meta & 0x00FFFFFFFFFFFFFFULL;
                for (int k = 0; k < 6; k++) row[5 + k] = dest[k];
                row[11] = 0xFFULL | (0xFFULL << 8) | (0xFFULL << 24);
            }
            if (!nonempty) continue;
            Philox4 r = philox4x32_10((u32)seed, (u32)(seed >> 32), step0 + (u32)t, 0u, (u32)gid, (u32)(gid >> 32));
            int from, to;
            int id = pick_random(g, dest, nonempty, r.x, r.y, from, to);
            apply_move(g, id, from, to);
            int win = winner_of(g);
            if (row) row[11] = (u64)from | ((u64)to << 8) | ((u64)win << 16) | ((u64)id << 24);
            if (win) { wins[win - 1]++; reset_start(g); }
        }
        store_game_h(st, n, i, g);
    }
}

void hc_play_greedy(u64 *st, int64_t n, int64_t gid0, uint64_t seed, int max_plies)
{
    for (int64_t i = 0; i < n; i++) {
        Game g = load_game_h(st, n, i);
        u64 lo = st[5 * n + i], hi = st[6 * n + i];
        u64 gid = (u64)(gid0 + i);
        int status = (int)(g.meta >> 56);
        for (int t = 0; status == CCX_ST_RUNNING && t < max_plies; t++) {
            u64 dest[6], cand[6];
            movegen(g.occ_me | g.occ_op, g.cells_me, dest);
            int total = greedy_candidates(g, dest, cand);
            if (total == 0) { status = CCX_ST_NO_MOVES; break; }
            u32 ply = (u32)((g.meta >> 32) & 0xFFFF);
            Philox4 r = philox4x32_10((u32)seed, (u32)(seed >> 32), ply, 1u, (u32)gid, (u32)(gid >> 32));
            int from, to;
            int id = pick_candidate(g, cand, total, r.x, from, to);
            apply_move(g, id, from, to);
            push_hist(lo, hi, to);
            int win = winner_of(g);
            if (win) { status = win; break; }
            if (((g.meta >> 32) & 0xFFFF) >= 16 && repetition_stop(lo, hi)) status = CCX_ST_REPETITION;
        }
        g.meta = (g.meta & 0x00FFFFFFFFFFFFFFULL) | ((u64)status << 56);
        store_game_h(st, n, i, g);
        st[5 * n + i] = lo; st[6 * n + i] = hi;
    }
}

}  // extern "C"
