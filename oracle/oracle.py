"""ctypes binding of the CPU oracle (oracle/ccx_oracle*.c) — TEST INFRASTRUCTURE, NOT PRODUCT.

Importers: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference leg.
The product package must never import this module.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libccx_oracle.so")

STATE_WORDS = 8
NACT = 294
TRACE_WORDS = 12
_u64p = ctypes.POINTER(ctypes.c_uint64)


def build(force=False):
    """Compile the oracle with gcc (seconds) and, where the read-only reference checkout exists (build container), stage its
    .py files as oracle/_ref/reference.zip so that they travel to the GPU box (refrun.stage; git-ignored)."""
    try:
        import refrun
        refrun.stage()
    except Exception as e:                      # staging is best effort: the port below is always available
        print("oracle: reference staging skipped: %s" % e)
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in srcs)):
        return _LIB_PATH
    subprocess.check_call(["make", "-s", "-C", _HERE, "-B"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        i64, i32, u64, u32, dbl, vp = (ctypes.c_int64, ctypes.c_int32, ctypes.c_uint64, ctypes.c_uint32,
                                       ctypes.c_double, ctypes.c_void_p)
        L.orc_movegen_batch.argtypes = [vp, i64, vp]
        L.orc_encode_batch.argtypes = [vp, i64, vp]
        L.orc_greedy_batch.argtypes = [vp, i64, vp]
        L.orc_step_random.argtypes = [vp, i64, i64, u64, u32, i32, vp, vp, i64, i32]
        L.orc_play_greedy.argtypes = [vp, i64, i64, u64, i32, i32]
        L.orc_philox.argtypes = [u32] * 6 + [vp]
        L.orc_movelist_batch.argtypes = [vp, i64, vp, vp]
        L.orc_apply_batch.argtypes = [vp, i64, vp, vp, vp]
        L.orc_info_batch.argtypes = [vp, i64, vp]
        L.orc_greedy_list_batch.argtypes = [vp, i64, vp, vp]
        L.orc_mcts_stub_batch.argtypes = [vp, i64, i32, dbl, dbl, i32, vp, vp, vp, i32]
        L.orc_mcts_batch.argtypes = [vp, i64, i32, dbl, dbl, i32, i32, vp, i32, vp, vp, vp, vp, i32]
        L.orc_mcts_batch_ties.argtypes = [vp, i64, i32, dbl, dbl, i32, i32, vp, i32, vp, vp, vp, vp, i32, u64, i64]
        L.orc_mcts_batch_ties.restype = None
        L.orc_mcts_stub_batch.restype = None
        L.orc_mcts_batch.restype = None
        for name in ("orc_movegen_batch", "orc_encode_batch", "orc_greedy_batch", "orc_step_random",
                     "orc_play_greedy", "orc_philox"):
            getattr(L, name).restype = None
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


# ---------------------------------------------------------------------------------------------
# packed state helpers (numpy, host side).  Layout: include/ccx.h "State layout".

def cell(r, c):
    return 8 * int(r) + int(c)


def cell_rc(b):
    return int(b) >> 3, int(b) & 7


START_P1 = [(6, 0), (5, 0), (6, 1), (4, 0), (5, 1), (6, 2)]   # board.py:43-44
START_P2 = [(0, 6), (1, 6), (0, 5), (2, 6), (1, 5), (0, 4)]   # board.py:45-46


def pack_state(p1, p2, to_move=0, ply=0, last_moves=(), hist_dests=(), status=0):
    """p1/p2: six (r,c) per player in id order; last_moves: [(from_rc,to_rc)] most recent first (<=2);
    hist_dests: destinations (r,c) most recent first (<=16).  Returns uint64[8]."""
    w = np.zeros(STATE_WORDS, dtype=np.uint64)
    occ = [0, 0]
    cells = [0, 0]
    for pl, pos in enumerate((p1, p2)):
        for i, (r, c) in enumerate(pos):
            occ[pl] |= 1 << cell(r, c)
            cells[pl] |= cell(r, c) << (8 * i)
    meta = 0
    for k in range(2):
        if k < len(last_moves):
            f, t = last_moves[k]
            meta |= cell(*f) << (16 * k) | cell(*t) << (16 * k + 8)
        else:
            meta |= 0xFFFF << (16 * k)
    meta |= (ply & 0xFFFF) << 32 | (to_move & 0xFF) << 48 | (status & 0xFF) << 56
    hist = [(1 << 64) - 1, (1 << 64) - 1]
    for k, d in enumerate(hist_dests[:16]):
        hist[k >> 3] &= ~(0xFF << (8 * (k & 7)))
        hist[k >> 3] |= cell(*d) << (8 * (k & 7))
    w[:7] = [occ[0], occ[1], cells[0], cells[1], meta, hist[0], hist[1]]
    return w


def start_states(n):
    w = pack_state(START_P1, START_P2)
    return np.ascontiguousarray(np.repeat(w[:, None], n, axis=1))


def random_states(n, seed):
    """Board(randomised=True)-style placements (board.py:61-85): 12 distinct cells, first 6 -> P1."""
    rng = np.random.default_rng(seed)
    st = np.zeros((STATE_WORDS, n), dtype=np.uint64)
    for i in range(n):
        chosen = rng.choice(49, size=12, replace=False)
        pos = [(int(x) // 7, int(x) % 7) for x in chosen]
        st[:, i] = pack_state(pos[:6], pos[6:], to_move=int(rng.integers(0, 2)))
    return st


def unpack_meta(meta):
    meta = int(meta)
    return dict(last_from=meta & 0xFF, last_to=(meta >> 8) & 0xFF, prev_from=(meta >> 16) & 0xFF,
                prev_to=(meta >> 24) & 0xFF, ply=(meta >> 32) & 0xFFFF, to_move=(meta >> 48) & 0xFF,
                status=(meta >> 56) & 0xFF)


# ---------------------------------------------------------------------------------------------
# batched oracle calls

def movegen(st):
    st = np.ascontiguousarray(st, dtype=np.uint64)
    n = st.shape[1]
    masks = np.zeros((6, n), dtype=np.uint64)
    lib().orc_movegen_batch(_ptr(st), n, _ptr(masks))
    return masks


def encode(st):
    st = np.ascontiguousarray(st, dtype=np.uint64)
    n = st.shape[1]
    out = np.zeros((n, 7, 7, 7), dtype=np.uint8)
    lib().orc_encode_batch(_ptr(st), n, _ptr(out))
    return out


def greedy_candidates(st):
    st = np.ascontiguousarray(st, dtype=np.uint64)
    n = st.shape[1]
    masks = np.zeros((6, n), dtype=np.uint64)
    lib().orc_greedy_batch(_ptr(st), n, _ptr(masks))
    return masks


def movelist(st):
    st = np.ascontiguousarray(st, dtype=np.uint64)
    n = st.shape[1]
    out = np.zeros((n, 6, 24), dtype=np.int8)
    cnt = np.zeros((n, 6), dtype=np.int8)
    lib().orc_movelist_batch(_ptr(st), n, _ptr(out), _ptr(cnt))
    return out, cnt


def apply(st, frm, to):
    st = np.array(st, dtype=np.uint64, order="C")
    n = st.shape[1]
    frm = np.ascontiguousarray(frm, dtype=np.uint8)
    to = np.ascontiguousarray(to, dtype=np.uint8)
    winner = np.zeros(n, dtype=np.uint8)
    lib().orc_apply_batch(_ptr(st), n, _ptr(frm), _ptr(to), _ptr(winner))
    return st, winner


def info(st):
    st = np.ascontiguousarray(st, dtype=np.uint64)
    n = st.shape[1]
    out = np.zeros((n, 5), dtype=np.int16)
    lib().orc_info_batch(_ptr(st), n, _ptr(out))
    return out


def greedy_list(st):
    st = np.ascontiguousarray(st, dtype=np.uint64)
    n = st.shape[1]
    out = np.zeros((n, 32, 2), dtype=np.int16)
    cnt = np.zeros(n, dtype=np.int16)
    lib().orc_greedy_list_batch(_ptr(st), n, _ptr(out), _ptr(cnt))
    return out, cnt


def step_random(st, seed, step0, plies, game_id0=0, trace_games=0, nthreads=1):
    """In-place on a copy; returns (new_state, wins[2], trace[plies, trace_games, 12] or None)."""
    st = np.array(st, dtype=np.uint64, order="C")
    n = st.shape[1]
    wins = np.zeros(2, dtype=np.uint64)
    trace = np.zeros((plies, trace_games, TRACE_WORDS), dtype=np.uint64) if trace_games else None
    lib().orc_step_random(_ptr(st), n, game_id0, seed, step0, plies, _ptr(wins),
                          _ptr(trace) if trace is not None else None, trace_games, nthreads)
    return st, wins, trace


def play_greedy(st, seed, max_plies=100000, game_id0=0, nthreads=1):
    st = np.array(st, dtype=np.uint64, order="C")
    lib().orc_play_greedy(_ptr(st), st.shape[1], game_id0, seed, max_plies, nthreads)
    return st


def philox(k0, k1, c0, c1, c2, c3):
    out = np.zeros(4, dtype=np.uint32)
    lib().orc_philox(k0, k1, c0, c1, c2, c3, _ptr(out))
    return out


EVAL_UNIFORM, EVAL_HASH = 0, 1


def mcts(st, num_itr=175, cpuct=3.5, tau=1.0, pre_expand=0, evaluator=EVAL_UNIFORM, noise=None, nthreads=1, ties=None):
    """MCTS.search on every root (canonical edge order).  ties=None: first-maximum tie-break; ties=(seed, uid0): the
    reference's epsilon-tie list (MCTS.py:65-72) with the engine's Philox draw (ccx_mcts_set_tiebreak mode 1).
    Returns (visits[n,294] u32, pi[n,294] f64, root Q[n,294] f64, node count[n])."""
    st = np.ascontiguousarray(st, dtype=np.uint64)
    n = st.shape[1]
    visits = np.zeros((n, NACT), dtype=np.uint32)
    pi = np.zeros((n, NACT), dtype=np.float64)
    q = np.zeros((n, NACT), dtype=np.float64)
    nodes = np.zeros(n, dtype=np.int32)
    stride = 0
    if noise is not None:
        noise = np.ascontiguousarray(noise, dtype=np.float64)
        stride = noise.shape[1]
    if ties is not None:
        lib().orc_mcts_batch_ties(_ptr(st), n, num_itr, cpuct, tau, pre_expand, evaluator,
                                  _ptr(noise) if noise is not None else None, stride, _ptr(visits), _ptr(pi), _ptr(nodes),
                                  _ptr(q), nthreads, int(ties[0]), int(ties[1]))
        return visits, pi, q, nodes
    lib().orc_mcts_batch(_ptr(st), n, num_itr, cpuct, tau, pre_expand, evaluator,
                         _ptr(noise) if noise is not None else None, stride, _ptr(visits), _ptr(pi), _ptr(nodes),
                         _ptr(q), nthreads)
    return visits, pi, q, nodes
