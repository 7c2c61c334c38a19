"""Mirror of the reference's MCTS.py object surface (Node, Edge, MCTS) on top of the device tree pools.

`MCTS(root, model, cpuct, num_itr, tree_tau)`, `.expandAndBackUp(root, [])`, `.search() -> (pi, sampled_edge)`
behave like MCTS.py:40-153 for the way selfplay.py:114-133 and player.py:157-166 use them: the root's edges
are materialised as Edge objects (stats N/W/Q/P, fromPos/toPos, outNode) so callers can mix noise into
stats['P'] before search() and walk to `sampled_edge.outNode` afterwards.  The tree itself lives on the GPU;
`model` is any object with `.predict(x (7,7,7)) -> (p[294], v)` (MCTS.py:93) — a ResidualCNN from this
package is evaluated on the device without leaving it.  Ties are broken towards the first maximal edge."""
import copy
import ctypes

import numpy as np
import torch

from . import utils
from .config import BOARD_HEIGHT, BOARD_WIDTH, C_PUCT, DTYPE_U8, MCTS_SIMULATIONS, NUM_CHECKERS, PLAYER_ONE, PLAYER_TWO, TREE_TAU
from .engine import BatchedEnv, _p

ROOT_STRIDE = 128


class Node:
    def __init__(self, state, currPlayer):
        self.state = state
        self.currPlayer = currPlayer
        self.edges = []
        self.pi = np.zeros(NUM_CHECKERS * BOARD_WIDTH * BOARD_HEIGHT, dtype='float64')

    def isLeaf(self):
        return len(self.edges) == 0


class Edge:
    def __init__(self, inNode, outNode, prior, fromPos, toPos):
        self.inNode, self.outNode = inNode, outNode
        self.currPlayer = inNode.currPlayer
        self.fromPos, self.toPos = fromPos, toPos
        self.stats = {'N': 0, 'W': 0, 'Q': 0, 'P': prior}


class MCTS:
    def __init__(self, root, model, cpuct=C_PUCT, num_itr=MCTS_SIMULATIONS, tree_tau=TREE_TAU):
        self.root, self.model, self.cpuct, self.num_itr, self.tree_tau = root, model, cpuct, num_itr, tree_tau
        self._eng = root.state._eng
        self._begun = False
        self._leaf = self._eng.empty((5, 1), torch.int64)

    # -- device plumbing -----------------------------------------------------------------------------------
    def _begin(self):
        st = torch.from_numpy(self.root.state._pack(self.root.currPlayer - 1).view(np.int64)).to(self._eng.device)
        self._eng.call("ccx_mcts_begin", 1, _p(st), self.num_itr + 2, 0, -1)
        self._begun = True

    def _evaluate(self):
        """Model.predict on the selected leaf (MCTS.py:93)."""
        if hasattr(self.model, "evaluate_states"):
            return self.model.evaluate_states(self._leaf)
        full = torch.zeros((8, 1), dtype=torch.int64, device=self._eng.device)
        full[:5] = self._leaf
        x = BatchedEnv(1, engine=self._eng, state=full).encode(DTYPE_U8)[0].cpu().numpy().astype(np.float64)
        p, v = self.model.predict(x)
        p = torch.from_numpy(np.ascontiguousarray(np.asarray(p, dtype=np.float64)).reshape(1, -1)).to(self._eng.device)
        v = torch.tensor([float(v)], dtype=torch.float64, device=self._eng.device)
        return p, v

    def _simulate(self):
        self._eng.call("ccx_mcts_select", 1, float(self.cpuct), _p(self._leaf))
        p, v = self._evaluate()
        self._eng.call("ccx_mcts_expand_backup", 1, _p(p), _p(v), None, 0, 0)

    def _pull_root(self):
        e = self._eng
        ne = e.empty((1,), torch.int32)
        mv = e.empty((1, ROOT_STRIDE), torch.int16)
        N = e.empty((1, ROOT_STRIDE), torch.int32)
        W = e.empty((1, ROOT_STRIDE), torch.float64)
        P = e.empty((1, ROOT_STRIDE), torch.float64)
        e.call("ccx_mcts_get_root", 1, ROOT_STRIDE, _p(ne), _p(mv), _p(N), _p(W), _p(P))
        k = int(ne.item())
        return k, mv[0, :k].cpu().numpy().astype(np.int64) & 0xFFFF, N[0, :k].cpu().numpy(), W[0, :k].cpu().numpy(), P[0, :k].cpu().numpy()

    def _materialise_root_edges(self):
        root = self.root
        k, mv, N, W, P = self._pull_root()
        if not root.edges:
            nxt = PLAYER_ONE + PLAYER_TWO - root.currPlayer
            for j in range(k):
                cid, to = int(mv[j]) >> 8, int(mv[j]) & 0xFF
                frm, dst = root.state.checkers_pos[root.currPlayer][cid], (to >> 3, to & 7)
                child = copy.deepcopy(root.state)
                child.place(root.currPlayer, frm, dst)                          # MCTS.py:104-105
                root.edges.append(Edge(root, Node(child, nxt), float(P[j]), frm, dst))
        for j, e in enumerate(root.edges):
            e.stats['N'] = int(N[j]); e.stats['W'] = float(W[j])
            e.stats['Q'] = float(W[j]) / int(N[j]) if N[j] else 0
            e.stats['P'] = float(P[j])

    # -- MCTS.py API -----------------------------------------------------------------------------------------
    def expandAndBackUp(self, leafNode, breadcrumbs):
        """Only the root expansion of selfplay.py:117 is driven from Python; expansions inside search()
        happen on the device."""
        assert leafNode is self.root and not breadcrumbs and leafNode.isLeaf()
        self._begin()
        self._simulate()
        self._materialise_root_edges()

    def search(self):
        if not self._begun:
            self._begin()
        elif self.root.edges:
            # callers may have edited stats['P'] (Dirichlet noise, selfplay.py:121-124): push it to the device
            P = np.zeros((1, ROOT_STRIDE), dtype=np.float64)
            P[0, :len(self.root.edges)] = [e.stats['P'] for e in self.root.edges]
            Pd = torch.from_numpy(P).to(self._eng.device)
            self._eng.call("ccx_mcts_set_root_priors", 1, ROOT_STRIDE, _p(Pd))
        for _ in range(self.num_itr):                                           # MCTS.py:123-125
            self._simulate()
        self._materialise_root_edges()
        root = self.root
        for edge in root.edges:                                                 # MCTS.py:131-137
            cid = root.state.checkers_id[root.currPlayer][edge.fromPos]
            root.pi[utils.encode_checker_index(cid, edge.toPos)] = pow(edge.stats['N'], 1. / self.tree_tau)
        root.pi /= np.sum(root.pi)
        sampled_index = np.random.choice(np.arange(len(root.pi)), p=root.pi)    # MCTS.py:140
        cid, sampled_to = utils.decode_checker_index(sampled_index)
        sampled_from = root.state.checkers_pos[root.currPlayer][cid]
        sampled_edge = None
        for edge in root.edges:
            if edge.fromPos == sampled_from and edge.toPos == sampled_to:
                sampled_edge = edge
                break
        assert sampled_edge is not None
        return root.pi, sampled_edge
