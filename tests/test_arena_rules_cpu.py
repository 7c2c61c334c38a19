"""Host-side rules of the arena mirrors (ai_vs_ai.py:45-52 55 % promotion rule, ai_vs_greedy.py:52-58, train.py:150-187
colour alternation) with the device part replaced by a stub.  CPU only."""
import pytest


class FakeArena:
    script = []

    def __init__(self, p1, p2, n, **kw):
        self.p1, self.p2, self.n, self.kw = p1, p2, n, kw
        FakeArena.calls.append(self)

    def play(self):
        w1, w2 = FakeArena.script.pop(0)
        return dict(p1_wins=w1, p2_wins=w2, stopped=self.n - w1 - w2, unfinished=0, games=self.n, plies=0)


@pytest.fixture()
def arena(monkeypatch):
    from chinesecheckersagent_b200 import arena as A
    FakeArena.calls = []
    monkeypatch.setattr(A, "BatchedArena", FakeArena)
    monkeypatch.setattr(A, "_load", lambda m, engine=None: m)
    return A


@pytest.mark.parametrize("w1,w2,expect", [(14, 10, "a"), (13, 11, None), (5, 14, "b"), (13, 13, None)])
def test_agent_match_55_percent_rule(arena, w1, w2, expect):
    FakeArena.script = [(w1, w2)]
    n = 24 if w1 + w2 <= 24 else 26
    assert arena.agent_match("a", "b", n) == expect          # winner needs more than int(0.55 * n) wins (ai_vs_ai.py:47-50)
    assert FakeArena.calls[0].p1 == "a" and FakeArena.calls[0].p2 == "b"


def test_evaluate_alternates_colours_and_enforces_move_limit(arena):
    FakeArena.script = [(7, 4), (6, 5)]                      # (best as P1: 7-4), (cur as P1: 6-5)
    cur, best, draws = arena.evaluate("best", "cur", num_games=24)
    assert (cur, best, draws) == (4 + 6, 7 + 5, 2)
    a, b = FakeArena.calls
    assert (a.p1, a.p2, a.n) == ("best", "cur", 12) and (b.p1, b.p2, b.n) == ("cur", "best", 12)
    assert a.kw["enforce_move_limit"] and b.kw["enforce_move_limit"] and b.kw["game_id0"] == 12


def test_agent_greedy_match_swaps_sides(arena):
    FakeArena.script = [(3, 2), (1, 4)]                      # model as P1: 3-2; greedy as P1: 1-4 (model wins 4)
    assert arena.agent_greedy_match("m", 10) == "m"
    a, b = FakeArena.calls
    assert (a.p1, a.p2) == ("m", arena.GREEDY) and (b.p1, b.p2) == (arena.GREEDY, "m")
    FakeArena.script = [(1, 4), (4, 1)]
    assert arena.agent_greedy_match("m", 10) == arena.GREEDY
