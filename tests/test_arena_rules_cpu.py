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


def test_game_referee_matches_rule_statement():
    """_Referee against a direct deque statement of game.py:70-92 on random destination streams."""
    import random
    from collections import deque
    from chinesecheckersagent_b200.config import PROGRESS_MOVE_LIMIT, TOTAL_HIST_MOVES, UNIQUE_DEST_LIMIT
    from chinesecheckersagent_b200.game import _Referee
    rng = random.Random(5)
    for trial in range(200):
        capped = bool(trial & 1)
        pool = [(rng.randrange(7), rng.randrange(7)) for _ in range(rng.choice([2, 3, 4, 6, 12]))]
        ref, hist, plies = _Referee(capped), deque(maxlen=TOTAL_HIST_MOVES), 0
        for _ in range(PROGRESS_MOVE_LIMIT + 5):
            d = rng.choice(pool)
            hist.append(d)
            own = {hist[i] for i in range(len(hist) - 1, -1, -2)}
            want = None
            if len(hist) == TOTAL_HIST_MOVES and len(own) <= UNIQUE_DEST_LIMIT:
                want = "rep"
            else:
                plies += 1
                if capped and plies >= PROGRESS_MOVE_LIMIT:
                    want = "cap"
            got = ref.verdict(d)
            assert (got is None) == (want is None)
            if want:
                assert ("Repetition" in got) == (want == "rep")
                break


def test_board_utils_round_trip_and_tips():
    from chinesecheckersagent_b200 import board_utils as bu
    seen = set()
    for i in range(7):
        for j in range(7):
            h = bu.np_index_to_human_coord((i, j))
            assert bu.human_coord_to_np_index(h) == (i, j)
            assert 1 <= h[0] <= 13 and 1 <= h[1] <= 7 - abs(h[0] - 7)
            seen.add(h)
    assert len(seen) == 49
    assert bu.np_index_to_human_coord((6, 0)) == (13, 1) and bu.np_index_to_human_coord((0, 6)) == (1, 1)
    assert bu.np_index_to_human_coord((0, 0)) == (7, 1) and bu.np_index_to_human_coord((6, 6)) == (7, 7)
    assert bu.is_valid_pos(6, 6) and not bu.is_valid_pos(7, 0) and not bu.is_valid_pos(0, -1)
    assert bu.convert_np_to_human_moves({(6, 0): [(5, 0), (6, 1)]}) == {(13, 1): [(12, 1), (12, 2)]}
