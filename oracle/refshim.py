"""Import shim that makes the UNMODIFIED reference importable in the build container.

TEST INFRASTRUCTURE ONLY.  This module is used by ``tests/golden/gen_golden.py`` to run the real
reference (``/root/reference``) and write golden fixtures; and by ``oracle/refrun.py``
(bench.py's CPU-baseline legs, the drop-in tests) to run the copy staged as ``oracle/_ref/reference.zip``
(git-ignored; ``/root/reference`` itself does not exist on the GPU box).  Nothing in the product package
imports it.

Why a shim is needed: ``utils.py:3`` imports h5py, ``model.py:3-7`` imports keras,
``loss.py:1`` imports tensorflow (none installed, no network) and ``utils.py:7`` imports
``collections.Mapping`` (removed in Python 3.10).  The stubs below are inert; the Keras net is never
instantiated through them.
"""
import collections
import collections.abc
import os
import sys
import types
import warnings

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "reference.zip")     # refrun.stage(): travels with gpurun
REFERENCE_DIR = os.environ.get("CCX_REFERENCE_DIR") or ("/root/reference" if os.path.isdir("/root/reference") else _STAGED)


class _Inert:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return self

    def __getattr__(self, n):
        return _Inert()


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_installed = False


def install():
    """Idempotently install the stubs and put the reference on sys.path."""
    global _installed
    if _installed:
        return
    if not os.path.exists(REFERENCE_DIR):
        raise RuntimeError("reference checkout not found at %s (golden generation only runs in the "
                           "build container)" % REFERENCE_DIR)
    warnings.filterwarnings("ignore")            # utils.py:13 has an invalid escape sequence
    sys.dont_write_bytecode = True               # the reference directory is read-only
    sys.path.insert(0, REFERENCE_DIR)
    collections.Mapping = collections.abc.Mapping        # utils.py:7
    collections.Container = collections.abc.Container
    _stub("h5py")                                        # utils.py:3
    _stub("tensorflow").nn = _Inert()                    # loss.py:1
    _stub("keras", regularizers=_Inert())                # model.py:3
    _stub("keras.optimizers", SGD=_Inert, Adam=_Inert)   # model.py:4
    _stub("keras.models", load_model=_Inert(), Model=_Inert)   # model.py:5-6
    _stub("keras.layers", **{k: _Inert for k in
          "Input Conv2D Flatten Dense BatchNormalization LeakyReLU Activation add".split()})  # model.py:7
    _installed = True


class FirstChoice:
    """Stand-in for the ``random`` module *inside MCTS.py only*: ``choice(seq) -> seq[0]``.

    ``MCTS.py:65-69`` resets ``chosen_edges`` on a strictly larger QU, so element 0 is always the first
    edge attaining the exact maximum; the batched engine implements first-index argmax.
    """
    @staticmethod
    def choice(seq):
        return seq[0]


def load():
    """Return a namespace with the reference modules (board, utils, MCTS, player, game, selfplay)."""
    install()
    import board as ref_board
    import board_utils as ref_board_utils
    import config as ref_config
    import utils as ref_utils
    import MCTS as ref_mcts
    import player as ref_player
    import game as ref_game
    import selfplay as ref_selfplay
    import data_generators as ref_datagen
    return types.SimpleNamespace(board=ref_board, board_utils=ref_board_utils, config=ref_config,
                                 utils=ref_utils, MCTS=ref_mcts, player=ref_player, game=ref_game,
                                 selfplay=ref_selfplay, data_generators=ref_datagen)
