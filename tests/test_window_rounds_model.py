"""The round rule of the window-search kernels (csrc/frame.cu), restated on the CPU (tools/model_window_rounds.py), against the
oracle's sequential SearchByProjection: host logic only, no GPU."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_round_rule_equals_sequential_walk(oracle, synth):
    import model_window_rounds as model
    import test_frame as tf
    rng = np.random.default_rng(11)
    most = 0
    for k in range(14):
        # every other scene is the hard kind: crowded, windows that span a large part of the frame
        ok, rounds, ctx = model.check_scene(oracle, synth, tf, rng, max_features=500, max_points=1500,
                                            th=15.0 if k % 2 else None, crowded=True if k % 2 else None)
        assert ok, ctx
        most = max(most, rounds)
    assert most >= 4      # the scenes really needed several rounds


def test_plain_taken_flag_rule_is_wrong_somewhere(oracle, synth):
    """The first form of the rule (a feature only knows THAT it was taken, not by whom) lets a later point that became final early
    hide a feature from an earlier unresolved point: the model must be able to tell the two rules apart."""
    import model_window_rounds as model
    import test_frame as tf
    rng = np.random.default_rng(5)
    wrong = 0
    for _ in range(80):
        ok, _, _ = model.check_scene(oracle, synth, tf, rng, crowded=True, flag_rule=True)
        wrong += not ok
        if wrong:
            break
    assert wrong, "80 crowded scenes and the taken-flag rule never differed from the oracle: the model lost its teeth"
