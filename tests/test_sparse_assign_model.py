"""The algorithm of the sparse certified solver (csrc/lapjv.cuh: sparse_assign_warp + arcs_unique), checked on
the CPU through its statement-by-statement model in tests/design/sparse_assign_model.py: on crowds without
ties it must return exactly the reference solver's assignment and certify it; with an equal-cost alternative it
must refuse (the kernel then runs the literal JV emulation).  The CUDA code itself is compared with the oracle
in tests/test_gpu_parity.py::test_lapjv_component_path_matches_oracle."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'design'))
from sparse_assign_model import sparse_assign  # noqa: E402

from oracle.lapjv import lapjv  # noqa: E402


def crowd(rng, R, C, ties=False):
    cost = rng.uniform(0.95, 1.3, (R, C)).astype(np.float32)
    n = min(R, C)
    perm, rows = rng.permutation(C)[:n], rng.permutation(R)[:n]
    for k in range(n):
        if rng.random() < 0.9:
            cost[rows[k], perm[k]] = rng.uniform(0.05, 0.6)
        if k + 1 < n and rng.random() < 0.6:
            cost[rows[k], perm[k + 1]] = rng.uniform(0.3, 0.89)
        if k > 0 and rng.random() < 0.3:
            cost[rows[k], perm[k - 1]] = rng.uniform(0.3, 0.89)
    if ties:
        cost[np.ix_(rows[:2], perm[:2])] = 0.5
    return cost


@pytest.mark.parametrize('R,C', [(170, 240), (70, 300), (90, 40), (8, 8), (1, 5)])
def test_model_returns_the_reference_assignment_on_tie_free_crowds(R, C):
    rng = np.random.default_rng(R * 7 + C)
    for _ in range(8):
        cost = crowd(rng, R, C)
        x, _ = lapjv(cost, 0.9)
        pairs, refused, _ = sparse_assign(cost, 0.9)
        assert not refused
        assert pairs == sorted((int(i), int(j)) for i, j in enumerate(x) if j >= 0)


def test_model_refuses_equal_cost_alternatives():
    rng = np.random.default_rng(3)
    for _ in range(8):
        _, refused, _ = sparse_assign(crowd(rng, 60, 80, ties=True), 0.9)
        assert refused


def test_model_handles_no_admissible_entry_and_single_pairs():
    cost = np.full((5, 7), 1.0, np.float32)
    pairs, refused, _ = sparse_assign(cost, 0.9)
    assert pairs == [] and not refused
    cost[3, 2] = 0.2
    pairs, refused, _ = sparse_assign(cost, 0.9)
    assert pairs == [(3, 2)] and not refused
