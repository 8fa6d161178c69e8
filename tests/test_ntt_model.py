"""CPU check of the NTT planner / tile / stage / twiddle-index arithmetic used by the CUDA kernel
(tools/ntt_model.py transcribes era_boojum_b200/csrc/ntt.cu) against the oracle."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import ntt_model as M  # noqa: E402
from oracle import oracle as O  # noqa: E402


@pytest.mark.parametrize("m", [4, 5, 7, 10, 13, 14])
@pytest.mark.parametrize("coset", [1, 7])
def test_model_matches_oracle(m, coset):
    a = O.random_field(np.random.default_rng(m), 1 << m)
    f, _ = M.transform(a, coset, False)
    assert np.array_equal(f, O.ntt_n2b(a, coset))
    g, _ = M.transform(a, coset, True)
    assert np.array_equal(g, O.intt_n2n(a, coset))


def test_plans_cover_all_rounds():
    for m in range(4, 31):
        for inv in (False, True):
            plan = M.make_plan(m, inv)
            assert sum(t for t, _ in plan) == m
            r0 = 0
            for i, (t, w) in enumerate(plan):
                last = i == len(plan) - 1
                assert 4 <= t + w <= 13 or m < 4
                if not (inv and last):
                    assert w <= m - r0 - t
                else:
                    assert w <= r0
                r0 += t
