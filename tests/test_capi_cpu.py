"""CPU-side checks of the product library: it loads, exports every symbol include/boojum_b200.h declares, its host
self-test hooks (the same gl64 / Poseidon2 source the kernels compile) agree with the oracle, and without a GPU it
fails loudly instead of falling back."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = O.P


@pytest.fixture(scope="module")
def native():
    from era_boojum_b200 import native as n
    return n


def test_exports_match_header(native):
    hdr = open(os.path.join(ROOT, "include", "boojum_b200.h")).read()
    declared = set(re.findall(r"BJ_API\s+[\w\s\*]+?\b(bj_\w+)\s*\(", hdr))
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(native.lib, name), "library does not export %s" % name
    assert declared == set(native.SIGNATURES), declared ^ set(native.SIGNATURES)


def test_host_field_ops(native):
    lib = native.lib
    r = np.random.default_rng(3)
    vals = [int(x) for x in r.integers(0, 2**64, size=500, dtype=np.uint64)]
    vals += [0, 1, P - 1, P, P + 1, 2**64 - 1, 2**32 - 1, 2**32, 2**63]
    for a in vals[:60]:
        for b in vals[-12:]:
            assert lib.bj_host_gl_mul(a, b) == (a % P) * (b % P) % P
            assert lib.bj_host_gl_add(a, b) == (a + b) % P
            assert lib.bj_host_gl_sub(a, b) == (a - b) % P
    for a in vals:
        for s in (0, 1, 3, 12, 31, 32, 33, 63):
            assert lib.bj_host_gl_mul_pow2(a, s) == (a % P) * (1 << s) % P
    for a in vals[:20]:
        if a % P:
            assert lib.bj_host_gl_inv(a) * (a % P) % P == 1
    x = (ctypes.c_uint64 * 2)(vals[0], vals[1])
    y = (ctypes.c_uint64 * 2)(vals[2], vals[3])
    o = (ctypes.c_uint64 * 2)()
    lib.bj_host_e2_mul(x, y, o)
    assert (o[0], o[1]) == O.ext_mul((vals[0], vals[1]), (vals[2], vals[3]))
    lib.bj_host_e2_inv(x, o)
    assert (o[0], o[1]) == O.ext_inv((vals[0], vals[1]))


def test_host_poseidon2_matches_oracle(native):
    r = np.random.default_rng(4)
    for _ in range(20):
        st = r.integers(0, 2**64, size=12, dtype=np.uint64)
        mine = st.copy()
        native.lib.bj_host_poseidon2_permutation(mine.ctypes.data_as(ctypes.c_void_p))
        assert np.array_equal(mine, O.poseidon2_permutation(st))


def test_no_cpu_fallback(native):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    st = native.lib.bj_ctx_create(0, None, ctypes.byref(h))
    assert st == native.BJ_ERR_NO_DEVICE
    import era_boojum_b200 as bj
    with pytest.raises(bj.BoojumError):
        bj.Context(0)


@pytest.mark.parametrize("log_n,cols", [(4, 1), (10, 60), (16, 93), (20, 155), (22, 93)])
def test_non_residues_for_copy_permutation_host(native, log_n, cols):
    """bj_non_residues_for_copy_permutation (host code; make_non_residues, src/cs/implementations/utils.rs:636-688) against the
    Python-int restatement: k_0 = 1, then the smallest quadratic non-residues whose n-th powers are pairwise distinct and != 1"""
    from oracle.stage2 import non_residues_for_copy_permutation
    out = (ctypes.c_uint64 * cols)()
    assert native.lib.bj_non_residues_for_copy_permutation(1 << log_n, cols, out) == 0
    assert list(out) == non_residues_for_copy_permutation(1 << log_n, cols)
