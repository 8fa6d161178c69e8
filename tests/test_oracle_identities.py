"""Differential / identity tests of the CPU oracle, mirroring the reference's own test strategy
(SURVEY.md section 4): NTT vs O(n^2) DFT (src/fft/mod.rs:1344-1384, 1591-1634), round trips (:1636-1709),
fold-by-value == fold-by-coefficients (src/cs/implementations/fri/mod.rs:960-1031), batch inverse vs single
inverse (src/cs/implementations/utils.rs:1750-1818), generator sanity (src/field/goldilocks/mod.rs:614-653).
CPU only."""
import numpy as np
import pytest

from oracle import oracle as O

P = O.P


def rng(seed=0):
    return np.random.default_rng(seed)


def test_field_mul_fast_equals_definition():
    r = rng(1)
    a = [int(x) for x in r.integers(0, 2**64, size=2000, dtype=np.uint64)]
    b = [int(x) for x in r.integers(0, 2**64, size=2000, dtype=np.uint64)]
    edge = [0, 1, P - 1, P, P + 1, 2**64 - 1, 2**32, 2**32 - 1, 2**63]
    for x in edge:
        for y in edge:
            assert O.mul(x, y) == (x % P) * (y % P) % P
    for x, y in zip(a, b):
        assert O.mul(x, y) == (x % P) * (y % P) % P
        assert O.add(x, y) == (x + y) % P
        assert O.sub(x, y) == (x - y) % P


def test_generators():
    assert O.pow_(7, (P - 1) >> 32) == 0x185629DCDA58878C
    assert O.omega(1) == P - 1
    for k in range(1, 33):
        w = O.omega(k)
        assert O.pow_(w, 1 << k) == 1 and O.pow_(w, 1 << (k - 1)) == P - 1


def test_ext_field():
    r = rng(2)
    a, b = O.random_field(r, 2), O.random_field(r, 2)
    ab = O.ext_mul(a, b)
    assert ab == ((int(a[0]) * int(b[0]) + 7 * int(a[1]) * int(b[1])) % P,
                  (int(a[0]) * int(b[1]) + int(a[1]) * int(b[0])) % P)
    assert O.ext_mul(a, O.ext_inv(a)) == (1, 0)


@pytest.mark.parametrize("log_n", range(1, 10))
@pytest.mark.parametrize("coset", [1, 7])
def test_ntt_vs_naive_dft(log_n, coset):
    a = O.random_field(rng(log_n), 1 << log_n)
    assert np.array_equal(O.ntt_n2b(a, coset), O.naive_dft_bitreversed(a, coset))


def test_ntt_noncanonical_inputs():
    a = np.array([P, P + 5, 2**64 - 1, 0, 1, P - 1, 2**63, 12345], dtype=np.uint64)
    canon = a % np.uint64(P)
    assert np.array_equal(O.ntt_n2b(a, 7), O.ntt_n2b(canon, 7))


@pytest.mark.parametrize("log_n", [1, 4, 10, 16])
@pytest.mark.parametrize("coset", [1, 7])
def test_intt_roundtrip(log_n, coset):
    a = O.random_field(rng(100 + log_n), 1 << log_n)
    fwd = O.ntt_n2b(a, coset)
    back = O.intt_n2n(O.bitreverse(fwd), coset)
    assert np.array_equal(back, a)


def test_lde_definition():
    """storage[j][i] = f(7 * w_{nL}^{bitrev_L(j)} * w_n^{bitrev_n(i)})  (SURVEY A.4)."""
    log_n, log_l = 5, 3
    n, L = 1 << log_n, 1 << log_l
    vals = O.random_field(rng(7), (2, n))
    out = O.lde(vals, log_l)
    mono = O.intt_n2n(vals)
    wn, wnl = O.omega(log_n), O.omega(log_n + log_l)
    br = lambda x, b: int(format(x, "0%db" % b)[::-1], 2) if b else 0
    for c in range(2):
        coeffs = [int(x) for x in mono[c]]
        for j in range(L):
            for i in range(n):
                x = 7 * pow(wnl, br(j, log_l), P) * pow(wn, br(i, log_n), P) % P
                acc = 0
                for co in reversed(coeffs):
                    acc = (acc * x + co) % P
                assert acc == int(out[c, j, i])
    # flat index t = j*n+i holds f(7 * w_{nL}^{bitrev_{nL}(t)})
    flat = out[0].reshape(-1)
    for t in (0, 1, 37, n * L - 1):
        x = 7 * pow(wnl, br(t, log_n + log_l), P) % P
        acc = 0
        for co in reversed([int(v) for v in mono[0]]):
            acc = (acc * x + co) % P
        assert acc == int(flat[t])
    # first d cosets are the factor-d LDE (subset_for_degree, polynomial/lde.rs:298-308)
    assert np.array_equal(O.lde(vals, 1), out[:, :2, :])


def test_fri_fold_by_value_equals_by_coefficients():
    log_n = 8
    n = 1 << log_n
    r = rng(11)
    m0, m1 = O.random_field(r, n), O.random_field(r, n)
    alpha = O.random_field(r, 2)
    # values on coset 7 in bit-reversed order
    v0, v1 = O.ntt_n2b(m0, 7), O.ntt_n2b(m1, 7)
    roots = O.twiddles(log_n, inverse=True)
    f0, f1 = O.fri_fold(v0, v1, alpha, roots, O.inv(7))
    # by coefficients: g = 2*(even + alpha*odd)  (no division by two in the reference's fold)
    e = [(int(m0[2 * i]), int(m1[2 * i])) for i in range(n // 2)]
    o = [(int(m0[2 * i + 1]), int(m1[2 * i + 1])) for i in range(n // 2)]
    al = (int(alpha[0]), int(alpha[1]))
    g0, g1 = [], []
    for (e0, e1), (o0, o1) in zip(e, o):
        t = O.ext_mul(al, (o0, o1))
        g0.append(2 * (e0 + t[0]) % P)
        g1.append(2 * (e1 + t[1]) % P)
    w0 = O.ntt_n2b(np.array(g0, dtype=np.uint64), 49)
    w1 = O.ntt_n2b(np.array(g1, dtype=np.uint64), 49)
    assert np.array_equal(w0, f0) and np.array_equal(w1, f1)


def test_batch_inverse():
    r = rng(5)
    a = O.random_field(r, 1000)
    a[a == 0] = 1
    inv = O.batch_inverse(a)
    for x, y in zip(a[:50], inv[:50]):
        assert int(x) * int(y) % P == 1
    c0, c1 = O.random_field(r, 100), O.random_field(r, 100)
    i0, i1 = O.batch_inverse_ext(c0, c1)
    for k in range(100):
        assert O.ext_mul((c0[k], c1[k]), (i0[k], i1[k])) == (1, 0)


def test_merkle_tree_and_paths():
    r = rng(9)
    cols = [O.random_field(r, 64) for _ in range(11)]
    lh, levels, cap = O.merkle_tree(cols, cap_size=4)
    assert lh.shape == (64, 4) and [l.shape[0] for l in levels] == [32, 16, 8, 4]
    row = np.array([c[5] for c in cols], dtype=np.uint64)
    assert np.array_equal(O.poseidon2_hash_leaf(row), lh[5])
    assert np.array_equal(O.poseidon2_hash_node(lh[10], lh[11]), levels[0][5])
    for idx in (0, 5, 63):
        path = O.merkle_path(lh, levels, idx)
        assert path.shape[0] == 4
        assert O.merkle_verify(lh[idx], path, cap, idx)
        assert not O.merkle_verify(lh[idx ^ 1], path, cap, idx)
    # chunked leaves (FRI oracles): leaf m = c0[m*k:(m+1)*k] || c1[m*k:(m+1)*k]
    c0, c1 = O.random_field(r, 64), O.random_field(r, 64)
    lh2 = O.merkle_leaf_hashes([c0, c1], elems_per_leaf=8)
    assert np.array_equal(lh2[3], O.poseidon2_hash_leaf(np.concatenate([c0[24:32], c1[24:32]])))


def test_oracle_do_fri_commit_then_verify():
    """oracle do_fri (prover side) is accepted by the verifier-side chain that the golden fixture pins."""
    from oracle import replay
    r = rng(21)
    log_n, log_lde, cap, schedule = 8, 3, 16, [3, 3, 1]
    n = 1 << log_n
    m = O.random_field(r, (2, n))
    lde = O.lde(m, log_lde, from_monomials=True)
    c0, c1 = lde[0].reshape(-1), lde[1].reshape(-1)
    t = replay.Poseidon2Transcript()
    t.witness_field_elements([1, 2, 3])
    w = replay.do_fri_oracle(c0, c1, t, schedule, log_lde, cap)
    assert len(w["monomials"][0]) == n >> sum(schedule)
    for idx in (0, 5, 1000, n * 8 - 1):
        qs, sub = [], idx
        for lvl, k in enumerate(schedule):
            l0, l1 = w["levels"][lvl]
            lh, lv = w["trees"][lvl]
            ti = sub >> k
            le = np.concatenate([l0[ti << k:(ti + 1) << k], l1[ti << k:(ti + 1) << k]])
            qs.append((le, O.merkle_path(lh, lv, ti)))
            sub >>= k
        assert replay.verify_fri_query(idx, log_n, log_lde, schedule, cap, w["caps"], w["challenges"], w["monomials"], qs,
                                       (int(c0[idx]), int(c1[idx])))
