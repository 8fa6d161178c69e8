"""GPU parity tests: the CUDA path, called through the C-ABI (era_boojum_b200 -> libboojum_b200.so), against the CPU
oracle on the same seeded inputs, against the reference's golden fixture, and - at BASELINE.json sizes - through
size-independent properties.  Bit-exact: all values are integers mod p compared after canonicalisation."""
import numpy as np
import pytest

from oracle import oracle as O
from oracle import replay

pytestmark = pytest.mark.gpu

P = O.P


@pytest.fixture(scope="module")
def bj():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import era_boojum_b200 as m
    return m


@pytest.fixture(scope="module")
def ctx(bj):
    c = bj.Context.on_current_stream(0)
    yield c
    c.synchronize()
    c.close()


def rng(seed):
    return np.random.default_rng(seed)


# ------------------------------------------------------------------------------------------- NTT -------
def test_twiddles_match_reference_table(bj, ctx):
    for log_n in (1, 2, 5, 12, 16):
        for inv in (False, True):
            got = bj.to_numpy(ctx.precompute_twiddles_for_fft(1 << log_n, inv))
            assert np.array_equal(got[: max(1, (1 << log_n) // 2)], O.twiddles(log_n, inv))


@pytest.mark.parametrize("log_n", list(range(0, 17)) + [18, 20])
@pytest.mark.parametrize("coset", [1, 7])
def test_ntt_forward_matches_oracle(bj, ctx, log_n, coset):
    n_cols = 3 if log_n <= 16 else 2
    a = O.random_field(rng(log_n * 2 + coset), (n_cols, 1 << log_n))
    d = bj.to_device(a)
    ctx.fft_natural_to_bitreversed(d, coset)
    assert np.array_equal(bj.to_numpy(d), O.ntt_n2b(a, coset))


@pytest.mark.parametrize("log_n", list(range(0, 17)) + [18, 20])
@pytest.mark.parametrize("coset", [1, 7])
def test_ntt_inverse_matches_oracle(bj, ctx, log_n, coset):
    n_cols = 3 if log_n <= 16 else 2
    a = O.random_field(rng(1000 + log_n * 2 + coset), (n_cols, 1 << log_n))
    d = bj.to_device(a)
    ctx.ifft_natural_to_natural(d, coset)
    assert np.array_equal(bj.to_numpy(d), O.intt_n2n(a, coset))


def test_ntt_config1_2pow16_forward_inverse(bj, ctx):
    """BASELINE config 1: 2^16-point forward + inverse, cosets 1 and 7, adversarial inputs incl. non-canonical
    values in [p, 2^64) (reference: src/field/goldilocks/generic_impl.rs:449-455) and 0/1 vectors (fft/mod.rs:1432)."""
    n = 1 << 16
    r = rng(16)
    cases = [
        O.random_field(r, n),
        np.zeros(n, np.uint64),
        np.full(n, P - 1, np.uint64),
        r.integers(P, 2**64, size=n, dtype=np.uint64),       # all non-canonical
        np.full(n, 2**64 - 1, np.uint64),
        (np.arange(n) % 2).astype(np.uint64),
        np.concatenate([np.ones(1, np.uint64), np.zeros(n - 1, np.uint64)]),
    ]
    a = np.stack(cases)
    for coset in (1, 7):
        d = bj.to_device(a)
        ctx.fft_natural_to_bitreversed(d, coset)
        fwd = bj.to_numpy(d)
        assert np.array_equal(fwd, O.ntt_n2b(a, coset))
        assert (fwd < np.uint64(P)).all()
        d = bj.to_device(a)
        ctx.ifft_natural_to_natural(d, coset)
        assert np.array_equal(bj.to_numpy(d), O.intt_n2n(a, coset))
        # round trip: iNTT(bitrev(NTT(a))) == a mod p
        d = bj.to_device(a)
        ctx.fft_natural_to_bitreversed(d, coset)
        ctx.bitreverse_enumeration_inplace(d)
        ctx.ifft_natural_to_natural(d, coset)
        assert np.array_equal(bj.to_numpy(d), a % np.uint64(P))


def test_ntt_strided_batch(bj, ctx):
    """columns separated by a stride larger than n (col_stride argument of the C-ABI)."""
    import ctypes
    import torch
    from era_boojum_b200 import native
    n, cols, stride = 1 << 13, 4, (1 << 13) + 64
    a = O.random_field(rng(5), (cols, n))
    buf = np.zeros(cols * stride, np.uint64)
    for c in range(cols):
        buf[c * stride: c * stride + n] = a[c]
    buf[n:stride] = 12345  # padding must stay untouched
    d = bj.to_device(buf)
    st = native.lib.bj_ntt_natural_to_bitreversed(ctx._h, ctypes.c_void_p(d.data_ptr()), 13, cols, stride, 7)
    assert st == 0
    out = bj.to_numpy(d)
    want = O.ntt_n2b(a, 7)
    for c in range(cols):
        assert np.array_equal(out[c * stride: c * stride + n], want[c])
    assert (out[n:stride] == 12345).all()
    torch.cuda.synchronize()


@pytest.mark.parametrize("log_n", [22, 24])
def test_ntt_full_size_vs_oracle_and_properties(bj, ctx, log_n):
    """BASELINE config 2 sizes: one column compared with the oracle directly, plus linearity and round trip."""
    n = 1 << log_n
    r = rng(log_n)
    a, b = O.random_field(r, n), O.random_field(r, n)
    d = bj.to_device(np.stack([a, b]))
    ctx.fft_natural_to_bitreversed(d, 7)
    fa, fb = bj.to_numpy(d)
    assert np.array_equal(fa, O.ntt_n2b(a, 7))
    # linearity: NTT(a + 3b) == NTT(a) + 3 NTT(b)
    lin = ((a.astype(object) + 3 * b.astype(object)) % P).astype(np.uint64)
    d2 = bj.to_device(lin)
    ctx.fft_natural_to_bitreversed(d2, 7)
    want = ((fa.astype(object) + 3 * fb.astype(object)) % P).astype(np.uint64)
    assert np.array_equal(bj.to_numpy(d2), want)
    # round trip
    ctx.bitreverse_enumeration_inplace(d)
    ctx.ifft_natural_to_natural(d, 7)
    back = bj.to_numpy(d)
    assert np.array_equal(back[0], a) and np.array_equal(back[1], b)
    # inverse against the oracle at full size too
    d3 = bj.to_device(a)
    ctx.ifft_natural_to_natural(d3, 7)
    assert np.array_equal(bj.to_numpy(d3), O.intt_n2n(a, 7))


# ------------------------------------------------------------------------------------------- LDE -------
@pytest.mark.parametrize("log_n,log_l,cols", [(4, 1, 2), (10, 3, 5), (13, 2, 3), (16, 3, 3)])
def test_lde_matches_oracle(bj, ctx, log_n, log_l, cols):
    a = O.random_field(rng(log_n + log_l), (cols, 1 << log_n))
    d = bj.to_device(a)
    out = ctx.transform_raw_storages_to_lde(d, 1 << log_l)
    assert np.array_equal(bj.to_numpy(out), O.lde(a, log_l))
    assert np.array_equal(bj.to_numpy(d), a)  # input untouched
    mono = O.intt_n2n(a)
    out2 = ctx.transform_raw_storages_to_lde(bj.to_device(mono), 1 << log_l, from_monomials=True)
    assert np.array_equal(bj.to_numpy(out2), O.lde(a, log_l))


def test_lde_large_first_coset_subset(bj, ctx):
    """n = 2^20, L = 8: the first 2 cosets equal the L = 2 LDE (subset_for_degree, polynomial/lde.rs:298-308) and one
    coset is checked against the oracle."""
    log_n = 20
    a = O.random_field(rng(77), (2, 1 << log_n))
    d = bj.to_device(a)
    out8 = bj.to_numpy(ctx.transform_raw_storages_to_lde(d, 8))
    out2 = bj.to_numpy(ctx.transform_raw_storages_to_lde(d, 2))
    assert np.array_equal(out8[:, :2, :], out2)
    want = O.lde(a[:1], 3)
    assert np.array_equal(out8[0], want[0])


# ------------------------------------------------------------------------------------ Poseidon2 / Merkle -----
def test_poseidon2_permutation_matches_oracle(bj, ctx):
    st = rng(8).integers(0, 2**64, size=(300, 12), dtype=np.uint64)  # includes non-canonical values
    d = bj.to_device(st)
    ctx.poseidon2_permute(d)
    got = bj.to_numpy(d)
    for i in range(0, 300, 7):
        assert np.array_equal(got[i], O.poseidon2_permutation(st[i]))


@pytest.mark.parametrize("row_len", [1, 4, 7, 8, 9, 16, 58, 93, 156])
def test_poseidon2_leaf_hash_rows(bj, ctx, row_len):
    rows = O.random_field(rng(row_len), (33, row_len))
    got = bj.to_numpy(ctx.poseidon2_hash_rows(bj.to_device(rows)))
    for i in range(33):
        assert np.array_equal(got[i], O.poseidon2_hash_leaf(rows[i]))


@pytest.mark.parametrize("n_cols,log_leaves,cap", [(1, 4, 1), (8, 6, 4), (11, 8, 16), (93, 12, 16), (100, 10, 1024)])
def test_merkle_tree_matches_oracle(bj, ctx, n_cols, log_leaves, cap):
    n = 1 << log_leaves
    cols = [O.random_field(rng(100 * n_cols + c), n) for c in range(n_cols)]
    tree = ctx.merkle_tree_construct([bj.to_device(c) for c in cols], cap)
    lh, levels, capd = O.merkle_tree(cols, cap)
    assert np.array_equal(bj.to_numpy(tree.leaf_hashes), lh)
    got_levels = tree.levels()
    assert len(got_levels) == len(levels)
    for g, w in zip(got_levels, levels):
        assert np.array_equal(bj.to_numpy(g), w)
    assert np.array_equal(tree.get_cap(), capd)
    if n > cap:
        for idx in (0, 1, n // 2 + 3, n - 1):
            leaf, path = tree.get_proof(idx)
            assert O.merkle_verify(leaf, path, capd, idx)


@pytest.mark.parametrize("k", [2, 4, 8])
def test_merkle_chunked_leaves_for_fri_oracles(bj, ctx, k):
    n = 1 << 10
    c0, c1 = O.random_field(rng(k), n), O.random_field(rng(k + 50), n)
    tree = ctx.merkle_tree_construct([bj.to_device(c0), bj.to_device(c1)], 4, elems_per_leaf=k)
    lh = O.merkle_leaf_hashes([c0, c1], elems_per_leaf=k)
    assert np.array_equal(bj.to_numpy(tree.leaf_hashes), lh)
    assert np.array_equal(tree.get_cap(), O.merkle_nodes(lh, 4)[-1])


def test_merkle_lde_layout_coset_major(bj, ctx):
    """witness-oracle shape: LDE of C columns, leaf t = coset*n + row absorbs the C column values at t."""
    log_n, L, C = 8, 4, 9
    a = O.random_field(rng(31), (C, 1 << log_n))
    out = ctx.transform_raw_storages_to_lde(bj.to_device(a), L)           # [C, L, n]
    tree = ctx.merkle_tree_construct([out[c].reshape(-1) for c in range(C)], 16)
    want_lde = O.lde(a, 2)
    lh, levels, cap = O.merkle_tree([want_lde[c].reshape(-1) for c in range(C)], 16)
    assert np.array_equal(bj.to_numpy(tree.leaf_hashes), lh)
    assert np.array_equal(tree.get_cap(), cap)


def test_golden_fixture_leaf_hashes_on_gpu(bj, ctx, golden_fixture):
    """The reference's proof.json: GPU leaf hashes of the opened rows verify against the fixture's caps."""
    c = replay.replay_proof(golden_fixture)  # oracle-side replay (also yields nothing GPU specific)
    assert c["merkle_paths"] > 0
    proof, vk = golden_fixture["proof"], golden_fixture["vk"]
    idxs = _query_indices(golden_fixture)
    for q, idx in zip(proof["queries_per_fri_repetition"], idxs):
        for name, cap in (("witness_query", proof["witness_oracle_cap"]), ("stage_2_query", proof["stage_2_oracle_cap"]),
                          ("quotient_query", proof["quotient_oracle_cap"]), ("setup_query", vk["setup_merkle_tree_cap"])):
            row = np.array(q[name]["leaf_elements"], dtype=np.uint64)[None, :]
            leaf = bj.to_numpy(ctx.poseidon2_hash_rows(bj.to_device(row)))[0]
            path = np.array(q[name]["proof"], dtype=np.uint64).reshape(-1, 4)
            assert O.merkle_verify(leaf, path, np.array(cap, dtype=np.uint64), idx)


def _query_indices(fx):
    vk, proof = fx["vk"], fx["proof"]
    tr = replay.Poseidon2Transcript()
    tr.witness_merkle_tree_cap(vk["setup_merkle_tree_cap"])
    for v in proof["public_inputs"]:
        tr.witness_field_elements([v])
    tr.witness_merkle_tree_cap(proof["witness_oracle_cap"])
    for _ in range(8):
        tr.get_challenge()
    tr.witness_merkle_tree_cap(proof["stage_2_oracle_cap"])
    tr.get_ext_challenge()
    tr.witness_merkle_tree_cap(proof["quotient_oracle_cap"])
    tr.get_ext_challenge()
    for g in ("values_at_z", "values_at_z_omega", "values_at_0"):
        for v in proof[g]:
            tr.witness_field_elements(v["coeffs"])
    tr.get_ext_challenge()
    for cap in [proof["fri_base_oracle_cap"]] + list(proof["fri_intermediate_oracles_caps"]):
        tr.witness_merkle_tree_cap(cap)
        tr.get_ext_challenge()
    tr.witness_field_elements(proof["final_fri_monomials"][0])
    tr.witness_field_elements(proof["final_fri_monomials"][1])
    bb = replay.BoolsBuffer(21)
    out = []
    for _ in proof["queries_per_fri_repetition"]:
        bits = bb.get_bits(tr, 21)
        out.append(sum(b << i for i, b in enumerate(bits)))
    return out


# ------------------------------------------------------------------------------------------- FRI -------
@pytest.mark.parametrize("log_m,log_fold", [(3, 1), (3, 3), (12, 1), (12, 2), (12, 3), (17, 3)])
def test_fri_fold_matches_oracle(bj, ctx, log_m, log_fold):
    m = 1 << log_m
    r = rng(log_m * 4 + log_fold)
    c0, c1 = O.random_field(r, m), O.random_field(r, m)
    alpha = [int(x) for x in O.random_field(r, 2)]
    kappa = O.inv(7)
    o0, o1, new_kappa = ctx.fri_fold(bj.to_device(c0), bj.to_device(c1), log_fold, alpha, kappa)
    roots = O.twiddles(log_m, inverse=True)
    w0, w1, a, k = c0, c1, tuple(alpha), kappa
    for _ in range(log_fold):
        w0, w1 = O.fri_fold(w0, w1, a, roots[: len(w0) // 2], k)
        a = O.ext_mul(a, a)
        k = O.mul(k, k)
    assert np.array_equal(bj.to_numpy(o0), w0) and np.array_equal(bj.to_numpy(o1), w1)
    assert new_kappa == k


def test_fri_fold_golden_fixture(bj, ctx, golden_fixture):
    """Fold the opened FRI leaves of the reference proof on the GPU: the result must be the element found in the
    next oracle's leaf (verifier.rs:2386-2510).  The leaf is embedded at its true position of a zero vector so the
    kernel uses the same root indices as the real codeword."""
    fx = golden_fixture
    c = replay.replay_proof(fx)
    sched = c["schedule"]
    idxs = _query_indices(fx)
    log_full = fx["vk"]["fixed_parameters"]["domain_size"].bit_length() - 1 + 1
    for q, idx in zip(fx["proof"]["queries_per_fri_repetition"][:3], idxs[:3]):
        kappa = O.inv(7)
        sub, log_m = idx, log_full
        for lvl in range(len(sched) - 1):
            k = sched[lvl]
            deg = 1 << k
            tree_idx = sub >> k
            le = q["fri_queries"][lvl]["leaf_elements"]
            c0 = np.zeros(1 << log_m, np.uint64)
            c1 = np.zeros(1 << log_m, np.uint64)
            c0[tree_idx * deg:(tree_idx + 1) * deg] = le[:deg]
            c1[tree_idx * deg:(tree_idx + 1) * deg] = le[deg:]
            alpha = c["challenges"]["fri"][lvl][0]
            o0, o1, kappa = ctx.fri_fold(bj.to_device(c0), bj.to_device(c1), k, alpha, kappa)
            nxt = q["fri_queries"][lvl + 1]["leaf_elements"]
            ndeg = 1 << sched[lvl + 1]
            pos = tree_idx % ndeg
            assert int(bj.to_numpy(o0[tree_idx:tree_idx + 1])[0]) == nxt[pos]
            assert int(bj.to_numpy(o1[tree_idx:tree_idx + 1])[0]) == nxt[ndeg + pos]
            sub, log_m = tree_idx, log_m - k


def test_device_field_selftest(bj, ctx):
    """Inline-PTX mul/add/sub vs the portable C versions inside one kernel (4M random + edge inputs)."""
    import ctypes
    from era_boojum_b200 import native
    bad = ctypes.c_uint64(123)
    st = native.lib.bj_selftest_field(ctx._h, 1 << 22, 20260924, ctypes.byref(bad))
    assert st == 0 and bad.value == 0
