"""ORACLE (test infrastructure, never on the product path): `prove_cpu_basic` restated end to end on the oracle primitives -
an independent CPU prover for SMALL circuits, so that the proof of the GPU library can be compared with it bit for bit.

Follows src/cs/implementations/prover.rs:153-2269 round by round (the line ranges are quoted at each step) for circuits made of
the three bench gates (oracle/gates.py), optionally with the lookup argument over specialised columns with the table id in a
constant column and with public inputs; every tree hasher (Poseidon2 / Blake2s256 / Keccak256) and transcript (Poseidon2,
Poseidon, Blake2s, Keccak256) of the reference - the bench's two type-parameter pairs included; no proof of work.
Where the GPU library evaluates on the LDE domain with kernels, this file loops over the points in Python with the point-wise
restatements of oracle/stage2.py, oracle/lookup.py and oracle/gates.py; NTT / LDE / Merkle / FRI folds / DEEP come from the C
restatement (oracle/oracle.c).  Openings are evaluated from the MONOMIAL forms by Horner's rule (the library uses barycentric
evaluation on a coset), the copy-permutation polynomials come from oracle/stage2.py's row loop (the library: chunked ratios,
one batch inversion, a scan) - the agreement of the two proofs is therefore a statement about two different computations.

    prove(...)  ->  (proof dict in the reference's serde shape (proof.rs:57-143), setup cap)
"""
import numpy as np

from . import gates as G
from . import lookup as LK
from . import oracle as O
from . import replay as R
from . import stage2 as S2

P = R.P


def _flat(lde3, c):
    return lde3[c].reshape(-1)


def _ext_dict(v):
    return {"coeffs": [int(v[0]), int(v[1])], "_marker": None}


def _combine(a, b):
    """f0(at) + u f1(at), u^2 = 7"""
    return ((a[0] + 7 * b[1]) % P, (a[1] + b[0]) % P)


def _ext_columns(vals):
    """list of Fp2 tuples -> (c0 column, c1 column) as uint64 arrays"""
    return (np.array([v[0] for v in vals], dtype=np.uint64), np.array([v[1] for v in vals], dtype=np.uint64))


TRANSCRIPTS = {"poseidon2": R.Poseidon2Transcript, "poseidon": R.PoseidonTranscript, "blake2s": R.Blake2sTranscript,
               "keccak256": R.Keccak256Transcript}


def _serde(digests, hasher):
    """TreeHasher::Output values (4 u64 words each) in serde form: [u64; 4] for Poseidon2, [u8; 32] for Blake2s256 / Keccak256"""
    a = np.ascontiguousarray(np.asarray(digests, dtype=np.uint64).reshape(-1, 4))
    if hasher == "poseidon2":
        return a.tolist()
    return a.astype("<u8").view(np.uint8).reshape(-1, 32).tolist()


def prove(variables, sigmas, constants, gates, quotient_degree, fri_lde_factor, cap_size, security_level=100, lookup=None,
          public_inputs=(), hasher="poseidon2", transcript=None):
    """variables / sigmas: [V, n], constants: [C, n] uint64 arrays in natural row order; gates: [(name, repetitions, selector
    path[, first variable column, first constant column[, recorded program]])] as oracle/gates.py::quotient_gates_row takes them,
    in the order of their quotient terms (specialised-column gates first); lookup: None or dict(width, num_repetitions, variables_offset,
    table_id_column, tables [width + 1, n], multiplicities [n]); public_inputs: [(column, row)]."""
    variables, sigmas, constants = (np.asarray(a, dtype=np.uint64) % np.uint64(P) for a in (variables, sigmas, constants))
    V, n = variables.shape
    C = constants.shape[0]
    Q, L = quotient_degree, fri_lde_factor
    log_n, log_q, log_l = n.bit_length() - 1, Q.bit_length() - 1, L.bit_length() - 1
    log_d = max(log_l, log_q)                      # used_lde_degree (prover.rs:178-196)
    D = 1 << log_d
    ncm = n * L                                    # the oracles commit to the first L cosets (subset_for_degree)
    lk = lookup
    tables = np.asarray(lk["tables"], dtype=np.uint64) if lk else np.zeros((0, n), np.uint64)
    mult = np.asarray(lk["multiplicities"], dtype=np.uint64) if lk else None
    T = tables.shape[0]
    tree = lambda cols: R.merkle_tree_with_hasher(cols, cap_size, 1, hasher)
    commit = lambda cols: tree([c[:ncm] for c in cols])

    # ---- setup oracle: sigmas | constants | lookup tables (setup.rs:1093-1255) ----
    setup_lde = O.lde(np.concatenate([sigmas, constants, tables]), log_d)
    setup_cols = [_flat(setup_lde, c) for c in range(V + C + T)]
    sigma_cols, const_cols, table_cols = setup_cols[:V], setup_cols[V:V + C], setup_cols[V + C:]
    setup_lh, setup_lv, setup_cap = commit(setup_cols)

    tr = TRANSCRIPTS[transcript or hasher]()
    tr.witness_merkle_tree_cap(setup_cap.tolist())                       # prover.rs:211
    public_values = [int(variables[c, r]) for c, r in public_inputs]
    for v in public_values:                                              # prover.rs:264-266
        tr.witness_field_elements([v])

    # ---- round 1: witness commitment (prover.rs:313-353) ----
    w_lde = O.lde(variables, log_d)
    w_cols = [_flat(w_lde, c) for c in range(V)]
    m_col = _flat(O.lde(mult[None, :], log_d), 0) if lk else None
    w_oracle_cols = w_cols + ([m_col] if lk else [])                     # variables | witness (none) | multiplicities
    w_lh, w_lv, w_cap = commit(w_oracle_cols)
    tr.witness_merkle_tree_cap(w_cap.tolist())

    # ---- round 2: copy-permutation grand product, partial products, lookup polynomials (prover.rs:360-554) ----
    beta, gamma = tr.get_ext_challenge(), tr.get_ext_challenge()
    if lk:
        lookup_beta, lookup_gamma = tr.get_ext_challenge(), tr.get_ext_challenge()          # prover.rs:402-406
        wdt, nsub, voff, tid = lk["width"], lk["num_repetitions"], lk["variables_offset"], lk["table_id_column"]
    var_rows = [[int(x) for x in variables[j]] for j in range(V)]
    sig_rows = [[int(x) for x in sigmas[j]] for j in range(V)]
    z, partials = S2.partial_products(var_rows, sig_rows, beta, gamma, Q)
    st2 = list(_ext_columns(z))
    for pp in partials:
        st2 += list(_ext_columns(pp))
    n_partial = len(partials)
    if lk:
        a_polys, b_poly = LK.lookup_polys([variables[voff + i] for i in range(wdt * nsub)], wdt, constants[tid],
                                          [tables[j] for j in range(T)], mult, lookup_beta, lookup_gamma)
        for ap in a_polys:
            st2 += list(_ext_columns(ap))
        st2 += list(_ext_columns(b_poly))
    st2 = np.stack(st2)
    s2_lde = O.lde(st2, log_d)
    s2_cols = [_flat(s2_lde, c) for c in range(st2.shape[0])]
    s2_lh, s2_lv, s2_cap = commit(s2_cols)
    tr.witness_merkle_tree_cap(s2_cap.tolist())
    a_off = 2 + 2 * n_partial

    # ---- round 3: quotient (prover.rs:560-1495), point by point over the first Q cosets of the LDE domain ----
    alpha = tr.get_ext_challenge()
    # one term per repetition for the three bench gates; a gate given as a recorded program pushes len(writes) per repetition
    n_gate_terms = sum(g[1] * (len(g[5]["writes"]) if len(g) > 5 else 1) for g in gates)
    n_lk_terms = nsub + 1 if lk else 0                                   # lookup terms come first (prover.rs:608-625)
    total_terms = n_lk_terms + n_gate_terms + 1 + 1 + n_partial
    powers = R.ext_powers(alpha, total_terms)
    lk_ch, gate_ch, rest_ch = powers[:n_lk_terms], powers[n_lk_terms:n_lk_terms + n_gate_terms], powers[n_lk_terms + n_gate_terms:]
    z_lde = (s2_cols[0], s2_cols[1])
    part_ldes = [(s2_cols[2 + 2 * c], s2_cols[3 + 2 * c]) for c in range(n_partial)]
    npts = n * Q
    q0, q1 = np.zeros(npts, np.uint64), np.zeros(npts, np.uint64)
    van = [S2.vanishing_inverse(log_n, log_q, j) for j in range(Q)]
    if lk:
        a_ldes = [(s2_cols[a_off + 2 * i], s2_cols[a_off + 2 * i + 1]) for i in range(nsub)]
        b_lde = (s2_cols[a_off + 2 * nsub], s2_cols[a_off + 2 * nsub + 1])
    for t in range(npts):
        acc = (0, 0)
        if lk:
            acc = LK.quotient_lookup_point(t, [w_cols[voff + i] for i in range(wdt * nsub)], wdt, const_cols[tid], table_cols, m_col,
                                           a_ldes, b_lde, lookup_beta, lookup_gamma, lk_ch)
        acc = R.e_add(acc, G.quotient_gates_row(gates, [int(c[t]) for c in w_cols], [int(c[t]) for c in const_cols], gate_ch))
        acc = R.e_add(acc, S2.quotient_copy_permutation_point(t, log_n, log_d, w_cols, sigma_cols, z_lde, part_ldes, beta, gamma,
                                                              rest_ch, Q))
        acc = R.e_mul_base(acc, van[t >> log_n])                        # divide by the vanishing polynomial (utils.rs:770-817)
        q0[t], q1[t] = acc
    # cosets -> natural order, one interpolation of size n*Q on the coset 7, Q chunks of n coefficients (prover.rs:1399-1467)
    mono_q = [O.intt_n2n(O.bitreverse(q), 7) for q in (q0, q1)]
    if int(mono_q[0][npts - 1]) or int(mono_q[1][npts - 1]):
        raise ValueError("unsatisfied: quotient is not a polynomial of degree < n * quotient_degree")   # prover.rs:1425-1438
    chunks = np.stack([mono_q[k][j * n:(j + 1) * n] for j in range(Q) for k in (0, 1)])   # c0, c1 of chunk 0, chunk 1, ...
    qt_lde = O.lde(chunks, log_l, from_monomials=True)
    qt_cols = [_flat(qt_lde, c) for c in range(2 * Q)]
    qt_lh, qt_lv, qt_cap = tree(qt_cols)
    tr.witness_merkle_tree_cap(qt_cap.tolist())

    # ---- round 4: openings from the monomial forms (prover.rs:1501-1802) ----
    zc = tr.get_ext_challenge()
    w_n = R.omega(log_n)
    z_omega = R.e_mul_base(zc, w_n)
    mono = lambda cols: O.intt_n2n(np.asarray(cols, dtype=np.uint64))
    var_m, const_m, sig_m, st2_m = mono(variables), mono(constants), mono(sigmas), mono(st2)
    at_base = lambda m, at: S2.horner_ext(m, at)
    at_ext = lambda m0, m1, at: _combine(S2.horner_ext(m0, at), S2.horner_ext(m1, at))
    # order (prover.rs:1549-1683): variables, witness, constants, sigmas, z, partial products, multiplicities, lookup A, lookup B,
    # lookup tables, quotient chunks
    values_at_z = [at_base(var_m[j], zc) for j in range(V)] + [at_base(const_m[j], zc) for j in range(C)] \
        + [at_base(sig_m[j], zc) for j in range(V)] + [at_ext(st2_m[2 * i], st2_m[2 * i + 1], zc) for i in range(1 + n_partial)]
    values_at_0 = []
    if lk:
        tab_m, mult_m = mono(tables), mono(mult[None, :])
        values_at_z += [at_base(mult_m[0], zc)]
        values_at_z += [at_ext(st2_m[a_off + 2 * i], st2_m[a_off + 2 * i + 1], zc) for i in range(nsub + 1)]
        values_at_z += [at_base(tab_m[j], zc) for j in range(T)]
        values_at_0 = [(int(st2_m[a_off + 2 * i][0]), int(st2_m[a_off + 2 * i + 1][0])) for i in range(nsub + 1)]
    values_at_z += [at_ext(chunks[2 * i], chunks[2 * i + 1], zc) for i in range(Q)]
    values_at_z_omega = [at_ext(st2_m[0], st2_m[1], z_omega)]
    for v in values_at_z + values_at_z_omega + values_at_0:
        tr.witness_field_elements(v)

    # ---- round 5: DEEP combination + FRI (prover.rs:1828-2102) ----
    pi_groups = []                                                       # prover.rs:1805-1821, 2010-2041
    for (col, row), val in zip(public_inputs, public_values):
        at = pow(w_n, row, P)
        for g in pi_groups:
            if g[0] == at:
                g[1].append((col, val))
                break
        else:
            pi_groups.append((at, [(col, val)]))
    c = tr.get_ext_challenge()
    cm = lambda col: col[:ncm]
    sources = [(cm(x), None) for x in w_cols] + [(cm(x), None) for x in const_cols] + [(cm(x), None) for x in sigma_cols]
    sources += [(cm(s2_cols[2 * i]), cm(s2_cols[2 * i + 1])) for i in range(1 + n_partial)]
    zero_sources = []
    if lk:
        sources += [(cm(m_col), None)]
        zero_sources = [(cm(s2_cols[a_off + 2 * i]), cm(s2_cols[a_off + 2 * i + 1])) for i in range(nsub + 1)]
        sources += zero_sources
        sources += [(cm(x), None) for x in table_cols]
    sources += [(qt_cols[2 * i], qt_cols[2 * i + 1]) for i in range(Q)]
    assert len(sources) == len(values_at_z)
    ch = R.ext_powers(c, len(values_at_z) + 1 + len(values_at_0) + len(public_values))
    d0, d1 = np.zeros(ncm, np.uint64), np.zeros(ncm, np.uint64)
    d0, d1 = O.deep_group(d0, d1, sources, values_at_z, ch[:len(sources)], zc)
    d0, d1 = O.deep_group(d0, d1, [(cm(s2_cols[0]), cm(s2_cols[1]))], values_at_z_omega, ch[len(sources):len(sources) + 1], z_omega)
    off = len(sources) + 1
    if lk:
        d0, d1 = O.deep_group(d0, d1, zero_sources, values_at_0, ch[off:off + len(values_at_0)], (0, 0))
        off += len(values_at_0)
    for at, members in pi_groups:
        d0, d1 = O.deep_group(d0, d1, [(cm(w_cols[col]), None) for col, _ in members], [(val, 0) for _, val in members],
                              ch[off:off + len(members)], (at, 0))
        off += len(members)
    new_pow, num_queries, schedule, final_degree = R.compute_fri_schedule(security_level, cap_size, 0, log_l, log_n)
    assert new_pow == 0, "the oracle prover does not grind"
    fri = R.do_fri_oracle(d0, d1, tr, schedule, log_l, cap_size, hasher)

    # ---- queries (prover.rs:2161-2266) ----
    max_bits = log_n + log_l
    bools = R.BoolsBuffer(max_bits)
    answer = lambda cols, lh, lv, idx: {"leaf_elements": [int(col[idx]) for col in cols], "proof": _serde(O.merkle_path(lh, lv, idx), hasher)}
    queries = []
    for _ in range(num_queries):
        bits = bools.get_bits(tr, max_bits)
        idx = sum(b << i for i, b in enumerate(bits))
        q = {"witness_query": answer(w_oracle_cols, w_lh, w_lv, idx), "stage_2_query": answer(s2_cols, s2_lh, s2_lv, idx),
             "quotient_query": answer(qt_cols, qt_lh, qt_lv, idx), "setup_query": answer(setup_cols, setup_lh, setup_lv, idx)}
        fqs, sub = [], idx
        for lvl, k in enumerate(schedule):
            c0_l, c1_l = fri["levels"][lvl]
            lh, lv = fri["trees"][lvl]
            leaf, deg = sub >> k, 1 << k
            le = [int(x) for x in c0_l[leaf * deg:(leaf + 1) * deg]] + [int(x) for x in c1_l[leaf * deg:(leaf + 1) * deg]]
            fqs.append({"leaf_elements": le, "proof": _serde(O.merkle_path(lh, lv, leaf), hasher)})
            sub >>= k
        q["fri_queries"] = fqs
        queries.append(q)
    proof = {
        "proof_config": {"fri_lde_factor": L, "merkle_tree_cap_size": cap_size, "fri_folding_schedule": None,
                         "security_level": security_level, "pow_bits": 0},
        "public_inputs": public_values,
        "witness_oracle_cap": _serde(w_cap, hasher), "stage_2_oracle_cap": _serde(s2_cap, hasher), "quotient_oracle_cap": _serde(qt_cap, hasher),
        "final_fri_monomials": [fri["monomials"][0].tolist(), fri["monomials"][1].tolist()],
        "values_at_z": [_ext_dict(v) for v in values_at_z], "values_at_z_omega": [_ext_dict(v) for v in values_at_z_omega],
        "values_at_0": [_ext_dict(v) for v in values_at_0],
        "fri_base_oracle_cap": _serde(fri["caps"][0], hasher), "fri_intermediate_oracles_caps": [_serde(cp, hasher) for cp in fri["caps"][1:]],
        "queries_per_fri_repetition": queries, "pow_challenge": 0, "_marker": None,
    }
    return proof, setup_cap
