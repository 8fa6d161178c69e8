"""Host driver of the five-round IOP over the B200 kernels, mirroring CSReferenceAssembly::prove_cpu_basic
(src/cs/implementations/prover.rs:153-2269) for circuits whose gates live on general-purpose columns (no lookup argument,
no specialised columns, no public inputs yet).  Every heavy step is a C-ABI call into libboojum_b200.so; the transcript,
the FRI schedule and the proof assembly stay on the host, as in the reference.  The proof is returned in the reference's
serde shape (src/cs/implementations/proof.rs:57-143).

    setup = Setup.from_columns(ctx, sigmas, constants, gates, ...)      # one-off (setup.rs:1093-1255 role)
    proof = prove(ctx, setup, variables, config)                       # prove_cpu_basic role
"""
import time

import numpy as np

from . import Transcript, to_numpy

P = 0xFFFFFFFF00000001


def e_mul(a, b):
    return ((a[0] * b[0] + 7 * a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def e_mul_base(a, b):
    return (a[0] * b % P, a[1] * b % P)


def _omega(log_n):
    w = 0x185629DCDA58878C
    for _ in range(log_n, 32):
        w = w * w % P
    return w


def _ext_dict(v):
    return {"coeffs": [int(v[0]), int(v[1])], "_marker": None}


def _combine_ext(a, b):
    """value of f0 + u f1 at an Fp2 point from the values a = f0(at), b = f1(at) (u^2 = 7)."""
    return ((a[0] + 7 * b[1]) % P, (a[1] + b[0]) % P)


class ProofConfig:
    """ProofConfig (prover.rs:55-73)."""

    def __init__(self, fri_lde_factor=8, merkle_tree_cap_size=16, security_level=100, pow_bits=0):
        self.fri_lde_factor, self.merkle_tree_cap_size = fri_lde_factor, merkle_tree_cap_size
        self.security_level, self.pow_bits = security_level, pow_bits

    def to_dict(self):
        return {"fri_lde_factor": self.fri_lde_factor, "merkle_tree_cap_size": self.merkle_tree_cap_size,
                "fri_folding_schedule": None, "security_level": self.security_level, "pow_bits": self.pow_bits}


class Setup:
    """SetupStorage + setup Merkle tree + the fixed parameters of the VerificationKey (setup.rs:1093-1255,
    verifier.rs:31-79): sigma and constant columns, their LDEs, the tree over [sigmas | constants]."""

    def __init__(self, ctx, sigmas, constants, gates, quotient_degree, config, lookup=None):
        """lookup: None or dict(width, num_repetitions, variables_offset, table_id_column (index into constants),
        tables=[width + 1, n] tensor of lookup-table setup columns, multiplicities are part of the witness)."""
        torch = ctx._torch
        self.gates = gates                      # [dict(name, program..., num_repetitions, selector_path, ...)]
        self.quotient_degree = quotient_degree
        self.config = config
        self.num_variables = sigmas.shape[0]
        self.num_constants = constants.shape[0]
        self.log_n = sigmas.shape[1].bit_length() - 1
        self.sigmas = sigmas                    # [V, n] natural order (needed by stage 2)
        self.constants_natural = constants      # [C, n] natural order (table-id column of the lookup argument)
        L = config.fri_lde_factor
        self.lookup = lookup
        parts = [sigmas, constants] + ([lookup["tables"]] if lookup else [])
        cols = torch.cat(parts, dim=0).contiguous()
        self.lde = ctx.transform_raw_storages_to_lde(cols, L)      # [V + C, L, n]
        self.tree = ctx.merkle_tree_construct([self.lde[c].reshape(-1) for c in range(cols.shape[0])],
                                              config.merkle_tree_cap_size)
        self.cap = self.tree.get_cap()

    def sigma_lde(self, j):
        return self.lde[j].reshape(-1)

    def constant_lde(self, j):
        return self.lde[self.num_variables + j].reshape(-1)

    def table_lde(self, j):
        return self.lde[self.num_variables + self.num_constants + j].reshape(-1)

    def vk(self):
        lk = None
        if self.lookup:
            lk = {k: self.lookup[k] for k in ("width", "num_repetitions", "variables_offset", "table_id_column")}
        return {"lookup": lk, "domain_size": 1 << self.log_n, "num_variables": self.num_variables, "num_constants": self.num_constants,
                "quotient_degree": self.quotient_degree, "fri_lde_factor": self.config.fri_lde_factor,
                "cap_size": self.config.merkle_tree_cap_size,
                "gates": [(g["name"], g["num_repetitions"], list(g["selector_path"])) for g in self.gates],
                "setup_merkle_tree_cap": self.cap.tolist()}


def prove(ctx, setup, variables, timings=None, multiplicities=None):
    """variables: [V, n] int64 CUDA tensor (copy-permutation columns incl. the lookup sub-argument columns, natural row
    order); multiplicities: [n] tensor when the setup has a lookup argument.  Returns the proof dict."""
    torch = ctx._torch
    cfg = setup.config
    L, cap = cfg.fri_lde_factor, cfg.merkle_tree_cap_size
    log_L = L.bit_length() - 1
    V, C, Q = setup.num_variables, setup.num_constants, setup.quotient_degree
    log_n, log_q = setup.log_n, Q.bit_length() - 1
    n = 1 << log_n
    dev = variables.device
    tm = timings if timings is not None else {}

    def mark(name, t0):
        if timings is not None:
            torch.cuda.synchronize()
            tm[name] = tm.get(name, 0.0) + time.perf_counter() - t0

    flat = lambda t: t.reshape(-1)
    tr = Transcript()
    tr.witness_merkle_tree_cap(setup.cap)                                   # prover.rs:211
    # ---- round 1: witness commitment (prover.rs:313-353) ----
    t0 = time.perf_counter()
    lk = setup.lookup
    w_lde = ctx.transform_raw_storages_to_lde(variables, L)                  # [V, L, n]
    w_cols = [flat(w_lde[c]) for c in range(V)]
    m_col = None
    if lk:
        m_lde = ctx.transform_raw_storages_to_lde(multiplicities.reshape(1, -1).contiguous(), L)
        m_col = flat(m_lde[0])
    w_tree = ctx.merkle_tree_construct(w_cols + ([m_col] if lk else []), cap)   # variables | witness (none) | multiplicities
    w_cap = w_tree.get_cap()
    tr.witness_merkle_tree_cap(w_cap)
    mark("1_witness_lde_commit", t0)
    # ---- round 2: copy-permutation products (prover.rs:360-554) ----
    t0 = time.perf_counter()
    beta = tr.get_multiple_challenges_fixed(2)
    gamma = tr.get_multiple_challenges_fixed(2)
    if lk:
        lookup_beta = tr.get_multiple_challenges_fixed(2)                    # prover.rs:402-406
        lookup_gamma = tr.get_multiple_challenges_fixed(2)
    z0, z1, partials = ctx.compute_partial_products_in_extension([variables[c] for c in range(V)],
                                                                 [setup.sigmas[c] for c in range(V)], beta, gamma, Q)
    lk_polys = []
    if lk:
        wdt, nsub, voff = lk["width"], lk["num_repetitions"], lk["variables_offset"]
        a_polys, b_poly = ctx.compute_lookup_poly_pairs_specialized(
            [variables[voff + i] for i in range(wdt * nsub)], wdt, setup.constants_natural[lk["table_id_column"]],
            [lk["tables"][j] for j in range(wdt + 1)], multiplicities, lookup_beta, lookup_gamma)
        lk_polys = [t for pr in a_polys for t in pr] + [b_poly[0], b_poly[1]]
    st2 = torch.stack([z0, z1] + [t for pr in partials for t in pr] + lk_polys).contiguous()
    s2_lde = ctx.transform_raw_storages_to_lde(st2, L)
    s2_cols = [flat(s2_lde[c]) for c in range(st2.shape[0])]
    s2_tree = ctx.merkle_tree_construct(s2_cols, cap)
    s2_cap = s2_tree.get_cap()
    tr.witness_merkle_tree_cap(s2_cap)
    n_partial = len(partials)
    mark("2_stage2_products_lde_commit", t0)
    # ---- round 3: quotient (prover.rs:560-1495) ----
    t0 = time.perf_counter()
    alpha = tr.get_multiple_challenges_fixed(2)
    n_gate_terms = sum(len(g["writes"]) * g["num_repetitions"] for g in setup.gates)
    n_lk_terms = (lk["num_repetitions"] + 1) if lk else 0      # lookup terms come first (prover.rs:608-625)
    total_terms = n_lk_terms + n_gate_terms + 1 + 1 + n_partial
    powers = [(1, 0)]
    for _ in range(1, total_terms):
        powers.append(e_mul(powers[-1], alpha))
    npts = n * Q
    q0 = torch.zeros(npts, dtype=torch.int64, device=dev)
    q1 = torch.zeros(npts, dtype=torch.int64, device=dev)
    const_cols = [setup.constant_lde(j) for j in range(C)]
    a_off = 2 + 2 * n_partial
    if lk:
        a_ldes = [(s2_cols[a_off + 2 * i], s2_cols[a_off + 2 * i + 1]) for i in range(nsub)]
        b_lde = (s2_cols[a_off + 2 * nsub], s2_cols[a_off + 2 * nsub + 1])
        ctx.quotient_lookup_specialized([w_cols[voff + i] for i in range(wdt * nsub)], wdt, const_cols[lk["table_id_column"]],
                                        [setup.table_lde(j) for j in range(wdt + 1)], m_col, a_ldes, b_lde, lookup_beta, lookup_gamma,
                                        powers[:n_lk_terms], q0, q1)
    ctx.evaluate_gates_over_general_purpose_columns(setup.gates, w_cols, [], const_cols,
                                                    powers[n_lk_terms:n_lk_terms + n_gate_terms], q0, q1)
    part_ldes = [(s2_cols[2 + 2 * c], s2_cols[3 + 2 * c]) for c in range(n_partial)]
    ctx.quotient_copy_permutation(w_cols, [setup.sigma_lde(j) for j in range(V)], (s2_cols[0], s2_cols[1]), part_ldes, beta, gamma,
                                  powers[n_lk_terms + n_gate_terms:], log_n, log_L, log_q, Q, q0, q1)
    ctx.divide_by_vanishing(q0, q1, log_n, log_q)
    # flatten the cosets into natural order, interpolate once at size n*Q on coset 7, split into Q chunks (prover.rs:1399-1467)
    qq = torch.stack([q0, q1]).contiguous()
    ctx.bitreverse_enumeration_inplace(qq)
    ctx.ifft_natural_to_natural(qq, 7)
    # the reference's satisfiability guard: the top coefficient of the interpolant must vanish (prover.rs:1425-1438)
    top = to_numpy(qq[:, npts - 1])
    if int(top[0]) != 0 or int(top[1]) != 0:
        raise ValueError("unsatisfied: quotient is not a polynomial of degree < n * quotient_degree")
    chunks = torch.stack([qq[k][j * n:(j + 1) * n] for j in range(Q) for k in (0, 1)]).contiguous()   # c0,c1 of chunk 0, ...
    qt_lde = ctx.transform_raw_storages_to_lde(chunks, L, from_monomials=True)
    qt_cols = [flat(qt_lde[c]) for c in range(2 * Q)]
    qt_tree = ctx.merkle_tree_construct(qt_cols, cap)
    qt_cap = qt_tree.get_cap()
    tr.witness_merkle_tree_cap(qt_cap)
    mark("3_quotient", t0)
    # ---- round 4: openings (prover.rs:1501-1802) ----
    t0 = time.perf_counter()
    z = tr.get_multiple_challenges_fixed(2)
    w_n = _omega(log_n)
    z_omega = e_mul_base(z, w_n)
    # opening order (prover.rs:1549-1683): variables, witness, constants, sigmas, z, partial products, multiplicities,
    # lookup A, lookup B, lookup tables, quotient chunks.  `sources` keeps (c0, c1-or-None) per opened polynomial.
    sources = [(c, None) for c in w_cols] + [(c, None) for c in const_cols] + [(setup.sigma_lde(j), None) for j in range(V)]
    sources += [(s2_cols[2 * i], s2_cols[2 * i + 1]) for i in range(1 + n_partial)]
    if lk:
        sources += [(m_col, None)]
        sources += [(s2_cols[a_off + 2 * i], s2_cols[a_off + 2 * i + 1]) for i in range(nsub + 1)]
        sources += [(setup.table_lde(j), None) for j in range(wdt + 1)]
    sources += [(qt_cols[2 * i], qt_cols[2 * i + 1]) for i in range(Q)]

    def open_at(srcs, at):
        flat_cols, spans = [], []
        for c0, c1 in srcs:
            spans.append((len(flat_cols), c1 is not None))
            flat_cols += [c0] + ([c1] if c1 is not None else [])
        ev = ctx.barycentric_evaluate(flat_cols, log_n, at)
        return [(_combine_ext(ev[i], ev[i + 1]) if is_ext else ev[i]) for i, is_ext in spans]

    values_at_z = open_at(sources, z)
    values_at_z_omega = open_at([(s2_cols[0], s2_cols[1])], z_omega)
    zero_sources = [(s2_cols[a_off + 2 * i], s2_cols[a_off + 2 * i + 1]) for i in range(nsub + 1)] if lk else []
    values_at_0 = open_at(zero_sources, (0, 0)) if lk else []
    for v in values_at_z + values_at_z_omega + values_at_0:
        tr.witness_field_elements(v)
    mark("4_openings", t0)
    # ---- round 5: DEEP + FRI (prover.rs:1828-2102) ----
    t0 = time.perf_counter()
    c = tr.get_multiple_challenges_fixed(2)
    n_ch = len(values_at_z) + 1 + len(values_at_0)
    ch = [(1, 0), c]
    for _ in range(2, n_ch):
        ch.append(e_mul(ch[-1], c))
    deep0 = torch.zeros(n * L, dtype=torch.int64, device=dev)
    deep1 = torch.zeros(n * L, dtype=torch.int64, device=dev)
    ctx.quotening_operation_in_extension(deep0, deep1, sources, values_at_z, z, ch[:len(sources)])
    ctx.quotening_operation_in_extension(deep0, deep1, [(s2_cols[0], s2_cols[1])], values_at_z_omega, z_omega,
                                         ch[len(sources):len(sources) + 1])
    if lk:
        ctx.quotening_operation_in_extension(deep0, deep1, zero_sources, values_at_0, (0, 0), ch[len(sources) + 1:])
    import ctypes
    from .native import lib
    np_, nq, sl, fd = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32()
    sched = (ctypes.c_uint32 * 32)()
    assert lib.bj_compute_fri_schedule(cfg.security_level, cap, cfg.pow_bits, log_L, log_n, ctypes.byref(np_), ctypes.byref(nq),
                                       sched, ctypes.byref(sl), ctypes.byref(fd)) == 0
    assert np_.value == 0, "PoW is not implemented (benches use NoPow)"
    schedule = list(sched[: sl.value])
    fri = ctx.do_fri(tr, deep0, deep1, schedule, L, cap)
    mono0, mono1 = fri.monomial_forms()
    mark("5_deep_fri", t0)
    # ---- queries (prover.rs:2161-2266) ----
    t0 = time.perf_counter()
    max_bits = log_n + log_L
    idxs = [tr.get_index_bits(max_bits, max_bits) for _ in range(nq.value)]
    setup_cols = [setup.lde[c].reshape(-1) for c in range(setup.lde.shape[0])]
    oracles = [("witness_query", w_cols + ([m_col] if lk else []), w_tree), ("stage_2_query", s2_cols, s2_tree), ("quotient_query", qt_cols, qt_tree),
               ("setup_query", setup_cols, setup.tree)]
    rows = {name: (ctx.query_leaf_elements(cols, idxs), ctx.merkle_paths(tree, idxs)) for name, cols, tree in oracles}
    queries = []
    for qi, idx in enumerate(idxs):
        q = {name: {"leaf_elements": rows[name][0][qi].tolist(), "proof": rows[name][1][qi].tolist()} for name, _, _ in oracles}
        fqs, sub = [], idx
        for lvl, k in enumerate(schedule):
            le, path = fri.query(lvl, sub >> k, k)
            fqs.append({"leaf_elements": le.tolist(), "proof": path.tolist()})
            sub >>= k
        q["fri_queries"] = fqs
        queries.append(q)
    mark("6_queries", t0)
    return {
        "proof_config": cfg.to_dict(), "public_inputs": [],
        "witness_oracle_cap": w_cap.tolist(), "stage_2_oracle_cap": s2_cap.tolist(), "quotient_oracle_cap": qt_cap.tolist(),
        "final_fri_monomials": [mono0.tolist(), mono1.tolist()],
        "values_at_z": [_ext_dict(v) for v in values_at_z], "values_at_z_omega": [_ext_dict(v) for v in values_at_z_omega],
        "values_at_0": [_ext_dict(v) for v in values_at_0],
        "fri_base_oracle_cap": fri.get_cap(0).tolist(),
        "fri_intermediate_oracles_caps": [fri.get_cap(i).tolist() for i in range(1, fri.num_oracles())],
        "queries_per_fri_repetition": queries, "pow_challenge": 0, "_marker": None,
    }
