"""ORACLE (test infrastructure): the "run verifier at z" block of Verifier::verify (src/cs/implementations/verifier.rs:1144-1828)
for circuits described the way the reference describes them - a VerificationKey (vk.json: geometry, lookup parameters,
selectors_placement tree, table-id column) plus the verifier's gate configuration (the ordered list of evaluators over
general-purpose columns and over specialised columns, cs_builder_verifier.rs:76-250).

Together with oracle/replay.py::replay_proof (transcript, query indices, Merkle paths, DEEP, FRI) this is a complete
restatement of Verifier::verify for the reference's own fixture proof.json / vk.json
(src/gadgets/recursion/recursive_verifier.rs:2281-2368): `verify_quotient_at_z(fixture, REFERENCE_FIXTURE_GATES)` must hold.
It pins, against reference-produced data: the selector tree semantics, the order of the alpha powers (lookup, specialised
gates, general-purpose gates, z(1) = 1, copy-permutation chunks), the lookup relations, the copy-permutation relations with the
non-residues, the quotient-chunk recombination, and every gate evaluator of era_boojum_b200/gate_library.py (each evaluator
contributes at z with a non-zero selector, so a wrong formula in any of them breaks the identity).

Values are Python ints; Fp2 elements are (c0, c1) tuples.
"""
from era_boojum_b200 import gate_library as GL
from era_boojum_b200 import placement as PL

from . import replay as R
from .stage2 import non_residues_for_copy_permutation

P = R.P


class Fp2Backend:
    """PrimeFieldLike over ExtensionField<F, 2, EXT> with base-field constants (what VerifierPolyStorage feeds the evaluators)."""

    @staticmethod
    def zero():
        return (0, 0)

    @staticmethod
    def one():
        return (1, 0)

    @staticmethod
    def constant(v):
        return (v % P, 0)

    add = staticmethod(R.e_add)
    sub = staticmethod(R.e_sub)
    mul = staticmethod(R.e_mul)

    @staticmethod
    def double(a):
        return R.e_add(a, a)

    @staticmethod
    def negate(a):
        return R.e_sub((0, 0), a)

    @staticmethod
    def square(a):
        return R.e_mul(a, a)


class BaseBackend:
    """PrimeFieldLike over the base field (row-wise evaluation on plain trace values)."""

    @staticmethod
    def zero():
        return 0

    @staticmethod
    def one():
        return 1

    @staticmethod
    def constant(v):
        return v % P

    add = staticmethod(lambda a, b: (a + b) % P)
    sub = staticmethod(lambda a, b: (a - b) % P)
    mul = staticmethod(lambda a, b: a * b % P)
    double = staticmethod(lambda a: 2 * a % P)
    negate = staticmethod(lambda a: (-a) % P)
    square = staticmethod(lambda a: a * a % P)


# the verifier configuration of test_recursive_verification for proof.json / vk.json (recursive_verifier.rs:2290-2368):
# evaluators over general-purpose columns in registration order, gates with the same evaluator (and parameters) merged
# (cs_builder_verifier.rs:103-147): UIntXAddGate<32/16/8> share one evaluator, PublicInputGate and NopGate share the NOP one.
REFERENCE_FIXTURE_GATES = {
    "general_purpose": [GL.CONSTANT_ALLOCATOR, GL.U8X4_FMA, GL.poseidon2_flattened_gate(130, 0), GL.DOT_PRODUCT4, GL.ZERO_CHECK,
                        GL.FMA, GL.UINTX_ADD, GL.SELECTION, GL.PARALLEL_SELECTION4, GL.NOP, GL.REDUCTION4],
    # after the lookup's own specialised columns (allow_lookup registers first): (gate, num_repetitions, share_constants)
    "specialized": [(GL.BOOLEAN, 1, False)],
}


def _ext(d):
    return (int(d["coeffs"][0]), int(d["coeffs"][1]))


def circuit_layout(fp, gates):
    """Sizes the verifier derives from the VerificationKey and its gate configuration (verifier.rs:661-735, 805-886)."""
    geo = fp["parameters"]
    lk = fp["lookup_parameters"]
    lookup = None
    if lk != "NoLookup":
        (kind, pars), = lk.items()
        assert kind == "UseSpecializedColumnsWithTableIdAsConstant", "only the bench / fixture lookup mode is restated"
        lookup = pars
    gp_vars = geo["num_columns_under_copy_permutation"]
    spec_vars, spec_consts = 0, 0
    lookup_offset = None
    if lookup:
        lookup_offset = gp_vars
        spec_vars += lookup["width"] * lookup["num_repetitions"]
        spec_consts += 1                                    # the shared table-id constant
    spec = []
    for gate, reps, share in gates["specialized"]:
        v, w, c = gate.width
        assert w == 0
        spec.append(dict(gate=gate, reps=reps, share=share, var_base=gp_vars + spec_vars, const_base=spec_consts))
        spec_vars += v * reps
        spec_consts += c if share else c * reps
    consts_gp = fp["extra_constant_polys_for_selectors"] + geo["num_constant_columns"]
    n = fp["domain_size"]
    n_mult = 0
    if lookup:
        n_mult = max(1, -(-fp["total_tables_len"] // n))
    return dict(num_variables=gp_vars + spec_vars, num_witnesses=geo["num_witness_columns"], num_constants=consts_gp + spec_consts,
                consts_gp=consts_gp, gp_vars=gp_vars, lookup=lookup, lookup_offset=lookup_offset, specialized=spec,
                num_multiplicities=n_mult, quotient_degree=fp["quotient_degree"])


def quotient_terms_at_z(fp, gates, values_at_z, values_at_z_omega, values_at_0, z, alpha, beta, gamma, lookup_beta, lookup_gamma):
    """t_accumulator and t_from_chunks of verifier.rs:1144-1808 (Fp2 tuples)."""
    lay = circuit_layout(fp, gates)
    V, W, C, Q = lay["num_variables"], lay["num_witnesses"], lay["num_constants"], lay["quotient_degree"]
    n = fp["domain_size"]
    lookup = lay["lookup"]
    n_sub = lookup["num_repetitions"] if lookup else 0
    n_mult = lay["num_multiplicities"]
    n_tab = (lookup["width"] + 1) if lookup else 0
    n_partial = 0 if V <= Q else -(-V // Q) - 1            # num_intermediate_partial_product_relations

    it = iter(values_at_z)
    take = lambda k: [next(it) for _ in range(k)]
    variables, witnesses, constants, sigmas = take(V), take(W), take(C), take(V)
    z_at_z = next(it)
    partials = take(n_partial)
    multiplicities, a_polys, b_polys, tables = take(n_mult), take(n_sub), take(n_mult), take(n_tab)
    chunks = list(it)
    assert len(chunks) == Q, (len(chunks), Q)
    z_at_z_omega = values_at_z_omega[0]

    gp = gates["general_purpose"]
    tree = fp["selectors_placement"]
    gp_terms = [g.terms * g.num_repetitions_in_geometry(lay["gp_vars"], W, fp["parameters"]["num_constant_columns"]) for g in gp]
    n_lookup_terms = (n_sub + n_mult) if lookup else 0
    n_spec_terms = sum(s["gate"].terms * s["reps"] for s in lay["specialized"])
    total = n_lookup_terms + n_spec_terms + sum(gp_terms) + 2 + n_partial
    powers = R.ext_powers(alpha, total)                     # materialize_powers_serial: 1, alpha, alpha^2, ...
    ch_lookup, rest = powers[:n_lookup_terms], powers[n_lookup_terms:]
    ch_spec, rest = rest[:n_spec_terms], rest[n_spec_terms:]
    ch_gp, ch_rest = rest[:sum(gp_terms)], rest[sum(gp_terms):]

    F = Fp2Backend
    t = (0, 0)
    # ---- lookup (verifier.rs:1236-1522), specialised columns, table id in a constant ----
    if lookup:
        assert R.e_sub(_sum(values_at_0[:n_sub]), _sum(values_at_0[n_sub:])) == (0, 0), "lookup sumcheck"
        width = lookup["width"]
        gp_pows = [(1, 0)]
        for _ in range(width):
            gp_pows.append(R.e_mul(gp_pows[-1], lookup_gamma))
        agg_tables = lookup_beta
        for g, col in zip(gp_pows, tables):
            agg_tables = R.e_add(agg_tables, R.e_mul(g, col))
        table_id = [constants[fp["table_ids_column_idxes"][0]]] if fp["table_ids_column_idxes"] else []
        ch = iter(ch_lookup)
        cols = variables[lay["lookup_offset"]: lay["lookup_offset"] + width * n_sub]
        for i, a_poly in enumerate(a_polys):
            contribution = lookup_beta
            for g, col in zip(gp_pows, cols[i * width:(i + 1) * width] + table_id):
                contribution = R.e_add(contribution, R.e_mul(g, col))
            contribution = R.e_sub(R.e_mul(contribution, a_poly), (1, 0))
            t = R.e_add(t, R.e_mul(contribution, next(ch)))
        for b_poly, m in zip(b_polys, multiplicities):
            contribution = R.e_sub(R.e_mul(agg_tables, b_poly), m)
            t = R.e_add(t, R.e_mul(contribution, next(ch)))
    # ---- gates over specialised columns (verifier.rs:1535-1651): no selector ----
    k = 0
    for s in lay["specialized"]:
        terms = GL.evaluate_gate_terms(s["gate"], F, lambda i: variables[i], lambda i: witnesses[i], lambda i: constants[i], s["reps"],
                                       var_base=s["var_base"], wit_base=0, const_base=lay["consts_gp"] + s["const_base"])
        for term in terms:
            t = R.e_add(t, R.e_mul(term, ch_spec[k]))
            k += 1
    assert k == n_spec_terms
    # ---- gates over general-purpose columns (verifier.rs:1653-1716) ----
    k = 0
    for gate_idx, (gate, n_terms) in enumerate(zip(gp, gp_terms)):
        path = PL.output_placement(tree, gate_idx)
        if n_terms == 0:
            continue                                         # NOP / markers: a place in the tree, nothing to evaluate
        assert path is not None, "gate %s has terms but no selector" % gate.name
        selector = (1, 0)
        for depth, bit in enumerate(path):
            selector = R.e_mul(selector, constants[depth] if bit else R.e_sub((1, 0), constants[depth]))
        reps = n_terms // gate.terms
        terms = GL.evaluate_gate_terms(gate, F, lambda i: variables[i], lambda i: witnesses[i], lambda i: constants[i], reps,
                                       const_base=len(path))
        acc = (0, 0)
        for term in terms:
            acc = R.e_add(acc, R.e_mul(term, ch_gp[k]))
            k += 1
        t = R.e_add(t, R.e_mul(acc, selector))               # VerifierRelationDestination multiplies every term by the selector
    assert k == sum(gp_terms)
    # ---- copy permutation (verifier.rs:1718-1789) ----
    z_n = _pow(z, n)
    vanishing = R.e_sub(z_n, (1, 0))
    ch = iter(ch_rest)
    l1 = R.e_mul(vanishing, R.e_inv(R.e_sub(z, (1, 0))))
    t = R.e_add(t, R.e_mul(R.e_mul(R.e_sub(z_at_z, (1, 0)), l1), next(ch)))
    non_res = non_residues_for_copy_permutation(n, V)
    lhs_seq = partials + [z_at_z_omega]
    rhs_seq = [z_at_z] + partials
    for c, (lhs, rhs) in enumerate(zip(lhs_seq, rhs_seq)):
        sl = slice(c * Q, min((c + 1) * Q, V))
        for var, sig in zip(variables[sl], sigmas[sl]):
            lhs = R.e_mul(lhs, R.e_add(R.e_add(R.e_mul(sig, beta), var), gamma))
        for nr, var in zip(non_res[sl], variables[sl]):
            rhs = R.e_mul(rhs, R.e_add(R.e_add(R.e_mul(R.e_mul_base(z, nr), beta), var), gamma))
        t = R.e_add(t, R.e_mul(R.e_sub(lhs, rhs), next(ch)))
    assert next(ch, None) is None, "must exhaust all the challenges"
    # ---- quotient chunks (verifier.rs:1791-1808) ----
    from_chunks, pw = (0, 0), (1, 0)
    for el in chunks:
        from_chunks = R.e_add(from_chunks, R.e_mul(el, pw))
        pw = R.e_mul(pw, z_n)
    return t, R.e_mul(from_chunks, vanishing)


def _sum(vals):
    acc = (0, 0)
    for v in vals:
        acc = R.e_add(acc, v)
    return acc


def _pow(a, e):
    r = (1, 0)
    while e:
        if e & 1:
            r = R.e_mul(r, a)
        a = R.e_mul(a, a)
        e >>= 1
    return r


def challenges_from_transcript(fixture):
    """beta, gamma, lookup_beta, lookup_gamma, alpha, z of the proof (verifier.rs:907-1062; Poseidon2 transcript)."""
    vk, proof = fixture["vk"], fixture["proof"]
    tr = R.Poseidon2Transcript()
    tr.witness_merkle_tree_cap(vk["setup_merkle_tree_cap"])
    for v in proof["public_inputs"]:
        tr.witness_field_elements([v])
    tr.witness_merkle_tree_cap(proof["witness_oracle_cap"])
    beta, gamma = tr.get_ext_challenge(), tr.get_ext_challenge()
    lookup_beta = lookup_gamma = (0, 0)
    if vk["fixed_parameters"]["lookup_parameters"] != "NoLookup":
        lookup_beta, lookup_gamma = tr.get_ext_challenge(), tr.get_ext_challenge()
    tr.witness_merkle_tree_cap(proof["stage_2_oracle_cap"])
    alpha = tr.get_ext_challenge()
    tr.witness_merkle_tree_cap(proof["quotient_oracle_cap"])
    z = tr.get_ext_challenge()
    return dict(beta=beta, gamma=gamma, lookup_beta=lookup_beta, lookup_gamma=lookup_gamma, alpha=alpha, z=z)


def verify_quotient_at_z(fixture, gates=None, proof=None):
    """True iff the quotient identity holds at z for the fixture's proof (or `proof`, same vk)."""
    gates = gates or REFERENCE_FIXTURE_GATES
    fx = dict(fixture)
    if proof is not None:
        fx["proof"] = proof
    pf = fx["proof"]
    ch = challenges_from_transcript(fx)
    t, want = quotient_terms_at_z(fx["vk"]["fixed_parameters"], gates, [_ext(v) for v in pf["values_at_z"]],
                                  [_ext(v) for v in pf["values_at_z_omega"]], [_ext(v) for v in pf["values_at_0"]],
                                  ch["z"], ch["alpha"], ch["beta"], ch["gamma"], ch["lookup_beta"], ch["lookup_gamma"])
    return t == want
