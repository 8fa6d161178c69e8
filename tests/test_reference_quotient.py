"""The quotient identity at z of Verifier::verify (src/cs/implementations/verifier.rs:1144-1828) on the reference's own
proof.json / vk.json (the fixture of src/gadgets/recursion/recursive_verifier.rs:2281-2368), with the verifier's gate
configuration of that test.  CPU only.  This pins, against reference-produced data, the gate evaluators of
era_boojum_b200/gate_library.py (the programs the device interpreter runs), the selector tree, the alpha-power order, the
lookup and copy-permutation relations and the quotient-chunk recombination."""
import copy

import pytest

from era_boojum_b200 import gate_library as GL
from oracle import verifier_reference as VR


def test_reference_proof_satisfies_the_quotient_identity_at_z(golden_fixture):
    assert VR.verify_quotient_at_z(golden_fixture)


def test_layout_matches_the_reference_proof(golden_fixture):
    fp, proof = golden_fixture["vk"]["fixed_parameters"], golden_fixture["proof"]
    lay = VR.circuit_layout(fp, VR.REFERENCE_FIXTURE_GATES)
    # 130 general-purpose + 8 x 3 lookup + 1 boolean column; 7 general-purpose constants + the table id
    assert (lay["num_variables"], lay["num_constants"], lay["num_multiplicities"]) == (155, 8, 1)
    q0 = proof["queries_per_fri_repetition"][0]
    assert len(q0["witness_query"]["leaf_elements"]) == lay["num_variables"] + lay["num_witnesses"] + lay["num_multiplicities"]
    assert len(q0["setup_query"]["leaf_elements"]) == lay["num_variables"] + lay["num_constants"] + 4
    n_partial = -(-lay["num_variables"] // 8) - 1
    assert len(proof["values_at_z"]) == 2 * lay["num_variables"] + lay["num_constants"] + 1 + n_partial + 1 + 8 + 1 + 4 + 8
    # 414 terms over general-purpose columns (the Poseidon2 flattened gate alone has 118)
    gp = VR.REFERENCE_FIXTURE_GATES["general_purpose"]
    assert sum(g.terms * g.num_repetitions_in_geometry(130, 0, 4) for g in gp) == 414


# position of one opening of every kind in values_at_z (layout above): variable (general purpose / lookup / boolean column),
# constant (selector / gate constant / table id), sigma, z, partial product, multiplicity, A, B, table column, quotient chunk
@pytest.mark.parametrize("pos", [0, 77, 129, 131, 154, 155, 158, 162, 163, 317, 318, 319, 337, 338, 339, 346, 347, 348, 351, 352, 359])
def test_a_flipped_opening_breaks_the_identity(golden_fixture, pos):
    bad = copy.deepcopy(golden_fixture["proof"])
    bad["values_at_z"][pos]["coeffs"][1] ^= 1
    assert not VR.verify_quotient_at_z(golden_fixture, proof=bad)


def test_flipped_z_omega_opening_breaks_the_identity(golden_fixture):
    bad = copy.deepcopy(golden_fixture["proof"])
    bad["values_at_z_omega"][0]["coeffs"][0] ^= 1
    assert not VR.verify_quotient_at_z(golden_fixture, proof=bad)


def test_every_gate_evaluator_is_pinned_by_the_identity(golden_fixture):
    """each evaluator contributes at z with a non-zero selector: perturbing ANY of them (one term off by one) must break
    the identity - so the fixture pins every transcription in gate_library.py, not only the gates a row happens to use."""
    cfg = VR.REFERENCE_FIXTURE_GATES

    def perturbed(gate):
        def ev(F, src, push, shared):
            first = []

            def push1(x):
                push(F.add(x, F.one()) if not first else x)
                first.append(1)
            gate.evaluate(F, src, push1, shared)
        g = copy.copy(gate)
        g.evaluate = ev
        return g

    for i, gate in enumerate(cfg["general_purpose"]):
        if gate.terms == 0:
            continue
        gp = list(cfg["general_purpose"])
        gp[i] = perturbed(gate)
        assert not VR.verify_quotient_at_z(golden_fixture, {"general_purpose": gp, "specialized": cfg["specialized"]}), gate.name
    spec = [(perturbed(cfg["specialized"][0][0]),) + tuple(cfg["specialized"][0][1:])]
    assert not VR.verify_quotient_at_z(golden_fixture, {"general_purpose": cfg["general_purpose"], "specialized": spec})
    # and the order of registration matters (alpha powers follow it)
    gp = list(cfg["general_purpose"])
    gp[3], gp[4] = gp[4], gp[3]
    assert not VR.verify_quotient_at_z(golden_fixture, {"general_purpose": gp, "specialized": cfg["specialized"]})


def test_recorded_programs_agree_with_direct_evaluation():
    """capture() (the SSA program handed to the device) replayed with Python ints equals the evaluator run directly over the
    base field, for every gate of the library, on random rows - the recorder itself is not a source of divergence."""
    import numpy as np
    from era_boojum_b200 import native as N
    rng = np.random.default_rng(3)
    gates = VR.REFERENCE_FIXTURE_GATES["general_purpose"] + [GL.BOOLEAN, GL.poseidon2_flattened_gate(100, 30)]
    P = GL.P
    for gate in gates:
        if gate.terms == 0:
            continue
        prog = GL.capture(gate)
        v, w, c = gate.width
        var = [int(x) % P for x in rng.integers(0, 2**63, size=v)]
        wit = [int(x) % P for x in rng.integers(0, 2**63, size=max(w, 1))]
        con = [int(x) % P for x in rng.integers(0, 2**63, size=max(c, gate.shared, 1))]
        want = GL.evaluate_gate_terms(gate, VR.BaseBackend, lambda i: var[i], lambda i: wit[i], lambda i: con[i], 1)
        tmp = {}

        def val(ix):
            kind, x = ix
            return {N.IDX_VARIABLE: lambda: var[x], N.IDX_WITNESS: lambda: wit[x], N.IDX_CONSTANT_POLY: lambda: con[x],
                    N.IDX_CONSTANT_POLY_SHARED: lambda: con[x], N.IDX_TEMPORARY: lambda: tmp[x], N.IDX_CONSTANT_VALUE: lambda: x}[kind]()
        for op, dst, a, b in prog["relations"]:
            x = val(a)
            tmp[dst] = {N.REL_ADD: lambda: (x + val(b)) % P, N.REL_SUB: lambda: (x - val(b)) % P, N.REL_MUL: lambda: x * val(b) % P,
                        N.REL_DOUBLE: lambda: 2 * x % P, N.REL_NEGATE: lambda: -x % P, N.REL_SQUARE: lambda: x * x % P}[op]()
        assert [val(wr) for wr in prog["writes"]] == want, gate.name
