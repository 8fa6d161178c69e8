"""The oracle's own CPU prover (oracle/prover.py: prove_cpu_basic restated end to end on the oracle primitives, openings by
Horner from monomials, quotient point by point) and the oracle verifier agree - the pair the GPU library's proofs are compared
with bit for bit in tests/test_gpu_prove.py::test_proof_equals_the_cpu_oracle_prover."""
import json

import pytest

from oracle import circuits, prover as OP, verifier as OV


def vk_of(c, L, cap, setup_cap, public_inputs=(), hasher="poseidon2", transcript="poseidon2"):
    lk = c["lookup"]
    return {"domain_size": c["variables"].shape[1], "num_variables": c["variables"].shape[0], "num_constants": c["constants"].shape[0],
            "quotient_degree": c["quotient_degree"], "fri_lde_factor": L, "cap_size": cap,
            "gates": [(name, reps, path, 0, len(path)) for name, reps, path in c["gates"]],
            "lookup": {k: lk[k] for k in ("width", "num_repetitions", "variables_offset", "table_id_column")} if lk else None,
            "public_inputs_locations": [list(p) for p in public_inputs], "hasher": hasher, "transcript": transcript,
            "setup_merkle_tree_cap": setup_cap.tolist()}


@pytest.mark.parametrize("log_n,V,L,cap,lookup,pis", [(5, 20, 8, 16, False, ()), (6, 20, 4, 8, True, ((1, 3), (5, 3), (0, 9))),
                                                        (6, 20, 2, 8, True, ())])
def test_cpu_prover_and_verifier_agree(log_n, V, L, cap, lookup, pis):
    c = circuits.sha_shaped(log_n, V, seed=log_n, lookup=lookup)
    proof, setup_cap = OP.prove(c["variables"], c["sigmas"], c["constants"], c["gates"], c["quotient_degree"], L, cap,
                                lookup=c["lookup"], public_inputs=pis)
    vk = vk_of(c, L, cap, setup_cap, pis)
    assert OV.verify(vk, proof)
    bad = json.loads(json.dumps(proof))
    bad["values_at_z"][2]["coeffs"][0] ^= 1
    with pytest.raises(AssertionError):
        OV.verify(vk, bad)


@pytest.mark.parametrize("hasher,transcript", [("blake2s", "blake2s"), ("poseidon2", "poseidon"), ("keccak256", "keccak256")])
def test_cpu_prover_other_type_parameters(hasher, transcript):
    """the bench's second pair (Blake2s256 tree + Blake2sTranscript, sha256/mod.rs:264-271), the recursive-mode pair (Poseidon2
    tree + Poseidon transcript, :286-293) and Keccak256"""
    c = circuits.sha_shaped(5, 20, seed=3, lookup=True)
    pis = ((0, 2),)
    proof, setup_cap = OP.prove(c["variables"], c["sigmas"], c["constants"], c["gates"], c["quotient_degree"], 4, 8, lookup=c["lookup"],
                                public_inputs=pis, hasher=hasher, transcript=transcript)
    vk = vk_of(c, 4, 8, setup_cap, pis, hasher, transcript)
    assert OV.verify(vk, proof)
    if hasher != "poseidon2":
        assert len(proof["witness_oracle_cap"][0]) == 32          # [u8; 32] digests
    bad = json.loads(json.dumps(proof))
    bad["queries_per_fri_repetition"][0]["witness_query"]["leaf_elements"][0] ^= 1
    with pytest.raises(AssertionError):
        OV.verify(vk, bad)


def test_cpu_prover_refuses_an_unsatisfied_witness():
    c = circuits.sha_shaped(5, 20, seed=1)
    w = c["variables"].copy()
    w[7, 11] ^= 1
    with pytest.raises((ValueError, AssertionError)):
        OP.prove(w, c["sigmas"], c["constants"], c["gates"], c["quotient_degree"], 8, 16)


def production_gates():
    """the 11 gates of the reference's vk.json as placed recorded programs: (GPU-side dicts, oracle-side tuples)"""
    from era_boojum_b200 import gate_library as GL
    gp = [GL.CONSTANT_ALLOCATOR, GL.U8X4_FMA, GL.poseidon2_flattened_gate(130, 0), GL.DOT_PRODUCT4, GL.ZERO_CHECK, GL.FMA, GL.UINTX_ADD,
          GL.SELECTION, GL.PARALLEL_SELECTION4, GL.NOP, GL.REDUCTION4]
    dicts = [GL.placed(GL.BOOLEAN, 1, [], constants_placement_offset=8, variables_initial_offset=154)]
    for gate, path in zip(gp, circuits.PRODUCTION_PATHS):
        if gate.terms:
            dicts.append(GL.placed(gate, gate.num_repetitions_in_geometry(130, 0, 4), path))
    tuples = [(g["name"], g["num_repetitions"], g["selector_path"], g["variables_initial_offset"], g["constants_placement_offset"], g)
              for g in dicts]
    return dicts, tuples


def test_cpu_prover_production_shaped_circuit():
    """the geometry of the reference's vk.json (11 gates as recorded programs incl. the Poseidon2 flattened gate and the
    specialised boolean gate, lookups of width 3, quotient degree 8 over LDE factor 2) through the CPU prover and the verifier"""
    c = circuits.production_shaped(4, seed=2)
    dicts, tuples = production_gates()
    assert sum(len(g["writes"]) * g["num_repetitions"] for g in dicts) == 415
    pis = ((0, 3), (1, 3))
    proof, setup_cap = OP.prove(c["variables"], c["sigmas"], c["constants"], tuples, 8, 2, 4, lookup=c["lookup"], public_inputs=pis)
    vk = vk_of(dict(c, gates=[]), 2, 4, setup_cap, pis)
    vk["gates"] = [t[:5] + ({k: t[5][k] for k in ("relations", "writes", "variables_offset", "constants_offset")},) for t in tuples]
    assert OV.verify(vk, proof)
    w = c["variables"].copy()
    w[154, 5] = 2                                    # breaks the boolean gate on its specialised column
    with pytest.raises((ValueError, AssertionError)):
        OP.prove(w, c["sigmas"], c["constants"], tuples, 8, 2, 4, lookup=c["lookup"])
