"""Pins the CPU oracle against the reference's own golden fixture (proof.json + vk.json, the data behind
src/gadgets/recursion/recursive_verifier.rs:2212-2476).  CPU only."""
import numpy as np

from oracle import oracle as O
from oracle import replay


def test_proof_json_replay(golden_fixture):
    """Transcript replay -> query indices -> all base/FRI Merkle paths, DEEP value, FRI fold chain and
    final-monomial evaluation reproduce the fixture (Poseidon2, sponge, Merkle layout, transcript,
    query-index derivation, DEEP combination, FRI constants and schedule are pinned bit-exactly)."""
    c = replay.replay_proof(golden_fixture)
    nq = len(golden_fixture["proof"]["queries_per_fri_repetition"])
    assert c["schedule"] == [3, 3, 3, 3, 3, 1]
    assert c["merkle_paths"] == nq * (4 + 6)
    assert c["deep"] == nq and c["final"] == nq and c["fri_levels"] == nq * 6


def test_fri_schedules():
    # security 100, L=8, cap 16: [3]*5 at 2^16 (+ remainder), 34 queries
    _, nq, sched, final = replay.compute_fri_schedule(100, 16, 0, 3, 16)
    assert nq == 34 and sum(sched) + (final.bit_length() - 1) == 16
    _, nq, sched, final = replay.compute_fri_schedule(100, 32, 0, 1, 20)
    assert nq == 100 and sched == [3, 3, 3, 3, 3, 1] and final == 16


def test_fri_fold_matches_replay_formula(golden_fixture):
    """orc_fri_fold (prover-side, table-driven) agrees with the verifier-side formula on one FRI leaf of
    the fixture: fold the 8 leaf values 3 times with roots taken from the size-nL inverse twiddle table."""
    fx = golden_fixture
    c = replay.replay_proof(fx)
    proof, fp = fx["proof"], fx["vk"]["fixed_parameters"]
    log_n = fp["domain_size"].bit_length() - 1
    log_full = log_n + 1
    # re-derive the first query index
    q = proof["queries_per_fri_repetition"][0]
    le0 = q["fri_queries"][0]["leaf_elements"]
    le1 = q["fri_queries"][1]["leaf_elements"]
    # find the tree index by replaying the transcript once more
    tr_idx = _first_query_index(fx)
    tree_idx = tr_idx >> 3
    roots = O.twiddles(log_full, inverse=True)
    c0 = np.array(le0[:8], dtype=np.uint64)
    c1 = np.array(le0[8:], dtype=np.uint64)
    coset_inv = O.inv(7)
    pos = tree_idx * 8
    for a in c["challenges"]["fri"][0]:
        r = roots[pos // 2: pos // 2 + len(c0) // 2]
        c0, c1 = O.fri_fold(c0, c1, a, r, coset_inv)
        coset_inv = O.mul(coset_inv, coset_inv)
        pos //= 2
    sub = tree_idx % 8
    assert (int(c0[0]), int(c1[0])) == (le1[sub], le1[8 + sub])


def _first_query_index(fx):
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
    bits = replay.BoolsBuffer(21).get_bits(tr, 21)
    return sum(b << i for i, b in enumerate(bits))
