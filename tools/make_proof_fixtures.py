"""Dump small proofs of the synthetic circuits (one per hasher / transcript configuration) together with their verification
keys: regression fixtures for tests/golden (the oracle verifier must keep accepting them on CPU, and the prover must keep
producing exactly these bytes on the GPU).  Run on a B200: python tools/make_proof_fixtures.py gpurun_out/"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import era_boojum_b200 as bj
from era_boojum_b200 import prover, synthetic

out_dir = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
ctx = bj.Context.on_current_stream(0)
CASES = [("poseidon2", 8, 20, False, 0, []), ("blake2s", 8, 20, True, 20, [(1, 3), (5, 3)]), ("keccak256", 7, 20, False, 0, [(0, 9)])]
for hasher, log_n, V, lookup, pow_bits, pis in CASES:
    gen = synthetic.generate(ctx, log_n, V, seed=77, lookup=lookup)
    lk = gen[5] if lookup else None
    variables, sigmas, constants, gates, Q = gen[:5]
    cfg = prover.ProofConfig(fri_lde_factor=8, merkle_tree_cap_size=16, security_level=100, pow_bits=pow_bits, hasher=hasher, transcript=hasher)
    setup = prover.Setup(ctx, sigmas, constants, gates, Q, cfg, lookup=lk, public_inputs=pis)
    m = lk["multiplicities"] if lk else None
    proof = prover.prove(ctx, setup, variables, multiplicities=m)
    nat = ctx.native_setup(sigmas, constants, gates, Q, cfg, lookup=lk, public_inputs=pis)
    assert json.dumps(nat.prove(variables, m), sort_keys=True) == json.dumps(proof, sort_keys=True)
    fixture = {"generator": {"hasher": hasher, "log_n": log_n, "num_variables": V, "lookup": lookup, "pow_bits": pow_bits,
                             "public_inputs": pis, "seed": 77}, "vk": setup.vk(), "proof": proof}
    path = os.path.join(out_dir, "synthetic_proof_%s.json" % hasher)
    with open(path, "w") as f:
        json.dump(fixture, f, separators=(",", ":"))
    print(hasher, os.path.getsize(path), "bytes")
