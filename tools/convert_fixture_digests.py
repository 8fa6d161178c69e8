"""One-off: rewrite the Blake2s / Keccak fixtures recorded in round 1 (digests as 4 u64) to the reference's serde form for
those hashers, [u8; 32] per digest (src/cs/oracle/mod.rs:180, 245).  The digests themselves are unchanged (same 32 bytes,
little-endian words -> bytes); tools/make_proof_fixtures.py now records this form directly."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def to_bytes(d):
    if len(d) == 32:
        return d
    assert len(d) == 4
    return [b for w in d for b in int(w).to_bytes(8, "little")]


def conv(obj):
    for k in ("witness_oracle_cap", "stage_2_oracle_cap", "quotient_oracle_cap", "fri_base_oracle_cap", "setup_merkle_tree_cap"):
        if k in obj:
            obj[k] = [to_bytes(d) for d in obj[k]]
    if "fri_intermediate_oracles_caps" in obj:
        obj["fri_intermediate_oracles_caps"] = [[to_bytes(d) for d in cap] for cap in obj["fri_intermediate_oracles_caps"]]
    for q in obj.get("queries_per_fri_repetition", []):
        for k, a in q.items():
            for ans in (a if k == "fri_queries" else [a]):
                ans["proof"] = [to_bytes(d) for d in ans["proof"]]


for name in sys.argv[1:] or ["blake2s", "keccak256"]:
    path = os.path.join(ROOT, "tests", "golden", "synthetic_proof_%s.json" % name)
    with open(path) as f:
        fx = json.load(f)
    conv(fx["proof"])
    conv(fx["vk"])
    with open(path, "w") as f:
        json.dump(fx, f, separators=(",", ":"))
    print("converted", path)
