"""Extract the golden fixture used to pin the oracle from the reference's own proof.json / vk.json.

Dev-time only (needs /root/reference).  The reference's test `test_recursive_verification`
(src/gadgets/recursion/recursive_verifier.rs:2212-2476) deserialises these two files and verifies
them natively, so they are the only numeric golden vectors the reference holds for this path.
We keep everything the transcript replay needs (caps, openings, FRI monomials, config, the VK) but
only the first N_QUERIES of the 100 query records to keep the fixture small.

    python tools/make_golden.py            # writes tests/golden/boojum_proof_fixture.json
"""
import json
import os

N_QUERIES = 6
REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "tests", "golden", "boojum_proof_fixture.json")


def main():
    proof = json.load(open(os.path.join(REF, "proof.json")))
    vk = json.load(open(os.path.join(REF, "vk.json")))
    total = len(proof["queries_per_fri_repetition"])
    proof["queries_per_fri_repetition"] = proof["queries_per_fri_repetition"][:N_QUERIES]
    fixture = {
        "source": "matter-labs/era-boojum proof.json + vk.json (v0.2.2), first %d of %d queries" % (N_QUERIES, total),
        "num_queries_total": total,
        "vk": vk,
        "proof": proof,
    }
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        json.dump(fixture, f, separators=(",", ":"))
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
