"""bj_prove on a circuit loaded from a reference dump (tools/rust_export/README.md).

  * with B200_REFERENCE_DUMP=<dir written by tools/rust_export/b200_export.rs on a cargo host>: the setup cap must equal the
    reference's vk.json and the proof must equal the reference's proof.json value for value - the byte-level check of
    BASELINE.json configs[3] / configs[4] (skipped when no dump is present: the reference cannot be built in this image);
  * always (GPU): the same path on a dump of the synthetic SHA-shaped circuit written in the exporter's format."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _prove_from_dump(path):
    import era_boojum_b200 as bj
    from era_boojum_b200 import reference_dump as RD
    ctx = bj.Context.on_current_stream(0)
    dump = RD.load(path)
    a = RD.to_device(dump, ctx)
    nat = ctx.native_setup(a["sigmas"], a["constants"], a["gates"], a["quotient_degree"], a["config"], lookup=a["lookup"],
                           public_inputs=a["public_inputs"])
    proof = nat.prove(a["variables"], a["lookup"]["multiplicities"] if a["lookup"] else None)
    cap = nat.get_cap()
    vk = nat.vk()
    nat.close()
    ctx.close()
    return dump, proof, cap, vk


@pytest.mark.skipif(not os.environ.get("B200_REFERENCE_DUMP"), reason="no reference dump (needs a host with cargo, tools/rust_export)")
def test_proof_equals_the_reference_proof_value_for_value():
    from era_boojum_b200 import prover
    dump, proof, cap, _ = _prove_from_dump(os.environ["B200_REFERENCE_DUMP"])
    ref_vk, ref_proof = dump["vk"], dump["proof"]
    assert ref_vk is not None and ref_proof is not None, "the dump must hold vk.json and proof.json"
    assert prover._digests(cap, dump["manifest"]["hasher"]) == ref_vk["setup_merkle_tree_cap"], "setup cap differs from the reference VK"
    for key in ref_proof:
        assert proof[key] == ref_proof[key], "proof field %s differs from the reference proof" % key
    assert json.dumps(proof, sort_keys=True) == json.dumps(ref_proof, sort_keys=True)


def test_dump_format_round_trip_through_the_native_prover(tmp_path):
    import era_boojum_b200 as bj
    from era_boojum_b200 import placement as PL
    from era_boojum_b200 import prover, synthetic, reference_dump as RD
    from oracle import verifier as OV
    ctx = bj.Context.on_current_stream(0)
    variables, sigmas, constants, gates, Q, lk = synthetic.generate(ctx, 9, 60, seed=12, lookup=True)
    cfg = prover.ProofConfig(fri_lde_factor=8, merkle_tree_cap_size=16, security_level=100, hasher="blake2s", transcript="blake2s")
    nat = ctx.native_setup(sigmas, constants, gates, Q, cfg, lookup=lk)
    want = nat.prove(variables, lk["multiplicities"])
    want_cap, vk = nat.get_cap(), nat.vk()
    nat.close()
    tree = PL.tree_from_paths([(i, g["selector_path"], 4, 2) for i, g in enumerate(gates)])
    RD.write(str(tmp_path), variables=bj.to_numpy(variables), sigmas=bj.to_numpy(sigmas), constants=bj.to_numpy(constants),
             tables=bj.to_numpy(lk["tables"]), multiplicities=bj.to_numpy(lk["multiplicities"]).reshape(1, -1), gates=gates,
             quotient_degree=Q, config=cfg, lookup=lk, selectors_placement=tree, proof=want)
    ctx.close()
    dump, proof, cap, vk2 = _prove_from_dump(str(tmp_path))
    assert np.array_equal(cap, want_cap)
    assert json.dumps(proof, sort_keys=True) == json.dumps(dump["proof"], sort_keys=True) == json.dumps(want, sort_keys=True)
    assert OV.verify(vk, proof)
