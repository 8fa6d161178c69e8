"""The reference-dump format (tools/rust_export/README.md) round trips through the loader: programs recorded from
gate_library, written by reference_dump.write and read back by reference_dump.load are the same programs (temporaries
renumbered), columns are bit-identical, selector paths come from the TreeNode JSON.  CPU only."""
import numpy as np


def test_dump_round_trip(tmp_path):
    from era_boojum_b200 import gate_library as GL
    from era_boojum_b200 import placement as PL
    from era_boojum_b200 import prover, reference_dump as RD
    r = np.random.default_rng(1)
    n, V, C = 64, 60 + 32, 7
    paths = [[True, True], [True, False], [False]]
    gates = [GL.placed(GL.CONSTANT_ALLOCATOR, 4, paths[0]), GL.placed(GL.FMA, 15, paths[1]), GL.placed(GL.REDUCTION4, 12, paths[2])]
    # the exporter numbers temporaries from a process-wide counter: shift them to make sure the loader renumbers
    shifted = []
    for g in gates:
        g2 = dict(g)
        off = 1000 + 7 * len(shifted)
        bump = lambda ix: (ix[0], ix[1] + off) if ix is not None and ix[0] == 3 else ix
        g2["relations"] = [(op, dst + off, bump(a), bump(b)) for op, dst, a, b in g["relations"]]
        g2["writes"] = [bump(w) for w in g["writes"]]
        shifted.append(g2)
    tree = PL.tree_from_paths([(i, p, 4, 2) for i, p in enumerate(paths)])
    cols = {k: r.integers(0, 2**63, size=s, dtype=np.uint64) for k, s in
            (("variables", (V, n)), ("sigmas", (V, n)), ("constants", (C, n)), ("tables", (5, n)), ("multiplicities", (1, n)))}
    cfg = prover.ProofConfig(fri_lde_factor=8, merkle_tree_cap_size=16, security_level=100, hasher="blake2s", transcript="blake2s")
    lk = dict(width=4, num_repetitions=8, variables_offset=60, table_id_column=6)
    RD.write(str(tmp_path), gates=shifted, quotient_degree=4, config=cfg, lookup=lk, selectors_placement=tree, **cols)
    d = RD.load(str(tmp_path))
    for k, v in cols.items():
        assert np.array_equal(d[k], v)
    assert d["log_n"] == 6 and d["lookup"] == lk and d["proof"] is None
    assert len(d["gates"]) == 3
    for got, want in zip(d["gates"], gates):
        assert got["relations"] == want["relations"] and got["writes"] == want["writes"]
        assert got["selector_path"] == want["selector_path"] and got["constants_placement_offset"] == len(want["selector_path"])
        assert (got["num_repetitions"], got["variables_offset"], got["constants_offset"]) == (
            want["num_repetitions"], want["variables_offset"], want["constants_offset"])
