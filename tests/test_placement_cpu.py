"""Selector-tree helpers against the reference's own vk.json (the `selectors_placement` of the recursive-verifier fixture)."""
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _tree():
    with open(os.path.join(HERE, "golden", "boojum_proof_fixture.json")) as f:
        return json.load(f)["vk"]["fixed_parameters"]


def test_paths_of_the_reference_vk_tree():
    from era_boojum_b200 import placement as PL
    fp = _tree()
    tree = fp["selectors_placement"]
    placed = PL.all_placements(tree)
    assert len(placed) >= 9 and all(PL.output_placement(tree, g) == path for g, (path, _) in placed.items())
    # left = true, right = false (setup.rs:1467-1479): the root's left child is gate 2 in this vk
    assert PL.output_placement(tree, 2) == [True]
    assert PL.output_placement(tree, 12345) is None
    paths = [tuple(p) for p, _ in placed.values()]
    assert len(set(paths)) == len(paths)
    for a in paths:                                              # prefix-free: exactly one selector is 1 on every row
        assert not any(a != b and b[:len(a)] == a for b in paths)
    assert sum(2.0 ** -len(p) for p in paths) == 1.0             # and the code is complete (Kraft equality)
    # stats the verification key itself records: quotient degree 8 bounds the selector-times-gate degree, and
    # `extra_constant_polys_for_selectors` + `num_constant_columns` bound the constants of the widest leaf
    degree, constants = PL.compute_stats(tree)
    assert degree <= fp["quotient_degree"]
    assert constants <= fp["parameters"]["num_constant_columns"] + fp["extra_constant_polys_for_selectors"]
    f = PL.gate_selector_fields(tree, 5)
    assert f["constants_placement_offset"] == len(f["selector_path"]) == len(PL.output_placement(tree, 5))
    with pytest.raises(KeyError):
        PL.gate_selector_fields(tree, 999)


def test_tree_from_paths_round_trip_for_the_synthetic_circuit():
    from era_boojum_b200 import placement as PL
    gates = [(0, [True, True], 4, 1), (1, [True, False], 2, 3), (2, [False], 4, 2)]     # allocator, fma, reduction
    tree = PL.tree_from_paths(gates)
    assert [PL.output_placement(tree, g) for g, *_ in gates] == [p for _, p, _, _ in gates]
    assert PL.compute_stats(tree) == (max(2 + 1, 2 + 3, 1 + 2), max(2 + 4, 2 + 2, 1 + 4))
    with pytest.raises(ValueError):
        PL.tree_from_paths([(0, [True], 1, 1), (1, [True, False], 1, 1)])
