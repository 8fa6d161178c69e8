"""Selector-tree helpers: turn the reference's `selectors_placement` (TreeNode, serde JSON as in vk.json) into the
selector_path / constants_placement_offset fields of bj_gate_desc, and back.

    enum TreeNode { Empty, GateOnly(GateDescription), Fork { left, right } }        src/cs/implementations/setup.rs:1392-1396
    TreeNode::output_placement: left = true, right = false, root first              setup.rs:1457-1485
    a gate's selector = prod_i (path[i] ? const_i : 1 - const_i)                    prover.rs:2775-2916 (compute_selector_subpath)
    a gate's own constants start at column len(path)                                prover.rs:1000-1013
    TreeNode::compute_stats -> (max degree, constants needed by the widest leaf)     setup.rs (degree_at_depth :1414-1422)
"""


def output_placement(tree, gate_idx):
    """TreeNode::output_placement (setup.rs:1457-1485): the path of `gate_idx`, or None."""
    if tree == "Empty" or tree is None:
        return None
    if "GateOnly" in tree:
        return [] if tree["GateOnly"]["gate_idx"] == gate_idx else None
    fork = tree["Fork"]
    for branch, bit in ((fork["left"], True), (fork["right"], False)):
        sub = output_placement(branch, gate_idx)
        if sub is not None:
            return [bit] + sub
    return None


def all_placements(tree, prefix=()):
    """{gate_idx: (path, GateDescription)} for every gate in the tree."""
    if tree == "Empty" or tree is None:
        return {}
    if "GateOnly" in tree:
        return {tree["GateOnly"]["gate_idx"]: (list(prefix), tree["GateOnly"])}
    out = {}
    out.update(all_placements(tree["Fork"]["left"], tuple(prefix) + (True,)))
    out.update(all_placements(tree["Fork"]["right"], tuple(prefix) + (False,)))
    return out


def compute_stats(tree, depth=0):
    """(max constraint degree incl. the selector product, constant columns used by general-purpose gates) as in
    TreeNode::compute_stats: a leaf at depth d has degree d + gate degree (a lookup leaf max(d, 2)) and needs d + num_constants."""
    if tree == "Empty" or tree is None:
        return 0, 0
    if "GateOnly" in tree:
        g = tree["GateOnly"]
        deg = max(depth, 2) if g.get("is_lookup") else depth + g["degree"]
        return deg, depth + g["num_constants"]
    l, r = compute_stats(tree["Fork"]["left"], depth + 1), compute_stats(tree["Fork"]["right"], depth + 1)
    return max(l[0], r[0]), max(l[1], r[1])


def gate_selector_fields(tree, gate_idx):
    """dict(selector_path=..., constants_placement_offset=...) for a gate dict / bj_gate_desc."""
    path = output_placement(tree, gate_idx)
    if path is None:
        raise KeyError("gate %d is not in the selector tree" % gate_idx)
    return {"selector_path": path, "constants_placement_offset": len(path)}


def tree_from_paths(gates):
    """Inverse for circuits described by paths (the synthetic generator): gates = [(gate_idx, path, num_constants, degree)]
    -> TreeNode JSON.  Paths must form a full binary prefix code."""
    gates = list(gates)
    if len(gates) == 1 and not gates[0][1]:
        g = gates[0]
        return {"GateOnly": {"gate_idx": g[0], "num_constants": g[2], "degree": g[3], "needs_selector": True, "is_lookup": False}}
    left = [(i, p[1:], c, d) for i, p, c, d in gates if p and p[0]]
    right = [(i, p[1:], c, d) for i, p, c, d in gates if p and not p[0]]
    if len(left) + len(right) != len(gates) or not left or not right:
        raise ValueError("selector paths are not a full binary prefix code")
    return {"Fork": {"left": tree_from_paths(left), "right": tree_from_paths(right)}}
