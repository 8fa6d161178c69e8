"""Loader for the dump written by tools/rust_export/b200_export.rs on a host with cargo (the reference's own SHA-256 bench
circuit: witness, sigma / constant / lookup-table columns, GPUDataCapture gate programs, VerificationKey and the reference
Proof), and a writer of the same format (used to exercise the loader without Rust).  File formats: tools/rust_export/README.md.

load(dir) returns a dict with numpy arrays and the gate list in the form Context.native_setup / prover.Setup take; the tensors
are moved to the GPU by to_device().  Nothing here needs a GPU."""
import json
import os

import numpy as np

from . import native as N
from . import placement as PL

_KIND = {"variable": N.IDX_VARIABLE, "witness": N.IDX_WITNESS, "constant": N.IDX_CONSTANT_POLY, "constant_shared": N.IDX_CONSTANT_POLY_SHARED,
         "temporary": N.IDX_TEMPORARY, "value": N.IDX_CONSTANT_VALUE}
_OP = {"add": N.REL_ADD, "double": N.REL_DOUBLE, "sub": N.REL_SUB, "negate": N.REL_NEGATE, "mul": N.REL_MUL, "square": N.REL_SQUARE,
       "inverse": N.REL_INVERSE}
_KIND_NAME = {v: k for k, v in _KIND.items()}
_OP_NAME = {v: k for k, v in _OP.items()}


def _columns(path, n_cols, n):
    if n_cols == 0:
        return np.zeros((0, n), np.uint64)
    a = np.fromfile(path, dtype="<u8")
    if a.size != n_cols * n:
        raise ValueError("%s: %d values, expected %d x %d" % (path, a.size, n_cols, n))
    return a.reshape(n_cols, n)


def translate_program(gate):
    """GPUDataCapture JSON (README.md) -> dict(relations=[(op, dst, a, b)], writes=[...], offsets).  The reference numbers its
    TemporaryValues from one process-wide counter (gpu_synthesizer/mod.rs:135-165); they are renumbered from 0 per gate."""
    renum = {}

    def ix(v):
        kind, val = v
        if kind == "temporary":
            return (N.IDX_TEMPORARY, renum[val])
        return (_KIND[kind], int(val))

    rel = []
    for op, dst, a, b in gate["relations"]:
        assert dst[0] == "temporary", "a relation defines a temporary"
        a_, b_ = ix(a), (ix(b) if b is not None else None)
        renum[dst[1]] = len(renum)
        rel.append((_OP[op], renum[dst[1]], a_, b_))
    return dict(name=gate["name"], relations=rel, writes=[ix(w) for w in gate["writes"]], variables_offset=gate["variables_offset"],
                witnesses_offset=gate["witnesses_offset"], constants_offset=gate["constants_offset"])


def load(path):
    with open(os.path.join(path, "manifest.json")) as f:
        m = json.load(f)
    n = m["domain_size"]
    out = {"manifest": m, "log_n": n.bit_length() - 1}
    for name, key in (("variables", "num_variables"), ("witness", "num_witness"), ("multiplicities", "num_multiplicities"),
                      ("sigmas", "num_variables"), ("constants", "num_constants"), ("tables", "num_tables")):
        out[name] = _columns(os.path.join(path, name + ".bin"), m[key], n)
    tree = m["selectors_placement"]
    gates = []
    for g in m["gates"]:
        if g["num_quotient_terms"] == 0:
            continue
        d = translate_program(g)
        d.update(PL.gate_selector_fields(tree, g["gate_idx"]), num_repetitions=g["num_repetitions"])
        assert len(d["writes"]) == g["num_quotient_terms"]
        gates.append(d)
    out["gates"] = gates
    lk = m["lookup_parameters"]
    out["lookup"] = None
    if lk != "NoLookup":
        (kind, pars), = lk.items()
        if kind != "UseSpecializedColumnsWithTableIdAsConstant":
            raise NotImplementedError("lookup mode %s" % kind)
        gp = m["geometry"]["num_columns_under_copy_permutation"]
        out["lookup"] = dict(width=pars["width"], num_repetitions=pars["num_repetitions"], variables_offset=gp,
                             table_id_column=m["table_ids_column_idxes"][0])
    for name in ("vk", "proof"):
        p = os.path.join(path, name + ".json")
        out[name] = json.load(open(p)) if os.path.exists(p) else None
    return out


def to_device(dump, ctx):
    """numpy columns -> contiguous int64 CUDA tensors on the context's device; returns the arguments of
    ctx.native_setup(sigmas, constants, gates, Q, config, lookup=..., public_inputs=...) and of prove()."""
    import torch
    from . import prover
    dev = "cuda:%d" % ctx.device
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).to(dev).contiguous()
    m = dump["manifest"]
    if m["num_witness"]:
        raise NotImplementedError("witness columns outside the copy permutation")
    cfg = prover.ProofConfig(fri_lde_factor=m["fri_lde_factor"], merkle_tree_cap_size=m["cap_size"], security_level=m["security_level"],
                             pow_bits=m["pow_bits"], hasher=m["hasher"], transcript=m["transcript"])
    lk = None
    if dump["lookup"]:
        lk = dict(dump["lookup"], tables=t(dump["tables"]), multiplicities=t(dump["multiplicities"][0]))
    return dict(variables=t(dump["variables"]), sigmas=t(dump["sigmas"]), constants=t(dump["constants"]), gates=dump["gates"],
                quotient_degree=m["quotient_degree"], config=cfg, lookup=lk,
                public_inputs=[tuple(p) for p in m["public_inputs_locations"]])


# ------------------------------------------------------------------------------------------------- writer (same format) -----
def _ix_json(ix):
    return [_KIND_NAME[ix[0]], int(ix[1])]


def write(path, *, variables, sigmas, constants, tables, multiplicities, gates, quotient_degree, config, lookup, selectors_placement,
          public_inputs=(), proof=None, vk=None, geometry=None):
    """Write a dump in the exporter's format from this repository's own circuit description (gates: list of gate dicts with
    relations / writes / selector_path as taken by native_setup; gate_idx = position in the list)."""
    os.makedirs(path, exist_ok=True)
    n = variables.shape[1]
    for name, arr in (("variables", variables), ("sigmas", sigmas), ("constants", constants), ("tables", tables),
                      ("multiplicities", multiplicities), ("witness", np.zeros((0, n), np.uint64))):
        np.ascontiguousarray(arr, dtype="<u8").tofile(os.path.join(path, name + ".bin"))
    gl = []
    for i, g in enumerate(gates):
        gl.append({"name": g.get("name", "gate%d" % i), "gate_idx": i, "num_repetitions": g["num_repetitions"], "num_quotient_terms": len(g["writes"]),
                   "variables_offset": g.get("variables_offset", 0), "witnesses_offset": g.get("witnesses_offset", 0),
                   "constants_offset": g.get("constants_offset", 0),
                   "relations": [[_OP_NAME[op], ["temporary", int(dst)], _ix_json(a), _ix_json(b) if b is not None else None]
                                 for op, dst, a, b in g["relations"]],
                   "writes": [_ix_json(w) for w in g["writes"]]})
    lkp = "NoLookup"
    if lookup:
        lkp = {"UseSpecializedColumnsWithTableIdAsConstant": {"width": lookup["width"], "num_repetitions": lookup["num_repetitions"], "share_table_id": True}}
    m = {"domain_size": n, "num_variables": variables.shape[0], "num_witness": 0, "num_multiplicities": multiplicities.shape[0],
         "num_constants": constants.shape[0], "num_tables": tables.shape[0],
         "geometry": geometry or {"num_columns_under_copy_permutation": lookup["variables_offset"] if lookup else variables.shape[0],
                                  "num_witness_columns": 0, "num_constant_columns": 4, "max_allowed_constraint_degree": quotient_degree},
         "lookup_parameters": lkp, "quotient_degree": quotient_degree, "fri_lde_factor": config.fri_lde_factor,
         "cap_size": config.merkle_tree_cap_size, "security_level": config.security_level, "pow_bits": config.pow_bits,
         "hasher": config.hasher, "transcript": config.transcript,
         "table_ids_column_idxes": [lookup["table_id_column"]] if lookup else [], "selectors_placement": selectors_placement,
         "public_inputs_locations": [list(p) for p in public_inputs], "extra_constant_polys_for_selectors": 0, "gates": gl}
    with open(os.path.join(path, "manifest.json"), "w") as f:
        json.dump(m, f)
    for name, obj in (("proof", proof), ("vk", vk)):
        if obj is not None:
            with open(os.path.join(path, name + ".json"), "w") as f:
                json.dump(obj, f)
