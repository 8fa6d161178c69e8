"""CUDA-event timing of the gate evaluator (bj_quotient_gates_general_purpose) on the gate set of the reference's own fixture
circuit (tests/golden/boojum_proof_fixture.json: 13 gate types, 130 variable columns, Poseidon2 flattened gate of ~9k relations,
415 quotient terms) - the production shape of section 8 row (c), next to the 3-gate bench circuit.  Prints one JSON line."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import era_boojum_b200 as bj
from era_boojum_b200 import gate_library as GL, placement as PL, synthetic
from oracle import verifier_reference as VR   # circuit layout of the fixture only (tooling, not the product path)

LOG_POINTS = int(os.environ.get("LOG_POINTS", "20"))
ctx = bj.Context.on_current_stream(0)      # BJ_GATE_POINTS_PER_THREAD (if set) applies to this context


def forced_ctx(k):
    os.environ["BJ_GATE_POINTS_PER_THREAD"] = str(k)
    try:
        return bj.Context.on_current_stream(0)
    finally:
        del os.environ["BJ_GATE_POINTS_PER_THREAD"]
fx = json.load(open(os.path.join(ROOT, "tests", "golden", "boojum_proof_fixture.json")))
fp = fx["vk"]["fixed_parameters"]
cfg = VR.REFERENCE_FIXTURE_GATES
lay = VR.circuit_layout(fp, cfg)
V, C = lay["num_variables"], lay["num_constants"]
tree = fp["selectors_placement"]


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


npts = 1 << LOG_POINTS
var_cols = [torch.randint(0, 2**63 - 1, (npts,), dtype=torch.int64, device="cuda:0") for _ in range(V)]
const_cols = [torch.randint(0, 2**63 - 1, (npts,), dtype=torch.int64, device="cuda:0") for _ in range(C)]
q0 = torch.zeros(npts, dtype=torch.int64, device="cuda:0")
q1 = torch.zeros(npts, dtype=torch.int64, device="cuda:0")
res = {"log_points": LOG_POINTS, "variables": V, "constants": C}
spec, gp = [], []
for s_ in lay["specialized"]:
    spec.append(GL.placed(s_["gate"], s_["reps"], [], constants_placement_offset=lay["consts_gp"] + s_["const_base"],
                          variables_initial_offset=s_["var_base"]))
for gate_idx, gate in enumerate(cfg["general_purpose"]):
    if gate.terms == 0:
        continue
    reps = gate.num_repetitions_in_geometry(lay["gp_vars"], 0, fp["parameters"]["num_constant_columns"])
    gp.append((gate.name, GL.placed(gate, reps, PL.output_placement(tree, gate_idx))))


def run(gates, ctx=ctx):
    n_terms = sum(len(g["writes"]) * g["num_repetitions"] for g in gates)
    alphas = [(3 + i, 5 + 2 * i) for i in range(n_terms)]
    import ctypes
    from era_boojum_b200.native import lib
    keep, descs = ctx._gate_descs(gates)          # ctypes conversion once, outside the timed region
    ptrs = lambda cols: (ctypes.c_void_p * max(1, len(cols)))(*[c.data_ptr() for c in cols])
    al = (ctypes.c_uint64 * (2 * n_terms))(*[int(x) for a in alphas for x in a])
    vp, wp, cp = ptrs(var_cols), ptrs([]), ptrs(const_cols)

    def call():
        st = lib.bj_quotient_gates_general_purpose(ctx._h, descs, len(gates), vp, len(var_cols), wp, 0, cp, len(const_cols), al, n_terms,
                                                   npts, q0.data_ptr(), q1.data_ptr())
        ctx._check(st)
    t = timed(call)
    n_rel = sum(len(g["relations"]) * g["num_repetitions"] for g in gates)
    return {"ms": round(t, 3), "terms": n_terms, "relations_per_point": n_rel, "gpoints_s": round(npts / t / 1e6, 3),
            "grelations_s": round(n_rel * npts / t / 1e6, 1)}


allg = spec + [g for _, g in gp]
if os.environ.get("ONLY"):                  # profiling runs: one gate, one warm-up + 3 timed launches
    res[os.environ["ONLY"]] = run([g for name, g in gp if name == os.environ["ONLY"]])
    print(json.dumps(res))
    sys.exit(0)
res["all_gates"] = run(allg)
if "BJ_GATE_POINTS_PER_THREAD" not in os.environ:
    # points per thread 1 / 2 / 4 forced: timing, and the results must be identical
    outs = {}
    for k in (1, 2, 4):
        ck = forced_ctx(k)
        q0.zero_(); q1.zero_()
        r = run(allg, ck)
        res["all_gates_k%d" % k] = r
        q0.zero_(); q1.zero_()
        ck.evaluate_gates_over_general_purpose_columns(allg, var_cols, [], const_cols, [(3 + i, 5 + 2 * i) for i in range(r["terms"])], q0, q1)
        outs[k] = (q0.clone(), q1.clone())
    res["k_variants_identical"] = all(torch.equal(outs[k][0], outs[1][0]) and torch.equal(outs[k][1], outs[1][1]) for k in (2, 4))
for name, g in gp:
    res[name] = run([g])
# the bench circuit's three gates on its own geometry (60 variable columns), same number of points
variables, sigmas, constants, gates, Q = synthetic.generate(ctx, 10, 60, seed=1)
res["bench_circuit_3_gates"] = run(gates)
print(json.dumps(res))
