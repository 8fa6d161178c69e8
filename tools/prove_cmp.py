import sys, time, json; sys.path.insert(0, '.')
import torch, era_boojum_b200 as bj
from era_boojum_b200 import prover, synthetic
ctx = bj.Context.on_current_stream(0)
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 22
variables, sigmas, constants, gates, Q, lk = synthetic.generate(ctx, log_n, 60, seed=42, lookup=True)
cfg = prover.ProofConfig(fri_lde_factor=8, merkle_tree_cap_size=16, security_level=100)
nat = ctx.native_setup(sigmas, constants, gates, Q, cfg, lookup=lk)
for i in range(3):
    tm = {}; torch.cuda.synchronize(); t0 = time.perf_counter(); p = nat.prove(variables, lk["multiplicities"], timings=tm); t = time.perf_counter() - t0
    print("native", round(t, 4), {k: round(v, 4) for k, v in tm.items()})
nat.close(); del nat; torch.cuda.empty_cache()
setup = prover.Setup(ctx, sigmas, constants, gates, Q, cfg, lookup=lk)
for i in range(3):
    tm = {}; torch.cuda.synchronize(); t0 = time.perf_counter(); p2 = prover.prove(ctx, setup, variables, timings=tm, multiplicities=lk["multiplicities"]); torch.cuda.synchronize(); t = time.perf_counter() - t0
    print("python", round(t, 4), {k: round(v, 4) for k, v in tm.items()})
print("same", json.dumps(p, sort_keys=True) == json.dumps(p2, sort_keys=True))
