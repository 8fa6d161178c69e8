"""One warm-up + one proof of the synthetic SHA-shaped circuit through bj_prove (for ncu launch lists).
usage: prove_once.py [log_n] [poseidon2|blake2s]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, era_boojum_b200 as bj
from era_boojum_b200 import prover, synthetic
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
hasher = sys.argv[2] if len(sys.argv) > 2 else "poseidon2"
ctx = bj.Context.on_current_stream(0)
variables, sigmas, constants, gates, Q, lk = synthetic.generate(ctx, log_n, 60, seed=42, lookup=True)
transcript = "poseidon" if hasher == "poseidon2" else hasher   # the reference benches' (H, TR) pairs (sha256/mod.rs:264-293)
cfg = prover.ProofConfig(fri_lde_factor=8, merkle_tree_cap_size=16, security_level=100, hasher=hasher, transcript=transcript)
nat = ctx.native_setup(sigmas, constants, gates, Q, cfg, lookup=lk)
if os.environ.get("WARM", "1") == "1":
    nat.prove(variables, lk["multiplicities"])
torch.cuda.synchronize()
tm = {}
t0 = time.perf_counter()
nat.prove(variables, lk["multiplicities"], timings=tm)
print(hasher, log_n, round(time.perf_counter() - t0, 4), {k: round(v, 4) for k, v in tm.items()})
