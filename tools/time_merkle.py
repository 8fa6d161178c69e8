"""CUDA-event timing of the Poseidon2 Merkle build of BASELINE config 3 (2^22 leaves x 100 columns, cap 16)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, era_boojum_b200 as bj
ctx = bj.Context.on_current_stream(0)
gen = torch.Generator(device="cuda:0"); gen.manual_seed(1)
srcs = [torch.randint(0, 2**63 - 1, (1 << 22,), dtype=torch.int64, device="cuda:0", generator=gen) for _ in range(100)]
tree = ctx.merkle_tree_construct(srcs, 16)
cap = tree.get_cap()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    ctx.merkle_tree_construct(srcs, 16)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
print(json.dumps({"variant": os.environ.get("BJ_LIB_VARIANT", "default"), "ms": round(ms, 3), "gperms_s": round(((1 << 22) * 14 - 16) / ms / 1e6, 4),
                  "cap0": [int(x) for x in cap[0]]}))
