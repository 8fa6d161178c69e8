"""Short driver for ncu captures: a few forward NTTs (2^22 x 32 columns, coset 7) + one Merkle build."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import era_boojum_b200 as bj

what = sys.argv[1] if len(sys.argv) > 1 else "ntt"
ctx = bj.Context.on_current_stream(0)
if what == "ntt":
    log_n, cols = 22, 32
    d = torch.randint(0, 2**63 - 1, (cols, 1 << log_n), dtype=torch.int64, device="cuda:0")
    for _ in range(4):
        ctx.fft_natural_to_bitreversed(d, 7)
elif what == "merkle":
    n, cols = 1 << 18, 100
    srcs = [torch.randint(0, 2**63 - 1, (n,), dtype=torch.int64, device="cuda:0") for _ in range(cols)]
    for _ in range(2):
        ctx.merkle_tree_construct(srcs, 16)
ctx.synchronize()
torch.cuda.synchronize()
