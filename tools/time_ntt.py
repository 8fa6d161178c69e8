"""CUDA-event timing of the NTT family per size (1 GiB batches): forward coset-7, plain forward, inverse, LDE x8.
Env toggles of the library (BJ_NTT_*) apply; used for A/B runs of kernel variants."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import era_boojum_b200 as bj

ctx = bj.Context.on_current_stream(0)
res = {}


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for m in (20, 21, 22, 23, 24):
    cols = 1 << (27 - m)
    d = torch.randint(0, 2**63 - 1, (cols, 1 << m), dtype=torch.int64, device="cuda:0")
    el = cols << m
    r = {}
    r["fwd_coset7"] = el / timed(lambda: ctx.fft_natural_to_bitreversed(d, 7)) / 1e6
    r["fwd_plain"] = el / timed(lambda: ctx.fft_natural_to_bitreversed(d, 1)) / 1e6
    r["inv_coset7"] = el / timed(lambda: ctx.ifft_natural_to_natural(d, 7)) / 1e6
    res["2^%d" % m] = {k: round(v, 2) for k, v in r.items()}
    del d
# the same transform on a batch that fits in L2 (2 columns of 2^22 = 64 MiB < 126 MB): if the kernel waited for HBM this would
# be much faster than the 1 GiB batch above - it is not (the passes are bound by the integer pipes)
m = 22
d = torch.randint(0, 2**63 - 1, (2, 1 << m), dtype=torch.int64, device="cuda:0")
res["2^22_l2_resident_2cols"] = {"fwd_plain": round((2 << m) / timed(lambda: ctx.fft_natural_to_bitreversed(d, 1), reps=50) / 1e6, 2),
                                 "fwd_coset7": round((2 << m) / timed(lambda: ctx.fft_natural_to_bitreversed(d, 7), reps=50) / 1e6, 2)}
del d
for m in (20, 22):
    cols = 1 << (24 - m)
    d = torch.randint(0, 2**63 - 1, (cols, 1 << m), dtype=torch.int64, device="cuda:0")
    out = torch.empty((cols, 8, 1 << m), dtype=torch.int64, device="cuda:0")
    t = timed(lambda: ctx.transform_raw_storages_to_lde(d, 8, out=out))
    res["lde8_2^%d" % m] = {"ms": round(t, 3), "out_gelem_s": round(8 * (cols << m) / t / 1e6, 2)}
    del d, out
print(json.dumps(res))
