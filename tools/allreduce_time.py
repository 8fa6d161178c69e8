"""torchrun micro-check: time of the quotient-coset exchange (int64 sum all-reduce of 2 x Q x n u64) vs an all-gather."""
import os, sys, time
import torch, torch.distributed as dist
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda:%d" % local))
world = dist.get_world_size()
n = 2 * 4 * (1 << 22)
t = torch.zeros(n, dtype=torch.int64, device="cuda")
for name, fn in (("all_reduce int64 sum 256 MiB", lambda: dist.all_reduce(t)),
                 ("all_gather of the 1/world slices", lambda: dist.all_gather_into_tensor(t, t[rank * (n // world):(rank + 1) * (n // world)].clone()))):
    for _ in range(2):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    if rank == 0:
        print(name, "world", world, "ms", round((time.perf_counter() - t0) / 5 * 1e3, 3))
dist.destroy_process_group()
