"""world_size-2 gloo test (CPU) of the multi-GPU commitment sharding logic: column-sharded iNTT -> all-gather of
monomials -> coset-sharded evaluation + Merkle subtrees -> all-gather of cap digests, with the oracle standing in for
the kernels.  The assembled cap must equal the single-process commitment."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, V, log_n, L, cap, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from era_boojum_b200 import parallel
    cols = O.random_field(np.random.default_rng(7), (V, 1 << log_n))
    blk = parallel.column_block(rank, world, V)
    local = torch.from_numpy(np.ascontiguousarray(cols[blk.start:blk.stop]).view(np.int64))
    res = parallel.commit_sharded(parallel.OracleBackend(), dist, local, V, L, cap)
    want_lde = O.lde(cols, L.bit_length() - 1)
    for j, ev in res["cosets"].items():
        assert np.array_equal(ev.numpy().view(np.uint64), want_lde[:, j, :]), ("coset", j)
    _, _, want_cap = O.merkle_tree([want_lde[c].reshape(-1) for c in range(V)], cap)
    assert np.array_equal(res["cap"].numpy().view(np.uint64), want_cap)
    assert sorted(res["cosets"]) == parallel.owned_cosets(rank, world, L)
    out[rank] = 1
    dist.destroy_process_group()


@pytest.mark.parametrize("V,log_n,L,cap", [(6, 6, 4, 8), (4, 5, 8, 16)])
def test_commit_sharded_world2_gloo(V, log_n, L, cap):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), V, log_n, L, cap, out), nprocs=world, join=True)
    assert dict(out) == {0: 1, 1: 1}


def test_sharding_maps():
    from era_boojum_b200 import parallel
    assert list(parallel.column_block(1, 4, 8)) == [2, 3]
    assert parallel.owned_cosets(1, 2, 8) == [1, 3, 5, 7]
    assert parallel.leaf_owner((5 << 10) + 3, 10, 4) == 1
    # single-process path (no process group) equals the oracle too
    cols = O.random_field(np.random.default_rng(1), (3, 32))
    res = parallel.commit_sharded(parallel.OracleBackend(), None, torch.from_numpy(cols.view(np.int64)), 3, 4, 4)
    want = O.lde(cols, 2)
    assert np.array_equal(res["cap"].numpy().view(np.uint64), O.merkle_tree([want[c].reshape(-1) for c in range(3)], 4)[2])
