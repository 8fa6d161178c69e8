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
from tests._oracle_backend import OracleBackend


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
    res = parallel.commit_sharded(OracleBackend(), dist, local, V, L, cap)
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
    res = parallel.commit_sharded(OracleBackend(), None, torch.from_numpy(cols.view(np.int64)), 3, 4, 4)
    want = O.lde(cols, 2)
    assert np.array_equal(res["cap"].numpy().view(np.uint64), O.merkle_tree([want[c].reshape(-1) for c in range(3)], 4)[2])


# ---- communicators and index maps of the coset-sharded PROVER (era_boojum_b200/prover.py with comm=...) ----
def _comm_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from era_boojum_b200 import parallel
    comm = parallel.TorchDistComm(dist)
    assert (comm.rank, comm.world) == (rank, world)
    got = comm.all_gather_host({"rank": rank, "a": np.arange(3) + rank})
    assert [g["rank"] for g in got] == list(range(world)) and np.array_equal(got[1]["a"], np.arange(3) + 1)
    assert comm.broadcast_host("from0" if rank == 0 else None, 0) == "from0"
    # the quotient exchange: every rank fills only its own cosets, the integer sum recombines them exactly
    L, Q, n = 8, 4, 16
    qq = torch.zeros((2, Q, n), dtype=torch.int64)
    full = torch.from_numpy(O.random_field(np.random.default_rng(5), (2, Q, n)).view(np.int64))
    for k in range(Q // world):
        qq[:, k * world + rank] = full[:, k * world + rank]
    comm.all_reduce_sum_(qq)
    assert torch.equal(qq, full)
    # cap assembly: local trees over the owned cosets [k][row] -> the cap of the tree over all cosets [j][row]
    cap = 16
    cols = O.random_field(np.random.default_rng(9), (3, L * n)).reshape(3, L, n)
    mine = parallel.owned_cosets(rank, world, L)
    _, _, local_cap = O.merkle_tree([np.ascontiguousarray(cols[c][mine].reshape(-1)) for c in range(3)], cap // world)
    _, _, want_cap = O.merkle_tree([cols[c].reshape(-1) for c in range(3)], cap)
    assert np.array_equal(parallel.assemble_cap(comm, local_cap, L, cap), want_cap)
    out[rank] = 1
    dist.destroy_process_group()


def test_prover_communicator_world2_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_comm_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: 1, 1: 1}


def test_local_leaf_index_and_thread_comm():
    import threading
    from era_boojum_b200 import parallel
    log_n, world = 5, 4
    seen = set()
    for t in range(8 << log_n):                      # every global leaf has exactly one (owner, local index)
        owner, loc = parallel.local_leaf_index(t, log_n, world)
        assert owner == (t >> log_n) % world and (loc & 31) == (t & 31) and (loc >> log_n) == (t >> log_n) // world
        seen.add((owner, loc))
    assert len(seen) == 8 << log_n
    assert parallel.LocalComm().all_gather_host(3) == [3]
    shared, res = parallel.ThreadComm(3), [None] * 3

    def run(r):
        c = shared.rank_view(r)
        res[r] = (c.all_gather_host(r * 10), c.broadcast_host("x" if r == 2 else None, 2))

    ts = [threading.Thread(target=run, args=(r,)) for r in range(3)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert res == [([0, 10, 20], "x")] * 3
