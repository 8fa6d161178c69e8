"""Multi-GPU sharding of one commitment stage: one process per GPU, `torch.distributed` (NCCL over
NVLink on the GPU box, gloo in the CPU tests) for the single bulk exchange.

The two natural partitions of the path do not coincide:
  * the iNTT that turns trace columns into monomials is independent per COLUMN (src/cs/implementations/utils.rs:295-304)
    -> rank r transforms the contiguous block of columns [r*V/G, (r+1)*V/G);
  * everything after it is per (coset, row) and needs every column of the row (leaf hashing absorbs the row in column
    order, src/cs/oracle/merkle_tree.rs:137-153) -> rank r owns the LDE cosets {j : j mod G == r}; because a leaf's index is
    coset*n + row (src/cs/implementations/proof.rs:89-91) the Merkle subtrees of different cosets are disjoint down to the
    cap, so each rank builds its subtrees locally.
The exchange between the two: ONE all-gather of the monomial coefficients (8*n*V bytes in total); afterwards only cap
digests move (32 bytes each).  The result is bit-identical to the single-GPU commitment.

Compute goes through a small backend object so the same sharding / gathering / assembly code runs on the GPU (Context
kernels, TorchBackend) and, in the CPU tests, on a test-side stand-in (tests/_oracle_backend.py; world_size 2, gloo).
"""
import numpy as np


def column_block(rank, world, n_cols):
    """contiguous block of columns owned by `rank` for the column-parallel iNTT (n_cols must divide evenly)."""
    assert n_cols % world == 0, "pad the column count to a multiple of the world size"
    per = n_cols // world
    return range(rank * per, (rank + 1) * per)


def owned_cosets(rank, world, lde_factor):
    """LDE cosets (in storage order, i.e. bit-reversed enumeration) owned by `rank`."""
    return [j for j in range(lde_factor) if j % world == rank]


def leaf_owner(leaf_index, log_n, world):
    """rank that holds leaf t = coset * n + row (and answers its query)."""
    return (leaf_index >> log_n) % world


def coset_shift(log_n, log_lde, j):
    """7 * w_{nL}^{bitrev_L(j)}: the coset evaluated by LDE slot j (src/cs/implementations/utils.rs:333-379)."""
    P = 0xFFFFFFFF00000001
    w = 0x185629DCDA58878C
    for _ in range(log_n + log_lde, 32):
        w = w * w % P
    r = int(format(j, "0%db" % log_lde)[::-1], 2) if log_lde else 0
    return 7 * pow(w, r, P) % P


class TorchBackend:
    """GPU backend: era_boojum_b200.Context kernels on torch CUDA int64 tensors."""

    def __init__(self, ctx):
        self.ctx, self.torch = ctx, ctx._torch

    def intt(self, cols):
        return self.ctx.ifft_natural_to_natural(cols.clone(), 1)

    def coset_ntt(self, monomials, shift):
        return self.ctx.fft_natural_to_bitreversed(monomials.clone(), shift)

    def subtree(self, cols_2d, cap):
        tree = self.ctx.merkle_tree_construct([cols_2d[c] for c in range(cols_2d.shape[0])], cap)
        return tree, self.torch.from_numpy(tree.get_cap().view(np.int64)).to(cols_2d.device)

    def empty(self, shape, like):
        return self.torch.empty(shape, dtype=self.torch.int64, device=like.device)


def commit_sharded(backend, dist, local_cols, n_cols_total, lde_factor, cap_size, group=None):
    """Commit to `n_cols_total` trace columns held column-sharded (local_cols = this rank's contiguous block, [V/G, n],
    Lagrange values in natural order).  Returns dict(cosets={j: [V, n] evaluations}, trees={j: subtree}, cap=[cap, 4]).
    Collectives: one all_gather of monomials, one all_gather of cap digests."""
    torch = backend.torch
    rank = dist.get_rank(group) if dist is not None else 0
    world = dist.get_world_size(group) if dist is not None else 1
    n = local_cols.shape[1]
    log_n, log_lde = n.bit_length() - 1, lde_factor.bit_length() - 1
    assert local_cols.shape[0] * world == n_cols_total
    # 1. column-parallel iNTT
    mono_local = backend.intt(local_cols)
    # 2. the exchange: all-gather monomials (rank blocks are contiguous, so the gathered tensor is in column order)
    mono = backend.empty((n_cols_total, n), mono_local)
    if world > 1:
        dist.all_gather_into_tensor(mono, mono_local.contiguous(), group=group)
    else:
        mono.copy_(mono_local)
    # 3. coset-parallel evaluation + Merkle subtrees
    per_coset_cap = max(1, cap_size // lde_factor)
    assert cap_size >= lde_factor or world <= lde_factor, "cap smaller than the LDE factor: top levels need one more gather"
    cosets, trees = {}, {}
    my_caps = []
    for j in owned_cosets(rank, world, lde_factor):
        ev = backend.coset_ntt(mono, coset_shift(log_n, log_lde, j))
        tree, capd = backend.subtree(ev, per_coset_cap)
        cosets[j], trees[j] = ev, tree
        my_caps.append(capd.reshape(per_coset_cap, 4))
    # 4. cap assembly: coset j contributes cap entries [j * cap/L, (j+1) * cap/L)
    n_mine = len(my_caps)
    mine = torch.stack(my_caps).contiguous() if n_mine else backend.empty((0, per_coset_cap, 4), mono)
    if world > 1:
        assert lde_factor % world == 0
        gathered = backend.empty((world * n_mine, per_coset_cap, 4), mono)
        dist.all_gather_into_tensor(gathered, mine, group=group)
        gathered = gathered.reshape(world, n_mine, per_coset_cap, 4)
    else:
        gathered = mine.reshape(1, n_mine, per_coset_cap, 4)
    cap = backend.empty((lde_factor, per_coset_cap, 4), mono)
    for r in range(world):
        for k, j in enumerate(owned_cosets(r, world, lde_factor)):
            cap[j] = gathered[r, k]
    cap = cap.reshape(lde_factor * per_coset_cap, 4)
    if cap_size < lde_factor:
        raise NotImplementedError("cap_size < lde_factor: hash the gathered coset roots down to the cap")
    return {"cosets": cosets, "trees": trees, "cap": cap}


# ---------------------------------------------------------------------------------------------------------------------
# Coset-sharded PROVING (era_boojum_b200.prover.prove with comm=...): every rank holds the full witness and the cosets
# j = rank (mod world) of every committed polynomial (Context.set_coset_shard).  What crosses ranks:
#   * cap digests of every oracle (32 bytes each) - all ranks replay the same transcript;
#   * the quotient values on the first Q cosets (2 * Q * n u64, one all-reduce) - the one step where the owned cosets have
#     to recombine, because the quotient is interpolated at size n*Q (src/cs/implementations/prover.rs:1399-1467);
#   * the openings at z (computed by the owner of coset 0), the last FRI codeword (a few KiB), and the query answers.
# ---------------------------------------------------------------------------------------------------------------------
class LocalComm:
    """world of one (single GPU): the collectives are identities."""
    rank, world = 0, 1

    def all_gather_host(self, obj):
        return [obj]

    def all_reduce_sum_(self, t):
        return t

    def broadcast_host(self, obj, src=0):
        return obj


class TorchDistComm:
    """one process per GPU over torch.distributed (NCCL on the GPU box)."""

    def __init__(self, dist, group=None):
        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def all_gather_host(self, obj):
        out = [None] * self.world
        self.dist.all_gather_object(out, obj, group=self.group)
        return out

    def all_reduce_sum_(self, t):
        self.dist.all_reduce(t, group=self.group)     # int64 sum; exact here because only one rank holds a non-zero value
        return t

    def broadcast_host(self, obj, src=0):
        box = [obj]
        self.dist.broadcast_object_list(box, src=src, group=self.group)
        return box[0]


class ThreadComm:
    """`world` ranks as threads of ONE process sharing one GPU - lets a single-GPU test run the sharded prover with
    real coset shards (each thread owns a Context with its own shard).  ThreadComm(world).rank_view(r) is rank r's handle."""

    def __init__(self, world):
        import threading
        self.world = world
        self._barrier = threading.Barrier(world)
        self._slots = [None] * world

    def rank_view(self, rank):
        return _ThreadCommRank(self, rank)


class _ThreadCommRank:
    def __init__(self, shared, rank):
        self._s, self.rank, self.world = shared, rank, shared.world

    def all_gather_host(self, obj):
        self._s._slots[self.rank] = obj
        self._s._barrier.wait()
        out = list(self._s._slots)
        self._s._barrier.wait()
        return out

    def all_reduce_sum_(self, t):
        parts = self.all_gather_host(t)
        total = parts[0].clone()
        for p in parts[1:]:
            total += p
        self._s._barrier.wait()      # every rank has enqueued its reads (one shared stream keeps them ordered)
        t.copy_(total)
        return t

    def broadcast_host(self, obj, src=0):
        return self.all_gather_host(obj if self.rank == src else None)[src]


def assemble_cap(comm, local_cap, lde_factor, cap_size):
    """global cap (cap_size digests) from the per-rank caps of the local trees.  A rank's tree covers its cosets
    [k][row]; with cap_size >= lde_factor every coset ends in cap_size / lde_factor cap nodes, and the cap node c of the
    global tree belongs to coset c // (cap_size / lde_factor) (leaf index = coset * n + row)."""
    per = cap_size // lde_factor
    parts = comm.all_gather_host(np.ascontiguousarray(local_cap))
    out = np.zeros((cap_size, 4), np.uint64)
    for r, part in enumerate(parts):
        part = np.asarray(part, dtype=np.uint64).reshape(-1, 4)
        for k in range(lde_factor // comm.world):
            j = k * comm.world + r
            out[j * per:(j + 1) * per] = part[k * per:(k + 1) * per]
    return out


def local_leaf_index(global_index, log_coset_len, world):
    """(owner rank, index in the owner's local [k][row] layout) of element `global_index` of a coset-major vector."""
    j, i = global_index >> log_coset_len, global_index & ((1 << log_coset_len) - 1)
    return j % world, ((j // world) << log_coset_len) | i
