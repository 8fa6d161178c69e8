"""Reference wire format (MemcopySerializable) of era_boojum_b200.serialization, checked on CPU against an independent
byte-level reader written from the Rust sources (fast_serialization.rs:17-47, 139-208, 269-330; merkle_tree.rs:36-73;
lde.rs:179-217; polynomial_storage.rs:80-122)."""
import io
import struct

import numpy as np
import pytest
import torch

from oracle import oracle as O

P = 0xFFFFFFFF00000001


def _u64s(buf, off, n):
    return list(struct.unpack_from("<%dQ" % n, buf, off)), off + 8 * n


def test_field_vec_and_polynomials_layout():
    from era_boojum_b200 import serialization as S
    cols = np.array([[1, 2, P + 5, 2**64 - 1], [7, 0, P - 1, P]], dtype=np.uint64)      # non-canonical inputs are reduced
    f = io.BytesIO()
    S.write_polynomials(f, torch.from_numpy(cols.view(np.int64)))
    buf = f.getvalue()
    (count,), off = _u64s(buf, 0, 1)
    assert count == 2
    for row in cols:
        (n,), off = _u64s(buf, off, 1)
        vals, off = _u64s(buf, off, n)
        assert n == 4 and vals == [int(v) % P for v in row]
    assert off == len(buf)
    back = S.read_polynomials(io.BytesIO(buf))
    assert [b.tolist() for b in back] == [[int(v) % P for v in row] for row in cols]
    with pytest.raises(EOFError):
        S.read_polynomials(io.BytesIO(buf[:-3]))


def test_merkle_tree_blob_layout_and_round_trip():
    from era_boojum_b200 import MerkleTreeWithCap, serialization as S
    cols = [O.random_field(np.random.default_rng(c), 64) for c in range(5)]
    lh, levels, cap = O.merkle_tree(cols, 4)
    nodes = np.concatenate(levels)
    tree = MerkleTreeWithCap(4, torch.from_numpy(lh.view(np.int64)), torch.from_numpy(nodes.view(np.int64)))
    f = io.BytesIO()
    S.write_merkle_tree(f, tree)
    buf = f.getvalue()
    # independent parse: cap_size | leaf_hashes (flattened length, words) | number of levels | every level
    (cap_size, flat), off = _u64s(buf, 0, 2)
    assert cap_size == 4 and flat == 64 * 4
    words, off = _u64s(buf, off, flat)
    assert words == lh.reshape(-1).tolist()
    (n_levels,), off = _u64s(buf, off, 1)
    assert n_levels == len(levels) == 4                     # 32, 16, 8, 4 nodes
    for lv in levels:
        (flat,), off = _u64s(buf, off, 1)
        words, off = _u64s(buf, off, flat)
        assert flat == lv.size and words == lv.reshape(-1).tolist()
    assert off == len(buf)
    back = S.read_merkle_tree(io.BytesIO(buf))
    assert np.array_equal(back.get_cap(), cap)
    for idx in (0, 17, 63):
        leaf, path = back.get_proof(idx)
        assert np.array_equal(path, O.merkle_path(lh, levels, idx)) and O.merkle_verify(leaf, path, cap, idx)
    # byte digests ([u8; 32] of Blake2s / Keccak trees): the same payload, lengths counted in bytes (fast_serialization.rs:343-389)
    g = io.BytesIO()
    S.write_merkle_tree(g, tree, hasher="blake2s")
    bbuf = g.getvalue()
    assert struct.unpack_from("<2Q", bbuf, 0) == (4, 64 * 32) and bbuf[16:16 + 64 * 32] == buf[16:16 + 64 * 32]
    assert np.array_equal(S.read_merkle_tree(io.BytesIO(bbuf), hasher="keccak256").get_cap(), cap)
    bad = bytearray(buf)
    bad[8:16] = struct.pack("<Q", 64 * 4 + 4)               # claims one more leaf than there is data for
    with pytest.raises((ValueError, EOFError)):
        S.read_merkle_tree(io.BytesIO(bytes(bad)))


def test_lde_storage_and_setup_columns():
    from era_boojum_b200 import serialization as S
    trace = O.random_field(np.random.default_rng(3), (2, 16))
    lde = O.lde(trace, 2)                                   # [2, 4, 16]
    f = io.BytesIO()
    S.write_lde_storage(f, lde[1])
    buf = f.getvalue()
    (n_cosets,), off = _u64s(buf, 0, 1)
    assert n_cosets == 4
    for j in range(4):
        (n,), off = _u64s(buf, off, 1)
        vals, off = _u64s(buf, off, n)
        assert vals == lde[1, j].tolist()
    assert np.array_equal(S.read_lde_storage(io.BytesIO(buf)), lde[1])
    with pytest.raises(ValueError):
        S.write_lde_storage(io.BytesIO(), lde[1][:3])
    f = io.BytesIO()
    S.write_setup_base_columns(f, trace, trace[:1], None)
    sig, con, tab = S.read_setup_base_columns(io.BytesIO(f.getvalue()))
    assert len(sig) == 2 and len(con) == 1 and tab == [] and np.array_equal(sig[1], trace[1])
