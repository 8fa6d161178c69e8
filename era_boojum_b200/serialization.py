"""Reference wire format for the bulky prover objects: the reference's `MemcopySerializable` (raw little-endian u64 behind
length prefixes), so that setup polynomials, LDE storages and Merkle trees built on the GPU can be cached on disk and read back
by the reference (or by this package).  Host-side plumbing only: tensors are brought to the host, no kernel is involved.

    Vec<F>                      u64 length (in base-field elements), then the elements          fast_serialization.rs:139-208
    Vec<[F; N]> (digests)       u64 flattened length, then the elements                          fast_serialization.rs:269-330
    Vec<[u8; 32]> (digests)     u64 flattened length IN BYTES, then the bytes (Blake2s / Keccak)  fast_serialization.rs:343-389
    Vec<T>                      u64 count, then every T                                           fast_serialization.rs:17-47
    GenericPolynomial           = its storage Vec<F>                                              polynomial/mod.rs:101-120
    ArcGenericLdeStorage        u64 number of cosets, then one polynomial per coset              polynomial/lde.rs:179-217
    MerkleTreeWithCap           u64 cap_size, leaf_hashes, Vec<Vec<digest>> levels               cs/oracle/merkle_tree.rs:36-73
    SetupBaseStorage (columns)  three Vec<polynomial>: copy-permutation, constants, lookup tables polynomial_storage.rs:80-122
                                (the two small bincode-encoded tails - table id indexes and the selector tree - belong to the
                                circuit description and are not produced here)
Field elements are written canonical (< p), which is what the reference holds in memory.
"""
import struct

import numpy as np

from . import MerkleTreeWithCap, to_numpy

P = 0xFFFFFFFF00000001


def _host_u64(a):
    """torch tensor (int64 bit patterns) or array-like -> contiguous little-endian u64 numpy array, canonical mod p."""
    if hasattr(a, "detach"):
        a = to_numpy(a)
    a = np.ascontiguousarray(np.asarray(a, dtype=np.uint64)).reshape(-1)
    return np.where(a >= np.uint64(P), a - np.uint64(P), a).astype("<u8")


def _read_exact(f, n):
    b = f.read(n)
    if len(b) != n:
        raise EOFError("truncated MemcopySerializable stream")
    return b


def _read_u64(f):
    return struct.unpack("<Q", _read_exact(f, 8))[0]


def write_field_vec(f, values):
    """Vec<F>: length prefix + raw elements (also a GenericPolynomial of any form)."""
    a = _host_u64(values)
    f.write(struct.pack("<Q", a.shape[0]))
    f.write(a.tobytes())


def read_field_vec(f):
    n = _read_u64(f)
    return np.frombuffer(_read_exact(f, 8 * n), dtype="<u8").copy()


def write_digest_vec(f, digests, byte_digests=False):
    """Vec<[F; 4]> (algebraic hasher: length counted in field elements) or Vec<[u8; 32]> (Blake2s / Keccak: length counted
    in bytes); the payload is the same 4 little-endian u64 per digest either way.  Digests are NOT reduced."""
    a = digests
    if hasattr(a, "detach"):
        a = to_numpy(a)
    a = np.ascontiguousarray(np.asarray(a, dtype=np.uint64)).reshape(-1).astype("<u8")
    f.write(struct.pack("<Q", a.shape[0] * (8 if byte_digests else 1)))
    f.write(a.tobytes())


def read_digest_vec(f, byte_digests=False):
    n = _read_u64(f)
    unit = 32 if byte_digests else 4
    if n % unit:
        raise ValueError("digest vector length is not a multiple of %d" % unit)
    words = n // 8 if byte_digests else n
    return np.frombuffer(_read_exact(f, 8 * words), dtype="<u8").copy().reshape(-1, 4)


def write_polynomials(f, columns):
    """Vec<Arc<GenericPolynomial>>: count, then every polynomial.  `columns`: [n_cols, n] tensor / array or list of rows."""
    cols = [columns[i] for i in range(len(columns))]
    f.write(struct.pack("<Q", len(cols)))
    for c in cols:
        write_field_vec(f, c)


def read_polynomials(f):
    return [read_field_vec(f) for _ in range(_read_u64(f))]


def write_lde_storage(f, lde_column):
    """ArcGenericLdeStorage of ONE polynomial: [L, n] (coset-major, bit-reversed inside a coset - the layout of bj_lde)."""
    cosets = [lde_column[j] for j in range(len(lde_column))]
    if len(cosets) & (len(cosets) - 1):
        raise ValueError("the number of cosets must be a power of two (lde.rs:203)")
    f.write(struct.pack("<Q", len(cosets)))
    for c in cosets:
        write_field_vec(f, c)


def read_lde_storage(f):
    n = _read_u64(f)
    if n == 0 or n & (n - 1):
        raise ValueError("the number of cosets must be a power of two (lde.rs:203)")
    return np.stack([read_field_vec(f) for _ in range(n)])


def write_merkle_tree(f, tree, hasher="poseidon2"):
    """MerkleTreeWithCap { cap_size, leaf_hashes, node_hashes_enumerated_from_leafs } (levels n/2, n/4, ..., cap_size).
    hasher: "poseidon2" (H::Output = [F; 4]) or "blake2s" / "keccak256" (H::Output = [u8; 32])."""
    byte_digests = hasher != "poseidon2"
    f.write(struct.pack("<Q", tree.cap_size))
    write_digest_vec(f, tree.leaf_hashes, byte_digests)
    levels = tree.levels()
    f.write(struct.pack("<Q", len(levels)))
    for lv in levels:
        write_digest_vec(f, lv, byte_digests)


def read_merkle_tree(f, to_tensor=None, hasher="poseidon2"):
    """-> MerkleTreeWithCap over numpy arrays (or over tensors if `to_tensor`, e.g. era_boojum_b200.to_device, is given)."""
    byte_digests = hasher != "poseidon2"
    cap_size = _read_u64(f)
    leaf_hashes = read_digest_vec(f, byte_digests)
    levels = [read_digest_vec(f, byte_digests) for _ in range(_read_u64(f))]
    n = leaf_hashes.shape[0]
    expect, cnt = [], n
    while cnt > cap_size:
        cnt //= 2
        expect.append(cnt)
    if [lv.shape[0] for lv in levels] != expect:
        raise ValueError("Merkle tree levels do not match the leaf count / cap size")
    nodes = np.concatenate(levels) if levels else np.zeros((0, 4), np.uint64)
    if to_tensor is not None:
        return MerkleTreeWithCap(cap_size, to_tensor(leaf_hashes).reshape(-1, 4), to_tensor(nodes).reshape(-1, 4))
    return _HostTree(cap_size, leaf_hashes, nodes)


class _HostTree(MerkleTreeWithCap):
    """MerkleTreeWithCap over host numpy arrays (get_cap / get_proof work without a device)."""

    def get_cap(self):
        lv = self.levels()
        return np.array(lv[-1] if lv else self.leaf_hashes, dtype=np.uint64)

    def get_proof(self, idx):
        lv = self.levels()
        layers = ([self.leaf_hashes] + lv[:-1]) if lv else []
        leaf, path = np.array(self.leaf_hashes[idx], dtype=np.uint64), []
        for layer in layers:
            path.append(np.array(layer[idx ^ 1], dtype=np.uint64))
            idx >>= 1
        return leaf, np.array(path, dtype=np.uint64).reshape(-1, 4)


def write_setup_base_columns(f, sigmas, constants, lookup_tables):
    """the polynomial part of SetupBaseStorage (polynomial_storage.rs:107-117): copy_permutation_polys, constant_columns,
    lookup_tables_columns, each a Vec of Lagrange-form polynomials in natural row order."""
    write_polynomials(f, sigmas)
    write_polynomials(f, constants)
    write_polynomials(f, lookup_tables if lookup_tables is not None else [])


def read_setup_base_columns(f):
    return read_polynomials(f), read_polynomials(f), read_polynomials(f)
