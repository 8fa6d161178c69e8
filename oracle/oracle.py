"""ORACLE (test infrastructure, NOT product code): ctypes/numpy front-end of oracle/liboracle.so.

Each wrapper names the reference function it restates (paths relative to /root/reference/src).
All arrays are numpy uint64, canonical Goldilocks values on output.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

P = 0xFFFFFFFF00000001
MULT_GEN = 7


def build(force=False):
    """Compile oracle/liboracle.so with the system gcc (oracle/Makefile)."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        u64, sz, vp = ctypes.c_uint64, ctypes.c_size_t, ctypes.c_void_p
        sig = {
            "orc_add": (u64, [u64, u64]), "orc_sub": (u64, [u64, u64]), "orc_mul": (u64, [u64, u64]),
            "orc_mul_slow": (u64, [u64, u64]), "orc_inv": (u64, [u64]), "orc_pow": (u64, [u64, u64]),
            "orc_omega": (u64, [ctypes.c_uint]),
            "orc_ext_mul": (None, [vp, vp, vp]), "orc_ext_inv": (None, [vp, vp]),
            "orc_num_threads": (ctypes.c_int, []), "orc_set_threads": (None, [ctypes.c_int]),
            "orc_bitreverse": (None, [vp, ctypes.c_uint]),
            "orc_twiddles": (None, [vp, ctypes.c_uint, ctypes.c_int, ctypes.c_int]),
            "orc_ntt_n2b": (None, [vp, ctypes.c_uint, sz, sz, u64]),
            "orc_intt_n2n": (None, [vp, ctypes.c_uint, sz, sz, u64]),
            "orc_lde": (None, [vp, vp, ctypes.c_uint, ctypes.c_uint, sz, ctypes.c_int]),
            "orc_naive_dft_bitreversed": (None, [vp, vp, ctypes.c_uint, u64]),
            "orc_poseidon2_permutation": (None, [vp]),
            "orc_poseidon2_hash_leaf": (None, [vp, sz, vp]),
            "orc_poseidon2_hash_node": (None, [vp, vp, vp]),
            "orc_merkle_leaf_hashes": (None, [vp, sz, sz, sz, vp]),
            "orc_merkle_nodes": (sz, [vp, sz, sz, vp]),
            "orc_merkle_verify": (ctypes.c_int, [vp, vp, sz, vp, sz, sz]),
            "orc_fri_fold": (None, [vp, vp, sz, vp, vp, u64, vp, vp]),
            "orc_batch_inverse": (None, [vp, sz]),
            "orc_batch_inverse_ext": (None, [vp, vp, sz]),
            "orc_deep_point": (None, [vp, vp, vp, vp, vp, vp, vp, sz, u64, vp]),
            "orc_deep_group": (None, [vp, vp, vp, vp, sz, vp, vp, vp, ctypes.c_uint]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


# ---------------------------------------------------------------- field -----
def add(a, b):
    return lib().orc_add(a, b)


def sub(a, b):
    return lib().orc_sub(a, b)


def mul(a, b):
    return lib().orc_mul(a, b)


def inv(a):
    return lib().orc_inv(a)


def pow_(a, e):
    return lib().orc_pow(a, e)


def omega(log_n):
    """domain_generator_for_size (cs/implementations/utils.rs:13-28)."""
    return lib().orc_omega(log_n)


def ext_mul(a, b):
    x, y, o = _u64(a), _u64(b), np.zeros(2, np.uint64)
    lib().orc_ext_mul(_p(x), _p(y), _p(o))
    return (int(o[0]), int(o[1]))


def ext_inv(a):
    x, o = _u64(a), np.zeros(2, np.uint64)
    lib().orc_ext_inv(_p(x), _p(o))
    return (int(o[0]), int(o[1]))


def set_threads(n):
    lib().orc_set_threads(n)


def num_threads():
    return lib().orc_num_threads()


# ------------------------------------------------------------------ NTT -----
def bitreverse(a):
    a = _u64(a).copy()
    log_n = int(a.shape[-1]).bit_length() - 1
    flat = a.reshape(-1, a.shape[-1])
    for row in flat:
        lib().orc_bitreverse(_p(row), log_n)
    return a


def twiddles(log_n, inverse=False, with_assert=False):
    """precompute_twiddles_for_fft (cs/implementations/utils.rs:88-125)."""
    t = np.zeros(max(1, (1 << log_n) // 2), np.uint64)
    lib().orc_twiddles(_p(t), log_n, int(inverse), int(with_assert))
    return t


def ntt_n2b(cols, coset=1):
    """fft_natural_to_bitreversed per column (fft/mod.rs:398-411); cols: [n_cols, n] or [n]."""
    a = _u64(cols).copy()
    n = a.shape[-1]
    ncols = a.size // n
    lib().orc_ntt_n2b(_p(a), n.bit_length() - 1, ncols, n, coset)
    return a


def intt_n2n(cols, coset=1):
    """ifft_natural_to_natural per column (fft/mod.rs:464-491)."""
    a = _u64(cols).copy()
    n = a.shape[-1]
    ncols = a.size // n
    lib().orc_intt_n2n(_p(a), n.bit_length() - 1, ncols, n, coset)
    return a


def lde(cols, log_lde, from_monomials=False):
    """transform_raw_storages_to_lde (cs/implementations/utils.rs:270-403) -> [n_cols, L, n]."""
    a = _u64(cols)
    if a.ndim == 1:
        a = a[None, :]
    ncols, n = a.shape
    out = np.zeros((ncols, 1 << log_lde, n), np.uint64)
    lib().orc_lde(_p(a), _p(out), n.bit_length() - 1, log_lde, ncols, int(from_monomials))
    return out


def naive_dft_bitreversed(a, coset=1):
    a = _u64(a)
    out = np.zeros_like(a)
    lib().orc_naive_dft_bitreversed(_p(a), _p(out), a.shape[0].bit_length() - 1, coset)
    return out


# ------------------------------------------------------------ Poseidon2 -----
def poseidon2_permutation(state):
    s = _u64(state).copy()
    assert s.shape == (12,)
    lib().orc_poseidon2_permutation(_p(s))
    return s


def poseidon2_hash_leaf(els):
    e, o = _u64(els), np.zeros(4, np.uint64)
    lib().orc_poseidon2_hash_leaf(_p(e), e.shape[0], _p(o))
    return o


def poseidon2_hash_node(l, r):
    a, b, o = _u64(l), _u64(r), np.zeros(4, np.uint64)
    lib().orc_poseidon2_hash_node(_p(a), _p(b), _p(o))
    return o


def merkle_leaf_hashes(sources, elems_per_leaf=1):
    """sources: list of flat uint64 arrays (each n_leaves*elems_per_leaf long), in leaf-preimage order."""
    srcs = [_u64(s).reshape(-1) for s in sources]
    n_leaves = srcs[0].shape[0] // elems_per_leaf
    ptrs = (ctypes.c_void_p * len(srcs))(*[s.ctypes.data for s in srcs])
    out = np.zeros((n_leaves, 4), np.uint64)
    lib().orc_merkle_leaf_hashes(ptrs, len(srcs), n_leaves, elems_per_leaf, _p(out))
    return out


def merkle_nodes(leaf_hashes, cap_size):
    """continue_from_leaf_hashes (cs/oracle/merkle_tree.rs:388-449): list of levels (bottom-up)."""
    lh = _u64(leaf_hashes)
    n = lh.shape[0]
    total = max(0, n - cap_size)
    nodes = np.zeros((max(total, 1), 4), np.uint64)
    w = lib().orc_merkle_nodes(_p(lh), n, cap_size, _p(nodes))
    assert w == total
    levels, off, cnt = [], 0, n
    while cnt > cap_size:
        cnt //= 2
        levels.append(nodes[off:off + cnt])
        off += cnt
    return levels


def merkle_tree(sources, cap_size, elems_per_leaf=1):
    lh = merkle_leaf_hashes(sources, elems_per_leaf)
    levels = merkle_nodes(lh, cap_size)
    cap = levels[-1] if levels else lh
    return lh, levels, cap


def merkle_path(lh, levels, idx):
    """get_proof (cs/oracle/merkle_tree.rs:462-480): siblings bottom-up, cap level excluded."""
    path = []
    layers = [lh] + list(levels[:-1]) if levels else []
    for layer in layers:
        path.append(layer[idx ^ 1].copy())
        idx >>= 1
    return np.array(path, dtype=np.uint64).reshape(-1, 4)


def merkle_verify(leaf_hash, path, cap, idx):
    lh, pa, ca = _u64(leaf_hash), _u64(path).reshape(-1, 4), _u64(cap).reshape(-1, 4)
    return bool(lib().orc_merkle_verify(_p(lh), _p(pa), pa.shape[0], _p(ca), ca.shape[0], idx))


# ------------------------------------------------------------------ FRI -----
def fri_fold(c0, c1, alpha, roots, coset_inv):
    """fold_multiple (cs/implementations/fri/mod.rs:362-474), one fold-by-2."""
    a, b, r = _u64(c0), _u64(c1), _u64(roots)
    al = _u64(alpha)
    m = a.shape[0]
    o0, o1 = np.zeros(m // 2, np.uint64), np.zeros(m // 2, np.uint64)
    lib().orc_fri_fold(_p(a), _p(b), m, _p(al), _p(r), coset_inv, _p(o0), _p(o1))
    return o0, o1


def batch_inverse(a):
    x = _u64(a).copy()
    lib().orc_batch_inverse(_p(x), x.shape[0])
    return x


def batch_inverse_ext(c0, c1):
    x, y = _u64(c0).copy(), _u64(c1).copy()
    lib().orc_batch_inverse_ext(_p(x), _p(y), x.shape[0])
    return x, y


def deep_point(acc, f, v, ch, x, at):
    """One quotening_operation group at one point (verifier.rs:2526-2565).  f, v, ch: lists of (c0,c1)."""
    f, v, ch = _u64(f).reshape(-1, 2), _u64(v).reshape(-1, 2), _u64(ch).reshape(-1, 2)
    cols = [np.ascontiguousarray(t[:, k]) for t in (f, v, ch) for k in (0, 1)]
    a, at_ = _u64(acc).copy(), _u64(at)
    lib().orc_deep_point(_p(a), *[_p(c) for c in cols], f.shape[0], x, _p(at_))
    return (int(a[0]), int(a[1]))


def deep_group(acc0, acc1, sources, values_at, challenges, at):
    """quotening_operation_in_extension over the whole LDE domain (prover.rs:2523-2706).
    sources: list of (c0_array, c1_array_or_None); returns updated (acc0, acc1)."""
    a0, a1 = _u64(acc0).copy(), _u64(acc1).copy()
    c0s = [_u64(s[0]).reshape(-1) for s in sources]
    c1s = [(_u64(s[1]).reshape(-1) if s[1] is not None else None) for s in sources]
    n = len(sources)
    p0 = (ctypes.c_void_p * n)(*[c.ctypes.data for c in c0s])
    p1 = (ctypes.c_void_p * n)(*[(c.ctypes.data if c is not None else None) for c in c1s])
    vals = _u64(np.array(values_at, dtype=np.uint64).reshape(-1))
    chs = _u64(np.array(challenges, dtype=np.uint64).reshape(-1))
    at_ = _u64(np.array(at, dtype=np.uint64))
    lib().orc_deep_group(_p(a0), _p(a1), p0, p1, n, _p(vals), _p(chs), _p(at_), a0.shape[0].bit_length() - 1)
    return a0, a1


def random_field(rng, shape):
    """Uniform canonical Goldilocks elements from a numpy Generator (rejection below p)."""
    a = rng.integers(0, 2**64, size=shape, dtype=np.uint64)
    bad = a >= np.uint64(P)
    while bad.any():
        a[bad] = rng.integers(0, 2**64, size=int(bad.sum()), dtype=np.uint64)
        bad = a >= np.uint64(P)
    return a
