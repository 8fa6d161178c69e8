"""ORACLE (test infrastructure): a small SHA-256-bench-shaped circuit generated with numpy / Python integers - the CPU
counterpart of era_boojum_b200/synthetic.py (which needs a GPU), written independently: gates ConstantsAllocator /
FmaGateInBaseFieldWithoutConstant / ReductionGate<4> selected per row through a 3-leaf selector tree, neighbouring repetitions
chained by copy constraints, optionally lookups over specialised columns (table id in a constant column)."""
import numpy as np

from .replay import P, omega
from .stage2 import non_residues_for_copy_permutation


def sha_shaped(log_n, num_variables=20, seed=0, lookup=False, width=4, num_repetitions=2):
    """-> dict(variables [V, n], sigmas [V, n], constants [C, n], gates [(name, reps, path)], quotient_degree, lookup | None)"""
    rng = np.random.default_rng(seed)
    n, gp = 1 << log_n, num_variables
    n_ca, n_fma, n_red = min(4, gp), gp // 4, gp // 5
    gates = [("constant_allocator", n_ca, [True, True]), ("fma", n_fma, [True, False]), ("reduction4", n_red, [False])]
    kind = rng.integers(0, 3, n)                       # 0 allocator, 1 fma, 2 reduction
    var = [[int(x) for x in rng.integers(0, 1 << 20, n)] for _ in range(gp)]
    con = [[0] * n for _ in range(6)]
    pairs = []                                         # cells tied by a copy constraint: ((col a, row), (col b, row))
    for r in range(n):
        if kind[r] == 0:
            con[0][r], con[1][r] = 1, 1
            for k in range(n_ca):
                var[k][r] = con[2 + k][r] = int(rng.integers(0, 1 << 30))
        elif kind[r] == 1:
            con[0][r], con[1][r] = 1, 0
            c0, c1 = int(rng.integers(1, 1 << 10)), int(rng.integers(0, 1 << 10))
            con[2][r], con[3][r] = c0, c1
            for k in range(n_fma):
                if k:
                    var[4 * k + 2][r] = var[4 * k - 1][r]
                    pairs.append(((4 * k + 2, r), (4 * k - 1, r)))
                var[4 * k + 3][r] = (c0 * var[4 * k][r] * var[4 * k + 1][r] + c1 * var[4 * k + 2][r]) % P
        else:
            con[0][r] = 0
            rc = [int(x) for x in rng.integers(0, 1 << 8, 4)]
            con[1][r], con[2][r], con[3][r], con[4][r] = rc
            for k in range(n_red):
                if k:
                    var[5 * k][r] = var[5 * k - 1][r]
                    pairs.append(((5 * k, r), (5 * k - 1, r)))
                var[5 * k + 4][r] = sum(rc[i] * var[5 * k + i][r] for i in range(4)) % P
    lk = None
    if lookup:
        T = min(n, 1 << 10)
        tables = [[0] * n for _ in range(width + 1)]
        for i in range(T):
            row = [i, (i * i + 3) % P, i ^ 0x155, 7 * i + 1][:width]
            for j in range(width):
                tables[j][i] = row[j]
            tables[width][i] = 1
        mult = [0] * n
        for s in range(num_repetitions):
            cols = [[0] * n for _ in range(width)]
            for r in range(n):
                pick = int(rng.integers(0, T))
                mult[pick] += 1
                for j in range(width):
                    cols[j][r] = tables[j][pick]
            var += cols
        con.append([1] * n)                            # table id
        lk = dict(width=width, num_repetitions=num_repetitions, variables_offset=gp, table_id_column=6,
                  tables=np.array(tables, dtype=np.uint64), multiplicities=np.array(mult, dtype=np.uint64))
    V = len(var)
    ks = non_residues_for_copy_permutation(n, V)
    w = omega(log_n)
    xs = [pow(w, i, P) for i in range(n)]
    sig = [[ks[j] * xs[i] % P for i in range(n)] for j in range(V)]
    for (ca, ra), (cb, rb) in pairs:                   # a transposition per tied pair (the pairs are disjoint)
        sig[ca][ra], sig[cb][rb] = sig[cb][rb], sig[ca][ra]
    return dict(variables=np.array(var, dtype=np.uint64), sigmas=np.array(sig, dtype=np.uint64), constants=np.array(con, dtype=np.uint64),
                gates=gates, quotient_degree=4, lookup=lk)


# selector tree of the reference's vk.json fixture: gate index in registration order -> TreeNode::output_placement path
# (ConstantsAllocator, U8x4FMA, Poseidon2Flattened, DotProduct<4>, ZeroCheck, Fma, UIntXAdd, Selection, ParallelSelection<4>,
# Nop / PublicInput, Reduction<4>); tests/test_placement_cpu.py derives the same paths from the fixture's TreeNode
PRODUCTION_PATHS = [[False, False, False], [False, True, True, False, True, True], [True], [False, True, True, True, False, True],
                    [False, True, True, True, False, False], [False, True, True, True, True], [False, True, False],
                    [False, True, True, False, False, True], [False, True, True, False, True, False],
                    [False, True, True, False, False, False], [False, False, True]]


def production_shaped(log_n, seed=0):
    """numpy counterpart of era_boojum_b200.synthetic.generate_production_shaped (geometry of the reference's vk.json: 130
    general-purpose columns, 8 x 3 lookup columns, one boolean column; 8 constant columns; rows are Nop, ConstantsAllocator, Fma
    or Reduction rows).  -> dict(variables [155, n], sigmas, constants [8, n], lookup, quotient_degree = 8); the gate list is the
    caller's (recorded programs of era_boojum_b200/gate_library.py placed on PRODUCTION_PATHS)."""
    rng = np.random.default_rng(seed)
    n, GP, W, NSUB = 1 << log_n, 130, 3, 8
    n_fma, n_red, n_ca = GP // 4, GP // 5, 4
    var = [[int(x) for x in rng.integers(0, 1 << 20, n)] for _ in range(GP)]
    con = [[0] * n for _ in range(8)]
    pairs = []
    for r in range(n):
        kind = int(rng.integers(0, 4))               # 0 nop, 1 allocator, 2 fma, 3 reduction
        if kind == 0:
            bits = PRODUCTION_PATHS[9]
        elif kind == 1:
            bits = PRODUCTION_PATHS[0]
            for k in range(n_ca):
                var[k][r] = con[3 + k][r] = int(rng.integers(0, 1 << 30))
        elif kind == 2:
            bits = PRODUCTION_PATHS[5]
            c0, c1 = int(rng.integers(1, 1 << 10)), int(rng.integers(0, 1 << 10))
            con[5][r], con[6][r] = c0, c1            # the gate's row-shared constants start at its path length 5
            for k in range(n_fma):
                if k:
                    var[4 * k + 2][r] = var[4 * k - 1][r]
                    pairs.append(((4 * k + 2, r), (4 * k - 1, r)))
                var[4 * k + 3][r] = (c0 * var[4 * k][r] * var[4 * k + 1][r] + c1 * var[4 * k + 2][r]) % P
        else:
            bits = PRODUCTION_PATHS[10]
            rc = [int(x) for x in rng.integers(0, 1 << 8, 4)]
            con[3][r], con[4][r], con[5][r], con[6][r] = rc
            for k in range(n_red):
                if k:
                    var[5 * k][r] = var[5 * k - 1][r]
                    pairs.append(((5 * k, r), (5 * k - 1, r)))
                var[5 * k + 4][r] = sum(rc[i] * var[5 * k + i][r] for i in range(4)) % P
        for i, b in enumerate(bits):
            con[i][r] = int(b)
        con[7][r] = 1                                # lookup table id
    T = min(n, 1 << 10)
    tables = [[0] * n for _ in range(W + 1)]
    for i in range(T):
        tables[0][i], tables[1][i], tables[2][i], tables[3][i] = i, i * i + 3, i ^ 0x155, 1
    mult = [0] * n
    for s in range(NSUB):
        cols = [[0] * n for _ in range(W)]
        for r in range(n):
            pick = int(rng.integers(0, T))
            mult[pick] += 1
            for j in range(W):
                cols[j][r] = tables[j][pick]
        var += cols
    var.append([int(x) for x in rng.integers(0, 2, n)])          # the boolean gate's specialised column
    V = len(var)
    ks = non_residues_for_copy_permutation(n, V)
    w = omega(log_n)
    xs = [pow(w, i, P) for i in range(n)]
    sig = [[ks[j] * xs[i] % P for i in range(n)] for j in range(V)]
    for (ca, ra), (cb, rb) in pairs:
        sig[ca][ra], sig[cb][rb] = sig[cb][rb], sig[ca][ra]
    lk = dict(width=W, num_repetitions=NSUB, variables_offset=GP, table_id_column=7, tables=np.array(tables, dtype=np.uint64),
              multiplicities=np.array(mult, dtype=np.uint64))
    return dict(variables=np.array(var, dtype=np.uint64), sigmas=np.array(sig, dtype=np.uint64), constants=np.array(con, dtype=np.uint64),
                quotient_degree=8, lookup=lk)
