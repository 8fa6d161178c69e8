"""The host compiler of recorded gate programs (csrc/gates.cu: validation, push placement, peephole, slot allocation, lowering
to the 32-byte step format) checked WITHOUT a GPU: bj_gate_programs_compile returns the steps the device interpreter would
run, an emulator of the documented step format (include/boojum_b200.h) executes them over Python integers, and the pushed
terms must equal the gate evaluators' own terms (era_boojum_b200/gate_library.py, pinned by the quotient identity on the
reference's proof.json) - every gate type, several repetitions, every peephole setting."""
import random

import numpy as np
import pytest

import era_boojum_b200 as bj
from era_boojum_b200 import gate_library as GL
from oracle import verifier_reference as VR

P = VR.P
GATES = [GL.CONSTANT_ALLOCATOR, GL.BOOLEAN, GL.FMA, GL.REDUCTION4, GL.DOT_PRODUCT4, GL.ZERO_CHECK, GL.UINTX_ADD, GL.SELECTION,
         GL.PARALLEL_SELECTION4, GL.U8X4_FMA, GL.poseidon2_flattened_gate(130, 0)]
T_, L_, I_ = 0, 1, 2


def emulate(records, first, last, cols, reps, n_terms):
    """runs records[first:last] for `reps` repetitions; cols = the unified column table of one point; returns the pushed terms
    in term-index order per repetition"""
    out = []
    for rep in range(reps):
        slots, terms = {}, {}
        i = first
        while i < last:
            w0, a_raw, b_raw, c_raw = (int(x) for x in records[i])
            code, push, dst = w0 & 0x7F, (w0 >> 7) & 1, (w0 >> 8) & 0xFFFFFF
            sa, sb = (w0 >> 32) & 0xFFFF, (w0 >> 48) & 0xFFFF

            def operand(kind, raw, stride):
                if kind == T_:
                    return slots[raw]
                if kind == L_:
                    return cols[raw + rep * stride] % P
                return raw % P
            if code == 48:
                n, acc = a_raw, b_raw % P
                stride = (w0 >> 32) & 0xFFFFFFFF
                for t in range(n):
                    word = int(records[i + 1 + t // 4][(t % 4)])
                    ref, k = word & 0xFFFFFFFF, word >> 32
                    assert k < (1 << 28)
                    x = cols[(ref & 0x7FFFFFFF) + rep * stride] % P if ref >> 31 else slots[ref]
                    acc = (acc + k * x) % P
                r = acc
                i += (n + 3) // 4
            elif code < 27:
                op, ka, kb = code // 9, (code % 9) // 3, code % 3
                a, b = operand(ka, a_raw, sa), operand(kb, b_raw, sb)
                r = [(a + b) % P, (a - b) % P, a * b % P][op]
            elif code < 42:
                op, ka = (code - 27) // 3, (code - 27) % 3
                a = operand(ka, a_raw, sa)
                r = [2 * a % P, (-a) % P, a * a % P, pow(a, P - 2, P), a][op]
            else:
                ka, kb = (code - 42) // 3, (code - 42) % 3
                assert 42 <= code < 48 and ka in (T_, L_)
                r = (operand(ka, a_raw, sa) * operand(kb, b_raw, sb) + slots[c_raw]) % P
            if push:
                assert dst not in terms
                terms[dst] = r
            else:
                slots[dst] = r
            i += 1
        assert sorted(terms) == list(range(n_terms))
        out += [terms[k] for k in range(n_terms)]
    return out


@pytest.mark.parametrize("peephole", [0, 1, 3, 7, 15, 13, 8])
@pytest.mark.parametrize("gate", GATES, ids=lambda g: g.name)
def test_compiled_steps_compute_the_gate_terms(gate, peephole):
    rnd = random.Random(hash(gate.name) & 0xFFFF)
    reps = min(3, gate.num_repetitions_in_geometry(130, 0, 4))
    var0, const0 = 5, 2
    n_vars, n_consts = var0 + 130 + 12, const0 + 12
    placed = GL.placed(gate, reps, [True] * const0, variables_initial_offset=var0)
    records, first, live = bj.compile_gate_programs([placed], n_vars, 0, n_consts, peephole)
    assert first[0] == 0 and first[1] == len(records) and live <= 128
    var_v = [rnd.randrange(P) for _ in range(n_vars)]
    const_v = [rnd.randrange(P) for _ in range(n_consts)]
    want = GL.evaluate_gate_terms(gate, VR.BaseBackend, lambda i: var_v[i], lambda i: 0, lambda i: const_v[i], reps,
                                  var_base=var0, const_base=const0)
    got = emulate(records, first[0], first[1], var_v + const_v, reps, gate.terms)
    assert got == [int(x) % P for x in want]


def test_peephole_shrinks_the_poseidon2_gate_and_keeps_it_within_32_slots():
    g = GL.placed(GL.poseidon2_flattened_gate(130, 0), 1, [True])
    sizes = {}
    for mode in (0, 1, 3, 7, 15):
        records, first, live = bj.compile_gate_programs([g], 130, 0, 8, mode)
        sizes[mode] = len(records)
        assert live <= 32
    assert sizes[0] == len(g["relations"]) + len(g["writes"]) == 9636 + 118
    assert sizes[1] == 6036 + 118                    # x * 1, x + 0 aliased away
    assert sizes[15] < sizes[7] < sizes[3] < sizes[1]
    assert sizes[15] < 3200                          # linear combinations: a matrix row is one step (+ its term records)


def test_compiler_rejects_bad_programs():
    N = bj.native
    V, T = N.IDX_VARIABLE, N.IDX_TEMPORARY
    base = dict(num_repetitions=1, constants_placement_offset=0, selector_path=[], variables_offset=0, constants_offset=0)
    bad_programs = [
        dict(base, relations=[(N.REL_ADD, 0, (V, 0), (T, 1))], writes=[(T, 0)]),                     # temporary used before definition
        dict(base, relations=[(N.REL_ADD, 0, (V, 0), (V, 1)), (N.REL_ADD, 0, (V, 0), (V, 1))], writes=[(T, 0)]),   # not SSA
        dict(base, relations=[(N.REL_ADD, 0, (V, 0), (V, 9))], writes=[(T, 0)]),                     # column out of range
        dict(base, relations=[(N.REL_ADD, 0, (V, 0), (V, 1))], writes=[(T, 3)]),                     # write of an undefined temporary
    ]
    for g in bad_programs:
        with pytest.raises(bj.BoojumError):
            bj.compile_gate_programs([g], 4, 0, 1)


def _random_program(rnd, n_rel, n_vars, n_consts_per_rep, n_shared):
    """a random SSA program in the recorded form: operands favour recent temporaries (single-use chains, the shapes the peephole
    passes rewrite), immediates include 0, 1, small and full-width values, some terms are bare columns / constants"""
    N = bj.native
    V, C, CS, T, K = N.IDX_VARIABLE, N.IDX_CONSTANT_POLY, N.IDX_CONSTANT_POLY_SHARED, N.IDX_TEMPORARY, N.IDX_CONSTANT_VALUE
    imms = [0, 1, 1, 2, 3, 7, 1 << 20, (1 << 28) - 1, 1 << 28, (1 << 32) - 1, 1 << 32, P - 1, rnd.randrange(P)]
    rel = []

    def operand(i):
        x = rnd.random()
        if i and x < 0.55:
            return (T, i - 1 - min(i - 1, int(rnd.expovariate(0.7))))
        if x < 0.70:
            return (V, rnd.randrange(n_vars))
        if x < 0.78 and n_consts_per_rep:
            return (C, rnd.randrange(n_consts_per_rep))
        if x < 0.84 and n_shared:
            return (CS, rnd.randrange(n_shared))
        return (K, rnd.choice(imms))
    for i in range(n_rel):
        op = rnd.choices([N.REL_ADD, N.REL_SUB, N.REL_MUL, N.REL_DOUBLE, N.REL_NEGATE, N.REL_SQUARE, N.REL_INVERSE],
                         weights=[34, 10, 34, 6, 5, 9, 2])[0]
        binary = op in (N.REL_ADD, N.REL_SUB, N.REL_MUL)
        rel.append((op, i, operand(i), operand(i) if binary else None))
    n_writes = rnd.randrange(1, 6)
    writes = [(T, rnd.randrange(n_rel)) if rnd.random() < 0.85 else rnd.choice([(V, rnd.randrange(n_vars)), (K, rnd.randrange(P))])
              for _ in range(n_writes)]
    if rnd.random() < 0.5:
        writes[-1] = (T, n_rel - 1)
    return rel, writes


def _interpret(rel, writes, var_v, const_v, var_base, shared_base, const_base):
    N = bj.native
    tmp = {}

    def fetch(o):
        kind, val = o
        return {N.IDX_VARIABLE: lambda: var_v[var_base + val], N.IDX_CONSTANT_POLY: lambda: const_v[const_base + val],
                N.IDX_CONSTANT_POLY_SHARED: lambda: const_v[shared_base + val], N.IDX_TEMPORARY: lambda: tmp[val],
                N.IDX_CONSTANT_VALUE: lambda: val % P}[kind]()
    for op, dst, a, b in rel:
        x = fetch(a)
        y = fetch(b) if b is not None else None
        tmp[dst] = {N.REL_ADD: lambda: x + y, N.REL_SUB: lambda: x - y, N.REL_MUL: lambda: x * y, N.REL_DOUBLE: lambda: 2 * x,
                    N.REL_NEGATE: lambda: -x, N.REL_SQUARE: lambda: x * x, N.REL_INVERSE: lambda: pow(x, P - 2, P)}[op]() % P
    return [fetch(w) % P for w in writes]


@pytest.mark.parametrize("seed", range(40))
def test_compiler_on_random_programs(seed):
    """the peephole passes, the push placement and the slot allocation on programs nobody wrote by hand: for random SSA programs
    the compiled steps (every combination of passes) compute what a direct interpretation of the recorded program computes"""
    rnd = random.Random(1000 + seed)
    n_vars_rep, n_consts_rep, n_shared, reps = rnd.randrange(2, 7), rnd.randrange(0, 3), rnd.randrange(0, 3), rnd.randrange(1, 4)
    rel, writes = _random_program(rnd, rnd.randrange(3, 160), n_vars_rep, n_consts_rep, n_shared)
    var0, place = rnd.randrange(0, 3), rnd.randrange(0, 3)
    gate = dict(name="random", relations=rel, writes=writes, num_repetitions=reps, variables_offset=n_vars_rep,
                constants_offset=n_consts_rep, constants_placement_offset=place, selector_path=[True] * place,
                variables_initial_offset=var0)
    n_vars = var0 + n_vars_rep * reps
    n_consts = place + max(n_shared, n_consts_rep * reps, 1)
    var_v = [rnd.randrange(P) for _ in range(n_vars)]
    const_v = [rnd.randrange(P) for _ in range(n_consts)]
    want = []
    for rep in range(reps):
        want += _interpret(rel, writes, var_v, const_v, var0 + rep * n_vars_rep, place, place + rep * n_consts_rep)
    for peephole in (0, 1, 2, 3, 4, 5, 7, 8, 9, 11, 13, 15):
        records, first, live = bj.compile_gate_programs([gate], n_vars, 0, n_consts, peephole)
        got = emulate(records, first[0], first[1], var_v + const_v, reps, len(writes))
        assert got == want, (seed, peephole)
