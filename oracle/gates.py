"""ORACLE (test infrastructure): the reference's gate evaluators and reducing destination restated in Python ints
(small cases only).

  FmaGateInBaseWithoutConstantConstraintEvaluator::evaluate_once   src/cs/gates/fma_gate_without_constant.rs:95-124
  ReductionGateConstraintEvaluator<N>::evaluate_once               src/cs/gates/reduction_gate.rs:104-128
  ConstantAllocatorConstraintEvaluator::evaluate_once              src/cs/gates/constant_allocator.rs:107-126
  RowwiseEvaluator (repetitions, PerChunkOffset)                   src/cs/traits/evaluator.rs:376-397
  push_evaluation_result / proceed_to_next_gate                    src/cs/implementations/buffering_source.rs:158-221, 304-362
  compute_selector_subpath                                         src/cs/implementations/prover.rs:2775-2916
"""
P = 0xFFFFFFFF00000001


def fma_terms(v, c):
    """one repetition: quadratic_coeff * a * b + linear_coeff * c - d ; v = 4 variables, c = [quad, lin]."""
    return [(c[0] * v[0] % P * v[1] + c[1] * v[2] - v[3]) % P]


def reduction_terms(v, c, n=4):
    """sum_i c_i * v_i - v_n."""
    return [(sum(c[i] * v[i] for i in range(n)) - v[n]) % P]


def constant_allocator_terms(v, c):
    return [(v[0] - c[0]) % P]


GATES = {
    # name: (terms fn, principal width in variables, constants used, per-chunk (vars, consts) offsets, shared constants)
    "fma": (fma_terms, 4, 2, (4, 0)),
    "reduction4": (reduction_terms, 5, 4, (5, 0)),
    "constant_allocator": (constant_allocator_terms, 1, 1, (1, 1)),
}


def selector(path, const_row):
    s = 1
    for i, bit in enumerate(path):
        s = s * (const_row[i] if bit else (1 - const_row[i])) % P
    return s


# Index / Relation numbering of recorded programs (gpu_synthesizer/mod.rs:115-133; include/boojum_b200.h BJ_IDX_* / BJ_REL_*)
_VAR, _WIT, _CONST, _TMP, _IMM, _SHARED = range(6)


def program_terms(prog, var_row, const_row, var_base, shared_base, const_base):
    """one repetition of a recorded SSA program (GPUDataCapture: dict(relations, writes)) over base-field values of one point"""
    tmp = {}

    def fetch(o):
        kind, val = o
        if kind == _VAR:
            return var_row[var_base + val]
        if kind == _CONST:
            return const_row[const_base + val]
        if kind == _SHARED:
            return const_row[shared_base + val]
        if kind == _TMP:
            return tmp[val]
        assert kind == _IMM, "witness columns are not used by these circuits"
        return int(val) % P

    for op, dst, a, b in prog["relations"]:
        x = fetch(a)
        if op == 0:
            r = x + fetch(b)
        elif op == 1:
            r = 2 * x
        elif op == 2:
            r = x - fetch(b)
        elif op == 3:
            r = -x
        elif op == 4:
            r = x * fetch(b)
        elif op == 5:
            r = x * x
        else:
            assert op == 6
            r = pow(x, P - 2, P)
        tmp[dst] = r % P
    return [fetch(w) for w in prog["writes"]]


def quotient_gates_row(gates, var_row, const_row, alphas):
    """gates: list of (name, num_repetitions, selector_path[, first variable column, first constant column[, program]]) - the
    three bench gates by name, any other gate as its recorded program (specialised-column gates have an empty path and their own
    first columns).  Returns the (c0, c1) contribution of one point."""
    q0 = q1 = 0
    k = 0
    for g in gates:
        name, reps, path = g[0], g[1], g[2]
        var0 = g[3] if len(g) > 3 else 0
        place = g[4] if len(g) > 4 else len(path)
        a0 = a1 = 0
        for rep in range(reps):
            if len(g) > 5:
                prog = g[5]
                terms = program_terms(prog, var_row, const_row, var0 + rep * prog["variables_offset"], place,
                                      place + rep * prog["constants_offset"])
            else:
                fn, width, _, (voff, coff) = GATES[name]
                terms = fn(var_row[var0 + rep * voff: var0 + rep * voff + width], const_row[place + rep * coff:])
            for term in terms:
                a0 = (a0 + term * alphas[k][0]) % P
                a1 = (a1 + term * alphas[k][1]) % P
                k += 1
        s = selector(path, const_row)
        q0 = (q0 + s * a0) % P
        q1 = (q1 + s * a1) % P
    return q0, q1
