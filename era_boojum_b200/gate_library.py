"""Host-side mirror of the reference's gate evaluators (GateConstraintEvaluator::evaluate_once, src/cs/traits/evaluator.rs:105-250)
for the gate set the reference captures for its GPU hook (src/gpu_synthesizer/mod.rs:826-838) and verifies proof.json with
(src/gadgets/recursion/recursive_verifier.rs:2290-2368).

Like the reference, every evaluator is written ONCE, generically over a field-like backend `F` (PrimeFieldLike,
src/field/traits/field_like.rs): the same function
  * runs over the recording backend `Recorder` below - the role of gpu_synthesizer::GPUVariablesContext (mod.rs:135-352) - and
    yields the SSA program (Index / Relation lists, GPUDataCapture, mod.rs:354-443) that bj_quotient_gates_general_purpose
    interprets on the device;
  * runs over plain field values (the tests' and the oracle verifier's backends: base field on random rows, Fp2 at the
    challenge point z) - which is how these transcriptions are pinned: the quotient identity at z on the reference's own
    proof.json / vk.json only holds if every evaluator below computes what the reference's evaluator computes
    (tests/test_reference_quotient.py).

A backend provides: zero(), one(), constant(u64), add(a, b), sub(a, b), mul(a, b), double(a), negate(a), square(a).
A trace source provides: var(i), wit(i), const(i)  (indices relative to the current repetition, as TraceSource does).
"""
from . import native as N

P = 0xFFFFFFFF00000001


# ---------------------------------------------------------------------------------------------- helpers over a backend -----
def mul_acc(F, acc, a, b):
    """PrimeFieldLike::mul_and_accumulate_into: acc += a * b"""
    return F.add(acc, F.mul(a, b))


def small_pow7(F, x):
    """PrimeFieldLike::small_pow(7) (x^7; Poseidon2's NONLINEARITY_DEGREE)"""
    x2 = F.square(x)
    x3 = F.mul(x2, x)
    x4 = F.square(x2)
    return F.mul(x4, x3)


# --------------------------------------------------------------------------------------------------- gate definitions -----
class Gate:
    """One evaluator instance.  width = (variables, witnesses, constants) of the principal instance; offsets = PerChunkOffset of
    GatePlacementType::MultipleOnRow (None = UniqueOnRow); shared = how many row-shared constants load_row_shared_constants
    reads (constants 0..shared-1 of the gate, read once per row); terms = num_quotient_terms per repetition."""

    def __init__(self, name, width, terms, degree, offsets, shared, evaluate, required_constants=None, cite=""):
        self.name, self.width, self.terms, self.degree = name, width, terms, degree
        self.offsets, self.shared, self.evaluate, self.cite = offsets, shared, evaluate, cite
        self.required_constants = width[2] if required_constants is None else required_constants

    def num_repetitions_in_geometry(self, num_variables, num_witnesses, num_constants):
        """num_repetitions_in_geometry of the evaluators below (all: columns available // principal width; the constants
        allocator is also bounded by the constant columns)."""
        if self.offsets is None:
            return 1
        v, w, c = self.width
        lim = num_variables // v if v else 1 << 30
        if w:
            lim = min(lim, num_witnesses // w)
        if self.name == "constant_allocator":
            lim = min(lim, num_constants)
        return lim


def _constant_allocator(F, src, push, shared):
    # src/cs/gates/constant_allocator.rs:107-126
    push(F.sub(src.var(0), src.const(0)))


def _boolean(F, src, push, shared):
    # src/cs/gates/boolean_allocator.rs:104-123:  a * (1 - a)
    a = src.var(0)
    push(F.mul(a, F.sub(F.one(), a)))


def _fma(F, src, push, shared):
    # src/cs/gates/fma_gate_without_constant.rs:95-124:  c * lin + quad * (a * b) - d
    a, b, c, d = (src.var(i) for i in range(4))
    quad, lin = shared
    contribution = F.mul(c, lin)
    contribution = mul_acc(F, contribution, quad, F.mul(a, b))
    push(F.sub(contribution, d))


def _make_reduction(n):
    def ev(F, src, push, shared):
        # src/cs/gates/reduction_gate.rs:104-128:  sum_i var_i * c_i - result
        acc = F.zero()
        for i in range(n):
            acc = mul_acc(F, acc, src.var(i), shared[i])
        push(F.sub(acc, src.var(n)))
    return ev


def _make_dot_product(n):
    def ev(F, src, push, shared):
        # src/cs/gates/dot_product_gate.rs:  sum_i a_i * b_i - result
        acc = F.zero()
        for i in range(n):
            acc = mul_acc(F, acc, src.var(2 * i), src.var(2 * i + 1))
        push(F.sub(acc, src.var(2 * n)))
    return ev


def _zero_check(F, src, push, shared):
    # src/cs/gates/zero_check.rs (use_witness_column_for_inversion = false): flag + input * inv - 1 ; input * flag
    inp, flag, inv = src.var(0), src.var(1), src.var(2)
    push(F.sub(mul_acc(F, flag, inp, inv), F.one()))
    push(F.mul(inp, flag))


def _uintx_add(F, src, push, shared):
    # src/cs/gates/uintx_add.rs: a + b + carry_in - c - shift * carry_out ; carry_out^2 - carry_out
    (shift,) = shared
    a, b, cin, c, cout = (src.var(i) for i in range(5))
    t = F.sub(F.add(F.add(a, b), cin), c)
    push(F.sub(t, F.mul(shift, cout)))
    push(F.sub(F.mul(cout, cout), cout))


def _select(F, a, b, selector, result):
    t = F.mul(a, selector)
    t = mul_acc(F, t, F.sub(F.one(), selector), b)
    return F.sub(t, result)


def _selection(F, src, push, shared):
    # src/cs/gates/selection_gate.rs: a * s + (1 - s) * b - result
    push(_select(F, src.var(0), src.var(1), src.var(2), src.var(3)))


def _make_parallel_selection(n):
    def ev(F, src, push, shared):
        # src/cs/gates/parallel_selection.rs: one selector, n (a, b, result) triples
        s = src.var(0)
        for i in range(n):
            push(_select(F, src.var(3 * i + 1), src.var(3 * i + 2), s, src.var(3 * i + 3)))
    return ev


def _u8x4_fma(F, src, push, shared):
    # src/cs/gates/u32_fma.rs: a * b + c + carry_in = low + 2^32 high over four u8 limbs each, two relations
    k = lambda v: F.constant(v % P)
    s8, s16, s24 = k(1 << 8), k(1 << 16), k(1 << 24)
    m1, m8, m16, m24, m32, m40 = k(P - 1), k(P - (1 << 8)), k(P - (1 << 16)), k(P - (1 << 24)), k(P - (1 << 32)), k(P - (1 << 40))
    a = [src.var(i) for i in range(4)]
    b = [src.var(4 + i) for i in range(4)]
    c = [src.var(8 + i) for i in range(4)]
    carry = [src.var(12 + i) for i in range(4)]
    low = [src.var(16 + i) for i in range(4)]
    high = [src.var(20 + i) for i in range(4)]
    pc0, pc1 = src.var(24), src.var(25)
    t = c[0]
    t = mul_acc(F, t, c[1], s8)
    t = mul_acc(F, t, c[2], s16)
    t = mul_acc(F, t, c[3], s24)
    t = F.add(t, carry[0])
    t = mul_acc(F, t, carry[1], s8)
    t = mul_acc(F, t, carry[2], s16)
    t = mul_acc(F, t, carry[3], s24)
    t = mul_acc(F, t, low[0], m1)
    t = mul_acc(F, t, low[1], m8)
    t = mul_acc(F, t, low[2], m16)
    t = mul_acc(F, t, low[3], m24)
    t = mul_acc(F, t, a[0], b[0])
    u = F.mul(a[1], b[0])
    u = mul_acc(F, u, a[0], b[1])
    t = mul_acc(F, t, u, s8)
    u = F.mul(a[2], b[0])
    u = mul_acc(F, u, a[1], b[1])
    u = mul_acc(F, u, a[0], b[2])
    t = mul_acc(F, t, u, s16)
    u = F.mul(a[3], b[0])
    u = mul_acc(F, u, a[2], b[1])
    u = mul_acc(F, u, a[1], b[2])
    u = mul_acc(F, u, a[0], b[3])
    t = mul_acc(F, t, u, s24)
    t = mul_acc(F, t, pc0, m32)
    t = mul_acc(F, t, pc1, m40)
    push(t)
    t = pc0
    t = mul_acc(F, t, pc1, s8)
    t = mul_acc(F, t, high[0], m1)
    t = mul_acc(F, t, high[1], m8)
    t = mul_acc(F, t, high[2], m16)
    t = mul_acc(F, t, high[3], m24)
    u = F.mul(a[3], b[1])
    u = mul_acc(F, u, a[2], b[2])
    u = mul_acc(F, u, a[1], b[3])
    t = F.add(t, u)
    u = F.mul(a[3], b[2])
    u = mul_acc(F, u, a[2], b[3])
    t = mul_acc(F, t, u, s8)
    t = mul_acc(F, t, F.mul(a[3], b[3]), s16)
    push(t)


# Poseidon2 over Goldilocks, t = 12: parameters (src/implementations/poseidon2/params.rs; constants from the library's own table)
_M4 = [[5, 7, 1, 3], [4, 6, 1, 1], [1, 3, 5, 7], [1, 1, 4, 6]]
POSEIDON2_EXTERNAL_MATRIX = [[(2 if i // 4 == j // 4 else 1) * _M4[i % 4][j % 4] for j in range(12)] for i in range(12)]
_DIAG_SHIFTS = [4, 14, 11, 8, 0, 5, 2, 9, 13, 6, 3, 12]
POSEIDON2_INTERNAL_MATRIX = [[(1 << _DIAG_SHIFTS[i]) + 1 if i == j else 1 for j in range(12)] for i in range(12)]


def poseidon_round_constants():
    """the 360 Poseidon constants the device kernels use (csrc/poseidon_rc.h), RC[round * 12 + i]"""
    import os
    import re
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "poseidon_rc.h")
    vals = [int(v, 16) for v in re.findall(r"0x([0-9a-f]{16})ull", open(path).read())]
    assert len(vals) == 360
    return vals


def _matmul(F, matrix, state):
    out = []
    for row in matrix:
        acc = F.zero()
        for coeff, s in zip(row, state):
            acc = mul_acc(F, acc, s, F.constant(coeff))
        out.append(acc)
    return out


def _make_poseidon2_flattened(num_copiable, num_witness):
    """Poseidon2RoundFunctionFlattenedEvaluator::evaluate_once (src/cs/gates/poseidon2.rs:166-403): state in 12 variables,
    output in the next 12, then one variable per degree reset (witness columns first, then copiable ones)."""
    rc = poseidon_round_constants()
    full = [rc[12 * r: 12 * r + 12] for r in (0, 1, 2, 3, 26, 27, 28, 29)]
    partial = [rc[12 * r] for r in range(4, 26)]

    def ev(F, src, push, shared):
        state = [src.var(i) for i in range(12)]
        output = [src.var(12 + i) for i in range(12)]
        pos = {"v": 24, "w": 0}

        def next_reset_var():
            if pos["w"] < num_witness:
                pos["w"] += 1
                return src.wit(pos["w"] - 1)
            assert pos["v"] < num_copiable
            pos["v"] += 1
            return src.var(pos["v"] - 1)

        def reset(state):
            new = []
            for s in state:
                v = next_reset_var()
                push(F.sub(s, v))
                new.append(v)
            return new

        for rnd in range(4):
            if rnd != 0:
                state = reset(state)
            else:
                state = _matmul(F, POSEIDON2_EXTERNAL_MATRIX, state)
            state = [small_pow7(F, F.add(s, F.constant(full[rnd][i]))) for i, s in enumerate(state)]
            state = _matmul(F, POSEIDON2_EXTERNAL_MATRIX, state)
        for rnd in range(22):
            s0 = F.add(state[0], F.constant(partial[rnd]))
            v = next_reset_var()
            push(F.sub(s0, v))
            state[0] = small_pow7(F, v)
            state = _matmul(F, POSEIDON2_INTERNAL_MATRIX, state)
        for rnd in range(4, 8):
            state = reset(state)
            state = [small_pow7(F, F.add(s, F.constant(full[rnd][i]))) for i, s in enumerate(state)]
            state = _matmul(F, POSEIDON2_EXTERNAL_MATRIX, state)
        for s, o in zip(state, output):
            push(F.sub(o, s))
    return ev


def poseidon2_flattened_gate(num_copiable=130, num_witness=0):
    """Poseidon2FlattenedGate<F, 8, 12, 4, Poseidon2Goldilocks> for a geometry (compute_strategy, poseidon2.rs:502-528):
    130 variables in total = 12 in + 12 out + 36 + 22 + 48 degree resets, 118 terms."""
    assert num_copiable + num_witness == 130 and num_copiable >= 24
    return Gate("poseidon2_flattened", (num_copiable, num_witness, 0), 118, 7, (num_copiable, num_witness, 0), 0,
                _make_poseidon2_flattened(num_copiable, num_witness), cite="src/cs/gates/poseidon2.rs:166-403")


CONSTANT_ALLOCATOR = Gate("constant_allocator", (1, 0, 1), 1, 1, (1, 0, 1), 0, _constant_allocator, cite="src/cs/gates/constant_allocator.rs")
BOOLEAN = Gate("boolean", (1, 0, 0), 1, 2, (1, 0, 0), 0, _boolean, cite="src/cs/gates/boolean_allocator.rs")
FMA = Gate("fma", (4, 0, 2), 1, 3, (4, 0, 0), 2, _fma, cite="src/cs/gates/fma_gate_without_constant.rs")
REDUCTION4 = Gate("reduction4", (5, 0, 4), 1, 2, (5, 0, 0), 4, _make_reduction(4), cite="src/cs/gates/reduction_gate.rs")
DOT_PRODUCT4 = Gate("dot_product4", (9, 0, 0), 1, 2, (9, 0, 0), 0, _make_dot_product(4), cite="src/cs/gates/dot_product_gate.rs")
ZERO_CHECK = Gate("zero_check", (3, 0, 0), 2, 2, (3, 0, 0), 0, _zero_check, cite="src/cs/gates/zero_check.rs")
UINTX_ADD = Gate("uintx_add", (5, 0, 1), 2, 2, (5, 0, 0), 1, _uintx_add, cite="src/cs/gates/uintx_add.rs")
SELECTION = Gate("selection", (4, 0, 0), 1, 2, (4, 0, 0), 0, _selection, cite="src/cs/gates/selection_gate.rs")
PARALLEL_SELECTION4 = Gate("parallel_selection4", (13, 0, 0), 4, 2, (13, 0, 0), 0, _make_parallel_selection(4), cite="src/cs/gates/parallel_selection.rs")
U8X4_FMA = Gate("u8x4_fma", (26, 0, 0), 2, 2, (26, 0, 0), 0, _u8x4_fma, cite="src/cs/gates/u32_fma.rs")
# markers: no quotient terms (they only take a place in the selector tree)
PUBLIC_INPUT = Gate("public_input", (1, 0, 0), 0, 0, None, 0, lambda F, src, push, shared: None, cite="src/cs/gates/public_input.rs")
NOP = Gate("nop", (0, 0, 0), 0, 0, None, 0, lambda F, src, push, shared: None, cite="src/cs/gates/nop_gate.rs")


# -------------------------------------------------------------------------------------------- evaluating over a backend -----
class _Source:
    """TraceSource view of one repetition: indices are relative to (var_base, wit_base, const_base)."""

    def __init__(self, var, wit, const):
        self.var, self.wit, self.const = var, wit, const


def evaluate_gate_terms(gate, F, get_var, get_wit, get_const, num_repetitions, var_base=0, wit_base=0, const_base=0):
    """RowwiseEvaluator / ColumnwiseEvaluator (src/cs/traits/evaluator.rs:376-397, 310-374): the terms of all repetitions of
    `gate` in push order.  get_*(absolute column index) -> backend value.  const_base = the gate's first constant column
    (selector path length for general-purpose placement); row-shared constants are read once at const_base."""
    shared = [get_const(const_base + i) for i in range(gate.shared)]
    out = []
    off = gate.offsets or (0, 0, 0)
    for rep in range(num_repetitions):
        vb, wb, cb = var_base + rep * off[0], wit_base + rep * off[1], const_base + rep * off[2]
        src = _Source(lambda i, vb=vb: get_var(vb + i), lambda i, wb=wb: get_wit(wb + i), lambda i, cb=cb: get_const(cb + i))
        before = len(out)
        gate.evaluate(F, src, out.append, shared)
        assert len(out) - before == gate.terms, (gate.name, len(out) - before)
    return out


# ----------------------------------------------------------------------------------------- the recording backend (SSA) -----
class Recorder:
    """Records Relation(dst temporary, a, b) lists with Index operands - what gpu_synthesizer::GPUVariablesContext collects.
    Values are Index tuples (kind, value)."""

    def __init__(self):
        self.relations = []

    def _emit(self, op, a, b=None):
        dst = len(self.relations)
        self.relations.append((op, dst, a, b))
        return (N.IDX_TEMPORARY, dst)

    def zero(self):
        return (N.IDX_CONSTANT_VALUE, 0)

    def one(self):
        return (N.IDX_CONSTANT_VALUE, 1)

    def constant(self, v):
        return (N.IDX_CONSTANT_VALUE, v % P)

    def add(self, a, b):
        return self._emit(N.REL_ADD, a, b)

    def sub(self, a, b):
        return self._emit(N.REL_SUB, a, b)

    def mul(self, a, b):
        return self._emit(N.REL_MUL, a, b)

    def double(self, a):
        return self._emit(N.REL_DOUBLE, a)

    def negate(self, a):
        return self._emit(N.REL_NEGATE, a)

    def square(self, a):
        return self._emit(N.REL_SQUARE, a)


def capture(gate):
    """GPUDataCapture::from_evaluator (src/gpu_synthesizer/mod.rs:387-443) for one repetition: dict(relations, writes, ...) in
    the form Context.evaluate_gates_over_general_purpose_columns / native_setup take."""
    rec = Recorder()
    shared = [(N.IDX_CONSTANT_POLY_SHARED, i) for i in range(gate.shared)]
    src = _Source(lambda i: (N.IDX_VARIABLE, i), lambda i: (N.IDX_WITNESS, i), lambda i: (N.IDX_CONSTANT_POLY, i))
    writes = []
    gate.evaluate(rec, src, writes.append, shared)
    assert len(writes) == gate.terms
    off = gate.offsets or (0, 0, 0)
    fixed = []
    for w in writes:      # a term that is a bare column / constant is routed through a temporary, as the reference's hook does
        if w[0] != N.IDX_TEMPORARY:
            w = rec.add(w, rec.zero())
        fixed.append(w)
    return dict(name=gate.name, relations=list(rec.relations), writes=fixed, variables_offset=off[0], witnesses_offset=off[1],
                constants_offset=off[2])


def placed(gate, num_repetitions, selector_path, constants_placement_offset=None, variables_initial_offset=0,
           witnesses_initial_offset=0):
    """capture() + placement data of one gate of a circuit (general-purpose: selector path, constants start behind it;
    specialised columns: empty path and the initial offsets)."""
    d = capture(gate)
    d.update(num_repetitions=num_repetitions, selector_path=list(selector_path),
             constants_placement_offset=len(selector_path) if constants_placement_offset is None else constants_placement_offset,
             variables_initial_offset=variables_initial_offset, witnesses_initial_offset=witnesses_initial_offset)
    return d
