"""The oracle verifier's interpreter of recorded gate programs (oracle/verifier.py::_program_terms - how it checks circuits whose
gates travel as GPUDataCapture programs, e.g. the production-shaped circuit) against the evaluators themselves run over Fp2
(era_boojum_b200/gate_library.py, pinned by the quotient identity on the reference's proof.json): every gate type of the
fixture, several repetitions, random Fp2 'openings'."""
import random

import pytest

from era_boojum_b200 import gate_library as GL
from oracle import verifier as OV
from oracle import verifier_reference as VR

P = VR.P
GATES = [GL.CONSTANT_ALLOCATOR, GL.BOOLEAN, GL.FMA, GL.REDUCTION4, GL.DOT_PRODUCT4, GL.ZERO_CHECK, GL.UINTX_ADD, GL.SELECTION,
         GL.PARALLEL_SELECTION4, GL.U8X4_FMA, GL.poseidon2_flattened_gate(130, 0)]


@pytest.mark.parametrize("gate", GATES, ids=lambda g: g.name)
def test_program_interpreter_equals_evaluator(gate):
    rnd = random.Random(hash(gate.name) & 0xFFFF)
    reps = min(3, gate.num_repetitions_in_geometry(130, 0, 4))
    var_base, const_base = 2, 3
    nv = var_base + 130 + 8
    var_v = [(rnd.randrange(P), rnd.randrange(P)) for _ in range(nv)]
    const_v = [(rnd.randrange(P), rnd.randrange(P)) for _ in range(const_base + 16)]
    want = GL.evaluate_gate_terms(gate, VR.Fp2Backend, lambda i: var_v[i], lambda i: (0, 0), lambda i: const_v[i], reps,
                                  var_base=var_base, const_base=const_base)
    prog = GL.capture(gate)
    got = []
    for rep in range(reps):
        got += OV._program_terms(prog, var_v, const_v, var_base + rep * prog["variables_offset"],
                                 (const_base, const_base + rep * prog["constants_offset"]))
    assert len(got) == gate.terms * reps and got == want


def test_vk_carries_programs_for_gates_the_verifier_does_not_know_by_name():
    from era_boojum_b200 import prover
    cfg = prover.ProofConfig(fri_lde_factor=2, merkle_tree_cap_size=32)
    gates = [GL.placed(GL.BOOLEAN, 1, [], constants_placement_offset=8, variables_initial_offset=154), GL.placed(GL.FMA, 32, [False, True])]
    import numpy as np
    vk = prover.verification_key(10, 155, 8, gates, 8, cfg, None, [], np.zeros((32, 4), np.uint64))
    assert len(vk["gates"][0]) == 6 and vk["gates"][0][5]["relations"] and len(vk["gates"][1]) == 5
