"""End to end on the GPU: synthetic SHA-bench-shaped circuit -> setup -> prove (every heavy step through the C-ABI) ->
the oracle's restatement of the reference verifier accepts the proof; tampered proofs are rejected.  This is the
prove-then-verify pattern of the reference's own integration tests (src/cs/implementations/cs.rs:1075-1203)."""
import copy
import json

import numpy as np
import pytest

from oracle import verifier as OV

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch
    assert torch.cuda.is_available()
    import era_boojum_b200 as bj
    from era_boojum_b200 import prover, synthetic
    ctx = bj.Context.on_current_stream(0)
    yield bj, ctx, prover, synthetic
    ctx.synchronize()
    ctx.close()


def _prove(env, log_n, V, lde=8, cap=16, seed=0):
    bj, ctx, prover, synthetic = env
    variables, sigmas, constants, gates, Q = synthetic.generate(ctx, log_n, V, seed)
    cfg = prover.ProofConfig(fri_lde_factor=lde, merkle_tree_cap_size=cap, security_level=100)
    setup = prover.Setup(ctx, sigmas, constants, gates, Q, cfg)
    proof = prover.prove(ctx, setup, variables)
    return setup.vk(), proof


@pytest.mark.parametrize("log_n,V,lde,cap", [(8, 20, 8, 16), (10, 60, 8, 16), (9, 60, 4, 8), (12, 60, 8, 16)])
def test_prove_then_verify(env, log_n, V, lde, cap):
    vk, proof = _prove(env, log_n, V, lde, cap, seed=log_n)
    json.dumps(proof)  # serialisable in the reference's shape
    assert OV.verify(vk, proof)
    assert len(proof["queries_per_fri_repetition"]) == -(-100 // (lde.bit_length() - 1))


@pytest.mark.parametrize("log_n", [8, 11])
def test_prove_then_verify_with_lookup_argument(env, log_n):
    """the bench geometry incl. the log-derivative lookup: 60 gp + 8x4 specialised lookup columns, table id in a constant."""
    bj, ctx, prover, synthetic = env
    variables, sigmas, constants, gates, Q, lk = synthetic.generate(ctx, log_n, 60, seed=log_n, lookup=True)
    cfg = prover.ProofConfig(fri_lde_factor=8, merkle_tree_cap_size=16, security_level=100)
    setup = prover.Setup(ctx, sigmas, constants, gates, Q, cfg, lookup=lk)
    proof = prover.prove(ctx, setup, variables, multiplicities=lk["multiplicities"])
    vk = setup.vk()
    assert len(proof["values_at_0"]) == 9 and len(proof["queries_per_fri_repetition"][0]["witness_query"]["leaf_elements"]) == 93
    assert OV.verify(vk, proof)
    bad = copy.deepcopy(proof)
    bad["values_at_0"][2]["coeffs"][1] ^= 1
    with pytest.raises(AssertionError):
        OV.verify(vk, bad)


def test_tampered_proofs_are_rejected(env):
    vk, proof = _prove(env, 8, 20, seed=5)
    assert OV.verify(vk, proof)
    bad = copy.deepcopy(proof)
    bad["values_at_z"][3]["coeffs"][0] ^= 1
    with pytest.raises(AssertionError):
        OV.verify(vk, bad)
    bad = copy.deepcopy(proof)
    bad["queries_per_fri_repetition"][0]["witness_query"]["leaf_elements"][0] ^= 1
    with pytest.raises(AssertionError):
        OV.verify(vk, bad)
    bad = copy.deepcopy(proof)
    bad["final_fri_monomials"][0][0] ^= 1
    with pytest.raises(AssertionError):
        OV.verify(vk, bad)
    bad = copy.deepcopy(proof)
    bad["quotient_oracle_cap"][0][0] ^= 1
    with pytest.raises(AssertionError):
        OV.verify(vk, bad)


def test_unsatisfied_circuit_fails_in_prover(env):
    """a trace that violates a gate trips the reference's guard on the top quotient coefficient (prover.rs:1425-1438)."""
    bj, ctx, prover, synthetic = env
    variables, sigmas, constants, gates, Q = synthetic.generate(ctx, 8, 20, 3)
    cfg = prover.ProofConfig()
    setup = prover.Setup(ctx, sigmas, constants, gates, Q, cfg)
    variables[3, 13] += 1   # column 3 is constrained by every gate type (and tied by a copy constraint in fma rows)
    with pytest.raises((ValueError, bj.BoojumError)):
        prover.prove(ctx, setup, variables)


def _prove_sharded_threads(bj, prover, synthetic, world, log_n, V, lde, cap, seed, lookup):
    """`world` coset shards as threads of this process on one GPU (parallel.ThreadComm): each thread owns a Context
    carrying its shard, a Setup and runs the same prove(); returns every rank's (vk, proof)."""
    import threading
    from era_boojum_b200 import parallel
    shared = parallel.ThreadComm(world)
    out, errs = [None] * world, []

    def run(rank):
        try:
            ctx = bj.Context(0)
            ctx.set_coset_shard(rank, world, lde)
            gen = synthetic.generate(ctx, log_n, V, seed=seed, lookup=lookup)
            lk = gen[5] if lookup else None
            variables, sigmas, constants, gates, Q = gen[:5]
            cfg = prover.ProofConfig(fri_lde_factor=lde, merkle_tree_cap_size=cap, security_level=100)
            setup = prover.Setup(ctx, sigmas, constants, gates, Q, cfg, lookup=lk, comm=shared.rank_view(rank))
            proof = prover.prove(ctx, setup, variables, multiplicities=lk["multiplicities"] if lk else None)
            out[rank] = (setup.vk(), proof)
            ctx.synchronize()
            ctx.close()
        except BaseException as e:       # a dead rank must not leave the others waiting at a barrier
            errs.append(e)
            shared._barrier.abort()

    ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    if errs:
        raise errs[0]
    return out


@pytest.mark.parametrize("world,log_n,V,lde,cap,lookup", [(2, 9, 20, 8, 16, False), (4, 8, 60, 8, 16, True), (8, 8, 20, 8, 16, False),
                                                          (2, 10, 60, 4, 8, True), (2, 10, 60, 2, 16, True)])   # last: LDE factor 2 < quotient degree 4
def test_coset_sharded_prover_equals_single_gpu(env, world, log_n, V, lde, cap, lookup):
    """the multi-GPU decomposition (every rank holds the cosets j = rank mod world; caps, quotient cosets, openings, the
    last FRI codeword and the query answers are exchanged) yields the SAME proof as the single-context prover."""
    bj, ctx, prover, synthetic = env
    gen = synthetic.generate(ctx, log_n, V, seed=3, lookup=lookup)
    lk = gen[5] if lookup else None
    variables, sigmas, constants, gates, Q = gen[:5]
    cfg = prover.ProofConfig(fri_lde_factor=lde, merkle_tree_cap_size=cap, security_level=100)
    setup = prover.Setup(ctx, sigmas, constants, gates, Q, cfg, lookup=lk)
    ref = prover.prove(ctx, setup, variables, multiplicities=lk["multiplicities"] if lk else None)
    assert OV.verify(setup.vk(), ref)
    res = _prove_sharded_threads(bj, prover, synthetic, world, log_n, V, lde, cap, 3, lookup)
    for vk, proof in res:
        assert vk == setup.vk()
        assert json.dumps(proof, sort_keys=True) == json.dumps(ref, sort_keys=True)


def _prove_native_sharded_threads(bj, synthetic, prover, world, log_n, V, lde, cap, seed, lookup, hasher="poseidon2", transcript="poseidon2",
                                  public_inputs=()):
    """the library's own sharded driver (bj_setup_create / bj_prove on contexts that carry a bj_comm): `world` ranks as threads
    on one GPU over the local transport; returns every rank's (cap, proof)."""
    import threading
    group = bj.Comm.local_group(world)
    out, errs = [None] * world, []

    def run(rank):
        try:
            ctx = bj.Context(0)
            comm = bj.Comm.local(ctx, group, rank, world, lde)
            gen = synthetic.generate(ctx, log_n, V, seed=seed, lookup=lookup)
            lk = gen[5] if lookup else None
            variables, sigmas, constants, gates, Q = gen[:5]
            cfg = prover.ProofConfig(fri_lde_factor=lde, merkle_tree_cap_size=cap, security_level=100, hasher=hasher, transcript=transcript)
            nat = ctx.native_setup(sigmas.contiguous(), constants.contiguous(), gates, Q, cfg, lookup=lk, public_inputs=public_inputs)
            proof = nat.prove(variables.contiguous(), lk["multiplicities"] if lk else None)
            out[rank] = (nat.get_cap(), proof)
            ctx.synchronize()
            nat.close()
            comm.close()
            ctx.close()
        except BaseException as e:
            errs.append(e)

    ts = [threading.Thread(target=run, args=(r,), daemon=True) for r in range(world)]
    [t.start() for t in ts]
    [t.join(timeout=600) for t in ts]
    if errs:
        raise errs[0]
    assert all(o is not None for o in out), "a rank did not finish"
    bj.Comm.destroy_local_group(group)
    return out


@pytest.mark.parametrize("world,log_n,V,lde,cap,lookup,hasher,transcript", [
    (2, 9, 20, 8, 16, False, "poseidon2", "poseidon2"), (4, 8, 60, 8, 16, True, "poseidon2", "poseidon"),
    (8, 8, 20, 8, 16, False, "blake2s", "blake2s"), (2, 10, 60, 4, 8, True, "blake2s", "blake2s"), (8, 9, 60, 8, 32, True, "poseidon2", "poseidon2"),
    (2, 10, 60, 2, 16, True, "poseidon2", "poseidon2")])   # last: LDE factor 2 < quotient degree 4 (the shape of the reference's proof.json)
def test_native_sharded_prover_equals_single_gpu(env, world, log_n, V, lde, cap, lookup, hasher, transcript):
    """bj_prove on coset-sharded contexts (communicator: local transport, the NCCL transport takes the same code path) returns
    on every rank exactly the single-GPU proof - world 2 / 4 / 8 incl. world > quotient degree (ranks that own no quotient
    coset), lookup argument, public inputs, both hashers."""
    bj, ctx, prover, synthetic = env
    gen = synthetic.generate(ctx, log_n, V, seed=3, lookup=lookup)
    lk = gen[5] if lookup else None
    variables, sigmas, constants, gates, Q = gen[:5]
    pis = [(1, 3), (5, 3)] if world == 4 else []
    cfg = prover.ProofConfig(fri_lde_factor=lde, merkle_tree_cap_size=cap, security_level=100, hasher=hasher, transcript=transcript)
    nat = ctx.native_setup(sigmas.contiguous(), constants.contiguous(), gates, Q, cfg, lookup=lk, public_inputs=pis)
    ref = nat.prove(variables.contiguous(), lk["multiplicities"] if lk else None)
    assert OV.verify(nat.vk(), ref)
    ref_cap = nat.get_cap()
    nat.close()
    res = _prove_native_sharded_threads(bj, synthetic, prover, world, log_n, V, lde, cap, 3, lookup, hasher, transcript, pis)
    for cap_r, proof in res:
        assert np.array_equal(cap_r, ref_cap)
        assert json.dumps(proof, sort_keys=True) == json.dumps(ref, sort_keys=True)


def test_comm_collectives_local_transport(env):
    """bj_comm_all_gather / all_gather_host / broadcast_host over the local transport (4 thread-ranks on one GPU)."""
    import threading
    import torch
    bj, ctx, prover, synthetic = env
    world = 4
    group = bj.Comm.local_group(world)
    res, errs = [None] * world, []

    def run(rank):
        try:
            c = bj.Context(0)
            comm = bj.Comm.local(c, group, rank, world, 8)
            send = torch.full((1000,), rank + 1, dtype=torch.int64, device="cuda:0")
            recv = torch.zeros((world, 1000), dtype=torch.int64, device="cuda:0")
            comm.all_gather(send, recv)
            host = comm.all_gather_host(np.array([rank * 10, rank * 10 + 1], dtype=np.uint64))
            b = comm.broadcast_host(np.array([7, 8, 9] if rank == 2 else [0, 0, 0], dtype=np.uint64), root=2)
            res[rank] = (recv.cpu().numpy(), host, b)
            comm.close()
            c.close()
        except BaseException as e:
            errs.append(e)

    ts = [threading.Thread(target=run, args=(r,), daemon=True) for r in range(world)]
    [t.start() for t in ts]
    [t.join(timeout=120) for t in ts]
    assert not errs, errs
    for recv, host, b in res:
        assert all((recv[r] == r + 1).all() for r in range(world))
        assert host.tolist() == [[r * 10, r * 10 + 1] for r in range(world)]
        assert b.tolist() == [7, 8, 9]
    bj.Comm.destroy_local_group(group)


def test_local_comm_python_fri_equals_cxx_driver(env):
    """world of one through the communicator path (Python-level FRI loop) == bj_do_fri driver."""
    bj, ctx, prover, synthetic = env
    from era_boojum_b200 import parallel
    variables, sigmas, constants, gates, Q = synthetic.generate(ctx, 9, 20, seed=4)
    cfg = prover.ProofConfig(fri_lde_factor=8, merkle_tree_cap_size=16, security_level=100)
    a = prover.prove(ctx, prover.Setup(ctx, sigmas, constants, gates, Q, cfg), variables)
    b = prover.prove(ctx, prover.Setup(ctx, sigmas, constants, gates, Q, cfg, comm=parallel.LocalComm()), variables)
    assert json.dumps(a, sort_keys=True) == json.dumps(b, sort_keys=True)


@pytest.mark.parametrize("log_n,lookup", [(9, False), (10, True)])
def test_native_cxx_prover_equals_python_driver(env, log_n, lookup):
    """bj_setup_create + bj_prove (host C++ inside the library, JSON in the reference's serde shape) produce the same proof
    as the Python driver over the same entry points, and the oracle verifier accepts it."""
    bj, ctx, prover, synthetic = env
    gen = synthetic.generate(ctx, log_n, 60, seed=11, lookup=lookup)
    lk = gen[5] if lookup else None
    variables, sigmas, constants, gates, Q = gen[:5]
    cfg = prover.ProofConfig(fri_lde_factor=8, merkle_tree_cap_size=16, security_level=100)
    setup = prover.Setup(ctx, sigmas, constants, gates, Q, cfg, lookup=lk)
    ref = prover.prove(ctx, setup, variables, multiplicities=lk["multiplicities"] if lk else None)
    nat = ctx.native_setup(sigmas.contiguous(), constants.contiguous(), gates, Q, cfg, lookup=lk)
    assert np.array_equal(nat.get_cap(), setup.cap)
    timings = {}
    got = nat.prove(variables.contiguous(), lk["multiplicities"] if lk else None, timings=timings)
    assert json.dumps(got, sort_keys=True) == json.dumps(ref, sort_keys=True)
    assert OV.verify(setup.vk(), got)
    assert len(timings) == 6 and all(v >= 0 for v in timings.values())
    nat.close()


@pytest.mark.parametrize("log_n,lde,cap,lookup,hasher", [(10, 2, 16, False, "poseidon2"), (10, 2, 16, True, "blake2s"), (9, 2, 4, True, "poseidon2")])
def test_quotient_degree_above_fri_lde_factor(env, log_n, lde, cap, lookup, hasher):
    """fri_lde_factor < quotient degree - the production shape (the reference's own proof.json has fri_lde_factor 2 with
    quotient degree 8): columns are evaluated at max(L, Q) cosets, the oracles commit to the first L (prover.rs:178-196
    `used_lde_degree`, `subset_for_degree`).  Both drivers agree, the verifier accepts, a tampered opening is rejected, and
    the committed caps equal those of a plain factor-L evaluation (the committed subset IS the factor-L domain)."""
    bj, ctx, prover, synthetic = env
    gen = synthetic.generate(ctx, log_n, 60, seed=21, lookup=lookup)
    lk = gen[5] if lookup else None
    variables, sigmas, constants, gates, Q = gen[:5]
    assert Q > lde
    cfg = prover.ProofConfig(fri_lde_factor=lde, merkle_tree_cap_size=cap, security_level=100, hasher=hasher,
                             transcript="blake2s" if hasher == "blake2s" else "poseidon2")
    setup = prover.Setup(ctx, sigmas, constants, gates, Q, cfg, lookup=lk)
    m = lk["multiplicities"] if lk else None
    ref = prover.prove(ctx, setup, variables, multiplicities=m)
    nat = ctx.native_setup(sigmas.contiguous(), constants.contiguous(), gates, Q, cfg, lookup=lk)
    assert np.array_equal(nat.get_cap(), setup.cap)
    got = nat.prove(variables.contiguous(), m)
    nat.close()
    assert json.dumps(got, sort_keys=True) == json.dumps(ref, sort_keys=True)
    assert got["proof_config"]["fri_lde_factor"] == lde and len(got["values_at_z"]) > 0
    assert OV.verify(setup.vk(), got)
    bad = json.loads(json.dumps(got))
    bad["values_at_z"][3]["coeffs"][0] ^= 1
    with pytest.raises(AssertionError):
        OV.verify(setup.vk(), bad)
    # the witness cap is the cap of a tree over the factor-L LDE of the same columns
    cols = [variables] + ([m.reshape(1, -1)] if lk else [])
    import torch
    plain = ctx.transform_raw_storages_to_lde(torch.cat(cols, dim=0).contiguous(), lde)
    tree = ctx.merkle_tree_construct([plain[c].reshape(-1) for c in range(plain.shape[0])], cap, hasher=hasher)
    want = prover._digests(tree.get_cap(), hasher)
    assert got["witness_oracle_cap"] == want


@pytest.mark.parametrize("hasher,transcript", [("poseidon2", "poseidon2"), ("blake2s", "blake2s")])
def test_production_shaped_circuit(env, hasher, transcript):
    """The geometry of the reference's own vk.json / proof.json fixture (a recursion-layer circuit): 130 general-purpose columns
    with its 11 evaluators behind the 6-level selector tree (incl. the Poseidon2 flattened gate: ~9.6k relations, 118 terms),
    8 lookup sub-arguments of width 3, a boolean gate on a specialised column, 4 public inputs, quotient degree 8 with
    fri_lde_factor 2 and cap 32.  Both drivers give the same proof; the verifier - which evaluates the recorded programs over
    Fp2 itself - accepts it and rejects tampered openings; a witness that breaks the specialised boolean gate, or an
    fma row, is refused by the prover."""
    bj, ctx, prover, synthetic = env
    c = synthetic.generate_production_shaped(ctx, 11, seed=5)
    cfg = prover.ProofConfig(fri_lde_factor=2, merkle_tree_cap_size=32, security_level=100, hasher=hasher, transcript=transcript)
    assert c["variables"].shape[0] == 155 and c["constants"].shape[0] == 8 and len(c["gates"]) == 11
    assert sum(len(g["writes"]) * g["num_repetitions"] for g in c["gates"]) == 415      # the fixture's number of gate terms
    nat = ctx.native_setup(c["sigmas"], c["constants"], c["gates"], 8, cfg, lookup=c["lookup"], public_inputs=c["public_inputs"])
    m = c["lookup"]["multiplicities"]
    proof = nat.prove(c["variables"], m)
    vk = nat.vk()
    assert OV.verify(vk, proof)
    assert len(proof["public_inputs"]) == 4 and len(proof["values_at_z"]) == 155 + 8 + 155 + 1 + 19 + 1 + 9 + 4 + 8
    setup = prover.Setup(ctx, c["sigmas"], c["constants"], c["gates"], 8, cfg, lookup=c["lookup"], public_inputs=c["public_inputs"])
    ref = prover.prove(ctx, setup, c["variables"], multiplicities=m)
    assert json.dumps(ref, sort_keys=True) == json.dumps(proof, sort_keys=True)
    bad = json.loads(json.dumps(proof))
    bad["values_at_z"][40]["coeffs"][1] ^= 1          # a general-purpose variable read by the Poseidon2 gate
    with pytest.raises(AssertionError):
        OV.verify(vk, bad)
    for col, val in ((154, 2), (3, 12345)):           # the specialised boolean column; the output of the first fma repetition
        w = c["variables"].clone()
        rows = (c["constants"][1] == 1) & (c["constants"][3] == 1) if col == 3 else None     # fma rows: path bits 0 1 1 1 1
        row = int(rows.nonzero()[0]) if rows is not None else 9
        w[col, row] = val
        with pytest.raises(bj.BoojumError):
            nat.prove(w, m)
    nat.close()


@pytest.mark.parametrize("log_n,V,L,cap,lookup,pis,hasher,transcript", [
    (5, 20, 8, 16, False, (), "poseidon2", "poseidon2"), (6, 20, 4, 8, True, ((1, 3), (5, 3), (0, 9)), "poseidon2", "poseidon2"),
    (6, 20, 2, 8, True, (), "poseidon2", "poseidon2"), (7, 40, 8, 16, True, ((2, 100),), "poseidon2", "poseidon2"),
    (9, 60, 8, 16, True, ((0, 1), (59, 511)), "poseidon2", "poseidon2"),
    (6, 20, 8, 16, True, ((3, 5),), "blake2s", "blake2s"), (6, 20, 8, 16, True, ((3, 5),), "poseidon2", "poseidon"),
    (5, 20, 4, 8, True, (), "keccak256", "keccak256")])
def test_proof_equals_the_cpu_oracle_prover(env, log_n, V, L, cap, lookup, pis, hasher, transcript):
    """bj_prove returns, bit for bit, the proof of the oracle's own CPU prover (oracle/prover.py: prove_cpu_basic restated end
    to end - quotient point by point in Python integers, openings by Horner from the monomial forms, C restatement of NTT /
    Merkle / FRI) on a circuit generated on the CPU (oracle/circuits.py): transcript order, challenges, caps, openings, the FRI
    chain, query indices and every Merkle path agree - incl. lookups, public inputs and an LDE factor below the quotient degree."""
    bj, ctx, prover, synthetic = env
    from oracle import circuits, prover as OP
    c = circuits.sha_shaped(log_n, V, seed=100 + log_n, lookup=lookup)
    want, want_setup_cap = OP.prove(c["variables"], c["sigmas"], c["constants"], c["gates"], c["quotient_degree"], L, cap,
                                    lookup=c["lookup"], public_inputs=pis, hasher=hasher, transcript=transcript)
    gates = synthetic.sha_shaped_gates(V)
    assert [(g["name"], g["num_repetitions"], g["selector_path"]) for g in gates] == [tuple(g) for g in c["gates"]]
    lk = None
    if lookup:
        lk = dict(c["lookup"], tables=bj.to_device(c["lookup"]["tables"]), multiplicities=bj.to_device(c["lookup"]["multiplicities"]))
    cfg = prover.ProofConfig(fri_lde_factor=L, merkle_tree_cap_size=cap, security_level=100, hasher=hasher, transcript=transcript)
    nat = ctx.native_setup(bj.to_device(c["sigmas"]), bj.to_device(c["constants"]), gates, c["quotient_degree"], cfg, lookup=lk,
                           public_inputs=list(pis))
    assert np.array_equal(nat.get_cap(), want_setup_cap)
    got = nat.prove(bj.to_device(c["variables"]), lk["multiplicities"] if lk else None)
    nat.close()
    for key in want:                      # field by field first (a readable failure), then the whole document
        assert got[key] == want[key], key
    assert json.dumps(got, sort_keys=True) == json.dumps(want, sort_keys=True)


@pytest.mark.parametrize("log_n,cap,hasher", [(5, 4, "poseidon2"), (4, 4, "blake2s")])
def test_production_shaped_proof_equals_the_cpu_oracle_prover(env, log_n, cap, hasher):
    """the production shape at proof level: the geometry of the reference's vk.json - 11 gates as recorded programs (the Poseidon2
    flattened gate, the boolean gate on a specialised column, ...), 8 lookups of width 3, quotient degree 8 over fri_lde_factor 2,
    public inputs - generated on the CPU; bj_prove (device interpreter with all peephole passes) returns bit for bit the proof of
    the oracle's CPU prover, which evaluates the recorded programs relation by relation in Python integers."""
    bj, ctx, prover, synthetic = env
    from oracle import circuits, prover as OP
    from tests.test_oracle_prover_cpu import production_gates
    c = circuits.production_shaped(log_n, seed=40 + log_n)
    dicts, tuples = production_gates()
    pis = ((0, 3), (1, 3), (2, 7))
    want, want_setup_cap = OP.prove(c["variables"], c["sigmas"], c["constants"], tuples, 8, 2, cap, lookup=c["lookup"], public_inputs=pis,
                                    hasher=hasher)
    lk = dict(c["lookup"], tables=bj.to_device(c["lookup"]["tables"]), multiplicities=bj.to_device(c["lookup"]["multiplicities"]))
    cfg = prover.ProofConfig(fri_lde_factor=2, merkle_tree_cap_size=cap, security_level=100, hasher=hasher, transcript=hasher)
    nat = ctx.native_setup(bj.to_device(c["sigmas"]), bj.to_device(c["constants"]), dicts, 8, cfg, lookup=lk, public_inputs=list(pis))
    assert np.array_equal(nat.get_cap(), want_setup_cap)
    got = nat.prove(bj.to_device(c["variables"]), lk["multiplicities"])
    nat.close()
    for key in want:
        assert got[key] == want[key], key
    assert json.dumps(got, sort_keys=True) == json.dumps(want, sort_keys=True)


def test_recursive_mode_poseidon2_type_parameters(env):
    """H = GoldilocksPoseidon2Sponge, TR = GoldilocksPoisedonTranscript (Poseidon v1 sponge): the type parameters of
    run_sha256_prover_recursive_mode_poseidon2 (src/gadgets/sha256/mod.rs:286-293, BASELINE configs[4]).  Both drivers give the
    same proof, the oracle verifier (Python-int Poseidon v1) accepts it, and it differs from the Poseidon2-transcript proof."""
    bj, ctx, prover, synthetic = env
    variables, sigmas, constants, gates, Q, lk = synthetic.generate(ctx, 9, 60, seed=5, lookup=True)
    cfg = prover.ProofConfig(fri_lde_factor=8, merkle_tree_cap_size=16, security_level=100, hasher="poseidon2", transcript="poseidon")
    setup = prover.Setup(ctx, sigmas, constants, gates, Q, cfg, lookup=lk)
    ref = prover.prove(ctx, setup, variables, multiplicities=lk["multiplicities"])
    nat = ctx.native_setup(sigmas.contiguous(), constants.contiguous(), gates, Q, cfg, lookup=lk)
    got = nat.prove(variables.contiguous(), lk["multiplicities"])
    assert json.dumps(got, sort_keys=True) == json.dumps(ref, sort_keys=True)
    assert nat.vk() == setup.vk() and setup.vk()["transcript"] == "poseidon"
    assert OV.verify(setup.vk(), got)
    cfg2 = prover.ProofConfig(fri_lde_factor=8, merkle_tree_cap_size=16, security_level=100)
    other = prover.prove(ctx, prover.Setup(ctx, sigmas, constants, gates, Q, cfg2, lookup=lk), variables, multiplicities=lk["multiplicities"])
    assert other["witness_oracle_cap"] == got["witness_oracle_cap"] and other["values_at_z"] != got["values_at_z"]
    bad = dict(setup.vk(), transcript="poseidon2")
    with pytest.raises(AssertionError):
        OV.verify(bad, got)
    nat.close()


def test_native_cxx_prover_rejects_unsatisfied_witness(env):
    bj, ctx, prover, synthetic = env
    variables, sigmas, constants, gates, Q = synthetic.generate(ctx, 8, 20, seed=2)
    cfg = prover.ProofConfig(fri_lde_factor=8, merkle_tree_cap_size=16, security_level=100)
    nat = ctx.native_setup(sigmas.contiguous(), constants.contiguous(), gates, Q, cfg)
    bad = variables.clone()
    bad[3, 5] += 1
    with pytest.raises(bj.BoojumError):
        nat.prove(bad.contiguous())
    nat.prove(variables.contiguous())


@pytest.mark.parametrize("lookup", [False, True])
def test_public_inputs_python_and_native_drivers(env, lookup):
    """public inputs (places in the variable columns): committed to first, enforced by quotening at w^row
    (prover.rs:264-266, 1805-1821, 2010-2041); two of the three share a row, so two opening points."""
    bj, ctx, prover, synthetic = env
    gen = synthetic.generate(ctx, 9, 60, seed=21, lookup=lookup)
    lk = gen[5] if lookup else None
    variables, sigmas, constants, gates, Q = gen[:5]
    cfg = prover.ProofConfig(fri_lde_factor=8, merkle_tree_cap_size=16, security_level=100)
    places = [(3, 17), (40, 17), (7, 300)]
    setup = prover.Setup(ctx, sigmas, constants, gates, Q, cfg, lookup=lk, public_inputs=places)
    m = lk["multiplicities"] if lk else None
    proof = prover.prove(ctx, setup, variables, multiplicities=m)
    vk = setup.vk()
    assert proof["public_inputs"] == [int(bj.to_numpy(variables[c, r].reshape(1))[0]) for c, r in places]
    assert OV.verify(vk, proof)
    nat = ctx.native_setup(sigmas, constants, gates, Q, cfg, lookup=lk, public_inputs=places)
    assert json.dumps(nat.prove(variables, m), sort_keys=True) == json.dumps(proof, sort_keys=True)
    bad = copy.deepcopy(proof)
    bad["public_inputs"][1] = (bad["public_inputs"][1] + 1) % bj.P
    with pytest.raises(AssertionError):
        OV.verify(vk, bad)
    # sharded prover agrees as well
    res = None
    if not lookup:
        import threading
        from era_boojum_b200 import parallel
        shared, outs, errs = parallel.ThreadComm(2), [None, None], []

        def run(rank):
            try:
                c2 = bj.Context(0)
                c2.set_coset_shard(rank, 2, 8)
                s2 = prover.Setup(c2, sigmas, constants, gates, Q, cfg, comm=shared.rank_view(rank), public_inputs=places)
                outs[rank] = prover.prove(c2, s2, variables)
                c2.synchronize()
                c2.close()
            except BaseException as e:
                errs.append(e)
                shared._barrier.abort()

        ts = [threading.Thread(target=run, args=(r,)) for r in range(2)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        assert not errs, errs
        assert all(json.dumps(o, sort_keys=True) == json.dumps(proof, sort_keys=True) for o in outs)


@pytest.mark.parametrize("lookup", [False, True])
def test_gates_over_specialized_columns(env, lookup):
    """an FMA gate placed on specialised columns (no selector, shared constants, terms ahead of the general-purpose gates;
    prover.rs:653-801) next to the selector-tree gates: prove (Python driver and native C++ driver) -> verify."""
    bj, ctx, prover, synthetic = env
    gen = synthetic.generate(ctx, 9, 60, seed=31, lookup=lookup)
    lk = gen[5] if lookup else None
    variables, sigmas, constants, gates, Q = gen[:5]
    variables, sigmas, constants, gates = synthetic.add_specialized_fma(ctx, variables, sigmas, constants, gates, repetitions=3, seed=1)
    cfg = prover.ProofConfig(fri_lde_factor=8, merkle_tree_cap_size=16, security_level=100)
    m = lk["multiplicities"] if lk else None
    setup = prover.Setup(ctx, sigmas, constants, gates, Q, cfg, lookup=lk)
    proof = prover.prove(ctx, setup, variables, multiplicities=m)
    assert OV.verify(setup.vk(), proof)
    nat = ctx.native_setup(sigmas, constants, gates, Q, cfg, lookup=lk)
    assert json.dumps(nat.prove(variables, m), sort_keys=True) == json.dumps(proof, sort_keys=True)
    bad = variables.clone()
    bad[variables.shape[0] - 1, 9] += 1            # break the last specialised repetition on one row
    with pytest.raises(ValueError):
        prover.prove(ctx, setup, bad, multiplicities=m)


@pytest.mark.parametrize("lookup,world", [(False, 1), (True, 1), (False, 2)])
def test_blake2s_hasher_and_transcript_non_recursive_config(env, lookup, world):
    """H = Blake2s256, TR = Blake2sTranscript - the type parameters of sha256_bench_non_recursive
    (src/gadgets/sha256/mod.rs:527): Python driver, native C++ driver and the coset-sharded prover agree, the oracle verifier
    (hashlib Blake2s) accepts, a Poseidon2-configured verifier does not."""
    bj, ctx, prover, synthetic = env
    gen = synthetic.generate(ctx, 9, 60, seed=41, lookup=lookup)
    lk = gen[5] if lookup else None
    variables, sigmas, constants, gates, Q = gen[:5]
    cfg = prover.ProofConfig(fri_lde_factor=8, merkle_tree_cap_size=16, security_level=100, hasher="blake2s", transcript="blake2s")
    m = lk["multiplicities"] if lk else None
    setup = prover.Setup(ctx, sigmas, constants, gates, Q, cfg, lookup=lk, public_inputs=[(2, 5)])
    proof = prover.prove(ctx, setup, variables, multiplicities=m)
    vk = setup.vk()
    assert vk["hasher"] == "blake2s" and OV.verify(vk, proof)
    wrong = dict(vk, transcript="poseidon2")
    with pytest.raises(AssertionError):
        OV.verify(wrong, proof)
    nat = ctx.native_setup(sigmas, constants, gates, Q, cfg, lookup=lk, public_inputs=[(2, 5)])
    assert json.dumps(nat.prove(variables, m), sort_keys=True) == json.dumps(proof, sort_keys=True)
    if world > 1:
        import threading
        from era_boojum_b200 import parallel
        shared, outs, errs = parallel.ThreadComm(world), [None] * world, []

        def run(rank):
            try:
                c2 = bj.Context(0)
                c2.set_coset_shard(rank, world, 8)
                s2 = prover.Setup(c2, sigmas, constants, gates, Q, cfg, comm=shared.rank_view(rank), public_inputs=[(2, 5)])
                outs[rank] = prover.prove(c2, s2, variables)
                c2.synchronize()
                c2.close()
            except BaseException as e:
                errs.append(e)
                shared._barrier.abort()

        ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        assert not errs, errs
        assert all(json.dumps(o, sort_keys=True) == json.dumps(proof, sort_keys=True) for o in outs)


def test_pow_blake2s_kernel_and_proof_with_pow_bits(env):
    """PoWRunner for Blake2s256 (pow.rs:52-147) on the GPU, then a proof with pow_bits = 20 (ProofConfig default,
    prover.rs:70): 98/3 -> fewer queries, nonce checked by the oracle verifier, both drivers agree."""
    import hashlib
    bj, ctx, prover, synthetic = env
    for seed, bits in ((b"", 1), (bytes(range(40)), 12), (bytes(range(7)), 20), (bytes(range(52)), 9), (b"abc", 0)):
        nonce = ctx.pow_blake2s(seed, bits)
        first = int.from_bytes(hashlib.blake2s(seed + nonce.to_bytes(8, "little"), digest_size=32).digest()[:8], "little")
        assert first & ((1 << bits) - 1) == 0
        if bits <= 16:   # the serial search of the reference returns the smallest solution (pow.rs:60-73)
            for c in range(nonce):
                f = int.from_bytes(hashlib.blake2s(seed + c.to_bytes(8, "little"), digest_size=32).digest()[:8], "little")
                assert f & ((1 << bits) - 1) != 0
    variables, sigmas, constants, gates, Q = synthetic.generate(ctx, 9, 60, seed=51)
    cfg = prover.ProofConfig(fri_lde_factor=8, merkle_tree_cap_size=16, security_level=100, pow_bits=20)
    setup = prover.Setup(ctx, sigmas, constants, gates, Q, cfg)
    proof = prover.prove(ctx, setup, variables)
    assert proof["pow_challenge"] != 0 and len(proof["queries_per_fri_repetition"]) < 34
    assert OV.verify(setup.vk(), proof)
    nat = ctx.native_setup(sigmas, constants, gates, Q, cfg)
    assert json.dumps(nat.prove(variables), sort_keys=True) == json.dumps(proof, sort_keys=True)
    bad = copy.deepcopy(proof)
    bad["pow_challenge"] += 1
    with pytest.raises(AssertionError):
        OV.verify(setup.vk(), bad)


def test_pow_keccak256_kernel(env):
    """PoWRunner for Keccak256 (pow.rs:140-230) on the GPU against the pure-Python Keccak-256 of the oracle: the nonce solves
    the puzzle and, for the serial range of the reference (<= 16 bits), is the smallest one."""
    from oracle.keccak import keccak256
    bj, ctx, prover, synthetic = env
    for seed, bits in ((b"", 1), (bytes(range(40)), 10), (bytes(range(7)), 18), (bytes(range(120)), 6), (b"abc", 0)):
        nonce = ctx.pow_keccak256(seed, bits)
        first = int.from_bytes(keccak256(seed + nonce.to_bytes(8, "little"))[:8], "little")
        assert first & ((1 << bits) - 1) == 0
        if bits <= 10:
            for c in range(nonce):
                f = int.from_bytes(keccak256(seed + c.to_bytes(8, "little"))[:8], "little")
                assert f & ((1 << bits) - 1) != 0
    with pytest.raises(bj.BoojumError):
        ctx.pow_keccak256(bytes(121), 4)


def test_keccak256_hasher_and_transcript(env):
    """H = Keccak256, TR = Keccak256Transcript (the third TreeHasher / transcript pair of the reference): both drivers agree
    and the oracle verifier (pure-Python Keccak-256) accepts."""
    bj, ctx, prover, synthetic = env
    variables, sigmas, constants, gates, Q = synthetic.generate(ctx, 8, 20, seed=61)
    cfg = prover.ProofConfig(fri_lde_factor=8, merkle_tree_cap_size=16, security_level=100, hasher="keccak256", transcript="keccak256")
    setup = prover.Setup(ctx, sigmas, constants, gates, Q, cfg)
    proof = prover.prove(ctx, setup, variables)
    assert OV.verify(setup.vk(), proof)
    nat = ctx.native_setup(sigmas, constants, gates, Q, cfg)
    assert json.dumps(nat.prove(variables), sort_keys=True) == json.dumps(proof, sort_keys=True)
    bad = copy.deepcopy(proof)
    bad["queries_per_fri_repetition"][1]["stage_2_query"]["proof"][0][0] ^= 1
    with pytest.raises(AssertionError):
        OV.verify(setup.vk(), bad)
