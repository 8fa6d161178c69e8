"""Host driver of the five-round IOP over the B200 kernels, mirroring CSReferenceAssembly::prove_cpu_basic
(src/cs/implementations/prover.rs:153-2269) for circuits whose gates live on general-purpose columns (optional lookup
argument over specialised columns, optional public inputs; no gates over specialised columns).  Every heavy step is a C-ABI call into libboojum_b200.so; the transcript,
the FRI schedule and the proof assembly stay on the host, as in the reference.  The proof is returned in the reference's
serde shape (src/cs/implementations/proof.rs:57-143).

    setup = Setup.from_columns(ctx, sigmas, constants, gates, ...)      # one-off (setup.rs:1093-1255 role)
    proof = prove(ctx, setup, variables, config)                       # prove_cpu_basic role
"""
import time

import numpy as np

from . import BoojumError, Transcript, to_numpy
from .parallel import LocalComm, assemble_cap, local_leaf_index

P = 0xFFFFFFFF00000001


def e_mul(a, b):
    return ((a[0] * b[0] + 7 * a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def e_mul_base(a, b):
    return (a[0] * b % P, a[1] * b % P)


def _omega(log_n):
    w = 0x185629DCDA58878C
    for _ in range(log_n, 32):
        w = w * w % P
    return w


def _digests(arr, hasher):
    """serde form of TreeHasher::Output values held as rows of 4 little-endian u64: Poseidon2 -> [u64; 4]; Blake2s256 /
    Keccak256 -> [u8; 32] (src/cs/oracle/mod.rs:180, 245)."""
    a = np.ascontiguousarray(np.asarray(arr, dtype=np.uint64).reshape(-1, 4))
    if hasher == "poseidon2":
        return a.tolist()
    return a.astype("<u8").view(np.uint8).reshape(-1, 32).tolist()


def _ext_dict(v):
    return {"coeffs": [int(v[0]), int(v[1])], "_marker": None}


def _combine_ext(a, b):
    """value of f0 + u f1 at an Fp2 point from the values a = f0(at), b = f1(at) (u^2 = 7)."""
    return ((a[0] + 7 * b[1]) % P, (a[1] + b[0]) % P)


class ProofConfig:
    """ProofConfig (prover.rs:55-73)."""

    def __init__(self, fri_lde_factor=8, merkle_tree_cap_size=16, security_level=100, pow_bits=0, hasher="poseidon2",
                 transcript="poseidon2"):
        """hasher / transcript: the H and TR type parameters of prove_cpu_basic - "poseidon2" for both is the recursive-mode
        bench (GoldilocksPoseidon2Sponge + GoldilocksPoisedon2Transcript), "blake2s" for both is sha256_bench_non_recursive
        (Blake2s256 + Blake2sTranscript, src/gadgets/sha256/mod.rs:527)."""
        self.fri_lde_factor, self.merkle_tree_cap_size = fri_lde_factor, merkle_tree_cap_size
        self.security_level, self.pow_bits = security_level, pow_bits
        self.hasher, self.transcript = hasher, transcript

    def to_dict(self):
        return {"fri_lde_factor": self.fri_lde_factor, "merkle_tree_cap_size": self.merkle_tree_cap_size,
                "fri_folding_schedule": None, "security_level": self.security_level, "pow_bits": self.pow_bits}


class Setup:
    """SetupStorage + setup Merkle tree + the fixed parameters of the VerificationKey (setup.rs:1093-1255,
    verifier.rs:31-79): sigma and constant columns, their LDEs, the tree over [sigmas | constants]."""

    def __init__(self, ctx, sigmas, constants, gates, quotient_degree, config, lookup=None, comm=None, public_inputs=()):
        """lookup: None or dict(width, num_repetitions, variables_offset, table_id_column (index into constants),
        tables=[width + 1, n] tensor of lookup-table setup columns, multiplicities are part of the witness).
        comm: None (one GPU) or a communicator of era_boojum_b200.parallel - every rank then keeps the LDE cosets
        j = rank (mod world) only and ctx must carry the matching coset shard (Context.set_coset_shard)."""
        torch = ctx._torch
        self.comm = comm
        self.public_inputs = [(int(c), int(r)) for c, r in public_inputs]   # (column, row) places (CSReferenceAssembly::public_inputs)
        world = comm.world if comm else 1
        assert ctx.shard_world == world, "Context.set_coset_shard(rank, world, L) must match the communicator"
        assert world == 1 or config.merkle_tree_cap_size >= config.fri_lde_factor, "sharded proving needs cap_size >= LDE factor"
        self.gates = gates                      # [dict(name, program..., num_repetitions, selector_path, ...)]
        self.quotient_degree = quotient_degree
        self.config = config
        self.num_variables = sigmas.shape[0]
        self.num_constants = constants.shape[0]
        self.log_n = sigmas.shape[1].bit_length() - 1
        self.sigmas = sigmas                    # [V, n] natural order (needed by stage 2)
        self.constants_natural = constants      # [C, n] natural order (table-id column of the lookup argument)
        L = config.fri_lde_factor
        # columns are evaluated at D = max(L, quotient degree) cosets; the oracle commits to the first L of them (the reference's
        # used_lde_degree / subset_for_degree, prover.rs:178-196): in bit-reversed coset order they ARE the factor-L domain
        D = max(L, quotient_degree)
        self.committed_len = (sigmas.shape[1] * L) // world
        self.lookup = lookup
        parts = [sigmas, constants] + ([lookup["tables"]] if lookup else [])
        cols = torch.cat(parts, dim=0).contiguous()
        self.lde = ctx.transform_raw_storages_to_lde(cols, D)      # [V + C, D / world, n]
        self.tree = ctx.merkle_tree_construct([self.lde[c].reshape(-1)[:self.committed_len] for c in range(cols.shape[0])],
                                              config.merkle_tree_cap_size // world, hasher=config.hasher)
        self.cap = self.tree.get_cap()
        if comm:
            self.cap = assemble_cap(comm, self.cap, L, config.merkle_tree_cap_size)

    def sigma_lde(self, j):
        return self.lde[j].reshape(-1)

    def constant_lde(self, j):
        return self.lde[self.num_variables + j].reshape(-1)

    def table_lde(self, j):
        return self.lde[self.num_variables + self.num_constants + j].reshape(-1)

    def vk(self):
        return verification_key(self.log_n, self.num_variables, self.num_constants, self.gates, self.quotient_degree, self.config,
                                self.lookup, self.public_inputs, self.cap)


def verification_key(log_n, num_variables, num_constants, gates, quotient_degree, config, lookup, public_inputs, cap):
    """The fixed parameters a verifier needs (VerificationKey, verifier.rs:31-79) for circuits of this driver, plus the setup cap
    in serde form; shared by the Python Setup and the native bj_setup wrapper."""
    lk = None
    if lookup:
        lk = {k: lookup[k] for k in ("width", "num_repetitions", "variables_offset", "table_id_column")}
    return {"lookup": lk, "domain_size": 1 << log_n, "num_variables": num_variables, "num_constants": num_constants,
            "quotient_degree": quotient_degree, "fri_lde_factor": config.fri_lde_factor,
            "cap_size": config.merkle_tree_cap_size,
            # (name, repetitions, selector path, first variable column, first constant column[, recorded program]): the three
            # bench evaluators are known to the verifier by name; any other gate travels as its SSA program
            "gates": [(g["name"], g["num_repetitions"], list(g["selector_path"]), g.get("variables_initial_offset", 0),
                       g["constants_placement_offset"]) +
                      (() if g["name"] in ("fma", "reduction4", "constant_allocator") else
                       ({k: g.get(k, 0) for k in ("relations", "writes", "variables_offset", "witnesses_offset", "constants_offset")},))
                      for g in gates],
            "public_inputs_locations": [list(p) for p in public_inputs],
            "hasher": config.hasher, "transcript": config.transcript,
            "setup_merkle_tree_cap": _digests(cap, config.hasher)}


def _commit(ctx, comm, cols, L, cap, hasher="poseidon2"):
    """Merkle oracle over LDE columns: local tree (this rank's cosets) + the global cap."""
    world = comm.world if comm else 1
    tree = ctx.merkle_tree_construct(cols, cap // world, hasher=hasher)
    local_cap = tree.get_cap()
    return tree, (assemble_cap(comm, local_cap, L, cap) if comm else local_cap)


class _ShardedFri:
    """do_fri (src/cs/implementations/fri/mod.rs:49-357) over coset shards: folds and oracle subtrees are local (a fold of
    2^k neighbours never leaves a coset), every oracle cap is gathered so that all ranks draw the same challenges, the last
    codeword (a few hundred elements) is gathered and interpolated by every rank."""

    def __init__(self, ctx, comm, tr, c0, c1, schedule, L, cap, hasher="poseidon2"):
        torch = ctx._torch
        self.levels, self.caps, self.schedule = [], [], list(schedule)
        kappa = pow(7, P - 2, P)                                              # coset_inverse (fri/mod.rs:194)
        cur0, cur1 = c0, c1
        for k in schedule:
            tree = ctx.merkle_tree_construct([cur0, cur1], cap // comm.world, elems_per_leaf=1 << k, hasher=hasher)
            gcap = assemble_cap(comm, tree.get_cap(), L, cap)
            tr.witness_merkle_tree_cap(gcap)
            alpha = tr.get_multiple_challenges_fixed(2)
            self.levels.append((cur0, cur1, tree, k))
            self.caps.append(gcap)
            cur0, cur1, kappa = ctx.fri_fold(cur0, cur1, k, alpha, kappa)
        # final codeword: local [L / world][m] -> global [L][m] -> monomials (fri/mod.rs:312-334)
        m = cur0.numel() // (L // comm.world)
        parts = comm.all_gather_host(np.stack([to_numpy(cur0), to_numpy(cur1)]).reshape(2, L // comm.world, m))
        full = np.zeros((2, L, m), np.uint64)
        for r, part in enumerate(parts):
            for kk in range(L // comm.world):
                full[:, kk * comm.world + r] = part[:, kk]
        fin = torch.from_numpy(full.reshape(2, L * m).view(np.int64)).to(c0.device).contiguous()
        ctx.bitreverse_enumeration_inplace(fin)
        ctx.ifft_natural_to_natural(fin, pow(kappa, P - 2, P))
        mono = to_numpy(fin)
        if mono[:, m:].any():
            raise ValueError("FRI: folded codeword is not of low degree")
        self.mono = (mono[0, :m].copy(), mono[1, :m].copy())
        tr.witness_field_elements(self.mono[0].tolist())
        tr.witness_field_elements(self.mono[1].tolist())


def prove(ctx, setup, variables, timings=None, multiplicities=None):
    """variables: [V, n] int64 CUDA tensor (copy-permutation columns incl. the lookup sub-argument columns, natural row
    order); multiplicities: [n] tensor when the setup has a lookup argument.  Returns the proof dict.
    With setup.comm set (coset-sharded proving, one process per GPU) every rank passes the same full witness, works on
    its own LDE cosets and returns the same proof."""
    torch = ctx._torch
    cfg = setup.config
    comm = setup.comm
    rank, world = (comm.rank, comm.world) if comm else (0, 1)
    L, cap = cfg.fri_lde_factor, cfg.merkle_tree_cap_size
    log_L = L.bit_length() - 1
    V, C, Q = setup.num_variables, setup.num_constants, setup.quotient_degree
    log_n, log_q = setup.log_n, Q.bit_length() - 1
    n = 1 << log_n
    L_loc = L // world
    D = max(L, Q)                      # evaluation domain of the committed columns (the quotient needs Q cosets of each)
    log_D = D.bit_length() - 1
    ncm = n * L_loc                    # committed prefix of every LDE column
    cm = lambda cols: [c[:ncm] for c in cols]
    Q_loc = Q // world if Q >= world else (1 if rank < Q else 0)     # owned cosets among the first Q
    dev = variables.device
    tm = timings if timings is not None else {}

    def mark(name, t0):
        if timings is not None:
            torch.cuda.synchronize()
            tm[name] = tm.get(name, 0.0) + time.perf_counter() - t0

    flat = lambda t: t.reshape(-1)
    tr = Transcript(cfg.transcript)
    tr.witness_merkle_tree_cap(setup.cap)                                   # prover.rs:211
    public_values = [int(to_numpy(variables[c, r].reshape(1))[0]) % P for c, r in setup.public_inputs]
    for v in public_values:                                                 # prover.rs:264-266
        tr.witness_field_elements([v])
    # ---- round 1: witness commitment (prover.rs:313-353) ----
    t0 = time.perf_counter()
    lk = setup.lookup
    w_lde = ctx.transform_raw_storages_to_lde(variables, D)                  # [V, D_loc, n]
    w_cols = [flat(w_lde[c]) for c in range(V)]
    m_col = None
    if lk:
        m_lde = ctx.transform_raw_storages_to_lde(multiplicities.reshape(1, -1).contiguous(), D)
        m_col = flat(m_lde[0])
    w_tree, w_cap = _commit(ctx, comm, cm(w_cols + ([m_col] if lk else [])), L, cap, cfg.hasher)   # variables | witness (none) | multiplicities
    tr.witness_merkle_tree_cap(w_cap)
    mark("1_witness_lde_commit", t0)
    # ---- round 2: copy-permutation products (prover.rs:360-554); the trace-domain part is replicated on every rank ----
    t0 = time.perf_counter()
    beta = tr.get_multiple_challenges_fixed(2)
    gamma = tr.get_multiple_challenges_fixed(2)
    if lk:
        lookup_beta = tr.get_multiple_challenges_fixed(2)                    # prover.rs:402-406
        lookup_gamma = tr.get_multiple_challenges_fixed(2)
    z0, z1, partials = ctx.compute_partial_products_in_extension([variables[c] for c in range(V)],
                                                                 [setup.sigmas[c] for c in range(V)], beta, gamma, Q)
    lk_polys = []
    if lk:
        wdt, nsub, voff = lk["width"], lk["num_repetitions"], lk["variables_offset"]
        a_polys, b_poly = ctx.compute_lookup_poly_pairs_specialized(
            [variables[voff + i] for i in range(wdt * nsub)], wdt, setup.constants_natural[lk["table_id_column"]],
            [lk["tables"][j] for j in range(wdt + 1)], multiplicities, lookup_beta, lookup_gamma)
        lk_polys = [t for pr in a_polys for t in pr] + [b_poly[0], b_poly[1]]
    st2 = torch.stack([z0, z1] + [t for pr in partials for t in pr] + lk_polys).contiguous()
    s2_lde = ctx.transform_raw_storages_to_lde(st2, D)
    s2_cols = [flat(s2_lde[c]) for c in range(st2.shape[0])]
    s2_tree, s2_cap = _commit(ctx, comm, cm(s2_cols), L, cap, cfg.hasher)
    tr.witness_merkle_tree_cap(s2_cap)
    n_partial = len(partials)
    mark("2_stage2_products_lde_commit", t0)
    # ---- round 3: quotient (prover.rs:560-1495) ----
    t0 = time.perf_counter()
    alpha = tr.get_multiple_challenges_fixed(2)
    n_gate_terms = sum(len(g["writes"]) * g["num_repetitions"] for g in setup.gates)
    n_lk_terms = (lk["num_repetitions"] + 1) if lk else 0      # lookup terms come first (prover.rs:608-625)
    total_terms = n_lk_terms + n_gate_terms + 1 + 1 + n_partial
    powers = [(1, 0)]
    for _ in range(1, total_terms):
        powers.append(e_mul(powers[-1], alpha))
    npts = n * Q
    const_cols = [setup.constant_lde(j) for j in range(C)]
    a_off = 2 + 2 * n_partial
    qq = torch.zeros((2, Q, n), dtype=torch.int64, device=dev)          # quotient values on the first Q cosets (global)
    if Q_loc:
        # the quotient lives on the first Q cosets = the first Q_loc local slots of every LDE column
        q0 = torch.zeros(n * Q_loc, dtype=torch.int64, device=dev)
        q1 = torch.zeros(n * Q_loc, dtype=torch.int64, device=dev)
        if lk:
            a_ldes = [(s2_cols[a_off + 2 * i], s2_cols[a_off + 2 * i + 1]) for i in range(nsub)]
            b_lde = (s2_cols[a_off + 2 * nsub], s2_cols[a_off + 2 * nsub + 1])
            ctx.quotient_lookup_specialized([w_cols[voff + i] for i in range(wdt * nsub)], wdt, const_cols[lk["table_id_column"]],
                                            [setup.table_lde(j) for j in range(wdt + 1)], m_col, a_ldes, b_lde, lookup_beta, lookup_gamma,
                                            powers[:n_lk_terms], q0, q1)
        ctx.evaluate_gates_over_general_purpose_columns(setup.gates, w_cols, [], const_cols,
                                                        powers[n_lk_terms:n_lk_terms + n_gate_terms], q0, q1)
        part_ldes = [(s2_cols[2 + 2 * c], s2_cols[3 + 2 * c]) for c in range(n_partial)]
        ctx.quotient_copy_permutation(w_cols, [setup.sigma_lde(j) for j in range(V)], (s2_cols[0], s2_cols[1]), part_ldes, beta, gamma,
                                      powers[n_lk_terms + n_gate_terms:], log_n, log_D, log_q, Q, q0, q1)
        ctx.divide_by_vanishing(q0, q1, log_n, log_q)
        for k in range(Q_loc):                                              # local slot k = global coset k * world + rank
            qq[0, k * world + rank] = q0[k * n:(k + 1) * n]
            qq[1, k * world + rank] = q1[k * n:(k + 1) * n]
        del q0, q1
    if comm:
        comm.all_reduce_sum_(qq)          # the cosets recombine here: every rank needs all Q of them for the interpolation
    # flatten the cosets into natural order, interpolate once at size n*Q on coset 7, split into Q chunks (prover.rs:1399-1467)
    qq = qq.reshape(2, npts)
    ctx.bitreverse_enumeration_inplace(qq)
    ctx.ifft_natural_to_natural(qq, 7)
    # the reference's satisfiability guard: the top coefficient of the interpolant must vanish (prover.rs:1425-1438)
    top = to_numpy(qq[:, npts - 1])
    if int(top[0]) != 0 or int(top[1]) != 0:
        raise ValueError("unsatisfied: quotient is not a polynomial of degree < n * quotient_degree")
    chunks = torch.stack([qq[k][j * n:(j + 1) * n] for j in range(Q) for k in (0, 1)]).contiguous()   # c0,c1 of chunk 0, ...
    qt_lde = ctx.transform_raw_storages_to_lde(chunks, L, from_monomials=True)
    qt_cols = [flat(qt_lde[c]) for c in range(2 * Q)]
    qt_tree, qt_cap = _commit(ctx, comm, qt_cols, L, cap, cfg.hasher)
    tr.witness_merkle_tree_cap(qt_cap)
    mark("3_quotient", t0)
    # ---- round 4: openings (prover.rs:1501-1802) ----
    t0 = time.perf_counter()
    z = tr.get_multiple_challenges_fixed(2)
    w_n = _omega(log_n)
    z_omega = e_mul_base(z, w_n)
    # opening order (prover.rs:1549-1683): variables, witness, constants, sigmas, z, partial products, multiplicities,
    # lookup A, lookup B, lookup tables, quotient chunks.  `sources` keeps (c0, c1-or-None) per opened polynomial.
    sources = [(c, None) for c in w_cols] + [(c, None) for c in const_cols] + [(setup.sigma_lde(j), None) for j in range(V)]
    sources += [(s2_cols[2 * i], s2_cols[2 * i + 1]) for i in range(1 + n_partial)]
    if lk:
        sources += [(m_col, None)]
        sources += [(s2_cols[a_off + 2 * i], s2_cols[a_off + 2 * i + 1]) for i in range(nsub + 1)]
        sources += [(setup.table_lde(j), None) for j in range(wdt + 1)]
    sources += [(qt_cols[2 * i], qt_cols[2 * i + 1]) for i in range(Q)]

    def open_at(srcs, at):
        flat_cols, spans = [], []
        for c0, c1 in srcs:
            spans.append((len(flat_cols), c1 is not None))
            flat_cols += [c0] + ([c1] if c1 is not None else [])
        ev = ctx.barycentric_evaluate(flat_cols, log_n, at)
        return [(_combine_ext(ev[i], ev[i + 1]) if is_ext else ev[i]) for i, is_ext in spans]

    zero_sources = [(s2_cols[a_off + 2 * i], s2_cols[a_off + 2 * i + 1]) for i in range(nsub + 1)] if lk else []
    opened = None
    if rank == 0:                          # barycentric evaluation reads coset 0, which rank 0 owns
        opened = (open_at(sources, z), open_at([(s2_cols[0], s2_cols[1])], z_omega), open_at(zero_sources, (0, 0)) if lk else [])
    if comm:
        opened = comm.broadcast_host(opened, 0)
    values_at_z, values_at_z_omega, values_at_0 = opened
    for v in values_at_z + values_at_z_omega + values_at_0:
        tr.witness_field_elements(v)
    mark("4_openings", t0)
    # ---- round 5: DEEP + FRI (prover.rs:1828-2102) ----
    t0 = time.perf_counter()
    # public inputs are enforced by quotening the variable columns at w^row, grouped by point (prover.rs:1805-1821, 2010-2041)
    pi_groups = []
    for (col, row), val in zip(setup.public_inputs, public_values):
        at = pow(w_n, row, P)
        for g in pi_groups:
            if g[0] == at:
                g[1].append((col, val))
                break
        else:
            pi_groups.append((at, [(col, val)]))
    c = tr.get_multiple_challenges_fixed(2)
    n_ch = len(values_at_z) + 1 + len(values_at_0) + len(public_values)
    ch = [(1, 0), c]
    for _ in range(2, n_ch):
        ch.append(e_mul(ch[-1], c))
    deep0 = torch.zeros(n * L_loc, dtype=torch.int64, device=dev)
    deep1 = torch.zeros(n * L_loc, dtype=torch.int64, device=dev)
    ctx.quotening_operation_in_extension(deep0, deep1, sources, values_at_z, z, ch[:len(sources)])
    ctx.quotening_operation_in_extension(deep0, deep1, [(s2_cols[0], s2_cols[1])], values_at_z_omega, z_omega,
                                         ch[len(sources):len(sources) + 1])
    off_ch = len(sources) + 1
    if lk:
        ctx.quotening_operation_in_extension(deep0, deep1, zero_sources, values_at_0, (0, 0), ch[off_ch:off_ch + len(values_at_0)])
        off_ch += len(values_at_0)
    for at, members in pi_groups:
        ctx.quotening_operation_in_extension(deep0, deep1, [(w_cols[col], None) for col, _ in members], [(val, 0) for _, val in members],
                                             (at, 0), ch[off_ch:off_ch + len(members)])
        off_ch += len(members)
    import ctypes
    from .native import lib
    np_, nq, sl, fd = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32()
    sched = (ctypes.c_uint32 * 32)()
    st = lib.bj_compute_fri_schedule(cfg.security_level, cap, cfg.pow_bits, log_L, log_n, ctypes.byref(np_), ctypes.byref(nq),
                                     sched, ctypes.byref(sl), ctypes.byref(fd))
    if st != 0:
        raise BoojumError(st, "bj_compute_fri_schedule: " + lib.bj_status_string(st).decode())
    schedule = list(sched[: sl.value])
    if not schedule:
        raise BoojumError(-1, "degenerate FRI instance: log_n + log_lde too small for cap size %d (no folding step)" % cap)
    if comm:
        fri = _ShardedFri(ctx, comm, tr, deep0, deep1, schedule, L, cap, cfg.hasher)
        mono0, mono1 = fri.mono
        fri_caps = fri.caps
    else:
        fri = ctx.do_fri(tr, deep0, deep1, schedule, L, cap, hasher=cfg.hasher)
        mono0, mono1 = fri.monomial_forms()
        fri_caps = [fri.get_cap(i) for i in range(fri.num_oracles())]
    # proof of work (prover.rs:2109-2132; POW = Blake2s256): seeded by 5 challenges, the nonce goes back into the transcript
    pow_challenge = 0
    if np_.value:
        seed = b"".join(int(tr.get_challenge()).to_bytes(8, "little") for _ in range(5))
        pow_challenge = ctx.pow_blake2s(seed, np_.value)      # every rank of a sharded prover finds the same nonce
        tr.witness_field_elements([pow_challenge & 0xFFFFFFFF, pow_challenge >> 32])
    mark("5_deep_fri", t0)
    # ---- queries (prover.rs:2161-2266): a query is answered by the rank that owns the coset of its index ----
    t0 = time.perf_counter()
    max_bits = log_n + log_L
    idxs = [tr.get_index_bits(max_bits, max_bits) for _ in range(nq.value)]
    setup_cols = [setup.lde[c].reshape(-1) for c in range(setup.lde.shape[0])]
    oracles = [("witness_query", cm(w_cols + ([m_col] if lk else [])), w_tree), ("stage_2_query", cm(s2_cols), s2_tree),
               ("quotient_query", qt_cols, qt_tree), ("setup_query", cm(setup_cols), setup.tree)]
    mine = [(qi, idx) for qi, idx in enumerate(idxs) if local_leaf_index(idx, log_n, world)[0] == rank]
    loc = [local_leaf_index(idx, log_n, world)[1] for _, idx in mine]
    rows = {name: (ctx.query_leaf_elements(cols, loc), ctx.merkle_paths(tree, loc)) for name, cols, tree in oracles} if mine else {}
    answered = {}
    for pos, (qi, idx) in enumerate(mine):
        q = {name: {"leaf_elements": rows[name][0][pos].tolist(), "proof": _digests(rows[name][1][pos], cfg.hasher)} for name, _, _ in oracles}
        fqs, sub, log_len = [], idx, log_n
        for lvl, k in enumerate(schedule):
            if comm:
                c0_l, c1_l, tree_l, _ = fri.levels[lvl]
                leaf = local_leaf_index(sub, log_len, world)[1] >> k
                le = ctx.query_leaf_elements([c0_l, c1_l], [leaf], elems_per_leaf=1 << k)[0]
                path = ctx.merkle_paths(tree_l, [leaf])[0]
            else:
                le, path = fri.query(lvl, sub >> k, k)
            fqs.append({"leaf_elements": le.tolist(), "proof": _digests(path, cfg.hasher)})
            sub >>= k
            log_len -= k
        q["fri_queries"] = fqs
        answered[qi] = q
    if comm:
        merged = {}
        for part in comm.all_gather_host(answered):
            merged.update(part)
        answered = merged
    queries = [answered[qi] for qi in range(len(idxs))]
    mark("6_queries", t0)
    return {
        "proof_config": cfg.to_dict(), "public_inputs": public_values,
        "witness_oracle_cap": _digests(w_cap, cfg.hasher), "stage_2_oracle_cap": _digests(s2_cap, cfg.hasher),
        "quotient_oracle_cap": _digests(qt_cap, cfg.hasher),
        "final_fri_monomials": [mono0.tolist(), mono1.tolist()],
        "values_at_z": [_ext_dict(v) for v in values_at_z], "values_at_z_omega": [_ext_dict(v) for v in values_at_z_omega],
        "values_at_0": [_ext_dict(v) for v in values_at_0],
        "fri_base_oracle_cap": _digests(fri_caps[0], cfg.hasher),
        "fri_intermediate_oracles_caps": [_digests(c_, cfg.hasher) for c_ in fri_caps[1:]],
        "queries_per_fri_repetition": queries, "pow_challenge": pow_challenge, "_marker": None,
    }
