"""ORACLE (test infrastructure): restatement of the reference verifier for circuits built from the bench's three gates
over general-purpose columns, optionally with the bench's lookup argument over specialised columns (table id in a constant
column), optional public inputs (the acceptance oracle of the prove -> verify tests).

Follows Verifier::verify (src/cs/implementations/verifier.rs:888-2510):
  transcript order :898-1068, alpha-power split :978-1023, quotient identity at z :1144-1828 (gates over general purpose
  columns :1646-1700, z(1)=1 :1708-1722, copy-permutation relations :1724-1768, t_from_chunks :1772-1790; lookup
  sumcheck at 0 and relations at z :1238-1560),
  DEEP regrouping + FRI chain :1817-2510 (shared with oracle/replay.py, which proof.json pins).
Gate terms at z use the same formulas as oracle/gates.py, lifted to Fp2.
"""
import numpy as np

from . import oracle as O
from . import replay as R
from .stage2 import non_residues_for_copy_permutation

P = O.P


def _ext(d):
    return (d["coeffs"][0], d["coeffs"][1])


def e_pow(a, e):
    r = (1, 0)
    while e:
        if e & 1:
            r = R.e_mul(r, a)
        a = R.e_mul(a, a)
        e >>= 1
    return r


# gate terms at an Fp2 point: v = opened variable values, c = opened constants (starting at the gate's placement)
def _fma(v, c):
    return [R.e_sub(R.e_add(R.e_mul(R.e_mul(c[0], v[0]), v[1]), R.e_mul(c[1], v[2])), v[3])]


def _reduction4(v, c):
    acc = (0, 0)
    for i in range(4):
        acc = R.e_add(acc, R.e_mul(c[i], v[i]))
    return [R.e_sub(acc, v[4])]


def _constant_allocator(v, c):
    return [R.e_sub(v[0], c[0])]


# Index / Relation numbering of the recorded programs (gpu_synthesizer/mod.rs:115-133; include/boojum_b200.h BJ_IDX_* / BJ_REL_*)
_IDX_VARIABLE, _IDX_WITNESS, _IDX_CONSTANT_POLY, _IDX_TEMPORARY, _IDX_CONSTANT_VALUE, _IDX_CONSTANT_POLY_SHARED = range(6)
_REL_ADD, _REL_DOUBLE, _REL_SUB, _REL_NEGATE, _REL_MUL, _REL_SQUARE, _REL_INVERSE = range(7)


def _program_terms(prog, var_v, const_v, var_base, const_base):
    """One repetition of a recorded SSA program (GPUDataCapture) over Fp2 values at z; returns the pushed terms."""
    tmp = {}

    def fetch(o):
        kind, val = o
        if kind == _IDX_VARIABLE:
            return var_v[var_base + val]
        if kind == _IDX_CONSTANT_POLY:
            return const_v[const_base[1] + val]
        if kind == _IDX_CONSTANT_POLY_SHARED:
            return const_v[const_base[0] + val]
        if kind == _IDX_TEMPORARY:
            return tmp[val]
        if kind == _IDX_CONSTANT_VALUE:
            return (int(val) % R.P, 0)
        raise AssertionError("witness columns are not part of this driver's circuits")

    for op, dst, a, b in prog["relations"]:
        x = fetch(tuple(a))
        if op == _REL_ADD:
            r = R.e_add(x, fetch(tuple(b)))
        elif op == _REL_DOUBLE:
            r = R.e_add(x, x)
        elif op == _REL_SUB:
            r = R.e_sub(x, fetch(tuple(b)))
        elif op == _REL_NEGATE:
            r = R.e_sub((0, 0), x)
        elif op == _REL_MUL:
            r = R.e_mul(x, fetch(tuple(b)))
        elif op == _REL_SQUARE:
            r = R.e_mul(x, x)
        elif op == _REL_INVERSE:
            r = R.e_inv(x)
        else:
            raise AssertionError("unknown relation %r" % (op,))
        assert dst not in tmp, "programs are SSA"
        tmp[dst] = r
    return [fetch(tuple(w)) for w in prog["writes"]]


GATES = {"fma": (_fma, 4, (4, 0)), "reduction4": (_reduction4, 5, (5, 0)), "constant_allocator": (_constant_allocator, 1, (1, 1))}


def verify(vk, proof):
    """vk: dict(domain_size, num_variables, num_constants, quotient_degree, gates=[(name, reps, path)], fri_lde_factor,
    cap_size, setup_merkle_tree_cap).  proof: dict in the reference's serde shape.  Returns True or raises AssertionError."""
    vk, proof = R.normalize_digests(vk), R.normalize_digests(proof)   # Blake2s / Keccak digests arrive as [u8; 32]
    n = vk["domain_size"]
    log_n = n.bit_length() - 1
    L = proof["proof_config"]["fri_lde_factor"]
    log_L = L.bit_length() - 1
    cap_size = proof["proof_config"]["merkle_tree_cap_size"]
    assert L == vk["fri_lde_factor"] and cap_size == vk["cap_size"]
    V, C, Q = vk["num_variables"], vk["num_constants"], vk["quotient_degree"]
    gates = vk["gates"]
    n_partial = 0 if V <= Q else (V + Q - 1) // Q - 1
    lk = vk.get("lookup")
    nsub, wdt = (lk["num_repetitions"], lk["width"]) if lk else (0, 0)
    n_lk_terms = nsub + 1 if lk else 0
    n_mult = 1 if lk else 0
    n_tables = wdt + 1 if lk else 0

    hasher = vk.get("hasher", "poseidon2")                # tree hasher and transcript of the proof system instance
    tr = {"poseidon2": R.Poseidon2Transcript, "blake2s": R.Blake2sTranscript, "keccak256": R.Keccak256Transcript,
          "poseidon": R.PoseidonTranscript}[vk.get("transcript", "poseidon2")]()
    leaf_fn, path_ok = R.hasher_functions(hasher)
    tr.witness_merkle_tree_cap(vk["setup_merkle_tree_cap"])
    pi_locations = vk.get("public_inputs_locations", [])
    assert len(proof["public_inputs"]) == len(pi_locations)
    for v in proof["public_inputs"]:                      # verifier.rs: public inputs enter right after the setup cap
        tr.witness_field_elements([v])
    tr.witness_merkle_tree_cap(proof["witness_oracle_cap"])
    beta = tr.get_ext_challenge()
    gamma = tr.get_ext_challenge()
    if lk:
        lookup_beta = tr.get_ext_challenge()
        lookup_gamma = tr.get_ext_challenge()
    tr.witness_merkle_tree_cap(proof["stage_2_oracle_cap"])
    alpha = tr.get_ext_challenge()
    # one term per repetition for the three bench gates; a gate that carries its recorded program pushes len(writes) per repetition
    n_gate_terms = sum(g[1] * (len(g[5]["writes"]) if len(g) > 5 else 1) for g in gates)
    total_terms = n_lk_terms + n_gate_terms + 1 + 1 + n_partial
    powers = [(1, 0)]
    for _ in range(1, total_terms):
        powers.append(R.e_mul(powers[-1], alpha))
    lk_ch = powers[:n_lk_terms]
    gp_ch, rest_ch = powers[n_lk_terms:n_lk_terms + n_gate_terms], powers[n_lk_terms + n_gate_terms:]
    tr.witness_merkle_tree_cap(proof["quotient_oracle_cap"])
    z = tr.get_ext_challenge()
    vals_z = [_ext(v) for v in proof["values_at_z"]]
    vals_zw = [_ext(v) for v in proof["values_at_z_omega"]]
    vals_0 = [_ext(v) for v in proof["values_at_0"]]
    for v in vals_z + vals_zw + vals_0:
        tr.witness_field_elements(v)
    assert len(vals_z) == V + C + V + 1 + n_partial + n_mult + n_lk_terms + n_tables + Q and len(vals_zw) == 1
    assert len(vals_0) == n_lk_terms

    # ---- quotient identity at z ----
    it = iter(vals_z)
    var_v = [next(it) for _ in range(V)]
    const_v = [next(it) for _ in range(C)]
    sigma_v = [next(it) for _ in range(V)]
    z_at_z = next(it)
    partial_v = [next(it) for _ in range(n_partial)]
    mult_v = [next(it) for _ in range(n_mult)]
    a_v = [next(it) for _ in range(nsub)]
    b_v = [next(it) for _ in range(n_mult)]
    table_v = [next(it) for _ in range(n_tables)]
    quot_v = [next(it) for _ in range(Q)]
    z_at_zw = vals_zw[0]
    t_acc = (0, 0)
    if lk:
        # log-derivative sumcheck from the openings at 0 (verifier.rs:1238-1257), then the relations at z (:1258-1560)
        sa, sb = (0, 0), (0, 0)
        for v in vals_0[:nsub]:
            sa = R.e_add(sa, v)
        for v in vals_0[nsub:]:
            sb = R.e_add(sb, v)
        assert sa == sb, "Lookup sumcheck is invalid"
        gp = [(1, 0)]
        for _ in range(wdt):
            gp.append(R.e_mul(gp[-1], lookup_gamma))
        voff, tid = lk["variables_offset"], lk["table_id_column"]
        for i in range(nsub):
            d = lookup_beta
            for j in range(wdt):
                d = R.e_add(d, R.e_mul(gp[j], var_v[voff + i * wdt + j]))
            d = R.e_add(d, R.e_mul(gp[wdt], const_v[tid]))
            t_acc = R.e_add(t_acc, R.e_mul(R.e_sub(R.e_mul(a_v[i], d), (1, 0)), lk_ch[i]))
        d = lookup_beta
        for j in range(wdt + 1):
            d = R.e_add(d, R.e_mul(gp[j], table_v[j]))
        t_acc = R.e_add(t_acc, R.e_mul(R.e_sub(R.e_mul(b_v[0], d), mult_v[0]), lk_ch[nsub]))
    k = 0
    for g in gates:
        # (name, repetitions, selector path[, first variable column, first constant column]); the optional pair places a
        # gate on specialised columns (prover.rs:653-801): no selector, terms listed before the general-purpose gates
        name, reps, path = g[0], g[1], g[2]
        var0 = g[3] if len(g) > 3 else 0
        const0 = g[4] if len(g) > 4 else len(path)
        sel = (1, 0)
        for i, bit in enumerate(path):
            sel = R.e_mul(sel, const_v[i] if bit else R.e_sub((1, 0), const_v[i]))
        acc = (0, 0)
        for rep in range(reps):
            if len(g) > 5:      # recorded program: PerChunkOffset per repetition, row-shared constants at the gate's first column
                prog = g[5]
                terms = _program_terms(prog, var_v, const_v, var0 + rep * prog["variables_offset"],
                                       (const0, const0 + rep * prog["constants_offset"]))
            else:
                fn, width, (voff, coff) = GATES[name]
                terms = fn(var_v[var0 + rep * voff: var0 + rep * voff + width], const_v[const0 + rep * coff:])
            for term in terms:
                acc = R.e_add(acc, R.e_mul(term, gp_ch[k]))
                k += 1
        t_acc = R.e_add(t_acc, R.e_mul(acc, sel))
    assert k == n_gate_terms
    z_n = e_pow(z, n)
    vanishing = R.e_sub(z_n, (1, 0))
    ch_it = iter(rest_ch)
    l1 = R.e_mul(vanishing, R.e_inv(R.e_sub(z, (1, 0))))
    t_acc = R.e_add(t_acc, R.e_mul(R.e_mul(R.e_sub(z_at_z, (1, 0)), l1), next(ch_it)))
    ks = non_residues_for_copy_permutation(n, V)
    lhs_l = partial_v + [z_at_zw]
    rhs_l = [z_at_z] + partial_v
    for c, (lhs, rhs) in enumerate(zip(lhs_l, rhs_l)):
        a = next(ch_it)
        for j in range(c * Q, min((c + 1) * Q, V)):
            lhs = R.e_mul(lhs, R.e_add(R.e_add(R.e_mul(sigma_v[j], beta), var_v[j]), gamma))
            rhs = R.e_mul(rhs, R.e_add(R.e_add(R.e_mul(R.e_mul_base(z, ks[j]), beta), var_v[j]), gamma))
        t_acc = R.e_add(t_acc, R.e_mul(R.e_sub(lhs, rhs), a))
    assert next(ch_it, None) is None
    t_chunks, pw = (0, 0), (1, 0)
    for q in quot_v:
        t_chunks = R.e_add(t_chunks, R.e_mul(q, pw))
        pw = R.e_mul(pw, z_n)
    assert t_acc == R.e_mul(t_chunks, vanishing), "Invalid quotient at Z"

    # ---- DEEP + FRI ----
    # public inputs grouped by opening point w^row, in order of first appearance (verifier.rs:1074-1108)
    w_n = R.omega(log_n)
    pi_groups = []
    for (col, row), val in zip(pi_locations, proof["public_inputs"]):
        at = pow(w_n, row, R.P)
        for g in pi_groups:
            if g[0] == at:
                g[1].append((col, val))
                break
        else:
            pi_groups.append((at, [(col, val)]))
    c = tr.get_ext_challenge()
    ch = R.ext_powers(c, len(vals_z) + len(vals_zw) + len(vals_0) + sum(len(g[1]) for g in pi_groups))
    new_pow, num_queries, schedule, final_degree = R.compute_fri_schedule(
        proof["proof_config"]["security_level"], cap_size, proof["proof_config"]["pow_bits"], log_L, log_n)
    assert num_queries == len(proof["queries_per_fri_repetition"])
    fri_caps = [proof["fri_base_oracle_cap"]] + list(proof["fri_intermediate_oracles_caps"])
    assert len(fri_caps) == len(schedule)
    fri_ch = []
    for cap in fri_caps:
        tr.witness_merkle_tree_cap(cap)
        fri_ch.append(tr.get_ext_challenge())
    mono = proof["final_fri_monomials"]
    assert len(mono[0]) == final_degree == len(mono[1])
    tr.witness_field_elements(mono[0])
    tr.witness_field_elements(mono[1])
    if new_pow:
        # PoWRunner for Blake2s256 (pow.rs:135-146, verifier.rs: the same 5 challenges seed the check)
        import hashlib
        seed = b"".join(int(tr.get_challenge()).to_bytes(8, "little") for _ in range(5))
        nonce = int(proof["pow_challenge"])
        first = int.from_bytes(hashlib.blake2s(seed + nonce.to_bytes(8, "little"), digest_size=32).digest()[:8], "little")
        assert first & ((1 << new_pow) - 1) == 0, "proof of work is invalid"
        tr.witness_field_elements([nonce & 0xFFFFFFFF, nonce >> 32])
    else:
        assert proof["pow_challenge"] == 0
    max_bits = log_n + log_L
    bools = R.BoolsBuffer(max_bits)
    w_n = R.omega(log_n)
    z_omega = R.e_mul_base(z, w_n)
    depth = max_bits - (cap_size.bit_length() - 1)
    for q in proof["queries_per_fri_repetition"]:
        bits = bools.get_bits(tr, max_bits)
        idx = sum(b << i for i, b in enumerate(bits))
        for name, cap in (("witness_query", proof["witness_oracle_cap"]), ("stage_2_query", proof["stage_2_oracle_cap"]),
                          ("quotient_query", proof["quotient_oracle_cap"]), ("setup_query", vk["setup_merkle_tree_cap"])):
            leaf = leaf_fn(q[name]["leaf_elements"])
            path = np.array(q[name]["proof"], dtype=np.uint64).reshape(-1, 4)
            assert path.shape[0] == depth
            assert path_ok(leaf, path, cap, idx), (name, idx)
        wq, sq = q["witness_query"]["leaf_elements"], q["stage_2_query"]["leaf_elements"]
        qq, uq = q["quotient_query"]["leaf_elements"], q["setup_query"]["leaf_elements"]
        assert len(wq) == V + n_mult and len(sq) == 2 * (1 + n_partial + n_lk_terms) and len(qq) == 2 * Q
        assert len(uq) == V + C + n_tables
        base = lambda els: [(e, 0) for e in els]
        ext = lambda els: [(els[i], els[i + 1]) for i in range(0, len(els), 2)]
        off_a = 2 + 2 * n_partial
        src = base(wq[:V]) + base(uq[V:V + C]) + base(uq[:V]) + ext(sq[0:2]) + ext(sq[2:off_a]) + base(wq[V:]) + ext(sq[off_a:]) \
            + base(uq[V + C:]) + ext(qq)
        x = 1
        for b, pw in zip(bits, [R.omega(i) for i in range(1, max_bits + 1)]):
            if b:
                x = R.fmul(x, pw)
        x_q = R.fmul(x, 7)
        acc = O.deep_point((0, 0), src, vals_z, ch[:len(src)], x_q, z)
        acc = O.deep_point(acc, ext(sq[0:2]), vals_zw, ch[len(src):len(src) + 1], x_q, z_omega)
        off_ch = len(src) + 1
        if lk:
            acc = O.deep_point(acc, ext(sq[off_a:]), vals_0, ch[off_ch:off_ch + len(vals_0)], x_q, (0, 0))
            off_ch += len(vals_0)
        for at, members in pi_groups:                     # (w_col(x) - value) / (x - w^row)
            acc = O.deep_point(acc, base([wq[col] for col, _ in members]), [(val, 0) for _, val in members],
                               ch[off_ch:off_ch + len(members)], x_q, (at, 0))
            off_ch += len(members)
        fqs = [(fq["leaf_elements"], fq["proof"]) for fq in q["fri_queries"]]
        R.verify_fri_query(idx, log_n, log_L, schedule, cap_size, fri_caps, fri_ch, (mono[0], mono[1]), fqs, start_value=acc, hasher=hasher)
    return True
