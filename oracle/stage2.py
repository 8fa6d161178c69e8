"""ORACLE (test infrastructure): stage-2 copy-permutation polynomials and the quotient-side relations, restated with
Python ints (small domains only).

  make_non_residues / non_residues_for_copy_permutation   src/cs/implementations/utils.rs:636-688, copy_permutation.rs:512-523
  pointwise_rational_in_extension                          src/cs/implementations/copy_permutation.rs:114-248
  shifted_grand_product_in_extension                       :425-510
  compute_partial_products_in_extension                    :649-766
"""
from .replay import P, e_add, e_inv, e_mul, e_mul_base, omega


def non_residues_for_copy_permutation(domain_size, num_columns):
    out, seen, cur = [1], [], 1
    while len(out) < num_columns:
        cur += 1
        if pow(cur, (P - 1) // 2, P) != P - 1:
            continue
        t = pow(cur, domain_size, P)
        if t == 1 or t in seen:
            continue
        seen.append(t)
        out.append(cur)
    return out


def partial_products(variables, sigmas, beta, gamma, max_degree):
    """variables / sigmas: lists of columns (lists of ints, natural order).  Returns (z, partials) with Fp2 tuples."""
    n_cols, n = len(variables), len(variables[0])
    log_n = n.bit_length() - 1
    w_n = omega(log_n)
    ks = non_residues_for_copy_permutation(n, n_cols)
    chunks = [range(s, min(s + max_degree, n_cols)) for s in range(0, n_cols, max_degree)]
    ratios = []
    for ch in chunks:
        col = []
        for i in range(n):
            x = pow(w_n, i, P)
            num, den = (1, 0), (1, 0)
            for j in ch:
                w = variables[j][i] % P
                a = e_add(e_add(e_mul_base(beta, ks[j] * x % P), (w, 0)), gamma)
                b = e_add(e_add(e_mul_base(beta, sigmas[j][i] % P), (w, 0)), gamma)
                num, den = e_mul(num, a), e_mul(den, b)
            col.append(e_mul(num, e_inv(den)))
        ratios.append(col)
    z, run = [], (1, 0)
    for i in range(n):
        z.append(run)
        for r in ratios:
            run = e_mul(run, r[i])
    assert run == (1, 0), "grand product must be one"
    partials, prev = [], z
    for r in ratios[:-1]:
        prev = [e_mul(p, q) for p, q in zip(prev, r)]
        partials.append(prev)
    return z, partials


def horner_ext(monomials, at):
    """f(at) for base-field coefficients (low degree first) at an Fp2 point."""
    acc = (0, 0)
    for c in reversed([int(x) for x in monomials]):
        acc = e_add(e_mul(acc, at), (c % P, 0))
    return acc


def lde_point(log_n, log_lde, t):
    """x(t) = 7 * w_{nL}^{bitrev_{nL}(t)} for the flat LDE index t."""
    bits = log_n + log_lde
    r = int(format(t, "0%db" % bits)[::-1], 2) if bits else 0
    return 7 * pow(omega(bits), r, P) % P


def quotient_copy_permutation_point(t, log_n, log_lde, var_ldes, sigma_ldes, z, partials, beta, gamma, alphas, chunk):
    """z(1)=1 term + copy-permutation relations at flat LDE index t (prover.rs:1189-1227, copy_permutation.rs:1000-1249).
    var_ldes / sigma_ldes: flat LDE columns; z, partials: (c0, c1) flat LDE columns."""
    n = 1 << log_n
    n_cols = len(var_ldes)
    ks = non_residues_for_copy_permutation(n, n_cols)
    x = lde_point(log_n, log_lde, t)
    coset, i = t >> log_n, t & (n - 1)
    br = lambda v: int(format(v, "0%db" % log_n)[::-1], 2) if log_n else 0
    tsh = (coset << log_n) | br((br(i) + 1) % n)
    zt = (int(z[0][t]), int(z[1][t]))
    l1 = (pow(x, n, P) - 1) * pow(x - 1, P - 2, P) % P
    q = e_mul(e_mul_base(((zt[0] - 1) % P, zt[1]), l1), alphas[0])
    chunks = [range(s, min(s + chunk, n_cols)) for s in range(0, n_cols, chunk)]
    for c, ch in enumerate(chunks):
        lhs = (int(partials[c][0][t]), int(partials[c][1][t])) if c + 1 < len(chunks) else (int(z[0][tsh]), int(z[1][tsh]))
        rhs = zt if c == 0 else (int(partials[c - 1][0][t]), int(partials[c - 1][1][t]))
        for j in ch:
            w = int(var_ldes[j][t])
            lhs = e_mul(lhs, e_add(e_add(e_mul_base(beta, int(sigma_ldes[j][t])), (w, 0)), gamma))
            rhs = e_mul(rhs, e_add(e_add(e_mul_base(beta, ks[j] * x % P), (w, 0)), gamma))
        d = ((lhs[0] - rhs[0]) % P, (lhs[1] - rhs[1]) % P)
        q = e_add(q, e_mul(d, alphas[c + 1]))
    return q


def vanishing_inverse(log_n, log_q, coset):
    n = 1 << log_n
    r = int(format(coset, "0%db" % log_q)[::-1], 2) if log_q else 0
    shift = 7 * pow(omega(log_n + log_q), r, P) % P
    return pow((pow(shift, n, P) - 1) % P, P - 2, P)
