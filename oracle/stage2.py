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
