"""ORACLE (test infrastructure): the log-derivative lookup argument over specialised columns with the table id in a
constant column (LookupParameters::UseSpecializedColumnsWithTableIdAsConstant), restated with Python ints.

  compute_lookup_poly_pairs_specialized           src/cs/implementations/lookup_argument_in_ext.rs:320-947
      aggregated table columns  t_0 + gamma t_1 + ... + beta                       :422-520
      A_i = 1 / (beta + sum_j gamma^j col_{i,j} + gamma^w table_id)                 :522-760 (batch inverse of the denominators)
      B   = multiplicity / (beta + sum_j gamma^j t_j)                               :762-947
  compute_quotient_terms_for_lookup_specialized   :949-1319
      alpha_i (A_i (beta + sum_j gamma^j col_{i,j} + gamma^w id) - 1)               :1115-1218
      alpha_k (B (beta + sum_j gamma^j t_j) - m)                                    :1221-1318
Small domains only (pure-Python loops).
"""
from .replay import P, e_add, e_inv, e_mul, e_mul_base


def _gamma_powers(gamma, k):
    out = [(1, 0)]
    for _ in range(1, k):
        out.append(e_mul(out[-1], gamma))
    return out


def _aggregate(cols, r, beta, gp):
    acc = beta
    for g, c in zip(gp, cols):
        acc = e_add(acc, e_mul_base(g, int(c[r]) % P))
    return acc


def lookup_polys(lookup_cols, width, table_id_col, table_cols, multiplicity, beta, gamma):
    """lookup_cols: n_sub*width columns; table_cols: width (+1 with the id) columns.  Returns ([A_i], B) as lists of Fp2."""
    n = len(multiplicity)
    n_sub = len(lookup_cols) // width
    gp = _gamma_powers(gamma, len(table_cols))
    A = []
    for i in range(n_sub):
        cols = list(lookup_cols[i * width:(i + 1) * width]) + ([table_id_col] if table_id_col is not None else [])
        assert len(cols) == len(gp)
        A.append([e_inv(_aggregate(cols, r, beta, gp)) for r in range(n)])
    B = [e_mul_base(e_inv(_aggregate(table_cols, r, beta, gp)), int(multiplicity[r]) % P) for r in range(n)]
    return A, B


def quotient_lookup_point(t, lookup_ldes, width, table_id_lde, table_ldes, multiplicity_lde, a_ldes, b_lde, beta, gamma, alphas):
    """the lookup terms of the quotient at flat LDE index t; a_ldes: [(c0 col, c1 col)], b_lde: (c0 col, c1 col)."""
    n_sub = len(lookup_ldes) // width
    gp = _gamma_powers(gamma, len(table_ldes))
    q = (0, 0)
    for i in range(n_sub):
        cols = list(lookup_ldes[i * width:(i + 1) * width]) + ([table_id_lde] if table_id_lde is not None else [])
        d = _aggregate(cols, t, beta, gp)
        term = e_mul(d, (int(a_ldes[i][0][t]), int(a_ldes[i][1][t])))
        term = ((term[0] - 1) % P, term[1])
        q = e_add(q, e_mul(term, alphas[i]))
    d = _aggregate(table_ldes, t, beta, gp)
    term = e_mul(d, (int(b_lde[0][t]), int(b_lde[1][t])))
    term = ((term[0] - int(multiplicity_lde[t])) % P, term[1])
    return e_add(q, e_mul(term, alphas[n_sub]))
