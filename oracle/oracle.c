/*
 * ORACLE (test infrastructure, NOT product code; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library).
 *
 * CPU restatement, in plain C, of the reference algorithms on Boojum's polynomial-commitment
 * hot path.  Every function cites the reference lines (relative to /root/reference/) it follows.
 * Parallelisation mirrors the reference's rayon chunking (one serial NTT per column spread over
 * threads; row-chunked hashing) via OpenMP so it can double as the "restated reference" CPU baseline.
 *
 * Parity status: Poseidon2 / sponge / Merkle layout / transcript / DEEP / FRI fold are pinned by the
 * reference's own golden fixture proof.json (tests/test_oracle_golden.py, fixture extracted by
 * tools/make_golden.py).  The NTT has no numeric golden vector in the reference tree (all its NTT
 * tests draw from thread_rng); it is pinned by definition: equality with the O(n^2) DFT in bit-reversed
 * order with and without coset 7, exactly the identity the reference's tests assert
 * (src/fft/mod.rs:1344-1384, 1591-1634), plus the FRI final-monomial check in proof.json which
 * exercises iNTT indirectly.
 */
#include "gl64.h"
#include "poseidon_rc.h"
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------------
 * scalar helpers exported for tests
 * ---------------------------------------------------------------------------------------------- */
API uint64_t orc_add(uint64_t a, uint64_t b) { return gl_add(gl_canon(a), gl_canon(b)); }
API uint64_t orc_sub(uint64_t a, uint64_t b) { return gl_sub(gl_canon(a), gl_canon(b)); }
API uint64_t orc_mul(uint64_t a, uint64_t b) { return gl_mul(gl_canon(a), gl_canon(b)); }
API uint64_t orc_mul_slow(uint64_t a, uint64_t b) { return gl_mul_slow(gl_canon(a), gl_canon(b)); }
API uint64_t orc_inv(uint64_t a) { return gl_inv(gl_canon(a)); }
API uint64_t orc_pow(uint64_t a, uint64_t e) { return gl_pow(gl_canon(a), e); }
API uint64_t orc_omega(unsigned log_n) { return gl_omega(log_n); }
API void orc_ext_mul(const uint64_t a[2], const uint64_t b[2], uint64_t out[2]) {
  gl2_t r = gl2_mul((gl2_t){gl_canon(a[0]), gl_canon(a[1])}, (gl2_t){gl_canon(b[0]), gl_canon(b[1])});
  out[0] = r.c0;
  out[1] = r.c1;
}
API void orc_ext_inv(const uint64_t a[2], uint64_t out[2]) {
  gl2_t r = gl2_inv((gl2_t){gl_canon(a[0]), gl_canon(a[1])});
  out[0] = r.c0;
  out[1] = r.c1;
}
API int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
API void orc_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ------------------------------------------------------------------------------------------------
 * NTT family
 * ---------------------------------------------------------------------------------------------- */

/* bitreverse_enumeration_inplace                         src/fft/mod.rs:41-155 (semantics only) */
API void orc_bitreverse(uint64_t *a, unsigned log_n) {
  size_t n = (size_t)1 << log_n;
  for (size_t i = 0; i < n; i++) {
    size_t j = gl_bitrev(i, log_n);
    if (i < j) {
      uint64_t t = a[i];
      a[i] = a[j];
      a[j] = t;
    }
  }
}

/* precompute_twiddles_for_fft: tab[i] = w^bitrev_{n/2}(i), i < n/2, w = omega_n or its inverse.
 *                                                         src/cs/implementations/utils.rs:88-125
 * (the reference's O(n) primitivity assert loop :107-110 is optional here: with_assert) */
API void orc_twiddles(uint64_t *tab, unsigned log_n, int inverse, int with_assert) {
  size_t n = (size_t)1 << log_n;
  uint64_t w = gl_omega(log_n);
  if (inverse) w = gl_inv(w);
  if (with_assert) {
    volatile uint64_t sink = 0;
    for (size_t i = 1; i < n; i++) sink ^= gl_pow(w, i); /* assert_ne!(omega^i, 1) */
    (void)sink;
  }
  size_t half = n / 2;
  if (half == 0) return;
  uint64_t cur = 1;
  for (size_t i = 0; i < half; i++) {
    tab[i] = cur;
    cur = gl_mul(cur, w);
  }
  if (log_n >= 2) orc_bitreverse(tab, log_n - 1);
}

/* serial_ct_ntt_natural_to_bitreversed                    src/fft/mod.rs:659-734 */
static void ntt_network(uint64_t *a, unsigned log_n, const uint64_t *tab) {
  size_t n = (size_t)1 << log_n;
  if (n == 1) return;
  size_t pairs = n / 2, groups = 1, dist = n / 2;
  while (groups < n) {
    for (size_t k = 0; k < groups; k++) {
      size_t i1 = k * pairs * 2, i2 = i1 + pairs;
      uint64_t s = tab[k];
      for (size_t j = i1; j < i2; j++) {
        uint64_t u = a[j], v = gl_mul(a[j + dist], s);
        a[j + dist] = gl_sub(u, v);
        a[j] = gl_add(u, v);
      }
    }
    pairs /= 2;
    groups *= 2;
    dist /= 2;
  }
}

/* distribute_powers: a[i] *= c^i                           src/fft/mod.rs:308-317 */
static void distribute_powers(uint64_t *a, size_t n, uint64_t c) {
  uint64_t s = 1;
  for (size_t i = 0; i < n; i++) {
    a[i] = gl_mul(a[i], s);
    s = gl_mul(s, c);
  }
}

static void canon_vec(uint64_t *a, size_t n) {
  for (size_t i = 0; i < n; i++) a[i] = gl_canon(a[i]);
}

/* fft_natural_to_bitreversed                               src/fft/mod.rs:398-411 */
API void orc_ntt_n2b_tab(uint64_t *a, unsigned log_n, uint64_t coset, const uint64_t *tab) {
  size_t n = (size_t)1 << log_n;
  canon_vec(a, n);
  coset = gl_canon(coset);
  if (coset != 1) distribute_powers(a, n, coset);
  ntt_network(a, log_n, tab);
}

/* ifft_natural_to_natural                                  src/fft/mod.rs:464-491 */
API void orc_intt_n2n_tab(uint64_t *a, unsigned log_n, uint64_t coset, const uint64_t *inv_tab) {
  size_t n = (size_t)1 << log_n;
  canon_vec(a, n);
  coset = gl_canon(coset);
  ntt_network(a, log_n, inv_tab);
  orc_bitreverse(a, log_n);
  if (coset != 1) distribute_powers(a, n, gl_inv(coset));
  if (n > 1) {
    uint64_t ninv = gl_inv((uint64_t)n % GL_P);
    for (size_t i = 0; i < n; i++) a[i] = gl_mul(a[i], ninv);
  }
}

/* batched drivers: n_cols columns of n elements, column c at a + c*stride.
 * One serial NTT per column, columns spread across threads (src/cs/implementations/utils.rs:295-304) */
API void orc_ntt_n2b(uint64_t *a, unsigned log_n, size_t n_cols, size_t stride, uint64_t coset) {
  size_t n = (size_t)1 << log_n;
  uint64_t *tab = (uint64_t *)malloc(sizeof(uint64_t) * (n / 2 + 1));
  orc_twiddles(tab, log_n, 0, 0);
#pragma omp parallel for schedule(static)
  for (long c = 0; c < (long)n_cols; c++) orc_ntt_n2b_tab(a + (size_t)c * stride, log_n, coset, tab);
  free(tab);
}
API void orc_intt_n2n(uint64_t *a, unsigned log_n, size_t n_cols, size_t stride, uint64_t coset) {
  size_t n = (size_t)1 << log_n;
  uint64_t *tab = (uint64_t *)malloc(sizeof(uint64_t) * (n / 2 + 1));
  orc_twiddles(tab, log_n, 1, 0);
#pragma omp parallel for schedule(static)
  for (long c = 0; c < (long)n_cols; c++) orc_intt_n2n_tab(a + (size_t)c * stride, log_n, coset, tab);
  free(tab);
}

/* transform_raw_storages_to_lde + transform_monomials_to_lde
 *                                                          src/cs/implementations/utils.rs:270-403
 * in : n_cols columns of n Lagrange values (natural order), column c at in + c*n
 * out: [col][coset j][n] with coset shift 7 * omega_{nL}^{bitrev_L(j)}, values bit-reversed in the coset.
 * from_monomials != 0 skips the iNTT (transform_monomials_to_lde). */
API void orc_lde(const uint64_t *in, uint64_t *out, unsigned log_n, unsigned log_lde, size_t n_cols,
                 int from_monomials) {
  size_t n = (size_t)1 << log_n, L = (size_t)1 << log_lde;
  uint64_t *ftab = (uint64_t *)malloc(sizeof(uint64_t) * (n / 2 + 1));
  uint64_t *itab = (uint64_t *)malloc(sizeof(uint64_t) * (n / 2 + 1));
  orc_twiddles(ftab, log_n, 0, 0);
  orc_twiddles(itab, log_n, 1, 0);
  uint64_t w_big = gl_omega(log_n + log_lde);
  uint64_t *mono = (uint64_t *)malloc(sizeof(uint64_t) * n * n_cols);
  memcpy(mono, in, sizeof(uint64_t) * n * n_cols);
#pragma omp parallel for schedule(static)
  for (long c = 0; c < (long)n_cols; c++) {
    if (!from_monomials) orc_intt_n2n_tab(mono + (size_t)c * n, log_n, 1, itab);
    else canon_vec(mono + (size_t)c * n, n);
  }
#pragma omp parallel for schedule(static)
  for (long job = 0; job < (long)(n_cols * L); job++) {
    size_t j = (size_t)job / n_cols, c = (size_t)job % n_cols; /* coset-major jobs, utils.rs:363-379 */
    uint64_t shift = gl_mul(GL_MULT_GEN, gl_pow(w_big, gl_bitrev(j, log_lde)));
    uint64_t *dst = out + (c * L + j) * n;
    memcpy(dst, mono + c * n, sizeof(uint64_t) * n);
    orc_ntt_n2b_tab(dst, log_n, shift, ftab);
  }
  free(mono);
  free(ftab);
  free(itab);
}

/* O(n^2) definition: out[bitrev(k)] = sum_i a_i (c w^k)^i  (what the reference tests compare against,
 * src/fft/mod.rs:1344-1384) */
API void orc_naive_dft_bitreversed(const uint64_t *a, uint64_t *out, unsigned log_n, uint64_t coset) {
  size_t n = (size_t)1 << log_n;
  uint64_t w = gl_omega(log_n);
  coset = gl_canon(coset);
  for (size_t k = 0; k < n; k++) {
    uint64_t x = gl_mul(coset, gl_pow(w, k)), acc = 0;
    for (size_t i = n; i-- > 0;) acc = gl_add(gl_mul(acc, x), gl_canon(a[i]));
    out[gl_bitrev(k, log_n)] = acc;
  }
}

/* ------------------------------------------------------------------------------------------------
 * Poseidon2 (t = 12, x^7, 4 + 22 + 4 rounds) and the overwrite-mode sponge
 * ---------------------------------------------------------------------------------------------- */

/* M4 block of the external matrix                         src/implementations/suggested_mds.rs:19-56 */
static inline void p2_block(uint64_t *x) {
  uint64_t t0 = gl_add(x[0], x[1]);
  uint64_t t1 = gl_add(x[2], x[3]);
  uint64_t t2 = gl_add(gl_dbl(x[1]), t1);
  uint64_t t3 = gl_add(gl_dbl(x[3]), t0);
  uint64_t t4 = gl_add(gl_dbl(gl_dbl(t1)), t3);
  uint64_t t5 = gl_add(gl_dbl(gl_dbl(t0)), t2);
  x[0] = gl_add(t3, t5);
  x[1] = t5;
  x[2] = gl_add(t2, t4);
  x[3] = t4;
}
/* M_E = circ(2 M4, M4, M4)                                 src/implementations/suggested_mds.rs:59-97 */
static inline void p2_external(uint64_t *s) {
  p2_block(s);
  p2_block(s + 4);
  p2_block(s + 8);
  for (int i = 0; i < 4; i++) {
    uint64_t sum = gl_add(gl_add(s[i], s[4 + i]), s[8 + i]);
    s[i] = gl_add(s[i], sum);
    s[4 + i] = gl_add(s[4 + i], sum);
    s[8 + i] = gl_add(s[8 + i], sum);
  }
}
static inline uint64_t p2_pow7(uint64_t x) {
  uint64_t x2 = gl_sqr(x), x3 = gl_mul(x2, x), x4 = gl_sqr(x2);
  return gl_mul(x4, x3);
}
/* internal matrix diag(2^s_i) + J                          state_generic_impl.rs:69-82, 171-200 */
static const unsigned P2_DIAG_SHIFT[12] = {4, 14, 11, 8, 0, 5, 2, 9, 13, 6, 3, 12};
static inline void p2_internal(uint64_t *s) {
  uint64_t sum = 0;
  for (int i = 0; i < 12; i++) sum = gl_add(sum, s[i]);
  for (int i = 0; i < 12; i++) s[i] = gl_add(gl_mul(s[i], (uint64_t)1 << P2_DIAG_SHIFT[i]), sum);
}
/* poseidon2_permutation: one global round counter 0..29    state_generic_impl.rs:158-233 */
API void orc_poseidon2_permutation(uint64_t *s) {
  for (int i = 0; i < 12; i++) s[i] = gl_canon(s[i]);
  p2_external(s);
  int r = 0;
  for (int k = 0; k < 4; k++, r++) {
    for (int i = 0; i < 12; i++) s[i] = p2_pow7(gl_add(s[i], ORACLE_POSEIDON_RC[r * 12 + i]));
    p2_external(s);
  }
  for (int k = 0; k < 22; k++, r++) {
    s[0] = p2_pow7(gl_add(s[0], ORACLE_POSEIDON_RC[r * 12]));
    p2_internal(s);
  }
  for (int k = 0; k < 4; k++, r++) {
    for (int i = 0; i < 12; i++) s[i] = p2_pow7(gl_add(s[i], ORACLE_POSEIDON_RC[r * 12 + i]));
    p2_external(s);
  }
}

/* GoldilocksPoseidon2Sponge<AbsorptionModeOverwrite>::hash_into_leaf
 *   src/cs/oracle/mod.rs:141-151, src/algebraic_props/sponge.rs:224-239 (absorb), :300-323 (finalize) */
API void orc_poseidon2_hash_leaf(const uint64_t *els, size_t n, uint64_t out[4]) {
  uint64_t st[12] = {0};
  size_t i = 0;
  for (; i + 8 <= n; i += 8) {
    for (int k = 0; k < 8; k++) st[k] = gl_canon(els[i + k]);
    orc_poseidon2_permutation(st);
  }
  if (i < n) {
    size_t f = n - i;
    for (size_t k = 0; k < f; k++) st[k] = gl_canon(els[i + k]);
    for (size_t k = f; k < 8; k++) st[k] = 0;
    orc_poseidon2_permutation(st);
  }
  memcpy(out, st, 4 * sizeof(uint64_t));
}
/* hash_into_node(l, r) = sponge over l || r = one permutation   src/cs/oracle/mod.rs:162-168 */
API void orc_poseidon2_hash_node(const uint64_t l[4], const uint64_t r[4], uint64_t out[4]) {
  uint64_t st[12] = {0};
  for (int k = 0; k < 4; k++) {
    st[k] = gl_canon(l[k]);
    st[4 + k] = gl_canon(r[k]);
  }
  orc_poseidon2_permutation(st);
  memcpy(out, st, 4 * sizeof(uint64_t));
}

/* Leaf hashing for MerkleTreeWithCap::construct (elems_per_leaf == 1, merkle_tree.rs:78-172) and
 * construct_by_chunking / _from_flat_sources (elems_per_leaf == 2^k, :176-386):
 * leaf m = concat over sources s (in order) of source_s[m*k .. (m+1)*k).
 * sources: n_src pointers to flat arrays of n_leaves*k u64 (a column's cosets flattened coset-major). */
API void orc_merkle_leaf_hashes(const uint64_t *const *sources, size_t n_src, size_t n_leaves,
                                size_t elems_per_leaf, uint64_t *leaf_hashes) {
#pragma omp parallel
  {
    uint64_t *buf = (uint64_t *)malloc(sizeof(uint64_t) * n_src * elems_per_leaf);
#pragma omp for schedule(static)
    for (long m = 0; m < (long)n_leaves; m++) {
      size_t w = 0;
      for (size_t s = 0; s < n_src; s++)
        for (size_t e = 0; e < elems_per_leaf; e++) buf[w++] = sources[s][(size_t)m * elems_per_leaf + e];
      orc_poseidon2_hash_leaf(buf, w, leaf_hashes + 4 * (size_t)m);
    }
    free(buf);
  }
}

/* continue_from_leaf_hashes: level by level until cap_size nodes remain  merkle_tree.rs:388-449
 * nodes: concatenated levels (n_leaves/2, n_leaves/4, ..., cap_size digests); returns #digests written.
 * The cap is the last level (or the leaf hashes themselves if n_leaves == cap_size). */
API size_t orc_merkle_nodes(const uint64_t *leaf_hashes, size_t n_leaves, size_t cap_size, uint64_t *nodes) {
  const uint64_t *prev = leaf_hashes;
  size_t cnt = n_leaves, written = 0;
  while (cnt > cap_size) {
    size_t next = cnt / 2;
    uint64_t *dst = nodes + 4 * written;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)next; i++)
      orc_poseidon2_hash_node(prev + 8 * (size_t)i, prev + 8 * (size_t)i + 4, dst + 4 * (size_t)i);
    prev = dst;
    written += next;
    cnt = next;
  }
  return written;
}

/* verify_proof_over_cap                                     merkle_tree.rs:482-504 */
API int orc_merkle_verify(const uint64_t leaf_hash[4], const uint64_t *path, size_t path_len,
                          const uint64_t *cap, size_t cap_size, size_t idx) {
  uint64_t cur[4];
  memcpy(cur, leaf_hash, sizeof cur);
  for (size_t d = 0; d < path_len; d++) {
    const uint64_t *sib = path + 4 * d;
    uint64_t nxt[4];
    if (idx & 1) orc_poseidon2_hash_node(sib, cur, nxt);
    else orc_poseidon2_hash_node(cur, sib, nxt);
    memcpy(cur, nxt, sizeof cur);
    idx >>= 1;
  }
  if (idx >= cap_size) return 0;
  for (int k = 0; k < 4; k++)
    if (gl_canon(cap[4 * idx + k]) != cur[k]) return 0;
  return 1;
}

/* ------------------------------------------------------------------------------------------------
 * FRI fold                                                   src/cs/implementations/fri/mod.rs:362-474
 * out[i] = (f[2i] + f[2i+1]) + alpha * (f[2i] - f[2i+1]) * roots[i] * coset_inv   (Fp2, no 1/2)
 * roots = inverse twiddle table of the FULL LDE domain (prefix reused every level, :509, :612)
 * ---------------------------------------------------------------------------------------------- */
API void orc_fri_fold(const uint64_t *c0, const uint64_t *c1, size_t m, const uint64_t alpha[2],
                      const uint64_t *roots, uint64_t coset_inv, uint64_t *o0, uint64_t *o1) {
  gl2_t al = {gl_canon(alpha[0]), gl_canon(alpha[1])};
  coset_inv = gl_canon(coset_inv);
#pragma omp parallel for schedule(static)
  for (long i = 0; i < (long)(m / 2); i++) {
    uint64_t a0 = gl_canon(c0[2 * i]), b0 = gl_canon(c0[2 * i + 1]);
    uint64_t a1 = gl_canon(c1[2 * i]), b1 = gl_canon(c1[2 * i + 1]);
    uint64_t r = gl_mul(gl_canon(roots[i]), coset_inv);
    gl2_t diff = {gl_mul(gl_sub(a0, b0), r), gl_mul(gl_sub(a1, b1), r)};
    gl2_t t = gl2_mul(diff, al);
    o0[i] = gl_add(gl_add(t.c0, a0), b0);
    o1[i] = gl_add(gl_add(t.c1, a1), b1);
  }
}

/* ------------------------------------------------------------------------------------------------
 * batch inverse (Montgomery trick)                           src/cs/implementations/utils.rs:405-600
 * zeros are not expected by the reference (it would panic on inverse of 0); here 0 -> 0 is NOT
 * special-cased either: callers must pass non-zero values.
 * ---------------------------------------------------------------------------------------------- */
API void orc_batch_inverse(uint64_t *a, size_t n) {
  if (n == 0) return;
  uint64_t *pre = (uint64_t *)malloc(sizeof(uint64_t) * n);
  uint64_t acc = 1;
  for (size_t i = 0; i < n; i++) {
    pre[i] = acc;
    acc = gl_mul(acc, gl_canon(a[i]));
  }
  uint64_t inv = gl_inv(acc);
  for (size_t i = n; i-- > 0;) {
    uint64_t ai = gl_canon(a[i]);
    a[i] = gl_mul(inv, pre[i]);
    inv = gl_mul(inv, ai);
  }
  free(pre);
}
API void orc_batch_inverse_ext(uint64_t *c0, uint64_t *c1, size_t n) {
  if (n == 0) return;
  gl2_t *pre = (gl2_t *)malloc(sizeof(gl2_t) * n);
  gl2_t acc = {1, 0};
  for (size_t i = 0; i < n; i++) {
    pre[i] = acc;
    acc = gl2_mul(acc, (gl2_t){gl_canon(c0[i]), gl_canon(c1[i])});
  }
  gl2_t inv = gl2_inv(acc);
  for (size_t i = n; i-- > 0;) {
    gl2_t ai = {gl_canon(c0[i]), gl_canon(c1[i])};
    gl2_t r = gl2_mul(inv, pre[i]);
    c0[i] = r.c0;
    c1[i] = r.c1;
    inv = gl2_mul(inv, ai);
  }
  free(pre);
}

/* ------------------------------------------------------------------------------------------------
 * DEEP quotening on one (coset,row):                         src/cs/implementations/prover.rs:2523-2706,
 * verifier-side statement src/cs/implementations/verifier.rs:2526-2565.
 * acc += (1/(x - at)) * sum_i ch_i * (f_i(x) - v_i), f_i in Fp2 given as (c0,c1).
 * ---------------------------------------------------------------------------------------------- */
API void orc_deep_point(uint64_t acc[2], const uint64_t *f_c0, const uint64_t *f_c1, const uint64_t *v_c0,
                        const uint64_t *v_c1, const uint64_t *ch_c0, const uint64_t *ch_c1, size_t n,
                        uint64_t x, const uint64_t at[2]) {
  gl2_t den = gl2_sub((gl2_t){gl_canon(x), 0}, (gl2_t){gl_canon(at[0]), gl_canon(at[1])});
  den = gl2_inv(den);
  gl2_t s = {0, 0};
  for (size_t i = 0; i < n; i++) {
    gl2_t d = gl2_sub((gl2_t){gl_canon(f_c0[i]), gl_canon(f_c1[i])}, (gl2_t){gl_canon(v_c0[i]), gl_canon(v_c1[i])});
    s = gl2_add(s, gl2_mul((gl2_t){gl_canon(ch_c0[i]), gl_canon(ch_c1[i])}, d));
  }
  s = gl2_mul(s, den);
  gl2_t r = gl2_add((gl2_t){gl_canon(acc[0]), gl_canon(acc[1])}, s);
  acc[0] = r.c0;
  acc[1] = r.c1;
}

/* DEEP accumulation over the whole LDE domain for one opening point (prover-side driver of the sum above):
 * quotening_operation_in_extension                              src/cs/implementations/prover.rs:2523-2706
 * for t in 0..2^log_rows: x = 7 * w_{nL}^{bitrev(t)};  acc[t] += sum_i ch_i (f_i[t] - v_i) / (x - at)
 * src_c1[i] == NULL marks a base-field polynomial. */
API void orc_deep_group(uint64_t *acc0, uint64_t *acc1, const uint64_t *const *src_c0, const uint64_t *const *src_c1,
                        size_t n_src, const uint64_t *values_at, const uint64_t *challenges, const uint64_t at[2],
                        unsigned log_rows) {
  size_t rows = (size_t)1 << log_rows;
  uint64_t w = gl_omega(log_rows);
  gl2_t a = {gl_canon(at[0]), gl_canon(at[1])};
#pragma omp parallel for schedule(static)
  for (long t = 0; t < (long)rows; t++) {
    uint64_t x = gl_mul(GL_MULT_GEN, gl_pow(w, gl_bitrev((size_t)t, log_rows)));
    gl2_t den = gl2_inv(gl2_sub((gl2_t){x, 0}, a));
    gl2_t s = {0, 0};
    for (size_t i = 0; i < n_src; i++) {
      gl2_t f = {gl_canon(src_c0[i][t]), src_c1[i] ? gl_canon(src_c1[i][t]) : 0};
      gl2_t v = {gl_canon(values_at[2 * i]), gl_canon(values_at[2 * i + 1])};
      gl2_t c = {gl_canon(challenges[2 * i]), gl_canon(challenges[2 * i + 1])};
      s = gl2_add(s, gl2_mul(c, gl2_sub(f, v)));
    }
    s = gl2_mul(s, den);
    acc0[t] = gl_add(gl_canon(acc0[t]), s.c0);
    acc1[t] = gl_add(gl_canon(acc1[t]), s.c1);
  }
}
