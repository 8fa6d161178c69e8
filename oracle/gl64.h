/*
 * ORACLE (test infrastructure, NOT product code).
 *
 * Goldilocks field p = 2^64 - 2^32 + 1 and its quadratic extension Fp[u]/(u^2 - 7),
 * restated in plain C from the reference's definitions:
 *   - scalar field, reduction of a 128-bit product   src/field/goldilocks/mod.rs:188-201
 *   - add / sub with wrap correction                 src/field/goldilocks/mod.rs:215-233, 309-327
 *   - generators (7, 0x185629dcda58878c, 2-adicity)  src/field/goldilocks/mod.rs:110-114
 *   - Fp2 non-residue 7                              src/field/goldilocks/extension.rs:15
 *   - Fp2 mul / square / inverse                     src/field/traits/field.rs:407-447, 484-510
 *
 * Unlike the reference (which tolerates values in [p, 2^64) internally) every function here
 * takes and returns CANONICAL values (< p); gl_canon() is applied at the input boundary.
 * All arithmetic is exact mod p, so canonical-vs-lazy representation cannot change results.
 */
#ifndef ORACLE_GL64_H
#define ORACLE_GL64_H
#include <stdint.h>
#include <stddef.h>

#define GL_P 0xFFFFFFFF00000001ull
#define GL_EPS 0xFFFFFFFFull
#define GL_MULT_GEN 7ull
#define GL_RADIX2_GEN 0x185629dcda58878cull /* order 2^32 */
#define GL_TWO_ADICITY 32

typedef unsigned __int128 gl_u128;

static inline uint64_t gl_canon(uint64_t a) { return a >= GL_P ? a - GL_P : a; }

static inline uint64_t gl_add(uint64_t a, uint64_t b) {
  uint64_t s = a + b;
  if (s < a || s >= GL_P) s -= GL_P; /* a,b < p so a+b < 2p: one subtraction suffices */
  return s;
}
static inline uint64_t gl_sub(uint64_t a, uint64_t b) { return a >= b ? a - b : a + (GL_P - b); }
static inline uint64_t gl_neg(uint64_t a) { return a ? GL_P - a : 0; }
/* definitional product (slow, used by tests to pin gl_mul) */
static inline uint64_t gl_mul_slow(uint64_t a, uint64_t b) { return (uint64_t)(((gl_u128)a * b) % GL_P); }
/* 128-bit product reduced with 2^64 = 2^32 - 1 and 2^96 = -1 (mod p), as in
 * from_u128_with_reduction, src/field/goldilocks/mod.rs:188-201; canonicalised at the end. */
static inline uint64_t gl_mul(uint64_t a, uint64_t b) {
  gl_u128 x = (gl_u128)a * b;
  uint64_t lo = (uint64_t)x, hi = (uint64_t)(x >> 64);
  uint64_t hh = hi >> 32, hl = hi & GL_EPS;
  uint64_t t0 = lo - hh;
  if (lo < hh) t0 -= GL_EPS; /* wrapped by 2^64 = EPS (mod p) */
  uint64_t t1 = hl * GL_EPS;
  uint64_t r = t0 + t1;
  if (r < t1) r += GL_EPS;
  return gl_canon(r);
}
static inline uint64_t gl_sqr(uint64_t a) { return gl_mul(a, a); }
static inline uint64_t gl_dbl(uint64_t a) { return gl_add(a, a); }

static inline uint64_t gl_pow(uint64_t b, uint64_t e) {
  uint64_t r = 1;
  while (e) {
    if (e & 1) r = gl_mul(r, b);
    b = gl_sqr(b);
    e >>= 1;
  }
  return r;
}
/* Fermat inverse (the reference uses a binary-GCD, inversion.rs:72+; same value) */
static inline uint64_t gl_inv(uint64_t a) { return gl_pow(a, GL_P - 2); }

/* omega_{2^k} = G^(2^(32-k))            src/cs/implementations/utils.rs:13-28 */
static inline uint64_t gl_omega(unsigned log_n) {
  uint64_t w = GL_RADIX2_GEN;
  for (unsigned i = log_n; i < GL_TWO_ADICITY; i++) w = gl_sqr(w);
  return w;
}

static inline size_t gl_bitrev(size_t x, unsigned bits) {
  size_t r = 0;
  for (unsigned i = 0; i < bits; i++) {
    r = (r << 1) | (x & 1);
    x >>= 1;
  }
  return r;
}

/* ---- Fp2 = Fp[u]/(u^2-7) ---- */
typedef struct {
  uint64_t c0, c1;
} gl2_t;

static inline gl2_t gl2_add(gl2_t a, gl2_t b) { return (gl2_t){gl_add(a.c0, b.c0), gl_add(a.c1, b.c1)}; }
static inline gl2_t gl2_sub(gl2_t a, gl2_t b) { return (gl2_t){gl_sub(a.c0, b.c0), gl_sub(a.c1, b.c1)}; }
static inline gl2_t gl2_mul(gl2_t a, gl2_t b) {
  /* (a0 + a1 u)(b0 + b1 u) = a0 b0 + 7 a1 b1 + (a0 b1 + a1 b0) u */
  uint64_t v0 = gl_mul(a.c0, b.c0), v1 = gl_mul(a.c1, b.c1);
  uint64_t c0 = gl_add(v0, gl_mul(7, v1));
  uint64_t c1 = gl_add(gl_mul(a.c0, b.c1), gl_mul(a.c1, b.c0));
  return (gl2_t){c0, c1};
}
static inline gl2_t gl2_mul_base(gl2_t a, uint64_t b) { return (gl2_t){gl_mul(a.c0, b), gl_mul(a.c1, b)}; }
static inline gl2_t gl2_inv(gl2_t a) {
  /* 1/(a0 + a1 u) = (a0 - a1 u)/(a0^2 - 7 a1^2) */
  uint64_t n = gl_sub(gl_sqr(a.c0), gl_mul(7, gl_sqr(a.c1)));
  uint64_t ni = gl_inv(n);
  return (gl2_t){gl_mul(a.c0, ni), gl_mul(gl_neg(a.c1), ni)};
}
#endif
