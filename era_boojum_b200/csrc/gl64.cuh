// Goldilocks field (p = 2^64 - 2^32 + 1) arithmetic for sm_100a device code.
//
// Semantics follow the reference's scalar field (src/field/goldilocks/mod.rs:188-201 reduce,
// :215-233 add, :309-327 sub) and Fp2 = Fp[u]/(u^2-7) (src/field/goldilocks/extension.rs:15,
// src/field/traits/field.rs:407-447).  Representation contract used by every kernel in this library:
//
//   * "canonical"  : value < p
//   * "lazy"       : any 64-bit value congruent to the element (the reference tolerates the same, mod.rs:147-171)
//   * gl::mul / gl::sqr / gl::mul_pow2   : lazy inputs  -> CANONICAL output
//   * gl::add(a, b) / gl::sub(a, b)      : a lazy, b CANONICAL (b <= p suffices) -> lazy output (single wrap
//                                          correction is exact)
//   * gl::add_lazy / gl::sub_lazy        : both lazy (canonicalises b first)
//   * gl::canon                          : lazy -> canonical; applied on every store that leaves the library
//
// All arithmetic is exact mod p, so the lazy/canonical choice can never change a result that is
// observed through gl::canon (which is how the reference observes them: serde, hashing, transcript).
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace gl {

typedef unsigned long long u64;
typedef unsigned int u32;

static constexpr u64 P = 0xFFFFFFFF00000001ull;
static constexpr u64 EPS = 0xFFFFFFFFull;  // 2^64 mod p
static constexpr u64 MULT_GEN = 7ull;
static constexpr u64 RADIX2_GEN = 0x185629dcda58878cull;  // order 2^32 (mod.rs:111)
static constexpr u64 INV7 = 0x249249246db6db6eull;         // 7^-1 mod p (checked in tests)

__host__ __device__ __forceinline__ u64 canon(u64 a) { return a >= P ? a - P : a; }

// portable C versions (host path; also the in-kernel reference of the PTX self-test)
__host__ __device__ __forceinline__ u64 add_c(u64 a, u64 b) {
  u64 s = a + b;
  return s < a ? s + EPS : s;
}
__host__ __device__ __forceinline__ u64 sub_c(u64 a, u64 b) {
  u64 d = a - b;
  return a < b ? d - EPS : d;
}

#if defined(__CUDA_ARCH__) && !defined(BJ_GL_PORTABLE)
#define BJ_GL_PTX 1
__device__ __forceinline__ u64 pack2(u32 lo, u32 hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(u64 a, u32& lo, u32& hi) { asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(a)); }
#endif

// a lazy, b <= p -> lazy.  (2^64 = EPS mod p: one wrap correction; it cannot wrap twice because the wrapped sum is < b)
__host__ __device__ __forceinline__ u64 add(u64 a, u64 b) {
#ifdef BJ_GL_PTX
  u32 a0, a1, b0, b1, lo, hi;
  unpack2(a, a0, a1);
  unpack2(b, b0, b1);
  // carry chains only mix add.cc/addc or sub.cc/subc (never add.cc -> subc, whose flag sense differs in SASS);
  // "+ c*EPS" is written as a multiply-add so that ptxas places it on the (under-used) FMA pipe
  asm("{\n\t.reg .u32 c;\n\t"
      "add.cc.u32 %0, %2, %4;\n\t"
      "addc.cc.u32 %1, %3, %5;\n\t"
      "addc.u32 c, 0, 0;\n\t"
      "mad.lo.cc.u32 %0, c, 0xffffffff, %0;\n\t"
      "madc.hi.u32 %1, c, 0xffffffff, %1;\n\t"
      "}"
      : "=&r"(lo), "=&r"(hi)
      : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
  return pack2(lo, hi);
#else
  return add_c(a, b);
#endif
}
// a lazy, b <= p -> lazy
__host__ __device__ __forceinline__ u64 sub(u64 a, u64 b) {
#ifdef BJ_GL_PTX
  u32 a0, a1, b0, b1, lo, hi;
  unpack2(a, a0, a1);
  unpack2(b, b0, b1);
  asm("{\n\t.reg .u32 m;\n\t"
      "sub.cc.u32 %0, %2, %4;\n\t"
      "subc.cc.u32 %1, %3, %5;\n\t"
      "subc.u32 m, 0, 0;\n\t"        // m = borrow ? 0xffffffff : 0
      "sub.cc.u32 %0, %0, m;\n\t"    // -= EPS on borrow
      "subc.u32 %1, %1, 0;\n\t"
      "}"
      : "=&r"(lo), "=&r"(hi)
      : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
  return pack2(lo, hi);
#else
  return sub_c(a, b);
#endif
}
// add() with the wrap correction on the ALU pipe (negate + add chain) instead of the FMA pipe (multiply-add by EPS).
// The NTT butterfly is FMA-pipe bound with every "* EPS" written as a multiply-add (50 FMA-pipe vs 38 ALU-pipe cycles per
// warp-butterfly, profiles/r1_ubench_integer_pipes.txt); moving ONE of its four corrections back to the ALU pipe balances the
// two pipes (44 / 44).  Same contract as add(): a lazy, b <= p -> lazy.
__host__ __device__ __forceinline__ u64 add_alu(u64 a, u64 b) {
#ifdef BJ_GL_PTX
  u32 a0, a1, b0, b1, lo, hi;
  unpack2(a, a0, a1);
  unpack2(b, b0, b1);
  asm("{\n\t.reg .u32 c, m;\n\t"
      "add.cc.u32 %0, %2, %4;\n\t"
      "addc.cc.u32 %1, %3, %5;\n\t"
      "addc.u32 c, 0, 0;\n\t"
      "neg.s32 m, c;\n\t"            // m = carry ? 0xffffffff : 0  (= carry * EPS, the low word; the high word of EPS is 0)
      "add.cc.u32 %0, %0, m;\n\t"
      "addc.u32 %1, %1, 0;\n\t"
      "}"
      : "=&r"(lo), "=&r"(hi)
      : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
  return pack2(lo, hi);
#else
  return add_c(a, b);
#endif
}
// sub() with the borrow correction written as multiply-adds (FMA pipe): t - EPS == t + 0xFFFFFFFF00000001 (mod 2^64), i.e.
// lo += bb, hi += bb * 0xffffffff + carry with bb = 1 on borrow.  Same contract as sub(): a lazy, b <= p -> lazy.
__host__ __device__ __forceinline__ u64 sub_fma(u64 a, u64 b) {
#ifdef BJ_GL_PTX
  u32 a0, a1, b0, b1, lo, hi;
  unpack2(a, a0, a1);
  unpack2(b, b0, b1);
  asm("{\n\t.reg .u32 m, bb;\n\t"
      "sub.cc.u32 %0, %2, %4;\n\t"
      "subc.cc.u32 %1, %3, %5;\n\t"
      "subc.u32 m, 0, 0;\n\t"          // m = borrow ? 0xffffffff : 0
      "and.b32 bb, m, 1;\n\t"
      "mad.lo.cc.u32 %0, bb, 0x1, %0;\n\t"
      "madc.lo.u32 %1, bb, 0xffffffff, %1;\n\t"
      "}"
      : "=&r"(lo), "=&r"(hi)
      : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
  return pack2(lo, hi);
#else
  return sub_c(a, b);
#endif
}
__host__ __device__ __forceinline__ u64 add_lazy(u64 a, u64 b) { return add(a, canon(b)); }
__host__ __device__ __forceinline__ u64 sub_lazy(u64 a, u64 b) { return sub(a, canon(b)); }
__host__ __device__ __forceinline__ u64 neg(u64 a) {
  a = canon(a);
  return a ? P - a : 0;
}
__host__ __device__ __forceinline__ u64 dbl(u64 a) { return add(a, canon(a)); }

// 128-bit (hi:lo) -> canonical, using 2^64 = 2^32 - 1 and 2^96 = -1 (mod p)
__host__ __device__ __forceinline__ u64 reduce128(u64 lo, u64 hi) {
  u64 hh = hi >> 32, hl = hi & EPS;
  u64 t0 = lo - hh;
  if (lo < hh) t0 -= EPS;
  u64 t1 = hl * EPS;  // (hl << 32) - hl, < 2^64
  u64 r = t0 + t1;
  if (r < t1) r += EPS;
  return canon(r);
}

__host__ __device__ __forceinline__ u64 mul_c(u64 a, u64 b) {
#ifdef __CUDA_ARCH__
  u64 lo = a * b;
  u64 hi = __umul64hi(a, b);
#else
  unsigned __int128 x = (unsigned __int128)a * b;
  u64 lo = (u64)x, hi = (u64)(x >> 64);
#endif
  return reduce128(lo, hi);
}

__host__ __device__ __forceinline__ u64 mul(u64 a, u64 b) {
#ifdef BJ_GL_PTX
  // 4-limb product from four 32x32->64 multiply-adds (no overflow in any partial sum), then
  // x = (r1:r0) + r2*EPS - r3 (2^64 = EPS, 2^96 = -1 mod p).  The "*EPS" steps are multiply-adds (FMA pipe); every wrap
  // is corrected at once (each can happen at most once, see the bounds in DESIGN.md 4.1); output canonical.
  u32 a0, a1, b0, b1, v0, v1;
  unpack2(a, a0, a1);
  unpack2(b, b0, b1);
  asm("{\n\t"
      ".reg .u64 p0, p1, p2, p3, z;\n\t"
      ".reg .u32 r0, r1, r2, r3, x, y, w, c, bb, ge;\n\t"
      ".reg .pred q;\n\t"
      "mul.wide.u32 p0, %2, %4;\n\t"
      "mov.b64 {r0, x}, p0;\n\t"
      "cvt.u64.u32 z, x;\n\t"
      "mad.wide.u32 p1, %2, %5, z;\n\t"
      "mov.b64 {x, y}, p1;\n\t"
      "cvt.u64.u32 z, x;\n\t"
      "mad.wide.u32 p2, %3, %4, z;\n\t"
      "mov.b64 {r1, w}, p2;\n\t"
      "cvt.u64.u32 z, y;\n\t"
      "mad.wide.u32 p3, %3, %5, z;\n\t"
      "cvt.u64.u32 z, w;\n\t"
      "add.u64 p3, p3, z;\n\t"
      "mov.b64 {r2, r3}, p3;\n\t"
      // t = (r1:r0) + r2*EPS with carry c ; t += c*EPS (cannot carry again)
      "mad.lo.cc.u32 %0, r2, 0xffffffff, r0;\n\t"
      "madc.hi.cc.u32 %1, r2, 0xffffffff, r1;\n\t"
      "addc.u32 c, 0, 0;\n\t"
      "mad.lo.cc.u32 %0, c, 0xffffffff, %0;\n\t"
      "madc.hi.u32 %1, c, 0xffffffff, %1;\n\t"
      // t -= r3 ; on borrow t -= EPS (cannot borrow again)
      "sub.cc.u32 %0, %0, r3;\n\t"
      "subc.cc.u32 %1, %1, 0;\n\t"
      "subc.u32 bb, 0, 0;\n\t"
      "sub.cc.u32 %0, %0, bb;\n\t"
      "subc.u32 %1, %1, 0;\n\t"
      // canonicalise: t >= p  <=>  hi == 0xffffffff and lo != 0 ; then t += EPS (mod 2^64) == t - p
      "setp.eq.u32 q, %1, 0xffffffff;\n\t"
      "setp.ne.and.u32 q, %0, 0, q;\n\t"
      "selp.u32 ge, 1, 0, q;\n\t"
      "mad.lo.cc.u32 %0, ge, 0xffffffff, %0;\n\t"
      "madc.hi.u32 %1, ge, 0xffffffff, %1;\n\t"
      "}"
      : "=&r"(v0), "=&r"(v1)
      : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
  return pack2(v0, v1);
#else
  return mul_c(a, b);
#endif
}
__host__ __device__ __forceinline__ u64 sqr(u64 a) { return mul(a, a); }


// Product with a LAZY result (< 2^64, congruent, not canonical): mul() without the final canonicalising step.
// Used where the consumer accepts lazy values (wide accumulators, further multiplications).
__host__ __device__ __forceinline__ u64 mul_lazy(u64 a, u64 b) {
#ifdef BJ_GL_PTX
  u32 a0, a1, b0, b1, v0, v1;
  unpack2(a, a0, a1);
  unpack2(b, b0, b1);
  asm("{\n\t"
      ".reg .u64 p0, p1, p2, p3, z;\n\t"
      ".reg .u32 r0, r1, r2, r3, x, y, w, c, bb;\n\t"
      "mul.wide.u32 p0, %2, %4;\n\t"
      "mov.b64 {r0, x}, p0;\n\t"
      "cvt.u64.u32 z, x;\n\t"
      "mad.wide.u32 p1, %2, %5, z;\n\t"
      "mov.b64 {x, y}, p1;\n\t"
      "cvt.u64.u32 z, x;\n\t"
      "mad.wide.u32 p2, %3, %4, z;\n\t"
      "mov.b64 {r1, w}, p2;\n\t"
      "cvt.u64.u32 z, y;\n\t"
      "mad.wide.u32 p3, %3, %5, z;\n\t"
      "cvt.u64.u32 z, w;\n\t"
      "add.u64 p3, p3, z;\n\t"
      "mov.b64 {r2, r3}, p3;\n\t"
      "mad.lo.cc.u32 %0, r2, 0xffffffff, r0;\n\t"
      "madc.hi.cc.u32 %1, r2, 0xffffffff, r1;\n\t"
      "addc.u32 c, 0, 0;\n\t"
      "mad.lo.cc.u32 %0, c, 0xffffffff, %0;\n\t"
      "madc.hi.u32 %1, c, 0xffffffff, %1;\n\t"
      "sub.cc.u32 %0, %0, r3;\n\t"
      "subc.cc.u32 %1, %1, 0;\n\t"
      "subc.u32 bb, 0, 0;\n\t"
      "sub.cc.u32 %0, %0, bb;\n\t"
      "subc.u32 %1, %1, 0;\n\t"
      "}"
      : "=&r"(v0), "=&r"(v1)
      : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
  return pack2(v0, v1);
#else
  return mul_c(a, b);
#endif
}

// ---- 96-bit accumulator (u64 + u32 carry word) for sums of lazy values with small coefficients: additions are plain
// integer carry chains, one reduction (2^64 = EPS) at the end.  Valid while the true sum stays below 2^96.
struct w96 {
  u64 lo;
  u32 hi;
};
__host__ __device__ __forceinline__ w96 w96_from(u64 a) { return {a, 0u}; }
__host__ __device__ __forceinline__ w96 w96_add(w96 a, w96 b) {
#ifdef BJ_GL_PTX
  w96 r;
  asm("add.cc.u64 %0, %2, %4;\n\taddc.u32 %1, %3, %5;" : "=l"(r.lo), "=r"(r.hi) : "l"(a.lo), "r"(a.hi), "l"(b.lo), "r"(b.hi));
  return r;
#else
  const u64 lo = a.lo + b.lo;
  return {lo, a.hi + b.hi + (lo < a.lo ? 1u : 0u)};
#endif
}
__host__ __device__ __forceinline__ w96 w96_add64(w96 a, u64 b) {
#ifdef BJ_GL_PTX
  w96 r;
  asm("add.cc.u64 %0, %2, %4;\n\taddc.u32 %1, %3, 0;" : "=l"(r.lo), "=r"(r.hi) : "l"(a.lo), "r"(a.hi), "l"(b));
  return r;
#else
  const u64 lo = a.lo + b;
  return {lo, a.hi + (lo < a.lo ? 1u : 0u)};
#endif
}
__host__ __device__ __forceinline__ w96 w96_shl(w96 a, unsigned s) {  // 0 < s < 32
  return {a.lo << s, (a.hi << s) | (u32)(a.lo >> (64 - s))};
}
// a * 2^s as a 96-bit value, 0 <= s < 32
__host__ __device__ __forceinline__ w96 w96_from_shl(u64 a, unsigned s) {
  if (s == 0) return {a, 0u};
  return {a << s, (u32)(a >> (64 - s))};
}
// -> lazy u64
__host__ __device__ __forceinline__ u64 w96_reduce(w96 a) {
#ifdef BJ_GL_PTX
  u32 l0, l1, v0, v1;
  unpack2(a.lo, l0, l1);
  asm("{\n\t.reg .u32 c;\n\t"
      "mad.lo.cc.u32 %0, %4, 0xffffffff, %2;\n\t"
      "madc.hi.cc.u32 %1, %4, 0xffffffff, %3;\n\t"
      "addc.u32 c, 0, 0;\n\t"
      "mad.lo.cc.u32 %0, c, 0xffffffff, %0;\n\t"
      "madc.hi.u32 %1, c, 0xffffffff, %1;\n\t"
      "}"
      : "=&r"(v0), "=&r"(v1)
      : "r"(l0), "r"(l1), "r"(a.hi));
  return pack2(v0, v1);
#else
  const u64 t1 = (u64)a.hi * EPS;
  u64 r = a.lo + t1;
  if (r < t1) r += EPS;  // hi * EPS < 2^64 - 2^33, so the wrapped sum + EPS cannot wrap again
  return r;
#endif
}

// a * k for a 32-bit k, exact, as a 96-bit value: two 32x32->64 multiply-adds instead of the four of a full product.
// Further small terms can be added with w96_add64 before the single w96_reduce (the sum must stay below 2^96).
__host__ __device__ __forceinline__ w96 mul_u32_wide(u64 a, u32 k) {
  const u64 p0 = (u64)(u32)a * k;
  const u64 p1 = (u64)(u32)(a >> 32) * k + (p0 >> 32);  // < 2^64: (2^32-1)^2 + 2^32 - 1
  return {(p1 << 32) | (u32)p0, (u32)(p1 >> 32)};
}

// a * b + c (all lazy) -> lazy: the addend enters the 128-bit product before its single reduction (a*b + c < 2^128)
__host__ __device__ __forceinline__ u64 fma_lazy(u64 a, u64 b, u64 c) {
#ifdef BJ_GL_PTX
  u32 a0, a1, b0, b1, c0, c1, v0, v1;
  unpack2(a, a0, a1);
  unpack2(b, b0, b1);
  unpack2(c, c0, c1);
  asm("{\n\t"
      ".reg .u64 p0, p1, p2, p3, z;\n\t"
      ".reg .u32 r0, r1, r2, r3, x, y, w, c, bb;\n\t"
      "mul.wide.u32 p0, %2, %4;\n\t"
      "mov.b64 {r0, x}, p0;\n\t"
      "cvt.u64.u32 z, x;\n\t"
      "mad.wide.u32 p1, %2, %5, z;\n\t"
      "mov.b64 {x, y}, p1;\n\t"
      "cvt.u64.u32 z, x;\n\t"
      "mad.wide.u32 p2, %3, %4, z;\n\t"
      "mov.b64 {r1, w}, p2;\n\t"
      "cvt.u64.u32 z, y;\n\t"
      "mad.wide.u32 p3, %3, %5, z;\n\t"
      "cvt.u64.u32 z, w;\n\t"
      "add.u64 p3, p3, z;\n\t"
      "mov.b64 {r2, r3}, p3;\n\t"
      // + c into the four limbs
      "add.cc.u32 r0, r0, %6;\n\t"
      "addc.cc.u32 r1, r1, %7;\n\t"
      "addc.cc.u32 r2, r2, 0;\n\t"
      "addc.u32 r3, r3, 0;\n\t"
      "mad.lo.cc.u32 %0, r2, 0xffffffff, r0;\n\t"
      "madc.hi.cc.u32 %1, r2, 0xffffffff, r1;\n\t"
      "addc.u32 c, 0, 0;\n\t"
      "mad.lo.cc.u32 %0, c, 0xffffffff, %0;\n\t"
      "madc.hi.u32 %1, c, 0xffffffff, %1;\n\t"
      "sub.cc.u32 %0, %0, r3;\n\t"
      "subc.cc.u32 %1, %1, 0;\n\t"
      "subc.u32 bb, 0, 0;\n\t"
      "sub.cc.u32 %0, %0, bb;\n\t"
      "subc.u32 %1, %1, 0;\n\t"
      "}"
      : "=&r"(v0), "=&r"(v1)
      : "r"(a0), "r"(a1), "r"(b0), "r"(b1), "r"(c0), "r"(c1));
  return pack2(v0, v1);
#else
  return add_c(mul_c(a, b), canon(c));
#endif
}

// a * 2^s for 0 <= s < 64 (s compile-time or uniform): 128-bit shift then reduce
__host__ __device__ __forceinline__ u64 mul_pow2(u64 a, unsigned s) {
  if (s == 0) return canon(a);
  u64 lo = a << s, hi = a >> (64 - s);
  return reduce128(lo, hi);
}

__host__ __device__ inline u64 pow(u64 b, u64 e) {
  u64 r = 1;
  while (e) {
    if (e & 1) r = mul(r, b);
    b = sqr(b);
    e >>= 1;
  }
  return r;
}
__host__ __device__ inline u64 inv(u64 a) { return pow(a, P - 2); }

// omega_{2^k} = G^(2^(32-k))           (src/cs/implementations/utils.rs:13-28)
__host__ __device__ inline u64 omega(unsigned log_n) {
  u64 w = RADIX2_GEN;
  for (unsigned i = log_n; i < 32; i++) w = sqr(w);
  return w;
}

// ---- Fp2 = Fp[u]/(u^2 - 7); components canonical on output of mul/inv, lazy after add/sub ----
struct e2 {
  u64 c0, c1;
};
__host__ __device__ __forceinline__ e2 e2_canon(e2 a) { return {canon(a.c0), canon(a.c1)}; }
// b canonical
__host__ __device__ __forceinline__ e2 e2_add(e2 a, e2 b) { return {add(a.c0, b.c0), add(a.c1, b.c1)}; }
__host__ __device__ __forceinline__ e2 e2_sub(e2 a, e2 b) { return {sub(a.c0, b.c0), sub(a.c1, b.c1)}; }
__host__ __device__ __forceinline__ u64 mul7(u64 a) {  // a canonical or lazy -> canonical
  // 7a = 8a - a
  return canon(sub(mul_pow2(a, 3), canon(a)));
}
__host__ __device__ __forceinline__ e2 e2_mul(e2 a, e2 b) {
  // Karatsuba as in field.rs:407-426: v0 = a0 b0, v1 = a1 b1, c1 = (a0+a1)(b0+b1) - v0 - v1, c0 = v0 + 7 v1
  u64 v0 = mul(a.c0, b.c0), v1 = mul(a.c1, b.c1);
  u64 s = mul(add_lazy(a.c0, a.c1), add_lazy(b.c0, b.c1));
  u64 c1 = canon(sub(sub(s, v0), v1));
  u64 c0 = canon(add(v0, mul7(v1)));
  return {c0, c1};
}
// Fp2 product with LAZY inputs and LAZY outputs (for chains of products whose end is canonicalised once): the two closing
// canonicalisations are dropped and 7 v1 + v0 is one 64 x 3 bit product kept in 96 bits with the addend (one reduction).
__host__ __device__ __forceinline__ e2 e2_mul_lazy(e2 a, e2 b) {
  const u64 v0 = mul(a.c0, b.c0), v1 = mul(a.c1, b.c1);
  const u64 s = mul(add_lazy(a.c0, a.c1), add_lazy(b.c0, b.c1));
  return {w96_reduce(w96_add64(mul_u32_wide(v1, 7u), v0)), sub(sub(s, v0), v1)};
}
__host__ __device__ __forceinline__ e2 e2_mul_base(e2 a, u64 b) { return {mul(a.c0, b), mul(a.c1, b)}; }
__host__ __device__ __forceinline__ e2 e2_sqr(e2 a) { return e2_mul(a, a); }
__host__ __device__ inline e2 e2_inv(e2 a) {
  u64 n = canon(sub(sqr(a.c0), mul7(sqr(a.c1))));
  u64 ni = inv(n);
  return {mul(a.c0, ni), mul(neg(a.c1), ni)};
}

}  // namespace gl
