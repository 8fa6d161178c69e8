// Variants of the Goldilocks butterfly arithmetic: correctness against portable C + cycles per warp-butterfly in a pure
// register stream.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mul_variants mul_variants.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;
static constexpr u64 P = 0xFFFFFFFF00000001ull, EPS = 0xFFFFFFFFull;
__host__ __device__ inline u64 canon(u64 a) { return a >= P ? a - P : a; }
__host__ __device__ inline u64 add_c(u64 a, u64 b) { u64 s = a + b; return s < a ? s + EPS : s; }
__host__ __device__ inline u64 sub_c(u64 a, u64 b) { u64 d = a - b; return a < b ? d - EPS : d; }
__host__ __device__ inline u64 reduce128(u64 lo, u64 hi) {
  u64 hh = hi >> 32, hl = hi & EPS;
  u64 t0 = lo - hh;
  if (lo < hh) t0 -= EPS;
  u64 t1 = hl * EPS;
  u64 r = t0 + t1;
  if (r < t1) r += EPS;
  return canon(r);
}
__device__ inline u64 mul_c(u64 a, u64 b) { return reduce128(a * b, __umul64hi(a, b)); }
__device__ __forceinline__ u64 pack2(u32 lo, u32 hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi)); return r; }
__device__ __forceinline__ void unpack2(u64 a, u32& lo, u32& hi) { asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(a)); }

#define PRODUCT                                    \
  ".reg .u64 p0, p1, p2, p3, z;\n\t"               \
  ".reg .u32 r0, r1, r2, r3, x, y, w, c, m, ul, uh;\n\t" \
  "mul.wide.u32 p0, %2, %4;\n\t"                   \
  "mov.b64 {r0, x}, p0;\n\t"                       \
  "cvt.u64.u32 z, x;\n\t"                          \
  "mad.wide.u32 p1, %2, %5, z;\n\t"                \
  "mov.b64 {x, y}, p1;\n\t"                        \
  "cvt.u64.u32 z, x;\n\t"                          \
  "mad.wide.u32 p2, %3, %4, z;\n\t"                \
  "mov.b64 {r1, w}, p2;\n\t"                       \
  "cvt.u64.u32 z, y;\n\t"                          \
  "mad.wide.u32 p3, %3, %5, z;\n\t"                \
  "cvt.u64.u32 z, w;\n\t"                          \
  "add.u64 p3, p3, z;\n\t"                         \
  "mov.b64 {r2, r3}, p3;\n\t"

// ---- variant 0: the arithmetic shipped before this experiment (IMAD/IMAD.HI corrections, canonical product)
struct V0 {
  static __device__ __forceinline__ u64 add(u64 a, u64 b) {
    u32 a0, a1, b0, b1, lo, hi; unpack2(a, a0, a1); unpack2(b, b0, b1);
    asm("{\n\t.reg .u32 c;\n\tadd.cc.u32 %0, %2, %4;\n\taddc.cc.u32 %1, %3, %5;\n\taddc.u32 c, 0, 0;\n\t"
        "mad.lo.cc.u32 %0, c, 0xffffffff, %0;\n\tmadc.hi.u32 %1, c, 0xffffffff, %1;\n\t}"
        : "=&r"(lo), "=&r"(hi) : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    return pack2(lo, hi);
  }
  static __device__ __forceinline__ u64 sub(u64 a, u64 b) {
    u32 a0, a1, b0, b1, lo, hi; unpack2(a, a0, a1); unpack2(b, b0, b1);
    asm("{\n\t.reg .u32 m;\n\tsub.cc.u32 %0, %2, %4;\n\tsubc.cc.u32 %1, %3, %5;\n\tsubc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 %0, %0, m;\n\tsubc.u32 %1, %1, 0;\n\t}"
        : "=&r"(lo), "=&r"(hi) : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    return pack2(lo, hi);
  }
  static __device__ __forceinline__ u64 mul(u64 a, u64 b, bool& bad) {
    u32 a0, a1, b0, b1, v0, v1; unpack2(a, a0, a1); unpack2(b, b0, b1);
    asm("{\n\t" PRODUCT ".reg .u32 bb, ge;\n\t.reg .pred q;\n\t"
        "mad.lo.cc.u32 %0, r2, 0xffffffff, r0;\n\tmadc.hi.cc.u32 %1, r2, 0xffffffff, r1;\n\taddc.u32 c, 0, 0;\n\t"
        "mad.lo.cc.u32 %0, c, 0xffffffff, %0;\n\tmadc.hi.u32 %1, c, 0xffffffff, %1;\n\t"
        "sub.cc.u32 %0, %0, r3;\n\tsubc.cc.u32 %1, %1, 0;\n\tsubc.u32 bb, 0, 0;\n\tsub.cc.u32 %0, %0, bb;\n\tsubc.u32 %1, %1, 0;\n\t"
        "setp.eq.u32 q, %1, 0xffffffff;\n\tsetp.ne.and.u32 q, %0, 0, q;\n\tselp.u32 ge, 1, 0, q;\n\t"
        "mad.lo.cc.u32 %0, ge, 0xffffffff, %0;\n\tmadc.hi.u32 %1, ge, 0xffffffff, %1;\n\t}"
        : "=&r"(v0), "=&r"(v1) : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    return pack2(v0, v1);
  }
};

// ---- variant 1: add/sub chains only (no IMAD in the corrections), lazy product + suspect flag.
//      r2*EPS = (r2 << 32) - r2 built with two subs; masks from the carry / borrow flags with subc m, 0, 0
//      (after an add chain this is m = -carry by the PTX definition d = a - (b + CC.CF)).
struct V1 {
  static __device__ __forceinline__ u64 add(u64 a, u64 b) {
    u32 a0, a1, b0, b1, lo, hi; unpack2(a, a0, a1); unpack2(b, b0, b1);
    asm("{\n\t.reg .u32 m;\n\tadd.cc.u32 %0, %2, %4;\n\taddc.cc.u32 %1, %3, %5;\n\tsubc.u32 m, 0, 0;\n\t"
        "add.cc.u32 %0, %0, m;\n\taddc.u32 %1, %1, 0;\n\t}"
        : "=&r"(lo), "=&r"(hi) : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    return pack2(lo, hi);
  }
  static __device__ __forceinline__ u64 sub(u64 a, u64 b) { return V0::sub(a, b); }
  static __device__ __forceinline__ u64 mul(u64 a, u64 b, bool& bad) {
    u32 a0, a1, b0, b1, v0, v1; unpack2(a, a0, a1); unpack2(b, b0, b1);
    asm("{\n\t" PRODUCT
        "sub.cc.u32 ul, 0, r2;\n\tsubc.u32 uh, r2, 0;\n\t"                 // (uh:ul) = r2 * EPS
        "add.cc.u32 %0, r0, ul;\n\taddc.cc.u32 %1, r1, uh;\n\tsubc.u32 m, 0, 0;\n\t"  // y, m = -carry
        "add.cc.u32 %0, %0, m;\n\taddc.u32 %1, %1, 0;\n\t"                 // + carry * EPS
        "sub.cc.u32 %0, %0, r3;\n\tsubc.cc.u32 %1, %1, 0;\n\tsubc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 %0, %0, m;\n\tsubc.u32 %1, %1, 0;\n\t}"               // - borrow * EPS
        : "=&r"(v0), "=&r"(v1) : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    bad |= v1 == 0xffffffffu;
    return pack2(v0, v1);
  }
};

// ---- variant 2: as 1 but the carry is materialised with addc (no add->subc mixing) and negated
struct V2 {
  static __device__ __forceinline__ u64 add(u64 a, u64 b) {
    u32 a0, a1, b0, b1, lo, hi; unpack2(a, a0, a1); unpack2(b, b0, b1);
    asm("{\n\t.reg .u32 c, m;\n\tadd.cc.u32 %0, %2, %4;\n\taddc.cc.u32 %1, %3, %5;\n\taddc.u32 c, 0, 0;\n\tneg.s32 m, c;\n\t"
        "add.cc.u32 %0, %0, m;\n\taddc.u32 %1, %1, 0;\n\t}"
        : "=&r"(lo), "=&r"(hi) : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    return pack2(lo, hi);
  }
  static __device__ __forceinline__ u64 sub(u64 a, u64 b) { return V0::sub(a, b); }
  static __device__ __forceinline__ u64 mul(u64 a, u64 b, bool& bad) {
    u32 a0, a1, b0, b1, v0, v1; unpack2(a, a0, a1); unpack2(b, b0, b1);
    asm("{\n\t" PRODUCT
        "sub.cc.u32 ul, 0, r2;\n\tsubc.u32 uh, r2, 0;\n\t"
        "add.cc.u32 %0, r0, ul;\n\taddc.cc.u32 %1, r1, uh;\n\taddc.u32 c, 0, 0;\n\tneg.s32 m, c;\n\t"
        "add.cc.u32 %0, %0, m;\n\taddc.u32 %1, %1, 0;\n\t"
        "sub.cc.u32 %0, %0, r3;\n\tsubc.cc.u32 %1, %1, 0;\n\tsubc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 %0, %0, m;\n\tsubc.u32 %1, %1, 0;\n\t}"
        : "=&r"(v0), "=&r"(v1) : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    bad |= v1 == 0xffffffffu;
    return pack2(v0, v1);
  }
};

// ---- variant 3: plain C (what nvcc makes of the portable code), lazy product + flag
struct V3 {
  static __device__ __forceinline__ u64 add(u64 a, u64 b) { return add_c(a, b); }
  static __device__ __forceinline__ u64 sub(u64 a, u64 b) { return sub_c(a, b); }
  static __device__ __forceinline__ u64 mul(u64 a, u64 b, bool& bad) {
    const u64 lo = a * b, hi = __umul64hi(a, b);
    u64 hh = hi >> 32, hl = hi & EPS;
    u64 t0 = lo - hh;
    if (lo < hh) t0 -= EPS;
    u64 t1 = hl * EPS;
    u64 r = t0 + t1;
    if (r < t1) r += EPS;
    bad |= (u32)(r >> 32) == 0xffffffffu;
    return r;
  }
};

// ---- variant 4: variant 1 with the r2*EPS product on the FMA pipe as ONE IMAD.WIDE (no carry needed: it is < 2^64)
struct V4 {
  static __device__ __forceinline__ u64 add(u64 a, u64 b) { return V1::add(a, b); }
  static __device__ __forceinline__ u64 sub(u64 a, u64 b) { return V0::sub(a, b); }
  static __device__ __forceinline__ u64 mul(u64 a, u64 b, bool& bad) {
    u32 a0, a1, b0, b1, v0, v1; unpack2(a, a0, a1); unpack2(b, b0, b1);
    asm("{\n\t" PRODUCT ".reg .u64 e;\n\t"
        "mul.wide.u32 e, r2, 0xffffffff;\n\tmov.b64 {ul, uh}, e;\n\t"
        "add.cc.u32 %0, r0, ul;\n\taddc.cc.u32 %1, r1, uh;\n\tsubc.u32 m, 0, 0;\n\t"
        "add.cc.u32 %0, %0, m;\n\taddc.u32 %1, %1, 0;\n\t"
        "sub.cc.u32 %0, %0, r3;\n\tsubc.cc.u32 %1, %1, 0;\n\tsubc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 %0, %0, m;\n\tsubc.u32 %1, %1, 0;\n\t}"
        : "=&r"(v0), "=&r"(v1) : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    bad |= v1 == 0xffffffffu;
    return pack2(v0, v1);
  }
};

template <class V, int ILP>
__global__ void __launch_bounds__(256) k_bfly(u64* out, u64 seed, int iters) {
  u64 a[ILP], b[ILP], w = seed | 1;
  bool bad = false;
#pragma unroll
  for (int i = 0; i < ILP; i++) { a[i] = seed + threadIdx.x * 977 + i * 131 + blockIdx.x; b[i] = a[i] * 3 + 1; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) {
      const u64 v = V::mul(b[i], w, bad);
      b[i] = V::sub(a[i], v);
      a[i] = V::add(a[i], v);
    }
  }
  u64 s = bad;
#pragma unroll
  for (int i = 0; i < ILP; i++) s ^= a[i] ^ b[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class V>
__global__ void k_check(u64 n, u64 seed, unsigned long long* mism) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  auto next = [](u64& s) { s += 0x9E3779B97F4A7C15ull; u64 z = s; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); };
  u64 st = seed + i * 0x632BE59BD9B4E019ull;
  const u64 edges[12] = {0, 1, P - 1, P, P + 1, ~0ull, EPS, EPS + 1, 1ull << 63, P - EPS, 2, P - 2};
  u64 a = next(st), b = next(st);
  if ((i & 7) == 3) {
    const u64 r = next(st);
    a = (1ull << ((i >> 3) & 63)) + (((i >> 9) & 3) == 1 ? 1 : 0) - (((i >> 9) & 3) == 2 ? 1 : 0);
    const unsigned sel = (unsigned)(i >> 11) & 7;
    u64 lo = r & 0xffffffffull, hi = r >> 32;
    if (sel & 1) lo = (sel & 4) ? 0xffffffffull : 0;
    if (sel & 2) hi = (sel & 4) ? 0xffffffffull : 0;
    b = (hi << 32) | lo;
  }
  if ((i & 63) == 5) { a = edges[(i >> 6) % 12]; b = edges[(i >> 6) / 12 % 12]; }
  unsigned bad = 0;
  bool sus = false;
  const u64 v = V::mul(a, b, sus);
  if (canon(v) != mul_c(a, b)) bad++;
  if (!sus && v > P) bad++;
  const u64 bc = canon(b);
  if (canon(V::add(a, bc)) != canon(add_c(a, bc))) bad++;
  if (canon(V::sub(a, bc)) != canon(sub_c(a, bc))) bad++;
  if (canon(V::add(a, P)) != canon(a)) bad++;
  if (canon(V::sub(a, P)) != canon(a)) bad++;
  if (bad) atomicAdd(mism, (unsigned long long)bad);
}

template <class V>
static void run(const char* name, u64* out, unsigned long long* d_m, int sms) {
  unsigned long long h = 0;
  cudaMemset(d_m, 0, 8);
  const u64 n = 1ull << 24;
  k_check<V><<<(unsigned)(n / 256), 256>>>(n, 99, d_m);
  cudaMemcpy(&h, d_m, 8, cudaMemcpyDeviceToHost);
  const int iters = 1024;
  for (int ctas = 4; ctas <= 8; ctas *= 2) {
    const int blocks = sms * ctas;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k_bfly<V, 8><<<blocks, 256>>>(out, 12345, iters);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k_bfly<V, 8><<<blocks, 256>>>(out, 12345, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double warp_bf = (double)blocks * 8 * iters * 8;
    printf("%-40s mismatches %llu  ctas/SM %d  %.1f SMSP-cycles per warp-butterfly\n", name, h, ctas, (ms * 1e-3) * 1.965e9 * sms * 4 / warp_bf);
  }
}

int main() {
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  const int sms = pr.multiProcessorCount;
  u64* out; cudaMalloc(&out, sizeof(u64) * 256 * sms * 8);
  unsigned long long* d_m; cudaMalloc(&d_m, 8);
  run<V0>("V0 shipped (IMAD.HI fixes, canonical)", out, d_m, sms);
  run<V1>("V1 add/sub chains, subc-after-add mask", out, d_m, sms);
  run<V2>("V2 add/sub chains, addc+neg mask", out, d_m, sms);
  run<V3>("V3 plain C", out, d_m, sms);
  run<V4>("V4 V1 with r2*EPS as IMAD.WIDE", out, d_m, sms);
  return 0;
}
