// Poseidon2-Goldilocks permutation (t = 12, x^7, 4 + 22 + 4 rounds) shared by device kernels and host code
// (transcript).  Parameters per the reference: src/implementations/poseidon2/state_generic_impl.rs:158-233
// (round structure, single round counter 0..29), :69-82 (internal diagonal 2^{4,14,11,8,0,5,2,9,13,6,3,12}),
// src/implementations/suggested_mds.rs:8-14,59-97 (external matrix circ(2 M4, M4, M4)).
// The state lives in 12 registers; every loop is fully unrolled.
#pragma once
#include "gl64.cuh"
#include "poseidon_rc.h"

namespace bj {
using gl::u64;

__constant__ u64 c_poseidon_rc[360];

#ifdef __CUDA_ARCH__
#define BJ_P2_RC(i) c_poseidon_rc[(i)]
#else
#define BJ_P2_RC(i) BJ_POSEIDON_RC_HOST[(i)]
#endif

// y = M4 x with M4 = [[5,7,1,3],[4,6,1,1],[1,3,5,7],[1,1,4,6]]; inputs lazy, outputs lazy
__host__ __device__ __forceinline__ void p2_m4(u64& x0, u64& x1, u64& x2, u64& x3) {
  const u64 a0 = gl::canon(x0), a1 = gl::canon(x1), a2 = gl::canon(x2), a3 = gl::canon(x3);
  const u64 t0 = gl::canon(gl::add(a0, a1));                 // x0 + x1
  const u64 t1 = gl::canon(gl::add(a2, a3));                 // x2 + x3
  const u64 t2 = gl::canon(gl::add(gl::add(a1, a1), t1));    // 2 x1 + t1
  const u64 t3 = gl::canon(gl::add(gl::add(a3, a3), t0));    // 2 x3 + t0
  const u64 t1_2 = gl::canon(gl::add(t1, t1));
  const u64 t0_2 = gl::canon(gl::add(t0, t0));
  const u64 t4 = gl::canon(gl::add(gl::add(t1_2, t1_2), t3));  // 4 t1 + t3
  const u64 t5 = gl::canon(gl::add(gl::add(t0_2, t0_2), t2));  // 4 t0 + t2
  x0 = gl::add(t3, t5);
  x1 = t5;
  x2 = gl::add(t2, t4);
  x3 = t4;
}

// s <- circ(2 M4, M4, M4) s
__host__ __device__ __forceinline__ void p2_external(u64 (&s)[12]) {
  p2_m4(s[0], s[1], s[2], s[3]);
  p2_m4(s[4], s[5], s[6], s[7]);
  p2_m4(s[8], s[9], s[10], s[11]);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const u64 a = gl::canon(s[i]), b = gl::canon(s[4 + i]), c = gl::canon(s[8 + i]);
    const u64 sum = gl::canon(gl::add(gl::add(a, b), c));
    s[i] = gl::add(a, sum);
    s[4 + i] = gl::add(b, sum);
    s[8 + i] = gl::add(c, sum);
  }
}

__host__ __device__ __forceinline__ u64 p2_pow7(u64 x) {
  const u64 x2 = gl::sqr(x), x3 = gl::mul(x2, x), x4 = gl::sqr(x2);
  return gl::mul(x4, x3);  // canonical
}

// s <- (diag(2^sh) + J) s
__host__ __device__ __forceinline__ void p2_internal(u64 (&s)[12]) {
  constexpr unsigned SH[12] = {4, 14, 11, 8, 0, 5, 2, 9, 13, 6, 3, 12};
  u64 sum = gl::canon(s[0]);
#pragma unroll
  for (int i = 1; i < 12; i++) sum = gl::add(sum, gl::canon(s[i]));
  sum = gl::canon(sum);
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = gl::add(gl::mul_pow2(s[i], SH[i]), sum);
}

__host__ __device__ __forceinline__ void poseidon2_permutation(u64 (&s)[12]) {
  p2_external(s);
#pragma unroll 1
  for (int r = 0; r < 4; r++) {
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = p2_pow7(gl::add(s[i], BJ_P2_RC(r * 12 + i)));
    p2_external(s);
  }
#pragma unroll 1
  for (int r = 4; r < 26; r++) {
    s[0] = p2_pow7(gl::add(s[0], BJ_P2_RC(r * 12)));
    p2_internal(s);
  }
#pragma unroll 1
  for (int r = 26; r < 30; r++) {
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = p2_pow7(gl::add(s[i], BJ_P2_RC(r * 12 + i)));
    p2_external(s);
  }
}

}  // namespace bj
