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

// All state values are LAZY (any u64 congruent mod p) throughout the permutation: the linear layers sum in 96-bit
// accumulators (plain carry chains, one 2^64 = 2^32 - 1 fold per output), the S-box uses mul_lazy; callers canonicalise
// the words they export.

// M4 = [[5,7,1,3],[4,6,1,1],[1,3,5,7],[1,1,4,6]] applied to one 4-block, result kept wide (coefficients sum to <= 16)
__host__ __device__ __forceinline__ void p2_m4_wide(u64 x0, u64 x1, u64 x2, u64 x3, gl::w96 (&y)[4]) {
  using gl::w96;
  const w96 t0 = gl::w96_add64(gl::w96_from(x0), x1);            // x0 + x1
  const w96 t1 = gl::w96_add64(gl::w96_from(x2), x3);            // x2 + x3
  const w96 t2 = gl::w96_add(gl::w96_shl(gl::w96_from(x1), 1), t1);  // 2 x1 + t1
  const w96 t3 = gl::w96_add(gl::w96_shl(gl::w96_from(x3), 1), t0);  // 2 x3 + t0
  const w96 t4 = gl::w96_add(gl::w96_shl(t1, 2), t3);            // 4 t1 + t3
  const w96 t5 = gl::w96_add(gl::w96_shl(t0, 2), t2);            // 4 t0 + t2
  y[0] = gl::w96_add(t3, t5);
  y[1] = t5;
  y[2] = gl::w96_add(t2, t4);
  y[3] = t4;
}

// s <- circ(2 M4, M4, M4) s  (+ optional round constants of the NEXT full round folded into the same reduction)
template <bool ADD_RC>
__host__ __device__ __forceinline__ void p2_external(u64 (&s)[12], int rc_base) {
  gl::w96 b[3][4];
  p2_m4_wide(s[0], s[1], s[2], s[3], b[0]);
  p2_m4_wide(s[4], s[5], s[6], s[7], b[1]);
  p2_m4_wide(s[8], s[9], s[10], s[11], b[2]);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const gl::w96 sum = gl::w96_add(gl::w96_add(b[0][i], b[1][i]), b[2][i]);
#pragma unroll
    for (int k = 0; k < 3; k++) {
      gl::w96 v = gl::w96_add(b[k][i], sum);
      if (ADD_RC) v = gl::w96_add64(v, BJ_P2_RC(rc_base + 4 * k + i));
      s[4 * k + i] = gl::w96_reduce(v);
    }
  }
}

__host__ __device__ __forceinline__ u64 p2_pow7(u64 x) {
  const u64 x2 = gl::mul_lazy(x, x), x3 = gl::mul_lazy(x2, x), x4 = gl::mul_lazy(x2, x2);
  return gl::mul_lazy(x4, x3);
}

// s <- (diag(2^sh) + J) s, optionally adding the next partial round's constant to s[0]
template <bool ADD_RC>
__host__ __device__ __forceinline__ void p2_internal(u64 (&s)[12], int rc_idx) {
  constexpr unsigned SH[12] = {4, 14, 11, 8, 0, 5, 2, 9, 13, 6, 3, 12};
  gl::w96 sum = gl::w96_from(s[0]);
#pragma unroll
  for (int i = 1; i < 12; i++) sum = gl::w96_add64(sum, s[i]);
  const u64 sr = gl::w96_reduce(sum);
#pragma unroll
  for (int i = 0; i < 12; i++) {
    gl::w96 v = gl::w96_add64(gl::w96_from_shl(s[i], SH[i]), sr);
    if (ADD_RC && i == 0) v = gl::w96_add64(v, BJ_P2_RC(rc_idx));
    s[i] = gl::w96_reduce(v);
  }
}

// rounds r = 0..29 with one running constant index (state_generic_impl.rs:219-233); the constants of round r are added
// inside the linear layer that ends round r-1
__host__ __device__ __forceinline__ void poseidon2_permutation(u64 (&s)[12]) {
  p2_external<true>(s, 0);  // initial M_E, + RC of full round 0
#pragma unroll 1
  for (int r = 0; r < 4; r++) {
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = p2_pow7(s[i]);
    if (r < 3) {
      p2_external<true>(s, (r + 1) * 12);
    } else {
      // next is partial round 4: only s[0] receives a constant
      p2_external<false>(s, 0);
      s[0] = gl::w96_reduce(gl::w96_add64(gl::w96_from(s[0]), BJ_P2_RC(4 * 12)));
    }
  }
#pragma unroll 1
  for (int r = 4; r < 26; r++) {
    s[0] = p2_pow7(s[0]);
    if (r < 25) {
      p2_internal<true>(s, (r + 1) * 12);
    } else {
      // next is full round 26: every word receives a constant
      p2_internal<false>(s, 0);
#pragma unroll
      for (int i = 0; i < 12; i++) s[i] = gl::w96_reduce(gl::w96_add64(gl::w96_from(s[i]), BJ_P2_RC(26 * 12 + i)));
    }
  }
#pragma unroll 1
  for (int r = 26; r < 30; r++) {
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = p2_pow7(s[i]);
    if (r < 29) p2_external<true>(s, (r + 1) * 12);
    else p2_external<false>(s, 0);
  }
}

}  // namespace bj
