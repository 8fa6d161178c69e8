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

// "Constants added after linear layer k": the round constants of round r are folded into the linear layer that ends
// round r-1 (layer 0 is the initial M_E).  Layers are numbered 0..30 in execution order: layer 0 = initial external,
// layers 1..4 = after full rounds 0..3, layers 5..26 = after partial rounds 4..25, layers 27..30 = after full rounds
// 26..29.  Row k holds the 12 constants to add (zeros where the next round takes none), so every layer has one code shape.
__constant__ u64 c_poseidon_layer_rc[31 * 12];

struct P2LayerConstants {
  u64 v[31 * 12];
};
inline P2LayerConstants make_p2_layer_constants() {
  P2LayerConstants t;
  for (int k = 0; k < 31; k++) {
    const int next_round = k;  // layer k is followed by round k (k = 30: nothing follows)
    for (int i = 0; i < 12; i++) {
      u64 c = 0;
      if (next_round < 30) {
        const bool full = next_round < 4 || next_round >= 26;
        if (full || i == 0) c = BJ_POSEIDON_RC_HOST[next_round * 12 + i];
      }
      t.v[k * 12 + i] = c;
    }
  }
  return t;
}
inline const P2LayerConstants& p2_layer_constants_host() {
  static const P2LayerConstants t = make_p2_layer_constants();
  return t;
}

#ifdef __CUDA_ARCH__
#define BJ_P2_LRC(i) c_poseidon_layer_rc[(i)]
#else
#define BJ_P2_LRC(i) p2_layer_constants_host().v[(i)]
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

// s <- circ(2 M4, M4, M4) s + constants of layer `layer`
__host__ __device__ __forceinline__ void p2_external(u64 (&s)[12], int layer) {
  gl::w96 b[3][4];
  p2_m4_wide(s[0], s[1], s[2], s[3], b[0]);
  p2_m4_wide(s[4], s[5], s[6], s[7], b[1]);
  p2_m4_wide(s[8], s[9], s[10], s[11], b[2]);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const gl::w96 sum = gl::w96_add(gl::w96_add(b[0][i], b[1][i]), b[2][i]);
#pragma unroll
    for (int k = 0; k < 3; k++)
      s[4 * k + i] = gl::w96_reduce(gl::w96_add64(gl::w96_add(b[k][i], sum), BJ_P2_LRC(layer * 12 + 4 * k + i)));
  }
}

__host__ __device__ __forceinline__ u64 p2_pow7(u64 x) {
  const u64 x2 = gl::mul_lazy(x, x), x3 = gl::mul_lazy(x2, x), x4 = gl::mul_lazy(x2, x2);
  return gl::mul_lazy(x4, x3);
}

// s <- (diag(2^sh) + J) s + constants of layer `layer`; `x7` = new s[0] (S-box output), rest_sum = s[1] + ... + s[11]
// (computed by the caller before the S-box so that it is off the critical path)
// ALL12 = false: only element 0 takes a constant (the layer is followed by another partial round: 21 of the 22 internal
// layers), which drops 11 constant loads + 96-bit additions per layer; ALL12 = true: the layer before the first closing full
// round (layer 26) adds all twelve.
template <bool ALL12>
__host__ __device__ __forceinline__ void p2_internal(u64 (&s)[12], gl::w96 rest_sum, int layer) {
  constexpr unsigned SH[12] = {4, 14, 11, 8, 0, 5, 2, 9, 13, 6, 3, 12};
  const u64 sr = gl::w96_reduce(gl::w96_add64(rest_sum, s[0]));
#pragma unroll
  for (int i = 0; i < 12; i++) {
    gl::w96 t = gl::w96_add64(gl::w96_from_shl(s[i], SH[i]), sr);
    if (ALL12 || i == 0) t = gl::w96_add64(t, BJ_P2_LRC(layer * 12 + i));
    s[i] = gl::w96_reduce(t);
  }
}

// 30 rounds (4 full, 22 partial, 4 full; state_generic_impl.rs:219-233) as ONE loop with two bodies so that the code
// stays inside the instruction cache; layer k's constants are those of round k.
__host__ __device__ __forceinline__ void poseidon2_permutation(u64 (&s)[12]) {
  p2_external(s, 0);
#pragma unroll 1
  for (int r = 0; r < 30; r++) {
    if (r < 4 || r >= 26) {
#pragma unroll
      for (int i = 0; i < 12; i++) s[i] = p2_pow7(s[i]);
      p2_external(s, r + 1);
    } else {
      gl::w96 rest = gl::w96_add64(gl::w96_from(s[1]), s[2]);
      gl::w96 rest2 = gl::w96_add64(gl::w96_from(s[3]), s[4]);
      gl::w96 rest3 = gl::w96_add64(gl::w96_from(s[5]), s[6]);
      gl::w96 rest4 = gl::w96_add64(gl::w96_from(s[7]), s[8]);
      rest = gl::w96_add64(gl::w96_add64(rest, s[9]), s[10]);
      rest2 = gl::w96_add64(gl::w96_add(rest2, rest3), s[11]);
      rest = gl::w96_add(gl::w96_add(rest, rest2), rest4);
      s[0] = p2_pow7(s[0]);
      if (r == 25) p2_internal<true>(s, rest, r + 1);
      else p2_internal<false>(s, rest, r + 1);
    }
  }
}

}  // namespace bj
