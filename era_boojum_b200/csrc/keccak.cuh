// Keccak-f[1600] and Keccak-256 (original padding 0x01 .. 0x80, rate 136 bytes = 17 lanes), shared by the device Merkle
// kernels and the host transcript.  The reference uses crate sha3's `Keccak256` (Cargo.toml: sha3_ce) as its third
// TreeHasher and transcript: src/cs/oracle/mod.rs:247-313, src/cs/implementations/transcript.rs:262-367.
#pragma once
#include <stdint.h>
#include <string.h>
#include "gl64.cuh"

namespace bj {

__host__ __device__ __forceinline__ uint64_t keccak_rol(uint64_t v, int r) { return (v << r) | (v >> (64 - r)); }

// lanes a[x + 5 y]
__host__ __device__ inline void keccak_f1600(uint64_t (&a)[25]) {
  constexpr uint64_t KECCAK_RC[24] = {
      0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808Aull, 0x8000000080008000ull, 0x000000000000808Bull,
      0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull, 0x000000000000008Aull, 0x0000000000000088ull,
      0x0000000080008009ull, 0x000000008000000Aull, 0x000000008000808Bull, 0x800000000000008Bull, 0x8000000000008089ull,
      0x8000000000008003ull, 0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800Aull, 0x800000008000000Aull,
      0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
#pragma unroll 1
  for (int round = 0; round < 24; round++) {
    uint64_t c[5], d;
#pragma unroll
    for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
#pragma unroll
    for (int x = 0; x < 5; x++) {
      d = c[(x + 4) % 5] ^ keccak_rol(c[(x + 1) % 5], 1);
#pragma unroll
      for (int y = 0; y < 25; y += 5) a[x + y] ^= d;
    }
    // rho + pi along the single 24-cycle of the lane permutation
    uint64_t cur = a[1];
#define BJ_KECCAK_RP(dst, rot)        \
  {                                   \
    const uint64_t t = a[dst];        \
    a[dst] = keccak_rol(cur, rot);    \
    cur = t;                          \
  }
    BJ_KECCAK_RP(10, 1) BJ_KECCAK_RP(7, 3) BJ_KECCAK_RP(11, 6) BJ_KECCAK_RP(17, 10) BJ_KECCAK_RP(18, 15) BJ_KECCAK_RP(3, 21)
    BJ_KECCAK_RP(5, 28) BJ_KECCAK_RP(16, 36) BJ_KECCAK_RP(8, 45) BJ_KECCAK_RP(21, 55) BJ_KECCAK_RP(24, 2) BJ_KECCAK_RP(4, 14)
    BJ_KECCAK_RP(15, 27) BJ_KECCAK_RP(23, 41) BJ_KECCAK_RP(19, 56) BJ_KECCAK_RP(13, 8) BJ_KECCAK_RP(12, 25) BJ_KECCAK_RP(2, 43)
    BJ_KECCAK_RP(20, 62) BJ_KECCAK_RP(14, 18) BJ_KECCAK_RP(22, 39) BJ_KECCAK_RP(9, 61) BJ_KECCAK_RP(6, 20) BJ_KECCAK_RP(1, 44)
#undef BJ_KECCAK_RP
#pragma unroll
    for (int y = 0; y < 25; y += 5) {
      const uint64_t b0 = a[y], b1 = a[y + 1], b2 = a[y + 2], b3 = a[y + 3], b4 = a[y + 4];
      a[y] = b0 ^ (~b1 & b2);
      a[y + 1] = b1 ^ (~b2 & b3);
      a[y + 2] = b2 ^ (~b3 & b4);
      a[y + 3] = b3 ^ (~b4 & b0);
      a[y + 4] = b4 ^ (~b0 & b1);
    }
    a[0] ^= KECCAK_RC[round];
  }
}

// streaming Keccak-256 on the host (transcript, self-test hook)
struct HostKeccak256 {
  uint64_t st[25];
  uint8_t buf[136];
  size_t buf_len = 0;
  HostKeccak256() { reset(); }
  void reset() {
    memset(st, 0, sizeof(st));
    buf_len = 0;
  }
  void absorb_block() {
    for (int i = 0; i < 17; i++) {
      uint64_t w = 0;
      for (int k = 0; k < 8; k++) w |= (uint64_t)buf[8 * i + k] << (8 * k);
      st[i] ^= w;
    }
    keccak_f1600(st);
    buf_len = 0;
  }
  void update(const uint8_t* data, size_t n) {
    while (n) {
      const size_t take = (136 - buf_len) < n ? (136 - buf_len) : n;
      memcpy(buf + buf_len, data, take);
      buf_len += take;
      data += take;
      n -= take;
      if (buf_len == 136) absorb_block();
    }
  }
  void finalize_reset(uint8_t out[32]) {
    memset(buf + buf_len, 0, 136 - buf_len);
    buf[buf_len] |= 0x01;
    buf[135] |= 0x80;
    buf_len = 136;
    absorb_block();
    for (int i = 0; i < 4; i++)
      for (int k = 0; k < 8; k++) out[8 * i + k] = (uint8_t)(st[i] >> (8 * k));
    reset();
  }
};

}  // namespace bj
