// Host-side Fiat-Shamir pieces of the path (they are sequential by nature and must stay on the host):
//   AlgebraicSpongeBasedTranscript<_, 8, 12, 4, Poseidon2, AbsorptionModeOverwrite> = GoldilocksPoisedon2Transcript
//                                                      src/cs/implementations/transcript.rs:62-129, 140-151
//   Blake2sTranscript                                  src/cs/implementations/transcript.rs:155-260
//   BoolsBuffer::get_bits (query index bits)           src/cs/implementations/transcript.rs:369-417
//   compute_fri_schedule                               src/cs/implementations/prover.rs:2281-2372
//   GoldilocksPoisedonTranscript = the same sponge transcript over the Poseidon (v1) permutation, the transcript of
//     run_sha256_prover_recursive_mode[_poseidon2] (src/gadgets/sha256/mod.rs:275-293; transcript.rs:131-138;
//     permutation src/implementations/poseidon_goldilocks_naive.rs:11-165)
// The Poseidon2 permutation is the same poseidon2.cuh source the kernels use, compiled for the host.
#include <cstring>
#include <vector>
#include "ctx.hpp"
#include "poseidon2.cuh"
#include "keccak.cuh"

namespace bj {
// Streaming Blake2s-256 on the host (RFC 7693, unkeyed, 32-byte digest) for Blake2sTranscript.
struct HostBlake2s {
  uint32_t h[8];
  uint64_t t = 0;
  uint8_t buf[64];
  size_t buf_len = 0;
  HostBlake2s() { reset(); }
  static uint32_t rotr(uint32_t x, int r) { return (x >> r) | (x << (32 - r)); }
  void reset() {
    static const uint32_t iv[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au, 0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
    for (int i = 0; i < 8; i++) h[i] = iv[i];
    h[0] ^= 0x01010020u;
    t = 0;
    buf_len = 0;
  }
  void compress(const uint8_t* block, bool last) {
    static const uint32_t iv[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au, 0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
    static const uint8_t S[10][16] = {
        {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
        {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
        {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
        {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
        {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};
    uint32_t m[16], v[16];
    for (int i = 0; i < 16; i++)
      m[i] = (uint32_t)block[4 * i] | ((uint32_t)block[4 * i + 1] << 8) | ((uint32_t)block[4 * i + 2] << 16) | ((uint32_t)block[4 * i + 3] << 24);
    for (int i = 0; i < 8; i++) {
      v[i] = h[i];
      v[8 + i] = iv[i];
    }
    v[12] ^= (uint32_t)t;
    v[13] ^= (uint32_t)(t >> 32);
    if (last) v[14] = ~v[14];
    auto G = [&](int a, int b, int c, int d, uint32_t x, uint32_t y) {
      v[a] = v[a] + v[b] + x;
      v[d] = rotr(v[d] ^ v[a], 16);
      v[c] = v[c] + v[d];
      v[b] = rotr(v[b] ^ v[c], 12);
      v[a] = v[a] + v[b] + y;
      v[d] = rotr(v[d] ^ v[a], 8);
      v[c] = v[c] + v[d];
      v[b] = rotr(v[b] ^ v[c], 7);
    };
    for (int r = 0; r < 10; r++) {
      G(0, 4, 8, 12, m[S[r][0]], m[S[r][1]]);
      G(1, 5, 9, 13, m[S[r][2]], m[S[r][3]]);
      G(2, 6, 10, 14, m[S[r][4]], m[S[r][5]]);
      G(3, 7, 11, 15, m[S[r][6]], m[S[r][7]]);
      G(0, 5, 10, 15, m[S[r][8]], m[S[r][9]]);
      G(1, 6, 11, 12, m[S[r][10]], m[S[r][11]]);
      G(2, 7, 8, 13, m[S[r][12]], m[S[r][13]]);
      G(3, 4, 9, 14, m[S[r][14]], m[S[r][15]]);
    }
    for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[8 + i];
  }
  void update(const uint8_t* data, size_t n) {
    while (n) {
      if (buf_len == 64) {  // a full buffer is compressed only when more data follows (the last block needs the flag)
        t += 64;
        compress(buf, false);
        buf_len = 0;
      }
      const size_t take = std::min<size_t>(64 - buf_len, n);
      memcpy(buf + buf_len, data, take);
      buf_len += take;
      data += take;
      n -= take;
    }
  }
  // digest of everything fed so far, then back to the initial state (blake2::Digest::finalize_reset)
  void finalize_reset(uint8_t out[32]) {
    t += buf_len;
    memset(buf + buf_len, 0, 64 - buf_len);
    compress(buf, true);
    for (int i = 0; i < 8; i++)
      for (int k = 0; k < 4; k++) out[4 * i + k] = (uint8_t)(h[i] >> (8 * k));
    reset();
  }
};
}  // namespace bj

namespace bj {
// Poseidon (v1) over Goldilocks, t = 12, x^7, 4 + 22 + 4 rounds (src/implementations/poseidon_goldilocks_naive.rs:123-165): every
// round adds its 12 round constants, applies the S-box to the whole state (full rounds) or to element 0 (partial rounds) and
// multiplies by the circulant MDS matrix with entries 2^MDS_MATRIX_EXPS[(column - row) mod 12] (:11-31, 67-90).
static void host_poseidon_v1_permutation(u64 state[12]) {
  static const int EXPS[12] = {0, 0, 1, 0, 3, 5, 1, 8, 12, 3, 16, 10};
  u64 s[12];
  for (int i = 0; i < 12; i++) s[i] = gl::canon(state[i]);
  for (int round = 0; round < 30; round++) {
    for (int i = 0; i < 12; i++) s[i] = gl::canon(gl::add_lazy(s[i], gl::canon(BJ_POSEIDON_RC_HOST[round * 12 + i])));
    const int n_sbox = (round < 4 || round >= 26) ? 12 : 1;
    for (int i = 0; i < n_sbox; i++) {
      const u64 x = s[i], x2 = gl::mul(x, x), x3 = gl::mul(x2, x), x4 = gl::mul(x2, x2);
      s[i] = gl::mul(x4, x3);
    }
    u64 r[12];
    for (int row = 0; row < 12; row++) {
      unsigned __int128 acc = 0;  // 12 terms below 2^(64+16): no overflow (MAX_ROW_VALUE_BITS == 81, :34-58)
      for (int col = 0; col < 12; col++) acc += (unsigned __int128)s[col] << EXPS[(col + 12 - row) % 12];
      // reduce hi * 2^64 + lo with 2^64 = 2^32 - 1 (hi < 2^17)
      const u64 lo = (u64)acc, hi = (u64)(acc >> 64);
      r[row] = gl::canon(gl::add_lazy(gl::canon(lo), gl::mul(hi, 0xFFFFFFFFull)));
    }
    for (int i = 0; i < 12; i++) s[i] = r[i];
  }
  for (int i = 0; i < 12; i++) state[i] = s[i];
}
}  // namespace bj

struct bj_transcript {
  // 0: Poseidon2 sponge, 1: Blake2sTranscript (transcript.rs:155-260), 2: Keccak256Transcript (:262-367),
  // 3: Poseidon (v1) sponge = GoldilocksPoisedonTranscript (:131-138)
  int kind = 0;
  bool algebraic() const { return kind == 0 || kind == 3; }
  void permute() {
    if (kind == 3) bj::host_poseidon_v1_permutation(state);
    else bj::poseidon2_permutation(state);
  }
  bj::HostBlake2s b2s;
  bj::HostKeccak256 keccak;
  std::vector<uint8_t> byte_buffer, byte_available;
  size_t byte_pos = 0;
  std::vector<bj::u64> buffer;
  std::vector<bj::u64> available;  // unread challenges, front first
  size_t avail_pos = 0;
  bj::u64 state[12] = {0};
  // BoolsBuffer
  std::vector<uint8_t> bits;
  size_t bits_pos = 0;
};

namespace bj {
// Blake2sTranscript: absorb the pending bytes, re-seed the hasher with its own output, expose the 32 output bytes
static void byte_hash_update(bj_transcript* t, const uint8_t* d, size_t n) {
  if (t->kind == 2) t->keccak.update(d, n);
  else t->b2s.update(d, n);
}
static void b2s_reseed(bj_transcript* t, bool keep_leftover) {
  uint8_t out[32];
  if (t->kind == 2) t->keccak.finalize_reset(out);
  else t->b2s.finalize_reset(out);
  byte_hash_update(t, out, 32);
  if (!keep_leftover) {
    t->byte_available.clear();
    t->byte_pos = 0;
  }
  t->byte_available.insert(t->byte_available.end(), out, out + 32);
}
static void b2s_challenge_bytes(bj_transcript* t, size_t num, uint8_t* dst) {
  if (!t->byte_buffer.empty()) {
    byte_hash_update(t, t->byte_buffer.data(), t->byte_buffer.size());
    t->byte_buffer.clear();
    b2s_reseed(t, false);
  }
  while (t->byte_available.size() - t->byte_pos < num) b2s_reseed(t, true);
  memcpy(dst, t->byte_available.data() + t->byte_pos, num);
  t->byte_pos += num;
  if (t->byte_pos == t->byte_available.size()) {
    t->byte_available.clear();
    t->byte_pos = 0;
  }
}
static u64 le64(const uint8_t* b) {
  u64 v = 0;
  for (int k = 0; k < 8; k++) v |= (u64)b[k] << (8 * k);
  return v;
}

static void transcript_refill(bj_transcript* t) {
  t->available.assign(t->state, t->state + 8);
  for (auto& v : t->available) v = gl::canon(v);
  t->avail_pos = 0;
}
}  // namespace bj

using namespace bj;

extern "C" {

bj_transcript* bj_transcript_new(void) { return new bj_transcript(); }
bj_transcript* bj_transcript_new_blake2s(void) {
  bj_transcript* t = new bj_transcript();
  t->kind = 1;
  return t;
}
bj_transcript* bj_transcript_new_keccak256(void) {
  bj_transcript* t = new bj_transcript();
  t->kind = 2;
  return t;
}
bj_transcript* bj_transcript_new_poseidon(void) {
  bj_transcript* t = new bj_transcript();
  t->kind = 3;
  return t;
}
void bj_host_poseidon_permutation(uint64_t state[12]) { bj::host_poseidon_v1_permutation((bj::u64*)state); }
void bj_transcript_free(bj_transcript* t) { delete t; }

void bj_transcript_witness_field_elements(bj_transcript* t, const uint64_t* els, size_t n) {
  if (!t || (!els && n)) return;
  if (!t->algebraic()) {  // el.as_u64_reduced().to_le_bytes()
    for (size_t i = 0; i < n; i++) {
      const u64 v = gl::canon(els[i]);
      for (int k = 0; k < 8; k++) t->byte_buffer.push_back((uint8_t)(v >> (8 * k)));
    }
    return;
  }
  for (size_t i = 0; i < n; i++) t->buffer.push_back(gl::canon(els[i]));
}

void bj_transcript_witness_merkle_tree_cap(bj_transcript* t, const uint64_t* cap, size_t n_digests) {
  if (t && !t->algebraic()) {  // caps are raw 32-byte digests (4 little-endian u64 each), not field elements
    if (!cap && n_digests) return;
    for (size_t i = 0; i < 4 * n_digests; i++)
      for (int k = 0; k < 8; k++) t->byte_buffer.push_back((uint8_t)(cap[i] >> (8 * k)));
    return;
  }
  bj_transcript_witness_field_elements(t, cap, 4 * n_digests);
}

uint64_t bj_transcript_get_challenge(bj_transcript* t) {
  if (!t) return 0;
  if (!t->algebraic()) {  // 8 challenge bytes, little endian, reduced (from_u64_with_reduction)
    uint8_t b[8];
    b2s_challenge_bytes(t, 8, b);
    return gl::canon(le64(b));
  }
  if (t->buffer.empty()) {
    if (t->avail_pos < t->available.size()) return t->available[t->avail_pos++];
    t->permute();  // run_round_function, then take the 8 rate elements
    transcript_refill(t);
    return t->available[t->avail_pos++];
  }
  // pad with 1, 0.. to a multiple of the rate, absorb in overwrite mode
  std::vector<u64> to_absorb;
  to_absorb.swap(t->buffer);
  to_absorb.push_back(1);
  while (to_absorb.size() % 8) to_absorb.push_back(0);
  for (size_t i = 0; i < to_absorb.size(); i += 8) {
    for (int k = 0; k < 8; k++) t->state[k] = to_absorb[i + k];
    t->permute();
  }
  transcript_refill(t);
  return t->available[t->avail_pos++];
}

// BoolsBuffer::get_bits for an algebraic transcript: each refill keeps the 64 - max_needed low bits of one challenge
uint64_t bj_transcript_get_index_bits(bj_transcript* t, uint32_t num_bits, uint32_t max_needed) {
  if (!t || num_bits > 64 || max_needed >= 64) return 0;
  while (t->bits.size() - t->bits_pos < num_bits) {
    if (!t->algebraic()) {  // non-algebraic transcript: 8 uniform bytes, all 64 bits (transcript.rs:401-413)
      uint8_t bb[8];
      b2s_challenge_bytes(t, 8, bb);
      const u64 el = le64(bb);
      for (uint32_t b = 0; b < 64; b++) t->bits.push_back((uint8_t)((el >> b) & 1));
      continue;
    }
    const u64 el = bj_transcript_get_challenge(t);
    for (uint32_t b = 0; b < 64 - max_needed; b++) t->bits.push_back((uint8_t)((el >> b) & 1));
  }
  u64 v = 0;
  for (uint32_t i = 0; i < num_bits; i++) v |= (u64)t->bits[t->bits_pos + i] << i;  // LSB first
  t->bits_pos += num_bits;
  if (t->bits_pos > 4096) {
    t->bits.erase(t->bits.begin(), t->bits.begin() + t->bits_pos);
    t->bits_pos = 0;
  }
  return v;
}

int32_t bj_compute_fri_schedule(uint32_t security_bits, uint32_t cap_size, uint32_t pow_bits, uint32_t rate_log_two,
                                uint32_t initial_degree_log_two, uint32_t* new_pow_bits, uint32_t* num_queries,
                                uint32_t* schedule, uint32_t* schedule_len, uint32_t* final_degree) {
  if (!new_pow_bits || !num_queries || !schedule || !schedule_len || !final_degree || rate_log_two == 0 ||
      security_bits <= pow_bits || cap_size == 0 || (cap_size & (cap_size - 1)))
    return BJ_ERR_INVALID_ARG;
  uint32_t raw = security_bits - pow_bits, np = pow_bits;
  if (raw % rate_log_two != 0) {
    if (np >= rate_log_two - (raw % rate_log_two)) np -= rate_log_two - (raw % rate_log_two);
  }
  raw = security_bits - np;
  uint32_t nq = raw / rate_log_two + ((raw % rate_log_two) ? 1 : 0);
  uint32_t stop = cap_size >> rate_log_two;
  if (stop < 1) stop = 1;
  uint32_t stop_log = 0;
  while ((1u << stop_log) < stop) stop_log++;
  uint32_t cap_log = 0;
  while ((1u << cap_log) < cap_size) cap_log++;
  uint32_t deg = initial_degree_log_two, n = 0;
  while (deg > stop_log) {
    if (deg + rate_log_two <= cap_log) break;
    if (n >= 32) return BJ_ERR_INVALID_ARG;
    if (deg - stop_log >= 3) {
      deg -= 3;
      schedule[n++] = 3;
    } else if (deg - stop_log == 2) {
      deg -= 2;
      schedule[n++] = 2;
    } else {
      deg -= 1;
      schedule[n++] = 1;
      break;
    }
    if (deg + rate_log_two <= cap_log) break;
  }
  *new_pow_bits = np;
  *num_queries = nq;
  *schedule_len = n;
  *final_degree = 1u << deg;
  return BJ_OK;
}

}  // extern "C"
