// Host-side Fiat-Shamir pieces of the path (they are sequential by nature and must stay on the host):
//   AlgebraicSpongeBasedTranscript<_, 8, 12, 4, Poseidon2, AbsorptionModeOverwrite> = GoldilocksPoisedon2Transcript
//                                                      src/cs/implementations/transcript.rs:62-129, 140-151
//   BoolsBuffer::get_bits (query index bits)           src/cs/implementations/transcript.rs:369-417
//   compute_fri_schedule                               src/cs/implementations/prover.rs:2281-2372
// The permutation is the same poseidon2.cuh source the kernels use, compiled for the host.
#include <vector>
#include "ctx.hpp"
#include "poseidon2.cuh"

struct bj_transcript {
  std::vector<bj::u64> buffer;
  std::vector<bj::u64> available;  // unread challenges, front first
  size_t avail_pos = 0;
  bj::u64 state[12] = {0};
  // BoolsBuffer
  std::vector<uint8_t> bits;
  size_t bits_pos = 0;
};

namespace bj {
static void transcript_refill(bj_transcript* t) {
  t->available.assign(t->state, t->state + 8);
  for (auto& v : t->available) v = gl::canon(v);
  t->avail_pos = 0;
}
}  // namespace bj

using namespace bj;

extern "C" {

bj_transcript* bj_transcript_new(void) { return new bj_transcript(); }
void bj_transcript_free(bj_transcript* t) { delete t; }

void bj_transcript_witness_field_elements(bj_transcript* t, const uint64_t* els, size_t n) {
  if (!t || (!els && n)) return;
  for (size_t i = 0; i < n; i++) t->buffer.push_back(gl::canon(els[i]));
}

void bj_transcript_witness_merkle_tree_cap(bj_transcript* t, const uint64_t* cap, size_t n_digests) {
  bj_transcript_witness_field_elements(t, cap, 4 * n_digests);
}

uint64_t bj_transcript_get_challenge(bj_transcript* t) {
  if (!t) return 0;
  if (t->buffer.empty()) {
    if (t->avail_pos < t->available.size()) return t->available[t->avail_pos++];
    poseidon2_permutation(t->state);  // run_round_function, then take the 8 rate elements
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
    poseidon2_permutation(t->state);
  }
  transcript_refill(t);
  return t->available[t->avail_pos++];
}

// BoolsBuffer::get_bits for an algebraic transcript: each refill keeps the 64 - max_needed low bits of one challenge
uint64_t bj_transcript_get_index_bits(bj_transcript* t, uint32_t num_bits, uint32_t max_needed) {
  if (!t || num_bits > 64 || max_needed >= 64) return 0;
  while (t->bits.size() - t->bits_pos < num_bits) {
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
