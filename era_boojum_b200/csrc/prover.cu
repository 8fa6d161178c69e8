// Host driver of the five-round IOP: the role of CSReferenceAssembly::prove_cpu_basic
// (src/cs/implementations/prover.rs:153-2269) and of the setup materialisation that feeds it (setup.rs:1093-1255), for
// circuits whose gates live on general-purpose columns, with the optional log-derivative lookup argument over specialised
// columns (table id in a constant column).  Plain host C++ over the device entry points of this library: every heavy step
// is one of the bj_* kernels; the transcript, the FRI schedule, query-index derivation and proof assembly stay on the host
// exactly as in the reference.  The proof is handed back in the reference's serde_json shape (proof.rs:57-143), so that a
// Rust shim can `serde_json::from_str::<Proof<..>>` it.
//
// Round structure (prover.rs line numbers):
//   1  witness LDE + oracle                      :313-353        4  openings at z, z*omega, 0             :1501-1802
//   2  copy-permutation / lookup polys + oracle  :360-554        5  DEEP combination + FRI                :1828-2102
//   3  quotient, interpolation, chunks + oracle  :560-1495       6  queries                               :2161-2266
#include <chrono>
#include <cstring>
#include <memory>
#include <string>
#include <vector>
#include "ctx.hpp"

namespace bj {

// stream-ordered device buffer
struct DevMem {
  bj_ctx* ctx = nullptr;
  u64* p = nullptr;
  DevMem() = default;
  DevMem(const DevMem&) = delete;
  DevMem& operator=(const DevMem&) = delete;
  ~DevMem() { release(); }
  void release() {
    if (p) cudaFreeAsync(p, ctx->stream);
    p = nullptr;
  }
  int32_t alloc(bj_ctx* c, size_t n_u64) {
    release();
    ctx = c;
    const size_t bytes = sizeof(u64) * (n_u64 ? n_u64 : 1);
    const cudaError_t e = c->pool ? cudaMallocFromPoolAsync((void**)&p, bytes, c->pool, c->stream) : cudaMallocAsync((void**)&p, bytes, c->stream);
    if (e != cudaSuccess) {
      cudaGetLastError();
      p = nullptr;
      BJ_FAIL(c, BJ_ERR_OOM, "prover: device allocation failed");
    }
    return BJ_OK;
  }
};

// a Merkle oracle over LDE columns
struct Oracle {
  std::vector<const uint64_t*> cols;
  DevMem leaf_hashes, nodes;
  u64 n_leaves = 0;
  u32 cap_size = 0;
  std::vector<u64> cap;  // host, 4 * cap_size
};

// n_leaves / cap_size are GLOBAL; on a coset-sharded context the tree covers this rank's cosets (n_leaves / world leaves,
// cap_size / world local cap nodes - same depth), and o.cap is the assembled global cap (lde_factor locates the cosets).
static int32_t oracle_build(bj_ctx* ctx, Oracle& o, u64 n_leaves, u32 cap_size, u32 hasher, u32 lde_factor) {
  const u32 world = comm_world(ctx);
  const u32 cap_global = cap_size;
  n_leaves /= world;
  cap_size /= world;
  o.n_leaves = n_leaves;
  o.cap_size = cap_size;
  if (cap_size == 0 || n_leaves < cap_size) BJ_FAIL(ctx, BJ_ERR_INVALID_ARG, "prover: oracle smaller than the cap");
  BJ_TRY(o.leaf_hashes.alloc(ctx, 4 * n_leaves));
  BJ_TRY(o.nodes.alloc(ctx, 4 * (n_leaves - cap_size)));
  BJ_TRY((hasher == BJ_HASHER_BLAKE2S ? bj_merkle_build_blake2s : hasher == BJ_HASHER_KECCAK256 ? bj_merkle_build_keccak256 : bj_merkle_build_poseidon2)(
      ctx, o.cols.data(), (u32)o.cols.size(), n_leaves, 1, cap_size, (uint64_t*)o.leaf_hashes.p, (uint64_t*)o.nodes.p));
  std::vector<u64> local(4 * (size_t)cap_size);
  const u64* src = n_leaves == cap_size ? o.leaf_hashes.p : o.nodes.p + 4 * (n_leaves - 2 * (u64)cap_size);
  BJ_CUDA(ctx, cudaMemcpyAsync(local.data(), src, sizeof(u64) * 4 * cap_size, cudaMemcpyDeviceToHost, ctx->stream));
  BJ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  o.cap.resize(4 * (size_t)cap_global);
  return comm_assemble_cap(ctx, local.data(), cap_global, lde_factor, o.cap.data());
}

// LDE of trace-domain columns (Lagrange values, natural order) onto this context's cosets.  One GPU: bj_lde.  Sharded: the
// two natural partitions meet here - the iNTT is independent per COLUMN, so rank r interpolates the block of
// columns [r * per, (r + 1) * per) only, ONE all-gather hands every rank all the monomials (8 * n * n_cols bytes in total), and
// each rank then evaluates them on its own cosets.  Without it every rank would repeat all the iNTTs (1/9 of the LDE work,
// which stops scaling).  Columns must be contiguous (stride n).
static int32_t lde_columns(bj_ctx* ctx, const uint64_t* d_in, uint64_t* d_out, u32 log_n, u32 log_l, u32 n_cols) {
  const u32 world = comm_world(ctx), rank = comm_rank(ctx);
  const u64 n = 1ull << log_n;
  if (world == 1 || n_cols < 2) return bj_lde(ctx, d_in, n, d_out, log_n, log_l, n_cols, 0);
  // software pipeline over groups of columns: while the all-gather of group g travels over NVLink (auxiliary stream), this
  // rank interpolates its share of group g + 1 and evaluates group g - 1 on its cosets
  const u64 out_stride = (n << log_l) / world;
  const u32 group = std::max<u32>(world, ((n_cols + 3) / 4 + world - 1) / world * world);  // <= 4 groups, a multiple of world
  struct Group {
    u32 c0, cnt, per;
    DevMem mono;
    cudaEvent_t gathered = nullptr;
  };
  std::vector<std::unique_ptr<Group>> groups;
  for (u32 c0 = 0; c0 < n_cols; c0 += group) {
    groups.emplace_back(new Group());
    groups.back()->c0 = c0;
    groups.back()->cnt = std::min(group, n_cols - c0);
    groups.back()->per = (groups.back()->cnt + world - 1) / world;
  }
  auto start = [&](Group& g) -> int32_t {
    BJ_TRY(g.mono.alloc(ctx, (size_t)world * g.per * n));
    const u32 first = std::min(rank * g.per, g.cnt), cnt = std::min(g.per, g.cnt - first);
    u64* mine = g.mono.p + (size_t)rank * g.per * n;
    if (cnt) {
      BJ_CUDA(ctx, cudaMemcpyAsync(mine, d_in + (size_t)(g.c0 + first) * n, sizeof(u64) * cnt * n, cudaMemcpyDeviceToDevice, ctx->stream));
      BJ_TRY(bj_intt_natural_to_natural(ctx, (uint64_t*)mine, log_n, cnt, n, 1));
    }
    if (cnt < g.per) BJ_CUDA(ctx, cudaMemsetAsync(mine + (size_t)cnt * n, 0, sizeof(u64) * (g.per - cnt) * n, ctx->stream));
    // in place: my block is my slot of the gathered array
    return comm_all_gather_overlapped(ctx->comm, mine, g.mono.p, (u64)g.per * n, &g.gathered);
  };
  auto finish = [&](Group& g) -> int32_t {
    BJ_TRY(comm_wait(ctx->comm, g.gathered));
    BJ_TRY(bj_lde(ctx, (const uint64_t*)g.mono.p, n, d_out + (size_t)g.c0 * out_stride, log_n, log_l, g.cnt, 1));
    g.mono.release();
    return BJ_OK;
  };
  for (size_t i = 0; i < groups.size(); i++) {
    BJ_TRY(start(*groups[i]));
    if (i) BJ_TRY(finish(*groups[i - 1]));
  }
  return finish(*groups.back());
}

int32_t copy_permutation_stage2_sharded(bj_ctx* ctx, const uint64_t* const* h_variable_cols, const uint64_t* const* h_sigma_cols, u32 n_cols,
                                        const uint64_t* h_non_residues, gl::e2 beta, gl::e2 gamma, u32 log_n, u32 chunk_size, u64* d_out);  // stage2.cu

struct GateCopy {
  std::vector<bj_gate_relation> relations;
  std::vector<bj_gate_index> writes;
  std::vector<uint8_t> path;
};

static inline void json_u64(std::string& s, u64 v) {  // decimal digits straight into the string (std::to_string allocates)
  char buf[24];
  int k = 24;
  do {
    buf[--k] = (char)('0' + v % 10);
    v /= 10;
  } while (v);
  s.append(buf + k, 24 - k);
}
static void json_u64_list(std::string& s, const u64* v, size_t n) {
  s += '[';
  for (size_t i = 0; i < n; i++) {
    if (i) s += ',';
    json_u64(s, v[i]);
  }
  s += ']';
}
// TreeHasher::Output in serde form: Poseidon2 digests are [GoldilocksField; 4] = 4 numbers; Blake2s256 / Keccak256 digests are
// [u8; 32] = 32 numbers (src/cs/oracle/mod.rs:180, 245).  Internally every digest is kept as 4 little-endian u64.
static void json_digests(std::string& s, const u64* v, size_t n_digests, bool as_bytes) {
  s += '[';
  for (size_t i = 0; i < n_digests; i++) {
    if (i) s += ',';
    if (!as_bytes) {
      json_u64_list(s, v + 4 * i, 4);
      continue;
    }
    s += '[';
    for (int b = 0; b < 32; b++) {
      if (b) s += ',';
      json_u64(s, (v[4 * i + b / 8] >> (8 * (b % 8))) & 0xff);
    }
    s += ']';
  }
  s += ']';
}
static void json_ext_list(std::string& s, const std::vector<gl::e2>& v) {
  s += '[';
  for (size_t i = 0; i < v.size(); i++) {
    if (i) s += ',';
    s += "{\"coeffs\":[";
    json_u64(s, v[i].c0);
    s += ',';
    json_u64(s, v[i].c1);
    s += "],\"_marker\":null}";
  }
  s += ']';
}

struct QueryAnswer {
  std::vector<u64> leaf_elements;
  std::vector<u64> path;  // 4 * depth
};

}  // namespace bj

struct bj_setup {
  bj_ctx* ctx = nullptr;
  bj_circuit c{};
  std::vector<bj::GateCopy> gate_store;
  std::vector<bj_gate_desc> gates;
  std::vector<uint32_t> pi_cols, pi_rows;
  const uint64_t *sigmas = nullptr, *constants = nullptr, *tables = nullptr;  // borrowed, natural row order
  uint32_t n_tables = 0;
  bj::DevMem lde;  // [V + C + T][D][n], D = max(L, quotient degree): the tree commits to the first L cosets of every column
  bj::Oracle tree;
  uint64_t col_len = 0;  // elements of one LDE column held by this context: n * (D / world)
  const uint64_t* col(uint32_t j) const { return (const uint64_t*)lde.p + (size_t)j * col_len; }
  uint32_t log_l() const {
    uint32_t l = 0;
    while ((1u << l) < c.fri_lde_factor) l++;
    return l;
  }
  // log2 of the LDE factor the columns are evaluated at.  The reference evaluates at max(fri_lde_factor, quotient degree) and
  // commits to the subset of the first fri_lde_factor cosets (prover.rs:178-196 `used_lde_degree`, `subset_for_degree`): in the
  // bit-reversed coset order the first L cosets of the factor-D domain ARE the factor-L domain, so the trees, DEEP, FRI and
  // the queries work on the prefix [0, n * L) of every column and only the quotient stage reads the cosets beyond it.
  uint32_t log_d() const {
    uint32_t l = log_l();
    while ((1u << l) < c.quotient_degree) l++;
    return l;
  }
};

struct bj_proof {
  bj_circuit c{};
  std::vector<bj::u64> witness_cap, stage2_cap, quotient_cap;
  std::vector<std::vector<bj::u64>> fri_caps;
  std::vector<bj::u64> mono_c0, mono_c1;
  std::vector<gl::e2> values_at_z, values_at_z_omega, values_at_0;
  std::vector<bj::u64> public_inputs;
  uint64_t pow_challenge = 0;
  // queries[q][oracle]: witness, stage 2, quotient, setup, then one per FRI oracle
  std::vector<std::vector<bj::QueryAnswer>> queries;
  double stage_seconds[6] = {0, 0, 0, 0, 0, 0};
  std::string json;
};

using namespace bj;

static bool is_pow2(uint32_t x) { return x && !(x & (x - 1)); }

extern "C" {

int32_t bj_setup_create(bj_ctx* ctx, const bj_circuit* circuit, const uint64_t* d_sigmas, const uint64_t* d_constants,
                        const uint64_t* d_lookup_tables, bj_setup** out) {
  bj::DeviceGuard device_guard(ctx);
  if (!ctx || !circuit || !d_sigmas || !out || circuit->num_variables == 0 || circuit->log_n == 0 || circuit->log_n > 28 ||
      !is_pow2(circuit->fri_lde_factor) || !is_pow2(circuit->merkle_tree_cap_size) || !is_pow2(circuit->quotient_degree) ||
      (circuit->num_constants && !d_constants) ||
      (circuit->n_gates && !circuit->gates) || circuit->tree_hasher > BJ_HASHER_KECCAK256 || circuit->transcript > 3)
    BJ_FAIL(ctx, BJ_ERR_INVALID_ARG, "bj_setup_create: bad argument");
  if (circuit->lookup_width && (!d_lookup_tables || circuit->lookup_table_id_column >= circuit->num_constants ||
                                circuit->lookup_variables_offset + circuit->lookup_width * circuit->lookup_num_repetitions >
                                    circuit->num_variables))
    BJ_FAIL(ctx, BJ_ERR_INVALID_ARG, "bj_setup_create: inconsistent lookup description");
  if (ctx->shard.log_stride && !ctx->comm)
    BJ_FAIL(ctx, BJ_ERR_UNSUPPORTED, "bj_setup_create: a coset-sharded context needs a communicator (bj_comm_create_*) for the native driver");
  if (ctx->comm && (circuit->merkle_tree_cap_size < circuit->fri_lde_factor || comm_world(ctx) > circuit->fri_lde_factor ||
                    (1u << ctx->shard_log_lde) != circuit->fri_lde_factor))
    BJ_FAIL(ctx, BJ_ERR_INVALID_ARG, "bj_setup_create: sharded proving needs cap_size >= LDE factor >= world and the LDE factor the communicator was created for");
  *out = nullptr;
  {
    // a proof needs at least one FRI folding step (the JSON has a fri_base_oracle_cap): reject circuits so small that
    // compute_fri_schedule (prover.rs:2281-2372) returns an empty schedule for this cap size, instead of failing inside bj_prove
    uint32_t log_l = 0, new_pow = 0, nq = 0, sched[32], sched_len = 0, fd = 0;
    while ((1u << log_l) < circuit->fri_lde_factor) log_l++;
    if (log_l == 0) BJ_FAIL(ctx, BJ_ERR_INVALID_ARG, "bj_setup_create: fri_lde_factor must be at least 2");
    BJ_TRY(bj_compute_fri_schedule(circuit->security_level, circuit->merkle_tree_cap_size, circuit->pow_bits, log_l, circuit->log_n, &new_pow, &nq,
                                   sched, &sched_len, &fd));
    if (sched_len == 0)
      BJ_FAIL(ctx, BJ_ERR_INVALID_ARG,
              "bj_setup_create: degenerate instance - 2^log_n * fri_lde_factor is too small for merkle_tree_cap_size (empty FRI schedule)");
  }
  std::unique_ptr<bj_setup> s(new bj_setup());
  s->ctx = ctx;
  s->c = *circuit;
  // deep copy of the gate programs (the caller's arrays need not outlive this call)
  s->gate_store.resize(circuit->n_gates);
  s->gates.resize(circuit->n_gates);
  for (uint32_t g = 0; g < circuit->n_gates; g++) {
    const bj_gate_desc& d = circuit->gates[g];
    GateCopy& gc = s->gate_store[g];
    gc.relations.assign(d.relations, d.relations + d.n_relations);
    gc.writes.assign(d.writes, d.writes + d.n_writes);
    if (d.selector_path_len) gc.path.assign(d.selector_path, d.selector_path + d.selector_path_len);
    s->gates[g] = d;
    s->gates[g].relations = gc.relations.data();
    s->gates[g].writes = gc.writes.data();
    s->gates[g].selector_path = gc.path.data();
  }
  s->c.gates = s->gates.data();
  if (circuit->n_public_inputs) {
    if (!circuit->public_input_columns || !circuit->public_input_rows) BJ_FAIL(ctx, BJ_ERR_INVALID_ARG, "bj_setup_create: public input places missing");
    s->pi_cols.assign(circuit->public_input_columns, circuit->public_input_columns + circuit->n_public_inputs);
    s->pi_rows.assign(circuit->public_input_rows, circuit->public_input_rows + circuit->n_public_inputs);
    for (uint32_t i = 0; i < circuit->n_public_inputs; i++)
      if (s->pi_cols[i] >= circuit->num_variables || s->pi_rows[i] >= (1ull << circuit->log_n))
        BJ_FAIL(ctx, BJ_ERR_INVALID_ARG, "bj_setup_create: public input place out of range");
  }
  s->c.public_input_columns = s->pi_cols.data();
  s->c.public_input_rows = s->pi_rows.data();
  s->sigmas = d_sigmas;
  s->constants = d_constants;
  s->tables = d_lookup_tables;
  s->n_tables = circuit->lookup_width ? circuit->lookup_width + 1 : 0;
  const uint32_t V = circuit->num_variables, C = circuit->num_constants, T = s->n_tables;
  const uint32_t log_n = circuit->log_n, log_l = s->log_l(), log_d = s->log_d();
  const u64 n = 1ull << log_n;
  s->col_len = (n << log_d) / comm_world(ctx);
  BJ_TRY(s->lde.alloc(ctx, (size_t)(V + C + T) * s->col_len));
  BJ_TRY(lde_columns(ctx, d_sigmas, (uint64_t*)s->lde.p, log_n, log_d, V));
  if (C) BJ_TRY(lde_columns(ctx, d_constants, (uint64_t*)s->lde.p + (size_t)V * s->col_len, log_n, log_d, C));
  if (T) BJ_TRY(lde_columns(ctx, d_lookup_tables, (uint64_t*)s->lde.p + (size_t)(V + C) * s->col_len, log_n, log_d, T));
  for (uint32_t j = 0; j < V + C + T; j++) s->tree.cols.push_back(s->col(j));
  BJ_TRY(oracle_build(ctx, s->tree, n << log_l, circuit->merkle_tree_cap_size, circuit->tree_hasher, circuit->fri_lde_factor));
  *out = s.release();
  return BJ_OK;
}

void bj_setup_free(bj_setup* s) {
  if (!s) return;
  bj::DeviceGuard device_guard(s->ctx);
  if (s->ctx) cudaStreamSynchronize(s->ctx->stream);
  delete s;
}

int32_t bj_setup_get_cap(const bj_setup* s, uint64_t* h_cap) {
  if (!s || !h_cap) return BJ_ERR_INVALID_ARG;
  memcpy(h_cap, s->tree.cap.data(), sizeof(uint64_t) * s->tree.cap.size());
  return BJ_OK;
}

int32_t bj_prove(bj_ctx* ctx, const bj_setup* setup, const uint64_t* d_variables, const uint64_t* d_multiplicities, bj_proof** out) {
  bj::DeviceGuard device_guard(ctx);
  if (!ctx || !setup || !d_variables || !out || setup->ctx != ctx) BJ_FAIL(ctx, BJ_ERR_INVALID_ARG, "bj_prove: bad argument");
  const bj_circuit& c = setup->c;
  const bool lk = c.lookup_width != 0;
  if (lk && !d_multiplicities) BJ_FAIL(ctx, BJ_ERR_INVALID_ARG, "bj_prove: the lookup argument needs the multiplicities column");
  *out = nullptr;
  std::unique_ptr<bj_proof> pf(new bj_proof());
  pf->c = c;
  pf->c.gates = nullptr;
  const uint32_t V = c.num_variables, C = c.num_constants, Q = c.quotient_degree, L = c.fri_lde_factor, cap = c.merkle_tree_cap_size;
  const uint32_t log_n = c.log_n, log_l = setup->log_l(), log_d = setup->log_d();
  uint32_t log_q = 0;
  while ((1u << log_q) < Q) log_q++;
  const u64 n = 1ull << log_n, nQ = n << log_q;
  // coset shard (multi-GPU): this context holds L / world cosets of every LDE column and Q_loc of the first Q cosets
  const uint32_t world = comm_world(ctx), rank = comm_rank(ctx);
  const u64 nL = (n << log_l) / world;                                   // LOCAL length of the committed part of an LDE column
  const u64 nD = (n << log_d) / world;                                   // LOCAL length (= stride) of an LDE column, D = max(L, Q)
  const u64 nQl = ctx->shard.local_cosets(Q) << log_n;                    // LOCAL quotient points
  if (setup->col_len != nD) BJ_FAIL(ctx, BJ_ERR_INVALID_ARG, "bj_prove: the setup was built with a different shard");
  const uint32_t T = setup->n_tables, wdt = c.lookup_width, nsub = c.lookup_num_repetitions, voff = c.lookup_variables_offset;
  auto t_prev = std::chrono::steady_clock::now();
  auto mark = [&](int stage) -> int32_t {
    BJ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const auto now = std::chrono::steady_clock::now();
    pf->stage_seconds[stage] += std::chrono::duration<double>(now - t_prev).count();
    t_prev = now;
    return BJ_OK;
  };
  struct TrGuard {
    bj_transcript* t;
    ~TrGuard() { bj_transcript_free(t); }
  } trg{c.transcript == 1 ? bj_transcript_new_blake2s() : c.transcript == 2 ? bj_transcript_new_keccak256() : c.transcript == 3 ? bj_transcript_new_poseidon() : bj_transcript_new()};
  bj_transcript* tr = trg.t;
  auto challenge2 = [&]() {
    gl::e2 r;
    r.c0 = bj_transcript_get_challenge(tr);
    r.c1 = bj_transcript_get_challenge(tr);
    return r;
  };
  bj_transcript_witness_merkle_tree_cap(tr, (const uint64_t*)setup->tree.cap.data(), cap);  // prover.rs:211
  // public inputs: read from the witness, committed to before anything else (prover.rs:264-266)
  const uint32_t n_pi = c.n_public_inputs;
  pf->public_inputs.resize(n_pi);
  for (uint32_t i = 0; i < n_pi; i++)
    BJ_CUDA(ctx, cudaMemcpyAsync(&pf->public_inputs[i], d_variables + ((size_t)setup->pi_cols[i] << c.log_n) + setup->pi_rows[i], sizeof(u64),
                                 cudaMemcpyDeviceToHost, ctx->stream));
  if (n_pi) BJ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (uint32_t i = 0; i < n_pi; i++) {
    pf->public_inputs[i] = gl::canon(pf->public_inputs[i]);
    const uint64_t v = pf->public_inputs[i];
    bj_transcript_witness_field_elements(tr, &v, 1);
  }

  // ---- round 1: witness commitment ----
  DevMem w_lde, m_lde;
  BJ_TRY(w_lde.alloc(ctx, (size_t)V * nD));
  BJ_TRY(lde_columns(ctx, d_variables, (uint64_t*)w_lde.p, log_n, log_d, V));
  std::vector<const uint64_t*> w_cols(V);
  for (uint32_t j = 0; j < V; j++) w_cols[j] = (const uint64_t*)w_lde.p + (size_t)j * nD;
  Oracle w_or;
  w_or.cols = w_cols;
  if (lk) {
    BJ_TRY(m_lde.alloc(ctx, nD));
    BJ_TRY(bj_lde(ctx, d_multiplicities, n, (uint64_t*)m_lde.p, log_n, log_d, 1, 0));
    w_or.cols.push_back((const uint64_t*)m_lde.p);  // variables | witness (none) | multiplicities
  }
  BJ_TRY(oracle_build(ctx, w_or, n << log_l, cap, c.tree_hasher, L));
  pf->witness_cap = w_or.cap;
  bj_transcript_witness_merkle_tree_cap(tr, (const uint64_t*)w_or.cap.data(), cap);
  BJ_TRY(mark(0));

  // ---- round 2: copy-permutation grand product, partial products, lookup polynomials ----
  const gl::e2 beta = challenge2(), gamma = challenge2();
  gl::e2 lookup_beta{0, 0}, lookup_gamma{0, 0};
  if (lk) {
    lookup_beta = challenge2();  // prover.rs:402-406
    lookup_gamma = challenge2();
  }
  const uint32_t n_chunks = (V + Q - 1) / Q, n_partial = n_chunks - 1;
  const uint32_t n_s2 = 2 + 2 * n_partial + (lk ? 2 * (nsub + 1) : 0);  // z | partials | A_i | B   (c0, c1 each)
  DevMem st2, s2_lde;
  BJ_TRY(st2.alloc(ctx, (size_t)n_s2 * n));
  {
    std::vector<const uint64_t*> vp(V), sp(V);
    for (uint32_t j = 0; j < V; j++) {
      vp[j] = d_variables + (size_t)j * n;
      sp[j] = setup->sigmas + (size_t)j * n;
    }
    std::vector<uint64_t> nr(V);
    BJ_TRY(bj_non_residues_for_copy_permutation(n, V, nr.data()));
    const uint64_t b[2] = {beta.c0, beta.c1}, g[2] = {gamma.c0, gamma.c1};
    if (world > 1 && n >= (u64)world * 16)  // rows split over the ranks, one all-gather (stage2.cu)
      BJ_TRY(copy_permutation_stage2_sharded(ctx, vp.data(), sp.data(), V, nr.data(), beta, gamma, log_n, Q, st2.p));
    else
      BJ_TRY(bj_copy_permutation_stage2(ctx, vp.data(), sp.data(), V, nr.data(), b, g, log_n, Q, (uint64_t*)st2.p, (uint64_t*)st2.p + n,
                                        (uint64_t*)st2.p + 2 * n));
    if (lk) {
      std::vector<const uint64_t*> lc(wdt * nsub), tc(T);
      for (uint32_t i = 0; i < wdt * nsub; i++) lc[i] = d_variables + (size_t)(voff + i) * n;
      for (uint32_t j = 0; j < T; j++) tc[j] = setup->tables + (size_t)j * n;
      const uint64_t lb[2] = {lookup_beta.c0, lookup_beta.c1}, lg[2] = {lookup_gamma.c0, lookup_gamma.c1};
      BJ_TRY(bj_lookup_polys_specialized(ctx, lc.data(), nsub, wdt, setup->constants + (size_t)c.lookup_table_id_column * n, tc.data(), T,
                                         d_multiplicities, lb, lg, log_n, (uint64_t*)st2.p + (size_t)(2 + 2 * n_partial) * n));
    }
  }
  BJ_TRY(s2_lde.alloc(ctx, (size_t)n_s2 * nD));
  BJ_TRY(lde_columns(ctx, (const uint64_t*)st2.p, (uint64_t*)s2_lde.p, log_n, log_d, n_s2));
  st2.release();
  std::vector<const uint64_t*> s2_cols(n_s2);
  for (uint32_t j = 0; j < n_s2; j++) s2_cols[j] = (const uint64_t*)s2_lde.p + (size_t)j * nD;
  Oracle s2_or;
  s2_or.cols = s2_cols;
  BJ_TRY(oracle_build(ctx, s2_or, n << log_l, cap, c.tree_hasher, L));
  pf->stage2_cap = s2_or.cap;
  bj_transcript_witness_merkle_tree_cap(tr, (const uint64_t*)s2_or.cap.data(), cap);
  const uint32_t a_off = 2 + 2 * n_partial;
  BJ_TRY(mark(1));

  // ---- round 3: quotient ----
  const gl::e2 alpha = challenge2();
  uint32_t n_gate_terms = 0;
  for (const auto& g : setup->gates) n_gate_terms += g.n_writes * g.num_repetitions;
  const uint32_t n_lk_terms = lk ? nsub + 1 : 0;  // lookup terms come first (prover.rs:608-625)
  const uint32_t total_terms = n_lk_terms + n_gate_terms + 1 + 1 + n_partial;
  std::vector<uint64_t> powers(2 * (size_t)total_terms);
  {
    gl::e2 cur{1, 0};
    for (uint32_t i = 0; i < total_terms; i++) {
      powers[2 * i] = cur.c0;
      powers[2 * i + 1] = cur.c1;
      cur = gl::e2_mul(cur, alpha);
    }
  }
  std::vector<const uint64_t*> const_cols(C), sigma_cols(V), table_cols(T);
  for (uint32_t j = 0; j < V; j++) sigma_cols[j] = setup->col(j);
  for (uint32_t j = 0; j < C; j++) const_cols[j] = setup->col(V + j);
  for (uint32_t j = 0; j < T; j++) table_cols[j] = setup->col(V + C + j);
  DevMem qq, qloc;  // qq: [2][nQ] global (c0 then c1); qloc: this rank's cosets among the first Q
  BJ_TRY(qq.alloc(ctx, 2 * nQ));
  uint64_t* q0 = (uint64_t*)qq.p;
  uint64_t* q1 = q0 + nQ;
  uint64_t* const gq0 = q0;
  uint64_t* const gq1 = q1;
  if (world > 1) {
    BJ_TRY(qloc.alloc(ctx, 2 * std::max<u64>(nQl, 1)));
    q0 = (uint64_t*)qloc.p;
    q1 = q0 + nQl;
    BJ_CUDA(ctx, cudaMemsetAsync(qloc.p, 0, sizeof(u64) * 2 * std::max<u64>(nQl, 1), ctx->stream));
  } else {
    BJ_CUDA(ctx, cudaMemsetAsync(qq.p, 0, sizeof(u64) * 2 * nQ, ctx->stream));
  }
  if (nQl) {
  if (lk) {
    std::vector<const uint64_t*> ll(wdt * nsub), al(2 * nsub);
    for (uint32_t i = 0; i < wdt * nsub; i++) ll[i] = w_cols[voff + i];
    for (uint32_t i = 0; i < 2 * nsub; i++) al[i] = s2_cols[a_off + i];
    const uint64_t lb[2] = {lookup_beta.c0, lookup_beta.c1}, lg[2] = {lookup_gamma.c0, lookup_gamma.c1};
    BJ_TRY(bj_quotient_lookup_specialized(ctx, ll.data(), nsub, wdt, const_cols[c.lookup_table_id_column], table_cols.data(), T,
                                          (const uint64_t*)m_lde.p, al.data(), s2_cols[a_off + 2 * nsub], s2_cols[a_off + 2 * nsub + 1], lb, lg,
                                          powers.data(), nQl, q0, q1));
  }
  if (n_gate_terms)
    BJ_TRY(bj_quotient_gates_general_purpose(ctx, setup->gates.data(), (uint32_t)setup->gates.size(), w_cols.data(), V, nullptr, 0,
                                             const_cols.data(), C, powers.data() + 2 * (size_t)n_lk_terms, n_gate_terms, nQl, q0, q1));
  {
    std::vector<uint64_t> nr(V);
    BJ_TRY(bj_non_residues_for_copy_permutation(n, V, nr.data()));
    const uint64_t b[2] = {beta.c0, beta.c1}, g[2] = {gamma.c0, gamma.c1};
    BJ_TRY(bj_quotient_copy_permutation(ctx, w_cols.data(), sigma_cols.data(), V, nr.data(), s2_cols[0], s2_cols[1],
                                        n_partial ? s2_cols.data() + 2 : nullptr, b, g, powers.data() + 2 * (size_t)(n_lk_terms + n_gate_terms),
                                        log_n, log_d, log_q, Q, q0, q1));
  }
  BJ_TRY(bj_quotient_divide_by_vanishing(ctx, q0, q1, log_n, log_q));
  }  // nQl
  if (world > 1) {
    // the one bulk exchange: the quotient cosets recombine (they are interpolated together at size n * Q).  Every rank sends
    // `per` = ceil(Q / world) coset slots (c0 | c1 per slot; ranks beyond Q send padding), one all-gather, then the owned
    // slots are scattered to their global coset positions j = k * world + r.
    const u64 per = std::max<u64>(1, Q / world);
    DevMem snd, rcv;
    BJ_TRY(snd.alloc(ctx, per * 2 * n));
    BJ_TRY(rcv.alloc(ctx, (u64)world * per * 2 * n));
    BJ_CUDA(ctx, cudaMemsetAsync(snd.p, 0, sizeof(u64) * per * 2 * n, ctx->stream));
    const u64 q_loc = nQl >> log_n;
    for (u64 k = 0; k < q_loc; k++) {
      BJ_CUDA(ctx, cudaMemcpyAsync(snd.p + (2 * k) * n, q0 + k * n, sizeof(u64) * n, cudaMemcpyDeviceToDevice, ctx->stream));
      BJ_CUDA(ctx, cudaMemcpyAsync(snd.p + (2 * k + 1) * n, q1 + k * n, sizeof(u64) * n, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    BJ_TRY(comm_all_gather(ctx->comm, snd.p, rcv.p, per * 2 * n));
    for (uint32_t r = 0; r < world; r++)
      for (u64 k = 0; k < per; k++) {
        const u64 j = k * world + r;
        if (j >= Q) continue;
        const u64* part = rcv.p + ((u64)r * per + k) * 2 * n;
        BJ_CUDA(ctx, cudaMemcpyAsync(gq0 + j * n, part, sizeof(u64) * n, cudaMemcpyDeviceToDevice, ctx->stream));
        BJ_CUDA(ctx, cudaMemcpyAsync(gq1 + j * n, part + n, sizeof(u64) * n, cudaMemcpyDeviceToDevice, ctx->stream));
      }
    q0 = gq0;
    q1 = gq1;
    qloc.release();
  }
  // cosets -> natural order, one interpolation of size n*Q on the coset 7, Q chunks of n coefficients (prover.rs:1399-1467)
  BJ_TRY(bj_bitreverse(ctx, q0, log_n + log_q, 2, nQ));
  BJ_TRY(bj_intt_natural_to_natural(ctx, q0, log_n + log_q, 2, nQ, gl::MULT_GEN));
  {
    // the reference's satisfiability guard: the top coefficient must vanish (prover.rs:1425-1438)
    uint64_t top[2];
    BJ_CUDA(ctx, cudaMemcpyAsync(&top[0], q0 + nQ - 1, sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    BJ_CUDA(ctx, cudaMemcpyAsync(&top[1], q1 + nQ - 1, sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    BJ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (top[0] || top[1]) BJ_FAIL(ctx, BJ_ERR_INVALID_ARG, "bj_prove: unsatisfied circuit (quotient is not a polynomial of degree < n * quotient_degree)");
  }
  DevMem chunks, qt_lde;  // chunk j: c0 then c1
  BJ_TRY(chunks.alloc(ctx, 2 * nQ));
  for (uint32_t j = 0; j < Q; j++) {
    BJ_CUDA(ctx, cudaMemcpyAsync(chunks.p + (size_t)(2 * j) * n, q0 + (size_t)j * n, sizeof(u64) * n, cudaMemcpyDeviceToDevice, ctx->stream));
    BJ_CUDA(ctx, cudaMemcpyAsync(chunks.p + (size_t)(2 * j + 1) * n, q1 + (size_t)j * n, sizeof(u64) * n, cudaMemcpyDeviceToDevice, ctx->stream));
  }
  qq.release();
  BJ_TRY(qt_lde.alloc(ctx, (size_t)(2 * Q) * nL));
  BJ_TRY(bj_lde(ctx, (const uint64_t*)chunks.p, n, (uint64_t*)qt_lde.p, log_n, log_l, 2 * Q, 1));
  chunks.release();
  std::vector<const uint64_t*> qt_cols(2 * Q);
  for (uint32_t j = 0; j < 2 * Q; j++) qt_cols[j] = (const uint64_t*)qt_lde.p + (size_t)j * nL;
  Oracle qt_or;
  qt_or.cols = qt_cols;
  BJ_TRY(oracle_build(ctx, qt_or, n << log_l, cap, c.tree_hasher, L));
  pf->quotient_cap = qt_or.cap;
  bj_transcript_witness_merkle_tree_cap(tr, (const uint64_t*)qt_or.cap.data(), cap);
  BJ_TRY(mark(2));

  // ---- round 4: openings.  Order (prover.rs:1549-1683): variables, witness, constants, sigmas, z, partial products,
  //      multiplicities, lookup A, lookup B, lookup tables, quotient chunks.
  const gl::e2 z = challenge2();
  const gl::e2 z_omega = gl::e2_mul_base(z, gl::omega(log_n));
  struct Src {
    const uint64_t* c0;
    const uint64_t* c1;  // nullptr: base-field polynomial
  };
  std::vector<Src> sources;
  for (uint32_t j = 0; j < V; j++) sources.push_back({w_cols[j], nullptr});
  for (uint32_t j = 0; j < C; j++) sources.push_back({const_cols[j], nullptr});
  for (uint32_t j = 0; j < V; j++) sources.push_back({sigma_cols[j], nullptr});
  for (uint32_t i = 0; i < 1 + n_partial; i++) sources.push_back({s2_cols[2 * i], s2_cols[2 * i + 1]});
  std::vector<Src> zero_sources;
  if (lk) {
    sources.push_back({(const uint64_t*)m_lde.p, nullptr});
    for (uint32_t i = 0; i < nsub + 1; i++) {
      sources.push_back({s2_cols[a_off + 2 * i], s2_cols[a_off + 2 * i + 1]});
      zero_sources.push_back({s2_cols[a_off + 2 * i], s2_cols[a_off + 2 * i + 1]});
    }
    for (uint32_t j = 0; j < T; j++) sources.push_back({table_cols[j], nullptr});
  }
  for (uint32_t i = 0; i < Q; i++) sources.push_back({qt_cols[2 * i], qt_cols[2 * i + 1]});
  auto open_at = [&](const std::vector<Src>& srcs, gl::e2 at, std::vector<gl::e2>& vals) -> int32_t {
    std::vector<const uint64_t*> flat;
    for (const auto& s : srcs) {
      flat.push_back(s.c0);
      if (s.c1) flat.push_back(s.c1);
    }
    vals.clear();
    if (flat.empty()) return BJ_OK;
    std::vector<uint64_t> ev(2 * flat.size());
    const uint64_t a[2] = {at.c0, at.c1};
    if (world == 1) {
      BJ_TRY(bj_barycentric_evaluate(ctx, flat.data(), (uint32_t)flat.size(), log_n, a, ev.data()));
    } else {
      // the columns are split over the ranks (every rank opens its block from the coset it owns), results are gathered
      const size_t per = (flat.size() + world - 1) / world, first = std::min(flat.size(), (size_t)rank * per);
      const size_t cnt = std::min(per, flat.size() - first);
      std::vector<uint64_t> mine(2 * per, 0), all(2 * per * world);
      if (cnt) BJ_TRY(bj_barycentric_evaluate(ctx, flat.data() + first, (uint32_t)cnt, log_n, a, mine.data()));
      BJ_TRY(comm_all_gather_host(ctx->comm, (const u64*)mine.data(), (u64*)all.data(), 2 * per));
      memcpy(ev.data(), all.data(), sizeof(uint64_t) * ev.size());   // rank blocks are contiguous: [r][per] == flat order
    }
    size_t k = 0;
    for (const auto& s : srcs) {
      if (s.c1) {  // f0 + u f1 at an Fp2 point (u^2 = 7)
        const gl::e2 a0{ev[2 * k], ev[2 * k + 1]}, b0{ev[2 * k + 2], ev[2 * k + 3]};
        vals.push_back({gl::canon(gl::add(a0.c0, gl::mul7(b0.c1))), gl::canon(gl::add(a0.c1, b0.c0))});
        k += 2;
      } else {
        vals.push_back({ev[2 * k], ev[2 * k + 1]});
        k += 1;
      }
    }
    return BJ_OK;
  };
  std::vector<Src> z_omega_sources{{s2_cols[0], s2_cols[1]}};
  BJ_TRY(open_at(sources, z, pf->values_at_z));
  BJ_TRY(open_at(z_omega_sources, z_omega, pf->values_at_z_omega));
  BJ_TRY(open_at(zero_sources, gl::e2{0, 0}, pf->values_at_0));
  for (const auto* vs : {&pf->values_at_z, &pf->values_at_z_omega, &pf->values_at_0})
    for (const auto& v : *vs) {
      const uint64_t e[2] = {v.c0, v.c1};
      bj_transcript_witness_field_elements(tr, e, 2);
    }
  BJ_TRY(mark(3));

  // ---- round 5: DEEP combination + FRI ----
  // public inputs grouped by opening point w^row in order of first appearance (prover.rs:1805-1821)
  struct PiGroup {
    u64 at;
    std::vector<Src> srcs;
    std::vector<gl::e2> vals;
  };
  std::vector<PiGroup> pi_groups;
  for (uint32_t i = 0; i < n_pi; i++) {
    const u64 at = gl::pow(gl::omega(log_n), setup->pi_rows[i]);
    PiGroup* g = nullptr;
    for (auto& e : pi_groups)
      if (e.at == at) g = &e;
    if (!g) {
      pi_groups.push_back({at, {}, {}});
      g = &pi_groups.back();
    }
    g->srcs.push_back({w_cols[setup->pi_cols[i]], nullptr});
    g->vals.push_back({pf->public_inputs[i], 0});
  }
  const gl::e2 ch0 = challenge2();
  const size_t n_ch = pf->values_at_z.size() + 1 + pf->values_at_0.size() + n_pi;
  std::vector<uint64_t> ch(2 * n_ch);
  {
    gl::e2 cur{1, 0};
    for (size_t i = 0; i < n_ch; i++) {
      ch[2 * i] = cur.c0;
      ch[2 * i + 1] = cur.c1;
      cur = gl::e2_mul(cur, ch0);
    }
  }
  DevMem deep;
  BJ_TRY(deep.alloc(ctx, 2 * nL));
  BJ_CUDA(ctx, cudaMemsetAsync(deep.p, 0, sizeof(u64) * 2 * nL, ctx->stream));
  auto deep_group = [&](const std::vector<Src>& srcs, const std::vector<gl::e2>& vals, gl::e2 at, const uint64_t* chs) -> int32_t {
    if (srcs.empty()) return BJ_OK;
    std::vector<const uint64_t*> p0(srcs.size()), p1(srcs.size());
    std::vector<uint64_t> v(2 * srcs.size());
    for (size_t i = 0; i < srcs.size(); i++) {
      p0[i] = srcs[i].c0;
      p1[i] = srcs[i].c1;
      v[2 * i] = vals[i].c0;
      v[2 * i + 1] = vals[i].c1;
    }
    const uint64_t a[2] = {at.c0, at.c1};
    return bj_deep_quotient_group(ctx, p0.data(), p1.data(), (uint32_t)srcs.size(), v.data(), chs, a, log_n + log_l /* global */,
                                  (uint64_t*)deep.p, (uint64_t*)deep.p + nL);
  };
  BJ_TRY(deep_group(sources, pf->values_at_z, z, ch.data()));
  BJ_TRY(deep_group(z_omega_sources, pf->values_at_z_omega, z_omega, ch.data() + 2 * sources.size()));
  BJ_TRY(deep_group(zero_sources, pf->values_at_0, gl::e2{0, 0}, ch.data() + 2 * (sources.size() + 1)));
  {
    size_t off = sources.size() + 1 + zero_sources.size();
    for (const auto& g : pi_groups) {  // prover.rs:2010-2041
      BJ_TRY(deep_group(g.srcs, g.vals, gl::e2{g.at, 0}, ch.data() + 2 * off));
      off += g.srcs.size();
    }
  }
  uint32_t new_pow = 0, num_queries = 0, sched[32], sched_len = 0, final_degree = 0;
  BJ_TRY(bj_compute_fri_schedule(c.security_level, cap, c.pow_bits, log_l, log_n, &new_pow, &num_queries, sched, &sched_len, &final_degree));
  bj_fri_oracles* fri = nullptr;
  BJ_TRY(bj_do_fri_with_hasher(ctx, tr, (const uint64_t*)deep.p, (const uint64_t*)deep.p + nL, log_n + log_l, sched, sched_len, log_l, cap,
                               c.tree_hasher, &fri));
  struct FriGuard {
    bj_fri_oracles* f;
    ~FriGuard() { bj_fri_oracles_free(f); }
  } frig{fri};
  const uint32_t n_fri = bj_fri_oracles_num_oracles(fri);
  pf->fri_caps.resize(n_fri);
  for (uint32_t i = 0; i < n_fri; i++) {
    pf->fri_caps[i].resize(4 * (size_t)cap);
    BJ_TRY(bj_fri_oracles_get_cap(fri, i, (uint64_t*)pf->fri_caps[i].data()));
  }
  const uint32_t n_mono = bj_fri_oracles_num_monomials(fri);
  pf->mono_c0.resize(n_mono);
  pf->mono_c1.resize(n_mono);
  BJ_TRY(bj_fri_oracles_get_monomials(fri, (uint64_t*)pf->mono_c0.data(), (uint64_t*)pf->mono_c1.data()));
  if (new_pow) {  // prover.rs:2109-2132 with POW = Blake2s256: 5 challenges seed the search, the nonce re-enters the transcript
    uint8_t seed[40];
    for (int i = 0; i < 5; i++) {
      const uint64_t e = bj_transcript_get_challenge(tr);
      for (int k = 0; k < 8; k++) seed[8 * i + k] = (uint8_t)(e >> (8 * k));
    }
    uint64_t nonce = 0;
    BJ_TRY(bj_pow_blake2s(ctx, seed, 40, new_pow, &nonce));
    pf->pow_challenge = nonce;
    const uint64_t lh[2] = {nonce & 0xffffffffull, nonce >> 32};
    bj_transcript_witness_field_elements(tr, lh, 2);
  }
  BJ_TRY(mark(4));

  // ---- queries ----
  const uint32_t max_bits = log_n + log_l;
  std::vector<uint64_t> idxs(num_queries);
  for (auto& i : idxs) i = bj_transcript_get_index_bits(tr, max_bits, max_bits);
  pf->queries.assign(num_queries, {});
  // a query is answered by the rank that owns the coset of its index (leaf t = coset * n + row lies in that rank's subtree);
  // the other ranks look up a dummy leaf, the answers are exchanged and every rank keeps the owner's
  std::vector<uint64_t> loc_idx(num_queries);
  std::vector<uint32_t> owner(num_queries);
  for (uint32_t q = 0; q < num_queries; q++) {
    const uint64_t j = idxs[q] >> log_n, i = idxs[q] & (n - 1);
    owner[q] = (uint32_t)(j % world);
    loc_idx[q] = owner[q] == rank ? (((j / world) << log_n) | i) : 0;
  }
  // every answer part (leaf elements / path of one oracle) is gathered locally first; ONE exchange then carries all of them
  struct Part {
    std::vector<uint64_t> data;  // [num_queries][rec_len]
    size_t rec_len;
  };
  std::vector<Part> parts;  // order: (rows, path) of the 4 base oracles, then (leaf elements, path) of every FRI level
  const Oracle* base[4] = {&w_or, &s2_or, &qt_or, &setup->tree};
  for (const Oracle* o : base) {
    const size_t row_len = o->cols.size();
    uint32_t depth = 0;
    while ((o->n_leaves >> depth) > o->cap_size) depth++;
    Part rows{std::vector<uint64_t>((size_t)num_queries * row_len), row_len};
    Part path{std::vector<uint64_t>((size_t)num_queries * depth * 4), (size_t)depth * 4};
    BJ_TRY(bj_query_leaf_elements(ctx, o->cols.data(), (uint32_t)row_len, 1, o->n_leaves, loc_idx.data(), num_queries, rows.data.data()));
    if (depth)
      BJ_TRY(bj_merkle_paths(ctx, (const uint64_t*)o->leaf_hashes.p, (const uint64_t*)o->nodes.p, o->n_leaves, o->cap_size, loc_idx.data(),
                             num_queries, path.data.data()));
    parts.push_back(std::move(rows));
    parts.push_back(std::move(path));
  }
  {
    uint32_t log_len = log_n;  // coset length of the level's codeword
    std::vector<uint64_t> sub(idxs);
    for (uint32_t lvl = 0; lvl < sched_len; lvl++) {
      const uint32_t k = sched[lvl];
      const size_t le_len = (size_t)2 << k;
      std::vector<uint64_t> locals(num_queries);
      for (uint32_t q = 0; q < num_queries; q++) {
        const uint64_t j = sub[q] >> log_len, i = sub[q] & ((1ull << log_len) - 1);
        locals[q] = owner[q] == rank ? ((((j / world) << log_len) | i) >> k) : 0;
        sub[q] >>= k;
      }
      uint32_t plen = 0;
      Part les{std::vector<uint64_t>((size_t)num_queries * le_len, 0), le_len};
      Part path{std::vector<uint64_t>((size_t)num_queries * 40 * 4, 0), 0};  // [num_queries][plen][4] after the call
      BJ_TRY(bj_fri_oracles_query_batch(fri, lvl, locals.data(), num_queries, les.data.data(), path.data.data(), &plen));
      path.rec_len = (size_t)plen * 4;
      path.data.resize((size_t)num_queries * path.rec_len);
      parts.push_back(std::move(les));
      parts.push_back(std::move(path));
      log_len -= k;
    }
  }
  if (world > 1) {
    size_t total = 0;
    for (const auto& pt : parts) total += pt.data.size();
    std::vector<uint64_t> mine(total), all((size_t)world * total);
    size_t off = 0;
    for (const auto& pt : parts) {
      if (!pt.data.empty()) memcpy(mine.data() + off, pt.data.data(), sizeof(uint64_t) * pt.data.size());
      off += pt.data.size();
    }
    BJ_TRY(comm_all_gather_host(ctx->comm, (const u64*)mine.data(), (u64*)all.data(), total));
    off = 0;
    for (auto& pt : parts) {
      for (uint32_t q = 0; q < num_queries && pt.rec_len; q++)
        memcpy(pt.data.data() + (size_t)q * pt.rec_len, all.data() + (size_t)owner[q] * total + off + (size_t)q * pt.rec_len,
               sizeof(uint64_t) * pt.rec_len);
      off += pt.data.size();
    }
  }
  for (size_t o = 0; o + 1 < parts.size(); o += 2) {
    const Part &le = parts[o], &pa = parts[o + 1];
    for (uint32_t q = 0; q < num_queries; q++) {
      QueryAnswer a;
      a.leaf_elements.assign(le.data.begin() + (size_t)q * le.rec_len, le.data.begin() + (size_t)(q + 1) * le.rec_len);
      a.path.assign(pa.data.begin() + (size_t)q * pa.rec_len, pa.data.begin() + (size_t)(q + 1) * pa.rec_len);
      pf->queries[q].push_back(std::move(a));
    }
  }
  BJ_TRY(mark(5));

  // ---- serde_json shape of Proof (proof.rs:57-143) ----
  std::string& s = pf->json;
  s.reserve(1 << 20);
  const bool digest_bytes = c.tree_hasher != BJ_HASHER_POSEIDON2;
  s += "{\"proof_config\":{\"fri_lde_factor\":" + std::to_string(L) + ",\"merkle_tree_cap_size\":" + std::to_string(cap) +
       ",\"fri_folding_schedule\":null,\"security_level\":" + std::to_string(c.security_level) + ",\"pow_bits\":" + std::to_string(c.pow_bits) +
       "},\"public_inputs\":";
  json_u64_list(s, pf->public_inputs.data(), pf->public_inputs.size());
  s += ",\"witness_oracle_cap\":";
  json_digests(s, pf->witness_cap.data(), cap, digest_bytes);
  s += ",\"stage_2_oracle_cap\":";
  json_digests(s, pf->stage2_cap.data(), cap, digest_bytes);
  s += ",\"quotient_oracle_cap\":";
  json_digests(s, pf->quotient_cap.data(), cap, digest_bytes);
  s += ",\"final_fri_monomials\":[";
  json_u64_list(s, pf->mono_c0.data(), pf->mono_c0.size());
  s += ',';
  json_u64_list(s, pf->mono_c1.data(), pf->mono_c1.size());
  s += "],\"values_at_z\":";
  json_ext_list(s, pf->values_at_z);
  s += ",\"values_at_z_omega\":";
  json_ext_list(s, pf->values_at_z_omega);
  s += ",\"values_at_0\":";
  json_ext_list(s, pf->values_at_0);
  s += ",\"fri_base_oracle_cap\":";
  json_digests(s, pf->fri_caps[0].data(), cap, digest_bytes);
  s += ",\"fri_intermediate_oracles_caps\":[";
  for (uint32_t i = 1; i < n_fri; i++) {
    if (i > 1) s += ',';
    json_digests(s, pf->fri_caps[i].data(), cap, digest_bytes);
  }
  s += "],\"queries_per_fri_repetition\":[";
  static const char* names[4] = {"witness_query", "stage_2_query", "quotient_query", "setup_query"};
  auto json_answer = [&](const QueryAnswer& a) {
    s += "{\"leaf_elements\":";
    json_u64_list(s, a.leaf_elements.data(), a.leaf_elements.size());
    s += ",\"proof\":";
    json_digests(s, a.path.data(), a.path.size() / 4, digest_bytes);
    s += '}';
  };
  for (uint32_t q = 0; q < num_queries; q++) {
    if (q) s += ',';
    s += '{';
    for (int o = 0; o < 4; o++) {
      s += std::string("\"") + names[o] + "\":";
      json_answer(pf->queries[q][o]);
      s += ',';
    }
    s += "\"fri_queries\":[";
    for (uint32_t lvl = 0; lvl < sched_len; lvl++) {
      if (lvl) s += ',';
      json_answer(pf->queries[q][4 + lvl]);
    }
    s += "]}";
  }
  s += "],\"pow_challenge\":" + std::to_string(pf->pow_challenge) + ",\"_marker\":null}";
  *out = pf.release();
  return BJ_OK;
}

void bj_proof_free(bj_proof* p) { delete p; }

int32_t bj_proof_to_json(const bj_proof* p, char* buf, size_t capacity, size_t* needed) {
  if (!p || !needed) return BJ_ERR_INVALID_ARG;
  *needed = p->json.size() + 1;
  if (!buf || capacity < *needed) return buf ? BJ_ERR_INVALID_ARG : BJ_OK;
  memcpy(buf, p->json.c_str(), *needed);
  return BJ_OK;
}

int32_t bj_proof_stage_seconds(const bj_proof* p, double out[6]) {
  if (!p || !out) return BJ_ERR_INVALID_ARG;
  for (int i = 0; i < 6; i++) out[i] = p->stage_seconds[i];
  return BJ_OK;
}

}  // extern "C"
