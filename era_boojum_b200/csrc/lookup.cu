// Log-derivative lookup argument over specialised columns with the table id in a constant column
// (LookupParameters::UseSpecializedColumnsWithTableIdAsConstant, the SHA-256 bench configuration).
//   stage 2   compute_lookup_poly_pairs_specialized           src/cs/implementations/lookup_argument_in_ext.rs:320-947
//       A_i[r] = 1 / (beta + sum_j gamma^j col_{i,j}[r] + gamma^w table_id[r])       one per sub-argument
//       B[r]   = m[r] / (beta + sum_j gamma^j t_j[r])                                 t = width + 1 table columns
//   quotient  compute_quotient_terms_for_lookup_specialized   :949-1319
//       alpha_i * (A_i * (beta + sum gamma^j col_{i,j} + gamma^w table_id) - 1),  alpha * (B * (beta + sum gamma^j t_j) - m)
// Stage 2 runs on the trace domain (natural row order); the quotient terms on the first Q cosets of the LDE.
#include <vector>
#include "ctx.hpp"

namespace bj {

constexpr int LK_MAX_SUB = 32;
constexpr int LK_MAX_WIDTH = 8;  // columns per tuple including the table id

struct LookupParams {
  const u64* const* lookup_cols;  // n_sub * width pointers (variable columns of the sub-arguments, in order)
  const u64* table_id_col;        // constant column holding the table id (may be null: no shared id)
  const u64* const* table_cols;   // width (+1 with id) pointers
  const u64* multiplicity;
  u32 n_sub, width, n_table_cols;
  gl::e2 beta;
  gl::e2 gamma_pows[LK_MAX_WIDTH];
  u64 n_points;
  // stage 2 outputs: [n_sub + 1][2][n]  (A_0.c0, A_0.c1, ..., B.c0, B.c1)
  u64* out;
  // quotient inputs / outputs
  const u64* const* a_polys;      // 2 * n_sub pointers (c0, c1)
  const u64* b_c0;
  const u64* b_c1;
  const u64* alphas;              // (n_sub + 1) Fp2
  u64* q_c0;
  u64* q_c1;
};

__device__ __forceinline__ gl::e2 lookup_aggregate(const LookupParams& p, const u64* const* cols, u32 n, const u64* extra, u64 t) {
  // beta + sum_j gamma^j * cols[j][t] (+ gamma^n * extra[t])
  gl::e2 acc = p.beta;
  for (u32 j = 0; j < n; j++) {
    const u64 v = cols[j][t];
    acc.c0 = gl::add(acc.c0, gl::mul(v, p.gamma_pows[j].c0));
    acc.c1 = gl::add(acc.c1, gl::mul(v, p.gamma_pows[j].c1));
  }
  if (extra) {
    const u64 v = extra[t];
    acc.c0 = gl::add(acc.c0, gl::mul(v, p.gamma_pows[n].c0));
    acc.c1 = gl::add(acc.c1, gl::mul(v, p.gamma_pows[n].c1));
  }
  return {gl::canon(acc.c0), gl::canon(acc.c1)};
}

__global__ void __launch_bounds__(128) lookup_polys_kernel(const LookupParams p) {
  const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= p.n_points) return;
  gl::e2 den[LK_MAX_SUB + 1], pre[LK_MAX_SUB + 1];
  for (u32 i = 0; i < p.n_sub; i++) den[i] = lookup_aggregate(p, p.lookup_cols + i * p.width, p.width, p.table_id_col, t);
  den[p.n_sub] = lookup_aggregate(p, p.table_cols, p.n_table_cols, nullptr, t);
  gl::e2 acc = {1, 0};
  for (u32 i = 0; i <= p.n_sub; i++) {
    pre[i] = acc;
    acc = gl::e2_mul(acc, den[i]);
  }
  gl::e2 inv = e2_inv_chain(acc);
  for (int i = (int)p.n_sub; i >= 0; i--) {
    gl::e2 r = gl::e2_mul(inv, pre[i]);
    inv = gl::e2_mul(inv, den[i]);
    if (i == (int)p.n_sub) r = gl::e2_mul_base(r, p.multiplicity[t]);
    p.out[((u64)i * 2) * p.n_points + t] = r.c0;
    p.out[((u64)i * 2 + 1) * p.n_points + t] = r.c1;
  }
}

__global__ void __launch_bounds__(128) quotient_lookup_kernel(const LookupParams p) {
  const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= p.n_points) return;
  gl::e2 q = {0, 0};
  for (u32 i = 0; i < p.n_sub; i++) {
    const gl::e2 d = lookup_aggregate(p, p.lookup_cols + i * p.width, p.width, p.table_id_col, t);
    gl::e2 v = gl::e2_mul(d, {p.a_polys[2 * i][t], p.a_polys[2 * i + 1][t]});
    v.c0 = gl::canon(gl::sub(v.c0, 1));
    v = gl::e2_mul(v, {__ldg(p.alphas + 2 * i), __ldg(p.alphas + 2 * i + 1)});
    q = {gl::canon(gl::add(q.c0, v.c0)), gl::canon(gl::add(q.c1, v.c1))};
  }
  {
    const gl::e2 d = lookup_aggregate(p, p.table_cols, p.n_table_cols, nullptr, t);
    gl::e2 v = gl::e2_mul(d, {p.b_c0[t], p.b_c1[t]});
    v.c0 = gl::canon(gl::sub(v.c0, gl::canon(p.multiplicity[t])));
    v = gl::e2_mul(v, {__ldg(p.alphas + 2 * p.n_sub), __ldg(p.alphas + 2 * p.n_sub + 1)});
    q = {gl::canon(gl::add(q.c0, v.c0)), gl::canon(gl::add(q.c1, v.c1))};
  }
  p.q_c0[t] = gl::canon(gl::add(p.q_c0[t], q.c0));
  p.q_c1[t] = gl::canon(gl::add(p.q_c1[t], q.c1));
}

static int32_t lookup_fill_common(bj_ctx* ctx, LookupParams& p, const uint64_t* const* h_lookup_cols, uint32_t n_sub, uint32_t width,
                                  const uint64_t* d_table_id_col, const uint64_t* const* h_table_cols, uint32_t n_table_cols,
                                  const uint64_t* d_multiplicity, const uint64_t h_beta[2], const uint64_t h_gamma[2]) {
  if (n_sub == 0 || n_sub > LK_MAX_SUB || width == 0 || n_table_cols == 0 || n_table_cols > LK_MAX_WIDTH ||
      width + (d_table_id_col ? 1 : 0) > LK_MAX_WIDTH || width + (d_table_id_col ? 1 : 0) != n_table_cols)
    BJ_FAIL(ctx, BJ_ERR_INVALID_ARG, "lookup: unsupported geometry (tuple width incl. table id must equal the number of table columns, <= 8)");
  void* d;
  BJ_TRY(param_upload(ctx, h_lookup_cols, sizeof(u64*) * n_sub * width, &d));
  p.lookup_cols = (const u64* const*)d;
  BJ_TRY(param_upload(ctx, h_table_cols, sizeof(u64*) * n_table_cols, &d));
  p.table_cols = (const u64* const*)d;
  p.table_id_col = (const u64*)d_table_id_col;
  p.multiplicity = (const u64*)d_multiplicity;
  p.n_sub = n_sub;
  p.width = width;
  p.n_table_cols = n_table_cols;
  p.beta = {gl::canon(h_beta[0]), gl::canon(h_beta[1])};
  const gl::e2 g = {gl::canon(h_gamma[0]), gl::canon(h_gamma[1])};
  gl::e2 cur = {1, 0};
  for (int j = 0; j < LK_MAX_WIDTH; j++) {
    p.gamma_pows[j] = cur;
    cur = gl::e2_mul(cur, g);
  }
  return BJ_OK;
}

}  // namespace bj

using namespace bj;

extern "C" {

int32_t bj_lookup_polys_specialized(bj_ctx* ctx, const uint64_t* const* h_lookup_cols, uint32_t n_subarguments, uint32_t width,
                                    const uint64_t* d_table_id_col, const uint64_t* const* h_table_cols, uint32_t n_table_cols,
                                    const uint64_t* d_multiplicity, const uint64_t h_beta[2], const uint64_t h_gamma[2],
                                    uint32_t log_n, uint64_t* d_out) {
  bj::DeviceGuard device_guard(ctx);
  if (!ctx || !h_lookup_cols || !h_table_cols || !d_multiplicity || !h_beta || !h_gamma || !d_out || log_n > 32)
    BJ_FAIL(ctx, BJ_ERR_INVALID_ARG, "bj_lookup_polys_specialized: bad argument");
  LookupParams p{};
  BJ_TRY(lookup_fill_common(ctx, p, h_lookup_cols, n_subarguments, width, d_table_id_col, h_table_cols, n_table_cols, d_multiplicity,
                            h_beta, h_gamma));
  p.n_points = 1ull << log_n;
  p.out = (u64*)d_out;
  lookup_polys_kernel<<<(unsigned)((p.n_points + 127) / 128), 128, 0, ctx->stream>>>(p);
  BJ_LAUNCH_CHECK(ctx);
  return BJ_OK;
}

int32_t bj_quotient_lookup_specialized(bj_ctx* ctx, const uint64_t* const* h_lookup_ldes, uint32_t n_subarguments, uint32_t width,
                                       const uint64_t* d_table_id_lde, const uint64_t* const* h_table_ldes, uint32_t n_table_cols,
                                       const uint64_t* d_multiplicity_lde, const uint64_t* const* h_a_ldes, const uint64_t* d_b_c0,
                                       const uint64_t* d_b_c1, const uint64_t h_beta[2], const uint64_t h_gamma[2],
                                       const uint64_t* h_alphas, uint64_t n_points, uint64_t* d_q_c0, uint64_t* d_q_c1) {
  bj::DeviceGuard device_guard(ctx);
  if (!ctx || !h_lookup_ldes || !h_table_ldes || !d_multiplicity_lde || !h_a_ldes || !d_b_c0 || !d_b_c1 || !h_beta || !h_gamma ||
      !h_alphas || !d_q_c0 || !d_q_c1 || n_points == 0)
    BJ_FAIL(ctx, BJ_ERR_INVALID_ARG, "bj_quotient_lookup_specialized: bad argument");
  LookupParams p{};
  BJ_TRY(lookup_fill_common(ctx, p, h_lookup_ldes, n_subarguments, width, d_table_id_lde, h_table_ldes, n_table_cols,
                            d_multiplicity_lde, h_beta, h_gamma));
  void* d;
  BJ_TRY(param_upload(ctx, h_a_ldes, sizeof(u64*) * 2 * n_subarguments, &d));
  p.a_polys = (const u64* const*)d;
  std::vector<u64> al(2 * (size_t)(n_subarguments + 1));
  for (size_t i = 0; i < al.size(); i++) al[i] = gl::canon(h_alphas[i]);
  BJ_TRY(param_upload(ctx, al.data(), sizeof(u64) * al.size(), &d));
  p.alphas = (const u64*)d;
  p.b_c0 = (const u64*)d_b_c0;
  p.b_c1 = (const u64*)d_b_c1;
  p.n_points = n_points;
  p.q_c0 = (u64*)d_q_c0;
  p.q_c1 = (u64*)d_q_c1;
  quotient_lookup_kernel<<<(unsigned)((n_points + 127) / 128), 128, 0, ctx->stream>>>(p);
  BJ_LAUNCH_CHECK(ctx);
  return BJ_OK;
}

}  // extern "C"
