// Stage-2 (copy-permutation argument) kernels on the trace domain.
//   compute_partial_products_in_extension     src/cs/implementations/copy_permutation.rs:649-766
//     pointwise_rational_in_extension         :114-248   R_c[i] = prod_{j in chunk c} (w_j + beta k_j x_i + gamma) / (w_j + beta sigma_j + gamma)
//     pointwise_product_in_extension          almost_z = prod_c R_c
//     shifted_grand_product_in_extension      :425-510   z[i] = prod_{t < i} almost_z[t]   (exclusive prefix product, z[0] = 1)
//     partial products                        :730-766   p_c = z * R_0 * ... * R_c   for all but the last chunk
//   non_residues_for_copy_permutation / make_non_residues       :512-523, src/cs/implementations/utils.rs:636-688
// All vectors are in natural row order over the trace domain (x_i = omega_n^i), Fp2 as (c0, c1) column pairs.
#include <vector>
#include "ctx.hpp"

namespace bj {

constexpr int CP_MAX_CHUNKS = 48;

struct CopyPermParams {
  const u64* const* vars;    // n_cols pointers, n values each
  const u64* const* sigmas;  // n_cols pointers
  const u64* non_residues;   // n_cols
  u32 n_cols;
  u32 chunk;                 // columns per chunk (quotient degree)
  u32 n_chunks;
  u64 n;                     // rows handled by this launch (the whole domain, or one rank's block of rows)
  u64 row0;                  // global index of the first row (the column pointers are already advanced by row0)
  gl::e2 beta, gamma;
  const u64* xw_lo;          // omega^i = xw_lo[i & mask] * xw_hi[i >> split]
  const u64* xw_hi;
  int xw_split;
  u64* ratios;               // [n_chunks][2][n]   R_c (c0 then c1)
  u64* az_c0;                // almost_z
  u64* az_c1;
};

__global__ void __launch_bounds__(128) copy_perm_ratios_kernel(const CopyPermParams p) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  const u64 gi = i + p.row0;
  const u64 x = gl::mul(__ldg(p.xw_lo + (gi & ((1ull << p.xw_split) - 1))), __ldg(p.xw_hi + (gi >> p.xw_split)));
  const gl::e2 bx = {gl::mul(p.beta.c0, x), gl::mul(p.beta.c1, x)};  // beta * x
  gl::e2 num[CP_MAX_CHUNKS], den[CP_MAX_CHUNKS];
  u32 col = 0;
  for (u32 c = 0; c < p.n_chunks; c++) {
    gl::e2 nu = {1, 0}, de = {1, 0};
    for (u32 j = 0; j < p.chunk && col < p.n_cols; j++, col++) {
      const u64 w = p.vars[col][i];
      const u64 s = p.sigmas[col][i];
      const u64 k = __ldg(p.non_residues + col);
      const u64 wg = gl::add_lazy(p.gamma.c0, w);  // w + gamma.c0 (lazy), shared by numerator and denominator
      // w + beta k x + gamma: k is a small integer (make_non_residues), 64 x 32 bit product kept in 96 bits with the addend
      gl::e2 a;
      if (k >> 31) {
        a = {gl::fma_lazy(bx.c0, k, wg), gl::fma_lazy(bx.c1, k, p.gamma.c1)};
      } else {
        a = {gl::w96_reduce(gl::w96_add64(gl::mul_u32_wide(bx.c0, (u32)k), wg)),
             gl::w96_reduce(gl::w96_add64(gl::mul_u32_wide(bx.c1, (u32)k), p.gamma.c1))};
      }
      // w + beta sigma + gamma: the addend rides in the product's 128 bits
      const gl::e2 b = {gl::fma_lazy(p.beta.c0, s, wg), gl::fma_lazy(p.beta.c1, s, p.gamma.c1)};
      nu = gl::e2_mul_lazy(nu, a);
      de = gl::e2_mul_lazy(de, b);
    }
    num[c] = gl::e2_canon(nu);
    den[c] = gl::e2_canon(de);
  }
  // invert all chunk denominators with one inversion (Montgomery trick inside the thread)
  gl::e2 pre[CP_MAX_CHUNKS];
  gl::e2 acc = {1, 0};
  for (u32 c = 0; c < p.n_chunks; c++) {
    pre[c] = acc;
    acc = gl::e2_mul(acc, den[c]);
  }
  gl::e2 inv = e2_inv_chain(acc);
  gl::e2 az = {1, 0};
  for (int c = (int)p.n_chunks - 1; c >= 0; c--) {
    const gl::e2 dinv = gl::e2_mul(inv, pre[c]);
    inv = gl::e2_mul(inv, den[c]);
    const gl::e2 r = gl::e2_mul(num[c], dinv);
    p.ratios[((u64)c * 2) * p.n + i] = r.c0;
    p.ratios[((u64)c * 2 + 1) * p.n + i] = r.c1;
    az = gl::e2_mul(az, r);
  }
  p.az_c0[i] = az.c0;
  p.az_c1[i] = az.c1;
}

// ---- exclusive prefix product in Fp2 (three launches: block products, scan of block products, apply) ----
constexpr int SCAN_T = 256;  // threads per block
constexpr int SCAN_E = 8;    // elements per thread (contiguous)

__device__ __forceinline__ gl::e2 e2_shfl_up(gl::e2 v, int d) {
  return {__shfl_up_sync(0xffffffffu, v.c0, d), __shfl_up_sync(0xffffffffu, v.c1, d)};
}

// block-wide exclusive scan of one value per thread; returns the exclusive prefix, *total = block product
__device__ __forceinline__ gl::e2 block_exclusive_scan(gl::e2 v, gl::e2* total) {
  __shared__ gl::e2 warp_tot[SCAN_T / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  gl::e2 inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const gl::e2 o = e2_shfl_up(inc, d);
    if (lane >= d) inc = gl::e2_mul(o, inc);
  }
  if (lane == 31) warp_tot[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    gl::e2 w = lane < SCAN_T / 32 ? warp_tot[lane] : gl::e2{1, 0};
#pragma unroll
    for (int d = 1; d < SCAN_T / 32; d <<= 1) {
      const gl::e2 o = e2_shfl_up(w, d);
      if (lane >= d) w = gl::e2_mul(o, w);
    }
    if (lane < SCAN_T / 32) warp_tot[lane] = w;  // inclusive over warps
  }
  __syncthreads();
  gl::e2 excl = e2_shfl_up(inc, 1);
  if (lane == 0) excl = {1, 0};
  if (wid > 0) excl = gl::e2_mul(warp_tot[wid - 1], excl);
  *total = warp_tot[SCAN_T / 32 - 1];
  return excl;
}

// pass 1: product of each block's SCAN_T * SCAN_E elements
__global__ void __launch_bounds__(SCAN_T) scan_block_products_kernel(const u64* __restrict__ c0, const u64* __restrict__ c1, u64 n,
                                                                      u64* __restrict__ bp) {
  const u64 base = ((u64)blockIdx.x * SCAN_T + threadIdx.x) * SCAN_E;
  gl::e2 v = {1, 0};
#pragma unroll
  for (int e = 0; e < SCAN_E; e++)
    if (base + e < n) v = gl::e2_mul(v, {c0[base + e], c1[base + e]});
  gl::e2 total;
  block_exclusive_scan(v, &total);
  if (threadIdx.x == 0) {
    bp[2 * blockIdx.x] = total.c0;
    bp[2 * blockIdx.x + 1] = total.c1;
  }
}

// pass 2: exclusive scan of the block products (single block, loops if there are more than SCAN_T of them)
__global__ void __launch_bounds__(SCAN_T) scan_of_block_products_kernel(u64* __restrict__ bp, u32 n_blocks, u64* __restrict__ grand_total) {
  gl::e2 carry = {1, 0};
  for (u32 start = 0; start < n_blocks; start += SCAN_T) {
    const u32 i = start + threadIdx.x;
    gl::e2 v = i < n_blocks ? gl::e2{bp[2 * i], bp[2 * i + 1]} : gl::e2{1, 0};
    gl::e2 total;
    gl::e2 excl = block_exclusive_scan(v, &total);
    excl = gl::e2_mul(carry, excl);
    if (i < n_blocks) {
      bp[2 * i] = excl.c0;
      bp[2 * i + 1] = excl.c1;
    }
    carry = gl::e2_mul(carry, total);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    grand_total[0] = carry.c0;
    grand_total[1] = carry.c1;
  }
}

// pass 3: out[i] = block_prefix * (exclusive prefix inside the block)
__global__ void __launch_bounds__(SCAN_T) scan_apply_kernel(const u64* __restrict__ c0, const u64* __restrict__ c1, u64 n,
                                                            const u64* __restrict__ bp, u64* __restrict__ o0, u64* __restrict__ o1) {
  const u64 base = ((u64)blockIdx.x * SCAN_T + threadIdx.x) * SCAN_E;
  gl::e2 el[SCAN_E];
  gl::e2 v = {1, 0};
#pragma unroll
  for (int e = 0; e < SCAN_E; e++) {
    el[e] = base + e < n ? gl::e2{c0[base + e], c1[base + e]} : gl::e2{1, 0};
    v = gl::e2_mul(v, el[e]);
  }
  gl::e2 total;
  gl::e2 run = block_exclusive_scan(v, &total);
  run = gl::e2_mul(gl::e2{bp[2 * blockIdx.x], bp[2 * blockIdx.x + 1]}, run);
#pragma unroll
  for (int e = 0; e < SCAN_E; e++) {
    if (base + e < n) {
      o0[base + e] = gl::canon(run.c0);
      o1[base + e] = gl::canon(run.c1);
    }
    run = gl::e2_mul(run, el[e]);
  }
}

// p_c[i] = z[i] * R_0[i] * ... * R_c[i], c < n_chunks - 1 ; out layout [n_chunks-1][2][n]
__global__ void __launch_bounds__(256) copy_perm_partials_kernel(const u64* __restrict__ z0, const u64* __restrict__ z1,
                                                                  const u64* __restrict__ ratios, u32 n_chunks, u64 n,
                                                                  u64* __restrict__ out) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  gl::e2 run = {z0[i], z1[i]};
  for (u32 c = 0; c + 1 < n_chunks; c++) {
    run = gl::e2_mul(run, {ratios[((u64)c * 2) * n + i], ratios[((u64)c * 2 + 1) * n + i]});
    out[((u64)c * 2) * n + i] = run.c0;
    out[((u64)c * 2 + 1) * n + i] = run.c1;
  }
}

// v[i] *= s (Fp2), for the row-sharded grand product: the local exclusive prefix times the product of the earlier blocks
__global__ void __launch_bounds__(256) e2_scale_kernel(u64* __restrict__ c0, u64* __restrict__ c1, u64 n, gl::e2 s) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const gl::e2 r = gl::e2_mul({c0[i], c1[i]}, s);
  c0[i] = gl::canon(r.c0);
  c1[i] = gl::canon(r.c1);
}

int32_t get_pow_tables_public(bj_ctx* ctx, u64 c, int log_n, u64 scale, PowTab* out);  // ntt.cu

// compute_partial_products_in_extension with the ROWS of the trace domain split over the ranks of the context's communicator
// (multi-GPU): rank r computes the chunk ratios, its exclusive prefix product and the partial products for rows
// [r * n / world, (r + 1) * n / world); the block totals (one Fp2 each) are exchanged so that every block starts from the
// product of the earlier ones (and the grand product is checked to be one), then ONE all-gather gives every rank all rows.
// d_out: [2 + 2 * (n_chunks - 1)][n]  = z.c0, z.c1, then (c0, c1) of every partial product, natural row order.
int32_t copy_permutation_stage2_sharded(bj_ctx* ctx, const uint64_t* const* h_variable_cols, const uint64_t* const* h_sigma_cols, u32 n_cols,
                                        const uint64_t* h_non_residues, gl::e2 beta, gl::e2 gamma, u32 log_n, u32 chunk_size, u64* d_out) {
  const u32 world = comm_world(ctx), rank = comm_rank(ctx);
  const u64 n = 1ull << log_n;
  const u32 n_chunks = (n_cols + chunk_size - 1) / chunk_size;
  if (n_chunks > CP_MAX_CHUNKS) BJ_FAIL(ctx, BJ_ERR_UNSUPPORTED, "copy permutation: more than 48 chunks");
  if (n % world || n / world < 1) BJ_FAIL(ctx, BJ_ERR_INVALID_ARG, "copy permutation: fewer rows than ranks");
  const u64 rows = n / world, row0 = (u64)rank * rows;
  const u32 n_out = 2 + 2 * (n_chunks - 1);
  CopyPermParams p{};
  void* d;
  std::vector<const u64*> vp(n_cols), sp(n_cols);
  for (u32 j = 0; j < n_cols; j++) {
    vp[j] = (const u64*)h_variable_cols[j] + row0;
    sp[j] = (const u64*)h_sigma_cols[j] + row0;
  }
  BJ_TRY(param_upload(ctx, vp.data(), sizeof(u64*) * n_cols, &d));
  p.vars = (const u64* const*)d;
  BJ_TRY(param_upload(ctx, sp.data(), sizeof(u64*) * n_cols, &d));
  p.sigmas = (const u64* const*)d;
  std::vector<u64> nr(n_cols);
  for (u32 i = 0; i < n_cols; i++) nr[i] = gl::canon(h_non_residues[i]);
  BJ_TRY(param_upload(ctx, nr.data(), sizeof(u64) * n_cols, &d));
  p.non_residues = (const u64*)d;
  p.n_cols = n_cols;
  p.chunk = chunk_size;
  p.n_chunks = n_chunks;
  p.n = rows;
  p.row0 = row0;
  p.beta = beta;
  p.gamma = gamma;
  PowTab pt;
  BJ_TRY(get_pow_tables_public(ctx, log_n ? gl::omega(log_n) : 1, (int)log_n, 1, &pt));
  p.xw_lo = pt.lo;
  p.xw_hi = pt.hi;
  p.xw_split = pt.split;
  const u64 per_block = (u64)SCAN_T * SCAN_E;
  const u32 n_blocks = (u32)((rows + per_block - 1) / per_block);
  // scratch: ratios [n_chunks][2][rows] | almost_z [2][rows] | block products | total | local result [n_out][rows] | gathered
  const size_t local_out = (size_t)n_out * rows;
  const size_t need = sizeof(u64) * ((size_t)n_chunks * 2 * rows + 2 * rows + 2 * (size_t)n_blocks + 2 + local_out + (size_t)world * local_out);
  BJ_TRY(ensure_scratch(ctx, need));
  u64* ratios = (u64*)ctx->scratch;
  u64* az0 = ratios + (size_t)n_chunks * 2 * rows;
  u64* az1 = az0 + rows;
  u64* bp = az1 + rows;
  u64* grand = bp + 2 * (size_t)n_blocks;
  u64* loc = grand + 2;                 // [n_out][rows]
  u64* all = loc + local_out;           // [world][n_out][rows]
  p.ratios = ratios;
  p.az_c0 = az0;
  p.az_c1 = az1;
  copy_perm_ratios_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, ctx->stream>>>(p);
  BJ_LAUNCH_CHECK(ctx);
  scan_block_products_kernel<<<n_blocks, SCAN_T, 0, ctx->stream>>>(az0, az1, rows, bp);
  BJ_LAUNCH_CHECK(ctx);
  scan_of_block_products_kernel<<<1, SCAN_T, 0, ctx->stream>>>(bp, n_blocks, grand);
  BJ_LAUNCH_CHECK(ctx);
  scan_apply_kernel<<<n_blocks, SCAN_T, 0, ctx->stream>>>(az0, az1, rows, bp, loc, loc + rows);
  BJ_LAUNCH_CHECK(ctx);
  u64 h_total[2];
  BJ_CUDA(ctx, cudaMemcpyAsync(h_total, grand, sizeof(h_total), cudaMemcpyDeviceToHost, ctx->stream));
  BJ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  std::vector<u64> totals(2 * (size_t)world);
  BJ_TRY(comm_all_gather_host(ctx->comm, h_total, totals.data(), 2));
  gl::e2 offset{1, 0}, all_prod{1, 0};
  for (u32 r = 0; r < world; r++) {
    const gl::e2 t{gl::canon(totals[2 * r]), gl::canon(totals[2 * r + 1])};
    if (r < rank) offset = gl::e2_mul(offset, t);
    all_prod = gl::e2_mul(all_prod, t);
  }
  if (gl::canon(all_prod.c0) != 1 || gl::canon(all_prod.c1) != 0)
    BJ_FAIL(ctx, BJ_ERR_INVALID_ARG, "bj_copy_permutation_stage2: grand product != 1 (copy constraints are not satisfied)");
  if (rank) {
    e2_scale_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, ctx->stream>>>(loc, loc + rows, rows, {gl::canon(offset.c0), gl::canon(offset.c1)});
    BJ_LAUNCH_CHECK(ctx);
  }
  if (n_chunks > 1) {
    copy_perm_partials_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, ctx->stream>>>(loc, loc + rows, ratios, n_chunks, rows, loc + 2 * rows);
    BJ_LAUNCH_CHECK(ctx);
  }
  BJ_TRY(comm_all_gather(ctx->comm, loc, all, (u64)local_out));
  // [rank][col][rows] -> [col][n]
  for (u32 r = 0; r < world; r++)
    BJ_CUDA(ctx, cudaMemcpy2DAsync(d_out + (size_t)r * rows, sizeof(u64) * n, all + (size_t)r * local_out, sizeof(u64) * rows, sizeof(u64) * rows, n_out,
                                   cudaMemcpyDeviceToDevice, ctx->stream));
  return BJ_OK;
}

}  // namespace bj

using namespace bj;

extern "C" {

// Legendre symbol by Euler's criterion; make_non_residues (utils.rs:636-688)
int32_t bj_non_residues_for_copy_permutation(uint64_t domain_size, uint32_t num_columns, uint64_t* h_out) {
  if (!h_out || num_columns == 0 || (domain_size & (domain_size - 1))) return BJ_ERR_INVALID_ARG;
  h_out[0] = 1;
  std::vector<u64> in_domain;  // k^domain_size of the accepted ones
  u64 current = 1;
  for (uint32_t c = 1; c < num_columns; c++) {
    for (;;) {
      current = gl::canon(gl::add(current, 1));
      if (gl::pow(current, (gl::P - 1) / 2) != gl::P - 1) continue;  // not a quadratic non-residue
      const u64 t = gl::pow(current, domain_size);
      if (t == 1) continue;  // inside the domain
      bool unique = true;
      for (u64 o : in_domain) unique &= (o != t);
      if (!unique) continue;
      in_domain.push_back(t);
      h_out[c] = current;
      break;
    }
  }
  return BJ_OK;
}

int32_t bj_copy_permutation_stage2(bj_ctx* ctx, const uint64_t* const* h_variable_cols, const uint64_t* const* h_sigma_cols,
                                   uint32_t n_cols, const uint64_t* h_non_residues, const uint64_t h_beta[2],
                                   const uint64_t h_gamma[2], uint32_t log_n, uint32_t chunk_size, uint64_t* d_z_c0,
                                   uint64_t* d_z_c1, uint64_t* d_partials) {
  bj::DeviceGuard device_guard(ctx);
  if (!ctx || !h_variable_cols || !h_sigma_cols || !h_non_residues || !h_beta || !h_gamma || !d_z_c0 || !d_z_c1 || n_cols == 0 ||
      chunk_size == 0 || log_n > 32)
    BJ_FAIL(ctx, BJ_ERR_INVALID_ARG, "bj_copy_permutation_stage2: bad argument");
  const u32 n_chunks = (n_cols + chunk_size - 1) / chunk_size;
  if (n_chunks > CP_MAX_CHUNKS) BJ_FAIL(ctx, BJ_ERR_UNSUPPORTED, "bj_copy_permutation_stage2: more than 48 chunks");
  if (n_chunks > 1 && !d_partials) BJ_FAIL(ctx, BJ_ERR_INVALID_ARG, "bj_copy_permutation_stage2: d_partials is NULL");
  const u64 n = 1ull << log_n;
  CopyPermParams p{};
  void* d;
  BJ_TRY(param_upload(ctx, h_variable_cols, sizeof(u64*) * n_cols, &d));
  p.vars = (const u64* const*)d;
  BJ_TRY(param_upload(ctx, h_sigma_cols, sizeof(u64*) * n_cols, &d));
  p.sigmas = (const u64* const*)d;
  std::vector<u64> nr(n_cols);
  for (u32 i = 0; i < n_cols; i++) nr[i] = gl::canon(h_non_residues[i]);
  BJ_TRY(param_upload(ctx, nr.data(), sizeof(u64) * n_cols, &d));
  p.non_residues = (const u64*)d;
  p.n_cols = n_cols;
  p.chunk = chunk_size;
  p.n_chunks = n_chunks;
  p.n = n;
  p.beta = {gl::canon(h_beta[0]), gl::canon(h_beta[1])};
  p.gamma = {gl::canon(h_gamma[0]), gl::canon(h_gamma[1])};
  PowTab pt;
  BJ_TRY(get_pow_tables_public(ctx, log_n ? gl::omega(log_n) : 1, (int)log_n, 1, &pt));
  p.xw_lo = pt.lo;
  p.xw_hi = pt.hi;
  p.xw_split = pt.split;
  // scratch: ratios [n_chunks][2][n] + almost_z [2][n] + block products
  const u64 per_block = (u64)SCAN_T * SCAN_E;
  const u32 n_blocks = (u32)((n + per_block - 1) / per_block);
  const size_t need = sizeof(u64) * ((size_t)n_chunks * 2 * n + 2 * n + 2 * (size_t)n_blocks + 2);
  BJ_TRY(ensure_scratch(ctx, need));
  u64* ratios = (u64*)ctx->scratch;
  u64* az0 = ratios + (size_t)n_chunks * 2 * n;
  u64* az1 = az0 + n;
  u64* bp = az1 + n;
  u64* grand = bp + 2 * (size_t)n_blocks;
  p.ratios = ratios;
  p.az_c0 = az0;
  p.az_c1 = az1;
  copy_perm_ratios_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(p);
  BJ_LAUNCH_CHECK(ctx);
  scan_block_products_kernel<<<n_blocks, SCAN_T, 0, ctx->stream>>>(az0, az1, n, bp);
  BJ_LAUNCH_CHECK(ctx);
  scan_of_block_products_kernel<<<1, SCAN_T, 0, ctx->stream>>>(bp, n_blocks, grand);
  BJ_LAUNCH_CHECK(ctx);
  scan_apply_kernel<<<n_blocks, SCAN_T, 0, ctx->stream>>>(az0, az1, n, bp, (u64*)d_z_c0, (u64*)d_z_c1);
  BJ_LAUNCH_CHECK(ctx);
  if (n_chunks > 1) {
    copy_perm_partials_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((const u64*)d_z_c0, (const u64*)d_z_c1, ratios, n_chunks, n,
                                                                                    (u64*)d_partials);
    BJ_LAUNCH_CHECK(ctx);
  }
  // the grand product over the whole domain must be one (assert_eq!(current, one), copy_permutation.rs:479)
  u64 h_grand[2];
  BJ_CUDA(ctx, cudaMemcpyAsync(h_grand, grand, sizeof(h_grand), cudaMemcpyDeviceToHost, ctx->stream));
  BJ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (gl::canon(h_grand[0]) != 1 || gl::canon(h_grand[1]) != 0)
    BJ_FAIL(ctx, BJ_ERR_INVALID_ARG, "bj_copy_permutation_stage2: grand product != 1 (copy constraints are not satisfied)");
  return BJ_OK;
}

}  // extern "C"
