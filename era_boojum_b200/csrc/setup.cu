// Setup / witness materialisation on the device: the two host loops that turn the circuit's placement
// data into the column polynomials the prover consumes.
//   materialize_variables_polynomials_from_dense_hint   src/cs/implementations/witness.rs:325-385
//       column[c][row] = all_values[hint[c][row]]; placeholders (Variable bit 63, src/cs/mod.rs:44, :173) and rows beyond the
//       hint stay zero.
//   create_permutation_polys                            src/cs/implementations/setup.rs:419-502
//       sigma starts as the identity id(c, row) = k_c * w^row (materialize_x_by_non_residue_polys) and every variable's
//       occurrences, visited column by column and row by row, are linked into one cycle: each occurrence receives the identity
//       value of the PREVIOUS occurrence, the first one that of the last (:446-489).  Here: a stable radix sort of the cells by
//       variable index (cub) puts every variable's occurrences next to each other in exactly that visiting order.
#include <cub/cub.cuh>
#include "ctx.hpp"

namespace bj {

constexpr u64 VAR_PLACEHOLDER_BIT = 1ull << 63;  // Variable::placeholder (src/cs/mod.rs:44)
constexpr u64 VAR_INDEX_MASK = (1ull << 48) - 1;

__global__ void __launch_bounds__(256) materialize_columns_kernel(const u64* __restrict__ values, u64 n_values, const u64* __restrict__ hint,
                                                                   u64 hint_rows, u64 n, u32 n_cols, u64* __restrict__ out,
                                                                   unsigned long long* __restrict__ out_of_range) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * n_cols) return;
  const u64 c = i / n, row = i % n;
  u64 v = 0;
  if (row < hint_rows) {
    const u64 h = hint[c * hint_rows + row];
    if (!(h & VAR_PLACEHOLDER_BIT)) {
      const u64 idx = h & VAR_INDEX_MASK;
      if (idx < n_values) v = gl::canon(values[idx]);
      else atomicAdd(out_of_range, 1ull);
    }
  }
  out[i] = v;
}

// id(c, row) = k_c * w^row with w^row = lo[row & mask] * hi[row >> split]
__device__ __forceinline__ u64 identity_value(u64 cell, int log_n, const u64* __restrict__ nr, const u64* __restrict__ lo,
                                              const u64* __restrict__ hi, int split) {
  const u64 row = cell & ((1ull << log_n) - 1);
  const u64 w = gl::mul(__ldg(lo + (row & ((1ull << split) - 1))), __ldg(hi + (row >> split)));
  return gl::mul(w, __ldg(nr + (cell >> log_n)));
}

__global__ void __launch_bounds__(256) sigma_keys_kernel(const u64* __restrict__ hint, u64 n_cells, u64* __restrict__ keys, u32* __restrict__ cells,
                                                          u64* __restrict__ sigma, int log_n, const u64* __restrict__ nr,
                                                          const u64* __restrict__ lo, const u64* __restrict__ hi, int split) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_cells) return;
  const u64 h = hint[i];
  keys[i] = (h & VAR_PLACEHOLDER_BIT) ? ~0ull : (h & VAR_INDEX_MASK);  // placeholders sort behind every variable
  cells[i] = (u32)i;
  sigma[i] = identity_value(i, log_n, nr, lo, hi, split);
}

__global__ void __launch_bounds__(256) sigma_group_start_kernel(const u64* __restrict__ keys, u64 n_cells, u32* __restrict__ start) {
  const u64 s = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_cells) return;
  start[s] = (s == 0 || keys[s - 1] != keys[s]) ? (u32)s : 0u;
}

__global__ void __launch_bounds__(256) sigma_link_kernel(const u64* __restrict__ keys, const u32* __restrict__ cells, const u32* __restrict__ start,
                                                          u64 n_cells, u64* __restrict__ sigma, int log_n, const u64* __restrict__ nr,
                                                          const u64* __restrict__ lo, const u64* __restrict__ hi, int split) {
  const u64 s = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_cells) return;
  const u64 k = keys[s];
  if (k == ~0ull) return;  // placeholder: never copied, keeps the identity
  const u64 me = identity_value(cells[s], log_n, nr, lo, hi, split);
  // the next occurrence receives this one's identity value; the last occurrence hands its value to the first of the group
  if (s + 1 < n_cells && keys[s + 1] == k) sigma[cells[s + 1]] = me;
  else sigma[cells[start[s]]] = me;
}

struct MaxU32 {
  __host__ __device__ __forceinline__ u32 operator()(u32 a, u32 b) const { return a > b ? a : b; }
};

struct TmpBuf {
  void* p = nullptr;
  ~TmpBuf() {
    if (p) cudaFree(p);
  }
};

int32_t get_pow_tables_public(bj_ctx* ctx, u64 c, int log_n, u64 scale, PowTab* out);

}  // namespace bj

using namespace bj;

extern "C" {

int32_t bj_materialize_columns(bj_ctx* ctx, const uint64_t* d_all_values, uint64_t n_values, const uint64_t* d_hint, uint32_t n_cols,
                               uint64_t hint_rows, uint32_t log_n, uint64_t* d_out) {
  bj::DeviceGuard device_guard(ctx);
  if (!ctx || !d_all_values || !d_hint || !d_out || n_cols == 0 || log_n > 32 || hint_rows > (1ull << log_n))
    BJ_FAIL(ctx, BJ_ERR_INVALID_ARG, "bj_materialize_columns: bad argument");
  const u64 n = 1ull << log_n, total = n * n_cols;
  TmpBuf flag;
  BJ_CUDA(ctx, cudaMalloc(&flag.p, sizeof(unsigned long long)));
  BJ_CUDA(ctx, cudaMemsetAsync(flag.p, 0, sizeof(unsigned long long), ctx->stream));
  materialize_columns_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>((const u64*)d_all_values, n_values, (const u64*)d_hint,
                                                                                      hint_rows, n, n_cols, (u64*)d_out,
                                                                                      (unsigned long long*)flag.p);
  BJ_LAUNCH_CHECK(ctx);
  unsigned long long bad = 0;
  BJ_CUDA(ctx, cudaMemcpyAsync(&bad, flag.p, sizeof(bad), cudaMemcpyDeviceToHost, ctx->stream));
  BJ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (bad) BJ_FAIL(ctx, BJ_ERR_INVALID_ARG, "bj_materialize_columns: a hint refers past the end of the witness vector");
  return BJ_OK;
}

int32_t bj_create_permutation_polys(bj_ctx* ctx, const uint64_t* d_placement, uint32_t n_cols, uint32_t log_n, uint64_t* d_sigmas) {
  bj::DeviceGuard device_guard(ctx);
  if (!ctx || !d_placement || !d_sigmas || n_cols == 0 || log_n > 32 || ((u64)n_cols << log_n) >= (1ull << 31))
    BJ_FAIL(ctx, BJ_ERR_INVALID_ARG, "bj_create_permutation_polys: bad argument (at most 2^31 - 1 cells)");
  const u64 n = 1ull << log_n, n_cells = n * n_cols;
  std::vector<uint64_t> nr(n_cols);
  BJ_TRY(bj_non_residues_for_copy_permutation(n, n_cols, nr.data()));
  void* d_nr;
  BJ_TRY(param_upload(ctx, nr.data(), sizeof(u64) * n_cols, &d_nr));
  PowTab pt;
  BJ_TRY(get_pow_tables_public(ctx, log_n ? gl::omega(log_n) : 1, (int)log_n, 1, &pt));
  TmpBuf keys_in, keys_out, cells_in, cells_out, start, tmp;
  BJ_CUDA(ctx, cudaMalloc(&keys_in.p, sizeof(u64) * n_cells));
  BJ_CUDA(ctx, cudaMalloc(&keys_out.p, sizeof(u64) * n_cells));
  BJ_CUDA(ctx, cudaMalloc(&cells_in.p, sizeof(u32) * n_cells));
  BJ_CUDA(ctx, cudaMalloc(&cells_out.p, sizeof(u32) * n_cells));
  BJ_CUDA(ctx, cudaMalloc(&start.p, sizeof(u32) * n_cells));
  const unsigned blocks = (unsigned)((n_cells + 255) / 256);
  sigma_keys_kernel<<<blocks, 256, 0, ctx->stream>>>((const u64*)d_placement, n_cells, (u64*)keys_in.p, (u32*)cells_in.p, (u64*)d_sigmas, (int)log_n,
                                                     (const u64*)d_nr, pt.lo, pt.hi, pt.split);
  BJ_LAUNCH_CHECK(ctx);
  // stable sort by variable index: equal keys keep the column-major visiting order of setup.rs:446-473
  size_t tmp_bytes = 0;
  BJ_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (const u64*)keys_in.p, (u64*)keys_out.p, (const u32*)cells_in.p,
                                                (u32*)cells_out.p, (int)n_cells, 0, 64, ctx->stream));
  size_t scan_bytes = 0;
  BJ_CUDA(ctx, cub::DeviceScan::InclusiveScan(nullptr, scan_bytes, (const u32*)start.p, (u32*)start.p, MaxU32(), (int)n_cells, ctx->stream));
  BJ_CUDA(ctx, cudaMalloc(&tmp.p, std::max(tmp_bytes, scan_bytes) + 16));
  BJ_CUDA(ctx, cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, (const u64*)keys_in.p, (u64*)keys_out.p, (const u32*)cells_in.p, (u32*)cells_out.p,
                                                (int)n_cells, 0, 64, ctx->stream));
  ctx->launches++;
  sigma_group_start_kernel<<<blocks, 256, 0, ctx->stream>>>((const u64*)keys_out.p, n_cells, (u32*)start.p);
  BJ_LAUNCH_CHECK(ctx);
  BJ_CUDA(ctx, cub::DeviceScan::InclusiveScan(tmp.p, scan_bytes, (const u32*)start.p, (u32*)start.p, MaxU32(), (int)n_cells, ctx->stream));
  ctx->launches++;
  sigma_link_kernel<<<blocks, 256, 0, ctx->stream>>>((const u64*)keys_out.p, (const u32*)cells_out.p, (const u32*)start.p, n_cells, (u64*)d_sigmas,
                                                     (int)log_n, (const u64*)d_nr, pt.lo, pt.hi, pt.split);
  BJ_LAUNCH_CHECK(ctx);
  BJ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the temporaries die with this scope
  return BJ_OK;
}

}  // extern "C"
