// Poseidon2 Merkle-tree kernels: leaf hashing (row gather across column-major sources -> overwrite sponge),
// node levels, cap.  Reference: MerkleTreeWithCap::construct / construct_by_chunking /
// construct_by_chunking_from_flat_sources / continue_from_leaf_hashes (src/cs/oracle/merkle_tree.rs:78-449);
// sponge = GoldilocksPoseidon2Sponge<AbsorptionModeOverwrite>: absorb (src/algebraic_props/sponge.rs:224-239),
// finalize (:300-323), hash_into_node (src/cs/oracle/mod.rs:162-168).
//
// One thread owns one leaf: it walks the sources in preimage order; thread m reads element m of every column,
// so a warp reads 256 contiguous bytes per column (the row gather the CPU does with a cache-hostile Flattener,
// src/field/traits/field_like.rs:301-352, is a coalesced column read here).
#include "ctx.hpp"
#include "poseidon2.cuh"

namespace bj {

// leaf m absorbs source_s[m*epl + e] for s = 0..n_src-1, e = 0..epl-1 (epl = 2^log_epl)
__global__ void __launch_bounds__(128) poseidon2_leaf_kernel(const u64* const* __restrict__ srcs, u32 n_src,
                                                              u64 n_leaves, int log_epl,
                                                              u64* __restrict__ digests) {
  const u64 m = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n_leaves) return;
  u64 st[12];
#pragma unroll
  for (int i = 0; i < 12; i++) st[i] = 0;
  const u64 total = (u64)n_src << log_epl;
  const u64 epl_mask = (1ull << log_epl) - 1;
  const u64 row0 = m << log_epl;
  // one loop (a single inlined copy of the permutation): the last block is zero-filled when it is partial
  for (u64 i = 0; i < total; i += 8) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const u64 idx = i + k;
      u64 v = 0;
      if (idx < total) v = srcs[idx >> log_epl][row0 + (idx & epl_mask)];
      st[k] = v;
    }
    poseidon2_permutation(st);
  }
  u64* out = digests + 4 * m;
#pragma unroll
  for (int k = 0; k < 4; k++) out[k] = gl::canon(st[k]);
}

// node i of the next level = H(prev[2i] || prev[2i+1]) : exactly one permutation
__global__ void __launch_bounds__(128) poseidon2_node_kernel(const u64* __restrict__ prev, u64 n_out,
                                                              u64* __restrict__ next) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_out) return;
  u64 st[12];
  const ulonglong2* in = reinterpret_cast<const ulonglong2*>(prev + 8 * i);
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const ulonglong2 v = in[k];
    st[2 * k] = v.x;
    st[2 * k + 1] = v.y;
  }
#pragma unroll
  for (int k = 8; k < 12; k++) st[k] = 0;
  poseidon2_permutation(st);
  u64* out = next + 4 * i;
#pragma unroll
  for (int k = 0; k < 4; k++) out[k] = gl::canon(st[k]);
}

// rows given contiguously (row-major), arbitrary length
__global__ void __launch_bounds__(128) poseidon2_rows_kernel(const u64* __restrict__ rows, u64 n_rows, u32 row_len,
                                                              u64* __restrict__ digests) {
  const u64 m = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n_rows) return;
  const u64* row = rows + m * row_len;
  u64 st[12];
#pragma unroll
  for (int i = 0; i < 12; i++) st[i] = 0;
  for (u32 i = 0; i < row_len; i += 8) {
#pragma unroll
    for (int k = 0; k < 8; k++) st[k] = (i + k < row_len) ? row[i + k] : 0;
    poseidon2_permutation(st);
  }
#pragma unroll
  for (int k = 0; k < 4; k++) digests[4 * m + k] = gl::canon(st[k]);
}

__global__ void __launch_bounds__(128) poseidon2_permute_kernel(u64* __restrict__ states, u64 n_states) {
  const u64 m = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n_states) return;
  u64 st[12];
#pragma unroll
  for (int i = 0; i < 12; i++) st[i] = states[12 * m + i];
  poseidon2_permutation(st);
#pragma unroll
  for (int i = 0; i < 12; i++) states[12 * m + i] = gl::canon(st[i]);
}

int32_t poseidon2_init_constants(bj_ctx* ctx) {
  BJ_CUDA(ctx, cudaMemcpyToSymbol(c_poseidon_layer_rc, p2_layer_constants_host().v, sizeof(u64) * 31 * 12));
  return BJ_OK;
}

int32_t merkle_nodes_poseidon2(bj_ctx* ctx, const u64* d_leaf_hashes, u64 n_leaves, u32 cap_size, u64* d_nodes) {
  const u64* prev = d_leaf_hashes;
  u64 cnt = n_leaves, written = 0;
  while (cnt > cap_size) {
    const u64 next = cnt / 2;
    u64* dst = d_nodes + 4 * written;
    poseidon2_node_kernel<<<(unsigned)((next + 127) / 128), 128, 0, ctx->stream>>>(prev, next, dst);
    BJ_LAUNCH_CHECK(ctx);
    prev = dst;
    written += next;
    cnt = next;
  }
  return BJ_OK;
}

}  // namespace bj

using namespace bj;

extern "C" {

int32_t bj_merkle_build_poseidon2(bj_ctx* ctx, const uint64_t* const* h_sources, uint32_t n_sources,
                                  uint64_t n_leaves, uint32_t elems_per_leaf, uint32_t cap_size,
                                  uint64_t* d_leaf_hashes, uint64_t* d_nodes) {
  bj::DeviceGuard device_guard(ctx);
  if (!ctx || !h_sources || !d_leaf_hashes || n_sources == 0 || n_leaves == 0)
    BJ_FAIL(ctx, BJ_ERR_INVALID_ARG, "bj_merkle_build_poseidon2: bad argument");
  if ((n_leaves & (n_leaves - 1)) || (cap_size & (cap_size - 1)) || cap_size == 0 || cap_size > n_leaves ||
      (elems_per_leaf & (elems_per_leaf - 1)) || elems_per_leaf == 0)
    BJ_FAIL(ctx, BJ_ERR_INVALID_ARG, "bj_merkle_build_poseidon2: sizes must be powers of two, cap <= leaves");
  if (n_leaves > cap_size && !d_nodes) BJ_FAIL(ctx, BJ_ERR_INVALID_ARG, "bj_merkle_build_poseidon2: d_nodes is NULL");
  void* d_src;
  BJ_TRY(param_upload(ctx, h_sources, sizeof(u64*) * n_sources, &d_src));
  int log_epl = 0;
  while ((1u << log_epl) < elems_per_leaf) log_epl++;
  poseidon2_leaf_kernel<<<(unsigned)((n_leaves + 127) / 128), 128, 0, ctx->stream>>>(
      (const u64* const*)d_src, n_sources, n_leaves, log_epl, (u64*)d_leaf_hashes);
  BJ_LAUNCH_CHECK(ctx);
  if (n_leaves > cap_size) BJ_TRY(merkle_nodes_poseidon2(ctx, (const u64*)d_leaf_hashes, n_leaves, cap_size, (u64*)d_nodes));
  return BJ_OK;
}

int32_t bj_poseidon2_hash_rows(bj_ctx* ctx, const uint64_t* d_rows, uint64_t n_rows, uint32_t row_len,
                               uint64_t* d_digests) {
  bj::DeviceGuard device_guard(ctx);
  if (!ctx || !d_rows || !d_digests) BJ_FAIL(ctx, BJ_ERR_INVALID_ARG, "bj_poseidon2_hash_rows: bad argument");
  if (n_rows == 0) return BJ_OK;
  poseidon2_rows_kernel<<<(unsigned)((n_rows + 127) / 128), 128, 0, ctx->stream>>>((const u64*)d_rows, n_rows, row_len,
                                                                                    (u64*)d_digests);
  BJ_LAUNCH_CHECK(ctx);
  return BJ_OK;
}

int32_t bj_poseidon2_permute(bj_ctx* ctx, uint64_t* d_states, uint64_t n_states) {
  bj::DeviceGuard device_guard(ctx);
  if (!ctx || !d_states) BJ_FAIL(ctx, BJ_ERR_INVALID_ARG, "bj_poseidon2_permute: bad argument");
  if (n_states == 0) return BJ_OK;
  poseidon2_permute_kernel<<<(unsigned)((n_states + 127) / 128), 128, 0, ctx->stream>>>((u64*)d_states, n_states);
  BJ_LAUNCH_CHECK(ctx);
  return BJ_OK;
}

void bj_host_poseidon2_permutation(uint64_t state[12]) {
  u64 st[12];
  for (int i = 0; i < 12; i++) st[i] = state[i];
  poseidon2_permutation(st);
  for (int i = 0; i < 12; i++) state[i] = gl::canon(st[i]);
}

}  // extern "C"
