// Library context: one device + one stream, cached twiddle / coset-power tables, scratch arena.
// Replaces the reference's Worker (src/worker/mod.rs:5-87) as the "data-parallel executor" handle and caches
// what the reference recomputes on every stage (precompute_twiddles_for_fft, utils.rs:88-125).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>
#include <string>
#include <vector>
#include "../../include/boojum_b200.h"
#include "gl64.cuh"

namespace bj {
using gl::u64;

struct PowTab {
  u64 coset;
  int log_n;
  u64 scale;
  int split;
  u64* lo;
  u64* hi;
  u64* full;  // s * c^i for all i < 2^log_n, or nullptr (built while the per-context budget lasts)
};

// Coset shard of a multi-GPU prover: LDE-domain buffers of this context hold only the cosets j = first + k * 2^log_stride
// (k = 0, 1, ...), stored [local coset k][row].  A local flat index [k | i] maps to the global index [k | first | i].
struct CosetShard {
  uint32_t first = 0;
  uint32_t log_stride = 0;
  __host__ __device__ __forceinline__ u64 global_index(u64 t_loc, int log_coset_len) const {
    if (log_stride == 0) return t_loc;
    const u64 i = t_loc & ((1ull << log_coset_len) - 1);
    const u64 k = t_loc >> log_coset_len;
    return ((((k << log_stride) | first)) << log_coset_len) | i;
  }
  // how many of the first `group` (power of two) global cosets are local
  __host__ __device__ __forceinline__ u64 local_cosets(u64 group) const {
    const u64 stride = 1ull << log_stride;
    if (group >= stride) return group >> log_stride;
    return first < group ? 1 : 0;
  }
};

}  // namespace bj

struct bj_comm;

struct bj_ctx {
  int device = 0;
  bj_comm* comm = nullptr;  // communicator of a coset-sharded multi-GPU prover (comm.cu); nullptr = single GPU
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::string last_error;
  // bit-reversed twiddle tables tab[k] = w^bitrev(k); prefix property makes one table serve all sizes
  bj::u64* tw_fwd = nullptr;
  bj::u64* tw_inv = nullptr;
  int tw_log = 0;  // tables hold 2^(tw_log-1) entries
  std::vector<bj::PowTab> pow_cache;
  size_t pow_full_bytes = 0;  // bytes held by the full power tables of pow_cache
  void* scratch = nullptr;
  size_t scratch_bytes = 0;
  void* ptr_table = nullptr;  // device copy of host pointer arrays (Merkle sources)
  size_t ptr_table_bytes = 0;
  // host-buffer entry points: auxiliary copy streams + device staging ring (created on first use)
  bool copy_streams_ready = false;
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
  cudaEvent_t ev_up[3] = {}, ev_done[3] = {}, ev_down[3] = {};
  void* host_ring = nullptr;
  size_t host_ring_bytes = 0;
  void* param_arena = nullptr;  // bump arena for small per-call parameter blocks
  size_t param_off = 0;
  uint64_t launches = 0;  // kernels launched by this library through this context
  int sm_count = 148;
  bool ntt_attr_set = false;
  std::vector<void*> attr_done;  // kernels whose smem attributes are set on this device
  int ntt_use_v2 = 1;            // BJ_NTT_V2=0 forces the generic pass kernel
  int ntt_max_tile_log = 13;  // tunables (env BJ_NTT_*)
  int ntt_pass1_w = -1;
  int ntt_chunk_mb = 0;
  int ntt_full_pow = 1;          // BJ_NTT_FULL_POW=0 keeps the two-level coset power tables only
  uint32_t one = 1;              // a 1 the compiler cannot see (passed as a kernel parameter): additions written as multiply-adds by it issue on the FMA pipe (blake2s.cu)
  int gate_peephole = 15;          // BJ_GATE_PEEPHOLE: bit 0 = alias x*1 / x+0 / x*0, 1 = multiply-add fusion, 2 = linear combinations, 3 = pushing steps (gates.cu)
  int gate_points_per_thread = 0;  // BJ_GATE_POINTS_PER_THREAD=1|2|4: force the gate interpreter's points per thread (0: by size)
  int ntt_l2_persist = 1;        // BJ_NTT_L2_PERSIST=0: do not pin the coset-power table in L2 during the scaled pass
  bool l2_limit_set = false;
  int ntt_bulk = 0;              // BJ_NTT_BULK=1: experiment, bulk-copy (TMA) staged contiguous pass (ntt_v2.cuh)
  cudaMemPool_t pool = nullptr;  // private stream-ordered pool of the prover driver (keeps freed blocks: no OS round trips per proof)
  bj::CosetShard shard;  // bj_ctx_set_coset_shard; default = the whole domain
  uint32_t shard_log_lde = 0;  // LDE factor the shard was declared for (locates the coset bits of flat indices)
};

#define BJ_FAIL(ctx, code, msg)          \
  do {                                   \
    if (ctx) (ctx)->last_error = (msg);  \
    return (code);                       \
  } while (0)

#define BJ_CUDA(ctx, expr)                                                                      \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      if (ctx) (ctx)->last_error = std::string(#expr) + ": " + cudaGetErrorString(_e);          \
      return BJ_ERR_CUDA;                                                                       \
    }                                                                                           \
  } while (0)

#define BJ_TRY(expr)                 \
  do {                               \
    int32_t _s = (expr);             \
    if (_s != BJ_OK) return _s;      \
  } while (0)

#define BJ_LAUNCH_CHECK(ctx)                                                              \
  do {                                                                                    \
    (ctx)->launches++;                                                                    \
    cudaError_t _e = cudaGetLastError();                                                  \
    if (_e != cudaSuccess) {                                                              \
      (ctx)->last_error = std::string("kernel launch: ") + cudaGetErrorString(_e);        \
      return BJ_ERR_CUDA;                                                                 \
    }                                                                                     \
  } while (0)

namespace bj {
// Every extern "C" entry that takes a context runs on THAT context's device whatever the caller's current device is (a
// process may hold contexts on several GPUs, or switch devices between calls); the caller's current device is restored.
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(const bj_ctx* ctx) {
    if (!ctx) return;
    if (cudaGetDevice(&prev) != cudaSuccess) {
      cudaGetLastError();
      prev = -1;
    }
    if (prev != ctx->device) switched = cudaSetDevice(ctx->device) == cudaSuccess && prev >= 0;
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};
int32_t ensure_scratch(bj_ctx* ctx, size_t bytes);
// collectives of the sharded prover (comm.cu); all no-ops / plain copies for a world of one
int32_t comm_all_gather(bj_comm* c, const u64* d_send, u64* d_recv, u64 n);
int32_t comm_all_gather_host(bj_comm* c, const u64* h_send, u64* h_recv, u64 n);
int32_t comm_all_gather_overlapped(bj_comm* c, const u64* d_send, u64* d_recv, u64 n, cudaEvent_t* done);
int32_t comm_wait(bj_comm* c, cudaEvent_t done);
int32_t comm_broadcast_host(bj_comm* c, u64* h_buf, u64 n, uint32_t root);
uint32_t comm_world(const bj_ctx* ctx);
uint32_t comm_rank(const bj_ctx* ctx);
// global cap (cap_size digests) of an oracle whose local tree (this rank's cosets [k][row]) ends in cap_size / world digests:
// cap node c of the global tree belongs to coset c / (cap_size / L) (leaf index = coset * n + row, proof.rs:89-91)
int32_t comm_assemble_cap(bj_ctx* ctx, const u64* h_local_cap, uint32_t cap_size, uint32_t lde_factor, u64* h_global_cap);
int32_t ensure_twiddles(bj_ctx* ctx, int log_n);
int32_t param_upload(bj_ctx* ctx, const void* host, size_t bytes, void** d_out);
}  // namespace bj
