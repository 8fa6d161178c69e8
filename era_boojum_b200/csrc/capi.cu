// Context management, device memory and host self-test hooks of the C-ABI (include/boojum_b200.h).
#include <cstdlib>
#include <cstring>
#include "ctx.hpp"

namespace bj {
int32_t poseidon2_init_constants(bj_ctx* ctx);
}

namespace bj {
// device self-test of the PTX field arithmetic against the portable C versions (same inputs, same thread)
__global__ void field_selftest_kernel(u64 n, u64 seed, unsigned long long* mismatches) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // splitmix64 stream + edge values
  auto next = [](u64& s) {
    s += 0x9E3779B97F4A7C15ull;
    u64 z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  };
  u64 st = seed + i * 0x632BE59BD9B4E019ull;
  const u64 edges[12] = {0, 1, gl::P - 1, gl::P, gl::P + 1, ~0ull, gl::EPS, gl::EPS + 1, 1ull << 63, gl::P - gl::EPS, 2, gl::P - 2};
  u64 a = next(st), b = next(st);
  if ((i & 7) == 0) a = edges[(i >> 3) % 12];
  if ((i & 15) < 2) b = edges[(i >> 4) % 12];
  if ((i & 7) == 3) {
    // structured operands that drive the rare wrap paths of the reduction: a = 2^s (+-1), b with all-zero / all-one
    // 32-bit halves (e.g. 2^63 * (r << 33) leaves x = -r3: the borrow path; 2^32 * ~0 takes the carry path)
    const u64 r = next(st);
    a = (1ull << ((i >> 3) & 63)) + (((i >> 9) & 3) == 1 ? 1 : 0) - (((i >> 9) & 3) == 2 ? 1 : 0);
    const unsigned sel = (unsigned)(i >> 11) & 7;
    u64 lo = r & 0xffffffffull, hi = r >> 32;
    if (sel & 1) lo = (sel & 4) ? 0xffffffffull : 0;
    if (sel & 2) hi = (sel & 4) ? 0xffffffffull : 0;
    b = (hi << 32) | lo;
  }
  if ((i & 63) == 5) {  // every pair of edge values
    a = edges[(i >> 6) % 12];
    b = edges[(i >> 6) / 12 % 12];
  }
  unsigned bad = 0;
  if (gl::mul(a, b) != gl::mul_c(a, b)) bad++;
  if (gl::canon(gl::mul_lazy(a, b)) != gl::mul_c(a, b)) bad++;
  const u64 bc = gl::canon(b);
  if (gl::canon(gl::add(a, bc)) != gl::canon(gl::add_c(a, bc))) bad++;
  if (gl::canon(gl::sub(a, bc)) != gl::canon(gl::sub_c(a, bc))) bad++;
  // b == p is allowed by the contract
  if (gl::canon(gl::add(a, gl::P)) != gl::canon(a)) bad++;
  if (gl::canon(gl::sub(a, gl::P)) != gl::canon(a)) bad++;
  if (gl::mul(a, b) >= gl::P) bad++;
  if (bad) atomicAdd(mismatches, (unsigned long long)bad);
}
}  // namespace bj

using namespace bj;

extern "C" {

int32_t bj_selftest_field(bj_ctx* ctx, uint64_t n, uint64_t seed, uint64_t* h_mismatches) {
  bj::DeviceGuard device_guard(ctx);
  if (!ctx || !h_mismatches) return BJ_ERR_INVALID_ARG;
  *h_mismatches = 0;
  if (n == 0) return BJ_OK;
  struct Counter {  // freed on every exit path
    unsigned long long* d = nullptr;
    ~Counter() {
      if (d) cudaFree(d);
    }
  } c;
  BJ_CUDA(ctx, cudaMalloc(&c.d, sizeof(unsigned long long)));
  BJ_CUDA(ctx, cudaMemsetAsync(c.d, 0, sizeof(unsigned long long), ctx->stream));
  field_selftest_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, seed, c.d);
  BJ_LAUNCH_CHECK(ctx);
  unsigned long long h = 0;
  BJ_CUDA(ctx, cudaMemcpyAsync(&h, c.d, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  BJ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *h_mismatches = h;
  return BJ_OK;
}

const char* bj_version(void) { return "boojum_b200 0.1.0 (sm_100a)"; }

const char* bj_status_string(int32_t s) {
  switch (s) {
    case BJ_OK: return "ok";
    case BJ_ERR_INVALID_ARG: return "invalid argument";
    case BJ_ERR_CUDA: return "CUDA error";
    case BJ_ERR_NO_DEVICE: return "no CUDA device (this library has no CPU fallback)";
    case BJ_ERR_OOM: return "out of device memory";
    case BJ_ERR_UNSUPPORTED: return "unsupported";
    default: return "unknown status";
  }
}

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}

int32_t bj_ctx_create(int32_t device, void* stream, bj_ctx** out_ctx) {
  if (!out_ctx) return BJ_ERR_INVALID_ARG;
  *out_ctx = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    cudaGetLastError();
    return BJ_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= count) return BJ_ERR_INVALID_ARG;
  bj_ctx* ctx = new bj_ctx();
  ctx->device = device;
  bj::DeviceGuard device_guard(ctx);  // the caller's current device is restored on return
  {
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess || cur != device) {
      cudaGetLastError();
      delete ctx;
      return BJ_ERR_CUDA;
    }
  }
  ctx->stream = (cudaStream_t)stream;  // NULL == the CUDA legacy default stream (what torch uses by default)
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
  ctx->ntt_max_tile_log = env_int("BJ_NTT_MAX_TILE_LOG", 13);
  ctx->gate_points_per_thread = env_int("BJ_GATE_POINTS_PER_THREAD", 0);
  ctx->gate_peephole = env_int("BJ_GATE_PEEPHOLE", 15);
  if (ctx->ntt_max_tile_log < 8) ctx->ntt_max_tile_log = 8;
  if (ctx->ntt_max_tile_log > 14) ctx->ntt_max_tile_log = 14;
  ctx->ntt_pass1_w = env_int("BJ_NTT_PASS1_W", -1);
  ctx->ntt_use_v2 = env_int("BJ_NTT_V2", 1);
  ctx->ntt_full_pow = env_int("BJ_NTT_FULL_POW", 1);
  ctx->ntt_bulk = env_int("BJ_NTT_BULK", 0);
  ctx->ntt_l2_persist = env_int("BJ_NTT_L2_PERSIST", 1);
  ctx->ntt_chunk_mb = env_int("BJ_NTT_CHUNK_MB", 0);
  {
    // the prover driver allocates tens of GB per proof: a private pool that never trims keeps the second and later
    // proofs free of cudaMalloc / page-mapping cost (the default pool returns memory to the OS at every synchronisation)
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    if (cudaMemPoolCreate(&ctx->pool, &props) == cudaSuccess) {
      uint64_t keep = ~0ull;
      cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &keep);
    } else {
      cudaGetLastError();
      ctx->pool = nullptr;
    }
  }
  int32_t st = poseidon2_init_constants(ctx);
  if (st != BJ_OK) {
    if (ctx->pool) cudaMemPoolDestroy(ctx->pool);
    delete ctx;
    return st;
  }
  *out_ctx = ctx;
  return BJ_OK;
}

int32_t bj_ctx_destroy(bj_ctx* ctx) {
  if (!ctx) return BJ_OK;
  bj::DeviceGuard device_guard(ctx);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->tw_fwd) cudaFree(ctx->tw_fwd);
  if (ctx->tw_inv) cudaFree(ctx->tw_inv);
  for (auto& e : ctx->pow_cache) {
    cudaFree(e.lo);
    cudaFree(e.hi);
    if (e.full) cudaFree(e.full);
  }
  if (ctx->pool) cudaMemPoolDestroy(ctx->pool);
  if (ctx->scratch) cudaFree(ctx->scratch);
  if (ctx->ptr_table) cudaFree(ctx->ptr_table);
  if (ctx->param_arena) cudaFree(ctx->param_arena);
  if (ctx->host_ring) cudaFree(ctx->host_ring);
  if (ctx->copy_streams_ready) {
    cudaStreamDestroy(ctx->h2d_stream);
    cudaStreamDestroy(ctx->d2h_stream);
    for (int i = 0; i < 3; i++) {
      cudaEventDestroy(ctx->ev_up[i]);
      cudaEventDestroy(ctx->ev_done[i]);
      cudaEventDestroy(ctx->ev_down[i]);
    }
  }
  delete ctx;
  return BJ_OK;
}

int32_t bj_ctx_set_stream(bj_ctx* ctx, void* stream) {
  bj::DeviceGuard device_guard(ctx);
  if (!ctx) return BJ_ERR_INVALID_ARG;
  BJ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->stream = (cudaStream_t)stream;
  return BJ_OK;
}

int32_t bj_ctx_set_coset_shard(bj_ctx* ctx, uint32_t rank, uint32_t world, uint32_t log_lde) {
  bj::DeviceGuard device_guard(ctx);
  if (!ctx) return BJ_ERR_INVALID_ARG;
  if (world == 0 || (world & (world - 1)) || rank >= world || log_lde > 16 || world > (1u << log_lde))
    BJ_FAIL(ctx, BJ_ERR_INVALID_ARG, "bj_ctx_set_coset_shard: world must be a power of two <= the LDE factor and rank < world");
  uint32_t ls = 0;
  while ((1u << ls) < world) ls++;
  ctx->shard_log_lde = log_lde;
  ctx->shard.first = rank;
  ctx->shard.log_stride = ls;
  return BJ_OK;
}

int32_t bj_ctx_synchronize(bj_ctx* ctx) {
  bj::DeviceGuard device_guard(ctx);
  if (!ctx) return BJ_ERR_INVALID_ARG;
  BJ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return BJ_OK;
}

const char* bj_last_error(const bj_ctx* ctx) { return ctx ? ctx->last_error.c_str() : "no context"; }
uint64_t bj_launch_count(const bj_ctx* ctx) { return ctx ? ctx->launches : 0; }

int32_t bj_alloc(bj_ctx* ctx, size_t bytes, void** d_ptr) {
  bj::DeviceGuard device_guard(ctx);
  if (!ctx || !d_ptr) return BJ_ERR_INVALID_ARG;
  cudaError_t e = cudaMalloc(d_ptr, bytes ? bytes : 1);
  if (e != cudaSuccess) {
    cudaGetLastError();
    *d_ptr = nullptr;
    BJ_FAIL(ctx, e == cudaErrorMemoryAllocation ? BJ_ERR_OOM : BJ_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
  }
  return BJ_OK;
}
int32_t bj_free(bj_ctx* ctx, void* d_ptr) {
  bj::DeviceGuard device_guard(ctx);
  if (!ctx) return BJ_ERR_INVALID_ARG;
  if (!d_ptr) return BJ_OK;
  BJ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  BJ_CUDA(ctx, cudaFree(d_ptr));
  return BJ_OK;
}
int32_t bj_upload(bj_ctx* ctx, void* d_dst, const void* h_src, size_t bytes) {
  bj::DeviceGuard device_guard(ctx);
  if (!ctx || (!d_dst && bytes) || (!h_src && bytes)) return BJ_ERR_INVALID_ARG;
  BJ_CUDA(ctx, cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return BJ_OK;
}
int32_t bj_download(bj_ctx* ctx, void* h_dst, const void* d_src, size_t bytes) {
  bj::DeviceGuard device_guard(ctx);
  if (!ctx || (!h_dst && bytes) || (!d_src && bytes)) return BJ_ERR_INVALID_ARG;
  BJ_CUDA(ctx, cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return BJ_OK;
}
int32_t bj_alloc_host_pinned(size_t bytes, void** h_ptr) {
  if (!h_ptr) return BJ_ERR_INVALID_ARG;
  if (cudaMallocHost(h_ptr, bytes ? bytes : 1) != cudaSuccess) {
    cudaGetLastError();
    *h_ptr = nullptr;
    return BJ_ERR_OOM;
  }
  return BJ_OK;
}
int32_t bj_free_host_pinned(void* h_ptr) {
  if (h_ptr && cudaFreeHost(h_ptr) != cudaSuccess) return BJ_ERR_CUDA;
  return BJ_OK;
}

// ---- host self-test hooks (same gl64 / poseidon2 source, host compilation) ----
uint64_t bj_host_gl_mul(uint64_t a, uint64_t b) { return gl::mul(a, b); }
uint64_t bj_host_gl_add(uint64_t a, uint64_t b) { return gl::canon(gl::add_lazy(a, b)); }
uint64_t bj_host_gl_sub(uint64_t a, uint64_t b) { return gl::canon(gl::sub_lazy(a, b)); }
uint64_t bj_host_gl_inv(uint64_t a) { return gl::inv(a); }
uint64_t bj_host_gl_mul_pow2(uint64_t a, uint32_t s) { return gl::mul_pow2(a, s); }
void bj_host_e2_mul(const uint64_t a[2], const uint64_t b[2], uint64_t out[2]) {
  gl::e2 r = gl::e2_mul({a[0], a[1]}, {b[0], b[1]});
  out[0] = r.c0;
  out[1] = r.c1;
}
void bj_host_e2_inv(const uint64_t a[2], uint64_t out[2]) {
  gl::e2 r = gl::e2_inv({gl::canon(a[0]), gl::canon(a[1])});
  out[0] = r.c0;
  out[1] = r.c1;
}

}  // extern "C"
