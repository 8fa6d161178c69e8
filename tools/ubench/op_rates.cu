// Per-opcode issue rates on sm_100a (per SMSP, warp instructions per cycle at 1.965 GHz) for the integer ops the
// Goldilocks kernels are made of.  8 CTAs x 256 threads per SM, 8 independent chains per thread.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;
#define ILP 8
#define KERNEL(name, DECL, INIT, BODY, FOLD)                                           \
  __global__ void __launch_bounds__(256) name(u64* out, u32 seed, int iters) {         \
    DECL;                                                                              \
    u32 y = seed | 1;                                                                  \
    _Pragma("unroll") for (int i = 0; i < ILP; i++) { INIT; }                          \
    for (int it = 0; it < iters; it++) {                                               \
      _Pragma("unroll") for (int i = 0; i < ILP; i++) { BODY; }                        \
    }                                                                                  \
    u64 s = 0;                                                                         \
    _Pragma("unroll") for (int i = 0; i < ILP; i++) { FOLD; }                          \
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;                                    \
  }

KERNEL(k_wide, u64 x[ILP], x[i] = seed + threadIdx.x + i,
       asm volatile("{.reg .u32 l, h; mov.b64 {l, h}, %0; mad.wide.u32 %0, l, %1, %0;}" : "+l"(x[i]) : "r"(y)), s ^= x[i])
KERNEL(k_hi, u32 x[ILP], x[i] = seed + threadIdx.x + i, asm volatile("mad.hi.u32 %0, %0, %1, %1;" : "+r"(x[i]) : "r"(y)), s ^= x[i])
KERNEL(k_lo, u32 x[ILP], x[i] = seed + threadIdx.x + i, asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(x[i]) : "r"(y)), s ^= x[i])
KERNEL(k_addcc, u64 x[ILP], x[i] = seed + threadIdx.x + i,
       asm volatile("{.reg .u32 l, h; mov.b64 {l, h}, %0; add.cc.u32 l, l, %1; addc.u32 h, h, %1; mov.b64 %0, {l, h};}" : "+l"(x[i]) : "r"(y)),
       s ^= x[i])
KERNEL(k_add3, u32 x[ILP], x[i] = seed + threadIdx.x + i, asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y)), s ^= x[i])
KERNEL(k_lop, u32 x[ILP], x[i] = seed + threadIdx.x + i, asm volatile("lop3.b32 %0, %0, %1, %1, 0x96;" : "+r"(x[i]) : "r"(y)), s ^= x[i])
KERNEL(k_shf, u32 x[ILP], x[i] = seed + threadIdx.x + i, asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(x[i]) : "r"(y)), s ^= x[i])
KERNEL(k_selp, u32 x[ILP], x[i] = seed + threadIdx.x + i,
       asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %1, %0, p;}" : "+r"(x[i]) : "r"(y)), s ^= x[i])
// what ptxas does with a 64-bit add and with the lop/add mix
KERNEL(k_add64, u64 x[ILP], x[i] = seed + threadIdx.x + i, x[i] += (u64)y * 0x100000001ull + x[(i + 1) % ILP], s ^= x[i])

typedef void (*K)(u64*, u32, int);
int main() {
  cudaDeviceProp pr;
  cudaGetDeviceProperties(&pr, 0);
  const int sms = pr.multiProcessorCount;
  u64* out;
  cudaMalloc(&out, sizeof(u64) * 256 * sms * 8);
  const int iters = 4096, blocks = sms * 8;
  struct {
    const char* n;
    K k;
    int ops;
  } ks[] = {{"IMAD.WIDE.U32 (mad.wide)", k_wide, 1}, {"IMAD.HI.U32 (mad.hi)", k_hi, 1}, {"IMAD (mad.lo)", k_lo, 1},
            {"add.cc+addc pair", k_addcc, 2},        {"add.u32", k_add3, 1},          {"lop3", k_lop, 1},
            {"shf", k_shf, 1},                       {"setp+selp pair", k_selp, 2},   {"u64 add mix", k_add64, 1}};
  for (auto& e : ks) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    e.k<<<blocks, 256>>>(out, 12345, iters);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    e.k<<<blocks, 256>>>(out, 12345, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double warp_ops = (double)blocks * 8 * iters * ILP * e.ops;
    printf("%-28s  %.3f PTX-ops / cycle / SMSP\n", e.n, warp_ops / (ms * 1e-3) / (sms * 4) / 1.965e9);
  }
  return 0;
}
