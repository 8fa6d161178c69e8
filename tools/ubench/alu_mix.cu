// Micro-benchmarks of the integer pipes on sm_100a for the Goldilocks kernels (what issue rate can the ALU/FMA mix of a
// modular multiplication reach with no memory traffic at all?).  Build: nvcc -arch=sm_100a -O3 -o alu_mix alu_mix.cu -I../../era_boojum_b200/csrc
#include <cstdio>
#include <cuda_runtime.h>
#include "gl64.cuh"
using gl::u64;
using gl::u32;

template <int ILP>
__global__ void __launch_bounds__(256) k_mul(u64* out, u64 seed, int iters) {
  u64 x[ILP], w = seed | 1;
#pragma unroll
  for (int i = 0; i < ILP; i++) x[i] = seed + threadIdx.x * 977 + i * 131 + blockIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = gl::mul(x[i], w);
  }
  u64 s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// butterfly stream: (a, b) -> (a + w b, a - w b), ILP independent pairs
template <int ILP>
__global__ void __launch_bounds__(256) k_bfly(u64* out, u64 seed, int iters) {
  u64 a[ILP], b[ILP], w = seed | 1;
#pragma unroll
  for (int i = 0; i < ILP; i++) {
    a[i] = seed + threadIdx.x * 977 + i * 131 + blockIdx.x;
    b[i] = a[i] * 3 + 1;
  }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) {
      const u64 v = gl::mul(b[i], w);
      b[i] = gl::sub(a[i], v);
      a[i] = gl::add(a[i], v);
    }
  }
  u64 s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s ^= a[i] ^ b[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void __launch_bounds__(256) k_iadd(u64* out, u64 seed, int iters) {
  u32 x[ILP], y = (u32)seed | 1;
#pragma unroll
  for (int i = 0; i < ILP; i++) x[i] = (u32)seed + threadIdx.x + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y));
  }
  u32 s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void __launch_bounds__(256) k_imad(u64* out, u64 seed, int iters) {
  u32 x[ILP], y = (u32)seed | 1;
#pragma unroll
  for (int i = 0; i < ILP; i++) x[i] = (u32)seed + threadIdx.x + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(x[i]) : "r"(y));
  }
  u32 s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// alternating independent IADD / IMAD streams
template <int ILP>
__global__ void __launch_bounds__(256) k_mix(u64* out, u64 seed, int iters) {
  u32 x[ILP], z[ILP], y = (u32)seed | 1;
#pragma unroll
  for (int i = 0; i < ILP; i++) {
    x[i] = (u32)seed + threadIdx.x + i;
    z[i] = x[i] * 7;
  }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) {
      asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y));
      asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(z[i]) : "r"(y));
    }
  }
  u32 s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s ^= x[i] ^ z[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float run(F f, int blocks) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  f(blocks);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  f(blocks);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

int main() {
  cudaDeviceProp pr;
  cudaGetDeviceProperties(&pr, 0);
  const int sms = pr.multiProcessorCount;
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  u64* out;
  cudaMalloc(&out, sizeof(u64) * 256 * sms * 16);
  const int iters = 4096;
  printf("SMs %d, clock attr %d kHz\n", sms, clk_khz);
  for (int wps = 1; wps <= 8; wps *= 2) {  // resident CTAs of 256 threads per SM -> warps per SMSP = 2 * ctas
    const int blocks = sms * wps;
    const double lanes = (double)blocks * 256;
    float t;
    t = run([&](int b) { k_iadd<8><<<b, 256>>>(out, 12345, iters); }, blocks);
    printf("ctas/SM %d  iadd  : %.1f Gop/s/SM-lane-eq  (%.3f ms)  per-SMSP IPC@1.965GHz %.3f\n", wps, lanes * iters * 8 / t / 1e6, t,
           lanes * iters * 8 / 32 / (t * 1e-3) / (sms * 4) / 1.965e9);
    t = run([&](int b) { k_imad<8><<<b, 256>>>(out, 12345, iters); }, blocks);
    printf("ctas/SM %d  imad  : per-SMSP IPC %.3f\n", wps, lanes * iters * 8 / 32 / (t * 1e-3) / (sms * 4) / 1.965e9);
    t = run([&](int b) { k_mix<8><<<b, 256>>>(out, 12345, iters); }, blocks);
    printf("ctas/SM %d  mix   : per-SMSP IPC %.3f\n", wps, lanes * iters * 16 / 32 / (t * 1e-3) / (sms * 4) / 1.965e9);
    t = run([&](int b) { k_mul<8><<<b, 256>>>(out, 12345, iters / 4); }, blocks);
    printf("ctas/SM %d  mul x8: %.2f Gmul/s  (%.1f SMSP-cycles per warp-mul)\n", wps, lanes * (iters / 4) * 8 / t / 1e6,
           (t * 1e-3) * 1.965e9 * sms * 4 / (lanes / 32 * (iters / 4) * 8));
    t = run([&](int b) { k_bfly<8><<<b, 256>>>(out, 12345, iters / 4); }, blocks);
    printf("ctas/SM %d  bfly x8: %.2f Gbfly/s (%.1f SMSP-cycles per warp-butterfly)\n", wps, lanes * (iters / 4) * 8 / t / 1e6,
           (t * 1e-3) * 1.965e9 * sms * 4 / (lanes / 32 * (iters / 4) * 8));
  }
  return 0;
}
