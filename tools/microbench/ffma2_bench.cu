// Micro-benchmark: FP32 FMA issue rate on sm_100a, scalar FFMA against packed FFMA2 (fma.rn.f32x2), at the warp counts the
// fused conv kernel runs with.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ void ffma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
  asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
__device__ __forceinline__ unsigned long long pack2(float x, float y) {
  unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y)); return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& x, float& y) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v));
}

template <int MODE>   // 0: 64 scalar FFMA per iteration, 1: 32 FFMA2 (scalar x pair), 2: 32 FFMA2 (pair x pair)
__global__ void k(float* out, int iters, float b0, float h0) {
  float acc[64];
  unsigned long long acc2[32];
#pragma unroll
  for (int i = 0; i < 64; ++i) acc[i] = (float)i;
#pragma unroll
  for (int i = 0; i < 32; ++i) acc2[i] = pack2((float)i, (float)i + 0.5f);
  float b[8], h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { b[i] = b0 + i + threadIdx.x; h[i] = h0 + 2 * i; }
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int j = 0; j < 8; ++j) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[k * 8 + j]) : "f"(b[k]), "f"(h[j]));
    } else if (MODE == 1) {
#pragma unroll
      for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j) ffma2(acc2[k * 4 + j], pack2(b[k], b[k]), pack2(h[2 * j], h[2 * j + 1]));
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j) ffma2(acc2[k * 4 + j], pack2(b[k], b[(k + 1) & 7]), pack2(h[2 * j], h[2 * j + 1]));
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 64; ++i) s += acc[i];
#pragma unroll
  for (int i = 0; i < 32; ++i) { float x, y; unpack2(acc2[i], x, y); s += x + y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  float* out; cudaMalloc(&out, 148 * 1024 * 4);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int iters = 20000;
  for (int warps : {4, 8, 12, 16, 32}) {
    for (int mode = 0; mode < 3; ++mode) {
      float ms = 0;
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(a);
        if (mode == 0) k<0><<<148, warps * 32>>>(out, iters, 1.f, 2.f);
        else if (mode == 1) k<1><<<148, warps * 32>>>(out, iters, 1.f, 2.f);
        else k<2><<<148, warps * 32>>>(out, iters, 1.f, 2.f);
        cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
      }
      double fma = 148.0 * warps * 32 * 64.0 * iters;
      printf("warps/SM %2d mode %d (%s): %.3f ms  %.1f TFLOP/s  %.1f FMA/clk/SM @1.965GHz\n", warps, mode,
             mode == 0 ? "FFMA" : (mode == 1 ? "FFMA2 scalar*pair" : "FFMA2 pair*pair"), ms, 2 * fma / ms / 1e9,
             fma / 148 / (ms * 1e-3 * 1.965e9));
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
