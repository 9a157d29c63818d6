// Feasibility micro-benchmark for DESIGN.md section 9 (1): the outer-product accumulation of the tensor-product convolution
//     A_s[u][j] = sum_e basis_e[u] * h_e[j]        (U = 276 basis rows, 72 hidden units, K = edges of the segment)
// as a 3xTF32 tcgen05.mma (split a = hi + lo, three MMAs hi*hi + hi*lo + lo*hi, fp32 accumulation in TMEM).
// One CTA, one M = 128 row tile, N = 80 (72 padded to the next multiple of 16), K = 8 per instruction:
//   * operands written by CUDA threads straight into the canonical K-major no-swizzle UMMA layout (8-row x 16-byte core
//     matrices; what a producer warp of k_conv_fused would write instead of feeding FFMA2),
//   * accuracy of 1xTF32 / 3xTF32 against an fp64 reference next to a plain fp32 FMA chain,
//   * cycles per K = 8 step of the 3-MMA group, and cycles to read the 128 x 80 accumulator back (tcgen05.ld).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_tf32x3 umma_tf32x3.cu ; run: ./umma_tf32x3 [K]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#include <cuda_runtime.h>

constexpr int M = 128, N = 80, NREAL = 72, KSTEP = 8;
constexpr int TMEM_COLS = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
// K-major, SWIZZLE_NONE shared-memory descriptor (cute/arch/mma_sm100_desc.hpp): start address, leading (K-direction)
// and stride (8-row-group direction) byte offsets in units of 16 bytes, version 1
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
  for (int it = 0; it < (1 << 22); ++it) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}

struct Params {
  const float* A;      // [M][K] basis values (row u, edge e)
  const float* B;      // [NREAL][K] hidden units (unit j, edge e)
  float* D;            // [M][N]
  long long* cycles;   // [0] MMA issue..commit-complete, [1] TMEM read-out, [2] operand fill
  int* status;
  int K, terms, reps, swap_offsets;
};

// operand tile of one K = 8 step: chunk kc (4 consecutive k) of row r at 16-byte unit r + ROWS * kc
template <int ROWS>
__device__ __forceinline__ uint32_t* elem(uint32_t* base, int kstep, int r, int k) {
  return base + (size_t)kstep * ROWS * 8 + ((size_t)(r + ROWS * (k >> 2)) << 2) + (k & 3);
}

__global__ void __launch_bounds__(128, 1) k_umma(Params p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int nk = p.K / KSTEP;
  uint32_t* Ahi = reinterpret_cast<uint32_t*>(smem);
  uint32_t* Alo = Ahi + (size_t)nk * M * 8;
  uint32_t* Bhi = Alo + (size_t)nk * M * 8;
  uint32_t* Blo = Bhi + (size_t)nk * N * 8;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // ---- operand fill (what a producer warp would do per chunk of edges)
  long long t0 = clock64();
  for (int i = tid; i < M * p.K; i += 128) {
    const int r = i / p.K, k = i % p.K;
    const float v = p.A[i];
    const uint32_t hi = to_tf32(v);
    *elem<M>(Ahi, k / KSTEP, r, k % KSTEP) = hi;
    *elem<M>(Alo, k / KSTEP, r, k % KSTEP) = to_tf32(v - __uint_as_float(hi));
  }
  for (int i = tid; i < N * p.K; i += 128) {
    const int r = i / p.K, k = i % p.K;
    const float v = r < NREAL ? p.B[r * p.K + k] : 0.f;
    const uint32_t hi = to_tf32(v);
    *elem<N>(Bhi, k / KSTEP, r, k % KSTEP) = hi;
    *elem<N>(Blo, k / KSTEP, r, k % KSTEP) = to_tf32(v - __uint_as_float(hi));
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  long long t1 = clock64();
  const uint32_t tmem = tmem_base_s;
  // instruction descriptor: D = f32, A = B = tf32, K-major both, N >> 3, M >> 4
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  const uint32_t a_lbo = p.swap_offsets ? 128 : M * 16, a_sbo = p.swap_offsets ? M * 16 : 128;
  const uint32_t b_lbo = p.swap_offsets ? 128 : N * 16, b_sbo = p.swap_offsets ? N * 16 : 128;

  long long t2 = 0, t3 = 0;
  bool ok = true;
  // reps tiles issued back to back by one thread, ONE commit at the end: steady-state cost per MMA (the tensor core executes
  // them in issue order); reps = 1 gives the latency of a single tile
  if (tid == 0) {
    t2 = clock64();
    // descriptors advance by one K-step tile (start-address field, 16-byte units): one 64-bit add per operand and step
    const uint64_t ah0 = make_desc(smem_u32(Ahi), a_lbo, a_sbo), al0 = make_desc(smem_u32(Alo), a_lbo, a_sbo);
    const uint64_t bh0 = make_desc(smem_u32(Bhi), b_lbo, b_sbo), bl0 = make_desc(smem_u32(Blo), b_lbo, b_sbo);
    constexpr uint64_t a_inc = (M * 8 * 4) >> 4, b_inc = (N * 8 * 4) >> 4;
    for (int rep = 0; rep < p.reps; ++rep) {
      uint64_t ah = ah0, al = al0, bh = bh0, bl = bl0;
      if (p.terms >= 3) {
#pragma unroll 4
        for (int ks = 0; ks < nk; ++ks) {
          mma_tf32(tmem, ah, bh, idesc, ks > 0);
          mma_tf32(tmem, ah, bl, idesc, 1);
          mma_tf32(tmem, al, bh, idesc, 1);
          ah += a_inc; al += a_inc; bh += b_inc; bl += b_inc;
        }
      } else {
#pragma unroll 4
        for (int ks = 0; ks < nk; ++ks) {
          mma_tf32(tmem, ah, bh, idesc, ks > 0);
          ah += a_inc; bh += b_inc;
        }
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  ok = mbar_wait_bounded(&bar, 0);
  if (tid == 0) t3 = clock64();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // ---- accumulator read-out: warp w owns TMEM lanes 32 w .. 32 w + 31 (rows), 8 columns per instruction
  // timed: 10 x (tcgen05.ld 32x32b.x8) per warp, all in flight before one wait; the stores to global memory are not timed
  uint32_t v[N];
  __syncthreads();
  long long t4 = clock64();
  if (ok) {
#pragma unroll
    for (int c0 = 0; c0 < N; c0 += 8) {
      const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16) + c0;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(v[c0]), "=r"(v[c0 + 1]), "=r"(v[c0 + 2]), "=r"(v[c0 + 3]), "=r"(v[c0 + 4]), "=r"(v[c0 + 5]), "=r"(v[c0 + 6]),
                     "=r"(v[c0 + 7]) : "r"(taddr));
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  long long t5 = clock64();
  if (ok) {
    const int row = 32 * warp + (tid & 31);
#pragma unroll
    for (int c = 0; c < N; ++c) p.D[row * N + c] = __uint_as_float(v[c]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS));
  if (tid == 0) {
    p.cycles[0] = t3 - t2; p.cycles[1] = t5 - t4; p.cycles[2] = t1 - t0;
    *p.status = ok ? 0 : 1;
  }
}

// the same product as a plain fp32 FMA chain (what the FFMA2 path computes, up to order)
__global__ void k_fma(const float* A, const float* B, float* D, int K) {
  const int r = blockIdx.x, j = threadIdx.x;
  if (j >= NREAL) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) acc = fmaf(A[r * K + k], B[j * K + k], acc);
  D[r * N + j] = acc;
}

int main(int argc, char** argv) {
  const int K = argc > 1 ? atoi(argv[1]) : 24;
  if (K % KSTEP || K <= 0 || K > 320) { printf("K must be a multiple of 8, <= 320\n"); return 1; }
  std::vector<float> A((size_t)M * K), B((size_t)NREAL * K);
  srand(1);
  auto rnd = [] { return (float)rand() / RAND_MAX * 2.f - 1.f; };
  for (auto& v : A) v = rnd() * 1.7f;                 // basis values: products of features and harmonics, O(1)
  for (auto& v : B) v = fmaxf(rnd() + 0.3f, 0.f);     // hidden units after ReLU
  std::vector<double> ref((size_t)M * NREAL), mag((size_t)M * NREAL);
  for (int r = 0; r < M; ++r)
    for (int j = 0; j < NREAL; ++j) {
      double s = 0, a = 0;
      for (int k = 0; k < K; ++k) { s += (double)A[r * K + k] * B[j * K + k]; a += fabs((double)A[r * K + k] * B[j * K + k]); }
      ref[r * NREAL + j] = s; mag[r * NREAL + j] = a;
    }
  float *dA, *dB, *dD; long long* dC; int* dS;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, (size_t)M * N * 4);
  cudaMalloc(&dC, 3 * sizeof(long long)); cudaMalloc(&dS, sizeof(int));
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = (size_t)(K / KSTEP) * (2 * M + 2 * N) * 8 * 4;
  if (cudaFuncSetAttribute(k_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { printf("smem %zu too large\n", smem); return 1; }
  std::vector<float> D((size_t)M * N);
  auto err = [&](const char* what) {
    double worst = 0, worst_rel = 0;
    for (int r = 0; r < M; ++r)
      for (int j = 0; j < NREAL; ++j) {
        const double e = fabs((double)D[r * N + j] - ref[r * NREAL + j]);
        worst = fmax(worst, e); worst_rel = fmax(worst_rel, e / mag[r * NREAL + j]);
      }
    printf("%-34s max |err| %.3e   max |err| / sum|a b| %.3e\n", what, worst, worst_rel);
    return worst_rel;
  };
  k_fma<<<M, 96>>>(dA, dB, dD, K);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  printf("M %d N %d (72 used) K %d, operand tiles %zu bytes of shared memory\n", M, N, K, smem);
  err("fp32 FMA chain");
  int swap_used = -1;
  for (int swap = 0; swap < 2 && swap_used < 0; ++swap) {          // which of the two offsets is the K-direction one
    Params p{dA, dB, dD, dC, dS, K, 3, 1, swap};
    cudaMemset(dD, 0, D.size() * 4);
    k_umma<<<1, 128, smem>>>(p);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("k_umma (swap %d): %s\n", swap, cudaGetErrorString(e)); return 2; }
    int st; cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
    if (st) { printf("k_umma (swap %d): the MMA never committed\n", swap); return 3; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    char name[64]; snprintf(name, sizeof name, "3xTF32 tcgen05 (offset order %d)", swap);
    if (err(name) < 1e-4) swap_used = swap;
  }
  if (swap_used < 0) { printf("no descriptor variant reproduced the product\n"); return 4; }
  {
    Params p{dA, dB, dD, dC, dS, K, 1, 1, swap_used};
    k_umma<<<1, 128, smem>>>(p); cudaDeviceSynchronize();
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    err("1xTF32 tcgen05");
  }
  for (int terms : {1, 3}) {
    long long c1[3], cn[3];
    const int reps = 32;
    Params p1{dA, dB, dD, dC, dS, K, terms, 1, swap_used};
    k_umma<<<1, 128, smem>>>(p1); cudaDeviceSynchronize();
    cudaMemcpy(c1, dC, sizeof c1, cudaMemcpyDeviceToHost);
    Params pn{dA, dB, dD, dC, dS, K, terms, reps, swap_used};
    k_umma<<<1, 128, smem>>>(pn); cudaDeviceSynchronize();
    cudaMemcpy(cn, dC, sizeof cn, cudaMemcpyDeviceToHost);
    const int per_tile = (K / KSTEP) * terms;
    printf("%dxTF32: one tile (%d MMAs M128 N80 K8) issue -> completion %lld cycles; %d tiles back to back %lld cycles = %.1f cycles per MMA "
           "(floor 128*80/256 = 40); read-out of the 128 x 80 fp32 tile by 4 warps %lld cycles (%.1f B/clk); fill + hi/lo split incl. "
           "global loads %lld cycles\n",
           terms, per_tile, c1[0], reps, cn[0], (double)(cn[0] - c1[0]) / ((reps - 1) * per_tile), cn[1], 128.0 * 80 * 4 / cn[1], cn[2]);
  }
  return 0;
}
