// Pacing of tcgen05.mma kind::tf32 (cta_group::1, K = 8 per instruction, both operands in shared memory, K-major no-swizzle) as a
// function of the tile shape: cycles per instruction for M in {64, 128} and N in {16 .. 256}, issued back to back by one thread
// with one commit at the end.  Input to the design of the tensor-core conv kernel: how much does a narrow N cost?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_pacing umma_pacing.cu ; run: ./umma_pacing
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) |
         ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mma_tf32(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_tf32_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

struct P { long long* out; int M, N, reps, ts; };

__global__ void __launch_bounds__(128, 1) k_pace(P p) {
  __shared__ __align__(128) uint32_t A[128 * 8];
  __shared__ __align__(128) uint32_t B[256 * 8];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_s;
  const int tid = threadIdx.x;
  for (int i = tid; i < 128 * 8; i += 128) A[i] = 0x3f800000u;
  for (int i = tid; i < 256 * 8; i += 128) B[i] = 0x3f800000u;
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_s;
  // A operand in tensor memory for the .ts form: columns 480..495 of every lane
  {
    const uint32_t ta = tmem + ((uint32_t)(32 * (tid >> 5)) << 16) + 480;
    const uint32_t one = 0x3f800000u;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(ta), "r"(one) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)(p.M >> 4) << 24);
  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    const uint64_t ad = make_desc(smem_u32(A), p.M * 16, 128), bd = make_desc(smem_u32(B), p.N * 16, 128);
    t0 = clock64();
    if (p.ts) for (int r = 0; r < p.reps; ++r) mma_tf32_ts(tmem, tmem + 480, bd, idesc, r > 0);
    else for (int r = 0; r < p.reps; ++r) mma_tf32(tmem, ad, bd, idesc, r > 0);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    for (int it = 0; it < (1 << 24) && !ok; ++it)
      asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    t1 = clock64();
    p.out[0] = ok ? t1 - t0 : -1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  const int Ns[] = {16, 32, 48, 64, 80, 96, 128, 160, 208, 256};
  for (int ts = 0; ts < 2; ++ts)
    for (int M : {128, 64}) {
      if (ts && M == 64) continue;
      for (int N : Ns) {
        if (M == 64 && N % 8) continue;
        long long c1, cn;
        P p1{d, M, N, 8, ts}, pn{d, M, N, 264, ts};
        k_pace<<<1, 128>>>(p1); if (cudaDeviceSynchronize() != cudaSuccess) { printf("M %d N %d: launch failed: %s\n", M, N, cudaGetErrorString(cudaGetLastError())); return 1; }
        cudaMemcpy(&c1, d, 8, cudaMemcpyDeviceToHost);
        k_pace<<<1, 128>>>(pn); cudaDeviceSynchronize();
        cudaMemcpy(&cn, d, 8, cudaMemcpyDeviceToHost);
        printf("%s M %3d N %3d K 8: %6.1f cycles per MMA (8 MMAs: %lld cycles; ideal %5.1f at 8192 MAC/clk... M*N*8/1891 = %5.1f)\n", ts ? "A in TMEM (.ts)" : "A in smem (.ss) ",
               M, N, (double)(cn - c1) / 256.0, c1, M * N * 8 / 2048.0, M * N * 8 / 1891.0);
      }
    }
  return 0;
}
