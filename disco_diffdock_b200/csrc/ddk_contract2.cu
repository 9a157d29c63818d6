// k_conv_contract2: out[node] = sum_{group} ( A_seg (*) W2p + Bsum_seg (*) b2p ), mean, batch-norm affine, residual.
//
// The contraction streams the outer-product scratch A (read once, ~80 KB per non-empty segment) against the packed
// second-layer weights, i.e. it is bound by how fast A can be read.  v1 (ddk_conv.cu) spends its issue slots on scalar
// FFMA + shared-memory operand loads and keeps too few bytes in flight; here the arithmetic runs on the tensor pipe as
// a split-precision (3xTF32) mma.sync GEMM  [16 nodes] x [K = F*72] x [O]  per warp:
//     a*w ~= a_hi*w_hi + a_lo*w_hi + a_hi*w_lo      (a = a_hi + a_lo, w = w_hi + w_lo, hi/lo in TF32, fp32 accumulate)
// which keeps fp32-level accuracy (the dropped a_lo*w_lo term is ~2^-22 relative) while freeing registers and issue
// slots so that every lane can keep several 16-byte loads of A in flight.  One CTA = 4 warps x 16 nodes of one type.
#include "ddk_conv.cuh"

namespace ddk {

constexpr int C2_WARPS = 4;
constexpr int C2_THREADS = C2_WARPS * 32;
constexpr int C2_TM = C2_WARPS * 16;   // nodes per CTA
constexpr int C2_KT = 128;             // K rows of W2p staged per pass
constexpr int C2_OPMAX = 26;           // padded row stride (floats) of the staged weights: 4*OP = 8 (mod 32) -> conflict-free

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = to_tf32(x);
  lo = to_tf32(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

struct C2Smem {
  alignas(16) float Whi[2][C2_KT][C2_OPMAX];
  alignas(16) float Wlo[2][C2_KT][C2_OPMAX];
  float Out[C2_TM][D];
  float Cnt[C2_TM];
};

// one irrep class for the 16 nodes of this warp; O = multiplicity of the output irrep, NC = 1 (scalars) or 3 (vectors)
template <int O, int NC>
__device__ __forceinline__ void contract2_class(const ConArgs& p, const ClassInfo& ci, bool lig, int U, const int (&sidx)[2][2],
                                                C2Smem& S, int w, int lane) {
  constexpr int NT = (O + 7) / 8;
  constexpr int PADC = (8 * NT - O) > 0 ? (8 * NT - O) : 1;     // padding columns of the last n-tile
  constexpr int SU = (NC == 1) ? 4 : 2;          // 16-wide K slabs whose loads are issued together
  constexpr int WPT = (C2_KT * O + C2_THREADS - 1) / C2_THREADS;   // staged weights per thread and chunk
  const int g = lane >> 2, t = lane & 3, tid = threadIdx.x;
  float acc[NC][NT][4];
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[c][nt][i] = 0.f;
  const int K = ci.F * HID;
  float (*Whi)[C2_KT][C2_OPMAX] = S.Whi;
  float (*Wlo)[C2_KT][C2_OPMAX] = S.Wlo;

  for (int which = 0; which < 2; ++which) {
    const int grp = lig ? which : 2 + which;
    const float* Wg = p.W2p[grp] + ci.woff;
    const float* rowp[2][NC];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int c = 0; c < NC; ++c)
        rowp[r][c] = sidx[r][which] >= 0 ? p.A + ((size_t)sidx[r][which] * U + ci.uoff + c * ci.F) * HID : nullptr;

    float wreg[WPT];
    auto load_w = [&](int c0) {
#pragma unroll
      for (int q = 0; q < WPT; ++q) {
        int i = tid + q * C2_THREADS;
        int r = i / O;
        wreg[q] = (i < C2_KT * O && c0 + r < K) ? Wg[(size_t)c0 * O + i] : 0.f;
      }
    };
    auto store_w = [&](int buf) {
#pragma unroll
      for (int q = 0; q < WPT; ++q) {
        int i = tid + q * C2_THREADS;
        if (i < C2_KT * O) {
          int r = i / O, o = i % O;
          uint32_t hi, lo;
          split_tf32(wreg[q], hi, lo);
          Whi[buf][r][o] = __uint_as_float(hi);
          Wlo[buf][r][o] = __uint_as_float(lo);
        }
      }
      if (O % 8 != 0)   // zero the padding columns of the last n-tile
        for (int i = tid; i < C2_KT * PADC; i += C2_THREADS) {
          int r = i / PADC, o = O + i % PADC;
          Whi[buf][r][o] = 0.f;
          Wlo[buf][r][o] = 0.f;
        }
    };

    __syncthreads();            // previous users of the staging buffers are done
    load_w(0);
    store_w(0);
    __syncthreads();
    int buf = 0;
    for (int c0 = 0; c0 < K; c0 += C2_KT, buf ^= 1) {
      const bool more = c0 + C2_KT < K;
      if (more) load_w(c0 + C2_KT);             // next chunk's weights travel while this chunk is multiplied
      const int nslab = min(C2_KT, K - c0) / 16;
      for (int s0 = 0; s0 < nslab; s0 += SU) {
        float4 av[SU][2][NC];
#pragma unroll
        for (int su = 0; su < SU; ++su)
#pragma unroll
          for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < NC; ++c) {
              const int kb = c0 + 16 * (s0 + su) + 4 * t;
              av[su][r][c] = (s0 + su < nslab && rowp[r][c] != nullptr)
                                 ? __ldg(reinterpret_cast<const float4*>(rowp[r][c] + kb))
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
        for (int su = 0; su < SU; ++su) {
          if (s0 + su >= nslab) break;
          const int rb = 16 * (s0 + su) + 4 * t;   // staged row of this lane's first k
          uint32_t ah[2][NC][4], al[2][NC][4];
#pragma unroll
          for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < NC; ++c) {
              split_tf32(av[su][r][c].x, ah[r][c][0], al[r][c][0]);
              split_tf32(av[su][r][c].y, ah[r][c][1], al[r][c][1]);
              split_tf32(av[su][r][c].z, ah[r][c][2], al[r][c][2]);
              split_tf32(av[su][r][c].w, ah[r][c][3], al[r][c][3]);
            }
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            uint32_t bh[4], bl[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              bh[i] = __float_as_uint(Whi[buf][rb + i][8 * nt + g]);
              bl[i] = __float_as_uint(Wlo[buf][rb + i][8 * nt + g]);
            }
#pragma unroll
            for (int c = 0; c < NC; ++c) {
              // k8 block 1: lane's k = (4t, 4t+1); block 2: (4t+2, 4t+3)
#pragma unroll
              for (int hb = 0; hb < 2; ++hb) {
                const int i0 = 2 * hb, i1 = 2 * hb + 1;
                mma_tf32(acc[c][nt], ah[0][c][i0], ah[1][c][i0], ah[0][c][i1], ah[1][c][i1], bh[i0], bh[i1]);
                mma_tf32(acc[c][nt], al[0][c][i0], al[1][c][i0], al[0][c][i1], al[1][c][i1], bh[i0], bh[i1]);
                mma_tf32(acc[c][nt], ah[0][c][i0], ah[1][c][i0], ah[0][c][i1], ah[1][c][i1], bl[i0], bl[i1]);
              }
            }
          }
        }
      }
      if (more) store_w(buf ^ 1);
      __syncthreads();
    }
    // bias path: sum_e basis_e (*) b2p  (K = F, plain FFMA on the accumulator fragments)
    const float* bg = p.b2p[grp] + ci.boff;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      if (sidx[r][which] < 0) continue;
      const float* bs = p.Bsum + (size_t)sidx[r][which] * U + ci.uoff;
      for (int uk = 0; uk < ci.F; ++uk) {
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int o = 8 * nt + 2 * t + j;
            if (o < O) {
              const float wv = bg[uk * O + o];
#pragma unroll
              for (int c = 0; c < NC; ++c) acc[c][nt][2 * r + j] += bs[c * ci.F + uk] * wv;
            }
          }
      }
    }
  }
  // accumulator fragment -> [node][feature] tile: c0:(row g, col 2t) c1:(g, 2t+1) c2:(g+8, 2t) c3:(g+8, 2t+1)
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = 16 * w + g + 8 * (i >> 1), o = 8 * nt + 2 * t + (i & 1);
        if (o < O) S.Out[row][ci.col0 + (NC == 3 ? 3 * o + c : o)] = acc[c][nt][i];
      }
}

__global__ void __launch_bounds__(C2_THREADS, 3) k_conv_contract2(ConArgs p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  C2Smem& S = *reinterpret_cast<C2Smem*>(smem_raw);
  const int nl = p.lig1 - p.lig0;
  const int nblk_l = (nl + C2_TM - 1) / C2_TM;
  const bool lig = (int)blockIdx.x < nblk_l;
  const int t0 = lig ? p.lig0 + blockIdx.x * C2_TM : p.rec0 + (blockIdx.x - nblk_l) * C2_TM;
  const int tend = lig ? p.lig1 : p.rec1;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2;
  for (int i = threadIdx.x; i < C2_TM * D; i += C2_THREADS) S.Out[i / D][i % D] = 0.f;
  int sidx[2][2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int nt_ = t0 + 16 * w + g + 8 * r;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      int s = -1;
      if (nt_ < tend) {
        const int seg = 2 * ((lig ? 0 : p.NL) + nt_) + which;
        if (p.seg_cnt[seg] > 0) s = p.seg_sidx[seg];
      }
      sidx[r][which] = s;
    }
  }
  if (threadIdx.x < C2_TM) {
    const int nt_ = t0 + threadIdx.x;
    float cn = 1.f;
    if (nt_ < tend) {
      const int seg = 2 * ((lig ? 0 : p.NL) + nt_);
      cn = fmaxf((float)(p.seg_cnt[seg] + p.seg_cnt[seg + 1]), 1.f);
    }
    S.Cnt[threadIdx.x] = cn;
  }
  for (int k = 0; k < p.li.ncls; ++k) {
    const ClassInfo ci = p.li.cls[k];
    if (ci.O == 24) contract2_class<24, 1>(p, ci, lig, p.li.U, sidx, S, w, lane);
    else contract2_class<6, 3>(p, ci, lig, p.li.U, sidx, S, w, lane);
  }
  __syncthreads();
  // mean over edges, batch-norm affine (eval), residual with the zero-padded input (tensor_layers.py:159-166)
  for (int i = threadIdx.x; i < C2_TM * D; i += C2_THREADS) {
    const int q = i / D, f = i % D, nt_ = t0 + q;
    if (nt_ >= tend) continue;
    const size_t row = (size_t)((lig ? 0 : p.NL) + nt_) * D;
    float v = 0.f;
    if (f < p.li.dout) v = (S.Out[q][f] / S.Cnt[q]) * p.bn_scale[f] + p.bn_shift[f] + p.x_in[row + f];
    p.x_out[row + f] = v;
  }
}

cudaError_t contract2_configure() {
  return cudaFuncSetAttribute(k_conv_contract2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(C2Smem));
}

void launch_conv_contract2(DdkCtx* c, const ConArgs& q, cudaStream_t st) {
  const int blocks = (q.lig1 - q.lig0 + C2_TM - 1) / C2_TM + (q.rec1 - q.rec0 + C2_TM - 1) / C2_TM;
  LaunchScope ls(c, PC_CONTRACT, st);
  k_conv_contract2<<<blocks, C2_THREADS, sizeof(C2Smem), st>>>(q);
}

}  // namespace ddk
