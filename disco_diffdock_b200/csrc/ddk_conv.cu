// Tensor-product convolution layers (the hot loop): /root/reference/models/tensor_layers.py:147-168
// (TensorProductConvLayer.forward) with FasterTensorProduct (:65-116) and the per-edge radial MLP
// (/root/reference/models/layers.py:15-22), called five times per reverse step from score_model.py:227-230.
//
// Re-associated evaluation (exact algebra).  For a node s and edge group g
//     sum_e TP(x_dst, sh_e; W2 h_e + b2) = W2p (*) A_s + b2p (*) Bsum_s,
//     A_s[u][j] = sum_e basis_e[u] * h_e[j]   (U x 72 outer-product accumulator),   Bsum_s[u] = sum_e basis_e[u]
// so the 72 x W second MLP layer runs once per (node, group) instead of once per edge:
//   k_node_proj    : per node, the x_src / x_dst parts of the first MLP layer (W1[:,24:48] x_s + b1, W1[:,48:72] x_d)
//   k_conv_accum   : one CTA per (node, group) segment: h_e, basis_e, rank-1 updates of A in registers -> scratch
//   k_conv_contract: per tile of 32 nodes: A (*) W2p for both groups, mean, batch-norm affine, residual
#include <algorithm>

#include "ddk_conv.cuh"

namespace ddk {

// ------------------------------------------------------------------------------------------------ node projections
struct ProjArgs {
  int NL, NR;
  const float* x_in;              // [N][84] (unused when FIRST)
  float* x0_out;                  // FIRST: the initial node embedding is written here
  const float* lig_static; const float* rec_static; const float* tb;
  const int* lig_graph; const int* rec_graph;
  const float* lig_uncond; const float* rec_uncond; const float* uncond_emb;   // [5][24]
  const float* W1[4];             // fc.g.0.weight of the layer, [72][72]
  const float* b1[4];
  float* proj;                    // [N][4][72]: {src group a, src group b, dst group a', dst group b'}
  int sliced, N;                  // sliced > 0: [72 / J][N][4][J] with J = sliced (hidden-unit slices of the fused conv kernel)
};

constexpr int PROJ_NODES = 16;

// lig nodes: src of groups (0,1), dst of groups (0,3);  rec nodes: src of (2,3), dst of (2,1)
template <bool FIRST>
__global__ void __launch_bounds__(288) k_node_proj(ProjArgs p) {
  __shared__ float sW[4][NS][HID];     // [slot][k][j]
  __shared__ float sB[2][HID];
  __shared__ float sx[PROJ_NODES][NS];
  // persistent blocks: the first gridDim.x / 2 (rounded by node share) walk the ligand tiles, the others the receptor tiles,
  // so the 27 KB of weights are staged once per block instead of once per 16 nodes
  const int nblk_l = (p.NL + PROJ_NODES - 1) / PROJ_NODES, nblk_r = (p.NR + PROJ_NODES - 1) / PROJ_NODES;
  int gl = (int)(((long long)gridDim.x * nblk_l + nblk_l + nblk_r - 1) / (nblk_l + nblk_r));
  gl = max(1, min(gl, (int)gridDim.x - 1));
  const bool lig = (int)blockIdx.x < gl;
  const int tile0 = lig ? blockIdx.x : blockIdx.x - gl, tstep = lig ? gl : (int)gridDim.x - gl, ntile = lig ? nblk_l : nblk_r;
  const int gs[2][4] = {{0, 1, 0, 3}, {2, 3, 2, 1}};
  const int* g4 = gs[lig ? 0 : 1];
  for (int i = threadIdx.x; i < 4 * NS * HID; i += blockDim.x) {
    // coalesced over the weight rows: i -> (slot s, hidden unit j, input k)
    int s = i / (NS * HID), j = (i / NS) % HID, k = i % NS;
    int col = (s < 2 ? NS : 2 * NS) + k;
    sW[s][k][j] = p.W1[g4[s]][j * HID + col];
  }
  for (int i = threadIdx.x; i < 2 * HID; i += blockDim.x) sB[i / HID][i % HID] = p.b1[g4[i / HID]][i % HID];
  for (int tile = tile0; tile < ntile; tile += tstep) {
    const int n0 = tile * PROJ_NODES;
    const int nn = min(PROJ_NODES, (lig ? p.NL : p.NR) - n0);
    __syncthreads();                      // weights staged / previous tile consumed
    for (int i = threadIdx.x; i < nn * NS; i += blockDim.x) {
      int q = i / NS, k = i % NS, n = n0 + q;
      float v;
      if (FIRST) {
        int g = lig ? p.lig_graph[n] : p.rec_graph[n];
        const float* st = lig ? p.lig_static : p.rec_static;
        const float* un = lig ? p.lig_uncond : p.rec_uncond;
        v = st[n * NS + k] + p.tb[((size_t)g * TB_COUNT + (lig ? TB_LIG_NODE : TB_REC_NODE)) * NS + k];
        if (un != nullptr) v += un[n] * p.uncond_emb[(lig ? 0 : 1) * NS + k];
      } else {
        v = p.x_in[(size_t)((lig ? 0 : p.NL) + n) * D + k];
      }
      sx[q][k] = v;
    }
    __syncthreads();
    if (FIRST) {
      for (int i = threadIdx.x; i < nn * D; i += blockDim.x) {
        int q = i / D, f = i % D;
        p.x0_out[(size_t)((lig ? 0 : p.NL) + n0 + q) * D + f] = f < NS ? sx[q][f] : 0.f;
      }
    }
    int s = threadIdx.x / HID, j = threadIdx.x % HID;   // 288 threads = 4 slots x 72
    for (int q = 0; q < nn; ++q) {
      float acc = s < 2 ? sB[s][j] : 0.f;
#pragma unroll
      for (int k = 0; k < NS; ++k) acc += sW[s][k][j] * sx[q][k];
      const size_t node = (size_t)((lig ? 0 : p.NL) + n0 + q);
      if (p.sliced) p.proj[(((size_t)(j / p.sliced) * p.N + node) * 4 + s) * p.sliced + (j % p.sliced)] = acc;
      else p.proj[(node * 4 + s) * HID + j] = acc;
    }
  }
}

// ------------------------------------------------------------------------------------------------ accumulate (v1)
template <int LV>
struct AccSmem {
  int slot[KC];
  int dst[KC];
  alignas(16) float Ea[KC][EA];
  alignas(16) float Sh[KC][4];
  alignas(16) float Xd[KC][AccCfg<LV>::DINP];
  alignas(16) float H[KC][HID];
  alignas(16) float B[KC][AccCfg<LV>::BS];
};

template <int LV>
__global__ void __launch_bounds__(ACC_THREADS, 2) k_conv_accum(AccArgs p) {
  constexpr int U = AccCfg<LV>::U;
  constexpr int NA = AccCfg<LV>::NA;
  constexpr int BS = AccCfg<LV>::BS;
  constexpr int DIN = AccCfg<LV>::DINP;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  AccSmem<LV>& S = *reinterpret_cast<AccSmem<LV>*>(smem_raw);
  int (&s_slot)[KC] = S.slot;
  int (&s_dst)[KC] = S.dst;
  float (&sEa)[KC][EA] = S.Ea;
  float (&sSh)[KC][4] = S.Sh;
  float (&sXd)[KC][DIN] = S.Xd;
  float (&sH)[KC][HID] = S.H;
  float (&sB)[KC][BS] = S.B;

  const int seg = p.seg_order[blockIdx.x];
  const int n = p.seg_cnt[seg];
  if (n == 0) return;
  const int sidx = p.seg_sidx[seg];
  const int node = seg >> 1, which = seg & 1;
  const int g = node < p.NL ? which : 2 + which;
  const int dslot = (g == 1 || g == 3) ? 3 : 2;
  const int base = p.seg_base[seg];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;

  // role A (first MLP layer): thread -> hidden unit j2, edge phase eg
  const int j2 = tid % HID, eg = tid / HID;
  float w1a[EA];
  {
    const float4* wr = reinterpret_cast<const float4*>(p.W1[g] + j2 * HID);
#pragma unroll
    for (int q = 0; q < EA / 4; ++q) {
      float4 v = wr[q];
      w1a[4 * q] = v.x; w1a[4 * q + 1] = v.y; w1a[4 * q + 2] = v.z; w1a[4 * q + 3] = v.w;
    }
  }
  const float ps = p.proj[((size_t)node * 4 + which) * HID + j2];
  // role B (basis): thread -> basis row u3
  int btype, bi0, bm;
  basis_desc<LV>(tid < U ? tid : 0, btype, bi0, bm);
  float bsum = 0.f;

  float acc[NA][8];
#pragma unroll
  for (int a = 0; a < NA; ++a)
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) acc[a][jj] = 0.f;

  for (int c0 = 0; c0 < n; c0 += KC) {
    const int kc = min(KC, n - c0);
    if (tid < kc) {
      int2 ent = p.seg_list[base + c0 + tid];
      s_slot[tid] = ent.x;
      s_dst[tid] = ent.y;
    }
    __syncthreads();
    // gather edge embedding, harmonics, destination features and destination projection
    for (int i = tid; i < kc * (EA / 4); i += ACC_THREADS) {
      int e = i / (EA / 4), q = i % (EA / 4);
      reinterpret_cast<float4*>(&sEa[e][0])[q] = reinterpret_cast<const float4*>(p.ea_pool + (size_t)s_slot[e] * EA)[q];
    }
    if (tid < kc) reinterpret_cast<float4*>(&sSh[tid][0])[0] = p.sh_pool[s_slot[tid]];
    for (int i = tid; i < kc * (DIN / 4); i += ACC_THREADS) {
      int e = i / (DIN / 4), q = i % (DIN / 4);
      reinterpret_cast<float4*>(&sXd[e][0])[q] = reinterpret_cast<const float4*>(p.x + (size_t)s_dst[e] * D)[q];
    }
    for (int i = tid; i < kc * (HID / 4); i += ACC_THREADS) {
      int e = i / (HID / 4), q = i % (HID / 4);
      reinterpret_cast<float4*>(&sH[e][0])[q] =
          reinterpret_cast<const float4*>(p.proj + ((size_t)s_dst[e] * 4 + dslot) * HID)[q];
    }
    __syncthreads();
    // first MLP layer: h = relu(W1[:, :24] ea + (W1[:,24:48] x_s + b1) + W1[:,48:72] x_d)
    for (int e = eg; e < kc; e += ACC_THREADS / HID) {
      float h = ps + sH[e][j2];
#pragma unroll
      for (int q = 0; q < EA / 4; ++q) {
        float4 v = reinterpret_cast<const float4*>(&sEa[e][0])[q];
        h += w1a[4 * q] * v.x + w1a[4 * q + 1] * v.y + w1a[4 * q + 2] * v.z + w1a[4 * q + 3] * v.w;
      }
      sH[e][j2] = fmaxf(h, 0.f);
    }
    // basis functions of the chunk (raw products; constants live in the packed weights)
    if (tid < BS) {
      for (int e = 0; e < kc; ++e) {
        float b = 0.f;
        if (tid < U) {
          const float* xd = &sXd[e][0];
          const float* sh = &sSh[e][0];
          if (btype == 0) {
            b = xd[bi0] * sh[bm];
          } else if (btype == 1) {
            b = xd[bi0] * sh[1] + xd[bi0 + 1] * sh[2] + xd[bi0 + 2] * sh[3];
          } else {
            int c = bm - 1, c1 = (c + 1) % 3, c2 = (c + 2) % 3;
            b = xd[bi0 + c1] * sh[1 + c2] - xd[bi0 + c2] * sh[1 + c1];
          }
          bsum += b;
        }
        sB[e][tid] = b;
      }
    }
    __syncthreads();
    // rank-1 updates: acc[u][j] += basis[u] * h[j]
#pragma unroll 2
    for (int e = 0; e < kc; ++e) {
      float4 h0 = reinterpret_cast<const float4*>(&sH[e][8 * w])[0];
      float4 h1 = reinterpret_cast<const float4*>(&sH[e][8 * w])[1];
#pragma unroll
      for (int a = 0; a < NA; ++a) {
        float b = sB[e][lane + 32 * a];
        acc[a][0] += b * h0.x; acc[a][1] += b * h0.y; acc[a][2] += b * h0.z; acc[a][3] += b * h0.w;
        acc[a][4] += b * h1.x; acc[a][5] += b * h1.y; acc[a][6] += b * h1.z; acc[a][7] += b * h1.w;
      }
    }
    __syncthreads();
  }
  float* Aout = p.A + (size_t)sidx * U * HID;
#pragma unroll
  for (int a = 0; a < NA; ++a) {
    int u = lane + 32 * a;
    if (u < U) {
      float4* dst = reinterpret_cast<float4*>(Aout + (size_t)u * HID + 8 * w);
      dst[0] = make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
      dst[1] = make_float4(acc[a][4], acc[a][5], acc[a][6], acc[a][7]);
    }
  }
  if (tid < U) p.Bsum[(size_t)sidx * U + tid] = bsum;
}

// ------------------------------------------------------------------------------------------------ contract
constexpr int CON_TM = 32;        // nodes per CTA
constexpr int CON_KT = 128;       // K rows of W2p staged per pass

template <int O, int NCOMP>
__device__ __forceinline__ void contract_class(const ConArgs& p, const ClassInfo& ci, bool lig, int U, const int (&sidx)[4][2],
                                               float* sW, float (*sOut)[D], int w, int lane) {
  constexpr int OP = (O == 24) ? 28 : 8;   // padded row stride of the staged weights (floats)
  constexpr int R = 4 * NCOMP;
  float acc[R][O];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int o = 0; o < O; ++o) acc[r][o] = 0.f;
  const int K = ci.F * HID;
  for (int which = 0; which < 2; ++which) {
    const int g = lig ? which : 2 + which;
    const float* Wg = p.W2p[g] + ci.woff;
    for (int c0 = 0; c0 < K; c0 += CON_KT) {
      __syncthreads();
      for (int i = threadIdx.x; i < CON_KT * O; i += blockDim.x) {
        int r = i / O, o = i % O;
        sW[r * OP + o] = (c0 + r < K) ? Wg[(size_t)(c0 + r) * O + o] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int q = 0; q < CON_KT / 32; ++q) {
        const int kk = c0 + lane + 32 * q;
        if (kk < K) {
          float wv[O];
          const float* wr = sW + (lane + 32 * q) * OP;
#pragma unroll
          for (int o4 = 0; o4 < (O + 3) / 4; ++o4) {
            float4 v = reinterpret_cast<const float4*>(wr)[o4];
            if (4 * o4 + 0 < O) wv[4 * o4 + 0] = v.x;
            if (4 * o4 + 1 < O) wv[4 * o4 + 1] = v.y;
            if (4 * o4 + 2 < O) wv[4 * o4 + 2] = v.z;
            if (4 * o4 + 3 < O) wv[4 * o4 + 3] = v.w;
          }
          float av[R];
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < NCOMP; ++c) {
              int s = sidx[r][which];
              av[r * NCOMP + c] = (s >= 0) ? p.A[((size_t)s * U + ci.uoff + c * ci.F) * HID + kk] : 0.f;
            }
#pragma unroll
          for (int r = 0; r < R; ++r)
#pragma unroll
            for (int o = 0; o < O; ++o) acc[r][o] += av[r] * wv[o];
        }
      }
    }
    // bias path: sum_e basis_e (*) b2p
    const float* bg = p.b2p[g] + ci.boff;
    for (int uk = lane; uk < ci.F; uk += 32) {
      float wv[O];
#pragma unroll
      for (int o = 0; o < O; ++o) wv[o] = bg[uk * O + o];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < NCOMP; ++c) {
          int s = sidx[r][which];
          float a = (s >= 0) ? p.Bsum[(size_t)s * U + ci.uoff + c * ci.F + uk] : 0.f;
#pragma unroll
          for (int o = 0; o < O; ++o) acc[r * NCOMP + c][o] += a * wv[o];
        }
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < NCOMP; ++c)
#pragma unroll
      for (int o = 0; o < O; ++o) {
        float v = warp_sum(acc[r * NCOMP + c][o]);
        if (lane == 0) sOut[4 * w + r][ci.col0 + (NCOMP == 3 ? 3 * o + c : o)] = v;
      }
}

__global__ void __launch_bounds__(256) k_conv_contract(ConArgs p) {
  __shared__ __align__(16) float sW[CON_KT * 28];
  __shared__ float sOut[CON_TM][D];
  __shared__ float sCnt[CON_TM];
  const int nl = p.lig1 - p.lig0;
  const int nblk_l = (nl + CON_TM - 1) / CON_TM;
  const bool lig = (int)blockIdx.x < nblk_l;
  const int t0 = lig ? p.lig0 + blockIdx.x * CON_TM : p.rec0 + (blockIdx.x - nblk_l) * CON_TM;   // index within type
  const int tend = lig ? p.lig1 : p.rec1;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < CON_TM * D; i += blockDim.x) sOut[i / D][i % D] = 0.f;
  int sidx[4][2];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    int nt = t0 + 4 * w + r;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      int s = -1;
      if (nt < tend) {
        int seg = 2 * ((lig ? 0 : p.NL) + nt) + which;
        if (p.seg_cnt[seg] > 0) s = p.seg_sidx[seg];
      }
      sidx[r][which] = s;
    }
  }
  if (threadIdx.x < CON_TM) {
    int nt = t0 + threadIdx.x;
    float cn = 1.f;
    if (nt < tend) {
      int seg = 2 * ((lig ? 0 : p.NL) + nt);
      cn = fmaxf((float)(p.seg_cnt[seg] + p.seg_cnt[seg + 1]), 1.f);
    }
    sCnt[threadIdx.x] = cn;
  }
  __syncthreads();
  for (int k = 0; k < p.li.ncls; ++k) {
    const ClassInfo ci = p.li.cls[k];
    if (ci.O == 24) contract_class<24, 1>(p, ci, lig, p.li.U, sidx, sW, sOut, w, lane);
    else contract_class<6, 3>(p, ci, lig, p.li.U, sidx, sW, sOut, w, lane);
  }
  __syncthreads();
  // mean over edges, batch-norm affine (eval), residual with the zero-padded input (tensor_layers.py:159-166)
  for (int i = threadIdx.x; i < CON_TM * D; i += blockDim.x) {
    int q = i / D, f = i % D, nt = t0 + q;
    if (nt >= tend) continue;
    size_t row = (size_t)((lig ? 0 : p.NL) + nt) * D;
    float v = 0.f;
    if (f < p.li.dout) v = (sOut[q][f] / sCnt[q]) * p.bn_scale[f] + p.bn_shift[f] + p.x_in[row + f];
    p.x_out[row + f] = v;
  }
}

// ------------------------------------------------------------------------------------------------ launchers
cudaError_t conv_configure() {
  cudaError_t e;
  e = cudaFuncSetAttribute(k_conv_accum<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AccSmem<0>));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_conv_accum<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AccSmem<1>));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_conv_accum<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AccSmem<2>));
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_conv_accum<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AccSmem<3>));
}

void launch_node_proj(DdkCtx* c, int layer, const float* x_in, float* x0_out, cudaStream_t st) {
  ProjArgs p;
  p.NL = c->NL; p.NR = c->NR;
  p.x_in = x_in; p.x0_out = x0_out;
  p.lig_static = ptr<float>(c->b_lig_static); p.rec_static = ptr<float>(c->b_rec_static);
  p.tb = ptr<float>(c->b_tb);
  p.lig_graph = ptr<int>(c->b_lig_graph); p.rec_graph = ptr<int>(c->b_rec_graph);
  p.lig_uncond = c->cfg.has_unconditional ? c->lig_uncond : nullptr;
  p.rec_uncond = c->cfg.has_unconditional ? c->rec_uncond : nullptr;
  p.uncond_emb = W(c, DDK_W_UNCOND);
  for (int g = 0; g < 4; ++g) {
    p.W1[g] = W(c, conv_id(layer, DDK_WL_W1 + g));
    p.b1[g] = W(c, conv_id(layer, DDK_WL_B1 + g));
  }
  p.proj = ptr<float>(c->b_proj);
  p.sliced = 0; p.N = c->N;
  int blocks = (c->NL + PROJ_NODES - 1) / PROJ_NODES + (c->NR + PROJ_NODES - 1) / PROJ_NODES;
  blocks = std::max(2, std::min(blocks, 4 * c->sm_count));   // persistent blocks, see k_node_proj
  LaunchScope ls(c, PC_PROJ, st);
  if (x0_out != nullptr) k_node_proj<true><<<blocks, 288, 0, st>>>(p);
  else k_node_proj<false><<<blocks, 288, 0, st>>>(p);
}

void launch_conv_layer(DdkCtx* c, int layer, const float* x_in, float* x_out, cudaStream_t st, int mode) {
  if (!c->conv_v1 && !c->conv_v2) {
    launch_edge_hidden(c, layer, st, mode);
    launch_conv_fused(c, layer, x_in, x_out, st, mode);
    return;
  }
  const LayerInfo& li = c->layers[layer];
  for (const Chunk& ch : c->chunks) {
    AccArgs a;
    a.NL = c->NL;
    a.seg_order = ptr<int>(c->b_seg_order) + ch.order_off;
    a.seg_sidx = ptr<int>(c->b_seg_sidx);
    a.seg_base = ptr<int>(c->b_seg_base); a.seg_cnt = ptr<int>(c->b_seg_cnt); a.seg_list = ptr<int2>(c->b_seg_list);
    a.x = x_in; a.proj = ptr<float>(c->b_proj);
    a.ea_pool = ptr<float>(c->b_ea_pool); a.sh_pool = ptr<float4>(c->b_sh_pool);
    for (int g = 0; g < 4; ++g) a.W1[g] = W(c, conv_id(layer, DDK_WL_W1 + g));
    a.A = ptr<float>(c->b_A); a.Bsum = ptr<float>(c->b_Bsum);
    if (!c->conv_v1) {
      launch_conv_accum2(c, li, ch, a, st);
    } else {
    LaunchScope ls(c, PC_ACC0 + li.lv, st);
    switch (li.lv) {
      case 0: k_conv_accum<0><<<ch.nseg, ACC_THREADS, sizeof(AccSmem<0>), st>>>(a); break;
      case 1: k_conv_accum<1><<<ch.nseg, ACC_THREADS, sizeof(AccSmem<1>), st>>>(a); break;
      case 2: k_conv_accum<2><<<ch.nseg, ACC_THREADS, sizeof(AccSmem<2>), st>>>(a); break;
      default: k_conv_accum<3><<<ch.nseg, ACC_THREADS, sizeof(AccSmem<3>), st>>>(a); break;
    }
    }
    ConArgs q;
    q.NL = c->NL;
    q.lig0 = ch.lig0; q.lig1 = ch.lig1; q.rec0 = ch.rec0; q.rec1 = ch.rec1;
    q.seg_sidx = ptr<int>(c->b_seg_sidx); q.seg_cnt = ptr<int>(c->b_seg_cnt);
    q.A = ptr<float>(c->b_A); q.Bsum = ptr<float>(c->b_Bsum);
    for (int g = 0; g < 4; ++g) {
      q.W2p[g] = W(c, conv_id(layer, DDK_WL_W2P + g));
      q.b2p[g] = W(c, conv_id(layer, DDK_WL_B2P + g));
    }
    q.bn_scale = W(c, conv_id(layer, DDK_WL_BN_SCALE));
    q.bn_shift = W(c, conv_id(layer, DDK_WL_BN_SHIFT));
    q.x_in = x_in; q.x_out = x_out;
    q.li = li;
    int blocks = (ch.lig1 - ch.lig0 + CON_TM - 1) / CON_TM + (ch.rec1 - ch.rec0 + CON_TM - 1) / CON_TM;
    if (!c->conv_v1) {
      launch_conv_contract2(c, q, st);
    } else {
      LaunchScope ls(c, PC_CONTRACT, st);
      k_conv_contract<<<blocks, 256, 0, st>>>(q);
    }
  }
}

}  // namespace ddk
