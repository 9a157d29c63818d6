// Tensor-product convolution layers (the hot loop): /root/reference/models/tensor_layers.py:147-168
// (TensorProductConvLayer.forward) with FasterTensorProduct (:65-116) and the per-edge radial MLP
// (/root/reference/models/layers.py:15-22), called five times per reverse step from score_model.py:227-230.
//
// Re-associated evaluation (exact algebra).  For a node s and edge group g
//     sum_e TP(x_dst, sh_e; W2 h_e + b2) = W2p (*) A_s + b2p (*) Bsum_s,
//     A_s[u][j] = sum_e basis_e[u] * h_e[j]   (U x 72 outer-product accumulator),   Bsum_s[u] = sum_e basis_e[u]
// so the 72 x W second MLP layer runs once per (node, group) instead of once per edge.  A layer is four launches:
//   k_node_proj     (here)          : per node, the x_src / x_dst parts of the first MLP layer
//   k_edge_hidden   (ddk_hidden.cu) : per listed edge, h_e = relu(edge part + the two node parts)
//   k_conv_fused    (ddk_conv3.cu)  : rank-1 accumulation of A_s slices in registers + contraction with W2p slices
//   k_conv_finalize (ddk_conv3.cu)  : sum of the slice partials, mean, batch-norm affine, residual
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
  int N;
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
      p.proj[(node * 4 + s) * HID + j] = acc;
    }
  }
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
  p.N = c->N;
  int blocks = (c->NL + PROJ_NODES - 1) / PROJ_NODES + (c->NR + PROJ_NODES - 1) / PROJ_NODES;
  blocks = std::max(2, std::min(blocks, 4 * c->sm_count));   // persistent blocks, see k_node_proj
  LaunchScope ls(c, PC_PROJ, st);
  if (x0_out != nullptr) k_node_proj<true><<<blocks, 288, 0, st>>>(p);
  else k_node_proj<false><<<blocks, 288, 0, st>>>(p);
}

void launch_conv_layer(DdkCtx* c, int layer, const float* x_in, float* x_out, cudaStream_t st, int mode) {
  launch_edge_hidden(c, layer, st, mode);
  if (conv_path() == 2) launch_conv_tcr(c, layer, x_in, x_out, st, mode);
  else launch_conv_fused(c, layer, x_in, x_out, st, mode);
}

}  // namespace ddk
