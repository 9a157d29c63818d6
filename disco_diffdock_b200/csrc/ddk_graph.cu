// Graph construction and embedding kernels:
//   - per-batch setup: atom / residue encoders (step-invariant part), receptor contact-edge geometry
//   - per step: sigma-embedding biases, radius-graph / cross-graph segment lists, edge embeddings + harmonics
// Reference: /root/reference/models/score_model.py:191-225 (embed), 310-408 (build_*_conv_graph),
//            /root/reference/models/layers.py:140-149 (AtomEncoder).
#include "ddk_device.cuh"

namespace ddk {

// ------------------------------------------------------------------------------------------------ setup
// Ligand atom encoder, step-invariant part: W[:, :24] . sum_i Emb_i[x_i] + W[:, 56:56+L] . latent + b.
// (the sigma-embedding columns 24:56 are added per step through TB_LIG_NODE)
__global__ void k_setup_lig_nodes(int NL, const int32_t* __restrict__ lig_x, const float* __restrict__ tables,
                                  const float* __restrict__ Wn, const float* __restrict__ bn, int wcols, int L,
                                  const float* __restrict__ latent, float* __restrict__ out) {
  const int tab_off[16] = {0, 119, 123, 135, 147, 155, 165, 171, 177, 179, 187, 189, 191, 193, 195, 197};
  int n = blockIdx.x * blockDim.y + threadIdx.y;
  int o = threadIdx.x;  // 0..23
  __shared__ float emb[8][NS];
  if (n < NL) {
    float e = 0.f;
#pragma unroll
    for (int i = 0; i < LIG_CAT; ++i) e += tables[(tab_off[i] + lig_x[n * LIG_CAT + i]) * NS + o];
    emb[threadIdx.y][o] = e;
  }
  __syncthreads();
  if (n >= NL) return;
  float acc = bn[o];
  for (int k = 0; k < NS; ++k) acc += Wn[o * wcols + k] * emb[threadIdx.y][k];
  for (int l = 0; l < L; ++l) acc += Wn[o * wcols + NS + SE + l] * latent[n * L + l];
  out[n * NS + o] = acc;
}

// Residue encoder, step-invariant part: one warp per residue; W is [24][24 + 1280 + 32 + L].
__global__ void k_setup_rec_nodes(int NR, const float* __restrict__ rec_x, const float* __restrict__ table,
                                  const float* __restrict__ Wn, const float* __restrict__ bn, int wcols, int L,
                                  const float* __restrict__ latent, float* __restrict__ out) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= NR) return;
  const float* xr = rec_x + (size_t)warp * REC_X;
  int aa = (int)xr[0];
  float acc[NS];
#pragma unroll
  for (int o = 0; o < NS; ++o) acc[o] = 0.f;
  for (int k = lane; k < NS + ESM; k += 32) {
    float v = (k < NS) ? table[aa * NS + k] : xr[1 + (k - NS)];
#pragma unroll
    for (int o = 0; o < NS; ++o) acc[o] += Wn[o * wcols + k] * v;
  }
  for (int l = lane; l < L; l += 32) {
    float v = latent[warp * L + l];
#pragma unroll
    for (int o = 0; o < NS; ++o) acc[o] += Wn[o * wcols + NS + ESM + SE + l] * v;
  }
#pragma unroll
  for (int o = 0; o < NS; ++o) acc[o] = warp_sum(acc[o]);
  if (lane < NS) {
    float r = 0.f;
#pragma unroll
    for (int o = 0; o < NS; ++o) r = (lane == o) ? acc[o] : r;
    out[warp * NS + lane] = r + bn[lane];
  }
}

// Receptor contact edges: harmonics (static) and the distance / latent part of the first edge-MLP layer.
__global__ void k_setup_rr_edges(int ER, const int* __restrict__ rr_src, const int* __restrict__ rr_dst,
                                 const float* __restrict__ rec_pos, const float* __restrict__ W1, int wcols, int L,
                                 const float* __restrict__ latent, const float* __restrict__ sm,
                                 float* __restrict__ rr_pre, float4* __restrict__ sh_rr) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ER) return;
  int s = rr_src[e], d = rr_dst[e];
  float vx = rec_pos[d * 3 + 0] - rec_pos[s * 3 + 0], vy = rec_pos[d * 3 + 1] - rec_pos[s * 3 + 1],
        vz = rec_pos[d * 3 + 2] - rec_pos[s * 3 + 2];
  float nrm;
  sh_rr[e] = sh_l01(vx, vy, vz, &nrm);
  float pre[EA];
#pragma unroll
  for (int o = 0; o < EA; ++o) pre[o] = 0.f;
  for (int k = 0; k < DE; ++k) {
    float g = smear1(sm, nrm, k);
#pragma unroll
    for (int o = 0; o < EA; ++o) pre[o] += W1[o * wcols + SE + k] * g;
  }
  for (int l = 0; l < L; ++l) {
    float a = latent[s * L + l], b = latent[d * L + l];
#pragma unroll
    for (int o = 0; o < EA; ++o) pre[o] += W1[o * wcols + SE + DE + l] * a + W1[o * wcols + SE + DE + L + l] * b;
  }
#pragma unroll
  for (int o = 0; o < EA; ++o) rr_pre[(size_t)e * EA + o] = pre[o];
}

// Static entries of the segment lists: bond e -> (slot e, dst atom), receptor contact e -> (slot_rr + e, NL + dst residue)
__global__ void k_fill_static_lists(int EB, int ER, int NL, int slot_rr, const int* __restrict__ static_pos,
                                    const int* __restrict__ bond_dst, const int* __restrict__ rr_dst, int2* __restrict__ seg_list) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < EB) seg_list[static_pos[e]] = make_int2(e, bond_dst[e]);
  else if (e < EB + ER) seg_list[static_pos[e]] = make_int2(slot_rr + (e - EB), NL + rr_dst[e - EB]);
}

// ------------------------------------------------------------------------------------------------ per step
struct TbArgs {
  const float* Wsrc[TB_COUNT];
  const float* bias[TB_COUNT];
  int cols[TB_COUNT];
  int col0[TB_COUNT];
};

// tb[g][k][o] = bias_k[o] + W_k[o][col0_k : col0_k + 32] . sigma_emb[g]
__global__ void k_step_consts(int B, const float* __restrict__ semb, TbArgs a, float* __restrict__ tb) {
  int g = blockIdx.x;
  int t = threadIdx.x;  // TB_COUNT * 24
  __shared__ float s[SE];
  if (t < SE) s[t] = semb[g * SE + t];
  __syncthreads();
  int k = t / NS, o = t % NS;
  if (k >= TB_COUNT) return;
  const float* wr = a.Wsrc[k] + o * a.cols[k] + a.col0[k];
  float acc = a.bias[k] ? a.bias[k][o] : 0.f;
#pragma unroll 8
  for (int i = 0; i < SE; ++i) acc += wr[i] * s[i];
  tb[((size_t)g * TB_COUNT + k) * NS + o] = acc;
}

struct ListArgs {
  int NL, NR;
  const int* lig_ptr; const int* rec_ptr; const int* lig_graph; const int* rec_graph;
  const int64_t* ll_off; const int64_t* lr_off;
  const int* seg_base; const int* seg_static; int* seg_cnt; int2* seg_list;
  const float* lig_pos; const float* rec_pos; const float* cutoff;
  int64_t slot_ll, slot_lr;
  float r2_lig, r2_cross;
  int dynamic;
  unsigned long long* edge_total;
};

// One warp per dynamic segment; ordered (ascending index) compaction so the lists -- and therefore every
// floating-point sum over them -- are deterministic.
//   task <  NL       : group 0, ligand atom j: radius-graph edges j -> i (torch_cluster.radius_graph keeps, for each
//                      centre i, the first 33 in-radius candidates in index order incl. i itself, then drops i)
//   task <  2 NL     : group 1, ligand atom a: cross edges a -> r
//   task <  2 NL + NR: group 3, residue r: reversed cross edges r -> a (same predicate, same slot)
__global__ void k_build_lists(ListArgs p) {
  int task = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (task >= 2 * p.NL + p.NR) return;
  if (task < p.NL) {
    int j = task, g = p.lig_graph[j], l0 = p.lig_ptr[g], l1 = p.lig_ptr[g + 1], nl = l1 - l0;
    int seg = 2 * j;
    int base = p.seg_base[seg], cnt = p.seg_static[seg];
    float jx = p.lig_pos[j * 3], jy = p.lig_pos[j * 3 + 1], jz = p.lig_pos[j * 3 + 2];
    for (int ib = l0; ib < l1; ib += 32) {
      int i = ib + lane;
      bool keep = false;
      if (i < l1 && i != j) {
        float ix = p.lig_pos[i * 3], iy = p.lig_pos[i * 3 + 1], iz = p.lig_pos[i * 3 + 2];
        if (dist2_unfused(ix, iy, iz, jx, jy, jz) < p.r2_lig) {
          int rank = 0;  // candidates k < j of centre i that are in radius (incl. k == i)
          for (int k = l0; k < j; ++k)
            rank += dist2_unfused(ix, iy, iz, p.lig_pos[k * 3], p.lig_pos[k * 3 + 1], p.lig_pos[k * 3 + 2]) < p.r2_lig;
          keep = rank < 33;
        }
      }
      unsigned m = __ballot_sync(0xffffffffu, keep);
      if (keep) {
        int pos = cnt + __popc(m & ((1u << lane) - 1));
        p.seg_list[base + pos] = make_int2((int)(p.slot_ll + p.ll_off[g] + (int64_t)(j - l0) * nl + (i - l0)), i);
      }
      cnt += __popc(m);
    }
    if (lane == 0) { p.seg_cnt[seg] = cnt; atomicAdd(p.edge_total, (unsigned long long)cnt); }
    return;
  }
  bool from_lig = task < 2 * p.NL;
  int node = from_lig ? task - p.NL : task - 2 * p.NL;   // ligand atom a or residue r (global index in its type)
  int g = from_lig ? p.lig_graph[node] : p.rec_graph[node];
  int l0 = p.lig_ptr[g], l1 = p.lig_ptr[g + 1], r0 = p.rec_ptr[g], r1 = p.rec_ptr[g + 1], nr = r1 - r0;
  int seg = from_lig ? 2 * node + 1 : 2 * (p.NL + node) + 1;
  int base = p.seg_base[seg], cnt = 0;
  float c = p.dynamic ? p.cutoff[g] : 1.f;
  float r2 = p.dynamic ? 1.f : p.r2_cross;
  const float* mypos = from_lig ? p.lig_pos + node * 3 : p.rec_pos + node * 3;
  float mx = mypos[0], my = mypos[1], mz = mypos[2];
  if (p.dynamic) { mx = __fdiv_rn(mx, c); my = __fdiv_rn(my, c); mz = __fdiv_rn(mz, c); }
  int o0 = from_lig ? r0 : l0, o1 = from_lig ? r1 : l1;
  const float* opos = from_lig ? p.rec_pos : p.lig_pos;
  for (int ob = o0; ob < o1; ob += 32) {
    int o = ob + lane;
    bool keep = false;
    if (o < o1) {
      float ox = opos[o * 3], oy = opos[o * 3 + 1], oz = opos[o * 3 + 2];
      if (p.dynamic) { ox = __fdiv_rn(ox, c); oy = __fdiv_rn(oy, c); oz = __fdiv_rn(oz, c); }
      keep = dist2_unfused(mx, my, mz, ox, oy, oz) < r2;
    }
    unsigned m = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      int pos = cnt + __popc(m & ((1u << lane) - 1));
      int a = from_lig ? node : o, r = from_lig ? o : node;
      int slot = (int)(p.slot_lr + p.lr_off[g] + (int64_t)(a - l0) * nr + (r - r0));
      p.seg_list[base + pos] = make_int2(slot, from_lig ? p.NL + r : a);
    }
    cnt += __popc(m);
  }
  if (lane == 0) { p.seg_cnt[seg] = cnt; atomicAdd(p.edge_total, (unsigned long long)cnt); }
}

struct EdgeArgs {
  int NL, NR;
  const int* lig_graph; const int* rec_graph;
  const int* seg_base; const int* seg_static; const int* seg_cnt; const int2* seg_list;
  const float* lig_pos; const float* rec_pos;
  const float* bond_attr;
  const float* tb;
  const float* W1; const float* W2; const float* b2; int wcols;   // edge MLP of this group
  const float* sm;                                                // smearing constants of this group
  const float* rr_pre; int64_t slot_rr;
  const float* lat; int L;                                        // ligand latents (group 0 only)
  const float* uncond; const float* uncond_emb;                   // per-node flag of the source type + embedding
  float* ea_pool; float4* sh_pool;
};

// Edge embedding + harmonics for every listed edge of one group (template: 0 = ligand-ligand, 1 = cross,
// 2 = receptor contacts).  One thread per edge; the weights of the group's MLP are staged transposed in shared
// memory so every FMA reads a warp-uniform address.
template <int GROUP>
#pragma nv_diag_suppress 128    // GROUP == 2 returns early: the generic loop behind it is unreachable in that instantiation only
__global__ void __launch_bounds__(256) k_edge_features(EdgeArgs p) {
  constexpr int NIN = (GROUP == 0) ? (4 + DE) : (GROUP == 1 ? DE : 0);
  __shared__ __align__(16) float sW1[(NIN > 0 ? NIN : 1) * EA];   // [k][o]
  __shared__ __align__(16) float sW2[EA * EA];                     // [k][o]
  __shared__ float sb2[EA];
  __shared__ float sLat[GROUP == 0 ? 4 * 8 * EA / 8 : 1];          // up to 2L = 4 latent columns x 24 (L <= 2)
  for (int i = threadIdx.x; i < NIN * EA; i += blockDim.x) {
    int k = i / EA, o = i % EA;
    int col = (GROUP == 0) ? (k < 4 ? k : 4 + SE + (k - 4)) : SE + k;
    sW1[i] = p.W1[o * p.wcols + col];
  }
  for (int i = threadIdx.x; i < EA * EA; i += blockDim.x) sW2[i] = p.W2[(i % EA) * EA + i / EA];
  if (threadIdx.x < EA) sb2[threadIdx.x] = p.b2[threadIdx.x];
  if (GROUP == 0 && p.L > 0)
    for (int i = threadIdx.x; i < 2 * p.L * EA; i += blockDim.x) sLat[i] = p.W1[(i % EA) * p.wcols + 4 + SE + DE + i / EA];
  __syncthreads();

  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  int nseg = (GROUP == 2) ? p.NR : p.NL;
  if (warp >= nseg) return;
  int node = warp;
  int seg = (GROUP == 0) ? 2 * node : (GROUP == 1 ? 2 * node + 1 : 2 * (p.NL + node));
  int g = (GROUP == 2) ? p.rec_graph[node] : p.lig_graph[node];
  int base = p.seg_base[seg], cnt = p.seg_cnt[seg], nstatic = p.seg_static[seg];
  const float* tbv = p.tb + ((size_t)g * TB_COUNT + (GROUP == 0 ? TB_LIG_EDGE : (GROUP == 1 ? TB_CROSS_EDGE : TB_REC_EDGE))) * NS;
  float un = (p.uncond != nullptr) ? p.uncond[node] : 0.f;
  if constexpr (GROUP == 2) {
    // receptor contacts: the first layer is step-invariant (k_setup_rr_edges) up to the sigma-embedding bias, so only the
    // 24 x 24 second layer is left; the hidden activations are consumed as they are formed (no pre[] array: with it this
    // instantiation spilled 2 KB per thread to local memory)
    for (int e = lane; e < cnt; e += 32) {
      const int2 ent = p.seg_list[base + e];
      const float4* rp4 = reinterpret_cast<const float4*>(p.rr_pre + (size_t)(ent.x - p.slot_rr) * EA);
      float out[EA];
#pragma unroll
      for (int o = 0; o < EA; ++o) out[o] = sb2[o] + un * p.uncond_emb[o];
#pragma unroll
      for (int q = 0; q < EA / 4; ++q) {
        const float4 v = __ldg(rp4 + q);
        const float r4[4] = {fmaxf(v.x + tbv[4 * q], 0.f), fmaxf(v.y + tbv[4 * q + 1], 0.f), fmaxf(v.z + tbv[4 * q + 2], 0.f),
                             fmaxf(v.w + tbv[4 * q + 3], 0.f)};
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int o = 0; o < EA; ++o) out[o] = fmaf(sW2[(4 * q + c) * EA + o], r4[c], out[o]);
      }
      float4* dst = reinterpret_cast<float4*>(p.ea_pool + (size_t)ent.x * EA);
#pragma unroll
      for (int q = 0; q < EA / 4; ++q) dst[q] = make_float4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]);
    }
    return;
  }
  for (int e = lane; e < cnt; e += 32) {
    int2 ent = p.seg_list[base + e];
    float pre[EA];
#pragma unroll
    for (int o = 0; o < EA; ++o) pre[o] = tbv[o];
    if (GROUP == 2) {
      const float* rp = p.rr_pre + (size_t)(ent.x - p.slot_rr) * EA;
#pragma unroll
      for (int o = 0; o < EA; ++o) pre[o] += rp[o];
    } else {
      const float* sp = p.lig_pos + node * 3;
      const float* dp = (GROUP == 0) ? p.lig_pos + ent.y * 3 : p.rec_pos + (ent.y - p.NL) * 3;
      float nrm;
      float4 sh = sh_l01(dp[0] - sp[0], dp[1] - sp[1], dp[2] - sp[2], &nrm);
      p.sh_pool[ent.x] = sh;
      if (GROUP == 0) {
        if (e < nstatic) {   // a covalent bond: one-hot bond type (radius edges carry zeros, score_model.py:317-320)
          const float* ba = p.bond_attr + (size_t)ent.x * 4;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float v = ba[k];
#pragma unroll
            for (int o = 0; o < EA; ++o) pre[o] += sW1[k * EA + o] * v;
          }
        }
        for (int l = 0; l < p.L; ++l) {
          float a = p.lat[node * p.L + l], b = p.lat[ent.y * p.L + l];
#pragma unroll
          for (int o = 0; o < EA; ++o) pre[o] += sLat[l * EA + o] * a + sLat[(p.L + l) * EA + o] * b;
        }
      }
      constexpr int K0 = (GROUP == 0) ? 4 : 0;
      for (int k = 0; k < DE; ++k) {
        float gk = smear1(p.sm, nrm, k);
#pragma unroll
        for (int o = 0; o < EA; ++o) pre[o] += sW1[(K0 + k) * EA + o] * gk;
      }
    }
    float out[EA];
#pragma unroll
    for (int o = 0; o < EA; ++o) out[o] = sb2[o] + un * p.uncond_emb[o];
#pragma unroll
    for (int k = 0; k < EA; ++k) {
      float r = fmaxf(pre[k], 0.f);
#pragma unroll
      for (int o = 0; o < EA; ++o) out[o] += sW2[k * EA + o] * r;
    }
    float4* dst = reinterpret_cast<float4*>(p.ea_pool + (size_t)ent.x * EA);
#pragma unroll
    for (int q = 0; q < EA / 4; ++q) dst[q] = make_float4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]);
  }
}

// ------------------------------------------------------------------------------------------------ launchers
void launch_setup(DdkCtx* c, const DdkBatch* b, const int32_t* lig_x, const float* rec_x, cudaStream_t st) {
  int L = c->cfg.latent_dim;
  if (c->EB + c->ER > 0) {
    LaunchScope ls(c, PC_SETUP, st);
    k_fill_static_lists<<<(c->EB + c->ER + 255) / 256, 256, 0, st>>>(c->EB, c->ER, c->NL, (int)c->slot_rr, ptr<int>(c->b_static_pos),
                                                                   ptr<int>(c->b_bond_dst), ptr<int>(c->b_rr_dst),
                                                                   ptr<int2>(c->b_seg_list));
  }
  {
    dim3 blk(NS, 8);
    LaunchScope ls(c, PC_SETUP, st);
    k_setup_lig_nodes<<<(c->NL + 7) / 8, blk, 0, st>>>(c->NL, lig_x, W(c, DDK_W_LIG_EMB_TABLES), W(c, DDK_W_LIG_NODE_W),
                                                     W(c, DDK_W_LIG_NODE_B), NS + SE + L, L, c->lig_latent,
                                                     ptr<float>(c->b_lig_static));
  }
  {
    int threads = 256, warps_per_block = threads / 32;
    LaunchScope ls(c, PC_SETUP, st);
    k_setup_rec_nodes<<<(c->NR + warps_per_block - 1) / warps_per_block, threads, 0, st>>>(
        c->NR, rec_x, W(c, DDK_W_REC_EMB_TABLE), W(c, DDK_W_REC_NODE_W), W(c, DDK_W_REC_NODE_B), NS + ESM + SE + L, L,
        c->rec_latent, ptr<float>(c->b_rec_static));
  }
  if (c->ER > 0) {
    LaunchScope ls(c, PC_SETUP, st);
    k_setup_rr_edges<<<(c->ER + 127) / 128, 128, 0, st>>>(c->ER, ptr<int>(c->b_rr_src), ptr<int>(c->b_rr_dst), c->rec_pos,
                                                        W(c, DDK_W_REC_EDGE_W1), SE + DE + 2 * L, L, c->rec_latent,
                                                        W(c, DDK_W_SMEAR) + 33 * 1, ptr<float>(c->b_rr_pre),
                                                        ptr<float4>(c->b_sh_pool) + c->slot_rr);
  }
}

void launch_step_consts(DdkCtx* c, const float* sigma_emb, cudaStream_t st) {
  int L = c->cfg.latent_dim;
  TbArgs a;
  auto set = [&](int k, int wid, int bid, int cols, int col0) {
    a.Wsrc[k] = W(c, wid);
    a.bias[k] = bid >= 0 ? W(c, bid) : nullptr;
    a.cols[k] = cols;
    a.col0[k] = col0;
  };
  set(TB_LIG_NODE, DDK_W_LIG_NODE_W, -1, NS + SE + L, NS);
  set(TB_REC_NODE, DDK_W_REC_NODE_W, -1, NS + ESM + SE + L, NS + ESM);
  set(TB_LIG_EDGE, DDK_W_LIG_EDGE_W1, DDK_W_LIG_EDGE_B1, 4 + SE + DE + 2 * L, 4);
  set(TB_REC_EDGE, DDK_W_REC_EDGE_W1, DDK_W_REC_EDGE_B1, SE + DE + 2 * L, 0);
  set(TB_CROSS_EDGE, DDK_W_CROSS_EDGE_W1, DDK_W_CROSS_EDGE_B1, SE + DE + 2 * L, 0);
  set(TB_CENTER, DDK_W_CENTER_EDGE_W1, DDK_W_CENTER_EDGE_B1, DE + SE, DE);
  set(TB_TR_FINAL, DDK_W_TR_FINAL_W1, DDK_W_TR_FINAL_B1, 1 + SE, 1);
  set(TB_ROT_FINAL, DDK_W_ROT_FINAL_W1, DDK_W_ROT_FINAL_B1, 1 + SE, 1);
  LaunchScope ls(c, PC_GRAPH, st);
  k_step_consts<<<c->B, TB_COUNT * NS, 0, st>>>(c->B, sigma_emb, a, ptr<float>(c->b_tb));
}

void launch_build_lists(DdkCtx* c, const float* lig_pos, const float* cutoff, cudaStream_t st) {
  ListArgs p;
  p.NL = c->NL; p.NR = c->NR;
  p.lig_ptr = ptr<int>(c->b_lig_ptr); p.rec_ptr = ptr<int>(c->b_rec_ptr);
  p.lig_graph = ptr<int>(c->b_lig_graph); p.rec_graph = ptr<int>(c->b_rec_graph);
  p.ll_off = ptr<int64_t>(c->b_ll_off); p.lr_off = ptr<int64_t>(c->b_lr_off);
  p.seg_base = ptr<int>(c->b_seg_base); p.seg_static = ptr<int>(c->b_seg_static);
  p.seg_cnt = ptr<int>(c->b_seg_cnt); p.seg_list = ptr<int2>(c->b_seg_list);
  p.lig_pos = lig_pos; p.rec_pos = c->rec_pos; p.cutoff = cutoff;
  p.slot_ll = c->slot_ll; p.slot_lr = c->slot_lr;
  p.r2_lig = c->r2_lig; p.r2_cross = c->r2_cross; p.dynamic = c->cfg.dynamic_max_cross;
  p.edge_total = ptr<unsigned long long>(c->b_edge_total);
  int tasks = 2 * c->NL + c->NR, threads = 256;
  LaunchScope ls(c, PC_GRAPH, st);
  k_build_lists<<<(tasks * 32 + threads - 1) / threads, threads, 0, st>>>(p);
}

void launch_edge_features(DdkCtx* c, const float* lig_pos, cudaStream_t st) {
  int L = c->cfg.latent_dim;
  EdgeArgs p;
  p.NL = c->NL; p.NR = c->NR;
  p.lig_graph = ptr<int>(c->b_lig_graph); p.rec_graph = ptr<int>(c->b_rec_graph);
  p.seg_base = ptr<int>(c->b_seg_base); p.seg_static = ptr<int>(c->b_seg_static);
  p.seg_cnt = ptr<int>(c->b_seg_cnt); p.seg_list = ptr<int2>(c->b_seg_list);
  p.lig_pos = lig_pos; p.rec_pos = c->rec_pos; p.bond_attr = c->bond_attr;
  p.tb = ptr<float>(c->b_tb);
  p.rr_pre = ptr<float>(c->b_rr_pre); p.slot_rr = c->slot_rr;
  p.lat = c->lig_latent; p.L = L;
  p.ea_pool = ptr<float>(c->b_ea_pool); p.sh_pool = ptr<float4>(c->b_sh_pool);
  const float* unc = W(c, DDK_W_UNCOND);
  int threads = 256, wpb = threads / 32;
  // ligand-ligand
  p.W1 = W(c, DDK_W_LIG_EDGE_W1); p.W2 = W(c, DDK_W_LIG_EDGE_W2); p.b2 = W(c, DDK_W_LIG_EDGE_B2);
  p.wcols = 4 + SE + DE + 2 * L; p.sm = W(c, DDK_W_SMEAR) + 33 * 0;
  p.uncond = c->cfg.has_unconditional ? c->lig_uncond : nullptr; p.uncond_emb = unc + 2 * NS;
  { LaunchScope ls(c, PC_GRAPH, st); k_edge_features<0><<<(c->NL + wpb - 1) / wpb, threads, 0, st>>>(p); }
  // cross
  p.W1 = W(c, DDK_W_CROSS_EDGE_W1); p.W2 = W(c, DDK_W_CROSS_EDGE_W2); p.b2 = W(c, DDK_W_CROSS_EDGE_B2);
  p.wcols = SE + DE + 2 * L; p.sm = W(c, DDK_W_SMEAR) + 33 * 2;
  p.uncond_emb = unc + 4 * NS;
  { LaunchScope ls(c, PC_GRAPH, st); k_edge_features<1><<<(c->NL + wpb - 1) / wpb, threads, 0, st>>>(p); }
  // receptor contacts
  p.W1 = W(c, DDK_W_REC_EDGE_W1); p.W2 = W(c, DDK_W_REC_EDGE_W2); p.b2 = W(c, DDK_W_REC_EDGE_B2);
  p.wcols = SE + DE + 2 * L; p.sm = W(c, DDK_W_SMEAR) + 33 * 1;
  p.uncond = c->cfg.has_unconditional ? c->rec_uncond : nullptr; p.uncond_emb = unc + 3 * NS;
  { LaunchScope ls(c, PC_GRAPH, st); k_edge_features<2><<<(c->NR + wpb - 1) / wpb, threads, 0, st>>>(p); }
}

}  // namespace ddk
