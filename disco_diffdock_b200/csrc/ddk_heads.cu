// Score heads: translation / rotation (final_conv over centroid->atom edges) and torsion (tor_bond_conv over
// bond-midpoint->atom edges).  Reference: /root/reference/models/score_model.py:269-308, 410-438.
// Both tensor products are e3nn FullyConnectedTensorProducts with per-edge weights; normalisation constants follow
// e3nn's 'component' / 'element' rules.  These heads are <1 % of the work of a step, so
// the kernels favour simplicity: one CTA per graph, one thread per edge.
#include "ddk_device.cuh"

namespace ddk {

struct TrRotArgs {
  const int* lig_ptr;
  const float* lig_pos; const float* x;   // x: [N][84], ligand rows first
  const float* tb;
  const float* sm;                         // center smearing
  const float* Wc1; const float* Wc2; const float* bc2;            // center_edge_embedding (W1 [24][64]: smear cols 0:32)
  const float* Wf1; const float* bf1; const float* Wf2; const float* bf2;   // final_conv.fc [48][48], [144][48]
  const float* bn;                         // [4]
  const float* Wt1; const float* Wt2; const float* bt2;            // tr_final_layer  [24][33], [24], [1]
  const float* Wr1; const float* Wr2; const float* br2;            // rot_final_layer
  const float* tr_sigma; const float* rot_scale;
  float* tr; float* rot;
};

constexpr int HEAD_THREADS = 128;

__global__ void __launch_bounds__(HEAD_THREADS) k_head_trrot(TrRotArgs p) {
  __shared__ float sWc1[DE][EA];        // [k][o]
  __shared__ float sWc2[EA][EA];        // [k][o]
  __shared__ float sWf1[48][48];        // [k][o]
  __shared__ float sWf2[48][144];       // [k][r]
  __shared__ float sbf2[144];
  __shared__ float sred[HEAD_THREADS / 32][16];
  __shared__ float scen[3];
  const int g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int l0 = p.lig_ptr[g], l1 = p.lig_ptr[g + 1], nl = l1 - l0;
  for (int i = tid; i < DE * EA; i += HEAD_THREADS) sWc1[i / EA][i % EA] = p.Wc1[(i % EA) * (DE + SE) + i / EA];
  for (int i = tid; i < EA * EA; i += HEAD_THREADS) sWc2[i / EA][i % EA] = p.Wc2[(i % EA) * EA + i / EA];
  for (int i = tid; i < 48 * 48; i += HEAD_THREADS) sWf1[i / 48][i % 48] = p.Wf1[(i % 48) * 48 + i / 48];
  for (int i = tid; i < 48 * 144; i += HEAD_THREADS) sWf2[i / 144][i % 144] = p.Wf2[(i % 144) * 48 + i / 144];
  for (int i = tid; i < 144; i += HEAD_THREADS) sbf2[i] = p.bf2[i];
  // centroid (build_center_conv_graph, score_model.py:414-416)
  float cx = 0.f, cy = 0.f, cz = 0.f;
  for (int i = l0 + tid; i < l1; i += HEAD_THREADS) { cx += p.lig_pos[i * 3]; cy += p.lig_pos[i * 3 + 1]; cz += p.lig_pos[i * 3 + 2]; }
  cx = warp_sum(cx); cy = warp_sum(cy); cz = warp_sum(cz);
  if (lane == 0) { sred[w][0] = cx; sred[w][1] = cy; sred[w][2] = cz; }
  __syncthreads();
  if (tid < 3) {
    float s = 0.f;
    for (int q = 0; q < HEAD_THREADS / 32; ++q) s += sred[q][tid];
    scen[tid] = s / (float)nl;
  }
  __syncthreads();
  const float* tbc = p.tb + ((size_t)g * TB_COUNT + TB_CENTER) * NS;

  float o12[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) o12[i] = 0.f;
  for (int n = l0 + tid; n < l1; n += HEAD_THREADS) {
    const float* xn = p.x + (size_t)n * D;
    float nrm;
    float4 sh = sh_l01(p.lig_pos[n * 3] - scen[0], p.lig_pos[n * 3 + 1] - scen[1], p.lig_pos[n * 3 + 2] - scen[2], &nrm);
    float s[3] = {sh.y, sh.z, sh.w};
    // center_edge_embedding
    float pre[EA];
#pragma unroll
    for (int o = 0; o < EA; ++o) pre[o] = tbc[o];
    for (int k = 0; k < DE; ++k) {
      float gk = smear1(p.sm, nrm, k);
#pragma unroll
      for (int o = 0; o < EA; ++o) pre[o] += sWc1[k][o] * gk;
    }
    float feat[48];
#pragma unroll
    for (int o = 0; o < EA; ++o) feat[o] = p.bc2[o];
#pragma unroll
    for (int k = 0; k < EA; ++k) {
      float r = fmaxf(pre[k], 0.f);
#pragma unroll
      for (int o = 0; o < EA; ++o) feat[o] += sWc2[k][o] * r;
    }
#pragma unroll
    for (int o = 0; o < NS; ++o) feat[EA + o] = xn[o];
    // final_conv.fc first layer
    float h[48];
#pragma unroll
    for (int o = 0; o < 48; ++o) h[o] = p.bf1[o];
#pragma unroll
    for (int k = 0; k < 48; ++k) {
      float v = feat[k];
#pragma unroll
      for (int o = 0; o < 48; ++o) h[o] += sWf1[k][o] * v;
    }
#pragma unroll
    for (int o = 0; o < 48; ++o) h[o] = fmaxf(h[o], 0.f);
    // second layer rows consumed on the fly by the tensor product
    auto wrow = [&](int r) {
      float a = sbf2[r];
#pragma unroll
      for (int k = 0; k < 48; ++k) a += sWf2[k][r] * h[k];
      return a;
    };
    const float pw = sqrtf(3.f / (float)(NS + 2 * NV));
    const float k1 = pw * 0.5773502691896258f, k2 = pw * 0.4082482904638631f;   // 1/sqrt3, 1/sqrt6
    float e1o[2][3] = {{0, 0, 0}, {0, 0, 0}}, e1e[2][3] = {{0, 0, 0}, {0, 0, 0}};
    int r = 0;
    for (int u = 0; u < NS; ++u)            // 0e (x) 1o -> 1o
      for (int wv = 0; wv < 2; ++wv, ++r) { float a = wrow(r) * xn[u] * k1; e1o[wv][0] += a * s[0]; e1o[wv][1] += a * s[1]; e1o[wv][2] += a * s[2]; }
    for (int u = 0; u < NV; ++u)            // 1o (x) 0e -> 1o
      for (int wv = 0; wv < 2; ++wv, ++r) { float a = wrow(r) * sh.x * k1; for (int c = 0; c < 3; ++c) e1o[wv][c] += a * xn[24 + 3 * u + c]; }
    for (int u = 0; u < NV; ++u) {          // 1o (x) 1o -> 1e
      const float* v = xn + 24 + 3 * u;
      float cr[3] = {v[1] * s[2] - v[2] * s[1], v[2] * s[0] - v[0] * s[2], v[0] * s[1] - v[1] * s[0]};
      for (int wv = 0; wv < 2; ++wv, ++r) { float a = wrow(r) * k2; for (int c = 0; c < 3; ++c) e1e[wv][c] += a * cr[c]; }
    }
    for (int u = 0; u < NV; ++u)            // 1e (x) 0e -> 1e
      for (int wv = 0; wv < 2; ++wv, ++r) { float a = wrow(r) * sh.x * k1; for (int c = 0; c < 3; ++c) e1e[wv][c] += a * xn[42 + 3 * u + c]; }
    for (int u = 0; u < NV; ++u) {          // 1e (x) 1o -> 1o
      const float* v = xn + 42 + 3 * u;
      float cr[3] = {v[1] * s[2] - v[2] * s[1], v[2] * s[0] - v[0] * s[2], v[0] * s[1] - v[1] * s[0]};
      for (int wv = 0; wv < 2; ++wv, ++r) { float a = wrow(r) * k2; for (int c = 0; c < 3; ++c) e1o[wv][c] += a * cr[c]; }
    }
    for (int u = 0; u < NS; ++u)            // 0o (x) 1o -> 1e
      for (int wv = 0; wv < 2; ++wv, ++r) { float a = wrow(r) * xn[60 + u] * k1; e1e[wv][0] += a * s[0]; e1e[wv][1] += a * s[1]; e1e[wv][2] += a * s[2]; }
#pragma unroll
    for (int wv = 0; wv < 2; ++wv)
#pragma unroll
      for (int c = 0; c < 3; ++c) { o12[wv * 3 + c] += e1o[wv][c]; o12[6 + wv * 3 + c] += e1e[wv][c]; }
  }
#pragma unroll
  for (int i = 0; i < 12; ++i) o12[i] = warp_sum(o12[i]);
  __syncthreads();
  if (lane == 0)
    for (int i = 0; i < 12; ++i) sred[w][i] = o12[i];
  __syncthreads();
  if (tid == 0) {
    float gp[12];
    for (int i = 0; i < 12; ++i) {
      float sacc = 0.f;
      for (int q = 0; q < HEAD_THREADS / 32; ++q) sacc += sred[q][i];
      gp[i] = sacc / (float)max(nl, 1) * p.bn[i / 3];       // mean over atoms, batch-norm scale (vectors: no shift)
    }
    float trv[3], rov[3];
    for (int c = 0; c < 3; ++c) { trv[c] = gp[c] + gp[6 + c]; rov[c] = gp[3 + c] + gp[9 + c]; }   // score_model.py:274-275
    const float* tbt = p.tb + ((size_t)g * TB_COUNT + TB_TR_FINAL) * NS;
    const float* tbr = p.tb + ((size_t)g * TB_COUNT + TB_ROT_FINAL) * NS;
    float tn = sqrtf(trv[0] * trv[0] + trv[1] * trv[1] + trv[2] * trv[2]);
    float rn = sqrtf(rov[0] * rov[0] + rov[1] * rov[1] + rov[2] * rov[2]);
    float mt = p.bt2[0], mr = p.br2[0];
    for (int o = 0; o < NS; ++o) {
      mt += p.Wt2[o] * fmaxf(p.Wt1[o * (1 + SE)] * tn + tbt[o], 0.f);
      mr += p.Wr2[o] * fmaxf(p.Wr1[o * (1 + SE)] * rn + tbr[o], 0.f);
    }
    for (int c = 0; c < 3; ++c) {
      p.tr[g * 3 + c] = trv[c] / tn * mt / p.tr_sigma[g];
      p.rot[g * 3 + c] = rov[c] / rn * mr * p.rot_scale[g];
    }
  }
}

// ------------------------------------------------------------------------------------------------ torsion head
struct TorArgs {
  const int* lig_ptr; const int* rot_ptr; const int* rot_u; const int* rot_v;
  const float* lig_pos; const float* x;
  const float* sm;                         // ligand smearing (reused for bond edges, score_model.py:433)
  const float* We1; const float* be1; const float* We2; const float* be2;   // final_edge_embedding [24][32],[24][24]
  const float* W1; const float* b1; const float* W2; const float* b2;       // tor_bond_conv.fc [72][72], [288][72]
  const float* bn_scale; const float* bn_shift;                             // [48]
  const float* Wt1; const float* Wt2;                                       // tor_final_layer [24][48], [24]
  const float* tor_scale;
  float r2_lig;
  float* tor;
};

constexpr int TOR_THREADS = 128;   // 4 warps; each warp takes rotatable bonds round-robin, one lane per edge

__global__ void __launch_bounds__(TOR_THREADS) k_head_tor(TorArgs p) {
  extern __shared__ __align__(16) float tsm[];
  float* sW1 = tsm;                    // [72 k][72 o]
  float* sW2 = sW1 + HID * HID;        // [72 k][288 r]
  float* sWe1 = sW2 + HID * 288;       // [32 k][24 o]
  float* sWe2 = sWe1 + DE * EA;        // [24 k][24 o]
  float* sfeat = sWe2 + EA * EA;       // [4 warps][48]
  const int g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int b0 = p.rot_ptr[g], b1 = p.rot_ptr[g + 1];
  if (b0 == b1) return;
  const int l0 = p.lig_ptr[g], l1 = p.lig_ptr[g + 1];
  for (int i = tid; i < HID * HID; i += TOR_THREADS) sW1[i] = p.W1[(i % HID) * HID + i / HID];
  for (int i = tid; i < HID * 288; i += TOR_THREADS) sW2[i] = p.W2[(i % 288) * HID + i / 288];
  for (int i = tid; i < DE * EA; i += TOR_THREADS) sWe1[i] = p.We1[(i % EA) * DE + i / EA];
  for (int i = tid; i < EA * EA; i += TOR_THREADS) sWe2[i] = p.We2[(i % EA) * EA + i / EA];
  __syncthreads();
  const float kpw = sqrtf(1.f / (float)NV) * 0.5773502691896258f;   // path weight sqrt(1/nv) * C(1,1,0) = 1/sqrt3
  for (int b = b0 + w; b < b1; b += TOR_THREADS / 32) {
    const int u = p.rot_u[b], v = p.rot_v[b];
    const float ux = p.lig_pos[u * 3], uy = p.lig_pos[u * 3 + 1], uz = p.lig_pos[u * 3 + 2];
    const float vx = p.lig_pos[v * 3], vy = p.lig_pos[v * 3 + 1], vz = p.lig_pos[v * 3 + 2];
    const float mx = (ux + vx) / 2.f, my = (uy + vy) / 2.f, mz = (uz + vz) / 2.f;   // bond midpoint (:428)
    // Y2 of the bond direction (e3nn 'component' normalisation)
    float bx = vx - ux, by = vy - uy, bz = vz - uz;
    float bn = fmaxf(sqrtf(bx * bx + by * by + bz * bz), 1e-12f);
    bx /= bn; by /= bn; bz /= bn;
    const float s3 = 1.7320508075688772f, s5 = 2.23606797749979f;
    const float y2[5] = {s5 * s3 * bx * bz, s5 * s3 * bx * by, s5 * (by * by - 0.5f * (bx * bx + bz * bz)), s5 * s3 * by * bz,
                         s5 * (s3 / 2.f) * (bz * bz - bx * bx)};
    float acc[48];
#pragma unroll
    for (int i = 0; i < 48; ++i) acc[i] = 0.f;
    int cnt = 0;
    for (int ab = l0; ab < l1 && cnt < 32; ab += 32) {
      int a = ab + lane;
      bool hit = false;
      float ax = 0, ay = 0, az = 0;
      if (a < l1) {
        ax = p.lig_pos[a * 3]; ay = p.lig_pos[a * 3 + 1]; az = p.lig_pos[a * 3 + 2];
        hit = dist2_unfused(mx, my, mz, ax, ay, az) < p.r2_lig;
      }
      unsigned m = __ballot_sync(0xffffffffu, hit);
      bool keep = hit && (cnt + __popc(m & ((1u << lane) - 1)) < 32);   // torch_cluster.radius max_num_neighbors=32
      cnt += __popc(m);
      if (keep) {
        const float* xa = p.x + (size_t)a * D;
        const float* xu = p.x + (size_t)u * D;
        const float* xv = p.x + (size_t)v * D;
        float nrm;
        float4 sh = sh_l01(ax - mx, ay - my, az - mz, &nrm);
        const float s[3] = {sh.y, sh.z, sh.w};
        // filter = 1o block of FullTensorProduct(sh, Y2): sqrt3 * C121[i][j][k] s_i Y2_j   (score_model.py:296)
        const float ca = 0.31622776601683794f, cb = 0.18257418583505536f;   // 1/sqrt10, 1/sqrt30
        float f[3];
        f[0] = s3 * (ca * (s[1] * y2[1] + s[2] * y2[0]) - ca * s[0] * y2[4] - cb * s[0] * y2[2]);
        f[1] = s3 * (ca * (s[0] * y2[1] + s[2] * y2[3]) + 2.f * cb * s[1] * y2[2]);
        f[2] = s3 * (ca * (s[0] * y2[0] + s[1] * y2[3] + s[2] * y2[4]) - cb * s[2] * y2[2]);
        // final_edge_embedding on the smeared distance
        float pre[EA];
#pragma unroll
        for (int o = 0; o < EA; ++o) pre[o] = p.be1[o];
        for (int k = 0; k < DE; ++k) {
          float gk = smear1(p.sm, nrm, k);
#pragma unroll
          for (int o = 0; o < EA; ++o) pre[o] += sWe1[k * EA + o] * gk;
        }
        float feat[HID];
#pragma unroll
        for (int o = 0; o < EA; ++o) feat[o] = p.be2[o];
#pragma unroll
        for (int k = 0; k < EA; ++k) {
          float r = fmaxf(pre[k], 0.f);
#pragma unroll
          for (int o = 0; o < EA; ++o) feat[o] += sWe2[k * EA + o] * r;
        }
#pragma unroll
        for (int o = 0; o < NS; ++o) { feat[EA + o] = xa[o]; feat[EA + NS + o] = xu[o] + xv[o]; }
        float h[HID];
#pragma unroll
        for (int o = 0; o < HID; ++o) h[o] = p.b1[o];
        for (int k = 0; k < HID; ++k) {
          float fv = feat[k];
#pragma unroll
          for (int o = 0; o < HID; ++o) h[o] += sW1[k * HID + o] * fv;
        }
#pragma unroll
        for (int o = 0; o < HID; ++o) h[o] = fmaxf(h[o], 0.f);
        // tensor product: out0e[w] = k * sum_u W[u*24+w] (x1o[u].f);  out0o[w] = k * sum_u W[144+u*24+w] (x1e[u].f)
        for (int blk = 0; blk < 2; ++blk) {
          for (int uu = 0; uu < NV; ++uu) {
            const float* xv3 = xa + (blk == 0 ? 24 : 42) + 3 * uu;
            float d = (xv3[0] * f[0] + xv3[1] * f[1] + xv3[2] * f[2]) * kpw;
            for (int wv = 0; wv < NS; ++wv) {
              int r = blk * (NV * NS) + uu * NS + wv;
              float a = p.b2[r];
              for (int k = 0; k < HID; ++k) a += sW2[k * 288 + r] * h[k];
              acc[(blk == 0 ? NS : 0) + wv] += a * d;    // output order: 24 x 0o then 24 x 0e (score_model.py:156)
            }
          }
        }
      }
    }
    cnt = min(cnt, 32);
#pragma unroll
    for (int i = 0; i < 48; ++i) acc[i] = warp_sum(acc[i]);
    if (lane == 0)
      for (int i = 0; i < 48; ++i) sfeat[w * 48 + i] = (acc[i] / (float)max(cnt, 1)) * p.bn_scale[i] + p.bn_shift[i];
    __syncwarp();
    // tor_final_layer: Linear(48->24, no bias) . tanh . Linear(24->1, no bias)
    float part = 0.f;
    if (lane < NS) {
      float a = 0.f;
      for (int k = 0; k < 48; ++k) a += p.Wt1[lane * 48 + k] * sfeat[w * 48 + k];
      part = p.Wt2[lane] * tanhf(a);
    }
    part = warp_sum(part);
    if (lane == 0) p.tor[b] = part * p.tor_scale[g];
    __syncwarp();
  }
}

size_t tor_smem_bytes() { return (size_t)(HID * HID + HID * 288 + DE * EA + EA * EA + 4 * 48) * sizeof(float); }

cudaError_t heads_configure() {
  return cudaFuncSetAttribute(k_head_tor, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tor_smem_bytes());
}

void launch_head_trrot(DdkCtx* c, const float* lig_pos, const float* x, const DdkStepInputs* in, float* tr, float* rot,
                       cudaStream_t st) {
  TrRotArgs p;
  p.lig_ptr = ptr<int>(c->b_lig_ptr);
  p.lig_pos = lig_pos; p.x = x; p.tb = ptr<float>(c->b_tb);
  p.sm = W(c, DDK_W_SMEAR) + 33 * 3;
  p.Wc1 = W(c, DDK_W_CENTER_EDGE_W1); p.Wc2 = W(c, DDK_W_CENTER_EDGE_W2); p.bc2 = W(c, DDK_W_CENTER_EDGE_B2);
  p.Wf1 = W(c, DDK_W_FINAL_CONV_W1); p.bf1 = W(c, DDK_W_FINAL_CONV_B1);
  p.Wf2 = W(c, DDK_W_FINAL_CONV_W2); p.bf2 = W(c, DDK_W_FINAL_CONV_B2);
  p.bn = W(c, DDK_W_FINAL_CONV_BN);
  p.Wt1 = W(c, DDK_W_TR_FINAL_W1); p.Wt2 = W(c, DDK_W_TR_FINAL_W2); p.bt2 = W(c, DDK_W_TR_FINAL_B2);
  p.Wr1 = W(c, DDK_W_ROT_FINAL_W1); p.Wr2 = W(c, DDK_W_ROT_FINAL_W2); p.br2 = W(c, DDK_W_ROT_FINAL_B2);
  p.tr_sigma = in->tr_sigma; p.rot_scale = in->rot_scale;
  p.tr = tr; p.rot = rot;
  LaunchScope ls(c, PC_HEADS, st);
  k_head_trrot<<<c->B, HEAD_THREADS, 0, st>>>(p);
}

void launch_head_tor(DdkCtx* c, const float* lig_pos, const float* x, const DdkStepInputs* in, float* tor, cudaStream_t st) {
  if (c->RB == 0 || c->cfg.no_torsion) return;
  TorArgs p;
  p.lig_ptr = ptr<int>(c->b_lig_ptr); p.rot_ptr = ptr<int>(c->b_rot_ptr);
  p.rot_u = ptr<int>(c->b_rot_u); p.rot_v = ptr<int>(c->b_rot_v);
  p.lig_pos = lig_pos; p.x = x;
  p.sm = W(c, DDK_W_SMEAR);
  p.We1 = W(c, DDK_W_FINAL_EDGE_W1); p.be1 = W(c, DDK_W_FINAL_EDGE_B1);
  p.We2 = W(c, DDK_W_FINAL_EDGE_W2); p.be2 = W(c, DDK_W_FINAL_EDGE_B2);
  p.W1 = W(c, DDK_W_TOR_CONV_W1); p.b1 = W(c, DDK_W_TOR_CONV_B1);
  p.W2 = W(c, DDK_W_TOR_CONV_W2); p.b2 = W(c, DDK_W_TOR_CONV_B2);
  p.bn_scale = W(c, DDK_W_TOR_CONV_BN_SCALE); p.bn_shift = W(c, DDK_W_TOR_CONV_BN_SHIFT);
  p.Wt1 = W(c, DDK_W_TOR_FINAL_W1); p.Wt2 = W(c, DDK_W_TOR_FINAL_W2);
  p.tor_scale = in->tor_scale;
  p.r2_lig = c->r2_lig;
  p.tor = tor;
  LaunchScope ls(c, PC_HEADS, st);
  k_head_tor<<<c->B, TOR_THREADS, tor_smem_bytes(), st>>>(p);
}

}  // namespace ddk
