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
constexpr int HEAD_ATOMS = 64;     // atoms per pass (two threads per atom)

// One CTA per graph; atoms in passes of 64, two threads per atom: each computes half of the outputs of the three small
// dense layers (hidden activations exchanged through shared memory) and half of the 72 (path, u) blocks of the tensor
// product, whose second-layer rows are formed on the fly from the atom's hidden vector in shared memory.
__global__ void __launch_bounds__(HEAD_THREADS) k_head_trrot(TrRotArgs p) {
  extern __shared__ __align__(16) float hsm[];
  float (*sWc1)[EA] = reinterpret_cast<float (*)[EA]>(hsm);                       // [32 k][24 o]
  float (*sWc2)[EA] = reinterpret_cast<float (*)[EA]>(hsm + DE * EA);             // [24 k][24 o]
  float (*sWf1)[48] = reinterpret_cast<float (*)[48]>(hsm + DE * EA + EA * EA);   // [48 k][48 o]
  float (*sWf2)[144] = reinterpret_cast<float (*)[144]>(hsm + DE * EA + EA * EA + 48 * 48);   // [48 k][144 r]
  float* sbf2 = hsm + DE * EA + EA * EA + 48 * 48 + 48 * 144;                     // [144]
  float (*sPre)[EA + 1] = reinterpret_cast<float (*)[EA + 1]>(sbf2 + 144);        // [64][25]
  float (*sFeat)[49] = reinterpret_cast<float (*)[49]>(sbf2 + 144 + HEAD_ATOMS * (EA + 1));
  float (*sHid)[49] = reinterpret_cast<float (*)[49]>(sbf2 + 144 + HEAD_ATOMS * (EA + 1) + HEAD_ATOMS * 49);
  __shared__ float sred[HEAD_THREADS / 32][16];
  __shared__ float scen[3];
  const int g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int l0 = p.lig_ptr[g], l1 = p.lig_ptr[g + 1], nl = l1 - l0;
  for (int i = tid; i < DE * EA; i += HEAD_THREADS) sWc1[i / EA][i % EA] = p.Wc1[(i % EA) * (DE + SE) + i / EA];
  for (int i = tid; i < EA * EA; i += HEAD_THREADS) { const int o = i / EA, k = i % EA; sWc2[k][o] = p.Wc2[i]; }
  for (int i = tid; i < 48 * 48; i += HEAD_THREADS) { const int o = i / 48, k = i % 48; sWf1[k][o] = p.Wf1[i]; }
  for (int i = tid; i < 48 * 144; i += HEAD_THREADS) { const int r = i / 48, k = i % 48; sWf2[k][r] = p.Wf2[i]; }
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
  const float pw = sqrtf(3.f / (float)(NS + 2 * NV));
  const float k1 = pw * 0.5773502691896258f, k2 = pw * 0.4082482904638631f;   // 1/sqrt3, 1/sqrt6

  float o12[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) o12[i] = 0.f;
  const int al = tid >> 1, half = tid & 1;
  for (int n0 = l0; n0 < l1; n0 += HEAD_ATOMS) {
    const int n = n0 + al;
    const bool on = n < l1;
    const float* xn = p.x + (size_t)(on ? n : l0) * D;
    float nrm = 0.f;
    float4 sh = make_float4(1.f, 0.f, 0.f, 0.f);
    if (on) sh = sh_l01(p.lig_pos[n * 3] - scen[0], p.lig_pos[n * 3 + 1] - scen[1], p.lig_pos[n * 3 + 2] - scen[2], &nrm);
    const float s[3] = {sh.y, sh.z, sh.w};
    __syncthreads();                          // previous pass consumed
    if (on) {                                 // center_edge_embedding, first layer: 12 outputs per thread
      float pre[12];
#pragma unroll
      for (int i = 0; i < 12; ++i) pre[i] = tbc[12 * half + i];
      for (int k = 0; k < DE; ++k) {
        const float gk = smear1(p.sm, nrm, k);
#pragma unroll
        for (int i = 0; i < 12; ++i) pre[i] += sWc1[k][12 * half + i] * gk;
      }
#pragma unroll
      for (int i = 0; i < 12; ++i) sPre[al][12 * half + i] = fmaxf(pre[i], 0.f);
    }
    __syncthreads();
    if (on) {                                 // second layer + concatenation with the atom's scalars
      float ft[12];
#pragma unroll
      for (int i = 0; i < 12; ++i) ft[i] = p.bc2[12 * half + i];
#pragma unroll
      for (int k = 0; k < EA; ++k) {
        const float r = sPre[al][k];
#pragma unroll
        for (int i = 0; i < 12; ++i) ft[i] += sWc2[k][12 * half + i] * r;
      }
#pragma unroll
      for (int i = 0; i < 12; ++i) { sFeat[al][12 * half + i] = ft[i]; sFeat[al][EA + 12 * half + i] = xn[12 * half + i]; }
    }
    __syncthreads();
    if (on) {                                 // final_conv.fc first layer: 24 outputs per thread
      float h[24];
#pragma unroll
      for (int i = 0; i < 24; ++i) h[i] = p.bf1[24 * half + i];
      for (int k = 0; k < 48; ++k) {
        const float v = sFeat[al][k];
#pragma unroll
        for (int i = 0; i < 24; ++i) h[i] += sWf1[k][24 * half + i] * v;
      }
#pragma unroll
      for (int i = 0; i < 24; ++i) sHid[al][24 * half + i] = fmaxf(h[i], 0.f);
    }
    __syncthreads();
    if (on) {
      // second layer rows consumed on the fly by the tensor product; blocks (path, u) of two rows (wv = 0, 1) in the order of
      // the e3nn instruction list: [0e x 1o -> 1o: 24 | 1o x 0e -> 1o: 6 | 1o x 1o -> 1e: 6 | 1e x 0e -> 1e: 6 | 1e x 1o -> 1o: 6 |
      // 0o x 1o -> 1e: 24]; this thread takes the blocks of its parity
      const float* hv = &sHid[al][0];
      float e1o[2][3] = {{0, 0, 0}, {0, 0, 0}}, e1e[2][3] = {{0, 0, 0}, {0, 0, 0}};
      for (int blk = half; blk < 72; blk += 2) {
        float a0 = sbf2[2 * blk], a1 = sbf2[2 * blk + 1];
#pragma unroll 8
        for (int k = 0; k < 48; ++k) { const float hk = hv[k]; a0 += sWf2[k][2 * blk] * hk; a1 += sWf2[k][2 * blk + 1] * hk; }
        float b3[3];
        bool to1o;
        if (blk < 24) { const float c = xn[blk] * k1; b3[0] = c * s[0]; b3[1] = c * s[1]; b3[2] = c * s[2]; to1o = true; }
        else if (blk < 30) { const float* v = xn + 24 + 3 * (blk - 24); const float c = sh.x * k1; b3[0] = c * v[0]; b3[1] = c * v[1]; b3[2] = c * v[2]; to1o = true; }
        else if (blk < 36) { const float* v = xn + 24 + 3 * (blk - 30);
          b3[0] = (v[1] * s[2] - v[2] * s[1]) * k2; b3[1] = (v[2] * s[0] - v[0] * s[2]) * k2; b3[2] = (v[0] * s[1] - v[1] * s[0]) * k2; to1o = false; }
        else if (blk < 42) { const float* v = xn + 42 + 3 * (blk - 36); const float c = sh.x * k1; b3[0] = c * v[0]; b3[1] = c * v[1]; b3[2] = c * v[2]; to1o = false; }
        else if (blk < 48) { const float* v = xn + 42 + 3 * (blk - 42);
          b3[0] = (v[1] * s[2] - v[2] * s[1]) * k2; b3[1] = (v[2] * s[0] - v[0] * s[2]) * k2; b3[2] = (v[0] * s[1] - v[1] * s[0]) * k2; to1o = true; }
        else { const float c = xn[60 + (blk - 48)] * k1; b3[0] = c * s[0]; b3[1] = c * s[1]; b3[2] = c * s[2]; to1o = false; }
        if (to1o) {
#pragma unroll
          for (int c = 0; c < 3; ++c) { e1o[0][c] += a0 * b3[c]; e1o[1][c] += a1 * b3[c]; }
        } else {
#pragma unroll
          for (int c = 0; c < 3; ++c) { e1e[0][c] += a0 * b3[c]; e1e[1][c] += a1 * b3[c]; }
        }
      }
#pragma unroll
      for (int wv = 0; wv < 2; ++wv)
#pragma unroll
        for (int c = 0; c < 3; ++c) { o12[wv * 3 + c] += e1o[wv][c]; o12[6 + wv * 3 + c] += e1e[wv][c]; }
    }
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
  const int* lig_ptr; const int* rot_graph; const int* rot_u; const int* rot_v;
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

constexpr int TOR_THREADS = 128;
constexpr int TOR_E = 32;          // torch_cluster.radius max_num_neighbors of the bond graph (score_model.py:430)
constexpr int TOR_FS = HID + 1;    // padded row of the per-edge feature / hidden tiles

// One CTA per rotatable bond, every phase spread over the 128 threads:
//   1. warp 0 lists the <= 32 atoms within 5 A of the bond midpoint (index order, as torch_cluster.radius does)
//   2. edge embedding + concatenated features feat[e][72], tensor-product coefficients d[e][blk,u] (4 threads per edge)
//   3. first layer of tor_bond_conv.fc: H = relu(feat W1^T + b1)                       (thread = edge x 18 columns)
//   4. re-association (same algebra as the conv layers): A[q][k] = sum_e d[e][q] H[e][k], Dsum[q] = sum_e d[e][q], so the
//      72 -> 288 second layer runs once per BOND instead of once per edge:
//      out[w] = sum_u ( W2[r(blk,u,w)] . A[blk,u] + b2[r] Dsum[blk,u] )
//   5. mean over the edges, batch-norm affine, tor_final_layer.
__global__ void __launch_bounds__(TOR_THREADS) k_head_tor(TorArgs p) {
  extern __shared__ __align__(16) float tsm[];
  float* sW1 = tsm;                         // [72 k][72 o]
  float* sWe1 = sW1 + HID * HID;            // [32 k][24 o]
  float* sWe2 = sWe1 + DE * EA;             // [24 k][24 o]
  float* sFeat = sWe2 + EA * EA;            // [32 e][73]
  float* sH = sFeat + TOR_E * TOR_FS;       // [32 e][73]
  float* sPre = sH + TOR_E * TOR_FS;        // [32 e][25]
  float* sD = sPre + TOR_E * (EA + 1);      // [32 e][12]
  float* sA = sD + TOR_E * 12;              // [12 q][72 k]
  float* sDsum = sA + 12 * HID;             // [12]
  float* sRow = sDsum + 12;                 // [288]
  float* sOut = sRow + 288;                 // [48]
  __shared__ int sIdx[TOR_E];
  __shared__ int sCnt[2];
  const int b = blockIdx.x, g = p.rot_graph[b], tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int l0 = p.lig_ptr[g], l1 = p.lig_ptr[g + 1];
  for (int i = tid; i < HID * HID; i += TOR_THREADS) { const int o = i / HID, k = i % HID; sW1[k * HID + o] = p.W1[i]; }
  for (int i = tid; i < DE * EA; i += TOR_THREADS) { const int o = i / DE, k = i % DE; sWe1[k * EA + o] = p.We1[i]; }
  for (int i = tid; i < EA * EA; i += TOR_THREADS) { const int o = i / EA, k = i % EA; sWe2[k * EA + o] = p.We2[i]; }
  const float kpw = sqrtf(1.f / (float)NV) * 0.5773502691896258f;   // path weight sqrt(1/nv) * C(1,1,0) = 1/sqrt3
  {
    const int u = p.rot_u[b], v = p.rot_v[b];
    const float ux = p.lig_pos[u * 3], uy = p.lig_pos[u * 3 + 1], uz = p.lig_pos[u * 3 + 2];
    const float vx = p.lig_pos[v * 3], vy = p.lig_pos[v * 3 + 1], vz = p.lig_pos[v * 3 + 2];
    const float mx = (ux + vx) / 2.f, my = (uy + vy) / 2.f, mz = (uz + vz) / 2.f;   // bond midpoint (:428)
    __syncthreads();                           // weights staged
    // ---- 1. neighbour list
    if (w == 0) {
      int cnt = 0;
      for (int ab = l0; ab < l1; ab += 32) {
        const int a = ab + lane;
        bool hit = false;
        if (a < l1) hit = dist2_unfused(mx, my, mz, p.lig_pos[a * 3], p.lig_pos[a * 3 + 1], p.lig_pos[a * 3 + 2]) < p.r2_lig;
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        const int pos = cnt + __popc(m & ((1u << lane) - 1));
        if (hit && pos < TOR_E) sIdx[pos] = a;
        cnt += __popc(m);
      }
      if (lane == 0) sCnt[0] = min(cnt, TOR_E);
    }
    __syncthreads();
    const int ne = sCnt[0];
    // ---- 2. per edge: harmonics / filter / coefficients (sub-thread 0), first embedding layer (6 outputs per sub-thread)
    const int e = tid >> 2, sub = tid & 3;
    float nrm = 0.f;
    int a = 0;
    if (e < ne) {
      a = sIdx[e];
      const float ax = p.lig_pos[a * 3], ay = p.lig_pos[a * 3 + 1], az = p.lig_pos[a * 3 + 2];
      const float4 sh = sh_l01(ax - mx, ay - my, az - mz, &nrm);
      if (sub == 0) {
        // Y2 of the bond direction (e3nn 'component' normalisation)
        float bx = vx - ux, by = vy - uy, bz = vz - uz;
        const float bn = fmaxf(sqrtf(bx * bx + by * by + bz * bz), 1e-12f);
        bx /= bn; by /= bn; bz /= bn;
        const float s3 = 1.7320508075688772f, s5 = 2.23606797749979f;
        const float y2[5] = {s5 * s3 * bx * bz, s5 * s3 * bx * by, s5 * (by * by - 0.5f * (bx * bx + bz * bz)), s5 * s3 * by * bz,
                             s5 * (s3 / 2.f) * (bz * bz - bx * bx)};
        const float s[3] = {sh.y, sh.z, sh.w};
        // filter = 1o block of FullTensorProduct(sh, Y2): sqrt3 * C121[i][j][k] s_i Y2_j   (score_model.py:296)
        const float ca = 0.31622776601683794f, cb = 0.18257418583505536f;   // 1/sqrt10, 1/sqrt30
        float f[3];
        f[0] = s3 * (ca * (s[1] * y2[1] + s[2] * y2[0]) - ca * s[0] * y2[4] - cb * s[0] * y2[2]);
        f[1] = s3 * (ca * (s[0] * y2[1] + s[2] * y2[3]) + 2.f * cb * s[1] * y2[2]);
        f[2] = s3 * (ca * (s[0] * y2[0] + s[1] * y2[3] + s[2] * y2[4]) - cb * s[2] * y2[2]);
        const float* xa = p.x + (size_t)a * D;
#pragma unroll
        for (int q = 0; q < 12; ++q) {         // q = blk * 6 + u: x1o[u] . f (blk 0), x1e[u] . f (blk 1)
          const float* x3 = xa + 24 + 3 * q;   // x1o at 24..41, x1e at 42..59
          sD[e * 12 + q] = (x3[0] * f[0] + x3[1] * f[1] + x3[2] * f[2]) * kpw;
        }
      }
      float pre[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) pre[i] = p.be1[6 * sub + i];
      for (int k = 0; k < DE; ++k) {
        const float gk = smear1(p.sm, nrm, k);
#pragma unroll
        for (int i = 0; i < 6; ++i) pre[i] += sWe1[k * EA + 6 * sub + i] * gk;
      }
#pragma unroll
      for (int i = 0; i < 6; ++i) sPre[e * (EA + 1) + 6 * sub + i] = fmaxf(pre[i], 0.f);
    }
    __syncthreads();
    if (e < ne) {
      float ft[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) ft[i] = p.be2[6 * sub + i];
#pragma unroll
      for (int k = 0; k < EA; ++k) {
        const float r = sPre[e * (EA + 1) + k];
#pragma unroll
        for (int i = 0; i < 6; ++i) ft[i] += sWe2[k * EA + 6 * sub + i] * r;
      }
      const float* xa = p.x + (size_t)a * D;
      const float* xu = p.x + (size_t)u * D;
      const float* xv = p.x + (size_t)v * D;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const int o = 6 * sub + i;
        sFeat[e * TOR_FS + o] = ft[i];
        sFeat[e * TOR_FS + EA + o] = xa[o];
        sFeat[e * TOR_FS + EA + NS + o] = xu[o] + xv[o];
      }
    }
    __syncthreads();
    // ---- 3. H = relu(feat W1^T + b1): thread = (edge, 18 of the 72 columns)
    if (e < ne) {
      float h[18];
#pragma unroll
      for (int i = 0; i < 18; ++i) h[i] = p.b1[18 * sub + i];
      for (int k = 0; k < HID; ++k) {
        const float fv = sFeat[e * TOR_FS + k];
        const float* wr = sW1 + k * HID + 18 * sub;
#pragma unroll
        for (int i = 0; i < 18; ++i) h[i] += wr[i] * fv;
      }
#pragma unroll
      for (int i = 0; i < 18; ++i) sH[e * TOR_FS + 18 * sub + i] = fmaxf(h[i], 0.f);
    }
    __syncthreads();
    // ---- 4. A[q][k] = sum_e d[e][q] H[e][k], Dsum[q]; then the 288 second-layer rows against A
    for (int idx = tid; idx < 12 * HID; idx += TOR_THREADS) {
      const int q = idx / HID, k = idx % HID;
      float acc = 0.f;
      for (int ee = 0; ee < ne; ++ee) acc += sD[ee * 12 + q] * sH[ee * TOR_FS + k];
      sA[idx] = acc;
    }
    if (tid < 12) {
      float acc = 0.f;
      for (int ee = 0; ee < ne; ++ee) acc += sD[ee * 12 + tid];
      sDsum[tid] = acc;
    }
    __syncthreads();
    for (int r = tid; r < 288; r += TOR_THREADS) {   // r = blk * 144 + u * 24 + w
      const int q = r / NS;
      const float4* wr = reinterpret_cast<const float4*>(p.W2 + (size_t)r * HID);
      const float* ar = sA + q * HID;
      float acc = p.b2[r] * sDsum[q];
#pragma unroll
      for (int k4 = 0; k4 < HID / 4; ++k4) {
        const float4 wv = __ldg(wr + k4);
        acc += wv.x * ar[4 * k4] + wv.y * ar[4 * k4 + 1] + wv.z * ar[4 * k4 + 2] + wv.w * ar[4 * k4 + 3];
      }
      sRow[r] = acc;
    }
    __syncthreads();
    // ---- 5. sum over u in a fixed order, mean over edges, batch norm; output order: 24 x 0o then 24 x 0e (score_model.py:156)
    if (tid < 48) {
      const int blk = tid < NS ? 1 : 0, wv = tid % NS;
      float acc = 0.f;
      for (int uu = 0; uu < NV; ++uu) acc += sRow[blk * 144 + uu * NS + wv];
      sOut[tid] = (acc / (float)max(ne, 1)) * p.bn_scale[tid] + p.bn_shift[tid];
    }
    __syncthreads();
    // tor_final_layer: Linear(48->24, no bias) . tanh . Linear(24->1, no bias)
    if (w == 0) {
      float part = 0.f;
      if (lane < NS) {
        float acc = 0.f;
        for (int k = 0; k < 48; ++k) acc += p.Wt1[lane * 48 + k] * sOut[k];
        part = p.Wt2[lane] * tanhf(acc);
      }
      part = warp_sum(part);
      if (lane == 0) p.tor[b] = part * p.tor_scale[g];
    }
  }
}

size_t tor_smem_bytes() {
  return (size_t)(HID * HID + DE * EA + EA * EA + 2 * TOR_E * TOR_FS + TOR_E * (EA + 1) + TOR_E * 12 + 12 * HID + 12 + 288 + 48) * sizeof(float);
}

size_t trrot_smem_bytes() {
  return (size_t)(DE * EA + EA * EA + 48 * 48 + 48 * 144 + 144 + HEAD_ATOMS * (EA + 1) + 2 * HEAD_ATOMS * 49) * sizeof(float);
}

cudaError_t heads_configure() {
  cudaError_t e = cudaFuncSetAttribute(k_head_trrot, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)trrot_smem_bytes());
  if (e != cudaSuccess) return e;
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
  k_head_trrot<<<c->B, HEAD_THREADS, trrot_smem_bytes(), st>>>(p);
}

void launch_head_tor(DdkCtx* c, const float* lig_pos, const float* x, const DdkStepInputs* in, float* tor, cudaStream_t st) {
  if (c->RB == 0 || c->cfg.no_torsion) return;
  TorArgs p;
  p.lig_ptr = ptr<int>(c->b_lig_ptr); p.rot_graph = ptr<int>(c->b_rot_graph);
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
  k_head_tor<<<c->RB, TOR_THREADS, tor_smem_bytes(), st>>>(p);
}

}  // namespace ddk
