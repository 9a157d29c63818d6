// k_conv_fused: one tensor-product convolution layer WITHOUT the outer-product scratch round trip.
//
// Algebra (ddk_conv.cu, /root/reference/models/tensor_layers.py:65-116, 147-168): for a (node s, edge group g) segment
//     out_s = W2p (*) A_s + b2p (*) Bsum_s,   A_s[u][j] = sum_e basis_e[u] * h_e[j],   Bsum_s[u] = sum_e basis_e[u].
// The two-kernel path (ddk_conv2.cu + ddk_contract2.cu) writes A_s (U x 72 fp32, ~80 KB per segment) to HBM and reads it
// back: 4.3 GB per launch at 75 poses (profiles/r01_*).  Here the hidden dimension j is cut into NSL = 9 slices of
// J3 = 8 units.  A CTA owns one (group, slice) pair at a time ("combo"), keeps that slice of the packed second-layer
// weights (W_l x 8 floats, <= 60 KB) resident in shared memory, and for every segment of the group
//   - an accumulate warp builds A_s[:, slice] (U x 8) in REGISTERS: the lane owns up to 9 basis rows u; per edge it
//     evaluates its basis values straight from the staged destination features / harmonics (no basis tile in shared
//     memory) and does 8 FFMA per row against the 8 hidden units h_e[slice] of the edge (first MLP layer, evaluated
//     by the same warp from the staged edge embedding and the per-node projections);
//   - at the end of the segment the U x 8 block (+ the Bsum column in slice 0) goes to a per-warp shared-memory slot;
//   - four contraction warps take the 8 slots of a batch together (so every weight read from shared memory is used for
//     8 segments), contract them against the resident weight slice and write the 84-wide PARTIAL output of
//     (segment, slice) to HBM: 336 B instead of 80 KB.
// k_conv_finalize adds the 2 x 9 partials of a node in a fixed order, applies mean / batch-norm / residual.
// Accumulate and contraction warps are decoupled with named barriers (one batch of slack), every warp gathers its own
// edge stream with cp.async one chunk ahead, and CTAs claim (combo, block of segments) tasks from per-combo counters,
// staying on a combo while it has work so the weight slice is reloaded only when a CTA migrates.
// Results do not depend on the claiming order: each (segment, slice) partial is computed by exactly one warp
// sequence in a fixed order.
#include <cuda_pipeline_primitives.h>

#include <algorithm>
#include <vector>

#include "ddk_conv.cuh"

namespace ddk {

constexpr int EAS = 28;            // padded row of the staged edge embedding (conflict-free LDS.128 across 8 edges)
constexpr int F3_NCOMBO = 4 * NSL;

enum { F3_BAR_FULL = 1, F3_BAR_EMPTY = 2, F3_BAR_CON = 3, F3_BAR_CON2 = 4 };

__device__ __forceinline__ void f3_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void f3_bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

template <int LV>
struct F3Cfg {
  static constexpr int U = AccCfg<LV>::U;
  static constexpr int DINP = AccCfg<LV>::DINP;
  static constexpr int NSLOT = (U + 31) / 32;
  static constexpr int W = LV == 0 ? 720 : (LV == 1 ? 936 : (LV == 2 ? 1152 : 1872));   // second-layer rows (sum F*O)
  // basis rows of type 0 (x[i] * sh[m]) by harmonic component m
  static constexpr int C0 = LV == 0 ? 24 : (LV == 1 ? 42 : (LV == 2 ? 60 : 84));
  static constexpr int CC = LV == 3 ? 48 : 24;
  static constexpr int NU0 = C0 / 32, NUC = CC / 32;
  static constexpr int NT0U = NU0 + 3 * NUC;                 // slots whose 32 rows share one compile-time m
  static constexpr int T0TOT = C0 + 3 * CC;
  static constexpr int NT0 = T0TOT / 32 - NT0U;              // slots of type-0 rows with a per-lane m
  static constexpr int NGEN = NSLOT - NT0U - NT0;            // slots evaluated with the generic 3-term formula
  __host__ __device__ static constexpr int slot_m(int k) { return k < NU0 ? 0 : 1 + (k - NU0) / (NUC > 0 ? NUC : 1); }
};

template <int LV>
struct F3Smem {
  alignas(16) float Wsl[F3Cfg<LV>::W * J3];                  // [class][f][jj][o] of the resident (group, slice)
  alignas(16) float Wb[F3Cfg<LV>::W];                        // packed second-layer bias (used by slice 0 only)
  alignas(16) float As[F3_ACC][F3Cfg<LV>::U * AST];          // one slot per accumulate warp: [u][jj | bsum]
  struct Stage {
    alignas(16) float X[2][KC3][F3Cfg<LV>::DINP];
    alignas(16) float SH[2][KC3][4];
    alignas(16) float EA[KC3][EAS];
    alignas(16) float PD[KC3][J3];
    alignas(16) float H[KC3][J3];
  } st[F3_ACC];
  alignas(16) float W1a[J3][EA];                             // first-layer rows of the slice, edge-embedding columns
  alignas(16) float tile[F3_CON][F3_ACC][D];                 // per contraction warp partial outputs of a batch
  int meta[F3_ACC];                                          // segment id of each slot of the batch in flight (-1: none)
  int task[8];                                               // g, r, idx0, nseg, reload, combo cursor
};

struct F3Args {
  int NL, N;
  int nb_segs;                       // segments per task (multiple of F3_ACC)
  const int4* glist;                 // per-group lists of non-empty segments: (seg, n, base, 0)
  int goff[4];
  const int* gcnt;                   // [4]
  int* counters;                     // [F3_NCOMBO] next block of each combo
  const int2* seg_list;
  const float* x;                    // [N][84] layer input
  const float* projs;                // [NSL][N][4][J3]
  const float* ea_pool; const float4* sh_pool;
  const float* W1[4];                // [72][72]
  const float* W2S[4];               // [NSL][W * J3]
  const float* b2p[4];               // [W]
  const BasisEnt* btab;              // [NSLOT * 32]
  float* part;                       // [2 N][NSL][84]
  ConSplit split;
  ClassInfo cls[4];
  int w8off[4];                      // offset of each class inside a weight slice (floats)
  int ncls;
};

// ---------------------------------------------------------------------------------------------- group work lists
// Ordered compaction of the non-empty segments of each edge group (block g = group g), node order.
__global__ void __launch_bounds__(1024) k_build_group_lists(int NL, int NR, const int* __restrict__ seg_cnt,
                                                            const int* __restrict__ seg_base, int4* __restrict__ glist,
                                                            int* __restrict__ gcnt) {
  __shared__ int wsum[32];
  __shared__ int base_s;
  const int g = blockIdx.x;
  const int nn = g < 2 ? NL : NR;
  const int off = g == 0 ? 0 : (g == 1 ? NL : (g == 2 ? 2 * NL : 2 * NL + NR));
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) base_s = 0;
  __syncthreads();
  for (int i0 = 0; i0 < nn; i0 += 1024) {
    const int i = i0 + tid;
    int seg = 0, n = 0;
    if (i < nn) { seg = g < 2 ? 2 * i + g : 2 * (NL + i) + (g - 2); n = seg_cnt[seg]; }
    const unsigned m = __ballot_sync(0xffffffffu, n > 0);
    if (lane == 0) wsum[w] = __popc(m);
    __syncthreads();
    int o = base_s;
    for (int q = 0; q < w; ++q) o += wsum[q];
    if (n > 0) glist[off + o + __popc(m & ((1u << lane) - 1))] = make_int4(seg, n, seg_base[seg], 0);
    __syncthreads();
    if (tid == 0) { int t = 0; for (int q = 0; q < 32; ++q) t += wsum[q]; base_s += t; }
    __syncthreads();
  }
  if (tid == 0) gcnt[g] = base_s;
}

// ---------------------------------------------------------------------------------------------- accumulate warps
template <int LV>
struct LaneBasis {
  const float* xp[F3Cfg<LV>::NSLOT];                                   // &X[buf][0][i] of the first term
  const float* sp[F3Cfg<LV>::NT0 + F3Cfg<LV>::NGEN > 0 ? F3Cfg<LV>::NT0 + F3Cfg<LV>::NGEN : 1];   // &SH[buf][0][m]
  const float* gx[F3Cfg<LV>::NGEN > 0 ? 2 * F3Cfg<LV>::NGEN : 1];      // generic slots: second / third term
  const float* gs[F3Cfg<LV>::NGEN > 0 ? 2 * F3Cfg<LV>::NGEN : 1];
  float gf[F3Cfg<LV>::NGEN > 0 ? 3 * F3Cfg<LV>::NGEN : 1];
};

struct ChunkD {
  int node, which, base, c0, kc, batch;
  bool first, last, valid, done;
};

template <int LV, bool BIAS>
__device__ __forceinline__ void f3_acc_task(const F3Args& p, F3Smem<LV>& S, LaneBasis<LV>& LB, const int g, const int r,
                                            const int idx0, const int nseg, const int w, const int lane) {
  using Cfg = F3Cfg<LV>;
  constexpr int NSLOT = Cfg::NSLOT, NT0U = Cfg::NT0U, NT0 = Cfg::NT0, NGEN = Cfg::NGEN, DINP = Cfg::DINP, U = Cfg::U;
  constexpr int XQ = DINP / 4;
  constexpr int XBUF = KC3 * DINP, SBUF = KC3 * 4;
  typename F3Smem<LV>::Stage& T = S.st[w];
  const int nb = (nseg + F3_ACC - 1) / F3_ACC;
  const int dslot = (g == 1 || g == 3) ? 3 : 2;
  const int4* wl = p.glist + p.goff[g] + idx0;
  const float* projr = p.projs + (size_t)r * p.N * 4 * J3;

  float acc[NSLOT][J3];
  float bs[NSLOT];
#pragma unroll
  for (int k = 0; k < NSLOT; ++k) {
    bs[k] = 0.f;
#pragma unroll
    for (int j = 0; j < J3; ++j) acc[k][j] = 0.f;
  }

  // ---- chunk generator: batches bi = 0..nb-1, this warp's segment of a batch is idx0 + 8 bi + w (or none)
  int bi = 0, n = 0, sbase = 0, seg = 0, c0 = 0;
  bool in_seg = false;
  int4 pre = (w < nseg) ? wl[w] : make_int4(-1, 0, 0, 0);
  auto next_cd = [&]() {
    ChunkD d;
    d.done = false; d.valid = false; d.first = true; d.last = true;
    d.node = 0; d.which = 0; d.base = 0; d.c0 = 0; d.kc = 0; d.batch = bi;
    if (!in_seg) {
      if (bi >= nb) { d.done = true; return d; }
      const int si = F3_ACC * bi + w;
      if (si >= nseg) { ++bi; return d; }           // no segment for this warp in the batch: empty slot
      seg = pre.x; n = pre.y; sbase = pre.z; c0 = 0; in_seg = true;
      const int sn = si + F3_ACC;
      pre = (sn < nseg) ? wl[sn] : make_int4(-1, 0, 0, 0);
    }
    d.valid = true;
    d.node = seg >> 1; d.which = seg & 1; d.base = sbase; d.c0 = c0;
    d.kc = min(KC3, n - c0);
    d.first = (c0 == 0);
    d.last = (c0 + d.kc >= n);
    c0 += d.kc;
    if (d.last) { in_seg = false; ++bi; }
    return d;
  };
  auto load_ent = [&](const ChunkD& d) {
    int2 e = make_int2(0, 0);
    if (d.valid && lane < d.kc) e = p.seg_list[d.base + d.c0 + lane];
    return e;
  };
  // source-side projection of the segment's node: lanes need hidden units jg and jg + 4 (jg = lane >> 3)
  auto load_ps = [&](const ChunkD& d, float& a, float& b) {
    if (d.valid) {
      const float* q = projr + ((size_t)d.node * 4 + d.which) * J3 + (lane >> 3);
      a = q[0]; b = q[4];
    }
  };
  auto gather = [&](const ChunkD& d, const int2 ent, const int buf) {
    if (d.valid) {
      const int kc = d.kc;
#pragma unroll
      for (int i = 0; i < (KC3 * XQ + 31) / 32; ++i) {
        const int pc = lane + 32 * i;
        const int e = pc / XQ, q = pc % XQ;
        const int dst = __shfl_sync(0xffffffffu, ent.y, e & 7);
        if (pc < KC3 * XQ && e < kc) __pipeline_memcpy_async(&T.X[buf][e][4 * q], p.x + (size_t)dst * D + 4 * q, 16);
      }
      {
        const int e = lane & 7;
        const int slot = __shfl_sync(0xffffffffu, ent.x, e);
        if (lane < 8 && e < kc) __pipeline_memcpy_async(&T.SH[buf][e][0], p.sh_pool + slot, 16);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int pc = lane + 32 * i;                    // 8 edges x 6 pieces = 48
        const int e = pc / 6, q = pc % 6;
        const int slot = __shfl_sync(0xffffffffu, ent.x, e & 7);
        if (pc < KC3 * 6 && e < kc) __pipeline_memcpy_async(&T.EA[e][4 * q], p.ea_pool + (size_t)slot * EA + 4 * q, 16);
      }
      {
        const int e = lane >> 1, q = lane & 1;            // 8 edges x 2 pieces
        const int dst = __shfl_sync(0xffffffffu, ent.y, e & 7);
        if (lane < 16 && e < kc)
          __pipeline_memcpy_async(&T.PD[e][4 * q], projr + ((size_t)dst * 4 + dslot) * J3 + 4 * q, 16);
      }
    }
    __pipeline_commit();
  };

  ChunkD cd0 = next_cd();
  ChunkD cd1 = cd0.done ? cd0 : next_cd();
  int2 ent1;
  float ps0 = 0.f, ps1 = 0.f, pn0 = 0.f, pn1 = 0.f;
  {
    const int2 ent0 = load_ent(cd0);
    load_ps(cd0, ps0, ps1);
    gather(cd0, ent0, 0);
    ent1 = load_ent(cd1);
  }
  int buf = 0;
  while (!cd0.done) {
    __pipeline_wait_prior(0);
    __syncwarp();
    if (cd0.valid) {
      // ---- first radial-MLP layer for the slice: h = relu(W1[:, :24] ea + (W1[:,24:48] x_s + b1) + W1[:,48:72] x_d)
      const int e = lane & 7, jg = lane >> 3;
      float h0 = ps0 + T.PD[e][jg], h1 = ps1 + T.PD[e][jg + 4];
#pragma unroll
      for (int q = 0; q < EA / 4; ++q) {
        const float4 ea = *reinterpret_cast<const float4*>(&T.EA[e][4 * q]);
        const float4 wa = *reinterpret_cast<const float4*>(&S.W1a[jg][4 * q]);
        const float4 wb = *reinterpret_cast<const float4*>(&S.W1a[jg + 4][4 * q]);
        h0 += wa.x * ea.x + wa.y * ea.y + wa.z * ea.z + wa.w * ea.w;
        h1 += wb.x * ea.x + wb.y * ea.y + wb.z * ea.z + wb.w * ea.w;
      }
      T.H[e][jg] = fmaxf(h0, 0.f);
      T.H[e][jg + 4] = fmaxf(h1, 0.f);
    }
    __syncwarp();
    // ---- next chunk's gathers travel while this chunk is accumulated (EA / PD of this chunk are consumed)
    if (!cd1.done) {
      if (cd1.first) load_ps(cd1, pn0, pn1);
      gather(cd1, ent1, buf ^ 1);
    } else {
      __pipeline_commit();
    }
    ChunkD cd2 = cd1.done ? cd1 : next_cd();
    const int2 ent2 = load_ent(cd2);

    if (cd0.valid) {
      const int kc = cd0.kc;
#pragma unroll
      for (int e = 0; e < KC3; ++e) {
        if (e < kc) {
          const float4 ha = *reinterpret_cast<const float4*>(&T.H[e][0]);
          const float4 hb = *reinterpret_cast<const float4*>(&T.H[e][4]);
          const float4 s4 = *reinterpret_cast<const float4*>(&T.SH[0][0][0] + buf * SBUF + e * 4);
#pragma unroll
          for (int k = 0; k < NSLOT; ++k) {
            float b;
            if (k < NT0U) {
              const int m = Cfg::slot_m(k);
              const float sm = m == 0 ? s4.x : (m == 1 ? s4.y : (m == 2 ? s4.z : s4.w));
              b = LB.xp[k][e * DINP] * sm;
            } else if (k < NT0U + NT0) {
              b = LB.xp[k][e * DINP] * LB.sp[k - NT0U][e * 4];
            } else {
              const int q = k - NT0U - NT0;
              b = LB.gf[3 * q] * (LB.xp[k][e * DINP] * LB.sp[k - NT0U][e * 4]);
              b += LB.gf[3 * q + 1] * (LB.gx[2 * q][e * DINP] * LB.gs[2 * q][e * 4]);
              b += LB.gf[3 * q + 2] * (LB.gx[2 * q + 1][e * DINP] * LB.gs[2 * q + 1][e * 4]);
            }
            acc[k][0] += b * ha.x; acc[k][1] += b * ha.y; acc[k][2] += b * ha.z; acc[k][3] += b * ha.w;
            acc[k][4] += b * hb.x; acc[k][5] += b * hb.y; acc[k][6] += b * hb.z; acc[k][7] += b * hb.w;
            if (BIAS) bs[k] += b;
          }
        }
      }
    }
    if (cd0.last) {
      // ---- hand the finished U x 8 block to the contraction warps
      if (cd0.batch > 0) f3_bar_sync(F3_BAR_EMPTY, F3_THREADS);      // they are done with the previous batch
      if (cd0.valid) {
        float* slot = &S.As[w][0];
#pragma unroll
        for (int k = 0; k < NSLOT; ++k) {
          const int u = p.btab[k * 32 + lane].u;
          if (u >= 0) {
#pragma unroll
            for (int j = 0; j < J3; ++j) slot[u * AST + j] = acc[k][j];
            if (BIAS) slot[u * AST + J3] = bs[k];
          }
          bs[k] = 0.f;
#pragma unroll
          for (int j = 0; j < J3; ++j) acc[k][j] = 0.f;
        }
      }
      if (lane == 0) S.meta[w] = cd0.valid ? (2 * cd0.node + cd0.which) : -1;
      __threadfence_block();
      f3_bar_arrive(F3_BAR_FULL, F3_THREADS);
    }
    // ---- rotate
    const int tog = buf ? -XBUF : XBUF;
    const int togs = buf ? -SBUF : SBUF;
#pragma unroll
    for (int k = 0; k < NSLOT; ++k) LB.xp[k] += tog;
#pragma unroll
    for (int k = 0; k < NT0 + NGEN; ++k) LB.sp[k] += togs;
#pragma unroll
    for (int k = 0; k < 2 * NGEN; ++k) { LB.gx[k] += tog; LB.gs[k] += togs; }
    buf ^= 1;
    if (cd1.first) { ps0 = pn0; ps1 = pn1; }
    cd0 = cd1; cd1 = cd2; ent1 = ent2;
  }
  // leave the pointers on buffer 0 for the next task
  if (buf) {
#pragma unroll
    for (int k = 0; k < NSLOT; ++k) LB.xp[k] -= XBUF;
#pragma unroll
    for (int k = 0; k < NT0 + NGEN; ++k) LB.sp[k] -= SBUF;
#pragma unroll
    for (int k = 0; k < 2 * NGEN; ++k) { LB.gx[k] -= XBUF; LB.gs[k] -= SBUF; }
  }
  __pipeline_wait_prior(0);
}

// ---------------------------------------------------------------------------------------------- contraction warps
// scalar output class (O = 24, one component): lane = (k-part kp = lane >> 2, output group og = lane & 3 -> 6 outputs),
// 8 segments per lane; rows (f, jj) of the class are dealt round-robin to the 8 k-parts.
template <bool BIAS, int ASLOT>
__device__ __forceinline__ void f3_con_scalar(const float* __restrict__ Wc, const float* __restrict__ Wbc,
                                              const float* __restrict__ As, int uoff, int f0, int f1, float* tile, int col0,
                                              int lane) {
  constexpr int JC = BIAS ? J3 + 1 : J3;
  const int kp = lane >> 2, og = lane & 3;
  float acc[F3_ACC][6];
#pragma unroll
  for (int s = 0; s < F3_ACC; ++s)
#pragma unroll
    for (int o = 0; o < 6; ++o) acc[s][o] = 0.f;
  const int nrows = (f1 - f0) * JC;
  int f = f0 + kp / JC, jj = kp % JC;
  for (int q = kp; q < nrows; q += 8) {
    const float* wp = (!BIAS || jj < J3) ? Wc + (f * J3 + jj) * 24 + 6 * og : Wbc + f * 24 + 6 * og;
    const float2 w0 = *reinterpret_cast<const float2*>(wp);
    const float2 w1 = *reinterpret_cast<const float2*>(wp + 2);
    const float2 w2 = *reinterpret_cast<const float2*>(wp + 4);
    const float* ap = As + (uoff + f) * AST + jj;
#pragma unroll
    for (int s = 0; s < F3_ACC; ++s) {
      const float a = ap[s * ASLOT];
      acc[s][0] += a * w0.x; acc[s][1] += a * w0.y; acc[s][2] += a * w1.x;
      acc[s][3] += a * w1.y; acc[s][4] += a * w2.x; acc[s][5] += a * w2.y;
    }
    jj += 8;
    if (jj >= JC) { jj -= JC; ++f; }
  }
#pragma unroll
  for (int s = 0; s < F3_ACC; ++s)
#pragma unroll
    for (int o = 0; o < 6; ++o) {
      float v = acc[s][o];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (kp == 0) tile[s * D + col0 + 6 * og + o] += v;
    }
}

// vector output class (O = 6, three components sharing the weights): lane = (kp = lane >> 2, sg = lane & 3 -> segments
// 2 sg, 2 sg + 1), 2 x 3 x 6 accumulators per lane.
template <bool BIAS, int ASLOT>
__device__ __forceinline__ void f3_con_vector(const float* __restrict__ Wc, const float* __restrict__ Wbc,
                                              const float* __restrict__ As, int uoff, int F, int f0, int f1, float* tile,
                                              int col0, int lane) {
  constexpr int JC = BIAS ? J3 + 1 : J3;
  const int kp = lane >> 2, sg = lane & 3;
  float acc[2][3][6];
#pragma unroll
  for (int s = 0; s < 2; ++s)
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int o = 0; o < 6; ++o) acc[s][c][o] = 0.f;
  const int nrows = (f1 - f0) * JC;
  int f = f0 + kp / JC, jj = kp % JC;
  const float* A0 = As + (2 * sg) * ASLOT;
  for (int q = kp; q < nrows; q += 8) {
    const float* wp = (!BIAS || jj < J3) ? Wc + (f * J3 + jj) * 6 : Wbc + f * 6;
    const float2 w0 = *reinterpret_cast<const float2*>(wp);
    const float2 w1 = *reinterpret_cast<const float2*>(wp + 2);
    const float2 w2 = *reinterpret_cast<const float2*>(wp + 4);
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float a = A0[s * ASLOT + (uoff + c * F + f) * AST + jj];
        acc[s][c][0] += a * w0.x; acc[s][c][1] += a * w0.y; acc[s][c][2] += a * w1.x;
        acc[s][c][3] += a * w1.y; acc[s][c][4] += a * w2.x; acc[s][c][5] += a * w2.y;
      }
    jj += 8;
    if (jj >= JC) { jj -= JC; ++f; }
  }
#pragma unroll
  for (int s = 0; s < 2; ++s)
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int o = 0; o < 6; ++o) {
        float v = acc[s][c][o];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        if (kp == 0) tile[(2 * sg + s) * D + col0 + 3 * o + c] += v;
      }
}

template <int LV, bool BIAS>
__device__ __forceinline__ void f3_con_task(const F3Args& p, F3Smem<LV>& S, const int r, const int nseg, const int cw,
                                            const int lane) {
  constexpr int ASLOT = F3Cfg<LV>::U * AST;
  const int nb = (nseg + F3_ACC - 1) / F3_ACC;
  const int ct = cw * 32 + lane;
  float* tile = &S.tile[cw][0][0];
  for (int b = 0; b < nb; ++b) {
    for (int i = lane; i < F3_ACC * D; i += 32) tile[i] = 0.f;
    f3_bar_sync(F3_BAR_FULL, F3_THREADS);            // the 8 slots of batch b are written
    __syncwarp();
#pragma unroll 1
    for (int k = 0; k < p.ncls; ++k) {
      const int f0 = p.split.f0[cw][k], f1 = p.split.f1[cw][k];
      if (f1 <= f0) continue;
      const ClassInfo& ci = p.cls[k];
      const float* Wc = S.Wsl + p.w8off[k];
      const float* Wbc = S.Wb + ci.boff;
      if (ci.ncomp == 1) f3_con_scalar<BIAS, ASLOT>(Wc, Wbc, &S.As[0][0], ci.uoff, f0, f1, tile, ci.col0, lane);
      else f3_con_vector<BIAS, ASLOT>(Wc, Wbc, &S.As[0][0], ci.uoff, ci.F, f0, f1, tile, ci.col0, lane);
    }
    f3_bar_sync(F3_BAR_CON, F3_CON * 32);             // every contraction warp's tile is complete
    for (int i = ct; i < F3_ACC * D; i += F3_CON * 32) {
      const int s = i / D, f = i % D;
      const int sid = S.meta[s];
      if (sid >= 0) {
        const float v = ((S.tile[0][s][f] + S.tile[1][s][f]) + S.tile[2][s][f]) + S.tile[3][s][f];
        p.part[((size_t)sid * NSL + r) * D + f] = v;
      }
    }
    if (b + 1 < nb) f3_bar_arrive(F3_BAR_EMPTY, F3_THREADS);   // slots and meta may be overwritten
    f3_bar_sync(F3_BAR_CON2, F3_CON * 32);            // tiles may be cleared
  }
}

// ---------------------------------------------------------------------------------------------- the kernel
template <int LV>
__global__ void __launch_bounds__(F3_THREADS, 1) k_conv_fused(const __grid_constant__ F3Args p) {
  using Cfg = F3Cfg<LV>;
  constexpr int NSLOT = Cfg::NSLOT, NT0U = Cfg::NT0U, NT0 = Cfg::NT0, NGEN = Cfg::NGEN;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  F3Smem<LV>& S = *reinterpret_cast<F3Smem<LV>*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const bool is_acc = w < F3_ACC;

  LaneBasis<LV> LB;
  if (is_acc) {
    typename F3Smem<LV>::Stage& T = S.st[w];
    const float* X0 = &T.X[0][0][0];
    const float* S0 = &T.SH[0][0][0];
#pragma unroll
    for (int k = 0; k < NSLOT; ++k) {
      const BasisEnt be = p.btab[k * 32 + lane];
      LB.xp[k] = X0 + be.ia;
      if (k >= NT0U) LB.sp[k - NT0U] = S0 + be.ma;
      if (k >= NT0U + NT0) {
        const int q = k - NT0U - NT0;
        LB.gx[2 * q] = X0 + be.ib; LB.gx[2 * q + 1] = X0 + be.ic;
        LB.gs[2 * q] = S0 + be.mb; LB.gs[2 * q + 1] = S0 + be.mc;
        LB.gf[3 * q] = be.fa; LB.gf[3 * q + 1] = be.fb; LB.gf[3 * q + 2] = be.fc;
      }
    }
  }
  if (tid == 0) { S.task[5] = blockIdx.x % F3_NCOMBO; S.task[6] = -1; }
  __syncthreads();

  for (;;) {
    if (tid == 0) {
      int combo = S.task[5], found = 0;
      for (int tries = 0; tries < F3_NCOMBO && !found; ++tries) {
        const int g = combo / NSL;
        const int nblk = (p.gcnt[g] + p.nb_segs - 1) / p.nb_segs;
        if (nblk > 0) {
          const int blk = atomicAdd(p.counters + combo, 1);
          if (blk < nblk) {
            S.task[0] = g; S.task[1] = combo % NSL; S.task[2] = blk * p.nb_segs;
            S.task[3] = min(p.nb_segs, p.gcnt[g] - blk * p.nb_segs);
            S.task[4] = (combo != S.task[6]);
            S.task[5] = combo; S.task[6] = combo;
            found = 1;
            break;
          }
        }
        combo = (combo + 1) % F3_NCOMBO;
      }
      if (!found) S.task[0] = -1;
    }
    __syncthreads();
    const int g = S.task[0];
    if (g < 0) break;
    const int r = S.task[1], idx0 = S.task[2], nseg = S.task[3];
    if (S.task[4]) {
      const float4* src = reinterpret_cast<const float4*>(p.W2S[g] + (size_t)r * Cfg::W * J3);
      float4* dst = reinterpret_cast<float4*>(S.Wsl);
      for (int i = tid; i < Cfg::W * J3 / 4; i += F3_THREADS) dst[i] = src[i];
      if (r == 0)
        for (int i = tid; i < Cfg::W / 4; i += F3_THREADS)
          reinterpret_cast<float4*>(S.Wb)[i] = reinterpret_cast<const float4*>(p.b2p[g])[i];
      for (int i = tid; i < J3 * EA; i += F3_THREADS) S.W1a[i / EA][i % EA] = p.W1[g][(J3 * r + i / EA) * HID + i % EA];
      __syncthreads();
    }
    if (is_acc) {
      if (r == 0) f3_acc_task<LV, true>(p, S, LB, g, r, idx0, nseg, w, lane);
      else f3_acc_task<LV, false>(p, S, LB, g, r, idx0, nseg, w, lane);
    } else {
      if (r == 0) f3_con_task<LV, true>(p, S, r, nseg, w - F3_ACC, lane);
      else f3_con_task<LV, false>(p, S, r, nseg, w - F3_ACC, lane);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------- finalize
struct FinArgs {
  int N, dout;
  const int* seg_cnt;
  const float* part;
  const float* bn_scale; const float* bn_shift;
  const float* x_in; float* x_out;
};

// x_out[node] = bn(mean over the edges of both groups) + x_in[node]  (tensor_layers.py:159-166); the 2 x 9 partial
// outputs of a node are added in a fixed order (group-major, slice ascending).
__global__ void __launch_bounds__(256) k_conv_finalize(FinArgs p) {
  const int q = threadIdx.x / D, f = threadIdx.x % D;
  if (q >= 3) return;
  const int node = blockIdx.x * 3 + q;
  if (node >= p.N) return;
  const int c0 = p.seg_cnt[2 * node], c1 = p.seg_cnt[2 * node + 1];
  float v = 0.f;
  if (f < p.dout) {
    float s = 0.f;
    if (c0 > 0) {
      const float* q0 = p.part + ((size_t)(2 * node) * NSL) * D + f;
#pragma unroll
      for (int r = 0; r < NSL; ++r) s += q0[r * D];
    }
    if (c1 > 0) {
      const float* q1 = p.part + ((size_t)(2 * node + 1) * NSL) * D + f;
#pragma unroll
      for (int r = 0; r < NSL; ++r) s += q1[r * D];
    }
    const float cn = fmaxf((float)(c0 + c1), 1.f);
    v = (s / cn) * p.bn_scale[f] + p.bn_shift[f] + p.x_in[(size_t)node * D + f];
  }
  p.x_out[(size_t)node * D + f] = v;
}

// ---------------------------------------------------------------------------------------------- host side
static void basis_desc_host(int lv, int u, int& type, int& i0, int& m) {
  const int F0e = lv >= 1 ? 30 : 24;
  const int F1o = lv >= 2 ? 36 : (lv == 1 ? 30 : 24);
  const int F1e = lv >= 3 ? 36 : (lv == 2 ? 12 : (lv == 1 ? 6 : 0));
  const int X1O = 24, X1E = 42, X0O = 60;
  type = 0; i0 = 0; m = 0;
  if (u < F0e) {
    if (u < 24) { type = 0; i0 = u; m = 0; } else { type = 1; i0 = X1O + 3 * (u - 24); }
    return;
  }
  u -= F0e;
  if (u < 3 * F1o) {
    int c = u / F1o, k = u % F1o;
    if (k < 24) { type = 0; i0 = k; m = 1 + c; }
    else if (k < 30) { type = 0; i0 = X1O + 3 * (k - 24) + c; m = 0; }
    else { type = 2; i0 = X1E + 3 * (k - 30); m = 1 + c; }
    return;
  }
  u -= 3 * F1o;
  if (F1e > 0 && u < 3 * F1e) {
    int c = u / F1e, k = u % F1e;
    if (k < 6) { type = 2; i0 = X1O + 3 * k; m = 1 + c; }
    else if (k < 12) { type = 0; i0 = X1E + 3 * (k - 6) + c; m = 0; }
    else { type = 0; i0 = X0O + (k - 12); m = 1 + c; }
    return;
  }
  u -= 3 * F1e;
  if (u < 6) { type = 1; i0 = X1E + 3 * u; } else { type = 0; i0 = X0O + (u - 6); m = 0; }
}

// (slot, lane) -> basis row.  Order: per harmonic component m the first floor(count_m / 32) * 32 type-0 rows (slots
// with a compile-time m), then the remaining type-0 rows, then the dot / cross rows; idle lanes get u = -1.
void build_basis_table(int lv, std::vector<BasisEnt>& tab) {
  const int U = lv == 0 ? 96 : (lv == 1 ? 138 : (lv == 2 ? 180 : 276));
  const int nslot = (U + 31) / 32;
  std::vector<int> t0[4], rest, gen, order;
  for (int u = 0; u < U; ++u) {
    int ty, i0, m;
    basis_desc_host(lv, u, ty, i0, m);
    if (ty == 0) t0[m].push_back(u); else gen.push_back(u);
  }
  for (int m = 0; m < 4; ++m) {
    const size_t nu = (t0[m].size() / 32) * 32;
    order.insert(order.end(), t0[m].begin(), t0[m].begin() + nu);
    rest.insert(rest.end(), t0[m].begin() + nu, t0[m].end());
  }
  order.insert(order.end(), rest.begin(), rest.end());
  order.insert(order.end(), gen.begin(), gen.end());
  tab.assign((size_t)nslot * 32, BasisEnt{-1, 0, 0, 0, 0, 0, 0, 0.f, 0.f, 0.f});
  for (size_t q = 0; q < order.size(); ++q) {
    const int u = order[q];
    int ty, i0, m;
    basis_desc_host(lv, u, ty, i0, m);
    BasisEnt e{u, i0, i0, i0, m, 0, 0, 1.f, 0.f, 0.f};
    if (ty == 1) { e.ia = i0; e.ib = i0 + 1; e.ic = i0 + 2; e.ma = 1; e.mb = 2; e.mc = 3; e.fa = e.fb = e.fc = 1.f; }
    if (ty == 2) {
      const int c = m - 1, c1 = (c + 1) % 3, c2 = (c + 2) % 3;
      e.ia = i0 + c1; e.ma = 1 + c2; e.fa = 1.f;
      e.ib = i0 + c2; e.mb = 1 + c1; e.fb = -1.f;
      e.ic = e.ia; e.mc = 0; e.fc = 0.f;
    }
    tab[q] = e;
  }
}

// contiguous, cost-balanced split of the (class, f) rows of a layer over the contraction warps
void build_con_split(const LayerInfo& li, ConSplit& sp) {
  double total = 0;
  for (int k = 0; k < li.ncls; ++k) total += (double)li.cls[k].F * li.cls[k].ncomp * li.cls[k].O;
  for (int w = 0; w < F3_CON; ++w)
    for (int k = 0; k < 4; ++k) { sp.f0[w][k] = 0; sp.f1[w][k] = 0; }
  int w = 0;
  double acc = 0;
  for (int k = 0; k < li.ncls; ++k) {
    const double row = (double)li.cls[k].ncomp * li.cls[k].O;
    int f = 0;
    while (f < li.cls[k].F) {
      const double room = total * (w + 1) / F3_CON - acc;
      int take = (w == F3_CON - 1) ? li.cls[k].F - f : (int)std::max(0.0, std::floor(room / row + 0.5));
      take = std::min(take, li.cls[k].F - f);
      if (take == 0) { if (w < F3_CON - 1) { ++w; continue; } take = li.cls[k].F - f; }
      if (sp.f1[w][k] == sp.f0[w][k]) sp.f0[w][k] = f;
      sp.f1[w][k] = f + take;
      f += take;
      acc += take * row;
      if (acc >= total * (w + 1) / F3_CON - 1e-9 && w < F3_CON - 1) ++w;
    }
  }
}

cudaError_t conv3_configure() {
  cudaError_t e;
  e = cudaFuncSetAttribute(k_conv_fused<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(F3Smem<0>));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_conv_fused<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(F3Smem<1>));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_conv_fused<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(F3Smem<2>));
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_conv_fused<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(F3Smem<3>));
}

void launch_build_group_lists(DdkCtx* c, cudaStream_t st) {
  LaunchScope ls(c, PC_GRAPH, st);
  k_build_group_lists<<<4, 1024, 0, st>>>(c->NL, c->NR, ptr<int>(c->b_seg_cnt), ptr<int>(c->b_seg_base),
                                          ptr<int4>(c->b_glist), ptr<int>(c->b_gcnt));
}

void launch_conv_fused(DdkCtx* c, int layer, const float* x_in, float* x_out, cudaStream_t st) {
  const LayerInfo& li = c->layers[layer];
  F3Args a;
  a.NL = c->NL; a.N = c->N;
  const int nsegs = 2 * c->N;
  int nb = (int)((int64_t)nsegs * NSL / (c->sm_count * 6)) / F3_ACC * F3_ACC;
  a.nb_segs = std::min(128, std::max(F3_ACC, nb));
  a.glist = ptr<int4>(c->b_glist);
  a.goff[0] = 0; a.goff[1] = c->NL; a.goff[2] = 2 * c->NL; a.goff[3] = 2 * c->NL + c->NR;
  a.gcnt = ptr<int>(c->b_gcnt);
  a.counters = ptr<int>(c->b_counters);
  a.seg_list = ptr<int2>(c->b_seg_list);
  a.x = x_in; a.projs = ptr<float>(c->b_proj);
  a.ea_pool = ptr<float>(c->b_ea_pool); a.sh_pool = ptr<float4>(c->b_sh_pool);
  for (int g = 0; g < 4; ++g) {
    a.W1[g] = W(c, conv_id(layer, DDK_WL_W1 + g));
    a.W2S[g] = c->w2s + c->w2s_off[layer * 4 + g];
    a.b2p[g] = W(c, conv_id(layer, DDK_WL_B2P + g));
  }
  a.btab = c->btab + c->btab_off[li.lv];
  a.part = ptr<float>(c->b_part);
  a.split = c->con_split[layer];
  a.ncls = li.ncls;
  int w8 = 0;
  for (int k = 0; k < 4; ++k) {
    a.cls[k] = li.cls[k < li.ncls ? k : 0];
    a.w8off[k] = w8;
    if (k < li.ncls) w8 += li.cls[k].F * J3 * li.cls[k].O;
  }
  cudaMemsetAsync(c->b_counters.p, 0, F3_NCOMBO * sizeof(int), st);
  const int grid = c->sm_count;
  {
    LaunchScope ls(c, PC_ACC0 + li.lv, st);
    switch (li.lv) {
      case 0: k_conv_fused<0><<<grid, F3_THREADS, sizeof(F3Smem<0>), st>>>(a); break;
      case 1: k_conv_fused<1><<<grid, F3_THREADS, sizeof(F3Smem<1>), st>>>(a); break;
      case 2: k_conv_fused<2><<<grid, F3_THREADS, sizeof(F3Smem<2>), st>>>(a); break;
      default: k_conv_fused<3><<<grid, F3_THREADS, sizeof(F3Smem<3>), st>>>(a); break;
    }
  }
  FinArgs f;
  f.N = c->N; f.dout = li.dout;
  f.seg_cnt = ptr<int>(c->b_seg_cnt);
  f.part = ptr<float>(c->b_part);
  f.bn_scale = W(c, conv_id(layer, DDK_WL_BN_SCALE));
  f.bn_shift = W(c, conv_id(layer, DDK_WL_BN_SHIFT));
  f.x_in = x_in; f.x_out = x_out;
  {
    LaunchScope ls(c, PC_CONTRACT, st);
    k_conv_finalize<<<(c->N + 2) / 3, 256, 0, st>>>(f);
  }
}

}  // namespace ddk
