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
//     memory) and does 4 packed FFMA2 (fma.rn.f32x2: scalar basis value x pair of hidden units) per row against the 8
//     hidden units h_e[slice] of the edge.  h_e (first radial-MLP layer) is produced once per layer for every listed
//     edge by k_edge_hidden (ddk_hidden.cu) in slice-major list order, so a chunk's slice is one contiguous block;
//   - at the end of the segment the U x 8 block (+ the Bsum column in slice 0) goes to a per-warp shared-memory slot;
//   - four contraction warps take the 8 slots of a batch together (so every weight read from shared memory is used for
//     8 segments), contract them against the resident weight slice and write the 84-wide PARTIAL output of
//     (segment, slice) to HBM: 336 B instead of 80 KB.
// k_conv_finalize adds the 2 x (72 / J) partials of a node in a fixed order, applies mean / batch-norm / residual.
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

constexpr int F3_NCOMBO_MAX = 4 * NSL_MAX;

enum { F3_BAR_FULL = 1, F3_BAR_EMPTY = 2, F3_BAR_CON = 3, F3_BAR_CON2 = 4 };

__device__ __forceinline__ void f3_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void f3_bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// Basis rows are laid out over (slot, lane) by SOURCE feature so that one shared-memory load feeds several rows:
//   slots 0..3  (A) : lane -> scalar source x[i] (x0e, at level 3 also the first 8 x0o); row of slot m is x[i] * sh[m]
//   slots 4,5   (B) : level 3, the other 16 x0o sources: lane = (i & 15, mh); rows x[i] * sh[mh], x[i] * sh[2 + mh]
//   slot  V0        : levels 2,3: the first 32 rows  x1o/x1e[k][c] * sh[0]
//   2 generic slots : everything else (dot / cross products and the left-over products), 3-term formula per lane
template <int LV>
struct F3Cfg {
  static constexpr int U = AccCfg<LV>::U;
  static constexpr int DINP = AccCfg<LV>::DINP;
  static constexpr int XQ = DINP / 4;
  static constexpr int W = LV == 0 ? 720 : (LV == 1 ? 936 : (LV == 2 ? 1152 : 1872));   // second-layer rows (sum F*O)
  static constexpr int J = f3_J(LV);
  static constexpr int JQ = J / 4;
  static constexpr int NSLV = HID / J;
  static constexpr int AST = J + 1;          // row stride of an A slot: J hidden units + the sum-of-basis (bias) column
  static constexpr bool HAS_B = LV == 3, HAS_V0 = LV >= 2;
  static constexpr int NGEN = LV == 0 ? 0 : 2;
  static constexpr int SLOT_B = 4, SLOT_V0 = 4 + (HAS_B ? 2 : 0), SLOT_G = SLOT_V0 + (HAS_V0 ? 1 : 0);
  static constexpr int NSLOT = SLOT_G + NGEN;
};

template <int LV>
struct F3Smem {
  alignas(16) float Wsl[F3Cfg<LV>::W * F3Cfg<LV>::J];        // [class][f][jj][o] of the resident (group, slice)
  alignas(16) float Wb[F3Cfg<LV>::W];                        // packed second-layer bias (used by slice 0 only)
  alignas(16) float As[F3_ACC][F3Cfg<LV>::U * F3Cfg<LV>::AST];   // one slot per accumulate warp: [u][jj | bsum]
  struct Stage {
    alignas(16) float X[2][KC3][F3Cfg<LV>::DINP];
    alignas(16) float SH[2][KC3][4];
    alignas(16) float H[2][KC3][F3Cfg<LV>::J];
  } st[F3_ACC];
  alignas(16) float tile[F3_CON][F3_ACC][D];                 // per contraction warp partial outputs of a batch
  int meta[F3_ACC];                                          // segment id of each slot of the batch in flight (-1: none)
  int task[8];                                               // g, r, idx0, nseg, reload, combo cursor, resident combo
};

struct F3Args {
  int NL, N;
  int nb_segs;                       // segments per task (multiple of F3_ACC)
  const int4* glist;                 // per-group lists of non-empty segments: (seg, n, base, 0)
  int goff[4];
  const int* gcnt;                   // [4]
  int* counters;                     // [4 * NSLV] next block of each combo
  const int2* seg_list;
  const float* x;                    // [N][84] layer input
  const float* hs;                   // [NSLV][LT][J] hidden units of every listed edge (k_edge_hidden)
  size_t LT;                         // capacity of seg_list
  const float4* sh_pool;
  const float* W2S[4];               // [NSLV][W * J]
  const float* b2p[4];               // [W]
  const BasisEnt* btab;              // [NSLOT * 32]
  float* part;                       // [2 N][NSLV][84]
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
  static constexpr int NG3 = F3Cfg<LV>::NGEN > 0 ? 3 * F3Cfg<LV>::NGEN : 1;
  int oA;                            // column of the lane's scalar source inside a staged feature row (slots A)
  int oB;                            // slots B
  int oV0;                           // slot V0
  int gx[NG3];                       // generic slots: x term columns
  int gs[NG3];                       //                harmonic term indices (0..3)
  float gf[NG3];                     //                coefficients (0, +1, -1)
  bool mh;                           // lane >> 4 (slots B)
};

struct ChunkD {                       // one gather chunk (<= KC3 consecutive list entries of a segment) of an accumulate warp
  int pos, kc, seg, flags;           // first list position, edges, segment id (valid when CD_LAST), CD_* flags
};
enum { CD_VALID = 1, CD_LAST = 2, CD_DONE = 4, CD_NOT_FIRST_BATCH = 8 };

__device__ __forceinline__ void f3_cp16(void* dst, const void* src) { __pipeline_memcpy_async(dst, src, 16); }

// packed fp32 pairs (sm_100a FFMA2): d.{x,y} += a.{x,y} * b.{x,y}; with a = (v, v) ptxas emits the scalar-broadcast form
typedef unsigned long long f32x2;
__device__ __forceinline__ void f3_ffma2(f32x2& d, const f32x2 a, const f32x2 b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
__device__ __forceinline__ f32x2 f3_pack2(const float x, const float y) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
  return r;
}
__device__ __forceinline__ void f3_unpack2(const f32x2 v, float& x, float& y) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v));
}

// one edge: the J hidden units of the slice (as J / 2 packed pairs) against every basis row of the lane
template <int LV, bool BIAS>
__device__ __forceinline__ void f3_edge(f32x2 (&acc)[F3Cfg<LV>::NSLOT][F3Cfg<LV>::J / 2], float (&bs)[F3Cfg<LV>::NSLOT],
                                        const LaneBasis<LV>& LB, const float* __restrict__ hrow,
                                        const float* __restrict__ shrow, const float* __restrict__ xrow) {
  using Cfg = F3Cfg<LV>;
  constexpr int J = Cfg::J, NSLOT = Cfg::NSLOT;
  // basis values of the lane's rows for this edge
  float b[NSLOT];
  {
    const float4 s4 = *reinterpret_cast<const float4*>(shrow);
    const float xa = xrow[LB.oA];
    b[0] = xa * s4.x; b[1] = xa * s4.y; b[2] = xa * s4.z; b[3] = xa * s4.w;
    if (Cfg::HAS_B) {
      const float xb = xrow[LB.oB];
      b[Cfg::SLOT_B] = xb * (LB.mh ? s4.y : s4.x);
      b[Cfg::SLOT_B + 1] = xb * (LB.mh ? s4.w : s4.z);
    }
    if (Cfg::HAS_V0) b[Cfg::SLOT_V0] = xrow[LB.oV0] * s4.x;
#pragma unroll
    for (int q = 0; q < Cfg::NGEN; ++q) {
      float v = LB.gf[3 * q] * (xrow[LB.gx[3 * q]] * shrow[LB.gs[3 * q]]);
      v += LB.gf[3 * q + 1] * (xrow[LB.gx[3 * q + 1]] * shrow[LB.gs[3 * q + 1]]);
      v += LB.gf[3 * q + 2] * (xrow[LB.gx[3 * q + 2]] * shrow[LB.gs[3 * q + 2]]);
      b[Cfg::SLOT_G + q] = v;
    }
  }
  if (BIAS) {
#pragma unroll
    for (int k = 0; k < NSLOT; ++k) bs[k] += b[k];
  }
  f32x2 h[J / 2];
#pragma unroll
  for (int q = 0; q < J / 4; ++q) {
    const float4 v = *reinterpret_cast<const float4*>(hrow + 4 * q);
    h[2 * q] = f3_pack2(v.x, v.y); h[2 * q + 1] = f3_pack2(v.z, v.w);
  }
#pragma unroll
  for (int k = 0; k < NSLOT; ++k) {
    const f32x2 bb = f3_pack2(b[k], b[k]);
#pragma unroll
    for (int j = 0; j < J / 2; ++j) f3_ffma2(acc[k][j], bb, h[j]);
  }
}

template <int LV, bool BIAS>
__device__ __forceinline__ void f3_acc_task(const F3Args& p, F3Smem<LV>& S, LaneBasis<LV>& LB, const int g, const int r,
                                            const int idx0, const int nseg, const int w, const int lane) {
  using Cfg = F3Cfg<LV>;
  constexpr int NSLOT = Cfg::NSLOT, DINP = Cfg::DINP, J = Cfg::J, AST = Cfg::AST, XQ = Cfg::XQ;
  constexpr int XBUF = KC3 * DINP, SBUF = KC3 * 4, HBUF = KC3 * J;
  typename F3Smem<LV>::Stage& T = S.st[w];
  const int nb = (nseg + F3_ACC - 1) / F3_ACC;
  const int4* wl = p.glist + p.goff[g] + idx0;
  const float* hsr = p.hs + (size_t)r * p.LT * J;   // slice r of the hidden units, list order
  const int ge = lane & 7, gsub = lane >> 3;        // gather role of the lane: edge, 16-byte sub-piece

  f32x2 acc[NSLOT][J / 2];
  float bs[NSLOT];
#pragma unroll
  for (int k = 0; k < NSLOT; ++k) {
    bs[k] = 0.f;
#pragma unroll
    for (int j = 0; j < J / 2; ++j) acc[k][j] = 0ull;
  }

  // ---- chunk generator: batches bi = 0..nb-1, this warp's segment of a batch is idx0 + 8 bi + w (or none)
  int bi = 0, n = 0, sbase = 0, seg = 0, c0 = 0;
  bool in_seg = false;
  int4 pre = (w < nseg) ? wl[w] : make_int4(-1, 0, 0, 0);
  auto next_cd = [&]() {
    ChunkD d;
    d.pos = 0; d.kc = 0; d.seg = -1;
    d.flags = CD_LAST | (bi > 0 ? CD_NOT_FIRST_BATCH : 0);
    if (!in_seg) {
      if (bi >= nb) { d.flags = CD_DONE; return d; }
      const int si = F3_ACC * bi + w;
      if (si >= nseg) { ++bi; return d; }           // no segment for this warp in the batch: empty slot
      seg = pre.x; n = pre.y; sbase = pre.z; c0 = 0; in_seg = true;
      const int sn = si + F3_ACC;
      pre = (sn < nseg) ? wl[sn] : make_int4(-1, 0, 0, 0);
    }
    d.pos = sbase + c0;
    d.kc = min(KC3, n - c0);
    d.seg = seg;
    c0 += d.kc;
    d.flags = CD_VALID | (bi > 0 ? CD_NOT_FIRST_BATCH : 0);
    if (c0 >= n) { d.flags |= CD_LAST; in_seg = false; ++bi; }
    return d;
  };
  auto load_ent = [&](const ChunkD& d) {
    int2 e = make_int2(0, 0);
    if (lane < d.kc) e = p.seg_list[d.pos + lane];
    return e;
  };
  // every lane copies fixed 16-byte pieces (q = gsub + 4 i) of the destination features of its edge ge (two shuffles per
  // chunk, immediate offsets); the chunk's hidden-unit slice is one contiguous block of kc * J floats
  auto gather = [&](const ChunkD& d, const int2 ent, const int buf) {
    if (d.kc > 0) {
      const int slot = __shfl_sync(0xffffffffu, ent.x, ge);
      const int dst = __shfl_sync(0xffffffffu, ent.y, ge);
      if (ge < d.kc) {
        const float* xs = p.x + (size_t)dst * D + 4 * gsub;
        float* xd = &T.X[buf][ge][4 * gsub];
#pragma unroll
        for (int i = 0; i < (XQ + 3) / 4; ++i)
          if (gsub + 4 * i < XQ) f3_cp16(xd + 16 * i, xs + 16 * i);
        if (gsub == 0) f3_cp16(&T.SH[buf][ge][0], p.sh_pool + slot);
      }
      if (lane < d.kc * (J / 4))
        f3_cp16(&T.H[buf][0][0] + 4 * lane, hsr + (size_t)d.pos * J + 4 * lane);
    }
    __pipeline_commit();
  };

  ChunkD cd0 = next_cd();
  ChunkD cd1 = (cd0.flags & CD_DONE) ? cd0 : next_cd();
  int2 ent1;
  {
    const int2 ent0 = load_ent(cd0);
    gather(cd0, ent0, 0);
    ent1 = load_ent(cd1);
  }
  int buf = 0;
  while (!(cd0.flags & CD_DONE)) {
    __pipeline_wait_prior(0);
    __syncwarp();
    // ---- next chunk's gathers travel while this chunk is accumulated (the other buffer was consumed last iteration)
    gather(cd1, ent1, buf ^ 1);
    ChunkD cd2 = (cd1.flags & CD_DONE) ? cd1 : next_cd();
    const int2 ent2 = load_ent(cd2);

    if (cd0.kc > 0) {
      const float* hb = &T.H[0][0][0] + buf * HBUF;
      const float* sb = &T.SH[0][0][0] + buf * SBUF;
      const float* xb = &T.X[0][0][0] + buf * XBUF;
      if (cd0.kc == KC3) {
#pragma unroll
        for (int e = 0; e < KC3; ++e) f3_edge<LV, BIAS>(acc, bs, LB, hb + e * J, sb + e * 4, xb + e * DINP);
      } else {
#pragma unroll 1
        for (int e = 0; e < cd0.kc; ++e) f3_edge<LV, BIAS>(acc, bs, LB, hb + e * J, sb + e * 4, xb + e * DINP);
      }
    }
    if (cd0.flags & CD_LAST) {
      // ---- hand the finished U x J block to the contraction warps
      if (cd0.flags & CD_NOT_FIRST_BATCH) f3_bar_sync(F3_BAR_EMPTY, F3_THREADS);   // they are done with the previous batch
      if (cd0.flags & CD_VALID) {
        float* slot = &S.As[w][0];
#pragma unroll
        for (int k = 0; k < NSLOT; ++k) {
          const int u = p.btab[k * 32 + lane].u;
          if (u >= 0) {
#pragma unroll
            for (int j = 0; j < J / 2; ++j) {
              float v0, v1;
              f3_unpack2(acc[k][j], v0, v1);
              slot[u * AST + 2 * j] = v0; slot[u * AST + 2 * j + 1] = v1;
            }
            if (BIAS) slot[u * AST + J] = bs[k];
          }
          bs[k] = 0.f;
#pragma unroll
          for (int j = 0; j < J / 2; ++j) acc[k][j] = 0ull;
        }
      }
      if (lane == 0) S.meta[w] = cd0.seg;
      __threadfence_block();
      f3_bar_arrive(F3_BAR_FULL, F3_THREADS);
    }
    // ---- rotate
    buf ^= 1;
    cd0 = cd1; cd1 = cd2; ent1 = ent2;
  }
  __pipeline_wait_prior(0);
}

// ---------------------------------------------------------------------------------------------- contraction warps
// scalar output class (O = 24, one component): lane = (k-part kp = lane >> 2, output group og = lane & 3 -> 6 outputs),
// 8 segments per lane; rows (f, jj) of the class are dealt round-robin to the 8 k-parts.
template <bool BIAS, int J, int ASLOT>
__device__ __forceinline__ void f3_con_scalar(const float* __restrict__ Wc, const float* __restrict__ Wbc,
                                              const float* __restrict__ As, int uoff, int f0, int f1, float* tile, int col0,
                                              int lane) {
  constexpr int JC = BIAS ? J + 1 : J;
  constexpr int AST = J + 1;
  const int kp = lane >> 2, og = lane & 3;
  f32x2 acc[F3_ACC][3];
#pragma unroll
  for (int s = 0; s < F3_ACC; ++s)
#pragma unroll
    for (int o = 0; o < 3; ++o) acc[s][o] = 0ull;
  const int nrows = (f1 - f0) * JC;
  int f = f0, jj = kp;                          // kp < 8 <= JC
  for (int q = kp; q < nrows; q += 8) {
    const float* wp = (!BIAS || jj < J) ? Wc + (f * J + jj) * 24 + 6 * og : Wbc + f * 24 + 6 * og;
    const f32x2 w0 = *reinterpret_cast<const f32x2*>(wp);
    const f32x2 w1 = *reinterpret_cast<const f32x2*>(wp + 2);
    const f32x2 w2 = *reinterpret_cast<const f32x2*>(wp + 4);
    const float* ap = As + (uoff + f) * AST + jj;
#pragma unroll
    for (int s = 0; s < F3_ACC; ++s) {
      const float a = ap[s * ASLOT];
      const f32x2 aa = f3_pack2(a, a);
      f3_ffma2(acc[s][0], aa, w0); f3_ffma2(acc[s][1], aa, w1); f3_ffma2(acc[s][2], aa, w2);
    }
    jj += 8;
    if (jj >= JC) { jj -= JC; ++f; }
  }
#pragma unroll
  for (int s = 0; s < F3_ACC; ++s)
#pragma unroll
    for (int o = 0; o < 6; ++o) {
      float v0, v1;
      f3_unpack2(acc[s][o >> 1], v0, v1);
      float v = (o & 1) ? v1 : v0;
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (kp == 0) tile[s * D + col0 + 6 * og + o] += v;
    }
}

// vector output class (O = 6, three components sharing the weights): lane = (kp = lane >> 2, sg = lane & 3 -> segments
// 2 sg, 2 sg + 1), 2 x 3 x 6 accumulators per lane.
template <bool BIAS, int J, int ASLOT>
__device__ __forceinline__ void f3_con_vector(const float* __restrict__ Wc, const float* __restrict__ Wbc,
                                              const float* __restrict__ As, int uoff, int F, int f0, int f1, float* tile,
                                              int col0, int lane) {
  constexpr int JC = BIAS ? J + 1 : J;
  constexpr int AST = J + 1;
  const int kp = lane >> 2, sg = lane & 3;
  f32x2 acc[2][3][3];
#pragma unroll
  for (int s = 0; s < 2; ++s)
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int o = 0; o < 3; ++o) acc[s][c][o] = 0ull;
  const int nrows = (f1 - f0) * JC;
  int f = f0, jj = kp;
  const float* A0 = As + (2 * sg) * ASLOT;
  for (int q = kp; q < nrows; q += 8) {
    const float* wp = (!BIAS || jj < J) ? Wc + (f * J + jj) * 6 : Wbc + f * 6;
    const f32x2 w0 = *reinterpret_cast<const f32x2*>(wp);
    const f32x2 w1 = *reinterpret_cast<const f32x2*>(wp + 2);
    const f32x2 w2 = *reinterpret_cast<const f32x2*>(wp + 4);
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float a = A0[s * ASLOT + (uoff + c * F + f) * AST + jj];
        const f32x2 aa = f3_pack2(a, a);
        f3_ffma2(acc[s][c][0], aa, w0); f3_ffma2(acc[s][c][1], aa, w1); f3_ffma2(acc[s][c][2], aa, w2);
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
        float v0, v1;
        f3_unpack2(acc[s][c][o >> 1], v0, v1);
        float v = (o & 1) ? v1 : v0;
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        if (kp == 0) tile[(2 * sg + s) * D + col0 + 3 * o + c] += v;
      }
}

template <int LV, bool BIAS>
__device__ __forceinline__ void f3_con_task(const F3Args& p, F3Smem<LV>& S, const int r, const int nseg, const int cw,
                                            const int lane) {
  using Cfg = F3Cfg<LV>;
  constexpr int ASLOT = Cfg::U * Cfg::AST;
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
      if (ci.ncomp == 1) f3_con_scalar<BIAS, Cfg::J, ASLOT>(Wc, Wbc, &S.As[0][0], ci.uoff, f0, f1, tile, ci.col0, lane);
      else f3_con_vector<BIAS, Cfg::J, ASLOT>(Wc, Wbc, &S.As[0][0], ci.uoff, ci.F, f0, f1, tile, ci.col0, lane);
    }
    f3_bar_sync(F3_BAR_CON, F3_CON * 32);             // every contraction warp's tile is complete
    for (int i = ct; i < F3_ACC * D; i += F3_CON * 32) {
      const int s = i / D, f = i % D;
      const int sid = S.meta[s];
      if (sid >= 0) {
        const float v = ((S.tile[0][s][f] + S.tile[1][s][f]) + S.tile[2][s][f]) + S.tile[3][s][f];
        p.part[((size_t)sid * Cfg::NSLV + r) * D + f] = v;
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
  constexpr int J = Cfg::J, NSLV = Cfg::NSLV, NCOMBO = 4 * Cfg::NSLV;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  F3Smem<LV>& S = *reinterpret_cast<F3Smem<LV>*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const bool is_acc = w < F3_ACC;

  LaneBasis<LV> LB;
  if (is_acc) {
    LB.oA = p.btab[lane].ia;
    LB.oB = Cfg::HAS_B ? p.btab[Cfg::SLOT_B * 32 + lane].ia : 0;
    LB.oV0 = Cfg::HAS_V0 ? p.btab[Cfg::SLOT_V0 * 32 + lane].ia : 0;
    LB.mh = (lane & 16) != 0;
#pragma unroll
    for (int q = 0; q < Cfg::NGEN; ++q) {
      const BasisEnt be = p.btab[(Cfg::SLOT_G + q) * 32 + lane];
      LB.gx[3 * q] = be.ia; LB.gx[3 * q + 1] = be.ib; LB.gx[3 * q + 2] = be.ic;
      LB.gs[3 * q] = be.ma; LB.gs[3 * q + 1] = be.mb; LB.gs[3 * q + 2] = be.mc;
      LB.gf[3 * q] = be.fa; LB.gf[3 * q + 1] = be.fb; LB.gf[3 * q + 2] = be.fc;
    }
  }
  if (tid == 0) { S.task[5] = blockIdx.x % NCOMBO; S.task[6] = -1; }
  __syncthreads();

  for (;;) {
    if (tid == 0) {
      int combo = S.task[5], found = 0;
      for (int tries = 0; tries < NCOMBO && !found; ++tries) {
        const int g = combo / NSLV;
        const int nblk = (p.gcnt[g] + p.nb_segs - 1) / p.nb_segs;
        if (nblk > 0) {
          const int blk = atomicAdd(p.counters + combo, 1);
          if (blk < nblk) {
            S.task[0] = g; S.task[1] = combo % NSLV; S.task[2] = blk * p.nb_segs;
            S.task[3] = min(p.nb_segs, p.gcnt[g] - blk * p.nb_segs);
            S.task[4] = (combo != S.task[6]);
            S.task[5] = combo; S.task[6] = combo;
            found = 1;
            break;
          }
        }
        combo = (combo + 1) % NCOMBO;
      }
      if (!found) S.task[0] = -1;
    }
    __syncthreads();
    const int g = S.task[0];
    if (g < 0) break;
    const int r = S.task[1], idx0 = S.task[2], nseg = S.task[3];
    if (S.task[4]) {
      const float4* src = reinterpret_cast<const float4*>(p.W2S[g] + (size_t)r * Cfg::W * J);
      float4* dst = reinterpret_cast<float4*>(S.Wsl);
      for (int i = tid; i < Cfg::W * J / 4; i += F3_THREADS) dst[i] = src[i];
      if (r == 0)
        for (int i = tid; i < Cfg::W / 4; i += F3_THREADS)
          reinterpret_cast<float4*>(S.Wb)[i] = reinterpret_cast<const float4*>(p.b2p[g])[i];
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
  int N, dout, nsl;
  const int* seg_cnt;
  const float* part;
  const float* bn_scale; const float* bn_shift;
  const float* x_in; float* x_out;
};

// x_out[node] = bn(mean over the edges of both groups) + x_in[node]  (tensor_layers.py:159-166); the 2 x (72 / J) partial
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
      const float* q0 = p.part + ((size_t)(2 * node) * p.nsl) * D + f;
      for (int r = 0; r < p.nsl; ++r) s += q0[r * D];
    }
    if (c1 > 0) {
      const float* q1 = p.part + ((size_t)(2 * node + 1) * p.nsl) * D + f;
      for (int r = 0; r < p.nsl; ++r) s += q1[r * D];
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

// (slot, lane) -> basis row, by source feature (see F3Cfg); idle lanes get u = -1.
void build_basis_table(int lv, std::vector<BasisEnt>& tab) {
  const int U = lv == 0 ? 96 : (lv == 1 ? 138 : (lv == 2 ? 180 : 276));
  const bool hasB = lv == 3, hasV0 = lv >= 2;
  const int ngen = lv == 0 ? 0 : 2;
  const int slotB = 4, slotV0 = 4 + (hasB ? 2 : 0), slotG = slotV0 + (hasV0 ? 1 : 0), nslot = slotG + ngen;
  std::vector<int> t0row(84 * 4, -1);
  std::vector<char> used(U, 0);
  for (int u = 0; u < U; ++u) {
    int ty, i0, m;
    basis_desc_host(lv, u, ty, i0, m);
    if (ty == 0) t0row[i0 * 4 + m] = u;
  }
  tab.assign((size_t)nslot * 32, BasisEnt{-1, 0, 0, 0, 0, 0, 0, 0.f, 0.f, 0.f});
  auto put_t0 = [&](int slot, int lane, int i0, int m) {
    const int u = t0row[i0 * 4 + m];
    tab[(size_t)slot * 32 + lane] = BasisEnt{u, i0, i0, i0, m, 0, 0, 1.f, 0.f, 0.f};
    if (u >= 0) used[u] = 1;
  };
  std::vector<int> ssrc, vsrc;
  for (int i = 0; i < 24; ++i) ssrc.push_back(i);                       // x0e
  if (lv == 3) for (int i = 0; i < 24; ++i) ssrc.push_back(60 + i);     // x0o
  if (lv >= 1) for (int i = 24; i < 42; ++i) vsrc.push_back(i);         // x1o[k][c]
  if (lv >= 2) for (int i = 42; i < 60; ++i) vsrc.push_back(i);         // x1e[k][c]
  for (int l = 0; l < 32 && l < (int)ssrc.size(); ++l)
    for (int m = 0; m < 4; ++m) put_t0(m, l, ssrc[l], m);
  if (hasB)
    for (int l = 0; l < 32; ++l) {
      const int src = ssrc[32 + (l & 15)], mh = l >> 4;
      put_t0(slotB, l, src, mh);
      put_t0(slotB + 1, l, src, 2 + mh);
    }
  if (hasV0)
    for (int l = 0; l < 32; ++l) put_t0(slotV0, l, vsrc[l], 0);
  int q = 0;
  for (int u = 0; u < U; ++u) {
    if (used[u]) continue;
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
    if (q < ngen * 32) tab[(size_t)(slotG + q / 32) * 32 + q % 32] = e;
    ++q;
  }
  if (q > ngen * 32) tab.clear();     // cannot happen for lv 0..3 (checked by ddk_create)
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
  const int J = f3_J(li.lv), nsl = f3_nsl(li.lv);
  int nb = (int)((int64_t)nsegs * nsl / (c->sm_count * 6)) / F3_ACC * F3_ACC;
  a.nb_segs = std::min(128, std::max(F3_ACC, nb));
  a.glist = ptr<int4>(c->b_glist);
  a.goff[0] = 0; a.goff[1] = c->NL; a.goff[2] = 2 * c->NL; a.goff[3] = 2 * c->NL + c->NR;
  a.gcnt = ptr<int>(c->b_gcnt);
  a.counters = ptr<int>(c->b_counters);
  a.seg_list = ptr<int2>(c->b_seg_list);
  a.x = x_in; a.hs = ptr<float>(c->b_hs); a.LT = (size_t)c->list_total;
  a.sh_pool = ptr<float4>(c->b_sh_pool);
  for (int g = 0; g < 4; ++g) {
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
    if (k < li.ncls) w8 += li.cls[k].F * J * li.cls[k].O;
  }
  cudaMemsetAsync(c->b_counters.p, 0, F3_NCOMBO_MAX * sizeof(int), st);
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
  f.N = c->N; f.dout = li.dout; f.nsl = nsl;
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
