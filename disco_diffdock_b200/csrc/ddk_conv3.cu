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
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ddk_conv.cuh"

namespace ddk {

constexpr int F3_NCOMBO_MAX = 5 * NSL_MAX;

enum { F3_BAR_CON = 1, F3_BAR_CON2 = 2, F3_BAR_PAIR0 = 3 };   // named barriers; + one per warp pair

__device__ __forceinline__ void f3_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// mbarriers between the accumulate warps and the contraction warps: FULL (one arrival per accumulate warp and batch) and
// EMPTY (one arrival per contraction warp and batch).  Unlike a named barrier an accumulate pair only waits for the
// contraction warps, never for the other pairs.
__device__ __forceinline__ unsigned f3_smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void f3_mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared.b64 [%0], %1;" ::"r"(f3_smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void f3_mbar_arrive(unsigned long long* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared.b64 st, [%0];\n\t}" ::"r"(f3_smem_addr(bar)) : "memory");
}
__device__ __forceinline__ bool f3_mbar_try(unsigned long long* bar, int parity) {
  unsigned ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(f3_smem_addr(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// Bulk asynchronous copy global -> shared (TMA engine, SASS UBLKCP) that completes on an mbarrier: the resident weight
// slice of a combo is one contiguous block of up to 69 KB, so a single elected thread moves it with one instruction
// instead of a 640-thread float4 loop.  size: multiple of 16 bytes; dst / src 16-byte aligned.
__device__ __forceinline__ void f3_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(f3_smem_addr(dst)), "l"(src), "r"(bytes), "r"(f3_smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void f3_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}"
               ::"r"(f3_smem_addr(bar)), "r"(bytes) : "memory");
}
// waiting warps back off between probes so they do not take issue slots from the warps they are waiting for
__device__ __forceinline__ void f3_mbar_wait(unsigned long long* bar, int parity) {
  if (f3_mbar_try(bar, parity)) return;
  while (!f3_mbar_try(bar, parity)) __nanosleep(128);
}

// Basis rows are laid out over (slot, lane) by SOURCE feature so that a few shared-memory loads feed many rows and every
// slot has the same instruction sequence on all lanes (no per-lane harmonic index):
//   mixed group (4 slots, m = 0..3): a lane is either a SCALAR lane -- source x[i]; row of slot m is x[i] * sh[m] (for a
//                     vector component x1o/x1e[k][c] only the m = 0 row exists) -- or a VECTOR lane -- source 3-vector v
//                     (x1o[k] or x1e[k]); slot 0 is the dot product v . s, slots 1..3 the cross product (v x s)_c.
//                     Slots 0..3 form the first group; level 3 has a second one (slots 4..7).
//   slot V0         : levels >= 1: rows x1o/x1e[k][c] * sh[0]
//   generic slot    : level 2 only: the 20 left-over rows, 3-term formula with per-lane indices
// A segment is accumulated by a PAIR of warps, each owning about half of the slots (half_of).
template <int LV>
struct F3Cfg {
  static constexpr int U = AccCfg<LV>::U;
  static constexpr int DINP = AccCfg<LV>::DINP;
  static constexpr int XQ = DINP / 4;
  static constexpr int W = LV == 0 ? 720 : (LV == 1 ? 936 : (LV == 2 ? 1152 : 1872));   // second-layer rows (sum F*O)
  static constexpr int J = f3_J(LV);
  static constexpr int NSLV = HID / J;
  static constexpr int AST = J + 1;          // row stride of an A slot: J hidden units + the sum-of-basis (bias) column
  static constexpr int KC = 8;               // edges per gather chunk of a warp pair (16 was measured slower at every level)
  static constexpr bool G0_VEC = LV == 1 || LV == 2;     // the first mixed group has vector lanes
  static constexpr bool HAS_M = LV == 3;                 // second mixed group
  static constexpr bool HAS_V0 = LV >= 1, HAS_G = LV == 2;
  static constexpr int SLOT_M = 4, SLOT_V0 = HAS_M ? 8 : 4, SLOT_G = 5;
  static constexpr int NSLOT = LV == 0 ? 4 : (LV == 1 ? 5 : (LV == 2 ? 6 : 9));
  // half_of(k) = the warp of the pair that owns slot k
  __host__ __device__ static constexpr int half_of(int k) {
    if (LV == 0) return k >= 2;                          // A0 A1 | A2 A3
    if (LV == 1) return (k == 2 || k == 3);              // A0 A1 V0 | A2 A3
    if (LV == 2) return (k == 2 || k == 3 || k == SLOT_G);   // A0 A1 V0 | A2 A3 G
    return k >= SLOT_M && k < SLOT_M + 4;                // A0..A3 V0 | M0..M3
  }
};

template <int LV>
struct F3Smem {
  alignas(16) float Wsl[F3Cfg<LV>::W * F3Cfg<LV>::J];        // [class][f][jj][o] of the resident (group, slice)
  alignas(16) float Wb[F3Cfg<LV>::W];                        // packed second-layer bias (used by slice 0 only)
  alignas(16) float As[F3_ACC][F3Cfg<LV>::U * F3Cfg<LV>::AST];   // one slot per accumulate warp pair: [u][jj | bsum]
  struct Stage {
    alignas(16) float X[2][F3Cfg<LV>::KC][F3Cfg<LV>::DINP];
    alignas(16) float SH[2][F3Cfg<LV>::KC][4];
    alignas(16) float H[2][F3Cfg<LV>::KC][F3Cfg<LV>::J];
  } st[F3_ACC];
  alignas(16) float tile[F3_CON][F3_ACC][D];                 // per contraction warp partial outputs of a batch
  alignas(8) unsigned long long bar_full, bar_empty;         // mbarriers, see f3_mbar_*
  alignas(8) unsigned long long bar_w;                       // completion of the weight-slice bulk copy
  int meta[F3_ACC];                                          // segment id of each slot of the batch in flight (-1: none)
  int task[8];                                               // g, r, idx0, nseg, reload, combo cursor, resident combo, reloads so far
  int ntc;                                                   // the task's segments were accumulated by k_acc_tc: load their A blocks
};

struct F3Args {
  int NL, N;
  int nb_segs;                       // segments per task (multiple of F3_ACC)
  int gmask;                         // bit g set: edge group g is processed (the last layer before the heads: ligand nodes only)
  const int4* glist;                 // per-group lists of non-empty segments: (seg, n, base, 0)
  int goff[4];
  int gci[4];                        // index of each group's count in gcnt (group 2 may use the filtered list)
  const int* gcnt;                   // [4]
  int* counters;                     // [5 * NSLV] segment cursor of each combo
  const int2* seg_list;
  const float* x;                    // [N][84] layer input
  const float* hs;                   // [NSLV][LT][J] hidden units of every listed edge (k_edge_hidden)
  size_t LT;                         // capacity of seg_list
  const float4* sh_pool;
  const float* W2S[4];               // [NSLV][W * J]
  const float* b2p[4];               // [W]
  const LaneTab* ltab;               // [32] lane table of the level
  float* part;                       // [2 N][NSLV][84]
  ConSplit split;
  ClassInfo cls[4];
  int w8off[4];                      // offset of each class inside a weight slice (floats)
  int ncls;
  const float* tc_scratch;           // A_s blocks of the long group-1 segments (k_acc_tc), or null
  const int* tc_n_long;              // how many entries at the head of the group-1 list they cover
  int tc_cap;
  unsigned long long* trace;         // DDK_CONV_TRACE: per CTA (start, end, tasks, weight reloads, reload ns, work ns, claim ns); else null
};

__device__ __forceinline__ unsigned long long f3_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// ---------------------------------------------------------------------------------------------- group work lists
// Non-empty segments of each edge group (block g = group g), bucketed by their number of 8-edge chunks, longest first:
// the eight segments a CTA accumulates side by side then take the same number of chunk iterations, and the long
// segments are claimed first.  The order inside a bucket is whatever the shared-memory cursors hand out; no result
// depends on it (every (segment, slice) partial is produced by one warp pair in a fixed order).
constexpr int GL_BUCKETS = 64;

// need[0][r] = residue r has a cross edge;  need[h][r] = need[h-1][r] or r is a receptor-contact neighbour of such a residue
__global__ void k_need_hop0(int NL, int NR, const int* __restrict__ seg_cnt, unsigned char* __restrict__ need) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < NR) need[r] = seg_cnt[2 * (NL + r) + 1] > 0;
}
__global__ void k_need_expand(int ER, const int* __restrict__ rr_src, const int* __restrict__ rr_dst,
                              const unsigned char* __restrict__ prev, unsigned char* __restrict__ next) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < ER && prev[rr_src[e]]) { next[rr_src[e]] = 1; next[rr_dst[e]] = 1; }   // every write stores 1: no race that matters
}

// blocks 0..3: the four edge groups; block 4 + h: group 2 restricted to the residues of need[h].
// split_sub > 0 (k_conv_tcr): the lig<-rec segments (list 1) are listed in PIECES of at most split_sub edges, in the region at
// split_off; entry = (segment, edges, first list position, record id) with record id = segment for piece 0 (and for every
// unsplit entry) and 2 N + (TCR_PMAX - 1) * ligand node + p - 1 for piece p > 0.
__global__ void __launch_bounds__(1024) k_build_group_lists(int NL, int NR, const int* __restrict__ seg_cnt,
                                                            const int* __restrict__ seg_base, int4* __restrict__ glist,
                                                            int* __restrict__ gcnt, unsigned long long* __restrict__ counters,
                                                            const unsigned char* __restrict__ need, const int tc_min_chunks,
                                                            const int split_sub, const int split_off) {
  __shared__ int hist[GL_BUCKETS], cursor[GL_BUCKETS];
  __shared__ int nedge, nreal;
  const int li = blockIdx.x;                       // work list
  const bool filt = li >= 4;
  const int g = filt ? 2 : li;
  const int nn = g < 2 ? NL : NR;
  const bool split = li == 1 && split_sub > 0;
  const int off = split ? split_off
                        : (li == 0 ? 0 : (li == 1 ? NL : (li == 2 ? 2 * NL : (li == 3 ? 2 * NL + NR : 2 * NL + 2 * NR + (li - 4) * NR))));
  const unsigned char* nd = filt ? need + (size_t)(li - 4) * NR : nullptr;
  const int tid = threadIdx.x;
  auto bucket = [](int n) { return min(GL_BUCKETS - 1, (n + KC3 - 1) / KC3); };
  auto pieces = [&](int n) { return split ? min(TCR_PMAX, (n + split_sub - 1) / split_sub) : 1; };
  auto piece_len = [&](int n, int P, int pc) { return pc < P - 1 ? split_sub : n - (P - 1) * (split ? split_sub : 0); };
  if (tid < GL_BUCKETS) hist[tid] = 0;
  if (tid == 0) { nedge = 0; nreal = 0; }
  __syncthreads();
  for (int i = tid; i < nn; i += 1024) {
    const int seg = g < 2 ? 2 * i + g : 2 * (NL + i) + (g - 2);
    int n = seg_cnt[seg];
    if (filt && !nd[i]) n = 0;
    if (n > 0) {
      const int P = pieces(n);
      for (int pc = 0; pc < P; ++pc) atomicAdd(&hist[bucket(piece_len(n, P, pc))], 1);
      atomicAdd(&nedge, n); atomicAdd(&nreal, 1);
    }
  }
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int bkt = GL_BUCKETS - 1; bkt >= 0; --bkt) { cursor[bkt] = run; run += hist[bkt]; }
    gcnt[li] = run;                                                         // list entries (pieces count as entries)
    gcnt[F3_NLIST + 4 + li] = nedge;                                        // listed edges of this work list in this step
    if (!filt) gcnt[F3_NLIST + li] = cursor[tc_min_chunks - 1];             // segments with >= tc_min_chunks chunks (list head)
    if (!filt) atomicAdd(counters + 1, (unsigned long long)nreal);          // all non-empty segments
    atomicAdd(counters + 2 + li, (unsigned long long)nedge);                // edges per work list
    atomicAdd(counters + 2 + F3_NLIST + li, (unsigned long long)nreal);     // segments per work list
  }
  __syncthreads();
  for (int i = tid; i < nn; i += 1024) {
    const int seg = g < 2 ? 2 * i + g : 2 * (NL + i) + (g - 2);
    int n = seg_cnt[seg];
    if (filt && !nd[i]) n = 0;
    if (n > 0) {
      const int P = pieces(n), base = seg_base[seg];
      for (int pc = 0; pc < P; ++pc) {
        const int len = piece_len(n, P, pc);
        const int o = atomicAdd(&cursor[bucket(len)], 1);
        const int rec = pc == 0 ? seg : 2 * (NL + NR) + (TCR_PMAX - 1) * i + pc - 1;
        glist[off + o] = make_int4(seg, len, base + pc * (split ? split_sub : 0), rec);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- accumulate warps
struct LaneBasis {                   // the part of the LaneTab row a warp keeps in registers
  int oS0, oV0g, oS1, oV1, oVz;      // scalar / vector source columns of the two mixed groups, source column of slot V0
  bool vec0, vec1;                   // the lane is a vector lane of group 0 / 1
  int gx[3], gs[3];                  // generic slot
  float gf[3];
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

// basis values of one mixed group (slots K0..K0+3) restricted to the slots of this half
template <int LV, int HALF, int K0, bool HASVEC>
__device__ __forceinline__ void f3_mixed(float (&b)[F3Cfg<LV>::NSLOT], const float* __restrict__ xrow, const float4 s4,
                                         const int oS, const int oV, const bool isvec) {
  using Cfg = F3Cfg<LV>;
  constexpr bool n0 = Cfg::half_of(K0) == HALF, n1 = Cfg::half_of(K0 + 1) == HALF, n2 = Cfg::half_of(K0 + 2) == HALF,
                 n3 = Cfg::half_of(K0 + 3) == HALF;
  if (!(n0 || n1 || n2 || n3)) return;
  const float xs = xrow[oS];
  float v0 = 0.f, v1 = 0.f, v2 = 0.f;
  if (HASVEC) { v0 = xrow[oV]; v1 = xrow[oV + 1]; v2 = xrow[oV + 2]; }
  if (n0) {
    float t = xs * s4.x;
    if (HASVEC) { float d = v0 * s4.y; d = fmaf(v1, s4.z, d); d = fmaf(v2, s4.w, d); t = isvec ? d : t; }
    b[K0] = t;
  }
  if (n1) {
    float t = xs * s4.y;
    if (HASVEC) { const float c = fmaf(v1, s4.w, -(v2 * s4.z)); t = isvec ? c : t; }
    b[K0 + 1] = t;
  }
  if (n2) {
    float t = xs * s4.z;
    if (HASVEC) { const float c = fmaf(v2, s4.y, -(v0 * s4.w)); t = isvec ? c : t; }
    b[K0 + 2] = t;
  }
  if (n3) {
    float t = xs * s4.w;
    if (HASVEC) { const float c = fmaf(v0, s4.z, -(v1 * s4.y)); t = isvec ? c : t; }
    b[K0 + 3] = t;
  }
}

// one edge: the J hidden units of the slice (as J / 2 packed pairs) against the basis rows of the lane that belong to
// this warp's half of the slots
template <int LV, bool BIAS, int HALF>
__device__ __forceinline__ void f3_edge(f32x2 (&acc)[F3Cfg<LV>::NSLOT][F3Cfg<LV>::J / 2], float (&bs)[F3Cfg<LV>::NSLOT],
                                        const LaneBasis& LB, const float* __restrict__ hrow,
                                        const float* __restrict__ shrow, const float* __restrict__ xrow) {
  using Cfg = F3Cfg<LV>;
  constexpr int J = Cfg::J, NSLOT = Cfg::NSLOT;
  // basis values of the lane's rows for this edge
  float b[NSLOT];
  {
    const float4 s4 = *reinterpret_cast<const float4*>(shrow);
    f3_mixed<LV, HALF, 0, Cfg::G0_VEC>(b, xrow, s4, LB.oS0, LB.oV0g, LB.vec0);
    if (Cfg::HAS_M) f3_mixed<LV, HALF, Cfg::SLOT_M, true>(b, xrow, s4, LB.oS1, LB.oV1, LB.vec1);
    if (Cfg::HAS_V0 && Cfg::half_of(Cfg::SLOT_V0) == HALF) b[Cfg::SLOT_V0] = xrow[LB.oVz] * s4.x;
    if (Cfg::HAS_G && Cfg::half_of(Cfg::SLOT_G) == HALF) {
      float v = LB.gf[0] * (xrow[LB.gx[0]] * shrow[LB.gs[0]]);
      v += LB.gf[1] * (xrow[LB.gx[1]] * shrow[LB.gs[1]]);
      v += LB.gf[2] * (xrow[LB.gx[2]] * shrow[LB.gs[2]]);
      b[Cfg::SLOT_G] = v;
    }
  }
  if (BIAS) {
#pragma unroll
    for (int k = 0; k < NSLOT; ++k)
      if (Cfg::half_of(k) == HALF) bs[k] += b[k];
  }
  constexpr int JH = J > 12 ? 12 : J;             // hidden units per pass (keeps the live set small at J = 24)
#pragma unroll
  for (int j0 = 0; j0 < J; j0 += JH) {
    f32x2 h[JH / 2];
#pragma unroll
    for (int q = 0; q < JH / 4; ++q) {
      const float4 v = *reinterpret_cast<const float4*>(hrow + j0 + 4 * q);
      h[2 * q] = f3_pack2(v.x, v.y); h[2 * q + 1] = f3_pack2(v.z, v.w);
    }
#pragma unroll
    for (int k = 0; k < NSLOT; ++k)
      if (Cfg::half_of(k) == HALF) {
        const f32x2 bb = f3_pack2(b[k], b[k]);
#pragma unroll
        for (int j = 0; j < JH / 2; ++j) f3_ffma2(acc[k][j0 / 2 + j], bb, h[j]);
      }
  }
}

// One warp of the pair `pr` (HALF = 0 / 1).  Both warps run the same chunk sequence in lockstep: the pair's stage is
// gathered cooperatively (cp.async, one chunk ahead) and handed over with one 64-thread named barrier per chunk.
template <int LV, bool BIAS, int HALF>
__device__ __forceinline__ void f3_acc_task(const F3Args& p, F3Smem<LV>& S, const LaneBasis& LB, const int g, const int r,
                                            const int idx0, const int nseg, const int pr, const int lane, int& nflush) {
  using Cfg = F3Cfg<LV>;
  constexpr int NSLOT = Cfg::NSLOT, DINP = Cfg::DINP, J = Cfg::J, AST = Cfg::AST, XQ = Cfg::XQ;
  constexpr int KC = Cfg::KC, GSUBS = 64 / KC;      // edges per chunk; 16-byte sub-pieces copied side by side per edge
  constexpr int XBUF = KC * DINP, SBUF = KC * 4, HBUF = KC * J;
  typename F3Smem<LV>::Stage& T = S.st[pr];
  const int nb = (nseg + F3_ACC - 1) / F3_ACC;
  const int4* wl = p.glist + p.goff[g] + idx0;
  const float* hsr = p.hs + (size_t)r * p.LT * J;   // slice r of the hidden units, list order
  const int pl = HALF * 32 + lane;                  // thread of the pair
  const int ge = pl % KC, gsub = pl / KC;           // gather role: edge, 16-byte sub-piece (0..GSUBS-1)

  f32x2 acc[NSLOT][J / 2];
  float bs[NSLOT];
  int urow[NSLOT];                                  // rows of the A block owned by this lane (this half's slots)
#pragma unroll
  for (int k = 0; k < NSLOT; ++k) {
    bs[k] = 0.f;
    urow[k] = Cfg::half_of(k) == HALF ? p.ltab[lane].u[k] : -1;
#pragma unroll
    for (int j = 0; j < J / 2; ++j) acc[k][j] = 0ull;
  }

  // ---- chunk generator: batches bi = 0..nb-1, this pair's segment of a batch is idx0 + 8 bi + pr (or none)
  int bi = 0, n = 0, sbase = 0, seg = 0, c0 = 0;
  bool in_seg = false;
  int4 pre = (pr < nseg) ? wl[pr] : make_int4(-1, 0, 0, 0);
  auto next_cd = [&]() {
    ChunkD d;
    d.pos = 0; d.kc = 0; d.seg = -1;
    d.flags = CD_LAST | (bi > 0 ? CD_NOT_FIRST_BATCH : 0);
    if (!in_seg) {
      if (bi >= nb) { d.flags = CD_DONE; return d; }
      const int si = F3_ACC * bi + pr;
      if (si >= nseg) { ++bi; return d; }           // no segment for this pair in the batch: empty slot
      seg = pre.x; n = pre.y; sbase = pre.z; c0 = 0; in_seg = true;
      const int sn = si + F3_ACC;
      pre = (sn < nseg) ? wl[sn] : make_int4(-1, 0, 0, 0);
    }
    d.pos = sbase + c0;
    d.kc = min(KC, n - c0);
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
  // the 64 threads of the pair copy fixed 16-byte pieces (q = gsub + GSUBS i) of the destination features of edge ge; the
  // chunk's hidden-unit slice is one contiguous block of kc * J floats (copied by the second warp)
  auto gather = [&](const ChunkD& d, const int2 ent, const int buf) {
    if (d.kc > 0) {
      const int slot = __shfl_sync(0xffffffffu, ent.x, ge);
      const int dst = __shfl_sync(0xffffffffu, ent.y, ge);
      if (ge < d.kc) {
        const float* xs = p.x + (size_t)dst * D + 4 * gsub;
        float* xd = &T.X[buf][ge][4 * gsub];
#pragma unroll
        for (int i = 0; i < (XQ + GSUBS - 1) / GSUBS; ++i)
          if (gsub + GSUBS * i < XQ) f3_cp16(xd + 4 * GSUBS * i, xs + 4 * GSUBS * i);
        if (gsub == 0) f3_cp16(&T.SH[buf][ge][0], p.sh_pool + slot);
      }
      if (HALF == 1) {
#pragma unroll
        for (int i = 0; i < (KC * J / 4 + 31) / 32; ++i)
          if (lane + 32 * i < d.kc * (J / 4))
            f3_cp16(&T.H[buf][0][0] + 4 * (lane + 32 * i), hsr + (size_t)d.pos * J + 4 * (lane + 32 * i));
      }
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
    f3_bar_sync(F3_BAR_PAIR0 + pr, 64);             // both halves of chunk cd0 have landed; buf ^ 1 is consumed
    // ---- next chunk's gathers travel while this chunk is accumulated
    gather(cd1, ent1, buf ^ 1);
    ChunkD cd2 = (cd1.flags & CD_DONE) ? cd1 : next_cd();
    const int2 ent2 = load_ent(cd2);

    if (cd0.kc > 0) {
      const float* hb = &T.H[0][0][0] + buf * HBUF;
      const float* sb = &T.SH[0][0][0] + buf * SBUF;
      const float* xb = &T.X[0][0][0] + buf * XBUF;
      if (cd0.kc == KC) {
#pragma unroll
        for (int e = 0; e < KC; ++e) f3_edge<LV, BIAS, HALF>(acc, bs, LB, hb + e * J, sb + e * 4, xb + e * DINP);
      } else if (KC > 8 && cd0.kc == 8) {           // the 8-edge tail of the 24-edge receptor-contact segments
#pragma unroll
        for (int e = 0; e < 8; ++e) f3_edge<LV, BIAS, HALF>(acc, bs, LB, hb + e * J, sb + e * 4, xb + e * DINP);
      } else {
#pragma unroll 1
        for (int e = 0; e < cd0.kc; ++e) f3_edge<LV, BIAS, HALF>(acc, bs, LB, hb + e * J, sb + e * 4, xb + e * DINP);
      }
    }
    if (cd0.flags & CD_LAST) {
      // ---- hand the finished U x J block to the contraction warps
      if (nflush > 0) f3_mbar_wait(&S.bar_empty, (nflush - 1) & 1);   // the contraction warps are done with the previous batch
      if (cd0.flags & CD_VALID) {
        float* slot = &S.As[pr][0];
#pragma unroll
        for (int k = 0; k < NSLOT; ++k)
          if (Cfg::half_of(k) == HALF) {
            const int u = urow[k];
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
      if (HALF == 0 && lane == 0) S.meta[pr] = cd0.seg;
      __syncwarp();
      if (lane == 0) f3_mbar_arrive(&S.bar_full);
      ++nflush;
    }
    // ---- rotate
    buf ^= 1;
    cd0 = cd1; cd1 = cd2; ent1 = ent2;
  }
  __pipeline_wait_prior(0);
}

// A task made only of segments that k_acc_tc accumulated on the tensor cores (the claim logic never mixes the two kinds): the
// pair copies the segment's (U x (J | bsum)) block of this slice from the scratch into its slot and hands it to the
// contraction warps with the same mbarrier protocol as f3_acc_task.  Kept apart so that the FFMA2 loop carries no extra state.
template <int LV>
__device__ __noinline__ void f3_tc_task(const F3Args& p, F3Smem<LV>& S, const int r, const int idx0, const int nseg, const int pr,
                                        const int pl, int& nflush) {
  using Cfg = F3Cfg<LV>;
  constexpr int NF = Cfg::U * Cfg::AST, NFP = (NF + 3) & ~3, VW = NF % 4 == 0 ? 4 : 2;
  const int nb = (nseg + F3_ACC - 1) / F3_ACC;
  const int4* wl = p.glist + p.goff[1] + idx0;
  for (int bi = 0; bi < nb; ++bi) {
    const int si = F3_ACC * bi + pr;
    if (nflush > 0) f3_mbar_wait(&S.bar_empty, (nflush - 1) & 1);
    if (si < nseg) {
      const float* src = p.tc_scratch + ((size_t)(idx0 + si) * Cfg::NSLV + r) * NFP;
      float* slot = &S.As[pr][0];
      if (VW == 4) {
        for (int i = pl; i < NF / 4; i += 64) reinterpret_cast<float4*>(slot)[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
      } else {
        for (int i = pl; i < NF / 2; i += 64) reinterpret_cast<float2*>(slot)[i] = __ldg(reinterpret_cast<const float2*>(src) + i);
      }
    }
    if (pl == 0) S.meta[pr] = si < nseg ? wl[si].x : -1;
    __syncwarp();
    if ((pl & 31) == 0) f3_mbar_arrive(&S.bar_full);
    ++nflush;
  }
}

// ---------------------------------------------------------------------------------------------- contraction warps
// scalar output class (O = 24, one component): lane = (k-part kp = lane >> 2, output group og = lane & 3 -> 6 outputs),
// 8 segments per lane; rows (f, jj) of the class are dealt round-robin to the 8 k-parts.  The loads of the next row
// travel while the current one is multiplied; the 8 k-parts are summed with a halving exchange (24 + 12 + 6 shuffles),
// after which lane (kp, og) holds the 6 outputs of segment kp.
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
  // rows are processed in ping-pong pairs: the loads of one row are in flight while the other is multiplied
#define F3_CS_LOAD(W_, A_)                                                                                              \
  {                                                                                                                     \
    const float* wp = (!BIAS || jj < J) ? Wc + (f * J + jj) * 24 + 6 * og : Wbc + f * 24 + 6 * og;                      \
    W_##0 = *reinterpret_cast<const f32x2*>(wp);                                                                        \
    W_##1 = *reinterpret_cast<const f32x2*>(wp + 2);                                                                    \
    W_##2 = *reinterpret_cast<const f32x2*>(wp + 4);                                                                    \
    const float* ap = As + (uoff + f) * AST + jj;                                                                       \
    A_##0 = ap[0]; A_##1 = ap[ASLOT]; A_##2 = ap[2 * ASLOT]; A_##3 = ap[3 * ASLOT];                                     \
    A_##4 = ap[4 * ASLOT]; A_##5 = ap[5 * ASLOT]; A_##6 = ap[6 * ASLOT]; A_##7 = ap[7 * ASLOT];                         \
    jj += 8;                                                                                                            \
    if (jj >= JC) { jj -= JC; ++f; }                                                                                    \
  }
#define F3_CS_FMA1(S_, AV_, W_)                                                                                         \
  {                                                                                                                     \
    const f32x2 t = f3_pack2(AV_, AV_);                                                                                 \
    f3_ffma2(acc[S_][0], t, W_##0); f3_ffma2(acc[S_][1], t, W_##1); f3_ffma2(acc[S_][2], t, W_##2);                     \
  }
#define F3_CS_FMA(W_, A_)                                                                                               \
  F3_CS_FMA1(0, A_##0, W_) F3_CS_FMA1(1, A_##1, W_) F3_CS_FMA1(2, A_##2, W_) F3_CS_FMA1(3, A_##3, W_)                   \
  F3_CS_FMA1(4, A_##4, W_) F3_CS_FMA1(5, A_##5, W_) F3_CS_FMA1(6, A_##6, W_) F3_CS_FMA1(7, A_##7, W_)
  static_assert(F3_ACC == 8, "f3_con_scalar is written for 8 segments per batch");
  f32x2 wa0 = 0ull, wa1 = 0ull, wa2 = 0ull, wb0 = 0ull, wb1 = 0ull, wb2 = 0ull;
  float aa0 = 0.f, aa1 = 0.f, aa2 = 0.f, aa3 = 0.f, aa4 = 0.f, aa5 = 0.f, aa6 = 0.f, aa7 = 0.f;
  float ab0 = 0.f, ab1 = 0.f, ab2 = 0.f, ab3 = 0.f, ab4 = 0.f, ab5 = 0.f, ab6 = 0.f, ab7 = 0.f;
  if (kp < nrows) F3_CS_LOAD(wa, aa)
  for (int q = kp; q < nrows; q += 16) {
    const bool hb = q + 8 < nrows;
    if (hb) F3_CS_LOAD(wb, ab)
    F3_CS_FMA(wa, aa)
    if (hb) {
      if (q + 16 < nrows) F3_CS_LOAD(wa, aa)
      F3_CS_FMA(wb, ab)
    }
  }
#undef F3_CS_LOAD
#undef F3_CS_FMA1
#undef F3_CS_FMA
  // ---- sum over the 8 k-parts: after the three exchanges lane kp holds segment s = kp
  float v[F3_ACC][6];
#pragma unroll
  for (int s = 0; s < F3_ACC; ++s)
#pragma unroll
    for (int o = 0; o < 3; ++o) f3_unpack2(acc[s][o], v[s][2 * o], v[s][2 * o + 1]);
  const bool b2 = (lane & 16) != 0, b1 = (lane & 8) != 0, b0 = (lane & 4) != 0;
  float r4[4][6], r2[2][6], r1[6];
#pragma unroll
  for (int s = 0; s < 4; ++s)
#pragma unroll
    for (int o = 0; o < 6; ++o) {
      const float send = b2 ? v[s][o] : v[s + 4][o];
      const float keep = b2 ? v[s + 4][o] : v[s][o];
      r4[s][o] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
  for (int s = 0; s < 2; ++s)
#pragma unroll
    for (int o = 0; o < 6; ++o) {
      const float send = b1 ? r4[s][o] : r4[s + 2][o];
      const float keep = b1 ? r4[s + 2][o] : r4[s][o];
      r2[s][o] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
  for (int o = 0; o < 6; ++o) {
    const float send = b0 ? r2[0][o] : r2[1][o];
    const float keep = b0 ? r2[1][o] : r2[0][o];
    r1[o] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  float* tp = tile + kp * D + col0 + 6 * og;    // segment 4 b2 + 2 b1 + b0 = kp
#pragma unroll
  for (int o = 0; o < 6; ++o) tp[o] += r1[o];
}

// vector output class (O = 6, three components sharing the weights): lane = (kp = lane >> 2, sg = lane & 3 -> segments
// 2 sg, 2 sg + 1), 2 x 3 x 6 accumulators per lane; same pipelining, halving exchanges over (segment, output half), then a
// plain exchange for the last step.
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
#define F3_CV_LOAD(W_, A_)                                                                                              \
  {                                                                                                                     \
    const float* wp = (!BIAS || jj < J) ? Wc + (f * J + jj) * 6 : Wbc + f * 6;                                          \
    W_##0 = *reinterpret_cast<const f32x2*>(wp);                                                                        \
    W_##1 = *reinterpret_cast<const f32x2*>(wp + 2);                                                                    \
    W_##2 = *reinterpret_cast<const f32x2*>(wp + 4);                                                                    \
    const float* ap = A0 + (uoff + f) * AST + jj;                                                                       \
    A_##0 = ap[0]; A_##1 = ap[F * AST]; A_##2 = ap[2 * F * AST];                                                        \
    A_##3 = ap[ASLOT]; A_##4 = ap[ASLOT + F * AST]; A_##5 = ap[ASLOT + 2 * F * AST];                                    \
    jj += 8;                                                                                                            \
    if (jj >= JC) { jj -= JC; ++f; }                                                                                    \
  }
#define F3_CV_FMA1(S_, C_, AV_, W_)                                                                                     \
  {                                                                                                                     \
    const f32x2 t = f3_pack2(AV_, AV_);                                                                                 \
    f3_ffma2(acc[S_][C_][0], t, W_##0); f3_ffma2(acc[S_][C_][1], t, W_##1); f3_ffma2(acc[S_][C_][2], t, W_##2);         \
  }
#define F3_CV_FMA(W_, A_)                                                                                               \
  F3_CV_FMA1(0, 0, A_##0, W_) F3_CV_FMA1(0, 1, A_##1, W_) F3_CV_FMA1(0, 2, A_##2, W_)                                   \
  F3_CV_FMA1(1, 0, A_##3, W_) F3_CV_FMA1(1, 1, A_##4, W_) F3_CV_FMA1(1, 2, A_##5, W_)
  f32x2 wa0 = 0ull, wa1 = 0ull, wa2 = 0ull, wb0 = 0ull, wb1 = 0ull, wb2 = 0ull;
  float aa0 = 0.f, aa1 = 0.f, aa2 = 0.f, aa3 = 0.f, aa4 = 0.f, aa5 = 0.f;
  float ab0 = 0.f, ab1 = 0.f, ab2 = 0.f, ab3 = 0.f, ab4 = 0.f, ab5 = 0.f;
  if (kp < nrows) F3_CV_LOAD(wa, aa)
  for (int q = kp; q < nrows; q += 16) {
    const bool hb = q + 8 < nrows;
    if (hb) F3_CV_LOAD(wb, ab)
    F3_CV_FMA(wa, aa)
    if (hb) {
      if (q + 16 < nrows) F3_CV_LOAD(wa, aa)
      F3_CV_FMA(wb, ab)
    }
  }
#undef F3_CV_LOAD
#undef F3_CV_FMA1
#undef F3_CV_FMA
  float v[2][3][6];
#pragma unroll
  for (int s = 0; s < 2; ++s)
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int o = 0; o < 3; ++o) f3_unpack2(acc[s][c][o], v[s][c][2 * o], v[s][c][2 * o + 1]);
  const bool b2 = (lane & 16) != 0, b1 = (lane & 8) != 0, b0 = (lane & 4) != 0;
  float r1[3][6], r2[3][3];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int o = 0; o < 6; ++o) {               // b2 selects the segment of the pair
      const float send = b2 ? v[0][c][o] : v[1][c][o];
      const float keep = b2 ? v[1][c][o] : v[0][c][o];
      r1[c][o] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int o = 0; o < 3; ++o) {               // b1 selects the output half
      const float send = b1 ? r1[c][o] : r1[c][o + 3];
      const float keep = b1 ? r1[c][o + 3] : r1[c][o];
      r2[c][o] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int o = 0; o < 3; ++o) r2[c][o] += __shfl_xor_sync(0xffffffffu, r2[c][o], 4);
  if (!b0) {
    float* tp = tile + (2 * sg + (b2 ? 1 : 0)) * D + col0 + 3 * (b1 ? 3 : 0);
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int o = 0; o < 3; ++o) tp[3 * o + c] += r2[c][o];
  }
}

template <int LV, bool BIAS>
__device__ __forceinline__ void f3_con_task(const F3Args& p, F3Smem<LV>& S, const int r, const int nseg, const int cw,
                                            const int lane, int& nbatch) {
  using Cfg = F3Cfg<LV>;
  constexpr int ASLOT = Cfg::U * Cfg::AST;
  const int nb = (nseg + F3_ACC - 1) / F3_ACC;
  const int ct = cw * 32 + lane;
  float* tile = &S.tile[cw][0][0];
  for (int b = 0; b < nb; ++b) {
    for (int i = lane; i < F3_ACC * D; i += 32) tile[i] = 0.f;
    f3_mbar_wait(&S.bar_full, nbatch & 1);           // the 8 slots of batch b are written
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
    __syncwarp();
    if (lane == 0) f3_mbar_arrive(&S.bar_empty);      // this warp is done with the slots and meta
    ++nbatch;
    f3_bar_sync(F3_BAR_CON2, F3_CON * 32);            // tiles may be cleared
  }
}

// ---------------------------------------------------------------------------------------------- the kernel
template <int LV>
__global__ void __launch_bounds__(F3_THREADS, 1) k_conv_fused(const __grid_constant__ F3Args p) {
  using Cfg = F3Cfg<LV>;
  constexpr int J = Cfg::J, NSLV = Cfg::NSLV, NCOMBO = 5 * Cfg::NSLV;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  F3Smem<LV>& S = *reinterpret_cast<F3Smem<LV>*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const bool is_acc = w < 2 * F3_ACC;
  const int pr = w >> 1, half = w & 1;

  LaneBasis LB;
  {
    const LaneTab& lt = p.ltab[lane];
    LB.oS0 = lt.oS[0]; LB.oV0g = lt.oV[0]; LB.vec0 = lt.isvec[0] != 0;
    LB.oS1 = lt.oS[1]; LB.oV1 = lt.oV[1]; LB.vec1 = lt.isvec[1] != 0;
    LB.oVz = lt.oV0;
#pragma unroll
    for (int k = 0; k < 3; ++k) { LB.gx[k] = lt.gi[k]; LB.gs[k] = lt.gm[k]; LB.gf[k] = lt.gf[k]; }
  }
  if (tid == 0) {
    S.task[5] = blockIdx.x % NCOMBO; S.task[6] = -1; S.task[7] = 0;
    f3_mbar_init(&S.bar_full, 2 * F3_ACC);
    f3_mbar_init(&S.bar_empty, F3_CON);
    f3_mbar_init(&S.bar_w, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");   // the init is visible to the async proxy
  }
  __syncthreads();
  int nbat = 0;                      // batches handed over so far (same count on both sides of the mbarriers)
#if DDK_CONV_TRACE     // build with DDK_NVCC_EXTRA=-DDDK_CONV_TRACE=1 (tools/conv_trace.sh); the counters cost registers
  unsigned long long tr_t0 = 0, tr_a = 0, tr_reload = 0, tr_work = 0, tr_claim = 0, tr_tasks = 0, tr_nrel = 0;
  const bool tracing = p.trace != nullptr && tid == 0;
  if (tracing) tr_t0 = f3_now();
#define F3_TRACE(...) if (tracing) { __VA_ARGS__ }
#else
#define F3_TRACE(...)
#endif

  for (;;) {
    F3_TRACE(tr_a = f3_now();)
    if (tid == 0) {
      // combos: (work range, slice).  Ranges 0..3 = the four edge groups; range 4 = the head of the group-1 list that k_acc_tc
      // accumulated on the tensor cores (range 1 then starts behind it), so a task holds one kind of segment only
      int combo = S.task[5], found = 0;
      for (int tries = 0; tries < NCOMBO && !found; ++tries) {
        const int g5 = combo / NSLV, g = g5 == 4 ? 1 : g5;
        const int gall = ((p.gmask >> g) & 1) ? p.gcnt[p.gci[g]] : 0;
        const int ntc = (g == 1 && p.tc_scratch != nullptr) ? min(min(*p.tc_n_long, p.tc_cap), gall) : 0;
        const int lo = g5 == 1 ? ntc : 0, gn = (g5 == 4 ? ntc : gall) - lo;
        // guided self-scheduling on the combo's segment cursor: blocks shrink as the combo runs out, so the CTAs of a
        // launch finish within a few segments of each other
        const int rem = gn - *reinterpret_cast<volatile int*>(p.counters + combo);
        if (rem > 0) {
          const int size = max(F3_ACC, min(p.nb_segs, (rem / 8) / F3_ACC * F3_ACC));
          const int start = atomicAdd(p.counters + combo, size);
          if (start < gn) {
            const int wcombo = g * NSLV + combo % NSLV;                  // which weight slice the task needs
            S.ntc = g5 == 4;
            S.task[0] = g; S.task[1] = combo % NSLV; S.task[2] = lo + start;
            S.task[3] = min(size, gn - start);
            S.task[4] = (wcombo != S.task[6]);
            if (wcombo != S.task[6]) S.task[7] += 1;
            S.task[5] = combo; S.task[6] = wcombo;
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
    F3_TRACE(const unsigned long long t = f3_now(); tr_claim += t - tr_a; tr_a = t; ++tr_tasks;)
    if (S.task[4]) {
      // every read of the previous slice happened before the __syncthreads that ended the previous task; the proxy fence
      // orders those generic-proxy reads before the async-proxy writes of the copy
      if (tid == 0) {
        constexpr unsigned wbytes = Cfg::W * J * sizeof(float), bbytes = Cfg::W * sizeof(float);
        static_assert(wbytes % 16 == 0 && bbytes % 16 == 0, "bulk copies move multiples of 16 bytes");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        f3_mbar_expect_tx(&S.bar_w, wbytes + (r == 0 ? bbytes : 0u));
        f3_bulk_g2s(S.Wsl, p.W2S[g] + (size_t)r * Cfg::W * J, wbytes, &S.bar_w);
        if (r == 0) f3_bulk_g2s(S.Wb, p.b2p[g], bbytes, &S.bar_w);
      }
      f3_mbar_wait(&S.bar_w, (S.task[7] - 1) & 1);
      F3_TRACE(const unsigned long long t = f3_now(); tr_reload += t - tr_a; tr_a = t; ++tr_nrel;)
    }
    if (is_acc && S.ntc) {
      f3_tc_task<LV>(p, S, r, idx0, nseg, pr, half * 32 + lane, nbat);
    } else if (is_acc) {
      if (half == 0) {
        if (r == 0) f3_acc_task<LV, true, 0>(p, S, LB, g, r, idx0, nseg, pr, lane, nbat);
        else f3_acc_task<LV, false, 0>(p, S, LB, g, r, idx0, nseg, pr, lane, nbat);
      } else {
        if (r == 0) f3_acc_task<LV, true, 1>(p, S, LB, g, r, idx0, nseg, pr, lane, nbat);
        else f3_acc_task<LV, false, 1>(p, S, LB, g, r, idx0, nseg, pr, lane, nbat);
      }
    } else {
      if (r == 0) f3_con_task<LV, true>(p, S, r, nseg, w - 2 * F3_ACC, lane, nbat);
      else f3_con_task<LV, false>(p, S, r, nseg, w - 2 * F3_ACC, lane, nbat);
    }
    __syncthreads();
    F3_TRACE(tr_work += f3_now() - tr_a;)
  }
  F3_TRACE(unsigned long long* o = p.trace + 8 * blockIdx.x;
           o[0] = tr_t0; o[1] = f3_now(); o[2] = tr_tasks; o[3] = tr_nrel; o[4] = tr_reload; o[5] = tr_work; o[6] = tr_claim;)
#undef F3_TRACE
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

// rows in PROCESSING order of k_acc_tc: plain products first, then dot products, then cross products; u = row of the A block
// (kernel order), -1 = padding (evaluates x[0] * sh[0]; the result is never read)
void build_tc_rows(int lv, TcRow* rows) {
  const int U = lv == 0 ? 96 : (lv == 1 ? 138 : (lv == 2 ? 180 : 276));
  int n = 0;
  for (int pass = 0; pass < 3; ++pass)
    for (int u = 0; u < U; ++u) {
      int ty, i0, m;
      basis_desc_host(lv, u, ty, i0, m);
      if (ty == pass) rows[n++] = TcRow{(short)ty, (short)i0, (short)m, (short)u};
    }
  for (; n < TC_MAXROWS; ++n) rows[n] = TcRow{0, 0, 0, -1};
}

// lane -> (rows, sources) of a basis level, see F3Cfg.  Returns false unless every row is owned exactly once.
bool build_lane_table(int lv, LaneTab* tab) {
  const int U = lv == 0 ? 96 : (lv == 1 ? 138 : (lv == 2 ? 180 : 276));
  const int nslot = lv == 0 ? 4 : (lv == 1 ? 5 : (lv == 2 ? 6 : 9));
  const int slotV0 = lv == 3 ? 8 : 4, slotG = 5;
  std::vector<int> t0row(84 * 4, -1), dotrow(84, -1), crossrow(84 * 3, -1);
  for (int u = 0; u < U; ++u) {
    int ty, i0, m;
    basis_desc_host(lv, u, ty, i0, m);
    if (ty == 0) t0row[i0 * 4 + m] = u;
    else if (ty == 1) dotrow[i0] = u;
    else crossrow[i0 * 3 + (m - 1)] = u;
  }
  for (int l = 0; l < 32; ++l) {
    LaneTab& t = tab[l];
    for (int k = 0; k < F3_MAXSLOT; ++k) t.u[k] = -1;
    for (int q = 0; q < 2; ++q) { t.oS[q] = 0; t.oV[q] = 0; t.isvec[q] = 0; }
    t.oV0 = 0;
    for (int k = 0; k < 3; ++k) { t.gi[k] = 0; t.gm[k] = 0; t.gf[k] = 0.f; }
  }
  std::vector<int> used(U, 0);
  auto own = [&](int lane, int slot, int u) { if (u >= 0) { tab[lane].u[slot] = u; used[u]++; } };
  auto scalar_lane = [&](int grp, int lane, int i) {
    tab[lane].oS[grp] = i; tab[lane].isvec[grp] = 0;
    for (int m = 0; m < 4; ++m) own(lane, 4 * grp + m, t0row[i * 4 + m]);
  };
  auto vector_lane = [&](int grp, int lane, int i0) {
    tab[lane].oV[grp] = i0; tab[lane].isvec[grp] = 1;
    own(lane, 4 * grp, dotrow[i0]);
    for (int c = 0; c < 3; ++c) own(lane, 4 * grp + 1 + c, crossrow[i0 * 3 + c]);
  };
  std::vector<int> vecs, vrows;                    // 3-vector sources (first column), vector components (x * sh[0] rows)
  if (lv >= 1) for (int k = 0; k < 6; ++k) vecs.push_back(24 + 3 * k);
  if (lv >= 2) for (int k = 0; k < 6; ++k) vecs.push_back(42 + 3 * k);
  if (lv >= 1) for (int i = 24; i < (lv >= 2 ? 60 : 42); ++i) vrows.push_back(i);
  for (int l = 0; l < 24; ++l) scalar_lane(0, l, l);                       // x0e
  size_t nv = 0, nr = 0;
  if (lv == 1 || lv == 2)
    for (int l = 24; l < 32 && nv < vecs.size(); ++l) vector_lane(0, l, vecs[nv++]);
  if (lv == 3) {
    for (int l = 24; l < 32; ++l) scalar_lane(0, l, 60 + (l - 24));         // x0o 0..7
    for (int l = 0; l < 12; ++l) vector_lane(1, l, vecs[nv++]);
    for (int l = 12; l < 28; ++l) scalar_lane(1, l, 68 + (l - 12));         // x0o 8..23
  }
  if (lv >= 1)
    for (int l = 0; l < 32 && nr < vrows.size(); ++l) {
      const int i = vrows[nr++];
      tab[l].oV0 = i;
      own(l, slotV0, t0row[i * 4]);
    }
  if (lv == 3)
    for (int l = 28; l < 32 && nr < vrows.size(); ++l) scalar_lane(1, l, vrows[nr++]);   // only their m = 0 rows exist
  if (lv == 2) {
    int l = 0;
    for (; nr < vrows.size(); ++nr, ++l) {
      const int i = vrows[nr];
      tab[l].gi[0] = tab[l].gi[1] = tab[l].gi[2] = i; tab[l].gm[0] = 0; tab[l].gf[0] = 1.f;
      own(l, slotG, t0row[i * 4]);
    }
    for (; nv < vecs.size(); ++nv) {
      const int i0 = vecs[nv];
      if (dotrow[i0] >= 0) {
        if (l >= 32) return false;
        for (int k = 0; k < 3; ++k) { tab[l].gi[k] = i0 + k; tab[l].gm[k] = 1 + k; tab[l].gf[k] = 1.f; }
        own(l++, slotG, dotrow[i0]);
      }
      for (int c = 0; c < 3; ++c, ++l) {
        if (l >= 32) return false;
        const int c1 = (c + 1) % 3, c2 = (c + 2) % 3;
        tab[l].gi[0] = i0 + c1; tab[l].gm[0] = 1 + c2; tab[l].gf[0] = 1.f;
        tab[l].gi[1] = i0 + c2; tab[l].gm[1] = 1 + c1; tab[l].gf[1] = -1.f;
        tab[l].gi[2] = i0; tab[l].gm[2] = 0; tab[l].gf[2] = 0.f;
        own(l, slotG, crossrow[i0 * 3 + c]);
      }
    }
  }
  if (nv != vecs.size() || nr != vrows.size()) return false;
  for (int u = 0; u < U; ++u)
    if (used[u] != 1) return false;
  for (int l = 0; l < 32; ++l)
    for (int k = nslot; k < F3_MAXSLOT; ++k)
      if (tab[l].u[k] >= 0) return false;
  return true;
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

void launch_build_group_lists(DdkCtx* c, cudaStream_t st, bool with_needed) {
  unsigned char* need = ptr<unsigned char>(c->b_need);
  const int nhop = with_needed ? c->nhop : 0;
  if (nhop > 0) {
    LaunchScope ls(c, PC_GRAPH, st);
    k_need_hop0<<<(c->NR + 255) / 256, 256, 0, st>>>(c->NL, c->NR, ptr<int>(c->b_seg_cnt), need);
  }
  for (int h = 1; h < nhop; ++h) {
    cudaMemsetAsync(need + (size_t)h * c->NR, 0, c->NR, st);
    if (c->ER == 0) { cudaMemcpyAsync(need + (size_t)h * c->NR, need + (size_t)(h - 1) * c->NR, c->NR, cudaMemcpyDeviceToDevice, st); continue; }   // no receptor contacts: the set does not grow (and a zero-size grid is an invalid launch)
    LaunchScope ls(c, PC_GRAPH, st);
    k_need_expand<<<(c->ER + 255) / 256, 256, 0, st>>>(c->ER, ptr<int>(c->b_rr_src), ptr<int>(c->b_rr_dst),
                                                      need + (size_t)(h - 1) * c->NR, need + (size_t)h * c->NR);
  }
  // DDK_TC_MIN_CHUNKS: shortest segment (in 8-edge chunks) that takes the tensor-core path (default TC_MIN_CHUNKS)
  static const int tc_min_chunks =
      std::min(GL_BUCKETS - 1, std::max(1, getenv("DDK_TC_MIN_CHUNKS") ? atoi(getenv("DDK_TC_MIN_CHUNKS")) : TC_MIN_CHUNKS));
  LaunchScope ls(c, PC_GRAPH, st);
  k_build_group_lists<<<4 + nhop, 1024, 0, st>>>(c->NL, c->NR, ptr<int>(c->b_seg_cnt), ptr<int>(c->b_seg_base),
                                                 ptr<int4>(c->b_glist), ptr<int>(c->b_gcnt),
                                                 ptr<unsigned long long>(c->b_edge_total), need, tc_min_chunks,
                                                 tcr_split_sub(), glist_split_off(c));
}

// k_conv_tcr accumulates the lig<-rec segments in pieces (list 1 lives in its own region of the work-list buffer then)
int tcr_split_sub() {
  static const int sub = getenv("DDK_TCR_SUB") ? std::max(0, atoi(getenv("DDK_TCR_SUB"))) / KC3 * KC3 : TCR_SUB;
  return conv_path() == 2 ? sub : 0;
}
int glist_split_off(const DdkCtx* c) { return 2 * c->N + c->nhop * c->NR; }
int glist_off_group1(const DdkCtx* c) { return tcr_split_sub() > 0 ? glist_split_off(c) : c->NL; }

void launch_conv_finalize(DdkCtx* c, int layer, const float* x_in, float* x_out, cudaStream_t st, bool lig_only, int nsl) {
  const LayerInfo& li = c->layers[layer];
  FinArgs f;
  f.N = lig_only ? c->NL : c->N; f.dout = li.dout; f.nsl = nsl;   // ligand nodes come first
  f.seg_cnt = ptr<int>(c->b_seg_cnt);
  f.part = ptr<float>(c->b_part);
  f.bn_scale = W(c, conv_id(layer, DDK_WL_BN_SCALE));
  f.bn_shift = W(c, conv_id(layer, DDK_WL_BN_SHIFT));
  f.x_in = x_in; f.x_out = x_out;
  LaunchScope ls(c, PC_CONTRACT, st);
  k_conv_finalize<<<(f.N + 2) / 3, 256, 0, st>>>(f);
}

void launch_conv_fused(DdkCtx* c, int layer, const float* x_in, float* x_out, cudaStream_t st, int mode) {
  const bool lig_only = mode == CONV_LIG;
  const LayerInfo& li = c->layers[layer];
  F3Args a;
  a.NL = c->NL; a.N = c->N;
  a.gmask = lig_only ? 0x3 : 0xf;
  const int nsegs = 2 * (lig_only ? c->NL : c->N);
  const int J = f3_J(li.lv), nsl = f3_nsl(li.lv);
  static const int tasks_per_cta = getenv("DDK_TASKS_PER_CTA") ? atoi(getenv("DDK_TASKS_PER_CTA")) : 12;
  int nb = (int)((int64_t)nsegs * nsl / (c->sm_count * tasks_per_cta)) / F3_ACC * F3_ACC;
  a.nb_segs = std::min(128, std::max(F3_ACC, nb));
  a.glist = ptr<int4>(c->b_glist);
  a.goff[0] = 0; a.goff[1] = c->NL; a.goff[2] = 2 * c->NL; a.goff[3] = 2 * c->NL + c->NR;
  for (int g = 0; g < 4; ++g) a.gci[g] = g;
  if (mode >= CONV_NEEDED) { const int h = mode - CONV_NEEDED; a.goff[2] = 2 * c->NL + 2 * c->NR + h * c->NR; a.gci[2] = 4 + h; }
  a.gcnt = ptr<int>(c->b_gcnt);
  a.counters = ptr<int>(c->b_counters);
  a.seg_list = ptr<int2>(c->b_seg_list);
  a.x = x_in; a.hs = ptr<float>(c->b_hs); a.LT = (size_t)c->list_total;
  a.sh_pool = ptr<float4>(c->b_sh_pool);
  for (int g = 0; g < 4; ++g) {
    a.W2S[g] = c->w2s + c->w2s_off[layer * 4 + g];
    a.b2p[g] = W(c, conv_id(layer, DDK_WL_B2P + g));
  }
  a.ltab = c->ltab + li.lv * 32;
  a.part = ptr<float>(c->b_part);
  a.split = c->con_split[layer];
  a.ncls = li.ncls;
  int w8 = 0;
  for (int k = 0; k < 4; ++k) {
    a.cls[k] = li.cls[k < li.ncls ? k : 0];
    a.w8off[k] = w8;
    if (k < li.ncls) w8 += li.cls[k].F * J * li.cls[k].O;
  }
  const bool tc = conv_path() == 1 && c->tc_cap > 0;
  a.tc_scratch = tc ? ptr<float>(c->b_tc_scratch) : nullptr;
  a.tc_n_long = ptr<int>(c->b_gcnt) + F3_NLIST + 1;
  a.tc_cap = c->tc_cap;
  if (tc) launch_acc_tc(c, layer, x_in, st);
  cudaMemsetAsync(c->b_counters.p, 0, F3_NCOMBO_MAX * sizeof(int), st);
  const int grid = c->sm_count;
#if DDK_CONV_TRACE
  static const bool trace_on = getenv("DDK_CONV_TRACE") != nullptr;     // per-CTA timeline summary on stderr (synchronises)
#else
  static const bool trace_on = false;
#endif
  static unsigned long long* trace_buf = nullptr;
  if (trace_on && !trace_buf) cudaMalloc(&trace_buf, (size_t)grid * 8 * sizeof(unsigned long long));
  a.trace = trace_on ? trace_buf : nullptr;
  {
    LaunchScope ls(c, PC_ACC0 + li.lv, st);
    switch (li.lv) {
      case 0: k_conv_fused<0><<<grid, F3_THREADS, sizeof(F3Smem<0>), st>>>(a); break;
      case 1: k_conv_fused<1><<<grid, F3_THREADS, sizeof(F3Smem<1>), st>>>(a); break;
      case 2: k_conv_fused<2><<<grid, F3_THREADS, sizeof(F3Smem<2>), st>>>(a); break;
      default: k_conv_fused<3><<<grid, F3_THREADS, sizeof(F3Smem<3>), st>>>(a); break;
    }
  }
  if (trace_on && trace_buf) {
    std::vector<unsigned long long> h((size_t)grid * 8);
    cudaStreamSynchronize(st);
    cudaMemcpy(h.data(), trace_buf, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    unsigned long long t0 = ~0ull, t1 = 0, first_end = ~0ull;
    double busy = 0, tasks = 0, nrel = 0, rel = 0, work = 0, claim = 0;
    for (int b = 0; b < grid; ++b) {
      const unsigned long long* o = &h[(size_t)b * 8];
      t0 = std::min(t0, o[0]); t1 = std::max(t1, o[1]); first_end = std::min(first_end, o[1]);
      busy += (double)(o[1] - o[0]); tasks += (double)o[2]; nrel += (double)o[3]; rel += (double)o[4]; work += (double)o[5]; claim += (double)o[6];
    }
    const double span = (double)(t1 - t0);
    fprintf(stderr, "[conv_trace] layer %d lv %d mode %2d B %4d span %8.1f us  resident %5.1f%%  first CTA done at %5.1f%%  per CTA: tasks %5.1f "
                    "reloads %4.1f  work %5.1f%%  reload %4.1f%%  claim %4.1f%% of span\n",
            layer, li.lv, mode, c->B, span / 1e3, 100.0 * busy / grid / span, 100.0 * (double)(first_end - t0) / span, tasks / grid,
            nrel / grid, 100.0 * work / grid / span, 100.0 * rel / grid / span, 100.0 * claim / grid / span);
  }
  launch_conv_finalize(c, layer, x_in, x_out, st, lig_only, nsl);
}

}  // namespace ddk
