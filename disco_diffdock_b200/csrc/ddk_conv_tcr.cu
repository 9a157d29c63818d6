// k_conv_tcr: one tensor-product convolution layer with the outer-product accumulation of EVERY segment on the 5th-generation
// tensor cores and the contraction against the second radial-MLP layer in the same CTA, straight from tensor memory.
//
// Algebra (ddk_conv.cu, /root/reference/models/tensor_layers.py:65-116, 147-168): for a (node s, edge group g) segment
//     out_s = W2p (*) A_s + b2p (*) Bsum_s,   A_s[u][j] = sum_e basis_e[u] * h_e[j],   Bsum_s[u] = sum_e basis_e[u].
// A_s (U x 73 fp32, 80 KB at level 3) is the accumulator of a GEMM with K = the edges of the segment; the contraction needs
// the whole packed weight block of the edge group (539 KB at level 3), which fits neither shared memory nor a second trip
// through HBM / L2 (89 KB per segment: k_acc_tc + k_conv_fused pay exactly that for the long segments).  Here the work of a
// layer is cut into ROLES = (a set of irrep classes, a range of hidden units) such that
//   * the basis rows of a role fit ONE M = 128 accumulator tile (<= 128 rows),
//   * its slice of the packed weights (<= 150 KB) stays RESIDENT in shared memory while a CTA works on the role,
//   * several segments' accumulators (N = 80 / 48 / 32 columns each) sit side by side in tensor memory.
// Level 3: roles {1o}, {1e} (108 rows x all 72 hidden units + the ones column) and {0e, 0o} x three thirds of the hidden units
// (60 rows x 24 + ones); the lower levels analogously with the scalar classes in halves (build_tcr_roles).  Every segment is visited
// once per role; each visit writes a partial record (TCR_PS floats) that k_conv_finalize_tcr adds in a fixed order.
// The long lig<-rec segments are listed in pieces of at most TCR_SUB edges (k_build_group_lists): the tensor core updates its fp32
// accumulator with truncation, and short chains keep that one-sided error at the level of the fp32 FMA chain.
// Scheduling: per (edge group, role) a cursor into the group's segment list; a CTA claims the next block of the role of its group
// whose cursor is furthest behind, so the roles walk the list side by side (later visits of an edge hit L2) and finish together.
// Warp roles inside a CTA (768 threads):
//   * 3 gather warps: list entries, destination feature rows, harmonics and the role's hidden units of 8 edges -> staging ring
//     (cp.async, completion on mbarriers; they run ahead across segment boundaries);
//   * 3 sets of 4 row warps: thread p evaluates basis row p of the 8 edges, splits it into TF32 hi + lo and writes it into TENSOR
//     MEMORY (tcgen05.st; the A operand never touches shared memory); they also split the hidden units into the B operand
//     (K-major no-swizzle UMMA layout in shared memory, one extra row of ones -> Bsum);
//   * 1 MMA warp: three tcgen05.mma kind::tf32 (.ts form) per chunk -- hi*hi + hi*lo + lo*hi -- into the segment's
//     accumulator slot; tcgen05.commit frees the operand stage and publishes the finished accumulator;
//   * 8 contraction warps: tcgen05.ld the accumulators of G finished segments (G = 2 for vector roles, 4 for scalar roles) and
//     contract them with the resident weights in packed FFMA2 -- every weight read from shared memory is used for G segments
//     (and, in vector roles, for the three components that sit in neighbouring lanes: broadcast) --, sum over the rows of each
//     (class, component) by shuffles in a fixed order and write their share of the partial record.
// Measured pacing of tcgen05.mma kind::tf32 on B200 (tools/microbench/umma_pacing.cu): max(46, N / 2) cycles per instruction for
// M = 64 and 128 alike, so an M = 128 x N = 80 tile costs the same as any narrower one -- the reason every role keeps the
// widest N its weights allow and one tile.
// Results do not depend on the claiming order: every (segment, role) partial is produced by one fixed instruction sequence.
#include <cuda_pipeline_primitives.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ddk_tc.cuh"

namespace ddk {

constexpr int TR_NSETS = 3;          // row-warp sets: set s produces the operands of chunks it = s (mod 3) -- three chunks in flight
#ifndef DDK_TCR_NST
#define DDK_TCR_NST 3
#endif
constexpr int TR_NST = DDK_TCR_NST;  // operand stages (A in tensor memory, B in shared memory): chunk it -> stage it % TR_NST
static_assert(TR_NST >= TR_NSETS && TR_NST <= 6, "operand stages");
#ifndef DDK_TCR_XR
#define DDK_TCR_XR 9
#endif
#ifndef DDK_TCR_NGW
#define DDK_TCR_NGW 3
#endif
constexpr int TR_XR = DDK_TCR_XR;    // staging ring of the gather warps
constexpr int TR_NGW = DDK_TCR_NGW;  // gather warps: chunk it -> warp it % TR_NGW
// Phase-parity waits are only safe while no waiter can run two phases ahead of its barrier.  Ring slot it % TR_XR is filled by
// gather warp it % TR_NGW -- always the same warp, because TR_NGW divides TR_XR -- and read by row set it % 3, which differs from round to
// round: a set that waits for chunk it has stored chunk it - 3, so the MMA warp has consumed every chunk <= it - 3 - TR_NST, and
// the previous fill of the slot (chunk it - TR_XR) is among them iff TR_XR >= TR_NST + 3.  (Measured: 8-, 10-chunk rings with 3, 4
// or 5 stages and a third gather warp all run at the same speed within 1 %; a 6-stage / 8-chunk build violates the bound and hangs.)
static_assert(TR_NGW >= 1 && TR_NGW <= 3 && TR_XR % TR_NGW == 0 && (TR_XR % TR_NSETS == 0 || TR_XR >= TR_NST + 3),
              "staging ring: single filler per slot, readers at most one phase ahead");
constexpr int TR_ROWW = 4;           // row warps per set = one 128-row tile
constexpr int TR_CONW = 8;           // contraction warps
constexpr int TR_W_MMA = TR_NSETS * TR_ROWW, TR_W_GATHER = TR_W_MMA + 1, TR_W_CON = TR_W_MMA + 4;   // warps 12 | 13.. | 16..23
constexpr int TR_THREADS = (TR_W_CON + TR_CONW) * 32;   // 768: warps 0-11 rows, 12 MMA, 13-15 gather, 16-23 contraction (a multiple of 4: quarter = warp & 3)
constexpr int TR_COLS = 512;         // tensor-memory columns allocated
constexpr int TR_ACOL = 512 - 16 * TR_NST;        // A operand stages: TR_NST x (8 hi + 8 lo) columns from here; accumulator slots below
constexpr int TR_NMAX = 80;          // widest MMA N
constexpr int TR_GV = 2, TR_GS = 4;  // segments contracted together: vector roles / scalar roles
constexpr int TR_MAXSEG = 128;       // segments per task
constexpr int TR_PF = 4;             // chunks per batch of prefetched list entries (TR_PF * KC3 = one warp)
static_assert(TR_PF * KC3 == 32, "a batch of prefetched list entries is one warp wide");

struct TrArgs {
  int NL, N;
  int nb_segs;                       // segments per task (multiple of 8)
  int couple;                        // 1: a CTA claims for the role of its edge group that is furthest behind (see the claim code)
  int gmask;                         // bit g set: edge group g is processed
  const int4* glist; int goff[4]; int gci[4];
  const int* gcnt;                   // [list] segments; [2 F3_NLIST + 4 ... ] see launch_build_group_lists
  const int* gedges;                 // [list] listed edges of this step
  int* counters;                     // [4 * nroles] segment cursor of each combo
  const int2* seg_list;
  const float* x;                    // [N][84] layer input
  const float* hs;                   // [list position][72] hidden units of every listed edge (k_edge_hidden, edge-major)
  const float4* sh_pool;
  const TcrRole* roles; int nroles;  // roles of the level
  const float* W[4][TCR_MAXROLES];   // resident weight slice of (group, role)
  float* part;                       // [2 N][nroles][84]
  long long* dbg;                    // DDK_TCR_TRACE build: [grid][32] cycle counters
};

template <int LV>
struct TrCfg {
  static constexpr int U = AccCfg<LV>::U, DINP = AccCfg<LV>::DINP, XQ = DINP / 4;
  static constexpr int J = f3_J(LV), NSL = HID / J;
};

template <int LV>
struct TrSmem {
  alignas(128) uint32_t Bhi[TR_NST][TR_NMAX * 8];
  alignas(128) uint32_t Blo[TR_NST][TR_NMAX * 8];
  alignas(16) float X[TR_XR][KC3][TrCfg<LV>::DINP];                     // destination feature rows of the chunk's edges
  alignas(16) float SH[TR_XR][KC3][4];
  alignas(16) float HS[TR_XR][KC3][HID];                                // the role's hidden units of each edge (first nj floats)
  alignas(16) unsigned char role_[offsetof(TcrRole, fsrc)];             // the resident role's tables (all but fsrc, which only the finalize kernel reads)
  __device__ const TcrRole& role() const { return *reinterpret_cast<const TcrRole*>(role_); }
  alignas(8) unsigned long long full[TR_NST], empty[TR_NST];            // operand stages: row warps <-> MMA thread
  alignas(8) unsigned long long sfull[TR_XR], sempty[TR_XR];            // staging ring: gather warps <-> row warps
  alignas(8) unsigned long long accfull[TCR_MAXACC], accempty[TCR_MAXACC];   // accumulator slots: MMA thread <-> contraction warps
  alignas(8) unsigned long long bar_w;                                  // completion of the weight-slice bulk copy
  uint32_t tmem_base;
  int task[8];                       // g, role, idx0, nseg, reload, combo cursor, resident combo (g * nroles + role), -
  int seg_id[TR_MAXSEG], seg_n[TR_MAXSEG], seg_base[TR_MAXSEG];   // the task's work-list entries (record id, edges, first list position)
  alignas(128) float Wsl[1];         // the resident weight slice follows (TCR_WMAX floats, dynamic)
};

#if DDK_TCR_TRACE     // build with DDK_NVCC_EXTRA=-DDDK_TCR_TRACE=1 (tools/tcr_trace.sh): per CTA cycle counters of one thread per warp role
#define TR_T(var) const long long var = clock64();
#define TR_ADD(slot, a, b) dbgacc[slot] += (b) - (a);
#else
#define TR_T(var)
#define TR_ADD(slot, a, b)
#endif


typedef unsigned long long tr_f32x2;
__device__ __forceinline__ void tr_ffma2(tr_f32x2& d, const tr_f32x2 a, const tr_f32x2 b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
__device__ __forceinline__ tr_f32x2 tr_pack2(const float x, const float y) {
  tr_f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
  return r;
}
__device__ __forceinline__ void tr_unpack2(const tr_f32x2 v, float& x, float& y) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v));
}
__device__ __forceinline__ void tr_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
}

__device__ __forceinline__ void tr_ld2(uint32_t taddr, uint32_t (&v)[2]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(v[0]), "=r"(v[1]) : "r"(taddr));
}
__device__ __forceinline__ void tr_ld4(uint32_t taddr, uint32_t (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr));
}

// ---------------------------------------------------------------------------------------------- contraction warps
// One group of G segments (slots sg .. sg + G - 1; nvalid of them real).  VEC: lane = basis row (class, component c, row f of the
// class) with the three components in neighbouring lanes, so the 6 weights of (f, j) are one broadcast 8-byte load per pair;
// the two warp sets split the columns.  Scalar roles: lane = row, the two warp sets split the 24 outputs.
template <int LV, bool VEC>
__device__ __forceinline__ void tr_con_group(const TrArgs& p, TrSmem<LV>& S, const float* __restrict__ Wsl, const uint32_t tmem,
                                             const int sg, const int nvalid, const int seg0, const int g_edge, const int role_id,
                                             const int cw, const int q, const int lane, long long* dbgp = nullptr) {
  constexpr int G = VEC ? TR_GV : TR_GS, NACC = 2 * G;
  constexpr int O = VEC ? 6 : 24;                              // outputs per basis row; every thread computes 6 of them
  const TcrRole& R = S.role();
  const int N = R.N, ncol = R.ncol;
  const int set = cw >> 2;                                    // warp set 0 / 1
  // VEC: lane = basis row, the two warp sets split the columns.  Scalar roles: every row sits in lanes l and l + 16 of its
  // quarter (the accumulator rows are duplicated by the row warps), and (warp set, lane half) selects 6 of its 24 outputs.
  const bool active = R.rows[32 * q + lane].u >= 0;
  int c0 = 0, o0 = 0;
  if (VEC) c0 = set ? 40 : 0; else o0 = 12 * set + 6 * (lane >> 4);
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const int s = sg + g, slot = s % NACC;
    tc_mbar_wait_sleep(&S.accfull[slot], (s / NACC) & 1);
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#if DDK_TCR_TRACE
  const long long tp0 = clock64();
#endif
  tr_f32x2 acc[G][3];
#pragma unroll
  for (int g = 0; g < G; ++g)
#pragma unroll
    for (int k = 0; k < 3; ++k) acc[g][k] = 0ull;
  const float* wrow = Wsl + (active ? R.woff[32 * q + lane] : 0) + o0;
  const uint32_t tlane = tmem + ((uint32_t)(32 * q) << 16);
  constexpr int CW = VEC ? 8 : 4;                            // accumulator columns per load (4 segments x 4 in scalar roles)
  // The hot loop has no predicates: every block is full (accumulator columns >= ncol are exact zeros -- the B rows behind the
  // ones row are zeroed at every role change -- and the weight slice is padded with zeros behind its last block), and the
  // slots of absent segments (g >= nvalid, last group of a task) are read and multiplied like the others; their results are
  // never stored.  (With `if (g < nvalid)` / `if (col < c1)` inside, every FFMA2 came with two predicated moves.)
  const int nblk = VEC ? 5 : (ncol + CW - 1) / CW;           // vector roles: 40 columns per warp set
  uint32_t v[VEC ? 2 : 1][G][CW];                            // scalar roles (4 segments per group) have no registers for a second block
  auto issue = [&](int cb, uint32_t (&dst)[G][CW]) {
#pragma unroll
    for (int g = 0; g < G; ++g) {
      if constexpr (VEC) tr_ld8(tlane + ((sg + g) % NACC) * N + cb, dst[g]);
      else tr_ld4(tlane + ((sg + g) % NACC) * N + cb, dst[g]);
    }
  };
  auto compute = [&](int cb, const uint32_t (&src)[G][CW]) {
    const float* w = wrow + cb * O;
#pragma unroll
    for (int jj = 0; jj < CW; ++jj) {
      tr_f32x2 wv[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) wv[k] = *reinterpret_cast<const tr_f32x2*>(w + jj * O + 2 * k);
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float a = __uint_as_float(src[g][jj]);
        const tr_f32x2 aa = tr_pack2(a, a);
#pragma unroll
        for (int k = 0; k < 3; ++k) tr_ffma2(acc[g][k], aa, wv[k]);
      }
    }
  };
  if constexpr (VEC) {
    issue(c0, v[0]);
#pragma unroll 1
    for (int bk = 0; bk < nblk; bk += 2) {
      const int cb = c0 + bk * CW;
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");           // block bk has landed in v[0]
      if (bk + 1 < nblk) issue(cb + CW, v[1]);
      compute(cb, v[0]);
      if (bk + 1 < nblk) {
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");         // block bk + 1 has landed in v[1]
        if (bk + 2 < nblk) issue(cb + 2 * CW, v[0]);
        compute(cb + CW, v[1]);
      }
    }
  } else {
#pragma unroll 1
    for (int bk = 0; bk < nblk; ++bk) {
      issue(bk * CW, v[0]);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      compute(bk * CW, v[0]);
    }
  }
  // every accumulator value this thread needs is in registers: the slots may be refilled
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncwarp();
  if (lane == 0) {
#pragma unroll
    for (int g = 0; g < G; ++g) tc_mbar_arrive(&S.accempty[(sg + g) % NACC]);
  }
#if DDK_TCR_TRACE
  const long long tp1 = clock64();
#endif
  // ---- sum over the basis rows inside the warp, by shuffles in a fixed order; no other warp is involved: the warp writes
  // its own share of the (segment, role) partial record and k_conv_finalize_tcr adds the shares (TcrRole::fsrc).
  float r[G][6];
#pragma unroll
  for (int g = 0; g < G; ++g)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      tr_unpack2(acc[g][k], r[g][2 * k], r[g][2 * k + 1]);
      if (!active) { r[g][2 * k] = 0.f; r[g][2 * k + 1] = 0.f; }     // padding rows evaluate x[0] * sh[0] against block 0
    }
  if constexpr (VEC) {
    // the three components of a basis row sit in neighbouring lanes and every class starts at a warp boundary: the rows
    // that feed one (component, output) are the lanes of one residue mod 3 -> lanes 0, 1, 2 end with the warp's sums
#pragma unroll
    for (int delta = 24; delta >= 3; delta >>= 1) {
#pragma unroll
      for (int g = 0; g < G; ++g)
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          const float t = __shfl_down_sync(0xffffffffu, r[g][k], delta);
          if (lane + delta < 32) r[g][k] += t;
        }
    }
    if (lane < 3) {
#pragma unroll
      for (int g = 0; g < G; ++g)
        if (g < nvalid) {
          float* dst = p.part + ((size_t)S.seg_id[seg0 + g] * p.nroles + role_id) * TCR_PS + (cw * 3 + lane) * 6;
#pragma unroll
          for (int k = 0; k < 3; ++k) *reinterpret_cast<float2*>(dst + 2 * k) = make_float2(r[g][2 * k], r[g][2 * k + 1]);
        }
    }
  } else {
    // lane = 16 h + l: row 16 q + l, outputs 12 set + 6 h + (0..5); the 16 rows of a quarter belong to one class
#pragma unroll
    for (int delta = 8; delta >= 1; delta >>= 1) {
#pragma unroll
      for (int g = 0; g < G; ++g)
#pragma unroll
        for (int k = 0; k < 6; ++k) r[g][k] += __shfl_xor_sync(0xffffffffu, r[g][k], delta);
    }
    if ((lane & 15) == 0) {
#pragma unroll
      for (int g = 0; g < G; ++g)
        if (g < nvalid) {
          float* dst = p.part + ((size_t)S.seg_id[seg0 + g] * p.nroles + role_id) * TCR_PS + (cw * 2 + (lane >> 4)) * 6;
#pragma unroll
          for (int k = 0; k < 3; ++k) *reinterpret_cast<float2*>(dst + 2 * k) = make_float2(r[g][2 * k], r[g][2 * k + 1]);
        }
    }
  }
#if DDK_TCR_TRACE
  if (dbgp) { const int k0 = VEC ? 3 : 5; dbgp[k0] += tp1 - tp0; dbgp[k0 + 1] += clock64() - tp1; dbgp[7] += VEC ? 1 : 0; }
#endif
}

// ---------------------------------------------------------------------------------------------- the kernel
template <int LV>
__global__ void __launch_bounds__(TR_THREADS, 1) k_conv_tcr(const __grid_constant__ TrArgs p) {
  using Cfg = TrCfg<LV>;
  constexpr int DINP = Cfg::DINP;
  extern __shared__ __align__(128) unsigned char tr_raw[];
  TrSmem<LV>& S = *reinterpret_cast<TrSmem<LV>*>(tr_raw);
  float* const Wsl = &S.Wsl[0];
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);     // warp-uniform by construction: role branches stay uniform
  const int ncombo = 4 * p.nroles;

  // ---- one-time setup
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem(&S.tmem_base)), "n"(TR_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == TR_W_MMA * 32) {
    for (int s = 0; s < TR_NST; ++s) { tc_mbar_init(&S.full[s], TR_ROWW); tc_mbar_init(&S.empty[s], 1); }
    for (int s = 0; s < TR_XR; ++s) { tc_mbar_init(&S.sfull[s], 32); tc_mbar_init(&S.sempty[s], TR_ROWW); }
    for (int s = 0; s < TCR_MAXACC; ++s) { tc_mbar_init(&S.accfull[s], 1); tc_mbar_init(&S.accempty[s], TR_CONW); }
    tc_mbar_init(&S.bar_w, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // first combo of this CTA: the combos share the CTAs in proportion to a rough cost model (per listed edge and per segment,
    // vector roles about twice the scalar ones), so few CTAs have to migrate -- and reload a weight slice -- before the tail
    auto combo_cost = [&](int cmb) {
      const int g = cmb / p.nroles, r = cmb % p.nroles;
      if (!((p.gmask >> g) & 1)) return 0.f;
      const float ne = (float)p.gedges[p.gci[g]], ns = (float)p.gcnt[p.gci[g]];
      return p.roles[r].isS ? ne * 10.f + ns * 400.f : ne * 18.f + ns * 500.f;
    };
    float total = 0.f;
    for (int cmb = 0; cmb < ncombo; ++cmb) total += combo_cost(cmb);
    const float pos = ((float)blockIdx.x + 0.5f) / (float)gridDim.x * total;
    int first = 0;
    float run = 0.f;
    for (int cmb = 0; cmb < ncombo; ++cmb) { run += combo_cost(cmb); first = cmb; if (run > pos) break; }
    S.task[5] = first; S.task[6] = -1;
  }
  {
    // operand tiles and staging rings start as zeros: the bulk copies of a partial chunk leave the rows of absent edges as they
    // are (their hidden units are masked to 0 in the B operand, so whatever FINITE values they hold contribute nothing)
    uint32_t* z = &S.Bhi[0][0];
    constexpr int nz = (int)((offsetof(TrSmem<LV>, role_) - offsetof(TrSmem<LV>, Bhi)) / 4);
    for (int i = tid; i < nz; i += TR_THREADS) z[i] = 0u;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = S.tmem_base;

#if DDK_TCR_TRACE
  long long dbgacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long tk0 = clock64();
  long long t_reload = 0, n_tasks = 0, n_reload = 0;
#endif
  // Chunk bookkeeping.  Every warp role walks the chunks of the CTA's tasks in the same order; chunk number `it` (never
  // materialised) belongs to row set it % 3, gather warp it % TR_NGW, operand stage it % TR_NST and ring slot it % TR_XR.  A row set / gather
  // warp steps straight from one of its chunks to the next (`skip` = offset of its first chunk in the next task) and keeps the
  // ring slot, ring phase and stage phase of that chunk incrementally; the MMA warp visits every chunk.
  int skip = 0, buf = 0, bph = 0, rst = 0, sph = 0;
  if (warp < TR_W_MMA) { skip = warp >> 2; buf = skip % TR_XR; rst = skip % TR_NST; }
  else if (warp >= TR_W_GATHER && warp < TR_W_GATHER + TR_NGW) { skip = warp - TR_W_GATHER; buf = skip; }
  int mstage = 0, mph = 0;   // MMA warp: operand stage and its phase
  int sg = 0;        // accumulator slots so far (MMA thread / contraction warps); a multiple of the group size between tasks
  int nwl = 0;       // weight-slice loads so far

  for (;;) {
    // ---- claim a task: (combo, block of segments)
    // couple = 1: inside its edge group the CTA takes the next block of the ROLE WHOSE CURSOR IS FURTHEST BEHIND (the resident one on
    // ties).  The roles of a group then walk the work list side by side -- the second to fifth visit of an edge find its hidden
    // units, list entries and harmonics in L2 instead of DRAM -- and finish together whatever their relative cost; the price is a
    // weight-slice reload (one bulk copy, ~2 k cycles) at most tasks.  couple = 0: stay on the resident combo while it has work.
    if (tid == 0) {
      int found = 0;
      if (p.couple) {
        int g = S.task[5] / p.nroles;
        const int rres = S.task[6] >= 0 && S.task[6] / p.nroles == g ? S.task[6] % p.nroles : -1;
        for (int gt = 0; gt < 4 && !found; ++gt, g = (g + 1) & 3) {
          const int gall = ((p.gmask >> g) & 1) ? p.gcnt[p.gci[g]] : 0;
          for (int retry = 0; retry <= p.nroles && !found; ++retry) {
            int best = -1, bcur = 0x7fffffff;
            for (int r = 0; r < p.nroles; ++r) {
              const int cur = *reinterpret_cast<volatile int*>(p.counters + g * p.nroles + r);
              if (cur < gall && (cur < bcur || (cur == bcur && r == rres))) { best = r; bcur = cur; }
            }
            if (best < 0) break;                                                // the group is exhausted
            const int combo = g * p.nroles + best, rem = gall - bcur;
            const int size = max(8, min(p.nb_segs, (rem / 8) / 8 * 8));         // guided self-scheduling
            const int start = atomicAdd(p.counters + combo, size);
            if (start < gall) {
              S.task[0] = g; S.task[1] = best; S.task[2] = start; S.task[3] = min(size, gall - start);
              S.task[4] = (combo != S.task[6]);
              S.task[5] = combo; S.task[6] = combo;
              found = 1;
            }
          }
        }
      } else {
        int combo = S.task[5];
        for (int tries = 0; tries < ncombo && !found; ++tries) {
          const int g = combo / p.nroles;
          const int gall = ((p.gmask >> g) & 1) ? p.gcnt[p.gci[g]] : 0;
          const int rem = gall - *reinterpret_cast<volatile int*>(p.counters + combo);
          if (rem > 0) {
            const int size = max(8, min(p.nb_segs, (rem / 8) / 8 * 8));         // guided self-scheduling
            const int start = atomicAdd(p.counters + combo, size);
            if (start < gall) {
              S.task[0] = g; S.task[1] = combo % p.nroles; S.task[2] = start; S.task[3] = min(size, gall - start);
              S.task[4] = (combo != S.task[6]);
              S.task[5] = combo; S.task[6] = combo;
              found = 1;
              break;
            }
          }
          combo = (combo + 1) % ncombo;
        }
      }
      if (!found) S.task[0] = -1;
    }
    __syncthreads();
    const int g = S.task[0];
    if (g < 0) break;
    const int role_id = S.task[1], idx0 = S.task[2], nseg = S.task[3];
#if DDK_TCR_TRACE
    ++n_tasks;
    const long long trl0 = clock64();
#endif
    if (S.task[4]) {
      // every read of the previous slice / role tables happened before the __syncthreads that ended the previous task
      const TcrRole* src = p.roles + role_id;
      if (tid == 0) {
        const uint32_t bytes = (uint32_t)src->wfloats * 4u;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_mbar_expect_tx(&S.bar_w, bytes);
        tc_bulk_g2s(Wsl, p.W[g][role_id], bytes, &S.bar_w);
      }
      {
        const int4* s4 = reinterpret_cast<const int4*>(src);
        int4* d4 = reinterpret_cast<int4*>(S.role_);
        static_assert(sizeof(TcrRole) % 16 == 0 && offsetof(TcrRole, fsrc) % 16 == 0 && TCR_MAXSRC == 8, "role tables are copied in 16-byte pieces");
        for (int i = tid; i < (int)(offsetof(TcrRole, fsrc) / 16); i += TR_THREADS) d4[i] = __ldg(s4 + i);
      }
      // the accumulator-slot ring depends on the role kind (4 slots of 80 columns / 8 slots of 32 or 48): every pipeline is
      // drained here, so the slot barriers restart from phase 0 together with the slot counter
      if (tid == TR_W_MMA * 32) {
        for (int sI = 0; sI < TCR_MAXACC; ++sI) {
          asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(tc_smem(&S.accfull[sI])) : "memory");
          asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(tc_smem(&S.accempty[sI])) : "memory");
          tc_mbar_init(&S.accfull[sI], 1); tc_mbar_init(&S.accempty[sI], TR_CONW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      }
      sg = 0;
      {
        // B rows [ncol, N) of every stage are exact zeros for this role (a previous role with another tile width may have left
        // hidden units there): the accumulator columns behind the ones column are then zeros, which the contraction relies on
        const int ncol_r = src->ncol, N_r = src->N;
        const int nz = (N_r - ncol_r) * 2 * TR_NST * 2;          // rows x k-halves x stages x (hi, lo)
        for (int i = tid; i < nz * 4; i += TR_THREADS) {
          const int wd = i & 3; int r = i >> 2;
          const int row = ncol_r + r % (N_r - ncol_r); r /= (N_r - ncol_r);
          const int hk = r & 1; r >>= 1;
          const int stg = r % TR_NST, hl = r / TR_NST;
          (hl ? S.Blo : S.Bhi)[stg][(row + N_r * hk) * 4 + wd] = 0u;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      }
      tc_mbar_wait_sleep(&S.bar_w, nwl & 1);
      ++nwl;
      __syncthreads();
#if DDK_TCR_TRACE
      ++n_reload; t_reload += clock64() - trl0;
#endif
    }
    {
      // the task's work-list entries -> shared memory: every warp role walks them, nobody waits on a global load per segment
      const int4* wle = p.glist + p.goff[g] + idx0;
      for (int i = tid; i < nseg; i += TR_THREADS) {
        const int4 e4 = __ldg(wle + i);                  // (segment, edges, first list position, record id)
        S.seg_id[i] = e4.w; S.seg_n[i] = e4.y; S.seg_base[i] = e4.z;
      }
    }
    __syncthreads();
    const TcrRole& R = S.role();
    const int N = R.N, nj = R.nj, j0 = R.j0;
    const bool vec = !R.isS;
    const int G = vec ? TR_GV : TR_GS;
    [[maybe_unused]] const int NACC = 2 * G;        // (trace build)

    if (warp < TR_W_MMA) {
      // ================================================================== row warps: A operand -> tensor memory, B -> smem
      // set = warp / 4 produces the chunks it = set (mod TR_NSETS) into operand stage `set`; thread (quarter, lane) = tile row
      const int rt = (warp & 3) * 32 + lane;
      const TcRow rd = R.rows[rt];
      const bool plain = __all_sync(0xffffffffu, rd.type == 0);
      // every basis row is  x[ia] sh[ma] + sb x[ib] sh[mb] + sc x[ic] sh[mc]:  plain product (sb = sc = 0), dot product of a
      // 3-vector with the l = 1 harmonics (sb = sc = 1), or one component of their cross product (sb = -1, sc = 0)
      int ia = rd.i0, ma = rd.m, ib = rd.i0, mb = 0, ic = rd.i0, mc = 0;
      float sb = 0.f, sc = 0.f;
      if (rd.type == 1) { ma = 1; ib = rd.i0 + 1; mb = 2; ic = rd.i0 + 2; mc = 3; sb = 1.f; sc = 1.f; }
      if (rd.type == 2) {
        const int a = rd.m % 3, c2 = (rd.m + 1) % 3;          // component m - 1 = x[a'] s[b'] - x[b'] s[a'] with (a', b') cyclic
        ia = rd.i0 + a; ma = 1 + c2; ib = rd.i0 + c2; mb = 1 + a; sb = -1.f;
      }
      // B operand work items of this thread: item w = (row n = w % (nj + 1) of the tile, half hk = w / (nj + 1) of the 8 edges)
      const int nb_items = 2 * (nj + 1);
      int bn[2], bh[2];
#pragma unroll
      for (int b2 = 0; b2 < 2; ++b2) {
        const int w = rt + 128 * b2;
        bn[b2] = w < nb_items ? w % (nj + 1) : -1;
        bh[b2] = w < nb_items ? w / (nj + 1) : 0;
      }
      int i = 0, c = skip, nch = (S.seg_n[0] + KC3 - 1) / KC3;
      auto norm = [&]() {
        while (i < nseg && c >= nch) { c -= nch; ++i; nch = i < nseg ? (S.seg_n[i] + KC3 - 1) / KC3 : 0; }
      };
      norm();
      while (i < nseg) {
        {
          const int buf_c = buf, rp_c = bph, stage = rst, sp_c = sph;
          const int kc = min(KC3, S.seg_n[i] - c * KC3);
          TR_T(ra)
          tc_mbar_wait(&S.sfull[buf_c], rp_c);
          TR_T(rb)
          // everything that only READS the staged chunk happens before the wait for the operand stage: the basis values, the B
          // rows of this thread and their TF32 splits sit in registers when the MMA warp releases the stage, and only the stores
          // remain on the stage's critical loop (release -> stores -> fences -> MMA of this chunk)
          float b[KC3];
          if (plain) {
            float xv[KC3], sv[KC3];
#pragma unroll
            for (int e = 0; e < KC3; ++e) { xv[e] = S.X[buf_c][e][rd.i0]; sv[e] = S.SH[buf_c][e][rd.m]; }
#pragma unroll
            for (int e = 0; e < KC3; ++e) b[e] = xv[e] * sv[e];
          } else {
#pragma unroll
            for (int e = 0; e < KC3; ++e) {
              const float* xr = &S.X[buf_c][e][0];
              const float* sr = &S.SH[buf_c][e][0];
              float r = xr[ia] * sr[ma];
              r = fmaf(sb * xr[ib], sr[mb], r);
              b[e] = fmaf(sc * xr[ic], sr[mc], r);
            }
          }
          uint32_t hi[KC3], lo[KC3];
#pragma unroll
          for (int e = 0; e < KC3; ++e) tc_split_rn(b[e], hi[e], lo[e]);
          uint32_t bh_[2][4], bl_[2][4];
#pragma unroll
          for (int b2 = 0; b2 < 2; ++b2) {
            const int nrow = bn[b2], hk = bh[b2];
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) { bh_[b2][qd] = 4 * hk + qd < kc ? 0x3f800000u : 0u; bl_[b2][qd] = 0u; }   // the ones row -> Bsum
            if (nrow >= 0 && nrow < nj) {
#pragma unroll
              for (int qd = 0; qd < 4; ++qd) {
                const float hld = S.HS[buf_c][4 * hk + qd][nrow];
                tc_split_rn(4 * hk + qd < kc ? hld : 0.f, bh_[b2][qd], bl_[b2][qd]);   // rows of absent edges hold stale (finite) data
              }
            }
          }
          __syncwarp();
          if (lane == 0) tc_mbar_arrive(&S.sempty[buf_c]);      // the ring slot is free again
          TR_T(rb2)
          tc_mbar_wait(&S.empty[stage], sp_c ^ 1);
          TR_T(rc)
          TR_ADD(0, ra, rb) TR_ADD(1, rb2, rc)
          {
            const uint32_t ta = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + TR_ACOL + stage * 16;
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                         ::"r"(ta), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]), "r"(hi[4]), "r"(hi[5]), "r"(hi[6]), "r"(hi[7]) : "memory");
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                         ::"r"(ta + 8), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]), "r"(lo[4]), "r"(lo[5]), "r"(lo[6]), "r"(lo[7]) : "memory");
          }
#pragma unroll
          for (int b2 = 0; b2 < 2; ++b2) {
            const int nrow = bn[b2], hk = bh[b2];
            if (nrow >= 0) {
              *reinterpret_cast<uint4*>(&S.Bhi[stage][(nrow + N * hk) * 4]) = make_uint4(bh_[b2][0], bh_[b2][1], bh_[b2][2], bh_[b2][3]);
              *reinterpret_cast<uint4*>(&S.Blo[stage][(nrow + N * hk) * 4]) = make_uint4(bl_[b2][0], bl_[b2][1], bl_[b2][2], bl_[b2][3]);
            }
          }
          TR_T(rd_)
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) tc_mbar_arrive(&S.full[stage]);
          TR_T(re)
          TR_ADD(2, rb, rb2) TR_ADD(2, rc, rd_) TR_ADD(3, rd_, re)
#if DDK_TCR_TRACE
          dbgacc[4] += 1;
          if (vec) { dbgacc[5] += 1; dbgacc[6] += (rb2 - rb) + (rd_ - rc); }
#endif
        }
        buf += TR_NSETS; if (buf >= TR_XR) { buf -= TR_XR; bph ^= 1; }
        rst += TR_NSETS; if (rst >= TR_NST) { rst -= TR_NST; sph ^= 1; }
        c += TR_NSETS;
        norm();
      }
      skip = c;
    } else if (warp == TR_W_MMA) {
      // ================================================================== MMA issue
      // The whole warp walks the chunks with warp-uniform values (shuffled from lane 0), and ONE elected lane issues the
      // tcgen05 instructions: their operands then live in uniform registers.  (Issued from inside `if (lane == 0)` every
      // UTCHMMA / UTCBAR was wrapped in an ELECT + 4 x R2UR + branch loop: ~390 cycles per chunk in the trace.)
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
      const int N_u = __shfl_sync(0xffffffffu, N, 0), G_u = __shfl_sync(0xffffffffu, G, 0), NACC_u = 2 * G_u;
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N_u >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t bhi0 = tc_smem(&S.Bhi[0][0]), blo0 = tc_smem(&S.Blo[0][0]);
      auto elect = []() {
        uint32_t pred;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
        return pred != 0;
      };
      // descriptors of operand stage 0; stage s is TR_NMAX * 32 bytes further (the start-address field counts 16-byte units)
      const uint64_t bhd0 = tc_desc(bhi0, N_u * 16, 128), bld0 = tc_desc(blo0, N_u * 16, 128);
      constexpr uint32_t dstep = (TR_NMAX * 8 * 4) >> 4;
      for (int i = 0; i < nseg; ++i, ++sg) {
        const int nch = (__shfl_sync(0xffffffffu, S.seg_n[i], 0) + KC3 - 1) / KC3;
        const int slot = sg % NACC_u;
        TR_T(ma)
        tc_mbar_wait(&S.accempty[slot], ((sg / NACC_u) & 1) ^ 1);                  // the contraction warps have read the slot's old content
        TR_T(mb)
        TR_ADD(0, ma, mb)
        const uint32_t d = tmem_u + slot * N_u;
        for (int c = 0; c < nch; ++c) {
          const int stage = mstage;
          TR_T(mc)
          tc_mbar_wait(&S.full[stage], mph);
          TR_T(md)
          TR_ADD(1, mc, md)
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t bhd = bhd0 + stage * dstep, bld = bld0 + stage * dstep;
          const uint32_t ah = tmem_u + TR_ACOL + stage * 16, al = ah + 8;
          if (elect()) {
            tc_mma_ts(d, ah, bhd, idesc, c > 0);
            tc_mma_ts(d, ah, bld, idesc, 1);
            tc_mma_ts(d, al, bhd, idesc, 1);
            tc_commit(&S.empty[stage]);
            if (c == nch - 1) tc_commit(&S.accfull[slot]);
          }
          __syncwarp();
          if (++mstage == TR_NST) { mstage = 0; mph ^= 1; }
          TR_T(me)
          TR_ADD(2, md, me)
        }
      }
      for (; sg % G_u != 0; ++sg) {                                                // pad the last group of the task with empty slots
        const int slot = sg % NACC_u;
        tc_mbar_wait(&S.accempty[slot], ((sg / NACC_u) & 1) ^ 1);
        if (lane == 0) tc_mbar_arrive(&S.accfull[slot]);
        __syncwarp();
      }
    } else if (warp >= TR_W_GATHER && warp < TR_W_GATHER + TR_NGW) {
      // ================================================================== gather warp: global -> staging ring
      // Levels 2, 3 (feature rows of 240 / 336 bytes): two passes of four edges, lane = (edge 4 * pass + lane / 8, sub = lane % 8); the
      // eight lanes of an edge copy its record in interleaved 16-byte pieces -- destination feature row, harmonics, the role's hidden units -- with cp.async (LDGSTS; bulk
      // copies are per-warp instructions and serialise when every lane has its own address: measured 1700 cycles per chunk).
      // A quarter-warp (the unit a 16-byte shared-memory store is processed in) then writes 128 CONTIGUOUS bytes of one row:
      // conflict-free whatever the row stride.  (With four lanes per edge a quarter-warp wrote 64 bytes of two rows 336 / 288 bytes
      // apart, which overlap modulo 128: 13.4 shared-memory wavefronts per LDGSTS instead of 4, a third of all wavefronts of the
      // kernel, whose shared-memory pipe is its busiest unit.)  Levels 0, 1 (96 / 176-byte rows): one pass, four lanes per edge.
      // Completion is signalled to the ring slot's mbarrier by
      // cp.async.mbarrier.arrive.noinc, so the warp never waits for its own copies.
      constexpr int XQ = DINP / 4;
      constexpr int LPE = XQ >= 12 ? 8 : 4, EPP = 32 / LPE;   // lanes per edge record (narrow rows: 4), edges per pass
      const int eq = lane / LPE, sub = lane % LPE;
      const int hq = nj / 4;
      // list entries (edge slot, destination node) are fetched a BATCH of TR_PF chunks at a time, one batch ahead: lane l holds the
      // entry of edge l % 8 of the batch's chunk l / 8 (one coalesced load), and the chunk's lanes get theirs by shuffle.  The
      // registers of a batch are touched again only TR_PF chunks after its load was issued -- a per-chunk register FIFO made
      // every chunk wait for the load issued one chunk earlier (the rotating moves read the newest entry).
      // The two gather warps take alternate chunks (gw = parity of the chunk counter); each prefetches the entries of ITS chunks.
      int i = 0, c = skip, nch = (S.seg_n[0] + KC3 - 1) / KC3;
      auto norm = [&](int& i_, int& c_, int& nch_) {
        while (i_ < nseg && c_ >= nch_) { c_ -= nch_; ++i_; nch_ = i_ < nseg ? (S.seg_n[i_] + KC3 - 1) / KC3 : 0; }
      };
      norm(i, c, nch);
      int hi_ = i, hc = c, hnch = nch;                  // head of the prefetch stream: this warp's next chunk not yet fetched
      auto batch_load = [&]() {
        int mypos = -1, myrem = 1;
#pragma unroll
        for (int b = 0; b < TR_PF; ++b) {
          if (hi_ < nseg) {
            if ((lane >> 3) == b) { mypos = S.seg_base[hi_] + hc * KC3; myrem = S.seg_n[hi_] - hc * KC3; }
            hc += TR_NGW;
            norm(hi_, hc, hnch);
          }
        }
        int2 v = make_int2(0, 0);
        if (mypos >= 0) v = __ldg(&p.seg_list[mypos + min(lane & 7, myrem - 1)]);
        return v;
      };
      int2 cur = batch_load(), nxt = batch_load();
      int kb = 0;                                       // chunk of the current batch
      while (i < nseg) {
        {
          const int n = S.seg_n[i];
          const int kc = min(KC3, n - c * KC3), pos0 = S.seg_base[i] + c * KC3;
          const int2 ent = cur;
          const int kb_c = kb;
          if (++kb == TR_PF) { kb = 0; cur = nxt; nxt = batch_load(); }
          TR_T(ga)
          tc_mbar_wait(&S.sempty[buf], bph ^ 1);
          TR_T(gb)
          TR_ADD(0, ga, gb)
#pragma unroll
          for (int ps = 0; ps < KC3 / EPP; ++ps) {
            const int e = EPP * ps + eq;
            const int slot = __shfl_sync(0xffffffffu, ent.x, kb_c * KC3 + e), dst = __shfl_sync(0xffffffffu, ent.y, kb_c * KC3 + e);
            if (e < kc) {                               // rows of absent edges keep their stale (finite) content
              const float* xs = p.x + (size_t)dst * D;
              float* xd = &S.X[buf][e][0];
#pragma unroll
              for (int k = 0; k < (XQ + LPE - 1) / LPE; ++k)
                if (sub + LPE * k < XQ) __pipeline_memcpy_async(xd + 4 * (sub + LPE * k), xs + 4 * (sub + LPE * k), 16);
              if (sub == LPE - 1) __pipeline_memcpy_async(&S.SH[buf][e][0], p.sh_pool + slot, 16);
              const float* hsrc = p.hs + (size_t)(pos0 + e) * HID + j0;
              float* hd = &S.HS[buf][e][0];
#pragma unroll
              for (int k = 0; k < (HID / 4 + LPE - 1) / LPE; ++k)
                if (sub + LPE * k < hq) __pipeline_memcpy_async(hd + 4 * (sub + LPE * k), hsrc + 4 * (sub + LPE * k), 16);
            }
          }
          // the barrier receives this thread's arrival when all of its copies above have landed
          asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(tc_smem(&S.sfull[buf])) : "memory");
          TR_T(gc)
          TR_ADD(1, gb, gc)
        }
        buf += TR_NGW; if (buf >= TR_XR) { buf -= TR_XR; bph ^= 1; }
        c += TR_NGW;
        norm(i, c, nch);
      }
      skip = c;
      asm volatile("cp.async.wait_all;" ::: "memory");
    } else if (warp >= TR_W_CON) {
      // ================================================================== contraction warps
      const int cw = warp - TR_W_CON, q = warp & 3;    // TR_W_CON is a multiple of 4: q = cw & 3
      const int ngroups = (nseg + G - 1) / G;
      for (int grp = 0; grp < ngroups; ++grp, sg += G) {
        const int nvalid = min(G, nseg - grp * G);
#if DDK_TCR_TRACE
        {                                             // time spent waiting for the group's accumulators (then the group itself)
          const long long ca = clock64();
          for (int g2 = 0; g2 < G; ++g2) tc_mbar_wait_sleep(&S.accfull[(sg + g2) % NACC], ((sg + g2) / NACC) & 1);
          dbgacc[0] += clock64() - ca;
        }
        const long long cb_ = clock64();
#endif
#if DDK_TCR_TRACE
        if (vec) tr_con_group<LV, true>(p, S, Wsl, tmem, sg, nvalid, grp * G, 0, role_id, cw, q, lane, dbgacc);
        else tr_con_group<LV, false>(p, S, Wsl, tmem, sg, nvalid, grp * G, 0, role_id, cw, q, lane, dbgacc);
#else
        if (vec) tr_con_group<LV, true>(p, S, Wsl, tmem, sg, nvalid, grp * G, 0, role_id, cw, q, lane);
        else tr_con_group<LV, false>(p, S, Wsl, tmem, sg, nvalid, grp * G, 0, role_id, cw, q, lane);
#endif
#if DDK_TCR_TRACE
        dbgacc[1] += clock64() - cb_; dbgacc[2] += 1;
#endif
      }
    }
    // every thread keeps only the counters its own warp role uses and advances them by walking the same task sequence, so the
    // warp roles agree on chunk and slot numbers without any exchange
    __syncthreads();
  }
#if DDK_TCR_TRACE
  if (p.dbg) {
    long long* o = p.dbg + (size_t)blockIdx.x * 32;
    if (tid == 0) { for (int k = 0; k < 5; ++k) o[k] = dbgacc[k]; o[5] = clock64() - tk0; o[6] = n_tasks; o[7] = n_reload; o[8] = t_reload; o[26] = dbgacc[5]; o[27] = dbgacc[6]; }
    if (tid == TR_W_MMA * 32) for (int k = 0; k < 3; ++k) o[10 + k] = dbgacc[k];
    if (tid == TR_W_GATHER * 32) for (int k = 0; k < 2; ++k) o[14 + k] = dbgacc[k];
    if (tid == TR_W_CON * 32) for (int k = 0; k < 8; ++k) o[18 + k] = dbgacc[k];
  }
#endif
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TR_COLS));
}

// ---------------------------------------------------------------------------------------------- finalize
// Sum of the partial records of a node's two segments over the roles (a fixed order: segment, role, share), mean over the
// node's edges, batch-norm affine, residual -- k_conv_finalize's arithmetic on k_conv_tcr's record layout.
struct FinTcrArgs {
  int N, dout, nroles;
  int NL, sub, extra;                // ligand nodes; edges per piece of a lig<-rec segment (0: unsplit); first record id of the pieces p > 0
  const int* seg_cnt; const float* part; const TcrRole* roles;
  const float* bn_scale; const float* bn_shift; const float* x_in; float* x_out;
};

__global__ void __launch_bounds__(256) k_conv_finalize_tcr(FinTcrArgs p) {
  const int q = threadIdx.x / D, f = threadIdx.x % D;
  if (q >= 3) return;
  const int node = blockIdx.x * 3 + q;
  if (node >= p.N) return;
  const int cnt[2] = {p.seg_cnt[2 * node], p.seg_cnt[2 * node + 1]};
  float v = 0.f;
  if (f < p.dout) {
    float s = 0.f;
    // the lig<-rec segment of a ligand node was accumulated in pieces (TCR_SUB): piece 0 in the segment's record, the others behind
    const int npiece = node < p.NL && p.sub > 0 && cnt[1] > 0 ? min(TCR_PMAX, (cnt[1] + p.sub - 1) / p.sub) : 1;
    for (int r = 0; r < p.nroles; ++r) {
      const int4 t4 = __ldg(reinterpret_cast<const int4*>(&p.roles[r].fsrc[f][0]));
      const int tw[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (cnt[h] <= 0) continue;
        const int np = h == 1 ? npiece : 1;
        for (int pc = 0; pc < np; ++pc) {
          const size_t rid = pc == 0 ? (size_t)(2 * node + h) : (size_t)p.extra + (size_t)(TCR_PMAX - 1) * node + pc - 1;
          const float* rec = p.part + (rid * p.nroles + r) * TCR_PS;
#pragma unroll
          for (int i = 0; i < TCR_MAXSRC; ++i) {
            const int o = (short)((tw[i >> 1] >> (16 * (i & 1))) & 0xffff);
            if (o >= 0) s += rec[o];
          }
        }
      }
    }
    const float cn = fmaxf((float)(cnt[0] + cnt[1]), 1.f);
    v = (s / cn) * p.bn_scale[f] + p.bn_shift[f] + p.x_in[(size_t)node * D + f];
  }
  p.x_out[(size_t)node * D + f] = v;
}

// ---------------------------------------------------------------------------------------------- host side
static int tcr_ncols16(int n) { return (n + 15) / 16 * 16; }

// Roles of basis level lv (see the header).  cls: the irrep classes of a layer of that level (build_layers).
int build_tcr_roles(int lv, const LayerInfo& li, TcrRole* roles) {
  std::vector<TcRow> all(TC_MAXROWS);
  build_tc_rows(lv, all.data());                      // (type, i0, m, u) of every basis row of the level
  std::vector<TcRow> byu(li.U);
  for (int p = 0; p < TC_MAXROWS; ++p) if (all[p].u >= 0) byu[all[p].u] = all[p];
  struct Spec { std::vector<int> cls; int j0, nj; bool isS; };
  std::vector<Spec> specs;
  std::vector<int> vcls, scls;
  for (int k = 0; k < li.ncls; ++k) (li.cls[k].ncomp == 3 ? vcls : scls).push_back(k);
  // vector roles: as many classes per role as fit one tile
  {
    std::vector<int> cur; int rows = 0;
    for (int k : vcls) {
      const int r = 3 * li.cls[k].F;
      if (!cur.empty() && rows + r > 128) { specs.push_back({cur, 0, HID, false}); cur.clear(); rows = 0; }
      cur.push_back(k); rows += r;
    }
    if (!cur.empty()) specs.push_back({cur, 0, HID, false});
  }
  // scalar roles: all scalar classes, the hidden units in halves where the weight slice fits (levels 0 - 2), else in thirds
  if (!scls.empty()) {
    const int parts = lv == 3 ? 3 : 2;
    for (int h = 0; h < parts; ++h) specs.push_back({scls, h * (HID / parts), HID / parts, true});
  }
  if ((int)specs.size() > TCR_MAXROLES) return -1;
  for (size_t r = 0; r < specs.size(); ++r) {
    const Spec& sp = specs[r];
    TcrRole& R = roles[r];
    memset(&R, 0, sizeof(R));
    R.isS = sp.isS; R.j0 = sp.j0; R.nj = sp.nj; R.ncol = sp.nj + 1; R.N = tcr_ncols16(R.ncol);
    if ((sp.j0 * 4) % 16 || (sp.nj * 4) % 16) return -1;   // bulk copies of the role's hidden units
    R.O = sp.isS ? 24 : 6;
    // weight block of one (class, f): ncol x O floats; the stride between blocks is = 2 (mod 4) floats, i.e. an odd number of
    // 8-byte words, so that the 8-byte loads of lanes working on different blocks fall into different banks
    const int blk = R.ncol * R.O;
    R.wstride = blk % 4 == 2 ? blk : blk + 2;
    // distinct rows: class-major (vector classes start at a multiple of 32 rows, scalar classes at a multiple of 16), inside a
    // class type-sorted (plain products first) with the three components of (class, f) in neighbouring rows
    // Vector roles, if the tile has room: the rows that are not plain products (cross products) start at a warp boundary of their
    // own, so that only their warps take the three-term evaluation path (48 instead of 16 shared-memory loads per chunk).
    const int align = sp.isS ? 16 : 32;
    int nd = 0, wblocks = 0;
    std::vector<TcRow> drow;
    std::vector<int> dwoff, dcls, dcomp;
    auto place = [&](bool split_types) {
      nd = 0; wblocks = 0;
      drow.assign(128, TcRow{0, 0, 0, -1}); dwoff.assign(128, 0); dcls.assign(128, -1); dcomp.assign(128, 0);
      for (int k : sp.cls) {
        const ClassInfo& ci = li.cls[k];
        if (ci.O != R.O) return false;
        nd = (nd + align - 1) / align * align;
        std::vector<int> blk_of(ci.F);
        for (int f = 0; f < ci.F; ++f) blk_of[f] = wblocks++;
        for (int pass = 0; pass < 3; ++pass) {
          bool first = true;
          for (int f = 0; f < ci.F; ++f) {
            if (byu[ci.uoff + f].type != pass) continue;
            if (first && pass > 0 && split_types && !sp.isS) nd = (nd + 31) / 32 * 32;
            first = false;
            for (int c = 0; c < ci.ncomp; ++c) {
              if (nd >= (sp.isS ? 64 : 128)) return false;
              const int u = ci.uoff + c * ci.F + f;
              if (byu[u].type != pass) return false;          // the components of a row share its type
              drow[nd] = byu[u]; dwoff[nd] = blk_of[f] * R.wstride; dcls[nd] = k; dcomp[nd] = c;
              ++nd;
            }
          }
        }
      }
      return true;
    };
    if (!place(true) && !place(false)) return -1;
    for (int i = 0; i < 128; ++i) { R.rows[i] = TcRow{0, 0, 0, -1}; R.woff[i] = 0; }
    std::vector<int> tcls(128, -1), tcomp(128, 0);           // (class, component) of every tile row
    if (sp.isS) {
      for (int d = 0; d < nd; ++d)
        for (int h = 0; h < 2; ++h) {
          const int t = 32 * (d / 16) + 16 * h + (d % 16);
          R.rows[t] = drow[d]; R.woff[t] = dwoff[d]; tcls[t] = dcls[d];
        }
      R.nrows = 128;
    } else {
      for (int d = 0; d < nd; ++d) { R.rows[d] = drow[d]; R.woff[d] = dwoff[d]; tcls[d] = dcls[d]; tcomp[d] = dcomp[d]; }
      R.nrows = nd;
    }
    R.wfloats = (wblocks * R.wstride + 8 * R.O + 3) / 4 * 4;     // + one block of 8 columns of zeros: the contraction reads full blocks
    if (R.wfloats > TCR_WMAX) return -1;
    // shares of every output column inside the partial record (see TcrRole)
    for (int f = 0; f < D; ++f) for (int i = 0; i < TCR_MAXSRC; ++i) R.fsrc[f][i] = -1;
    auto add_src = [&](int f, int idx) {
      for (int i = 0; i < TCR_MAXSRC; ++i) if (R.fsrc[f][i] < 0) { R.fsrc[f][i] = (short)idx; return true; }
      return false;
    };
    for (int cw = 0; cw < TR_CONW; ++cw) {
      const int set = cw >> 2, q = cw & 3;
      if (sp.isS) {
        int k = -1;                                             // class of the quarter's rows
        for (int l = 0; l < 16; ++l) {
          const int kk = tcls[32 * q + l];
          if (kk < 0) continue;
          if (k >= 0 && kk != k) return -1;
          k = kk;
        }
        if (k < 0) continue;
        for (int h = 0; h < 2; ++h)
          for (int o6 = 0; o6 < 6; ++o6)
            if (!add_src(li.cls[k].col0 + 12 * set + 6 * h + o6, (cw * 2 + h) * 6 + o6)) return -1;
      } else {
        for (int l3 = 0; l3 < 3; ++l3) {
          int k = -1, c = -1;                                   // (class, component) of the lanes = l3 (mod 3) of this warp
          for (int l = l3; l < 32; l += 3) {
            const int t = 32 * q + l;
            if (tcls[t] < 0) continue;
            if (k >= 0 && (tcls[t] != k || tcomp[t] != c)) return -1;
            k = tcls[t]; c = tcomp[t];
          }
          if (k < 0) continue;
          for (int o = 0; o < 6; ++o)
            if (!add_src(li.cls[k].col0 + 3 * o + c, (cw * 3 + l3) * 6 + o)) return -1;
        }
      }
    }
    const int G = sp.isS ? TR_GS : TR_GV;
    if (2 * G > TCR_MAXACC || 2 * G * R.N > TR_ACOL) return -1;
  }
  return (int)specs.size();
}

// The weight slice of (layer, group, role): for every (class, f) block of the role [ncol][O] floats (the role's hidden units,
// then the bias row that multiplies Bsum; zero for roles that do not start at hidden unit 0 -- their Bsum share is counted once)
void build_tcr_weights(const LayerInfo& li, const TcrRole& R, const float* w2p, const float* b2p, float* out) {
  std::fill(out, out + R.wfloats, 0.f);
  std::vector<char> done(128, 0);
  for (int i = 0; i < R.nrows; ++i) {
    const int u = R.rows[i].u;
    if (u < 0) continue;
    int k = 0;
    while (k + 1 < li.ncls && u >= li.cls[k + 1].uoff) ++k;
    const ClassInfo& ci = li.cls[k];
    const int f = (u - ci.uoff) % ci.F;
    float* blk = out + R.woff[i];
    for (int jj = 0; jj < R.nj; ++jj)
      for (int o = 0; o < ci.O; ++o) blk[jj * ci.O + o] = w2p[ci.woff + ((int64_t)f * HID + (R.j0 + jj)) * ci.O + o];
    if (R.j0 == 0)
      for (int o = 0; o < ci.O; ++o) blk[R.nj * ci.O + o] = b2p[ci.boff + (int64_t)f * ci.O + o];
  }
}

size_t tcr_smem_bytes(int lv) {
  const size_t base = lv == 0 ? offsetof(TrSmem<0>, Wsl) : lv == 1 ? offsetof(TrSmem<1>, Wsl) : lv == 2 ? offsetof(TrSmem<2>, Wsl)
                                                                                                       : offsetof(TrSmem<3>, Wsl);
  return base + (size_t)TCR_WMAX * sizeof(float);
}

cudaError_t conv_tcr_configure() {
  cudaError_t e;
  e = cudaFuncSetAttribute(k_conv_tcr<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tcr_smem_bytes(0));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_conv_tcr<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tcr_smem_bytes(1));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_conv_tcr<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tcr_smem_bytes(2));
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_conv_tcr<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tcr_smem_bytes(3));
}

void launch_conv_tcr(DdkCtx* c, int layer, const float* x_in, float* x_out, cudaStream_t st, int mode) {
  const bool lig_only = mode == CONV_LIG;
  const LayerInfo& li = c->layers[layer];
  const int nroles = c->tcr_nroles[li.lv];
  TrArgs a;
  a.NL = c->NL; a.N = c->N;
  a.gmask = lig_only ? 0x3 : 0xf;
  const int nsegs = 2 * (lig_only ? c->NL : c->N);
  static const int tasks_per_cta = getenv("DDK_TCR_TASKS_PER_CTA") ? atoi(getenv("DDK_TCR_TASKS_PER_CTA")) : 24;   // 40-pose calls: 12 -> 214, 24 -> 219, 48 -> 205 poses/s (400 poses: capped by DDK_TCR_NB)
  static const int couple = getenv("DDK_TCR_COUPLE") ? atoi(getenv("DDK_TCR_COUPLE")) : 1;
  static const int nb_max = getenv("DDK_TCR_NB") ? atoi(getenv("DDK_TCR_NB")) : 128;
  int nb = (int)((int64_t)nsegs * nroles / (c->sm_count * tasks_per_cta)) / 8 * 8;
  a.nb_segs = std::min(std::min(TR_MAXSEG, std::max(8, nb_max / 8 * 8)), std::max(8, nb));
  a.couple = couple;
  a.glist = ptr<int4>(c->b_glist);
  a.goff[0] = 0; a.goff[1] = glist_off_group1(c); a.goff[2] = 2 * c->NL; a.goff[3] = 2 * c->NL + c->NR;
  for (int g = 0; g < 4; ++g) a.gci[g] = g;
  if (mode >= CONV_NEEDED) { const int h = mode - CONV_NEEDED; a.goff[2] = 2 * c->NL + 2 * c->NR + h * c->NR; a.gci[2] = 4 + h; }
  a.gcnt = ptr<int>(c->b_gcnt);
  a.gedges = ptr<int>(c->b_gcnt) + F3_NLIST + 4;
  a.counters = ptr<int>(c->b_counters);
  a.seg_list = ptr<int2>(c->b_seg_list);
  a.x = x_in; a.hs = ptr<float>(c->b_hs);
  a.sh_pool = ptr<float4>(c->b_sh_pool);
  a.roles = c->tcr_roles + (size_t)li.lv * TCR_MAXROLES; a.nroles = nroles;
  for (int g = 0; g < 4; ++g)
    for (int r = 0; r < TCR_MAXROLES; ++r)
      a.W[g][r] = r < nroles ? c->w2r + c->w2r_off[((size_t)layer * 4 + g) * TCR_MAXROLES + r] : nullptr;
  a.part = ptr<float>(c->b_part);
  cudaMemsetAsync(c->b_counters.p, 0, 4 * TCR_MAXROLES * sizeof(int), st);
  const int grid = c->sm_count;
  a.dbg = nullptr;
#if DDK_TCR_TRACE
  static long long* dbg = nullptr;
  if (!dbg) cudaMalloc(&dbg, 256 * 32 * sizeof(long long));
  cudaMemsetAsync(dbg, 0, 256 * 32 * sizeof(long long), st);
  a.dbg = dbg;
#endif
  {
    LaunchScope ls(c, PC_TC0 + li.lv, st);
    switch (li.lv) {
      case 0: k_conv_tcr<0><<<grid, TR_THREADS, tcr_smem_bytes(0), st>>>(a); break;
      case 1: k_conv_tcr<1><<<grid, TR_THREADS, tcr_smem_bytes(1), st>>>(a); break;
      case 2: k_conv_tcr<2><<<grid, TR_THREADS, tcr_smem_bytes(2), st>>>(a); break;
      default: k_conv_tcr<3><<<grid, TR_THREADS, tcr_smem_bytes(3), st>>>(a); break;
    }
  }
#if DDK_TCR_TRACE
  {
    std::vector<long long> h((size_t)grid * 32);
    cudaStreamSynchronize(st);
    cudaMemcpy(h.data(), dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    double s_[32] = {0};
    for (int b = 0; b < grid; ++b) for (int k = 0; k < 32; ++k) s_[k] += (double)h[(size_t)b * 32 + k] / grid;
    fprintf(stderr, "[tcr_trace] layer %d lv %d mode %d kcycles per CTA: total %.0f, tasks %.1f, reloads %.1f (%.0f) | row warp 0 (set 0): chunks %.0f, "
                    "wait staging %.0f, wait stage %.0f, operands %.0f, fences %.0f | mma: wait slot %.0f, wait operands %.0f, issue %.0f | "
                    "gather: wait ring %.0f, issue %.0f | contraction warp 0: wait acc %.0f, groups %.0f (%.0f) | vector roles: %.0f groups, fma %.0f, "
                    "reduce %.0f; scalar roles: fma %.0f, reduce %.0f | row warp 0 in vector roles: chunks %.0f, operands %.0f\n",
            layer, li.lv, mode, s_[5] / 1e3, s_[6], s_[7], s_[8] / 1e3, s_[4], s_[0] / 1e3, s_[1] / 1e3, s_[2] / 1e3, s_[3] / 1e3,
            s_[10] / 1e3, s_[11] / 1e3, s_[12] / 1e3, s_[14] / 1e3, s_[15] / 1e3, s_[18] / 1e3, s_[19] / 1e3, s_[20],
            s_[25], s_[21] / 1e3, s_[22] / 1e3, s_[23] / 1e3, s_[24] / 1e3, s_[26], s_[27] / 1e3);
  }
#endif
  {
    FinTcrArgs f;
    f.N = lig_only ? c->NL : c->N; f.dout = li.dout; f.nroles = nroles;   // ligand nodes come first
    f.NL = c->NL; f.sub = tcr_split_sub(); f.extra = 2 * c->N;
    f.seg_cnt = ptr<int>(c->b_seg_cnt);
    f.part = ptr<float>(c->b_part);
    f.roles = a.roles;
    f.bn_scale = W(c, conv_id(layer, DDK_WL_BN_SCALE));
    f.bn_shift = W(c, conv_id(layer, DDK_WL_BN_SHIFT));
    f.x_in = x_in; f.x_out = x_out;
    LaunchScope ls(c, PC_CONTRACT, st);
    k_conv_finalize_tcr<<<(f.N + 2) / 3, 256, 0, st>>>(f);
  }
}

// ---- host checks (callable without a GPU): every basis row of a level is owned by exactly one role row, the weight slices
// reproduce the packed weights, the shapes fit the kernel's budgets, and the partial records added up through TcrRole::fsrc equal the
// direct contraction.  Returns 0 if all four levels pass, else 1 + level (tables) or 10 + level (records).
int host_tcr_roles_check() {
  for (int lv = 0; lv < 4; ++lv) {
    // class tables of a layer of this level (as build_layers in ddk_api.cu)
    const int lin = lv, lout = std::min(lv + 1, 3);
    int mi0e = NS, mi1o = lin >= 1 ? NV : 0, mi1e = lin >= 2 ? NV : 0, mi0o = lin >= 3 ? NS : 0;
    int mo[4] = {NS, lout >= 1 ? NV : 0, lout >= 2 ? NV : 0, lout >= 3 ? NS : 0};
    int F[4] = {mi0e + mi1o, mi0e + mi1o + mi1e, mi1o + mi1e + mi0o, mi1e + mi0o};
    int ncomp[4] = {1, 3, 3, 1}, col0[4] = {0, 24, 42, 60};
    LayerInfo li{};
    li.lv = lv;
    int uoff = 0; int64_t woff = 0, boff = 0;
    for (int k = 0; k < 4; ++k) {
      if (F[k] == 0 || mo[k] == 0) continue;
      ClassInfo ci{};
      ci.F = F[k]; ci.O = mo[k]; ci.ncomp = ncomp[k]; ci.uoff = uoff; ci.col0 = col0[k]; ci.woff = woff; ci.boff = boff;
      li.cls[li.ncls++] = ci;
      uoff += ncomp[k] * F[k]; woff += (int64_t)F[k] * HID * mo[k]; boff += (int64_t)F[k] * mo[k];
    }
    li.U = uoff;
    std::vector<TcrRole> roles(TCR_MAXROLES);
    const int nr = build_tcr_roles(lv, li, roles.data());
    if (nr <= 0) return 1 + lv;
    // coverage: every (row u, hidden unit j) and every (row u, bias) exactly once
    std::vector<int> cover((size_t)li.U * (HID + 1), 0);
    std::vector<float> w2p((size_t)woff), b2p((size_t)boff);
    for (size_t i = 0; i < w2p.size(); ++i) w2p[i] = 1.f + (float)i;
    for (size_t i = 0; i < b2p.size(); ++i) b2p[i] = -1.f - (float)i;
    for (int r = 0; r < nr; ++r) {
      const TcrRole& R = roles[r];
      std::vector<float> sl((size_t)R.wfloats);
      build_tcr_weights(li, R, w2p.data(), b2p.data(), sl.data());
      for (int i = 0; i < R.nrows; ++i) {
        const int u = R.rows[i].u;
        if (u < 0 || (R.isS && (i & 16))) continue;          // padding lanes; the duplicate lanes of scalar roles
        int k = 0;
        while (k + 1 < li.ncls && u >= li.cls[k + 1].uoff) ++k;
        const ClassInfo& ci = li.cls[k];
        const int f = (u - ci.uoff) % ci.F;
        for (int jj = 0; jj < R.nj; ++jj) {
          cover[(size_t)u * (HID + 1) + R.j0 + jj]++;
          for (int o = 0; o < ci.O; ++o)
            if (sl[R.woff[i] + jj * ci.O + o] != w2p[ci.woff + ((int64_t)f * HID + R.j0 + jj) * ci.O + o]) return 1 + lv;
        }
        if (R.j0 == 0) {
          cover[(size_t)u * (HID + 1) + HID]++;
          for (int o = 0; o < ci.O; ++o)
            if (sl[R.woff[i] + R.nj * ci.O + o] != b2p[ci.boff + (int64_t)f * ci.O + o]) return 1 + lv;
        }
      }
    }
    for (int v : cover) if (v != 1) return 1 + lv;
    // partial records: emulate what the contraction warps write for one segment (random accumulator A[u][0..72], column 72 = Bsum)
    // and what k_conv_finalize_tcr adds up through fsrc; compare with the direct contraction  out = W2p (*) A + b2p (*) Bsum
    {
      unsigned long long st = 88172645463325252ull + (unsigned)lv;
      auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (double)(st % 20001) / 10000.0 - 1.0; };
      std::vector<double> A((size_t)li.U * (HID + 1));
      for (auto& v : A) v = rnd();
      for (auto& v : w2p) v = (float)rnd();
      for (auto& v : b2p) v = (float)rnd();
      std::vector<double> direct(D, 0.0), emul(D, 0.0);
      for (int k = 0; k < li.ncls; ++k) {
        const ClassInfo& ci = li.cls[k];
        for (int c = 0; c < ci.ncomp; ++c)
          for (int o = 0; o < ci.O; ++o) {
            double acc = 0.0;
            for (int f = 0; f < ci.F; ++f) {
              const int u = ci.uoff + c * ci.F + f;
              for (int j = 0; j < HID; ++j) acc += A[(size_t)u * (HID + 1) + j] * w2p[ci.woff + ((int64_t)f * HID + j) * ci.O + o];
              acc += A[(size_t)u * (HID + 1) + HID] * b2p[ci.boff + (int64_t)f * ci.O + o];
            }
            direct[ci.col0 + (ci.ncomp == 3 ? 3 * o + c : o)] = acc;
          }
      }
      for (int r = 0; r < nr; ++r) {
        const TcrRole& R = roles[r];
        std::vector<float> sl((size_t)R.wfloats);
        build_tcr_weights(li, R, w2p.data(), b2p.data(), sl.data());
        auto acol = [&](int u, int col) {
          return col < R.nj ? A[(size_t)u * (HID + 1) + R.j0 + col] : (col == R.nj ? A[(size_t)u * (HID + 1) + HID] : 0.0);
        };
        std::vector<double> rec(TCR_PS, 0.0);
        for (int cw = 0; cw < TR_CONW; ++cw) {
          const int set = cw >> 2, q = cw & 3;
          for (int lane = 0; lane < 32; ++lane) {
            const int t = 32 * q + lane, u = R.rows[t].u;
            if (u < 0) continue;
            if (R.isS) {
              const int h = lane >> 4;
              for (int k6 = 0; k6 < 6; ++k6) {
                const int o = 12 * set + 6 * h + k6;
                double acc = 0.0;
                for (int col = 0; col < R.ncol; ++col) acc += acol(u, col) * sl[R.woff[t] + col * 24 + o];
                rec[(cw * 2 + h) * 6 + k6] += acc;
              }
            } else {
              const int c0 = set ? 40 : 0;
              for (int o = 0; o < 6; ++o) {
                double acc = 0.0;
                for (int col = c0; col < std::min(c0 + 40, R.ncol); ++col) acc += acol(u, col) * sl[R.woff[t] + col * 6 + o];
                rec[(cw * 3 + lane % 3) * 6 + o] += acc;
              }
            }
          }
        }
        for (int f = 0; f < D; ++f)
          for (int i = 0; i < TCR_MAXSRC && R.fsrc[f][i] >= 0; ++i) emul[f] += rec[R.fsrc[f][i]];
      }
      for (int f = 0; f < D; ++f)
        if (std::abs(emul[f] - direct[f]) > 1e-6 * (1.0 + std::abs(direct[f]))) return 10 + lv;
    }
  }
  return 0;
}

}  // namespace ddk
