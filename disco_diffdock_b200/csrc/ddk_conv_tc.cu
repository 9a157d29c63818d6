// k_acc_tc: the outer-product accumulation of LONG lig<-rec segments on the 5th-generation tensor cores.
//
//     A_s[u][j] = sum_e basis_e[u] * h_e[j],   Bsum_s[u] = sum_e basis_e[u]        (ddk_conv3.cu, tensor_layers.py:65-116)
// is a GEMM with M = U basis rows (96 .. 276), N = 72 hidden units (+ a column of ones that yields Bsum), K = the edges of
// the segment.  For the cross segments of a ligand atom (K = every residue inside the cut-off, up to N_r) the FFMA2 path
// of k_conv_fused spends ~310 cycles per edge; here the segment is one accumulator in tensor memory:
//   * two gather warps move everything a chunk of 8 edges needs (list entries, destination feature rows, harmonics, the 72
//     hidden units from k_edge_hidden) from global memory into a 6-deep staging ring with cp.async; completion is signalled
//     with cp.async.mbarrier.arrive.noinc, so they run ahead across segment boundaries and nobody else waits on global memory;
//   * 4 warps per 128-row tile evaluate the basis values of the 8 edges (rows sorted by type: branch-free), split them into
//     TF32 hi + lo and write them straight into TENSOR MEMORY with tcgen05.st (the thread's TMEM lane is its row): the A
//     operand never touches shared memory; the hidden units are split the same way into the B operand in shared memory
//     (canonical K-major no-swizzle UMMA layout: 8-row x 16-byte core matrices), row 72 = 1 for valid edges (-> Bsum);
//   * one thread issues tcgen05.mma kind::tf32 (.ts form: A from TMEM, B from shared memory) three times per tile and chunk
//     (hi*hi + hi*lo + lo*hi: fp32-level accuracy, tools/microbench/umma_tf32x3.cu), M = 128, N = 80, K = 8, accumulating in
//     TMEM; tcgen05.commit releases the operand stage (3-stage ring, mbarriers) and, after the last chunk, publishes the
//     accumulator;
//   * the row warps read the accumulator back (tcgen05.ld), drop it slice by slice into a staging block in exactly the slot
//     layout [u][J | bsum] the contraction warps of k_conv_fused consume, and one thread sends each block to the scratch with a
//     bulk asynchronous store; k_conv_fused then loads these blocks instead of accumulating (f3_tc_task), so scheduling,
//     contraction, partial outputs and k_conv_finalize are shared with the FFMA2 path.
// Which segments: the group-1 work list is sorted by length; its first gcnt[F3_NLIST + 1] entries have >= TC_MIN_CHUNKS
// chunks (k_build_group_lists).  Short segments stay on the FFMA2 path, where the per-segment cost dominates anyway.
// The choice depends only on the segment's own length, so results do not depend on batch composition.
#include <cuda_pipeline_primitives.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ddk_conv.cuh"
#include "ddk_tc.cuh"

namespace ddk {

constexpr int TC_N = 80;          // MMA N: 72 hidden units, the ones column, padding to a multiple of 16
constexpr int TC_ONES = HID;      // B row that is 1 for valid edges
constexpr int TC_NST = 3;         // operand stages
constexpr int TC_COLS = 512;      // TMEM columns allocated: accumulators (3 tiles x 80) at 0, A operand stages from TC_ACOL
constexpr int TC_ACOL = 256;      // A operand (hi 8 + lo 8 columns per tile and stage): written with tcgen05.st, read by the MMA (.ts form)
constexpr int TC_BAR_ROWS = 1;    // named barrier of the row warps
constexpr int TC_XR = 6;          // staging ring (feature rows, harmonics, hidden units of a chunk) filled by the gather warps
constexpr int TC_GW = 2;          // gather warps: all global -> shared traffic (cp.async) lives here, because the row threads
                                  // execute fence.proxy.async every chunk and that fence waits for the thread's own copies

template <int LV>
struct TcCfg {
  static constexpr int U = AccCfg<LV>::U, DINP = AccCfg<LV>::DINP, XQ = DINP / 4;
  static constexpr int TILES = (U + 127) / 128;
  static constexpr int ROWT = TILES * 128;           // row threads
  static constexpr int THREADS = ROWT + 32 + 32 * TC_GW;   // + the MMA warp + the gather warps
  static constexpr int J = f3_J(LV), NSL = HID / J, AST = J + 1;
  static constexpr int NF = U * AST, NFP = (NF + 3) & ~3;   // floats of one (segment, slice) block; padded stride in the scratch
  static constexpr int A_WORDS = TILES * 128 * 8;    // one stage of A (hi or lo): [tile][k chunk 2][128 rows][4]
  static constexpr int B_WORDS = TC_N * 8;
};

template <int LV>
struct TcSmem {
  alignas(128) uint32_t Bhi[TC_NST][TcCfg<LV>::B_WORDS];
  alignas(128) uint32_t Blo[TC_NST][TcCfg<LV>::B_WORDS];
  alignas(16) float X[TC_XR][KC3][TcCfg<LV>::DINP];
  alignas(16) float SH[TC_XR][KC3][4];
  alignas(16) float HS[TC_XR][HID / TcCfg<LV>::J][KC3 * TcCfg<LV>::J + 8];   // [slice][edge * J + jj] as k_edge_hidden stores them;
                                                                        // 8 floats of padding per slice: the B-operand reads of a
                                                                        // warp span 4 slices at the same (edge, jj)
  alignas(128) float OUT[2][TcCfg<LV>::NFP];                           // read-out staging: one slice block, double-buffered
  alignas(8) unsigned long long full[TC_NST], empty[TC_NST], accfull;
  alignas(8) unsigned long long sfull[TC_XR], sempty[TC_XR];           // staging ring: gather warps <-> row warps
  uint32_t tmem_base;
};

struct TcArgs {
  const int4* glist;          // group-1 work list (seg, n, base, 0), longest first
  const int* n_long;          // entries with >= TC_MIN_CHUNKS chunks
  int cap;                    // scratch capacity (segments)
  const int2* seg_list;
  const float* x;             // [N][84]
  const float* hs; size_t LT;
  const float4* sh_pool;
  const TcRow* rows;          // [U] of the level
  float* scratch;             // [segment][slice][U][AST]
  long long* dbg;             // DDK_TC_TRACE build: per CTA cycle counters
};

#if DDK_TC_TRACE
#define TC_T(var) const long long var = clock64();
#define TC_ADD(slot, a, b) if (tid == 0 || tid == ROWT) dbgacc[slot] += (b) - (a);
#else
#define TC_T(var)
#define TC_ADD(slot, a, b)
#endif

template <int LV>
__global__ void __launch_bounds__(TcCfg<LV>::THREADS, 1) k_acc_tc(const __grid_constant__ TcArgs p) {
  using Cfg = TcCfg<LV>;
  constexpr int XQ = Cfg::XQ, TILES = Cfg::TILES, ROWT = Cfg::ROWT;
  constexpr int J = Cfg::J, NSL = Cfg::NSL, AST = Cfg::AST;
  extern __shared__ __align__(128) unsigned char tc_raw[];
  TcSmem<LV>& S = *reinterpret_cast<TcSmem<LV>*>(tc_raw);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // warp-uniform by construction
  const bool is_row = warp < ROWT / 32;

  // ---- one-time setup: TMEM, barriers, zeroed operand tiles (rows >= U and B rows > 72 stay zero for ever)
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem(&S.tmem_base)), "n"(TC_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == ROWT) {
    for (int s = 0; s < TC_NST; ++s) { tc_mbar_init(&S.full[s], ROWT / 32); tc_mbar_init(&S.empty[s], 1); }
    for (int s = 0; s < TC_XR; ++s) { tc_mbar_init(&S.sfull[s], 32 * TC_GW); tc_mbar_init(&S.sempty[s], ROWT / 32); }
    tc_mbar_init(&S.accfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {
    uint32_t* z = &S.Bhi[0][0];
    constexpr int nz = 2 * TC_NST * Cfg::B_WORDS;
    static_assert(offsetof(TcSmem<LV>, X) == nz * sizeof(uint32_t), "operand tiles are contiguous");
    for (int i = tid; i < nz; i += Cfg::THREADS) z[i] = 0u;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = S.tmem_base;
  const int n_long = min(*p.n_long, p.cap);

#if DDK_TC_TRACE
  long long dbgacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long tk0 = clock64();
#endif
  if (is_row) {
    // ================================================================== row warps: operands, read-out
    // thread p evaluates basis row rows[p].u; rows are sorted by type (build_tc_rows) so that all but two warps evaluate plain
    // products x[i0] * sh[m] without any branch; the accumulator keeps this order, the read-out restores the kernel's row order
    const TcRow rd = p.rows[tid];
    const int u = rd.u;
    const bool valid = u >= 0;
    const int tile = tid >> 7;
    const bool plain = __all_sync(0xffffffffu, rd.type == 0);
    const int i1 = rd.type == 0 ? rd.i0 : rd.i0 + 1, i2 = rd.type == 0 ? rd.i0 : rd.i0 + 2;
    const float w_t0 = rd.type == 0 ? 1.f : 0.f, w_dt = rd.type == 1 ? 1.f : 0.f;
    const float w_c1 = (rd.type == 2 && rd.m == 1) ? 1.f : 0.f, w_c2 = (rd.type == 2 && rd.m == 2) ? 1.f : 0.f,
                w_c3 = (rd.type == 2 && rd.m == 3) ? 1.f : 0.f;
    int it = 0, nseg_done = 0;
    for (int i = blockIdx.x; i < n_long; i += gridDim.x) {
      const int4 ge = load_seg_entry(p.glist + i);
      const int n = ge.y;
      const int nch = (n + KC3 - 1) / KC3;
      for (int c = 0; c < nch; ++c, ++it) {
        const int kc = min(KC3, n - c * KC3);
        const int buf = it % TC_XR, stage = it % TC_NST;
        TC_T(ta)
        tc_mbar_wait(&S.sfull[buf], (it / TC_XR) & 1);                                // the chunk's rows / harmonics / hidden units landed
        TC_T(tb)
        tc_mbar_wait(&S.empty[stage], ((it / TC_NST) & 1) ^ 1);                       // the MMAs that read this operand stage are done
        TC_T(tc_)
        TC_ADD(0, ta, tb) TC_ADD(1, tb, tc_)
        // ---- basis values of the 8 edges for this row (branch-free, the 8 edges are independent instruction streams)
        float b[KC3];
        if (plain) {
          float xv[KC3], sv[KC3];
#pragma unroll
          for (int e = 0; e < KC3; ++e) { xv[e] = S.X[buf][e][rd.i0]; sv[e] = S.SH[buf][e][rd.m]; }   // 16 independent loads
#pragma unroll
          for (int e = 0; e < KC3; ++e) b[e] = xv[e] * sv[e];
        } else {
#pragma unroll
          for (int e = 0; e < KC3; ++e) {
            const float4 s4 = *reinterpret_cast<const float4*>(&S.SH[buf][e][0]);
            const float v0 = S.X[buf][e][rd.i0], v1 = S.X[buf][e][i1], v2 = S.X[buf][e][i2];
            const float t0 = v0 * S.SH[buf][e][rd.m];
            float dt = v0 * s4.y; dt = fmaf(v1, s4.z, dt); dt = fmaf(v2, s4.w, dt);
            const float c1 = fmaf(v1, s4.w, -(v2 * s4.z)), c2 = fmaf(v2, s4.y, -(v0 * s4.w)), c3 = fmaf(v0, s4.z, -(v1 * s4.y));
            // selections as arithmetic on per-thread one-hot weights: no branch, no divergence inside the mixed warps
            b[e] = w_t0 * t0 + w_dt * dt + w_c1 * c1 + w_c2 * c2 + w_c3 * c3;
          }
        }
        {
          uint32_t hi[KC3], lo[KC3];
#pragma unroll
          for (int e = 0; e < KC3; ++e) tc_split(b[e], hi[e], lo[e]);
          // A operand straight into tensor memory: this thread's TMEM lane is its row of the tile (no shared-memory round trip:
          // with both operands in shared memory the MMAs' own operand reads saturated the shared-memory bandwidth)
          const uint32_t ta = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + TC_ACOL + (stage * TILES + tile) * 16;
          asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                       ::"r"(ta), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]), "r"(hi[4]), "r"(hi[5]), "r"(hi[6]), "r"(hi[7]) : "memory");
          asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                       ::"r"(ta + 8), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]), "r"(lo[4]), "r"(lo[5]), "r"(lo[6]), "r"(lo[7]) : "memory");
        }
        // ---- B operand: work item w = (unit j = w % 72, half hk = w / 72 of the 8 edges): row j, 4 edges per 16-byte unit;
        //      items 144 / 145 = the two halves of the ones row (NB items per thread: 2 only at level 0)
        constexpr int NB = (2 * HID + 2 + ROWT - 1) / ROWT;
#pragma unroll
        for (int b2 = 0; b2 < NB; ++b2) {
          const int w = tid + b2 * ROWT;
          if (w < 2 * HID) {
            const int hj = w % HID, hk = w / HID;
            uint32_t h[4], l[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float hld = S.HS[buf][hj / J][(4 * hk + q) * J + hj % J];   // padded edges hold a copy of the last edge: masked here
              const float hvq = 4 * hk + q < kc ? hld : 0.f;
              tc_split(hvq, h[q], l[q]);
            }
            *reinterpret_cast<uint4*>(&S.Bhi[stage][(hj + TC_N * hk) * 4]) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(&S.Blo[stage][(hj + TC_N * hk) * 4]) = make_uint4(l[0], l[1], l[2], l[3]);
          } else if (w < 2 * HID + 2) {
            const int k2 = w - 2 * HID;
            const uint32_t one = 0x3f800000u;
            *reinterpret_cast<uint4*>(&S.Bhi[stage][(TC_ONES + TC_N * k2) * 4]) =
                make_uint4(4 * k2 < kc ? one : 0u, 4 * k2 + 1 < kc ? one : 0u, 4 * k2 + 2 < kc ? one : 0u, 4 * k2 + 3 < kc ? one : 0u);
          }
        }
        __syncwarp();
        if (lane == 0) tc_mbar_arrive(&S.sempty[buf]);                               // this warp is done with the staging buffer
        TC_T(td)
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");       // this thread's A rows are in tensor memory
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // this thread's B operand stores -> async proxy
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        TC_T(te)
        TC_ADD(2, tc_, td) TC_ADD(3, td, te)
        __syncwarp();
        if (lane == 0) tc_mbar_arrive(&S.full[stage]);
      }
      // ---- read the accumulator back and write the per-slice slot blocks
      TC_T(tf)
      tc_mbar_wait(&S.accfull, nseg_done & 1);
      TC_T(tg)
      TC_ADD(4, tf, tg)
      ++nseg_done;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t v[TC_N];
#pragma unroll
      for (int c0 = 0; c0 < TC_N; c0 += 8) {
        const uint32_t taddr = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + tile * TC_N + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[c0]), "=r"(v[c0 + 1]), "=r"(v[c0 + 2]), "=r"(v[c0 + 3]), "=r"(v[c0 + 4]), "=r"(v[c0 + 5]),
                       "=r"(v[c0 + 6]), "=r"(v[c0 + 7]) : "r"(taddr));
      }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      // slice by slice: every row thread drops its J (+ bsum) values into the staging block, one thread sends the block to the
      // scratch as ONE bulk asynchronous store (scattered 4-byte stores at a 36-byte stride cost more than the MMAs)
#pragma unroll
      for (int r = 0; r < NSL; ++r) {
        float* o = &S.OUT[r & 1][0];
        if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the store that last read this buffer is done
        asm volatile("bar.sync %0, %1;" ::"n"(TC_BAR_ROWS), "n"(ROWT) : "memory");
        if (valid) {
#pragma unroll
          for (int jj = 0; jj < J; ++jj) o[u * AST + jj] = __uint_as_float(v[r * J + jj]);
          o[u * AST + J] = r == 0 ? __uint_as_float(v[TC_ONES]) : 0.f;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync %0, %1;" ::"n"(TC_BAR_ROWS), "n"(ROWT) : "memory");
        if (tid == 0) {
          float* dstp = p.scratch + ((size_t)i * NSL + r) * Cfg::NFP;
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dstp), "r"(tc_smem(o)), "n"(Cfg::NFP * 4) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      TC_T(th)
      TC_ADD(5, tg, th)
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");   // ordered before the next segment's full-barrier arrivals
    }
  } else if (warp > ROWT / 32) {
    // ================================================================== gather warps: global -> staging ring, cp.async only
    // Padded edges of a partial chunk re-copy the chunk's last edge: their hidden units and ones entry are zeroed by the row
    // threads, so whatever finite basis values they produce contribute nothing.
    const int gw = warp - ROWT / 32 - 1;               // 0: feature rows + harmonics, 1: hidden units
    int it = 0;
    for (int i = blockIdx.x; i < n_long; i += gridDim.x) {
      const int4 ge = load_seg_entry(p.glist + i);
      const int n = ge.y, base = ge.z;
      const int nch = (n + KC3 - 1) / KC3;
      int2 ent_next = make_int2(0, 0);
      if (gw == 0 && lane < KC3) ent_next = p.seg_list[base + min(lane, n - 1)];
      for (int c = 0; c < nch; ++c, ++it) {
        const int kc = min(KC3, n - c * KC3), pos0 = base + c * KC3;
        const int buf = it % TC_XR;
        const int2 ent = ent_next;
        if (gw == 0 && lane < KC3 && c + 1 < nch) ent_next = p.seg_list[pos0 + KC3 + min(lane, n - (c + 1) * KC3 - 1)];
        tc_mbar_wait(&S.sempty[buf], ((it / TC_XR) & 1) ^ 1);                          // every row warp has read the buffer's old content
        if (gw == 0) {
          constexpr int NP = KC3 * XQ + KC3;            // 16-byte pieces: feature rows, then one harmonics record per edge
#pragma unroll
          for (int k = 0; k < (NP + 31) / 32; ++k) {    // uniform trip count: the shuffles need the whole warp
            const int q0 = lane + 32 * k;
            const bool on = q0 < NP, isx = q0 < KC3 * XQ;
            const int e = !on ? 0 : (isx ? q0 / XQ : q0 - KC3 * XQ), q = isx ? q0 % XQ : 0;
            const int es = min(e, kc - 1);
            const int slot = __shfl_sync(0xffffffffu, ent.x, es), dst = __shfl_sync(0xffffffffu, ent.y, es);
            if (on) {
              if (isx) __pipeline_memcpy_async(&S.X[buf][e][4 * q], p.x + (size_t)dst * D + 4 * q, 16);
              else __pipeline_memcpy_async(&S.SH[buf][e][0], p.sh_pool + slot, 16);
            }
          }
        } else {
          constexpr int PJ = J / 4;                     // 16-byte pieces per (slice, edge)
          for (int q0 = lane; q0 < NSL * KC3 * PJ; q0 += 32) {
            const int r = q0 / (KC3 * PJ), e = (q0 / PJ) % KC3, q = q0 % PJ;
            const int es = min(e, kc - 1);
            __pipeline_memcpy_async(&S.HS[buf][r][e * J + 4 * q], p.hs + ((size_t)r * p.LT + pos0 + es) * J + 4 * q, 16);
          }
        }
        // the barrier receives this thread's arrival when all of its copies above have landed
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(tc_smem(&S.sfull[buf])) : "memory");
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  } else {
    // ================================================================== MMA issue
    // The whole warp walks the chunks with warp-uniform values and ONE elected lane issues the tcgen05 instructions, so their
    // operands live in uniform registers (issued from inside `if (lane == 0)` every UTCHMMA / UTCBAR sat in an ELECT + R2UR +
    // branch loop: ~40 cycles apiece, 9 MMAs per chunk at level 3).
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    const int n_long_u = __shfl_sync(0xffffffffu, n_long, 0);
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t bhi0 = tc_smem(&S.Bhi[0][0]), blo0 = tc_smem(&S.Blo[0][0]);
    int it = 0;
    for (int i = blockIdx.x; i < n_long_u; i += gridDim.x) {
      const int nch = (__shfl_sync(0xffffffffu, load_seg_entry(p.glist + i).y, 0) + KC3 - 1) / KC3;
      for (int c = 0; c < nch; ++c, ++it) {
        const int stage = it % TC_NST;
        TC_T(ma)
        tc_mbar_wait(&S.full[stage], (it / TC_NST) & 1);
        TC_T(mb)
        TC_ADD(6, ma, mb)
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t bh = tc_desc(bhi0 + stage * (Cfg::B_WORDS * 4), TC_N * 16, 128);
        const uint64_t bl = tc_desc(blo0 + stage * (Cfg::B_WORDS * 4), TC_N * 16, 128);
        uint32_t elected;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(elected));
        if (elected) {
#pragma unroll
          for (int t = 0; t < TILES; ++t) {
            const uint32_t ah = tmem_u + TC_ACOL + (stage * TILES + t) * 16, al = ah + 8;
            const uint32_t d = tmem_u + t * TC_N;
            tc_mma_ts(d, ah, bh, idesc, c > 0);
            tc_mma_ts(d, ah, bl, idesc, 1);
            tc_mma_ts(d, al, bh, idesc, 1);
          }
          tc_commit(&S.empty[stage]);                    // the stage may be refilled once these MMAs have read it
          if (c == nch - 1) tc_commit(&S.accfull);       // the segment's accumulator is complete
        }
        __syncwarp();
        TC_T(mc)
        TC_ADD(7, mb, mc)
      }
    }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");            // every block has reached the scratch
#if DDK_TC_TRACE
  if (p.dbg && tid == 0) { for (int k = 0; k < 6; ++k) p.dbg[blockIdx.x * 10 + k] = dbgacc[k]; p.dbg[blockIdx.x * 10 + 8] = clock64() - tk0; }
  if (p.dbg && tid == ROWT) { p.dbg[blockIdx.x * 10 + 6] = dbgacc[6]; p.dbg[blockIdx.x * 10 + 7] = dbgacc[7]; }
#endif
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TC_COLS));
}

// ---------------------------------------------------------------------------------------------- host side
void host_tc_split(float a, uint32_t* hi, uint32_t* lo) { tc_split(a, *hi, *lo); }

static int g_tc_override = -1;      // ddk_debug_set_tc: -1 = follow the environment
int tc_set_override(int on) { const int prev = g_tc_override; g_tc_override = on; return prev; }

// DDK_TC: 0 = FFMA2 kernels only (k_conv_fused), 1 = k_conv_fused + k_acc_tc for the long lig<-rec segments, 2 = k_conv_tcr: every
// accumulation on the tensor cores, contraction from tensor memory, no scratch round trip (default: the fastest measured; as
// close to the reference as mode 1 with the shipped checkpoints, mode 0 is the strict one)
int conv_path() {
  static const int env = getenv("DDK_TC") == nullptr ? 2 : atoi(getenv("DDK_TC"));
  const int v = g_tc_override < 0 ? env : g_tc_override;
  return v < 0 ? 0 : (v > 2 ? 2 : v);
}
bool tc_enabled() { return conv_path() == 1; }
void host_tc_split_rn(float a, uint32_t* hi, uint32_t* lo) { tc_split_rn(a, *hi, *lo); }

size_t tc_scratch_floats_per_segment() {
  size_t m = 0;
  for (int lv = 0; lv < 4; ++lv) {
    const int U = lv == 0 ? 96 : (lv == 1 ? 138 : (lv == 2 ? 180 : 276));
    m = std::max(m, (size_t)f3_nsl(lv) * (((size_t)U * (f3_J(lv) + 1) + 3) & ~(size_t)3));
  }
  return m;
}

cudaError_t conv_tc_configure() {
  cudaError_t e;
  e = cudaFuncSetAttribute(k_acc_tc<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TcSmem<0>));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_acc_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TcSmem<1>));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_acc_tc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TcSmem<2>));
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_acc_tc<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TcSmem<3>));
}

void launch_acc_tc(DdkCtx* c, int layer, const float* x_in, cudaStream_t st) {
  const LayerInfo& li = c->layers[layer];
  TcArgs a;
  a.glist = ptr<int4>(c->b_glist) + c->NL;             // group 1
  a.n_long = ptr<int>(c->b_gcnt) + F3_NLIST + 1;
  a.cap = c->tc_cap;
  a.seg_list = ptr<int2>(c->b_seg_list);
  a.x = x_in;
  a.hs = ptr<float>(c->b_hs); a.LT = (size_t)c->list_total;
  a.sh_pool = ptr<float4>(c->b_sh_pool);
  a.rows = c->tc_rows + li.lv * TC_MAXROWS;
  a.scratch = ptr<float>(c->b_tc_scratch);
  const int grid = std::min(c->sm_count, std::max(1, c->NL));
  a.dbg = nullptr;
#if DDK_TC_TRACE
  static long long* dbg = nullptr;
  if (!dbg) cudaMalloc(&dbg, 148 * 10 * sizeof(long long));
  cudaMemsetAsync(dbg, 0, 148 * 10 * sizeof(long long), st);
  a.dbg = dbg;
#endif
  LaunchScope ls(c, PC_TC0 + li.lv, st);
  switch (li.lv) {
    case 0: k_acc_tc<0><<<grid, TcCfg<0>::THREADS, sizeof(TcSmem<0>), st>>>(a); break;
    case 1: k_acc_tc<1><<<grid, TcCfg<1>::THREADS, sizeof(TcSmem<1>), st>>>(a); break;
    case 2: k_acc_tc<2><<<grid, TcCfg<2>::THREADS, sizeof(TcSmem<2>), st>>>(a); break;
    default: k_acc_tc<3><<<grid, TcCfg<3>::THREADS, sizeof(TcSmem<3>), st>>>(a); break;
  }
#if DDK_TC_TRACE
  {
    std::vector<long long> h(148 * 10);
    cudaStreamSynchronize(st);
    cudaMemcpy(h.data(), dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    double s[10] = {0};
    for (int b = 0; b < grid; ++b) for (int k = 0; k < 10; ++k) s[k] += (double)h[b * 10 + k] / grid;
    fprintf(stderr, "[tc_trace] layer %d lv %d kcycles per CTA: total %.0f | row thread 0: wait staging %.0f, wait operand stage %.0f, "
                    "compute+store %.0f, proxy fence %.0f, wait accumulator %.0f, read-out %.0f | mma thread: wait full %.0f, issue %.0f\n",
            layer, li.lv, s[8] / 1e3, s[0] / 1e3, s[1] / 1e3, s[2] / 1e3, s[3] / 1e3, s[4] / 1e3, s[5] / 1e3, s[6] / 1e3, s[7] / 1e3);
  }
#endif
}

}  // namespace ddk
