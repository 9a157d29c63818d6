// k_conv_accum2: persistent, warp-specialised outer-product accumulation (see ddk_conv.cu for the algebra).
//
// One CTA per SM walks the (node, group) segments of a chunk round-robin in LPT order.  Inside the CTA
//   - 6 producer warps gather the edge embedding / harmonics / destination features of the next 32 edges with
//     cp.async (double buffered, one chunk ahead), evaluate the first radial-MLP layer h_e (72) and the basis
//     functions basis_e (U) and publish them in a shared-memory stage;
//   - 9 consumer warps do nothing but the rank-1 updates A[u][j] += basis_e[u] * h_e[j] out of that stage
//     (72 independent FFMA per thread and edge), and write A to the scratch when a segment ends.
// Producers and consumers hand stages over with named barriers (bar.sync / bar.arrive), so the gather latency, the
// first MLP layer and the basis evaluation overlap the FMA stream instead of alternating with it.
#include <cuda_pipeline_primitives.h>

#include "ddk_conv.cuh"

namespace ddk {

constexpr int NCONS = 288;   // 9 consumer warps (warp w: columns 8w..8w+7, lane: rows lane + 32 a)
constexpr int NPROD = 192;   // 6 producer warps
constexpr int ACC2_THREADS = NCONS + NPROD;
constexpr int EA_STRIDE = 28;  // padded row of the staged edge embedding (conflict-free LDS.128 across edges)

enum { BAR_FULL0 = 1, BAR_FULL1 = 2, BAR_EMPTY0 = 3, BAR_EMPTY1 = 4, BAR_PROD = 5 };

__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

template <int LV>
struct Acc2Smem {
  struct Stage {
    alignas(16) float H[KC][HID];
    alignas(16) float B[KC][AccCfg<LV>::BS];
    int kc;          // edges in this stage, -1 = no more work
    int sidx_last;   // scratch index of the segment when this stage is its last one, else -1
  } st[2];
  struct Gather {
    alignas(16) float Ea[KC][EA_STRIDE];
    alignas(16) float Sh[KC][4];
    alignas(16) float Xd[KC][AccCfg<LV>::DINP];
    alignas(16) float Pd[KC][HID];
    alignas(16) float Ps[HID];
  } ga[3];
  int2 ent[3][KC];                      // (slot, dst) entries of the chunks whose gathers are being issued
  alignas(16) float W1aT[4][EA][HID];   // [group][k][j] = W1_g[j][k], k < 24 (edge-embedding columns)
};

struct ChunkDesc {
  int seg, node, which, g, base, c0, kc, last, sidx, done;
};

// Ordered compaction of the non-empty segments of one graph chunk into packed descriptors (seg, n, base, sidx), so
// the accumulate kernel fetches a segment with ONE 16-byte load that it can issue a whole segment ahead of time.
__global__ void __launch_bounds__(1024) k_build_worklist(const int* __restrict__ seg_order, int nseg, const int* __restrict__ seg_cnt,
                                                         const int* __restrict__ seg_base, const int* __restrict__ seg_sidx,
                                                         int4* __restrict__ work, int* __restrict__ n_work) {
  __shared__ int wsum[32];
  __shared__ int base_s;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) base_s = 0;
  __syncthreads();
  for (int i0 = 0; i0 < nseg; i0 += 1024) {
    const int i = i0 + tid;
    int seg = 0, n = 0;
    if (i < nseg) { seg = seg_order[i]; n = seg_cnt[seg]; }
    const unsigned m = __ballot_sync(0xffffffffu, n > 0);
    if (lane == 0) wsum[w] = __popc(m);
    __syncthreads();
    int off = base_s;
    for (int q = 0; q < w; ++q) off += wsum[q];
    if (n > 0) work[off + __popc(m & ((1u << lane) - 1))] = make_int4(seg, n, seg_base[seg], seg_sidx[seg]);
    __syncthreads();
    if (tid == 0) { int t = 0; for (int q = 0; q < 32; ++q) t += wsum[q]; base_s += t; }
    __syncthreads();
  }
  if (tid == 0) *n_work = base_s;
}

template <int LV>
__global__ void __launch_bounds__(ACC2_THREADS, 1) k_conv_accum2(AccArgs p, const int4* __restrict__ work, const int* __restrict__ n_work) {
  constexpr int U = AccCfg<LV>::U;
  constexpr int NA = AccCfg<LV>::NA;
  constexpr int BS = AccCfg<LV>::BS;
  constexpr int DINP = AccCfg<LV>::DINP;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Acc2Smem<LV>& S = *reinterpret_cast<Acc2Smem<LV>*>(smem_raw);
  const int tid = threadIdx.x;

  if (tid >= NCONS) {
    // ============================================================== producers
    const int t = tid - NCONS;
    for (int i = t; i < 4 * EA * HID; i += NPROD) {
      int g = i / (EA * HID), k = (i / HID) % EA, j = i % HID;
      S.W1aT[g][k][j] = p.W1[g][j * HID + k];
    }
    // basis rows of this thread (u = t and, for t + 192 < BS, u = t + 192)
    int btype[2], bi0[2], bm[2];
    basis_desc<LV>(t < U ? t : 0, btype[0], bi0[0], bm[0]);
    basis_desc<LV>(t + NPROD < U ? t + NPROD : 0, btype[1], bi0[1], bm[1]);
    float bsum[2] = {0.f, 0.f};

    // static round-robin walk over the compacted, LPT-ordered work list, one chunk of <= KC edges at a time; the
    // descriptor of the next segment is loaded one segment ahead so its latency never sits on the critical path
    const int nwork = *n_work;
    int oi = blockIdx.x;               // work index of the descriptor held in `pre`
    int4 pre = oi < nwork ? work[oi] : make_int4(-1, 0, 0, 0);
    int seg = -1, n = 0, c0 = 0, sbase = 0, ssidx = 0;
    auto next_desc = [&]() {
      ChunkDesc d;
      d.done = 0;
      if (seg < 0 || c0 >= n) {
        if (pre.x < 0) { d.done = 1; d.kc = -1; d.seg = -1; d.node = d.which = d.g = d.base = d.c0 = d.last = d.sidx = 0; return d; }
        seg = pre.x; n = pre.y; sbase = pre.z; ssidx = pre.w;
        c0 = 0;
        oi += gridDim.x;
        pre = oi < nwork ? work[oi] : make_int4(-1, 0, 0, 0);
      }
      d.seg = seg; d.node = seg >> 1; d.which = seg & 1;
      d.g = d.node < p.NL ? d.which : 2 + d.which;
      d.base = sbase;
      d.c0 = c0;
      d.kc = min(KC, n - c0);
      d.last = (c0 + d.kc >= n);
      d.sidx = ssidx;
      c0 += d.kc;
      return d;
    };
    // chunk-invariant (edge, 16-byte piece) coordinates of the copies this thread issues
    constexpr int XQ = DINP / 4, PQ = HID / 4, EQ = EA / 4;
    constexpr int NX = (KC * XQ + NPROD - 1) / NPROD, NP = (KC * PQ + NPROD - 1) / NPROD;
    int xe[NX], xq[NX], pe[NP], pq[NP];
#pragma unroll
    for (int r = 0; r < NX; ++r) { int i = t + r * NPROD; xe[r] = i / XQ; xq[r] = i % XQ; }
#pragma unroll
    for (int r = 0; r < NP; ++r) { int i = t + r * NPROD; pe[r] = i / PQ; pq[r] = i % PQ; }
    const int ee = t / EQ, eq = t % EQ;     // KC * EQ == NPROD: exactly one edge-embedding piece per thread
    static_assert(KC * (EA / 4) == NPROD, "one Ea piece per producer thread");
    auto issue_gather = [&](const ChunkDesc& d, typename Acc2Smem<LV>::Gather& G, const int2* lst) {
      const int dslot = (d.g == 1 || d.g == 3) ? 3 : 2;
      if (ee < d.kc) __pipeline_memcpy_async(&G.Ea[ee][4 * eq], p.ea_pool + (size_t)lst[ee].x * EA + 4 * eq, 16);
      if (t < d.kc) __pipeline_memcpy_async(&G.Sh[t][0], p.sh_pool + lst[t].x, 16);
#pragma unroll
      for (int r = 0; r < NX; ++r)
        if (xe[r] < d.kc) __pipeline_memcpy_async(&G.Xd[xe[r]][4 * xq[r]], p.x + (size_t)lst[xe[r]].y * D + 4 * xq[r], 16);
#pragma unroll
      for (int r = 0; r < NP; ++r)
        if (pe[r] < d.kc)
          __pipeline_memcpy_async(&G.Pd[pe[r]][4 * pq[r]], p.proj + ((size_t)lst[pe[r]].y * 4 + dslot) * HID + 4 * pq[r], 16);
      if (t < HID / 4) __pipeline_memcpy_async(&G.Ps[4 * t], p.proj + ((size_t)d.node * 4 + d.which) * HID + 4 * t, 16);
      __pipeline_commit();
    };
    // the (slot, dst) entries themselves are fetched by one warp a full iteration before the gathers that need them
    int2 ent_reg = make_int2(0, 0);
    auto load_ent = [&](const ChunkDesc& d) {
      if (!d.done && t < d.kc) ent_reg = p.seg_list[d.base + d.c0 + t];
    };

    // gathers run two chunks ahead of the math (three gather buffers), entry lists three chunks ahead
    ChunkDesc dq[4];
    dq[0] = next_desc();
    dq[1] = dq[0].done ? dq[0] : next_desc();
    dq[2] = dq[1].done ? dq[1] : next_desc();
#pragma unroll
    for (int k = 0; k < 2; ++k) {                    // prologue: chunks 0 and 1 pay the entry latency once
      load_ent(dq[k]);
      if (t < KC) S.ent[k][t] = ent_reg;
      bar_sync(BAR_PROD, NPROD);
      if (!dq[k].done) issue_gather(dq[k], S.ga[k], S.ent[k]); else __pipeline_commit();
    }
    load_ent(dq[2]);
    for (int it = 0;; ++it) {
      const int s = it & 1;
      const ChunkDesc cur = dq[0];
      __pipeline_wait_prior(1);                      // this thread's copies of chunk `it` have landed
      if (t < KC) S.ent[(it + 2) % 3][t] = ent_reg;  // entries of chunk it+2 (loaded during the previous iteration)
      bar_sync(BAR_PROD, NPROD);                     // ... everybody's have, and everybody is done with chunk it-1's buffer
      dq[3] = dq[2].done ? dq[2] : next_desc();
      load_ent(dq[3]);                               // entries of chunk it+3 start travelling
      if (!dq[2].done) issue_gather(dq[2], S.ga[(it + 2) % 3], S.ent[(it + 2) % 3]); else __pipeline_commit();
      if (it >= 2) bar_sync(s ? BAR_EMPTY1 : BAR_EMPTY0, ACC2_THREADS);   // consumers released stage s
      typename Acc2Smem<LV>::Stage& T = S.st[s];
      if (cur.done) {
        if (t == 0) { T.kc = -1; T.sidx_last = -1; }
        __threadfence_block();
        bar_arrive(s ? BAR_FULL1 : BAR_FULL0, ACC2_THREADS);
        break;
      }
      typename Acc2Smem<LV>::Gather& G = S.ga[it % 3];
      // ---- first radial-MLP layer: h = relu(W1[:, :24] ea + (W1[:,24:48] x_s + b1) + W1[:,48:72] x_d)
      {
        const int e = t & 31, jb = t >> 5;           // 12 hidden units per thread
        if (e < cur.kc) {
          float h[12];
          const float4* pd = reinterpret_cast<const float4*>(&G.Pd[e][12 * jb]);
          const float4* ps = reinterpret_cast<const float4*>(&G.Ps[12 * jb]);
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            float4 a = pd[q], b = ps[q];
            h[4 * q] = a.x + b.x; h[4 * q + 1] = a.y + b.y; h[4 * q + 2] = a.z + b.z; h[4 * q + 3] = a.w + b.w;
          }
#pragma unroll
          for (int q = 0; q < EA / 4; ++q) {
            float4 ea = reinterpret_cast<const float4*>(&G.Ea[e][0])[q];
            const float eav[4] = {ea.x, ea.y, ea.z, ea.w};
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const float4* wr = reinterpret_cast<const float4*>(&S.W1aT[cur.g][4 * q + kk][12 * jb]);
#pragma unroll
              for (int r = 0; r < 3; ++r) {
                float4 wv = wr[r];
                h[4 * r] += wv.x * eav[kk]; h[4 * r + 1] += wv.y * eav[kk]; h[4 * r + 2] += wv.z * eav[kk]; h[4 * r + 3] += wv.w * eav[kk];
              }
            }
          }
          float4* out = reinterpret_cast<float4*>(&T.H[e][12 * jb]);
#pragma unroll
          for (int q = 0; q < 3; ++q)
            out[q] = make_float4(fmaxf(h[4 * q], 0.f), fmaxf(h[4 * q + 1], 0.f), fmaxf(h[4 * q + 2], 0.f), fmaxf(h[4 * q + 3], 0.f));
        }
      }
      // ---- basis functions (raw products; the constant factors live in the packed weights).  The per-row type is
      //      hoisted out of the edge loop so the loop bodies are straight-line LDS / FMUL / STS with constant strides.
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int u = t + r * NPROD;
        if (u < BS) {
          float* bp = &T.B[0][u];
          if (u < U) {
            const int ty = btype[r], i0 = bi0[r], m = bm[r];
            const float* xp = &G.Xd[0][i0];
            const float* sp = &G.Sh[0][0];
            float acc = 0.f;
            if (ty == 0) {
#pragma unroll 4
              for (int e = 0; e < cur.kc; ++e) {
                float b = xp[e * DINP] * sp[e * 4 + m];
                bp[e * BS] = b;
                acc += b;
              }
            } else if (ty == 1) {
#pragma unroll 4
              for (int e = 0; e < cur.kc; ++e) {
                float b = xp[e * DINP] * sp[e * 4 + 1] + xp[e * DINP + 1] * sp[e * 4 + 2] + xp[e * DINP + 2] * sp[e * 4 + 3];
                bp[e * BS] = b;
                acc += b;
              }
            } else {
              const int c = m - 1, c1 = (c + 1) % 3, c2 = (c + 2) % 3;
#pragma unroll 4
              for (int e = 0; e < cur.kc; ++e) {
                float b = xp[e * DINP + c1] * sp[e * 4 + 1 + c2] - xp[e * DINP + c2] * sp[e * 4 + 1 + c1];
                bp[e * BS] = b;
                acc += b;
              }
            }
            bsum[r] += acc;
          } else {
            for (int e = 0; e < cur.kc; ++e) bp[e * BS] = 0.f;
          }
        }
      }
      if (cur.last) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int u = t + r * NPROD;
          if (u < U) p.Bsum[(size_t)cur.sidx * U + u] = bsum[r];
          bsum[r] = 0.f;
        }
      }
      if (t == 0) { T.kc = cur.kc; T.sidx_last = cur.last ? cur.sidx : -1; }
      __threadfence_block();
      bar_arrive(s ? BAR_FULL1 : BAR_FULL0, ACC2_THREADS);
      dq[0] = dq[1]; dq[1] = dq[2]; dq[2] = dq[3];
    }
    return;
  }

  // ================================================================ consumers
  const int lane = tid & 31, w = tid >> 5;
  float acc[NA][8];
#pragma unroll
  for (int a = 0; a < NA; ++a)
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) acc[a][jj] = 0.f;
  for (int it = 0;; ++it) {
    const int s = it & 1;
    bar_sync(s ? BAR_FULL1 : BAR_FULL0, ACC2_THREADS);
    const typename Acc2Smem<LV>::Stage& T = S.st[s];
    const int kc = T.kc;
    if (kc < 0) break;
    const int sidx_last = T.sidx_last;
#pragma unroll 2
    for (int e = 0; e < kc; ++e) {
      float4 h0 = reinterpret_cast<const float4*>(&T.H[e][8 * w])[0];
      float4 h1 = reinterpret_cast<const float4*>(&T.H[e][8 * w])[1];
#pragma unroll
      for (int a = 0; a < NA; ++a) {
        float b = T.B[e][lane + 32 * a];
        acc[a][0] += b * h0.x; acc[a][1] += b * h0.y; acc[a][2] += b * h0.z; acc[a][3] += b * h0.w;
        acc[a][4] += b * h1.x; acc[a][5] += b * h1.y; acc[a][6] += b * h1.z; acc[a][7] += b * h1.w;
      }
    }
    bar_arrive(s ? BAR_EMPTY1 : BAR_EMPTY0, ACC2_THREADS);   // stage s (and its header) fully read
    if (sidx_last >= 0) {
      float* Aout = p.A + (size_t)sidx_last * U * HID;
#pragma unroll
      for (int a = 0; a < NA; ++a) {
        int u = lane + 32 * a;
        if (u < U) {
          float4* dst = reinterpret_cast<float4*>(Aout + (size_t)u * HID + 8 * w);
          dst[0] = make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
          dst[1] = make_float4(acc[a][4], acc[a][5], acc[a][6], acc[a][7]);
        }
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) acc[a][jj] = 0.f;
      }
    }
  }
}

cudaError_t conv2_configure() {
  cudaError_t e;
  e = cudaFuncSetAttribute(k_conv_accum2<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Acc2Smem<0>));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_conv_accum2<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Acc2Smem<1>));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_conv_accum2<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Acc2Smem<2>));
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_conv_accum2<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Acc2Smem<3>));
}

void launch_build_worklist(DdkCtx* c, cudaStream_t st) {
  int ci = 0;
  for (const Chunk& ch : c->chunks) {
    LaunchScope ls(c, PC_GRAPH, st);
    k_build_worklist<<<1, 1024, 0, st>>>(ptr<int>(c->b_seg_order) + ch.order_off, ch.nseg, ptr<int>(c->b_seg_cnt),
                                         ptr<int>(c->b_seg_base), ptr<int>(c->b_seg_sidx),
                                         ptr<int4>(c->b_work) + ch.order_off, ptr<int>(c->b_nwork) + ci);
    ++ci;
  }
}

void launch_conv_accum2(DdkCtx* c, const LayerInfo& li, const Chunk& ch, const AccArgs& a, cudaStream_t st) {
  static int sms = 0;
  if (sms == 0) {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, c->device);
    sms = prop.multiProcessorCount;
  }
  const int grid = std::min(ch.nseg, sms);
  const int ci = (int)(&ch - c->chunks.data());
  const int4* work = ptr<int4>(c->b_work) + ch.order_off;
  const int* nwork = ptr<int>(c->b_nwork) + ci;
  LaunchScope ls(c, PC_ACC0 + li.lv, st);
  switch (li.lv) {
    case 0: k_conv_accum2<0><<<grid, ACC2_THREADS, sizeof(Acc2Smem<0>), st>>>(a, work, nwork); break;
    case 1: k_conv_accum2<1><<<grid, ACC2_THREADS, sizeof(Acc2Smem<1>), st>>>(a, work, nwork); break;
    case 2: k_conv_accum2<2><<<grid, ACC2_THREADS, sizeof(Acc2Smem<2>), st>>>(a, work, nwork); break;
    default: k_conv_accum2<3><<<grid, ACC2_THREADS, sizeof(Acc2Smem<3>), st>>>(a, work, nwork); break;
  }
}

}  // namespace ddk
