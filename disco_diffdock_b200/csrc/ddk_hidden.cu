// k_edge_hidden: first layer of the per-edge radial MLP of a conv layer, once per listed edge
// (/root/reference/models/tensor_layers.py:154-155 -> FCBlock, /root/reference/models/layers.py:15-22):
//     h_e = relu(W1[:, 0:24] ea_e + (W1[:, 24:48] x_src[:24] + b1) + W1[:, 48:72] x_dst[:24])
// The two node terms come from k_node_proj (one 72-vector per node and role); this kernel adds the 24 -> 72 edge term for
// every entry of every non-empty (node, group) segment and writes the 72 hidden units in the layout the fused conv
// kernel streams: slice-major, list order, hs[r][list position][J] with J = hidden units per slice of the layer's level,
// so the slice of an 8-edge chunk is one contiguous block.
// Mapping: a CTA of 96 threads = 24 hidden-unit triples (j, j + 24, j + 48) x 4 entry lanes; the thread keeps its 3 x 24
// first-layer weights in registers and reads the staged edge embeddings as warp-broadcast 16-byte loads (72 FMA per 6
// shared-memory loads).  CTAs visit the four edge groups in rotated order, so the weights are loaded four times per CTA
// and the groups' very different sizes still balance.
#include <algorithm>

#include "ddk_device.cuh"

namespace ddk {

constexpr int HT = 16;             // list entries per tile
constexpr int HID_THREADS = 96;
constexpr int EAP = 28;            // padded row of a staged edge embedding

struct HidArgs {
  int NL;
  const int4* glist; int goff[4]; int gci[4]; const int* gcnt;   // non-empty segments per group: (seg, n, base, 0)
  const int2* seg_list;
  const float* ea_pool;            // [P][24]
  const float* proj;               // [N][4][72]
  const float* W1[4];              // [72][72] row-major, edge-embedding columns 0..23
  float* hs; size_t LT; int J;
  int gmask;                       // bit g set: edge group g is processed
};

struct HidDesc {                   // a tile of <= HT consecutive list entries of one segment
  int kc, pos;                     // entries, first list position (kc = 0: no more tiles)
  const float* psr;                // source-node term of the segment (this thread's hidden units)
};
struct HidData {                   // this thread's pieces of a tile, still in registers
  float4 ea, pd[3];
  float ps[3];
};

__global__ void __launch_bounds__(HID_THREADS) k_edge_hidden(const __grid_constant__ HidArgs p) {
  __shared__ __align__(16) float sEA[HT][EAP];
  __shared__ __align__(16) float sPD[HT][HID + 4];
  const int t = threadIdx.x, jg = t % 24, es = t / 24;
  const int te = t / 6, tsub = t % 6;              // staging role: entry of the tile, 16-byte piece (6 ea + 18 pd pieces per entry)
  const int nq = gridDim.x, q0 = blockIdx.x;       // every CTA strides over the segments of every group
  const int J = p.J;
  size_t hoff[3];                  // offset of hidden unit j = jg + 24 i inside an entry's slice-major record
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int j = jg + 24 * i;
    hoff[i] = (size_t)(j / J) * p.LT * J + (j % J);
  }
  for (int gi = 0; gi < 4; ++gi) {
    const int g = (blockIdx.x + gi) & 3;
    const int nsg = ((p.gmask >> g) & 1) ? p.gcnt[p.gci[g]] : 0;
    if (q0 >= nsg) continue;
    float w[3][EA];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float4* wr = reinterpret_cast<const float4*>(p.W1[g] + (size_t)(jg + 24 * i) * HID);
#pragma unroll
      for (int k = 0; k < EA / 4; ++k) {
        const float4 v = __ldg(wr + k);
        w[i][4 * k] = v.x; w[i][4 * k + 1] = v.y; w[i][4 * k + 2] = v.z; w[i][4 * k + 3] = v.w;
      }
    }
    const int dslot = (g == 1 || g == 3) ? 3 : 2;
    const int4* gl = p.glist + p.goff[g];
    // Three tiles are in flight: the list entries of tile t + 2 and, through the entries fetched one iteration earlier,
    // the edge embeddings / destination-node terms of tile t + 1 travel while tile t is computed from shared memory.
    int si = q0, c0 = 0;
    int4 sg = load_seg_entry(gl + si);
    auto gen = [&](HidDesc& d, int2& ent) {
      d.kc = 0; d.pos = 0; d.psr = p.proj;
      ent = make_int2(0, 0);
      if (si >= nsg) return;
      d.pos = sg.z + c0;
      d.kc = min(HT, sg.y - c0);
      d.psr = p.proj + ((size_t)(sg.x >> 1) * 4 + (sg.x & 1)) * HID + jg;
      if (te < d.kc) ent = p.seg_list[d.pos + te];
      c0 += HT;
      if (c0 >= sg.y) {
        si += nq; c0 = 0;
        if (si < nsg) sg = load_seg_entry(gl + si);
      }
    };
    auto load = [&](const HidDesc& d, const int2 ent, HidData& R) {
      if (d.kc == 0) return;
#pragma unroll
      for (int i = 0; i < 3; ++i) R.ps[i] = __ldg(d.psr + 24 * i);
      if (te < d.kc) {
        R.ea = __ldg(reinterpret_cast<const float4*>(p.ea_pool + (size_t)ent.x * EA) + tsub);
        const float4* pdr = reinterpret_cast<const float4*>(p.proj + ((size_t)ent.y * 4 + dslot) * HID) + tsub;
#pragma unroll
        for (int k = 0; k < 3; ++k) R.pd[k] = __ldg(pdr + 6 * k);
      }
    };
    HidDesc d0, d1;
    int2 e0, e1;
    HidData R;
    gen(d0, e0);
    gen(d1, e1);
    load(d0, e0, R);
    while (d0.kc > 0) {
      __syncthreads();                              // the previous tile is consumed
      if (te < d0.kc) {
        *reinterpret_cast<float4*>(&sEA[te][4 * tsub]) = R.ea;
#pragma unroll
        for (int k = 0; k < 3; ++k) *reinterpret_cast<float4*>(&sPD[te][4 * (tsub + 6 * k)]) = R.pd[k];
      }
      __syncthreads();
      const int kc = d0.kc, pos = d0.pos;
      const float ps[3] = {R.ps[0], R.ps[1], R.ps[2]};
      d0 = d1; e0 = e1;
      gen(d1, e1);                                  // entries of tile t + 2
      load(d0, e0, R);                              // data of tile t + 1 (its entries arrived during the last iteration)
#pragma unroll
      for (int q = 0; q < HT / 4; ++q) {
        const int e = es + 4 * q;
        if (e < kc) {
          float a[3];
#pragma unroll
          for (int i = 0; i < 3; ++i) a[i] = ps[i] + sPD[e][jg + 24 * i];
#pragma unroll
          for (int k = 0; k < EA / 4; ++k) {
            const float4 ea = *reinterpret_cast<const float4*>(&sEA[e][4 * k]);
#pragma unroll
            for (int i = 0; i < 3; ++i)
              a[i] = fmaf(w[i][4 * k + 3], ea.w, fmaf(w[i][4 * k + 2], ea.z, fmaf(w[i][4 * k + 1], ea.y, fmaf(w[i][4 * k], ea.x, a[i]))));
          }
          float* out = p.hs + (size_t)(pos + e) * J;
#pragma unroll
          for (int i = 0; i < 3; ++i) out[hoff[i]] = fmaxf(a[i], 0.f);
        }
      }
    }
  }
}

void launch_edge_hidden(DdkCtx* c, int layer, cudaStream_t st, int mode) {
  const bool lig_only = mode == CONV_LIG;
  HidArgs a;
  a.NL = c->NL;
  a.glist = ptr<int4>(c->b_glist);
  a.goff[0] = 0; a.goff[1] = glist_off_group1(c); a.goff[2] = 2 * c->NL; a.goff[3] = 2 * c->NL + c->NR;
  for (int g = 0; g < 4; ++g) a.gci[g] = g;
  if (mode >= CONV_NEEDED) { const int h = mode - CONV_NEEDED; a.goff[2] = 2 * c->NL + 2 * c->NR + h * c->NR; a.gci[2] = 4 + h; }
  a.gcnt = ptr<int>(c->b_gcnt);
  a.seg_list = ptr<int2>(c->b_seg_list);
  a.ea_pool = ptr<float>(c->b_ea_pool);
  a.proj = ptr<float>(c->b_proj);
  for (int g = 0; g < 4; ++g) a.W1[g] = W(c, conv_id(layer, DDK_WL_W1 + g));
  a.gmask = lig_only ? 0x3 : 0xf;
  a.hs = ptr<float>(c->b_hs); a.LT = (size_t)c->list_total; a.J = conv_path() == 2 ? HID : f3_J(c->layers[layer].lv);   // k_conv_tcr reads edge-major records [list position][72]
  const int nsegs = 2 * c->N;
  const int grid = std::min(c->sm_count * 4, std::max(1, nsegs / 8));   // 4 CTAs of 96 threads are resident per SM
  LaunchScope ls(c, PC_HIDDEN, st);
  k_edge_hidden<<<grid, HID_THREADS, 0, st>>>(a);
}

}  // namespace ddk
