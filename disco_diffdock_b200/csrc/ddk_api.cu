// C ABI of libddk (include/ddk.h): context / batch management and the step drivers.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <cstdlib>

#include "ddk_conv.cuh"

using namespace ddk;

namespace {

std::string g_create_error;

#define DDK_CUDA_TRY(ctx, expr)                                                                     \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      (ctx)->err = std::string(#expr) + ": " + cudaGetErrorString(_e);                              \
      return DDK_ERR_CUDA;                                                                          \
    }                                                                                               \
  } while (0)

int fail(DdkCtx* c, int code, const std::string& msg) {
  if (c) c->err = msg; else g_create_error = msg;
  return code;
}

int ensure(DdkCtx* c, Buf& b, size_t bytes) {
  if (bytes == 0) bytes = 16;
  if (b.bytes >= bytes) return DDK_OK;
  if (b.p) DDK_CUDA_TRY(c, cudaFree(b.p));
  b.p = nullptr; b.bytes = 0;
  size_t want = bytes + bytes / 8;   // a little head-room so slowly growing batches do not reallocate every time
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e != cudaSuccess) { c->err = std::string("cudaMalloc: ") + cudaGetErrorString(e); return DDK_ERR_NOMEM; }
  b.bytes = want;
  return DDK_OK;
}

template <typename T>
int upload(DdkCtx* c, Buf& b, const std::vector<T>& v, cudaStream_t st) {
  int rc = ensure(c, b, v.size() * sizeof(T));
  if (rc) return rc;
  if (!v.empty()) DDK_CUDA_TRY(c, cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
  return DDK_OK;
}

void build_layers(DdkCtx* c) {
  c->layers.clear();
  for (int l = 0; l < c->cfg.num_conv_layers; ++l) {
    int lin = std::min(l, 3), lout = std::min(l + 1, 3);
    int mi0e = NS, mi1o = lin >= 1 ? NV : 0, mi1e = lin >= 2 ? NV : 0, mi0o = lin >= 3 ? NS : 0;
    int mo[4] = {NS, lout >= 1 ? NV : 0, lout >= 2 ? NV : 0, lout >= 3 ? NS : 0};
    int F[4] = {mi0e + mi1o, mi0e + mi1o + mi1e, mi1o + mi1e + mi0o, mi1e + mi0o};
    int ncomp[4] = {1, 3, 3, 1}, col0[4] = {0, 24, 42, 60};
    LayerInfo li{};
    li.lv = lin;
    li.dout = NS + 3 * mo[1] + 3 * mo[2] + mo[3];
    int uoff = 0; int64_t woff = 0, boff = 0; li.ncls = 0;
    for (int k = 0; k < 4; ++k) {
      if (F[k] == 0 || mo[k] == 0) continue;
      ClassInfo ci{};
      ci.F = F[k]; ci.O = mo[k]; ci.ncomp = ncomp[k]; ci.uoff = uoff; ci.col0 = col0[k]; ci.woff = woff; ci.boff = boff;
      li.cls[li.ncls++] = ci;
      uoff += ncomp[k] * F[k];
      woff += (int64_t)F[k] * HID * mo[k];
      boff += (int64_t)F[k] * mo[k];
    }
    li.U = uoff;
    li.NA = (uoff + 31) / 32;
    c->layers.push_back(li);
  }
}

void free_buf(Buf& b) {
  if (b.p) cudaFree(b.p);
  b.p = nullptr; b.bytes = 0;
}

}  // namespace

extern "C" {

int ddk_abi_version(void) { return DDK_ABI_VERSION; }

const char* ddk_last_error(const DdkCtx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int ddk_create(const DdkConfig* cfg, const float* weights_h, size_t n_floats, const int64_t* offsets_h, int32_t n_offsets,
               int32_t device, DdkCtx** out) {
  if (!cfg || !weights_h || !offsets_h || !out) return fail(nullptr, DDK_ERR_INVALID, "null argument");
  if (cfg->abi_version != DDK_ABI_VERSION) return fail(nullptr, DDK_ERR_INVALID, "ABI version mismatch");
  if (cfg->ns != NS || cfg->nv != NV) return fail(nullptr, DDK_ERR_INVALID, "kernels are compiled for ns=24, nv=6");
  if (cfg->num_conv_layers < 3 || cfg->num_conv_layers > 8) return fail(nullptr, DDK_ERR_INVALID, "num_conv_layers must be 3..8");
  if (cfg->latent_dim < 0 || cfg->latent_dim > 2) return fail(nullptr, DDK_ERR_INVALID, "latent_dim must be 0..2");
  if (n_offsets != DDK_W_CONV_BASE + cfg->num_conv_layers * DDK_W_CONV_STRIDE)
    return fail(nullptr, DDK_ERR_INVALID, "offset table has the wrong length");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, DDK_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail(nullptr, DDK_ERR_INVALID, "device index out of range");
  DdkCtx* c = new DdkCtx();
  c->cfg = *cfg;
  c->device = device;
  c->off.assign(offsets_h, offsets_h + n_offsets);
  for (int64_t o : c->off)
    if (o >= (int64_t)n_floats) { delete c; return fail(nullptr, DDK_ERR_INVALID, "offset beyond the weight blob"); }
  double rl = cfg->lig_max_radius, rc = cfg->cross_max_distance;
  c->r2_lig = (float)(rl * rl);
  c->r2_cross = (float)(rc * rc);
  build_layers(c);
  auto bail = [&](const char* what, cudaError_t err) {
    g_create_error = std::string(what) + ": " + cudaGetErrorString(err);
    if (c->w) cudaFree(c->w);
    if (c->w2s) cudaFree(c->w2s);
    if (c->ltab) cudaFree(c->ltab);
    delete c;
    return DDK_ERR_CUDA;
  };
  if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
  if ((e = cudaMalloc(&c->w, n_floats * sizeof(float))) != cudaSuccess) return bail("cudaMalloc(weights)", e);
  if ((e = cudaMemcpy(c->w, weights_h, n_floats * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess)
    return bail("cudaMemcpy(weights)", e);
  const size_t ncounter = (2 + 2 * F3_NLIST) * sizeof(unsigned long long);
  if ((e = cudaMalloc(&c->b_edge_total.p, ncounter)) != cudaSuccess) return bail("cudaMalloc(counter)", e);
  c->b_edge_total.bytes = ncounter;   // cumulative: [0] dynamic edges, [1] non-empty segments, then edges / segments per work list
  if ((e = cudaMemset(c->b_edge_total.p, 0, ncounter)) != cudaSuccess) return bail("cudaMemset(counter)", e);
  if ((e = heads_configure()) != cudaSuccess) return bail("cudaFuncSetAttribute(heads)", e);
  if ((e = conv3_configure()) != cudaSuccess) return bail("cudaFuncSetAttribute(conv3)", e);
  if ((e = conv_tc_configure()) != cudaSuccess) return bail("cudaFuncSetAttribute(conv_tc)", e);
  if ((e = conv_tcr_configure()) != cudaSuccess) return bail("cudaFuncSetAttribute(conv_tcr)", e);
  {
    // k_conv_tcr: role tables per basis level, weight slices per (layer, group, role)
    std::vector<TcrRole> roles((size_t)4 * TCR_MAXROLES);
    for (int lv = 0; lv < 4; ++lv) {
      const LayerInfo* li = nullptr;
      for (const LayerInfo& l : c->layers) if (l.lv == lv) { li = &l; break; }
      if (!li) continue;
      c->tcr_nroles[lv] = build_tcr_roles(lv, *li, roles.data() + (size_t)lv * TCR_MAXROLES);
      if (c->tcr_nroles[lv] <= 0) {
        g_create_error = "internal: tensor-core role tables of a basis level are inconsistent";
        cudaFree(c->w); delete c;
        return DDK_ERR_STATE;
      }
    }
    if ((e = cudaMalloc(&c->tcr_roles, roles.size() * sizeof(TcrRole))) != cudaSuccess) return bail("cudaMalloc(tcr_roles)", e);
    if ((e = cudaMemcpy(c->tcr_roles, roles.data(), roles.size() * sizeof(TcrRole), cudaMemcpyHostToDevice)) != cudaSuccess)
      return bail("cudaMemcpy(tcr_roles)", e);
    std::vector<float> w2r;
    c->w2r_off.assign((size_t)c->cfg.num_conv_layers * 4 * TCR_MAXROLES, 0);
    for (int l = 0; l < c->cfg.num_conv_layers; ++l) {
      const LayerInfo& li = c->layers[l];
      for (int g = 0; g < 4; ++g)
        for (int r = 0; r < c->tcr_nroles[li.lv]; ++r) {
          const TcrRole& R = roles[(size_t)li.lv * TCR_MAXROLES + r];
          c->w2r_off[((size_t)l * 4 + g) * TCR_MAXROLES + r] = (int64_t)w2r.size();
          w2r.resize(w2r.size() + (size_t)R.wfloats);
          build_tcr_weights(li, R, weights_h + c->off[conv_id(l, DDK_WL_W2P + g)], weights_h + c->off[conv_id(l, DDK_WL_B2P + g)],
                            w2r.data() + w2r.size() - (size_t)R.wfloats);
        }
    }
    if ((e = cudaMalloc(&c->w2r, w2r.size() * sizeof(float))) != cudaSuccess) return bail("cudaMalloc(w2r)", e);
    if ((e = cudaMemcpy(c->w2r, w2r.data(), w2r.size() * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess)
      return bail("cudaMemcpy(w2r)", e);
  }
  {
    std::vector<TcRow> rows((size_t)4 * TC_MAXROWS);
    for (int lv = 0; lv < 4; ++lv) build_tc_rows(lv, rows.data() + (size_t)lv * TC_MAXROWS);
    if ((e = cudaMalloc(&c->tc_rows, rows.size() * sizeof(TcRow))) != cudaSuccess) return bail("cudaMalloc(tc_rows)", e);
    if ((e = cudaMemcpy(c->tc_rows, rows.data(), rows.size() * sizeof(TcRow), cudaMemcpyHostToDevice)) != cudaSuccess)
      return bail("cudaMemcpy(tc_rows)", e);
  }
  {
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return bail("cudaGetDeviceProperties", e);
    c->sm_count = prop.multiProcessorCount;
  }
  {
    // second-layer weights re-sliced for the fused kernel: W2S[l][g][r][class][f][jj][o] = W2p[l][g][class][(f, J r + jj)][o], J = f3_J(level)
    std::vector<float> w2s;
    c->w2s_off.assign((size_t)c->cfg.num_conv_layers * 4, 0);
    c->con_split.resize(c->cfg.num_conv_layers);
    for (int l = 0; l < c->cfg.num_conv_layers; ++l) {
      const LayerInfo& li = c->layers[l];
      build_con_split(li, c->con_split[l]);
      for (int g = 0; g < 4; ++g) {
        const float* src = weights_h + c->off[conv_id(l, DDK_WL_W2P + g)];
        c->w2s_off[(size_t)l * 4 + g] = (int64_t)w2s.size();
        const int J = f3_J(li.lv);
        for (int r = 0; r < HID / J; ++r)
          for (int k = 0; k < li.ncls; ++k) {
            const ClassInfo& ci = li.cls[k];
            for (int f = 0; f < ci.F; ++f)
              for (int jj = 0; jj < J; ++jj)
                for (int o = 0; o < ci.O; ++o)
                  w2s.push_back(src[ci.woff + ((int64_t)f * HID + (J * r + jj)) * ci.O + o]);
          }
      }
    }
    if ((e = cudaMalloc(&c->w2s, w2s.size() * sizeof(float))) != cudaSuccess) return bail("cudaMalloc(w2s)", e);
    if ((e = cudaMemcpy(c->w2s, w2s.data(), w2s.size() * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess)
      return bail("cudaMemcpy(w2s)", e);
    std::vector<LaneTab> tabs(4 * 32);
    for (int lv = 0; lv < 4; ++lv)
      if (!build_lane_table(lv, tabs.data() + lv * 32)) {
        g_create_error = "internal: basis rows of a level are not covered exactly once by the lane table";
        cudaFree(c->w); cudaFree(c->w2s); delete c;
        return DDK_ERR_STATE;
      }
    if ((e = cudaMalloc(&c->ltab, tabs.size() * sizeof(LaneTab))) != cudaSuccess) return bail("cudaMalloc(ltab)", e);
    if ((e = cudaMemcpy(c->ltab, tabs.data(), tabs.size() * sizeof(LaneTab), cudaMemcpyHostToDevice)) != cudaSuccess)
      return bail("cudaMemcpy(ltab)", e);
  }
  *out = c;
  return DDK_OK;
}

int ddk_destroy(DdkCtx* c) {
  if (!c) return DDK_OK;
  cudaSetDevice(c->device);
  Buf* all[] = {&c->b_lig_ptr, &c->b_rec_ptr, &c->b_lig_graph, &c->b_rec_graph, &c->b_bond_src, &c->b_bond_dst, &c->b_rr_src,
                &c->b_rr_dst, &c->b_rot_u, &c->b_rot_v, &c->b_rot_ptr, &c->b_rot_graph, &c->b_mr_off, &c->b_ll_off,
                &c->b_lr_off, &c->b_seg_base, &c->b_seg_static, &c->b_seg_cnt, &c->b_seg_list,
                &c->b_lig_static, &c->b_rec_static, &c->b_rr_pre, &c->b_ea_pool, &c->b_sh_pool, &c->b_tb,
                &c->b_xa, &c->b_xb, &c->b_proj, &c->b_tr, &c->b_rot, &c->b_tor, &c->b_pos, &c->b_step, &c->b_edge_total,
                &c->b_glist, &c->b_gcnt, &c->b_counters, &c->b_part, &c->b_hs, &c->b_need, &c->b_static_pos,
                &c->b_tc_scratch};
  for (Buf* b : all) free_buf(*b);
  for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
  if (c->w) cudaFree(c->w);
  if (c->w2s) cudaFree(c->w2s);
  if (c->ltab) cudaFree(c->ltab);
  if (c->tc_rows) cudaFree(c->tc_rows);
  if (c->tcr_roles) cudaFree(c->tcr_roles);
  if (c->w2r) cudaFree(c->w2r);
  delete c;
  return DDK_OK;
}

int ddk_set_batch(DdkCtx* c, const DdkBatch* b, void* stream) {
  if (!c || !b) return DDK_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  DDK_CUDA_TRY(c, cudaSetDevice(c->device));
  c->has_batch = false;
  if (b->B <= 0 || b->NL <= 0 || b->NR <= 0) return fail(c, DDK_ERR_INVALID, "empty batch");
  if (!b->lig_ptr_h || !b->rec_ptr_h || !b->bond_ptr_h || !b->rec_edge_ptr_h || !b->lig_x || !b->rec_x || !b->rec_pos)
    return fail(c, DDK_ERR_INVALID, "missing batch arrays");
  if (b->EB > 0 && (!b->bond_index_h || !b->edge_mask_h || !b->bond_attr)) return fail(c, DDK_ERR_INVALID, "missing bond arrays");
  if (b->ER > 0 && !b->rec_index_h) return fail(c, DDK_ERR_INVALID, "missing receptor edges");
  if (c->cfg.latent_dim > 0 && (!b->lig_latent || !b->rec_latent)) return fail(c, DDK_ERR_INVALID, "latents required");
  if (b->lig_ptr_h[b->B] != b->NL || b->rec_ptr_h[b->B] != b->NR || b->bond_ptr_h[b->B] != b->EB ||
      b->rec_edge_ptr_h[b->B] != b->ER)
    return fail(c, DDK_ERR_INVALID, "ptr arrays do not match the totals");
  const int B = b->B, NL = b->NL, NR = b->NR, EB = b->EB, ER = b->ER;
  static const bool timing = getenv("DDK_TIMING") != nullptr;   // host phases of this call on stderr
  const auto tm0 = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (timing) fprintf(stderr, "[ddk_set_batch] %-10s %8.3f ms\n", what,
                        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tm0).count());
  };
  c->B = B; c->NL = NL; c->NR = NR; c->EB = EB; c->ER = ER; c->N = NL + NR;
  c->rec_pos = b->rec_pos; c->mask_rotate = b->mask_rotate; c->bond_attr = b->bond_attr;
  c->lig_latent = b->lig_latent; c->rec_latent = b->rec_latent;
  c->lig_uncond = b->lig_uncond; c->rec_uncond = b->rec_uncond;

  std::vector<int> lig_ptr(b->lig_ptr_h, b->lig_ptr_h + B + 1), rec_ptr(b->rec_ptr_h, b->rec_ptr_h + B + 1);
  std::vector<int> lig_graph(NL), rec_graph(NR);
  std::vector<int64_t> ll_off(B), lr_off(B);
  int64_t LL = 0, LR = 0;
  c->maxNl = 0; c->maxNr = 0;
  for (int g = 0; g < B; ++g) {
    int nl = lig_ptr[g + 1] - lig_ptr[g], nr = rec_ptr[g + 1] - rec_ptr[g];
    if (nl <= 0 || nr <= 0) return fail(c, DDK_ERR_INVALID, "graph without ligand atoms or residues");
    for (int n = lig_ptr[g]; n < lig_ptr[g + 1]; ++n) lig_graph[n] = g;
    for (int n = rec_ptr[g]; n < rec_ptr[g + 1]; ++n) rec_graph[n] = g;
    ll_off[g] = LL; lr_off[g] = LR;
    LL += (int64_t)nl * nl; LR += (int64_t)nl * nr;
    c->maxNl = std::max(c->maxNl, nl); c->maxNr = std::max(c->maxNr, nr);
  }
  c->LLtot = LL; c->LRtot = LR;
  c->slot_ll = EB; c->slot_lr = EB + LL; c->slot_rr = EB + LL + LR; c->P = c->slot_rr + ER;
  if (c->P >= ((int64_t)1 << 31)) return fail(c, DDK_ERR_INVALID, "batch too large for 32-bit edge slots");

  // bonds, rotatable bonds.  The edge arrays of the caller are read in place (they are uploaded from there too): at 400
  // poses the receptor contacts alone are 2.9 M edges, and copies / separate passes showed up in the end-to-end time.
  const int* bond_src = b->bond_index_h; const int* bond_dst = b->bond_index_h + EB;
  const int* rr_src = b->rec_index_h;    const int* rr_dst = b->rec_index_h + ER;
  std::vector<int> rot_u, rot_v, rot_graph, rot_ptr(B + 1, 0);
  const int nsegs = 2 * (NL + NR);
  std::vector<int> seg_static(nsegs, 0), seg_base(nsegs + 1, 0);
  std::vector<int> static_pos((size_t)EB + ER);       // position of every static edge inside its segment (edge order)
  for (int g = 0; g < B; ++g) {
    const int l0 = lig_ptr[g], l1 = lig_ptr[g + 1];
    for (int e = b->bond_ptr_h[g]; e < b->bond_ptr_h[g + 1]; ++e) {
      const int u = bond_src[e], v = bond_dst[e];
      if (u < l0 || u >= l1 || v < l0 || v >= l1) return fail(c, DDK_ERR_INVALID, "bond crosses a graph boundary");
      if (b->edge_mask_h[e]) { rot_u.push_back(u); rot_v.push_back(v); rot_graph.push_back(g); }
      static_pos[e] = seg_static[2 * u]++;
    }
    rot_ptr[g + 1] = (int)rot_u.size();
  }
  c->RB = (int)rot_u.size();
  if (b->RB != c->RB) return fail(c, DDK_ERR_INVALID, "RB does not match edge_mask");
  if (c->RB > 0 && !c->cfg.no_torsion && (!b->mask_rotate || !b->mask_rotate_off_h))
    return fail(c, DDK_ERR_INVALID, "mask_rotate required");
  std::vector<int64_t> mr_off(B, 0);
  if (b->mask_rotate_off_h) mr_off.assign(b->mask_rotate_off_h, b->mask_rotate_off_h + B);
  for (int g = 0; g < B; ++g) {
    const int r0 = rec_ptr[g], r1 = rec_ptr[g + 1];
    const int e0 = b->rec_edge_ptr_h[g], e1 = b->rec_edge_ptr_h[g + 1];
    if (e0 < 0 || e1 < e0 || e1 > ER) return fail(c, DDK_ERR_INVALID, "bad receptor edge offsets");
    for (int e = e0; e < e1; ++e) {
      const int u = rr_src[e], v = rr_dst[e];
      if (u < r0 || u >= r1 || v < r0 || v >= r1) return fail(c, DDK_ERR_INVALID, "receptor edge crosses a graph boundary");
      static_pos[(size_t)EB + e] = seg_static[2 * (NL + u)]++;
    }
  }

  // segments: (node, group) -> list of (edge slot, destination node); capacity = static edges + every possible dynamic one
  {
    int64_t total = 0;
    for (int g = 0; g < B; ++g) {
      const int nl = lig_ptr[g + 1] - lig_ptr[g], nr = rec_ptr[g + 1] - rec_ptr[g];
      for (int n = lig_ptr[g]; n < lig_ptr[g + 1]; ++n) {
        seg_base[2 * n] = (int)total; total += seg_static[2 * n] + nl - 1;
        seg_base[2 * n + 1] = (int)total; total += nr;
      }
    }
    for (int g = 0; g < B; ++g) {
      const int nl = lig_ptr[g + 1] - lig_ptr[g];
      for (int r = rec_ptr[g]; r < rec_ptr[g + 1]; ++r) {
        seg_base[2 * (NL + r)] = (int)total; total += seg_static[2 * (NL + r)];
        seg_base[2 * (NL + r) + 1] = (int)total; total += nl;
      }
    }
    if (total >= ((int64_t)1 << 31)) return fail(c, DDK_ERR_INVALID, "batch too large for 32-bit list offsets");
    c->list_total = total;
  }
  // static list entries (covalent bonds, receptor contacts) sit at the head of their segment, in edge order; only their
  // positions travel to the device (k_fill_static_lists writes them), not the whole capacity-sized list
  for (int e = 0; e < EB; ++e) static_pos[e] += seg_base[2 * bond_src[e]];
  for (int e = 0; e < ER; ++e) static_pos[(size_t)EB + e] += seg_base[2 * (NL + rr_src[e])];
  const int64_t total = c->list_total;
  lap("host lists");
  int rc;
#define UP(buf, vec) if ((rc = upload(c, buf, vec, st)) != DDK_OK) return rc
  UP(c->b_lig_ptr, lig_ptr); UP(c->b_rec_ptr, rec_ptr); UP(c->b_lig_graph, lig_graph); UP(c->b_rec_graph, rec_graph);
#define UPP(buf, p_, n_)                                                                                   \
  if ((rc = ensure(c, buf, (size_t)(n_) * sizeof(int))) != DDK_OK) return rc;                              \
  if ((n_) > 0) DDK_CUDA_TRY(c, cudaMemcpyAsync(buf.p, p_, (size_t)(n_) * sizeof(int), cudaMemcpyHostToDevice, st))
  UPP(c->b_bond_src, bond_src, EB); UPP(c->b_bond_dst, bond_dst, EB); UPP(c->b_rr_src, rr_src, ER); UPP(c->b_rr_dst, rr_dst, ER);
#undef UPP
  UP(c->b_rot_u, rot_u); UP(c->b_rot_v, rot_v); UP(c->b_rot_ptr, rot_ptr); UP(c->b_rot_graph, rot_graph); UP(c->b_mr_off, mr_off);
  UP(c->b_ll_off, ll_off); UP(c->b_lr_off, lr_off);
  seg_base.resize(nsegs);
  UP(c->b_seg_base, seg_base); UP(c->b_seg_static, seg_static); UP(c->b_seg_cnt, seg_static); UP(c->b_static_pos, static_pos);
  if ((rc = ensure(c, c->b_seg_list, (size_t)std::max<int64_t>(total, 1) * sizeof(int2))) != DDK_OK) return rc;
#undef UP
#define EN(buf, bytes) if ((rc = ensure(c, buf, (size_t)(bytes))) != DDK_OK) return rc
  EN(c->b_lig_static, (size_t)NL * NS * 4); EN(c->b_rec_static, (size_t)NR * NS * 4); EN(c->b_rr_pre, (size_t)ER * EA * 4);
  EN(c->b_ea_pool, (size_t)c->P * EA * 4); EN(c->b_sh_pool, (size_t)c->P * 16);
  EN(c->b_tb, (size_t)B * TB_COUNT * NS * 4);
  EN(c->b_xa, (size_t)c->N * D * 4); EN(c->b_xb, (size_t)c->N * D * 4); EN(c->b_proj, (size_t)c->N * 4 * HID * 4);
  c->nhop = std::min(F3_MAXHOP, c->cfg.num_conv_layers - 1);
  // work lists | needed-hop lists | pieces of the lig<-rec segments (k_conv_tcr)
  EN(c->b_glist, ((size_t)nsegs + (size_t)c->nhop * NR + (size_t)NL * TCR_PMAX) * 16); EN(c->b_gcnt, (2 * F3_NLIST + 4) * 4); EN(c->b_counters, 5 * NSL_MAX * 4);
  EN(c->b_need, (size_t)std::max(1, c->nhop) * NR);
  static_assert(TCR_MAXROLES * TCR_PS <= NSL_MAX * D, "the partial records of k_conv_tcr fit the partial rows of k_conv_fused");
  EN(c->b_part, ((size_t)nsegs + (size_t)NL * (TCR_PMAX - 1)) * NSL_MAX * D * 4);
  EN(c->b_hs, (size_t)std::max<int64_t>(total, 1) * HID * 4);
  c->tc_cap = 0;
  if (conv_path() == 1) {   // A_s scratch of the round-1 tensor-core path (k_acc_tc): one block per ligand atom, capped at 4 GB
    const size_t per = tc_scratch_floats_per_segment() * sizeof(float);
    c->tc_cap = (int)std::min<size_t>((size_t)NL, ((size_t)4 << 30) / per);
    if (ensure(c, c->b_tc_scratch, (size_t)c->tc_cap * per) != DDK_OK) {   // optional: without the scratch every segment
      c->tc_cap = 0;                                                       // stays on the FFMA2 path
      c->err.clear();
      cudaGetLastError();
    }
  }
  EN(c->b_tr, (size_t)B * 3 * 4); EN(c->b_rot, (size_t)B * 3 * 4); EN(c->b_tor, (size_t)std::max(c->RB, 1) * 4);
#undef EN
  lap("uploads");
  launch_setup(c, b, b->lig_x, b->rec_x, st);
  DDK_CUDA_TRY(c, cudaGetLastError());
  DDK_CUDA_TRY(c, cudaStreamSynchronize(st));   // host vectors above go out of scope
  lap("setup+sync");
  c->has_batch = true;
  c->x_final = nullptr;
  return DDK_OK;
}

// heads_only: the caller reads only ligand node features afterwards (score heads), so the last conv layer skips the
// segments of receptor nodes (edge groups 2, 3) and the layer before it those of residues without a cross edge (the
// last layer reads receptor features only through cross edges); ddk_embed needs every node
static int run_embed(DdkCtx* c, const float* lig_pos, const DdkStepInputs* in, cudaStream_t st, bool heads_only) {
  launch_step_consts(c, in->sigma_emb, st);
  launch_build_lists(c, lig_pos, in->cross_cutoff, st);
  launch_edge_features(c, lig_pos, st);
  launch_build_group_lists(c, st, heads_only);
  float* xa = ptr<float>(c->b_xa);
  float* xb = ptr<float>(c->b_xb);
  launch_node_proj(c, 0, nullptr, xa, st);
  float* xin = xa; float* xout = xb;
  for (int l = 0; l < c->cfg.num_conv_layers; ++l) {
    const int L = c->cfg.num_conv_layers;
    launch_conv_layer(c, l, xin, xout, st, !heads_only ? CONV_ALL : (l == L - 1 ? CONV_LIG : CONV_NEEDED + (L - 2 - l)));
    if (l + 1 < c->cfg.num_conv_layers) launch_node_proj(c, l + 1, xout, nullptr, st);
    std::swap(xin, xout);
  }
  c->x_final = xin;
  c->x_final_all = !heads_only;
  return DDK_OK;
}

static int check_step(DdkCtx* c, const float* lig_pos, const DdkStepInputs* in) {
  if (!c || !lig_pos || !in) return DDK_ERR_INVALID;
  if (!c->has_batch) return fail(c, DDK_ERR_STATE, "ddk_set_batch has not been called");
  if (!in->sigma_emb || !in->tr_sigma || !in->rot_scale || (c->cfg.dynamic_max_cross && !in->cross_cutoff))
    return fail(c, DDK_ERR_INVALID, "missing step inputs");
  if (c->RB > 0 && !c->cfg.no_torsion && !in->tor_scale) return fail(c, DDK_ERR_INVALID, "missing tor_scale");
  return DDK_OK;
}

int ddk_embed(DdkCtx* c, const float* lig_pos, const DdkStepInputs* in, void* stream) {
  int rc = check_step(c, lig_pos, in);
  if (rc) return rc;
  DDK_CUDA_TRY(c, cudaSetDevice(c->device));
  run_embed(c, lig_pos, in, (cudaStream_t)stream, false);
  DDK_CUDA_TRY(c, cudaGetLastError());
  return DDK_OK;
}

int ddk_score(DdkCtx* c, const float* lig_pos, const DdkStepInputs* in, float* tr, float* rot, float* tor, void* stream) {
  int rc = check_step(c, lig_pos, in);
  if (rc) return rc;
  if (!tr || !rot) return fail(c, DDK_ERR_INVALID, "null output");
  cudaStream_t st = (cudaStream_t)stream;
  DDK_CUDA_TRY(c, cudaSetDevice(c->device));
  run_embed(c, lig_pos, in, st, true);
  launch_head_trrot(c, lig_pos, c->x_final, in, tr, rot, st);
  if (tor) launch_head_tor(c, lig_pos, c->x_final, in, tor, st);
  DDK_CUDA_TRY(c, cudaGetLastError());
  return DDK_OK;
}

int ddk_get_node_features(DdkCtx* c, float* lig_out, float* rec_out, void* stream) {
  if (!c || !c->has_batch || !c->x_final) return fail(c, DDK_ERR_STATE, "no embedding available");
  if (rec_out && !c->x_final_all) return fail(c, DDK_ERR_STATE, "receptor features are only complete after ddk_embed (ddk_score skips them in the last layer)");
  cudaStream_t st = (cudaStream_t)stream;
  if (lig_out) DDK_CUDA_TRY(c, cudaMemcpyAsync(lig_out, c->x_final, (size_t)c->NL * D * 4, cudaMemcpyDeviceToDevice, st));
  if (rec_out)
    DDK_CUDA_TRY(c, cudaMemcpyAsync(rec_out, c->x_final + (size_t)c->NL * D, (size_t)c->NR * D * 4, cudaMemcpyDeviceToDevice, st));
  return DDK_OK;
}

int ddk_update(DdkCtx* c, float* lig_pos, const float* tr, const float* rot, const float* tor, const float* z_tr,
               const float* z_rot, const float* z_tor, const DdkStepCoef* coef_h, void* stream) {
  if (!c || !lig_pos || !tr || !rot || !coef_h) return DDK_ERR_INVALID;
  if (!c->has_batch) return fail(c, DDK_ERR_STATE, "ddk_set_batch has not been called");
  DDK_CUDA_TRY(c, cudaSetDevice(c->device));
  launch_update(c, lig_pos, tr, rot, tor, z_tr, z_rot, z_tor, *coef_h, (cudaStream_t)stream);
  DDK_CUDA_TRY(c, cudaGetLastError());
  return DDK_OK;
}

int ddk_sample(DdkCtx* c, float* lig_pos, int32_t n_steps, const DdkStepInputs* si, const float* z_tr, const float* z_rot,
               const float* z_tor, const DdkStepCoef* coef_h, void* stream) {
  int rc = check_step(c, lig_pos, si);
  if (rc) return rc;
  if (n_steps <= 0 || !coef_h) return fail(c, DDK_ERR_INVALID, "bad step count / coefficients");
  cudaStream_t st = (cudaStream_t)stream;
  DDK_CUDA_TRY(c, cudaSetDevice(c->device));
  float* tr = ptr<float>(c->b_tr);
  float* rot = ptr<float>(c->b_rot);
  float* tor = (c->RB > 0 && !c->cfg.no_torsion) ? ptr<float>(c->b_tor) : nullptr;
  for (int s = 0; s < n_steps; ++s) {
    DdkStepInputs in;
    in.sigma_emb = si->sigma_emb + (size_t)s * c->B * SE;
    in.cross_cutoff = si->cross_cutoff ? si->cross_cutoff + (size_t)s * c->B : nullptr;
    in.tr_sigma = si->tr_sigma + (size_t)s * c->B;
    in.rot_scale = si->rot_scale + (size_t)s * c->B;
    in.tor_scale = si->tor_scale ? si->tor_scale + (size_t)s * c->B : nullptr;
    run_embed(c, lig_pos, &in, st, true);
    launch_head_trrot(c, lig_pos, c->x_final, &in, tr, rot, st);
    if (tor) launch_head_tor(c, lig_pos, c->x_final, &in, tor, st);
    launch_update(c, lig_pos, tr, rot, tor, z_tr ? z_tr + (size_t)s * c->B * 3 : nullptr,
                  z_rot ? z_rot + (size_t)s * c->B * 3 : nullptr, (z_tor && tor) ? z_tor + (size_t)s * c->RB : nullptr,
                  coef_h[s], st);
  }
  DDK_CUDA_TRY(c, cudaGetLastError());
  return DDK_OK;
}

int ddk_sample_host(DdkCtx* c, float* lig_pos_h, int32_t n_steps, const DdkStepInputs* si_h, const float* z_tr_h,
                    const float* z_rot_h, const float* z_tor_h, const DdkStepCoef* coef_h) {
  if (!c || !lig_pos_h || !si_h) return DDK_ERR_INVALID;
  if (!c->has_batch) return fail(c, DDK_ERR_STATE, "ddk_set_batch has not been called");
  if (n_steps <= 0) return fail(c, DDK_ERR_INVALID, "bad step count");
  DDK_CUDA_TRY(c, cudaSetDevice(c->device));
  cudaStream_t st = 0;
  const size_t B = c->B, S = n_steps;
  const size_t n_emb = S * B * SE, n_b = S * B, n_z3 = S * B * 3, n_zt = S * (size_t)c->RB, n_pos = (size_t)c->NL * 3;
  size_t total = n_emb + 4 * n_b + 2 * n_z3 + n_zt + n_pos;
  int rc = ensure(c, c->b_step, total * sizeof(float));
  if (rc) return rc;
  float* base = ptr<float>(c->b_step);
  float* d_emb = base; float* d_cut = d_emb + n_emb; float* d_trs = d_cut + n_b; float* d_rot = d_trs + n_b;
  float* d_tor = d_rot + n_b; float* d_ztr = d_tor + n_b; float* d_zrot = d_ztr + n_z3; float* d_ztor = d_zrot + n_z3;
  float* d_pos = d_ztor + n_zt;
  auto h2d = [&](float* dst, const float* src, size_t n) -> cudaError_t {
    if (!src || n == 0) return cudaSuccess;
    return cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyHostToDevice, st);
  };
  DDK_CUDA_TRY(c, h2d(d_emb, si_h->sigma_emb, n_emb));
  DDK_CUDA_TRY(c, h2d(d_cut, si_h->cross_cutoff, n_b));
  DDK_CUDA_TRY(c, h2d(d_trs, si_h->tr_sigma, n_b));
  DDK_CUDA_TRY(c, h2d(d_rot, si_h->rot_scale, n_b));
  DDK_CUDA_TRY(c, h2d(d_tor, si_h->tor_scale, n_b));
  DDK_CUDA_TRY(c, h2d(d_ztr, z_tr_h, n_z3));
  DDK_CUDA_TRY(c, h2d(d_zrot, z_rot_h, n_z3));
  DDK_CUDA_TRY(c, h2d(d_ztor, z_tor_h, n_zt));
  DDK_CUDA_TRY(c, h2d(d_pos, lig_pos_h, n_pos));
  DdkStepInputs si;
  si.sigma_emb = d_emb; si.cross_cutoff = si_h->cross_cutoff ? d_cut : nullptr; si.tr_sigma = d_trs; si.rot_scale = d_rot;
  si.tor_scale = si_h->tor_scale ? d_tor : nullptr;
  rc = ddk_sample(c, d_pos, n_steps, &si, z_tr_h ? d_ztr : nullptr, z_rot_h ? d_zrot : nullptr, z_tor_h ? d_ztor : nullptr,
                  coef_h, st);
  if (rc) return rc;
  DDK_CUDA_TRY(c, cudaMemcpyAsync(lig_pos_h, d_pos, n_pos * sizeof(float), cudaMemcpyDeviceToHost, st));
  DDK_CUDA_TRY(c, cudaStreamSynchronize(st));
  return DDK_OK;
}

int64_t ddk_kernel_launches(const DdkCtx* c) { return c ? c->launches : 0; }

int64_t ddk_last_edge_count(DdkCtx* c) {
  if (!c || !c->has_batch) return -1;
  cudaSetDevice(c->device);
  int nsegs = 2 * (c->NL + c->NR);
  std::vector<int> cnt(nsegs);
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  if (cudaMemcpy(cnt.data(), c->b_seg_cnt.p, nsegs * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  int64_t t = 0;
  for (int v : cnt) t += v;
  return t;
}

int64_t ddk_edge_total(DdkCtx* c) {
  if (!c) return -1;
  cudaSetDevice(c->device);
  unsigned long long v = 0;
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  if (cudaMemcpy(&v, c->b_edge_total.p, 8, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return (int64_t)v;
}

int64_t ddk_segment_total(DdkCtx* c) {
  if (!c) return -1;
  cudaSetDevice(c->device);
  unsigned long long v = 0;
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  if (cudaMemcpy(&v, ptr<unsigned long long>(c->b_edge_total) + 1, 8, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return (int64_t)v;
}

int ddk_group_totals(DdkCtx* c, int64_t* edges, int64_t* segments) {
  if (!c || !edges || !segments) return DDK_ERR_INVALID;
  cudaSetDevice(c->device);
  unsigned long long v[2 + 2 * F3_NLIST];
  DDK_CUDA_TRY(c, cudaDeviceSynchronize());
  DDK_CUDA_TRY(c, cudaMemcpy(v, c->b_edge_total.p, sizeof(v), cudaMemcpyDeviceToHost));
  static_assert(F3_NLIST == DDK_WORK_LISTS, "work list count");
  for (int g = 0; g < F3_NLIST; ++g) { edges[g] = (int64_t)v[2 + g]; segments[g] = (int64_t)v[2 + F3_NLIST + g]; }
  return DDK_OK;
}

int ddk_debug_read(DdkCtx* c, const char* name, void* dst_h, size_t max_bytes, size_t* n_bytes) {
  if (!c || !name || !c->has_batch) return DDK_ERR_INVALID;
  cudaSetDevice(c->device);
  const void* src = nullptr;
  size_t n = 0;
  std::string s(name);
  const int nsegs = 2 * (c->NL + c->NR);
  if (s == "x_final") { src = c->x_final; n = (size_t)c->N * D * 4; }
  else if (s == "xa") { src = c->b_xa.p; n = (size_t)c->N * D * 4; }
  else if (s == "xb") { src = c->b_xb.p; n = (size_t)c->N * D * 4; }
  else if (s == "proj") { src = c->b_proj.p; n = (size_t)c->N * 4 * HID * 4; }
  else if (s == "tb") { src = c->b_tb.p; n = (size_t)c->B * TB_COUNT * NS * 4; }
  else if (s == "seg_cnt") { src = c->b_seg_cnt.p; n = (size_t)nsegs * 4; }
  else if (s == "seg_base") { src = c->b_seg_base.p; n = (size_t)nsegs * 4; }
  else if (s == "seg_list") { src = c->b_seg_list.p; n = (size_t)c->list_total * 8; }
  else if (s == "ea_pool") { src = c->b_ea_pool.p; n = (size_t)c->P * EA * 4; }
  else if (s == "sh_pool") { src = c->b_sh_pool.p; n = (size_t)c->P * 16; }
  else if (s == "lig_static") { src = c->b_lig_static.p; n = (size_t)c->NL * NS * 4; }
  else if (s == "rec_static") { src = c->b_rec_static.p; n = (size_t)c->NR * NS * 4; }
  else if (s == "tr") { src = c->b_tr.p; n = (size_t)c->B * 12; }
  else if (s == "rot") { src = c->b_rot.p; n = (size_t)c->B * 12; }
  else if (s == "tor") { src = c->b_tor.p; n = (size_t)c->RB * 4; }
  else return fail(c, DDK_ERR_INVALID, "unknown debug buffer " + s);
  if (n_bytes) *n_bytes = n;
  if (!dst_h) return DDK_OK;
  if (!src) return fail(c, DDK_ERR_STATE, "buffer not available yet");
  DDK_CUDA_TRY(c, cudaDeviceSynchronize());
  DDK_CUDA_TRY(c, cudaMemcpy(dst_h, src, std::min(n, max_bytes), cudaMemcpyDeviceToHost));
  return DDK_OK;
}

int ddk_debug_set_tc(int32_t on) { return tc_set_override(on < 0 ? -1 : (on > 2 ? 2 : on)); }

int ddk_profile_enable(DdkCtx* c, int32_t on) {
  if (!c) return DDK_ERR_INVALID;
  c->prof = on != 0;
  return DDK_OK;
}

int ddk_profile_read(DdkCtx* c, double* ms, int64_t* launches) {
  if (!c || !ms || !launches) return DDK_ERR_INVALID;
  static_assert(PC_COUNT == DDK_PROFILE_CLASSES, "profile class count");
  cudaSetDevice(c->device);
  DDK_CUDA_TRY(c, cudaDeviceSynchronize());
  for (int i = 0; i < PC_COUNT; ++i) { ms[i] = 0.0; launches[i] = 0; }
  for (const ProfRec& r : c->prof_recs) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) { ms[r.cls] += t; launches[r.cls]++; }
  }
  c->prof_recs.clear();
  c->ev_used = 0;
  return DDK_OK;
}

int ddk_host_kabsch(const float* a_h, const float* b_h, int32_t n, float* R9_h, float* t3_h) {
  if (!a_h || !b_h || n <= 0 || !R9_h || !t3_h) return DDK_ERR_INVALID;
  host_kabsch(a_h, b_h, n, R9_h, t3_h);
  return DDK_OK;
}

int ddk_host_axis_angle_to_matrix(const float* aa, float* R9_h) {
  if (!aa || !R9_h) return DDK_ERR_INVALID;
  host_axis_angle(aa, R9_h);
  return DDK_OK;
}

int ddk_host_lane_tables_check(void) {
  LaneTab tab[32];
  for (int lv = 0; lv < 4; ++lv)
    if (!build_lane_table(lv, tab)) return 1 + lv;
  return 0;
}

// Host evaluation of the row table k_acc_tc works from (ddk_conv_tc.cu): basis_out[u] for one destination feature row x[84]
// and one harmonics record sh[4], u in kernel order.  Returns the number of rows of the level, or -1 if a row is missing /
// duplicated or the table is not sorted by row type (the kernel relies on that for branch-free warps).
int ddk_host_tc_rows_eval(int32_t lv, const float* x84, const float* sh4, float* basis_out) {
  if (lv < 0 || lv > 3 || !x84 || !sh4 || !basis_out) return -1;
  const int U = lv == 0 ? 96 : (lv == 1 ? 138 : (lv == 2 ? 180 : 276));
  std::vector<TcRow> rows(TC_MAXROWS);
  build_tc_rows(lv, rows.data());
  std::vector<int> seen(U, 0);
  int last_type = 0;
  for (int p = 0; p < TC_MAXROWS; ++p) {
    const TcRow& r = rows[p];
    if (r.u < 0) { if (p < U) return -1; continue; }
    if (p >= U || r.u >= U || seen[r.u]++ || r.type < last_type) return -1;
    last_type = r.type;
    const float* s = sh4;
    float v;
    if (r.type == 0) v = x84[r.i0] * s[r.m];
    else {
      const float v0 = x84[r.i0], v1 = x84[r.i0 + 1], v2 = x84[r.i0 + 2];
      if (r.type == 1) v = v0 * s[1] + v1 * s[2] + v2 * s[3];
      else if (r.m == 1) v = v1 * s[3] - v2 * s[2];
      else if (r.m == 2) v = v2 * s[1] - v0 * s[3];
      else v = v0 * s[2] - v1 * s[1];
    }
    basis_out[r.u] = v;
  }
  return U;
}

int ddk_host_tcr_roles_check(void) { return host_tcr_roles_check(); }

int ddk_host_tc_split_rn(const float* a_h, int32_t n, uint32_t* hi_h, uint32_t* lo_h) {
  if (!a_h || !hi_h || !lo_h || n < 0) return -1;
  for (int i = 0; i < n; ++i) host_tc_split_rn(a_h[i], hi_h + i, lo_h + i);
  return 0;
}

// Host build of the TF32 split k_acc_tc applies to both MMA operands: a = hi + lo exactly, hi on the TF32 grid.
int ddk_host_tc_split(const float* a_h, int32_t n, uint32_t* hi_h, uint32_t* lo_h) {
  if (!a_h || !hi_h || !lo_h || n < 0) return -1;
  for (int i = 0; i < n; ++i) host_tc_split(a_h[i], hi_h + i, lo_h + i);
  return 0;
}

}  // extern "C"
