// Internal declarations shared by the libddk translation units (not part of the C ABI).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/ddk.h"

namespace ddk {

constexpr int NS = 24;          // scalar multiplicity (ns)
constexpr int NV = 6;           // vector multiplicity (nv)
constexpr int D = 84;           // node feature width: [0e 24 | 1o 6x3 | 1e 6x3 | 0o 24]
constexpr int HID = 72;         // radial MLP width (3 ns)
constexpr int SE = 32;          // sigma embedding width
constexpr int DE = 32;          // distance embedding width
constexpr int ESM = 1280;
constexpr int REC_X = 1 + ESM;  // receptor row: amino-acid index + ESM embedding
constexpr int LIG_CAT = 16;
constexpr int EA = 24;          // edge embedding width (ns)

// per-graph per-step bias vectors derived from the sigma embedding (k_step_consts)
enum { TB_LIG_NODE = 0, TB_REC_NODE, TB_LIG_EDGE, TB_REC_EDGE, TB_CROSS_EDGE, TB_CENTER, TB_TR_FINAL, TB_ROT_FINAL, TB_COUNT };

struct ClassInfo {   // one irrep class of a conv layer's output (0e, 1o, 1e, 0o)
  int F, O, ncomp, uoff, col0;
  int64_t woff, boff;  // offsets inside the packed W2p / b2p of a group
};

struct LayerInfo {
  int lv;            // basis level = min(layer, 3)
  int U;             // basis size
  int NA;            // ceil(U / 32)
  int dout;          // valid output width (24, 42, 60, 84)
  ClassInfo cls[4];
  int ncls;
};

struct Buf {
  void* p = nullptr;
  size_t bytes = 0;
};

// ---- fused conv kernel (ddk_conv3.cu): hidden units are processed in slices of J (per basis level)
constexpr int NSL_MAX = 9;          // most slices of any level (72 / 8)
__host__ __device__ constexpr int f3_J(int lv) { return lv == 0 ? 24 : (lv == 3 ? 8 : 12); }   // hidden units per slice (as wide as shared memory and registers allow)
__host__ __device__ constexpr int f3_nsl(int lv) { return HID / f3_J(lv); }
constexpr int F3_ACC = 8;           // accumulate warp PAIRS per CTA (= segments per contraction batch)
constexpr int F3_CON = 4;           // contraction warps per CTA
constexpr int F3_THREADS = (2 * F3_ACC + F3_CON) * 32;
constexpr int KC3 = 8;              // edges per gather chunk of one accumulate warp

constexpr int F3_MAXSLOT = 9;
struct LaneTab {                    // what one lane of an accumulate warp evaluates at a basis level (k_conv_fused)
  int u[F3_MAXSLOT];                // row of the A block (kernel order of ddk_conv.cuh) owned in each slot, -1 = none
  int oS[2], oV[2], isvec[2];       // mixed slot groups (slots 0..3; 4..7 at level 3): column of the scalar source, first
                                    // column of the 3-vector source, and which of the two the lane evaluates
  int oV0;                          // slot V0: column of the source (x * sh[0])
  int gi[3], gm[3];                 // generic slot (level 2): fa x[ia] sh[ma] + fb x[ib] sh[mb] + fc x[ic] sh[mc]
  float gf[3];
};

// tensor-core accumulation of long lig<-rec segments (ddk_conv_tc.cu)
struct TcRow { short type, i0, m, u; };   // basis row: 0 = x[i0] * sh[m]; 1 = x[i0..i0+2] . sh[1..3]; 2 = component m-1 of x[i0..i0+2] x sh[1..3]
constexpr int TC_MAXROWS = 384;
constexpr int TC_MIN_CHUNKS = 8;            // segments with at least this many 8-edge chunks take the tensor-core path

// tensor-core conv with resident contraction (ddk_conv_tcr.cu): a ROLE = (irrep classes, range of hidden units) of a basis level
constexpr int TCR_MAXROLES = 5;     // roles per level
constexpr int TCR_MAXACC = 8;       // accumulator slots in tensor memory
constexpr int TCR_WMAX = 36240 + 192;     // floats of the largest resident weight slice (level 3, scalar classes, 24 hidden units + bias)
constexpr int TCR_SUB = 64;         // k_conv_tcr: a lig<-rec segment is accumulated in pieces of at most this many edges (short accumulation chains: the
                                    // tensor core adds with truncation); piece 0 writes the segment's record, piece p > 0 the
                                    // record 2 N + (TCR_PMAX - 1) * ligand node + p - 1
constexpr int TCR_PMAX = 16;        // pieces per segment at most (the last piece takes the rest)
constexpr int TCR_PS = 144;         // floats of the partial record of one (segment, role): 8 contraction warps x 18
constexpr int TCR_MAXSRC = 8;       // shares of one output column inside a record
struct alignas(16) TcrRole {
  int nrows;                        // tile rows in use (<= 128)
  int ncol;                         // contracted accumulator columns: nj hidden units + the ones column (Bsum)
  int N;                            // MMA N: ncol rounded up to 16
  int j0, nj;                       // hidden units [j0, j0 + nj)
  int O;                            // outputs per row: 6 (vector classes) or 24 (scalar classes)
  int isS;                          // scalar-class role
  int wstride, wfloats;             // floats of one (class, f) weight block incl. padding; floats of the whole slice
  int pad_[3];
  // Tile rows.  Vector roles: the three components of a basis row (class, f) in neighbouring rows, every class starting at a
  // multiple of 32 rows (one contraction warp never mixes classes).  Scalar roles: distinct row d = 16 q + l sits in tile rows
  // 32 q + l and 32 q + 16 + l (both halves of a warp), every class starting at a multiple of 16 distinct rows.
  TcRow rows[128];                  // basis row evaluated by row thread p (u = -1: padding)
  int woff[128];                    // offset of row p's weight block inside the slice
  // Partial record of a (segment, role), TCR_PS floats, written by the contraction warps (cw = 4 * set + quarter):
  //   vector roles: [cw][lane 0..2][6 outputs]  -- lane l holds the sum over the warp's rows in lanes = l (mod 3);
  //   scalar roles: [cw][lane half h][6]        -- outputs 12 set + 6 h + (0..5) summed over the 16 rows of the quarter.
  // fsrc[f]: the record entries that add up to node-feature column f (ascending, -1 terminated), none if another role owns f.
  short fsrc[D][TCR_MAXSRC];
};
static_assert(sizeof(TcrRole) % 16 == 0, "copied in 16-byte pieces");

struct ConSplit {                   // rows [f0, f1) of each irrep class handled by each contraction warp
  int f0[F3_CON][4], f1[F3_CON][4];
};

// kernel classes for the optional per-launch CUDA-event timing (ddk_profile_*)
enum ProfClass { PC_SETUP = 0, PC_GRAPH, PC_PROJ, PC_ACC0, PC_ACC1, PC_ACC2, PC_ACC3, PC_CONTRACT, PC_HEADS, PC_UPDATE, PC_HIDDEN, PC_TC0, PC_TC1, PC_TC2, PC_TC3, PC_COUNT };

struct ProfRec {
  int cls;
  cudaEvent_t a, b;
};

}  // namespace ddk

struct DdkCtx {
  DdkConfig cfg;
  int device = 0;
  std::string err;
  float* w = nullptr;                 // device weight blob
  std::vector<int64_t> off;           // offsets (floats)
  int64_t launches = 0;
  std::vector<ddk::LayerInfo> layers;
  float r2_lig = 25.f, r2_cross = 6400.f;

  // ---- batch (static) ----
  bool has_batch = false;
  int B = 0, NL = 0, NR = 0, EB = 0, ER = 0, RB = 0, N = 0;
  int maxNl = 0, maxNr = 0, maxRot = 0;
  int64_t LLtot = 0, LRtot = 0, P = 0, list_total = 0;
  int64_t slot_ll = 0, slot_lr = 0, slot_rr = 0;
  const float* rec_pos = nullptr;     // caller memory (must stay alive while the batch is current)
  const uint8_t* mask_rotate = nullptr;
  const float* lig_latent = nullptr; const float* rec_latent = nullptr;
  const float* lig_uncond = nullptr; const float* rec_uncond = nullptr;
  const float* bond_attr = nullptr;

  // device arrays owned by the context (grow-only)
  ddk::Buf b_lig_ptr, b_rec_ptr, b_lig_graph, b_rec_graph, b_bond_src, b_bond_dst, b_rr_src, b_rr_dst;
  ddk::Buf b_rot_u, b_rot_v, b_rot_ptr, b_rot_graph, b_mr_off, b_ll_off, b_lr_off;
  ddk::Buf b_seg_base, b_seg_static, b_seg_cnt, b_seg_list, b_static_pos;
  ddk::Buf b_lig_static, b_rec_static, b_rr_pre, b_ea_pool, b_sh_pool, b_tb;
  ddk::Buf b_xa, b_xb, b_proj;
  ddk::Buf b_tr, b_rot, b_tor, b_pos;
  ddk::Buf b_step;                    // staging for ddk_sample_host
  ddk::Buf b_edge_total;              // device uint64: edges of every combined graph built so far
  float* x_final = nullptr;           // points into xa or xb after the last conv layer
  bool x_final_all = false;           // receptor rows of x_final are valid (ddk_embed); ddk_score / ddk_sample skip them

  int sm_count = 148;
  float* w2s = nullptr;               // second-layer weights re-sliced per hidden-unit slice: [layer][group][72 / J][W * J]
  std::vector<int64_t> w2s_off;       // [layer * 4 + group] (floats)
  ddk::LaneTab* ltab = nullptr;       // [4 basis levels][32 lanes]: lane -> basis rows / sources of the fused conv kernel
  std::vector<ddk::ConSplit> con_split;   // per layer
  ddk::Buf b_glist, b_gcnt, b_counters, b_part;
  ddk::Buf b_need;                    // [hops][NR] uint8: receptor nodes whose features are read downstream (see ConvMode)
  int nhop = 0;                       // hops computed per step = num_conv_layers - 1
  ddk::Buf b_hs;                      // [72 / J][list_total][J]: hidden units of every listed edge of the current layer
  ddk::TcRow* tc_rows = nullptr;      // [4 levels][TC_MAXROWS]
  ddk::TcrRole* tcr_roles = nullptr;  // [4 levels][TCR_MAXROLES] (device)
  int tcr_nroles[4] = {0, 0, 0, 0};
  float* w2r = nullptr;               // weight slices of k_conv_tcr: [layer][group][role]
  std::vector<int64_t> w2r_off;       // [(layer * 4 + group) * TCR_MAXROLES + role] (floats)
  ddk::Buf b_tc_scratch;              // [segment][slice][U][J + 1]: A_s blocks of the tensor-core path
  int tc_cap = 0;                     // segments the scratch holds (0: tensor-core path off)

  // optional profiling (off by default)
  bool prof = false;
  std::vector<ddk::ProfRec> prof_recs;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
};

namespace ddk {

inline const float* W(const DdkCtx* c, int id) { return c->w + c->off[id]; }
inline int conv_id(int layer, int k) { return DDK_W_CONV_BASE + layer * DDK_W_CONV_STRIDE + k; }

template <typename T>
inline T* ptr(const Buf& b) { return reinterpret_cast<T*>(b.p); }

// RAII scope around one kernel launch: counts it and, when profiling is on, brackets it with CUDA events on `st`.
struct LaunchScope {
  DdkCtx* c; cudaStream_t st; int idx;
  LaunchScope(DdkCtx* ctx, int cls, cudaStream_t stream) : c(ctx), st(stream), idx(-1) {
    c->launches++;
    if (!c->prof) return;
    auto get = [&]() {
      if (c->ev_used == c->ev_pool.size()) { cudaEvent_t e; cudaEventCreate(&e); c->ev_pool.push_back(e); }
      return c->ev_pool[c->ev_used++];
    };
    ProfRec r{cls, get(), get()};
    cudaEventRecord(r.a, st);
    idx = (int)c->prof_recs.size();
    c->prof_recs.push_back(r);
  }
  ~LaunchScope() { if (idx >= 0) cudaEventRecord(c->prof_recs[idx].b, st); }
};

cudaError_t conv3_configure();
bool build_lane_table(int lv, LaneTab* tab32);
void build_con_split(const LayerInfo& li, ConSplit& sp);
void launch_build_group_lists(DdkCtx* c, cudaStream_t st, bool with_needed);
// Which segments a conv layer processes.  Before the score heads only ligand features are read, so working backwards from
// the last layer the set of receptor nodes whose features matter shrinks: the last layer needs ligand nodes only
// (CONV_LIG); the layer before it additionally the residues with a cross edge (hop 0); each earlier layer the previous set
// plus its receptor-contact neighbours (hop h).  CONV_NEEDED + h selects work list 4 + h for edge group 2.
enum ConvMode { CONV_LIG = -1, CONV_ALL = 0, CONV_NEEDED = 1 };
constexpr int F3_MAXHOP = 7;        // conv layers 0 .. L-2 <-> hops L-2 .. 0 (num_conv_layers <= 8)
constexpr int F3_NLIST = 4 + F3_MAXHOP;
void launch_conv_fused(DdkCtx* c, int layer, const float* x_in, float* x_out, cudaStream_t st, int mode);
void launch_edge_hidden(DdkCtx* c, int layer, cudaStream_t st, int mode);
cudaError_t heads_configure();
bool tc_enabled();                                   // DDK_TC=0 switches the tensor-core path off
int tc_set_override(int on);                         // run-time override of DDK_TC (-1: environment); returns the previous value
cudaError_t conv_tc_configure();
size_t tc_scratch_floats_per_segment();
void build_tc_rows(int lv, TcRow* rows);
void host_tc_split(float a, uint32_t* hi, uint32_t* lo);   // host build of the TF32 hi / lo split of k_acc_tc (tests)
void launch_acc_tc(DdkCtx* c, int layer, const float* x_in, cudaStream_t st);
int conv_path();                                     // 0: FFMA2 only, 1: FFMA2 + k_acc_tc for the long segments, 2: k_conv_tcr
cudaError_t conv_tcr_configure();
int build_tcr_roles(int lv, const LayerInfo& li, TcrRole* roles);
void build_tcr_weights(const LayerInfo& li, const TcrRole& R, const float* w2p, const float* b2p, float* out);
void launch_conv_tcr(DdkCtx* c, int layer, const float* x_in, float* x_out, cudaStream_t st, int mode);
void launch_conv_finalize(DdkCtx* c, int layer, const float* x_in, float* x_out, cudaStream_t st, bool lig_only, int nsl);
int host_tcr_roles_check();
int tcr_split_sub();                       // edges per piece of a lig<-rec segment (0: segments are not split; only k_conv_tcr splits)
int glist_split_off(const DdkCtx* c);      // region of b_glist behind the work lists and the needed-hop lists
int glist_off_group1(const DdkCtx* c);     // where work list 1 lives in b_glist (its own region while it is listed in pieces)
void host_tc_split_rn(float a, uint32_t* hi, uint32_t* lo);
void host_kabsch(const float* A, const float* Bp, int N, float* R9, float* t3);   // host build of the device routine (tests)
void host_axis_angle(const float* aa, float* R9);   // opt-in dynamic shared memory sizes (once per process / device)

// launchers (each enqueues on `st` and bumps ctx->launches)
void launch_setup(DdkCtx* c, const DdkBatch* b, const int32_t* lig_x, const float* rec_x, cudaStream_t st);
void launch_step_consts(DdkCtx* c, const float* sigma_emb, cudaStream_t st);
void launch_build_lists(DdkCtx* c, const float* lig_pos, const float* cutoff, cudaStream_t st);
void launch_edge_features(DdkCtx* c, const float* lig_pos, cudaStream_t st);
void launch_node_proj(DdkCtx* c, int layer, const float* x_in, float* x0_out, cudaStream_t st);
void launch_conv_layer(DdkCtx* c, int layer, const float* x_in, float* x_out, cudaStream_t st, int mode = 0);
void launch_head_trrot(DdkCtx* c, const float* lig_pos, const float* x, const DdkStepInputs* in, float* tr, float* rot,
                       cudaStream_t st);
void launch_head_tor(DdkCtx* c, const float* lig_pos, const float* x, const DdkStepInputs* in, float* tor, cudaStream_t st);
void launch_update(DdkCtx* c, float* lig_pos, const float* tr, const float* rot, const float* tor, const float* z_tr,
                   const float* z_rot, const float* z_tor, DdkStepCoef coef, cudaStream_t st);

}  // namespace ddk
