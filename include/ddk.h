/*
 * libddk -- C ABI of the B200-native reverse-diffusion docking sampler.
 *
 * Drop-in boundary for the hot path of gcorso/disco-diffdock.  Every entry point replaces a
 * piece of the reference's Python surface; file:line citations are into /root/reference:
 *
 *   ddk_create / ddk_destroy   <- TensorProductScoreModel.__init__ + load_state_dict
 *                                 (models/score_model.py:14-167, utils/model_utils.py:24-68, evaluate.py:160-174)
 *   ddk_set_batch              <- the step-invariant part of TensorProductScoreModel.embed(): atom / residue
 *                                 encoders, receptor graph geometry (models/score_model.py:191-199, 346-373;
 *                                 models/layers.py:140-149) for one PyG-style batch (utils/sampling.py:56-67)
 *   ddk_score                  <- TensorProductScoreModel.forward(data) -> (tr_pred, rot_pred, tor_pred)
 *                                 (models/score_model.py:259-308; called at utils/sampling.py:116)
 *   ddk_get_node_features      <- TensorProductScoreModel.embed(data)[:2] (models/score_model.py:169-257;
 *                                 called by models/pretrained_score_encoder.py:62)
 *   ddk_update                 <- the perturbation arithmetic + modify_conformer_batch
 *                                 (utils/sampling.py:137-198, utils/diffusion_utils.py:37-55, utils/torsion.py:71-86,
 *                                 utils/geometry.py:38-85, 126-156)
 *   ddk_sample                 <- the whole reverse-diffusion loop body of sampling() for one batch
 *                                 (utils/sampling.py:105-198), stream-ordered, no host synchronisation
 *   ddk_sample_host            <- same, HOST buffers in / out (the copies utils/sampling.py:67, 200-203 imply)
 *
 * Conventions
 *   - plain pointers and sizes only; all functions return 0 on success or a negative DdkStatus; the message of the
 *     last failure is available from ddk_last_error().  No C++ exception crosses the boundary.
 *   - pointers named *_h are HOST memory, everything else is DEVICE memory on the context's device.  The library
 *     never frees caller memory; it owns only the workspace it allocates in ddk_create / ddk_set_batch.
 *   - all work is enqueued on the cudaStream_t passed as `void* stream` (0 = legacy default stream); nothing
 *     synchronises the host except ddk_create, ddk_set_batch, ddk_sample_host and ddk_debug_*.
 *   - one context per device, one host thread per context (one process per GPU under torchrun).
 *   - floating point data is fp32, indices are int32.
 */
#ifndef DDK_H_
#define DDK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DDK_ABI_VERSION 1

typedef enum DdkStatus {
  DDK_OK = 0,
  DDK_ERR_INVALID = -1,     /* bad argument / unsupported configuration */
  DDK_ERR_CUDA = -2,        /* a CUDA runtime call failed */
  DDK_ERR_STATE = -3,       /* call order violated (e.g. ddk_score before ddk_set_batch) */
  DDK_ERR_NOMEM = -4
} DdkStatus;

/* Hyper-parameters (model_parameters.yml of the shipped checkpoints + constructor defaults). */
typedef struct DdkConfig {
  int32_t abi_version;        /* DDK_ABI_VERSION */
  int32_t ns, nv;             /* must be 24, 6 */
  int32_t num_conv_layers;    /* 3..8 (the heads consume the full 0e+1o+1e+0o representation) */
  int32_t latent_dim;         /* 0, or the number of equivariant latents (DisCo: 2) */
  int32_t has_unconditional;  /* latent_droprate > 0: *_unconditional_embedding present */
  int32_t dynamic_max_cross;  /* cross cutoff = 3*sigma_tr + 20 per graph (score_model.py:202-205) */
  int32_t scale_by_sigma;
  int32_t no_torsion;
  float lig_max_radius, rec_max_radius, cross_max_distance, center_max_distance;
  int64_t scratch_bytes;      /* reserved (was: cap of an outer-product scratch; the fused conv kernel has none), pass 0 */
} DdkConfig;

/* Offsets (in floats) of each packed tensor inside the weight blob; layout documented in
 * disco_diffdock_b200/weights.py, which is the only producer.  L = layer, G = edge group. */
typedef enum DdkWeightId {
  DDK_W_LIG_EMB_TABLES = 0,   /* [199][24] concatenated categorical tables */
  DDK_W_LIG_NODE_W,           /* [24][56+latent] additional_features_embedder.weight */
  DDK_W_LIG_NODE_B,
  DDK_W_REC_EMB_TABLE,        /* [38][24] */
  DDK_W_REC_NODE_W,           /* [24][1336+latent] */
  DDK_W_REC_NODE_B,
  DDK_W_LIG_EDGE_W1, DDK_W_LIG_EDGE_B1, DDK_W_LIG_EDGE_W2, DDK_W_LIG_EDGE_B2,       /* [24][68+2 latent], [24], [24][24], [24] */
  DDK_W_REC_EDGE_W1, DDK_W_REC_EDGE_B1, DDK_W_REC_EDGE_W2, DDK_W_REC_EDGE_B2,       /* [24][64+2 latent] ... */
  DDK_W_CROSS_EDGE_W1, DDK_W_CROSS_EDGE_B1, DDK_W_CROSS_EDGE_W2, DDK_W_CROSS_EDGE_B2,
  DDK_W_UNCOND,               /* [5][24]: lig_node, rec_node, lig_edge, rec_edge, cross_edge (zeros if absent) */
  DDK_W_CENTER_EDGE_W1, DDK_W_CENTER_EDGE_B1, DDK_W_CENTER_EDGE_W2, DDK_W_CENTER_EDGE_B2,
  DDK_W_FINAL_EDGE_W1, DDK_W_FINAL_EDGE_B1, DDK_W_FINAL_EDGE_W2, DDK_W_FINAL_EDGE_B2,
  DDK_W_FINAL_CONV_W1, DDK_W_FINAL_CONV_B1, DDK_W_FINAL_CONV_W2, DDK_W_FINAL_CONV_B2, /* [48][48],[48],[144][48],[144] */
  DDK_W_FINAL_CONV_BN,        /* [4] scale = weight / sqrt(running_var + eps) */
  DDK_W_TR_FINAL_W1, DDK_W_TR_FINAL_B1, DDK_W_TR_FINAL_W2, DDK_W_TR_FINAL_B2,         /* [24][33],[24],[24],[1] */
  DDK_W_ROT_FINAL_W1, DDK_W_ROT_FINAL_B1, DDK_W_ROT_FINAL_W2, DDK_W_ROT_FINAL_B2,
  DDK_W_TOR_CONV_W1, DDK_W_TOR_CONV_B1, DDK_W_TOR_CONV_W2, DDK_W_TOR_CONV_B2,         /* [72][72],[72],[288][72],[288] */
  DDK_W_TOR_CONV_BN_SCALE, DDK_W_TOR_CONV_BN_SHIFT,                                   /* [48],[48] */
  DDK_W_TOR_FINAL_W1, DDK_W_TOR_FINAL_W2,                                             /* [24][48],[24] */
  DDK_W_SMEAR,                /* [4][33]: lig, rec, cross, center GaussianSmearing: 32 offsets + coeff */
  DDK_W_CONV_BASE,            /* first per-layer entry; per layer DDK_W_CONV_STRIDE entries follow */
  DDK_W_COUNT_FIXED = DDK_W_CONV_BASE
} DdkWeightId;

/* Per conv layer L the offsets table continues with (index = DDK_W_CONV_BASE + L*DDK_W_CONV_STRIDE + k): */
enum {
  DDK_WL_W1 = 0,      /* 4 entries: group g  [72][72] fc.g.0.weight (row-major [out][in]) */
  DDK_WL_B1 = 4,      /* 4 entries: [72] */
  DDK_WL_W2P = 8,     /* 4 entries: packed second-layer weights, see weights.py: per irrep class [Fk*72][Ok] */
  DDK_WL_B2P = 12,    /* 4 entries: packed second-layer bias: per class [Fk][Ok] */
  DDK_WL_BN_SCALE = 16, /* [84] */
  DDK_WL_BN_SHIFT = 17, /* [84] */
  DDK_W_CONV_STRIDE = 18
};

typedef struct DdkCtx DdkCtx;

/* One PyG-style batch (disjoint union of B graphs, nodes of a graph contiguous). */
typedef struct DdkBatch {
  int32_t B;                  /* graphs */
  int32_t NL, NR;             /* total ligand atoms / receptor residues */
  int32_t EB, ER;             /* total directed bond edges / receptor contact edges */
  int32_t RB;                 /* total rotatable bonds (= edge_mask.sum()) */
  /* host-side index data (small) */
  const int32_t* lig_ptr_h;   /* [B+1] */
  const int32_t* rec_ptr_h;   /* [B+1] */
  const int32_t* bond_index_h;/* [2][EB] global ligand indices (data['ligand','ligand'].edge_index) */
  const int32_t* bond_ptr_h;  /* [B+1] bonds of graph g are [bond_ptr[g], bond_ptr[g+1]) */
  const uint8_t* edge_mask_h; /* [EB] */
  const int32_t* rec_index_h; /* [2][ER] global receptor indices */
  const int32_t* rec_edge_ptr_h; /* [B+1] */
  const int64_t* mask_rotate_off_h; /* [B] offset (in bytes) of graph g's [R_g][N_g] mask inside mask_rotate */
  /* device-side payload */
  const int32_t* lig_x;       /* [NL][16] categorical atom features */
  const float* bond_attr;     /* [EB][4] */
  const float* rec_x;         /* [NR][1281]: amino-acid index (as float) then ESM embedding */
  const float* rec_pos;       /* [NR][3] */
  const uint8_t* mask_rotate; /* concatenated per-complex masks, row-major [R][N] */
  const float* lig_latent;    /* [NL][latent_dim] or NULL */
  const float* rec_latent;    /* [NR][latent_dim] or NULL */
  const float* lig_uncond;    /* [NL] or NULL (treated as 0) */
  const float* rec_uncond;    /* [NR] or NULL */
} DdkBatch;

/* Per-graph, per-step scalars the host derives from t (utils/diffusion_utils.py:12-16, 58-69; utils/so3.py:91-95;
 * utils/torus.py:79-83).  All DEVICE pointers, length given in brackets. */
typedef struct DdkStepInputs {
  const float* sigma_emb;     /* [B][32] sinusoidal embedding of t_tr (node_t['tr'] == complex_t['tr']) */
  const float* cross_cutoff;  /* [B] 3*sigma_tr+20, or cross_max_distance */
  const float* tr_sigma;      /* [B] sigma_tr; tr_pred is divided by it (pass 1 if !scale_by_sigma) */
  const float* rot_scale;     /* [B] so3.score_norm(sigma_rot) (or 1) */
  const float* tor_scale;     /* [B] sqrt(torus.score_norm(sigma_tor)) (or 1) */
} DdkStepInputs;

/* a*score + b*z coefficients of one reverse step (utils/sampling.py:137-192), already cast to fp32. */
typedef struct DdkStepCoef {
  float a_tr, b_tr, a_rot, b_rot, a_tor, b_tor;
} DdkStepCoef;

int ddk_abi_version(void);

/* weights_h: host blob of n_floats fp32; offsets_h: n_offsets entries indexed by DdkWeightId (+ per-layer part). */
int ddk_create(const DdkConfig* cfg, const float* weights_h, size_t n_floats, const int64_t* offsets_h,
               int32_t n_offsets, int32_t device, DdkCtx** out);
int ddk_destroy(DdkCtx* ctx);
const char* ddk_last_error(const DdkCtx* ctx);   /* ctx may be NULL: returns the creation error */

int ddk_set_batch(DdkCtx* ctx, const DdkBatch* batch, void* stream);

/* forward(): pos [NL][3] -> tr [B][3], rot [B][3], tor [RB] (tor may be NULL when RB == 0 or no_torsion). */
int ddk_score(DdkCtx* ctx, const float* lig_pos, const DdkStepInputs* in, float* tr, float* rot, float* tor,
              void* stream);

/* node features after the last conv layer: lig [NL][84], rec [NR][84].  The ligand rows are valid after ddk_score or ddk_embed; the
 * receptor rows only after ddk_embed (before the score heads the last conv layer skips receptor nodes, whose features the
 * heads never read: models/score_model.py:269-307) -- asking for rec_out after ddk_score returns DDK_ERR_STATE. */
int ddk_embed(DdkCtx* ctx, const float* lig_pos, const DdkStepInputs* in, void* stream);
int ddk_get_node_features(DdkCtx* ctx, float* lig_out, float* rec_out, void* stream);

/* pos <- modify_conformer_batch(pos, a*score + b*z ...) in place; z_* may be NULL (treated as zeros). */
int ddk_update(DdkCtx* ctx, float* lig_pos, const float* tr, const float* rot, const float* tor,
               const float* z_tr, const float* z_rot, const float* z_tor, const DdkStepCoef* coef_h, void* stream);

/* The reverse-diffusion loop: n_steps x (score, update).  step_inputs point to arrays with a leading [n_steps]
 * dimension (e.g. sigma_emb [n_steps][B][32]); z_* [n_steps][...] or NULL; coef_h [n_steps] on the host. */
int ddk_sample(DdkCtx* ctx, float* lig_pos, int32_t n_steps, const DdkStepInputs* step_inputs,
               const float* z_tr, const float* z_rot, const float* z_tor, const DdkStepCoef* coef_h, void* stream);

/* Same with HOST buffers: copies pos / noise / step inputs in, runs, copies the final pos back, synchronises. */
int ddk_sample_host(DdkCtx* ctx, float* lig_pos_h, int32_t n_steps, const DdkStepInputs* step_inputs_h,
                    const float* z_tr_h, const float* z_rot_h, const float* z_tor_h, const DdkStepCoef* coef_h);

/* Introspection for tests / benchmarks (synchronising). */
int64_t ddk_kernel_launches(const DdkCtx* ctx);        /* kernels launched by this context so far */
int64_t ddk_last_edge_count(DdkCtx* ctx);              /* edges of the combined graph in the last ddk_score (sync) */
int64_t ddk_edge_total(DdkCtx* ctx);                   /* cumulative *dynamic* (ligand radius + cross) edges listed since ddk_create (sync);
                                                          the static bond / receptor-contact edges are not included */
int64_t ddk_segment_total(DdkCtx* ctx);                /* cumulative non-empty (node, edge group) segments since ddk_create (sync): the
                                                          second radial-MLP layer runs once per segment and conv layer */
#define DDK_WORK_LISTS 11
int ddk_group_totals(DdkCtx* ctx, int64_t* edges, int64_t* segments);   /* cumulative listed edges / non-empty segments per work
                                                          list ([DDK_WORK_LISTS] each): edge groups 0 lig-lig, 1 lig<-rec, 2 rec-rec,
                                                          3 rec<-lig, and 4 + h = group 2 restricted to the residues within h
                                                          receptor-contact hops of a residue with a cross edge; all steps (sync) */
/* Process-wide run-time choice of the conv-layer kernels (same meaning as the DDK_TC environment variable):
 *   2 = k_conv_tcr (the default, fastest measured): EVERY outer-product accumulation as 3xTF32 tcgen05.mma with the contraction
 *       against the resident second radial-MLP layer straight from tensor memory (no scratch round trip);
 *   1 = k_conv_fused (FFMA2) + k_acc_tc (only the long lig<-rec segments as 3xTF32 tcgen05.mma, A blocks through a scratch buffer);
 *   0 = k_conv_fused only (all fp32 FMA: the strictest rounding, for ill-conditioned trajectories);
 *  -1 = follow the environment again.
 * Takes effect at the next ddk_create / ddk_set_batch.  Returns the previous override.  Parity tests run the trajectories in
 * every mode. */
int ddk_debug_set_tc(int32_t on);
int ddk_debug_read(DdkCtx* ctx, const char* name, void* dst_h, size_t max_bytes, size_t* n_bytes);

/* Optional per-launch timing with CUDA events on the launching stream (used by bench.py for the roofline line).
 * Kernel classes: 0 setup, 1 graph (lists + edge embeddings), 2 node projections, 3..6 conv accumulate (basis level
 * 0..3), 7 conv contract / finalize, 8 score heads, 9 update, 10 first radial-MLP layer per listed edge (k_edge_hidden),
 * 11..14 tensor-core conv kernels (k_conv_tcr, or k_acc_tc in mode 1; basis level 0..3).
 * ddk_profile_read synchronises the device, adds the elapsed milliseconds / launch counts since the last read into
 * ms[15] / launches[15] and clears the records. */
#define DDK_PROFILE_CLASSES 15
int ddk_profile_enable(DdkCtx* ctx, int32_t on);
int ddk_profile_read(DdkCtx* ctx, double* ms, int64_t* launches);

/* Host builds of two device routines of the update kernel, callable without a GPU (unit tests):
 * rigid alignment R a_n + t ~ b_n (utils/geometry.py:126-156) and axis-angle -> matrix (utils/geometry.py:38-85). */
int ddk_host_kabsch(const float* a_h, const float* b_h, int32_t n, float* R9_h, float* t3_h);
int ddk_host_axis_angle_to_matrix(const float* axis_angle3_h, float* R9_h);

/* Host-side self check of the fused conv kernel's lane tables (callable without a GPU): for basis level 0..3 every row of
 * the FasterTensorProduct basis (models/tensor_layers.py:65-116) must be owned by exactly one (slot, lane).  Returns 0 if
 * all four levels are covered, otherwise 1 + the first failing level. */
int ddk_host_lane_tables_check(void);

/* Host evaluation of the basis-row table of the tensor-core accumulation kernel (k_acc_tc; models/tensor_layers.py:65-116): the
 * raw basis values (no constant factors) of basis level lv for one destination feature row x[84] and one harmonics record sh[4],
 * written in kernel row order.  Returns the number of rows (96 / 138 / 180 / 276), or -1 if the table is inconsistent. */
int ddk_host_tc_rows_eval(int32_t lv, const float* x84_h, const float* sh4_h, float* basis_out_h);

/* Host build of the TF32 split of k_acc_tc's operands (3xTF32 product hi*hi + hi*lo + lo*hi): for every a[i], hi[i] = a rounded to
 * the TF32 grid (13 low mantissa bits zero) and lo[i] = a - hi[i] (exact in fp32), both as raw fp32 bit patterns. */
int ddk_host_tc_split(const float* a_h, int32_t n, uint32_t* hi_h, uint32_t* lo_h);

/* Host self check of k_conv_tcr's role tables (callable without a GPU): at every basis level each (basis row, hidden unit) pair and
 * each (basis row, bias) pair is owned by exactly one role, the per-role weight slices reproduce the packed second-layer weights
 * (models/tensor_layers.py:154-156 re-associated, disco_diffdock_b200/weights.py) and every shape fits the kernel's tensor-memory /
 * shared-memory budgets; and, for a random accumulator block, the partial records the contraction warps would write, added up
 * through the finalize table (TcrRole::fsrc), equal the direct contraction W2p (*) A + b2p (*) Bsum.  Returns 0, 1 + the first
 * level whose tables fail, or 10 + the first level whose records fail. */
int ddk_host_tcr_roles_check(void);

/* Host build of k_conv_tcr's operand split: hi and lo both on the TF32 grid, |a - hi - lo| <= 2^-22 |a|. */
int ddk_host_tc_split_rn(const float* a_h, int32_t n, uint32_t* hi_h, uint32_t* lo_h);

#ifdef __cplusplus
}
#endif
#endif /* DDK_H_ */
