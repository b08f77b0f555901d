/* evdeblur_b200.h -- C ABI of the B200-native EvDeblurNeRF render / blur-loss hot path.
 *
 * Plain pointers and sizes only (no torch types).  Every pointer is a DEVICE pointer on the current CUDA device
 * unless the parameter name ends in `_host`.  Every entry point returns 0 on success, a negative EDN_E_* code
 * otherwise (edn_last_error() gives the text) and enqueues its work on `stream` (a cudaStream_t passed as void*;
 * NULL = legacy default stream) without synchronising.
 *
 * The reference (uzh-rpg/EvDeblurNeRF @ 4111020) has no FFI: its "operator interface" is the Python method surface
 * listed in SURVEY.md section 8(b).  Each entry point below names the reference file:line it replaces; the Python
 * mirror of that surface (same names / argument meaning) lives in evdeblurnerf_b200/ and binds these symbols with
 * ctypes (INTEGRATION.md shows the binding a reference maintainer would add).
 */
#ifndef EVDEBLUR_B200_H_
#define EVDEBLUR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EDN_ABI_VERSION 1

enum {
  EDN_OK = 0,
  EDN_E_INVALID = -1,   /* bad argument (shape / dim not supported, null pointer) */
  EDN_E_CUDA = -2,      /* a CUDA runtime call failed */
  EDN_E_UNSUPPORTED = -3
};

/* precision / storage codes.  EDN_TC32 (edn_render_coarse_fwd without feature_map, edn_render_fine_fwd): the tensor-core PARITY mode -- bf16 x 3 operand splitting
 * (hi.hi + lo.hi + hi.lo into fp32 TMEM accumulators), fp32 everywhere else: fp32-grade results on tcgen05. */
enum { EDN_F32 = 0, EDN_BF16 = 1, EDN_TC32 = 2 };

/* flags for the render entry points */
enum {
  EDN_FLAG_LINDISP = 1,       /* renderer.py:166-167 */
  EDN_FLAG_TRAIN = 2,         /* module.training: disables the rmnearplane mask (voxnerf.py:181) */
  EDN_FLAG_RELU_RGB = 4,      /* CRR coarse field: rgb_activate='relu' on top of the sigmoid (voxnerf.py:266) */
  EDN_FLAG_WHITE_BKGD = 8     /* nerf.py:126-127 */
};

/* VM-decomposed feature grid of one PDRF field (networks/pdrf/voxnerf.py:99-118, 132-151).
 * Planes / lines are in RENDER LAYOUT: channel-last, plane i = [H_i][W_i][C_i], line i = [L_i][C_i]
 * (edn_pack_vm_plane converts from the reference's [1,C,H,W] parameters). */
typedef struct edn_vm_grid {
  const void* plane[3];
  const void* line[3];
  int32_t plane_h[3];      /* gridSize[matMode[i][1]] */
  int32_t plane_w[3];      /* gridSize[matMode[i][0]] */
  int32_t line_len[3];     /* gridSize[vecMode[i]]    */
  int32_t n_comp[3];       /* must be {64,16,16}      */
  int32_t dtype;           /* EDN_F32 | EDN_BF16      */
  const float* basis_t;    /* basis_mat.weight transposed: [96][32] fp32 */
  float aabb_min[3];
  float aabb_max[3];
} edn_vm_grid;

/* One PDRF field's MLP weights, TRANSPOSED ([in][out], out contiguous) and zero padded as stated.
 * coarse (CRR): sigma0_t [96][64] (row 95 zero), sigma1_t [64][16], color0_t [42][64], color1_t [64][64], color2_t [64][4]
 * fine   (FVR): sigma0_t [128][256] (row 127 zero), sigma1_t [256][128] (the geo_feat columns 1..128 of sigma_net.1),
 *               sigma1_v [256] (its sigma column 0), color0_t [155][256], color1_t [256][256], color2_t [256][4]
 * color*_b: bias vectors or NULL (--rgb_add_bias, voxnerf.py:80). */
typedef struct edn_field_mlp {
  const float* sigma0_t;
  const float* sigma1_t;
  const float* sigma1_v;   /* fine only; NULL for the coarse field */
  const float* color0_t;
  const float* color1_t;
  const float* color2_t;
  const float* color0_b;
  const float* color1_b;
  const float* color2_b;
  int32_t hidden;          /* 64 (coarse) | 256 (fine) */
  int32_t geo_feat;        /* 15 (coarse) | 128 (fine) */
  const void* tc_blob;     /* bf16 tensor-core operand blob (edn_pack_fine_tc / edn_pack_coarse_tc) or NULL */
} edn_field_mlp;

const char* edn_last_error(void);
int edn_abi_version(void);

/* [1,C,H,W] fp32 (reference parameter layout, voxnerf.py:107-117) -> channel-last [H][W][C] fp32 or bf16. */
int edn_pack_vm_plane(const float* src_chw, void* dst_hwc, int32_t C, int32_t H, int32_t W, int32_t dst_dtype,
                      void* stream);

/* Counter-based RNG (Philox4x32-10) for the render path's random draws: stratified jitter t_rand (renderer.py:176), pdf
 * samples u (utils/rays.py:162) -- uniform [0,1) -- and the density noise randn * raw_noise_std (voxnerf.py:175, normal = 1,
 * scale = raw_noise_std).  Element i depends only on (seed, stream_id, i). */
int edn_fill_random(float* out, int64_t n, uint64_t seed, uint32_t stream_id, int32_t normal, float scale, void* stream);

/* VoxelNeRFBase.sample (voxnerf.py:203-208, 132-151): pts [n,3] -> feat [n,32].  Unit-test / drop-in entry. */
int edn_vm_sample(const edn_vm_grid* grid, const float* pts, float* feat, int64_t n, void* stream);

/* Coarse pass of NeRFAll.render_rays (renderer.py:157-188): sample placement, VM lookup, PE, CRR field
 * (voxnerf.py:210-259) and sigma->alpha compositing (voxnerf.py:153-201), fused.
 *   ray_batch [R][11] = o,d,near,far,viewdirs (renderer.py:443-446)
 *   t_vals [n_samples]        = linspace(0,1,n_samples)
 *   t_rand [R][n_samples]     or NULL (perturb == 0)            (renderer.py:176)
 *   noise  [R][n_samples-1]   or NULL (already * raw_noise_std)  (voxnerf.py:175)
 * precision: EDN_F32 = fp32 SIMT parity path, EDN_BF16 = tcgen05 tensor-core path (32 <= n_samples <= 128).
 * outputs: z_vals [R][S], weights [R][S], rgb [R][3], depth [R], acc [R];
 *          feat [R][S][15] or NULL (feature_map, voxnerf.py:221). */
int edn_render_coarse_fwd(const edn_vm_grid* grid, const edn_field_mlp* mlp, const float* ray_batch,
                          const float* t_vals, const float* t_rand, const float* noise, int64_t n_rays,
                          int32_t n_samples, int32_t flags, float rmnearplane, int32_t precision, float* z_vals,
                          float* weights, float* rgb, float* depth, float* acc, float* feat, void* stream);

/* bf16 UMMA operand blob of the coarse field (EDN_BF16 precision of edn_render_coarse_fwd): size and packer. */
int64_t edn_coarse_tc_blob_bytes(void);
int64_t edn_coarse_tc_pack_workspace_floats(void);
int edn_pack_coarse_tc(const edn_field_mlp* mlp, const float* basis_t, float* workspace, void* blob, void* stream);

/* sample_pdf (utils/rays.py:149-193) as called at renderer.py:199-203 + merge/sort (renderer.py:205) + z_std (:250).
 *   u_det [n_importance] = linspace(0,1,n_importance) (perturb == 0) or NULL;  u_rand [R][n_importance] or NULL.
 * outputs: z_samples [R][Ni], inds [R][Ni] int64 or NULL, z_vals [R][Nc+Ni] sorted (stable), order [R][Nc+Ni] int64
 * or NULL, z_std [R] or NULL.  cdf accumulated sequentially in fp64 and rounded (== torch CPU cumsum). */
int edn_sample_pdf_merge(const float* z_vals0, const float* weights0, const float* u_det, const float* u_rand,
                         int64_t n_rays, int32_t n_samples, int32_t n_importance, float* z_samples, int64_t* inds,
                         float* z_vals, int64_t* order, float* z_std, void* stream);

/* Size in bytes of the tensor-core operand blob of the fine field, and its packer: every K=16 slice of every layer
 * (and of the two basis_mat's) as a bf16 UMMA K-major core-matrix tile, in the order the fine kernel streams them -- for
 * both schedules: "full" (layer by layer, emits depth_feature) and "lean" (basis_mat folded into sigma_net.0 and
 * sigma_net.1's geo columns folded into color_net.0: linear maps composed in fp32 at pack time). */
int64_t edn_fine_tc_blob_bytes(void);
int64_t edn_fine_tc_pack_workspace_floats(void);
int edn_pack_fine_tc(const edn_field_mlp* mlp, const float* basis_t_coarse, const float* basis_t_fine, float* workspace,
                     void* blob, void* stream);

/* Fine pass of render_rays (renderer.py:190-217): VM lookup of both grids at the merged samples, PE, FVR field,
 * compositing.  precision: EDN_F32 = fp32 SIMT parity path, EDN_BF16 = tcgen05 tensor-core path (bf16 operands), EDN_TC32 =
 * tcgen05 parity path (bf16 x 3 split operands, fp32-grade; n_samples <= 128, feat must be NULL; needs mlp->tc_blob).
 *   z_vals [R][S] sorted merged depths;  noise [R][S-1] or NULL.
 * outputs: weights [R][S], rgb [R][3], depth [R], acc [R], feat [R][S][128] or NULL (depth_feature for AWP). */
int edn_render_fine_fwd(const edn_vm_grid* grid_coarse, const edn_vm_grid* grid_fine, const edn_field_mlp* mlp,
                        const float* ray_batch, const float* z_vals, const float* noise, int64_t n_rays,
                        int32_t n_samples, int32_t flags, float rmnearplane, int32_t precision, float* weights,
                        float* rgb, float* depth, float* acc, float* feat, void* stream);

/* NaN / Inf guard of render_rays' result dict (renderer.py:259-263: `torch.isnan(ret[k]).any()` / `isinf` per returned tensor,
 * printed, not raised).  One launch scans n_tensors (<= EDN_GUARD_MAX_TENSORS) fp32 device tensors -- pointer and element-count
 * arrays live on the HOST -- and ORs bit t into flags[0] if tensor t holds a NaN, into flags[1] if it holds an Inf.  flags:
 * 2 x uint32 in device memory, zeroed by the caller once; it is read back lazily (no synchronisation here). */
#define EDN_GUARD_MAX_TENSORS 16
int edn_check_finite(const float* const* tensors_host, const int64_t* sizes_host, int32_t n_tensors, uint32_t* flags,
                     void* stream);

/* ---- backward of the render path (the reference relies on torch autograd: loss.backward() at run_nerf.py:594) -------- */

/* One PDRF field's weights -- or their gradients -- in the reference's own nn.Linear layout ([out][in], fp32), i.e. the
 * state_dict tensors as they are: basis[g] = basis_mat.weight [32][96] of grid g feeding the field (coarse field: its own
 * grid; fine field: the coarse grid, then the fine grid, renderer.py:194-195), sigma0 [hidden][32*n_grids+63],
 * sigma1 [1+geo_feat][hidden], color0 [hidden][geo_feat+27], color1 [hidden][hidden], color2 [3][hidden], biases or NULL. */
typedef struct edn_field_weights {
  float* basis[2];
  float* sigma0;
  float* sigma1;
  float* color0;
  float* color1;
  float* color2;
  float* color0_b;
  float* color1_b;
  float* color2_b;
  int32_t hidden;
  int32_t geo_feat;
  int32_t n_grids;
} edn_field_weights;

/* Gradient planes / lines of one VM grid in RENDER LAYOUT (channel-last fp32, same dims as the edn_vm_grid). */
typedef struct edn_vm_grid_grad {
  float* plane[3];
  float* line[3];
} edn_vm_grid_grad;

/* Backward of one field's render pass: VoxelNeRFBase.sample + forward + raw2outputs (voxnerf.py:132-259) evaluated at
 * pts = o + d * z_vals (z_vals are constants: the coarse depths carry no gradient and z_samples is detached,
 * renderer.py:203).  Coarse field: grid1 = NULL, z_vals = the coarse depths; fine field: grid0 / grid1 = coarse / fine grid,
 * z_vals = merged depths.  Upstream gradients d_rgb [R][3], d_depth [R], d_acc [R], d_weights [R][S], d_feat [R][S][geo]
 * (each may be NULL = zero).  Everything below is ACCUMULATED into (+=): grad_w (same layout as w), grad_grid*, and
 * d_ray_batch [R][11] (columns o, d, viewdirs).  precision: EDN_F32 = fp32 GEMMs (parity), EDN_BF16 = TF32 tensor-core GEMMs.
 * The activations are recomputed chunk by chunk into `workspace`; edn_field_bwd_workspace_bytes gives the size for a chunk
 * of `chunk_rays` rays (any workspace holding >= 1 ray works; larger chunks run faster). */
/* Optional merge of the two passes' coarse-grid scatters (NULL = off).  The fine pass samples the coarse grid at all merged depths;
 * n_coarse of them per ray are the coarse pass's own positions (merged `order` < n_coarse, edn_sample_pdf_merge).  Fine-field call:
 * `order` [R][S] given -> those rows of d P are written to moved[ray][order][96] (storage type of `precision`: fp32 / bf16) instead
 * of being scattered.  Coarse-field call (n_coarse == n_samples, order ignored): `moved` is added to its own d P before the
 * scatter.  Call the fine field first.  Same gradients, a third fewer coarse-grid reds. */
typedef struct edn_field_bwd_merge {
  const int64_t* order;
  int32_t n_coarse;
  void* moved;        /* [R][n_coarse][96] of the activation storage type, 16-byte aligned */
} edn_field_bwd_merge;

int64_t edn_field_bwd_workspace_bytes(int32_t n_grids, int32_t hidden, int32_t geo_feat, int64_t chunk_rays, int32_t n_samples);
int edn_render_field_bwd(const edn_vm_grid* grid0, const edn_vm_grid* grid1, const edn_field_weights* w,
                         const float* ray_batch, const float* z_vals, const float* noise, int64_t n_rays, int32_t n_samples,
                         int32_t precision, const float* d_rgb, const float* d_depth, const float* d_acc,
                         const float* d_weights, const float* d_feat, const edn_field_weights* grad_w,
                         const edn_vm_grid_grad* grad_grid0, const edn_vm_grid_grad* grad_grid1, float* d_ray_batch,
                         void* workspace, int64_t workspace_bytes, const edn_field_bwd_merge* merge, void* stream);

/* Channel-last gradient plane [H][W][C] -> += into the reference layout [1,C,H,W] (inverse of edn_pack_vm_plane). */
int edn_unpack_vm_plane_grad(const float* src_hwc, float* dst_chw, int32_t C, int32_t H, int32_t W, int32_t accumulate, void* stream);

/* ---- mode = nerf ("run_network") -------------------------------------------------------------------------------------- */

/* Vanilla NeRF field weights (networks/nerf.py:23-44), TRANSPOSED [in][out] fp32:
 * pts_t[0] [64][256] (row 63 zero), pts_t[1..4,6,7] [256][256], pts_t[5] [320][256] = rows 0..62 PE part, row 63 zero,
 * rows 64..319 h part (skip concat [input_pts, h], nerf.py:138-139); alpha_w [256]; feature_t [256][256];
 * views_t [283][128] (rows 0..255 feature, 256..282 view-dir PE); rgb_t [128][4] (col 3 zero); rgb_b [3] or NULL. */
typedef struct edn_nerf_mlp {
  const float* pts_t[8];
  const float* pts_b[8];
  const float* alpha_w; const float* alpha_b;
  const float* feature_t; const float* feature_b;
  const float* views_t; const float* views_b;
  const float* rgb_t; const float* rgb_b;
} edn_nerf_mlp;

/* NeRF.mlpforward + NeRF.eval (nerf.py:46-72, 131-162) at pts = o + d * z_vals: raw [R][S][4] = rgb(3), sigma;
 * feature [R][S][256] or NULL (h when feature_after_linear = 0, feature_linear output otherwise). */
int edn_nerf_mlp_fwd(const edn_nerf_mlp* mlp, const float* ray_batch, const float* z_vals, int64_t n_rays, int32_t n_samples,
                     int32_t feature_after_linear, float* raw, float* feature, void* stream);

/* NeRF.raw2outputs (nerf.py:74-129). flags: EDN_FLAG_TRAIN, EDN_FLAG_WHITE_BKGD. */
int edn_nerf_raw2outputs(const float* raw, const float* z_vals, const float* ray_batch, const float* noise, int64_t n_rays,
                         int32_t n_samples, int32_t flags, float rmnearplane, float* weights, float* rgb, float* depth,
                         float* acc, void* stream);

/* Vanilla NeRF field weights -- or their gradients -- in the reference's nn.Linear layout ([out][in] fp32, the state_dict tensors as
 * they are): pts_w[0] [256][63], pts_w[1..4,6,7] [256][256], pts_w[5] [256][319], alpha_w [1][256], feature_w [256][256],
 * views_w [128][283], rgb_w [3][128]; rgb_b may be NULL. */
typedef struct edn_nerf_weights {
  float* pts_w[8];
  float* pts_b[8];
  float* alpha_w; float* alpha_b;
  float* feature_w; float* feature_b;
  float* views_w; float* views_b;
  float* rgb_w; float* rgb_b;
} edn_nerf_weights;

/* Backward of edn_nerf_mlp_fwd + edn_nerf_raw2outputs (nerf.py:46-72, 131-162, 74-129; autograd at run_nerf.py:594) at
 * pts = o + d * z_vals.  Upstream gradients d_rgb [R][3], d_depth [R], d_acc [R], d_weights [R][S], d_feat [R][S][256]
 * (each may be NULL; for EDN_FLAG_WHITE_BKGD pass d_acc - sum_c d_rgb_c as d_acc, nerf.py:126-127); feature_after_linear
 * selects which tensor d_feat refers to (feature_linear output, or the trunk output h).  grad_w and d_ray_batch [R][11] are
 * ACCUMULATED.  precision: EDN_F32 = fp32 GEMMs, EDN_BF16 = TF32 tensor-core GEMMs. */
int64_t edn_nerf_bwd_workspace_bytes(int64_t chunk_rays, int32_t n_samples);
int edn_nerf_field_bwd(const edn_nerf_weights* w, const float* ray_batch, const float* z_vals, const float* noise, int64_t n_rays,
                       int32_t n_samples, int32_t flags, int32_t precision, const float* d_rgb, const float* d_depth,
                       const float* d_acc, const float* d_weights, const float* d_feat, int32_t feature_after_linear,
                       const edn_nerf_weights* grad_w, float* d_ray_batch, void* workspace, int64_t workspace_bytes, void* stream);

/* Coarse sample placement alone (renderer.py:163-178), bit-exact: z_vals [R][n_samples]. */
int edn_place_samples(const float* ray_batch, const float* t_vals, const float* t_rand, int64_t n_rays, int32_t n_samples,
                      int32_t flags, float* z_vals, void* stream);

/* ---- blur kernel + render() prologue ------------------------------------------------------------------------------ */

/* DP-NeRF rigid blur kernel weights (nn.Linear layout [out][in], fp32), names as in the reference state_dict
 * (SURVEY Appendix A): kernelsnet.view_embed_module.img_embed [n_img][32], {r,v,w}_branch.0 [32][32]+[32],
 * r_linear / v_linear [3*num_motion][32]+[3*num_motion], w_linear [num_motion+1][32]+[num_motion+1]. */
typedef struct edn_rbk_params {
  const float* img_embed;
  const float* r_branch_w; const float* r_branch_b;
  const float* v_branch_w; const float* v_branch_b;
  const float* w_branch_w; const float* w_branch_b;
  const float* r_linear_w; const float* r_linear_b;
  const float* v_linear_w; const float* v_linear_b;
  const float* w_linear_w; const float* w_linear_b;
  int32_t num_motion;      /* kernel_ptnum - 1 */
  int32_t n_img;
  float rv_window;         /* 0.1 */
} edn_rbk_params;

/* RigidBlurringModel.forward + rbk_warp (dpnerf/blurmodel.py:129-173, 51-82; utils/rigid_warping.py:18-49, 72-154) fused
 * with the render() prologue (renderer.py:423-446) and get_ndc_rays (utils/rays.py:104-145).
 *   rays [N][3][2], images_idx [N] int64
 * outputs: new_rays [N][E][3][2] or NULL, weight [N][E], img_embed [N][32] or NULL, ray_batch [N*E][11] or NULL. */
int edn_rbk_warp_ndc_fwd(const edn_rbk_params* p, const float* rays, const int64_t* images_idx, int64_t n_rays,
                         int32_t H, int32_t W, float focal, float near, float far, int32_t ndc, float* new_rays,
                         float* weight, float* img_embed, float* ray_batch, void* stream);

/* render() prologue alone (renderer.py:423-446): rays [R][3][2] -> ray_batch [R][11]. */
int edn_build_ray_batch(const float* rays, int64_t n_rays, int32_t H, int32_t W, float focal, float near, float far,
                        int32_t ndc, float* ray_batch, void* stream);

/* Gradients of the DP-NeRF kernel-net parameters, same names / layouts as edn_rbk_params (r / v entries may be NULL when
 * num_motion == 0). */
typedef struct edn_rbk_grads {
  float* img_embed;
  float* r_branch_w; float* r_branch_b;
  float* v_branch_w; float* v_branch_b;
  float* w_branch_w; float* w_branch_b;
  float* r_linear_w; float* r_linear_b;
  float* v_linear_w; float* v_linear_b;
  float* w_linear_w; float* w_linear_b;
} edn_rbk_grads;

/* Backward of edn_rbk_warp_ndc_fwd (dpnerf/blurmodel.py:129-173, utils/rigid_warping.py:18-154, renderer.py:423-446,
 * utils/rays.py:104-145): d_ray_batch [N*E][11] (what edn_render_field_bwd accumulated; NULL only if num_motion == 0),
 * d_weight [N][E] or NULL, d_img_embed [N][32] or NULL (gradient of the returned view latents, used by AWP)
 * -> ACCUMULATES the kernel-net parameter gradients into `grads`.
 * workspace: edn_rbk_bwd_workspace_floats(n_rays, num_motion) floats. */
int64_t edn_rbk_bwd_workspace_floats(int64_t n_rays, int32_t num_motion);
int edn_rbk_warp_ndc_bwd(const edn_rbk_params* p, const float* rays, const int64_t* images_idx, int64_t n_rays, int32_t H,
                         int32_t W, float focal, int32_t ndc, const float* d_ray_batch, const float* d_weight,
                         const float* d_img_embed, const edn_rbk_grads* grads, float* workspace, void* stream);

/* Backward of edn_weighted_sum: d_out [N][C] -> d_x [N*E][C] (or NULL), d_w [N][E] (or NULL); both overwritten. */
int edn_weighted_sum_bwd(const float* x, const float* w, const float* d_out, int64_t n, int32_t n_exposure, int64_t channels,
                         float* d_x, float* d_w, void* stream);

/* Backward of edn_build_ray_batch (renderer.py:423-446, utils/rays.py:104-145): d_ray_batch [R][11] -> d_rays [R][3][2]
 * (overwritten).  Lets gradients reach rays that come from a learned blur kernel other than RBK (the DSK rays below). */
int edn_build_ray_batch_bwd(const float* rays, int64_t n_rays, int32_t H, int32_t W, float focal, int32_t ndc,
                            const float* d_ray_batch, float* d_rays, void* stream);

/* ---- deformable sparse kernel (DSK) ---------------------------------------------------------------------------------------- */

/* BlurModel parameters with kernel_type = DSK, use_pattern_pos = True, depth_embed = 0 (networks/pdrf/blurmodel.py:9-107), fp32,
 * nn.Linear layout [out][in]: img_embed = img_embed.img_embed [n_img][embed]; pattern_pos [n_pat][n_pt][2] (n_pat = 1 when
 * isglobal, else n_img); pattern_trans [n_pat][n_pt][2] or NULL (optim_trans); lin_w[l] / lin_b[l] = linears.{2l}
 * ([wide][in_cnl] for l = 0, then [wide][wide]), l < num_hidden <= EDN_DSK_MAX_HIDDEN; out0 = linears1.0 ([wide][wide], or
 * [wide][in_cnl + wide] when short_cut: input columns first); out1 = linears1.2 ([3][wide], or [5][wide] when optim_sv_trans).
 * in_cnl = (2 + 4 in_embed) + embed + (spatial_embed ? 2 + 4 spatial_embed : 0). */
#define EDN_DSK_MAX_HIDDEN 4
typedef struct edn_dsk_params {
  const float* img_embed;
  const float* pattern_pos;
  const float* pattern_trans;
  const float* lin_w[EDN_DSK_MAX_HIDDEN]; const float* lin_b[EDN_DSK_MAX_HIDDEN];
  const float* out0_w; const float* out0_b;
  const float* out1_w; const float* out1_b;
  int32_t n_img, n_pt, embed, in_embed, spatial_embed, num_hidden, wide, short_cut, isglobal, optim_sv_trans;
  float kernel_hwindow;
} edn_dsk_params;

typedef struct edn_dsk_grads {
  float* img_embed;
  float* pattern_pos;
  float* pattern_trans;
  float* lin_w[EDN_DSK_MAX_HIDDEN]; float* lin_b[EDN_DSK_MAX_HIDDEN];
  float* out0_w; float* out0_b;
  float* out1_w; float* out1_b;
} edn_dsk_grads;

/* BlurModel.forward, kernel_type = DSK (pdrf/blurmodel.py:109-224): canonical kernel positions (+ noise [N][n_pt][2] =
 * randn * random_hwindow, or NULL) -> PE | view embedding | PE(pixel position) -> MLP -> position offsets, optional origin
 * offsets, softmax weights -> rays through the offset pixels with the per-ray camera poses [N][3][4].
 *   rays_x, rays_y [N] pixel coordinates; images_idx [N] int64; fx, fy, cx, cy = K[0][0], K[1][1], K[0][2], K[1][2]
 * outputs: new_rays [N][n_pt][3][2], weight [N][n_pt], align [1] (blurmodel.py:190-191).
 * workspace: edn_dsk_workspace_floats(p, n_rays) floats. */
int64_t edn_dsk_workspace_floats(const edn_dsk_params* p, int64_t n_rays);
int edn_dsk_rays_fwd(const edn_dsk_params* p, const float* rays_x, const float* rays_y, const int64_t* images_idx,
                     const float* poses, const float* noise, int64_t n_rays, int32_t H, int32_t W, float fx, float fy, float cx,
                     float cy, float* new_rays, float* weight, float* align, float* workspace, void* stream);

/* Backward of edn_dsk_rays_fwd: d_new_rays [N][n_pt][3][2] (or NULL), d_weight [N][n_pt] (or NULL), d_align device scalar (or
 * NULL) -> ACCUMULATES the parameter gradients into `grads` (same layouts as edn_dsk_params).  Same workspace size. */
int edn_dsk_rays_bwd(const edn_dsk_params* p, const float* rays_x, const float* rays_y, const int64_t* images_idx,
                     const float* poses, const float* noise, int64_t n_rays, int32_t H, int32_t W, float fx, float fy, float cx,
                     float cy, const float* d_new_rays, const float* d_weight, const float* d_align, const edn_dsk_grads* grads,
                     float* workspace, void* stream);

/* ---- adaptive weight proposal -------------------------------------------------------------------------------------------- */

/* AdaptiveWeightProposal weights (networks/dpnerf/awp.py:37-47, mam.py:13-65), fp32:
 * sample_t[l] = sample_feature_embed_layer.l.weight TRANSPOSED ([128][64], then [64][64] x3), sample_b[l] [64];
 * motion_w[l] = motion_feature_embed_layer.l.weight ([32][111], [32][32], nn.Linear layout), motion_b[l] [32];
 * mam_linear_t = MAM.linear.weight transposed [64][32]; line_conv_att [32]; conva/convb/convc [16][32]; convn/convl [16][16];
 * convd_w = MAM.Corr.convd.0.weight [32][32]; bn_weight / bn_bias = MAM.Corr.convd.1 [32]; w_linear [E][32] + [E]. */
typedef struct edn_awp_params {
  const float* sample_t[4]; const float* sample_b[4];
  const float* motion_w[2]; const float* motion_b[2];
  const float* mam_linear_t; const float* mam_linear_b;
  const float* line_conv_att;
  const float* conva; const float* convb; const float* convc; const float* convn; const float* convl;
  const float* convd_w; const float* bn_weight; const float* bn_bias;
  const float* w_linear_w; const float* w_linear_b;
  int32_t input_ch;   /* channels of depth_feature = in-features of sample_feature_embed_layer.0: 128 (c2f geo features; 0 means 128)
                       * or 256 (mode = nerf, run_nerf.py:203-212).  Widths other than 128 run the materialised (GEMM) path: the
                       * caller sets edn_awp_options.keep_activations. */
} edn_awp_params;

/* AdaptiveWeightProposal.forward (awp.py:79-117) in train mode (BatchNorm1d uses batch statistics, mam.py:24-27):
 * depth_feature [N*E][S][input_ch], z_vals [N*E][S], rays_d rows of 3 floats with row stride rays_d_stride (e.g. ray_batch + 3,
 * stride 11), view_feature [N][32] -> ccw [N][E].  workspace: edn_awp_workspace_floats() floats.
 * Options: see edn_awp_options below. */
typedef struct edn_awp_options {
  int32_t precision;         /* EDN_F32 = fused fp32 SIMT kernels (parity path); EDN_BF16 = per-sample MLP as tall TF32 GEMMs */
  int32_t keep_activations;  /* forward: run the per-sample MLP as GEMMs (fp32 ones under EDN_F32) and keep the layer activations
                                in the workspace, so that edn_awp_bwd(forward_in_workspace = 1) does not recompute them */
  int32_t phase;             /* synchronised BatchNorm across ranks: 0 = whole pass with this call's batch sums;
                                1 = stop after the local batch sums (forward: 66 doubles at edn_awp_stats_offset_floats --
                                64 sums, the row count behind them, one pad; backward: 64 doubles at
                                edn_awp_bwd_sums_offset_floats -- inside the workspace; the caller all-reduces them in place);
                                2 = finish from the sums in the workspace */
  int64_t bn_rows_total;     /* rows behind the batch sums = N * E summed over all ranks; 0 = this call's N * E in phase 0 and, in
                                phases 1 / 2, the all-reduced row count the forward left in the workspace (ragged shards are fine) */
} edn_awp_options;

int64_t edn_awp_workspace_floats(int64_t n_rays, int32_t n_exposure, int32_t n_samples, const edn_awp_options* opt);
int64_t edn_awp_stats_offset_floats(int64_t n_rays, int32_t n_exposure, int32_t n_samples);
int edn_awp_fwd(const edn_awp_params* p, const float* depth_feature, const float* z_vals, const float* rays_d,
                int32_t rays_d_stride, const float* view_feature, int64_t n_rays, int32_t n_exposure, int32_t n_samples,
                float bn_eps, const edn_awp_options* opt, float* workspace, float* ccw, void* stream);

/* Gradients of the AWP parameters, same field layout as edn_awp_params (sample_t / mam_linear_t TRANSPOSED like the weights). */
typedef struct edn_awp_grads {
  float* sample_t[4]; float* sample_b[4];
  float* motion_w[2]; float* motion_b[2];
  float* mam_linear_t; float* mam_linear_b;
  float* line_conv_att;
  float* conva; float* convb; float* convc; float* convn; float* convl;
  float* convd_w; float* bn_weight; float* bn_bias;
  float* w_linear_w; float* w_linear_b;
} edn_awp_grads;

/* Backward of edn_awp_fwd (awp.py:79-117, 49-77; mam.py:13-84; train-mode BatchNorm): d_ccw [N][E] ->
 *   grads (ACCUMULATED), d_depth_feature [N*E][S][128] (overwritten; feed it to edn_render_field_bwd's d_feat),
 *   d_rays_d rows of 3 floats with row stride d_rays_d_stride (ACCUMULATED; may be NULL), d_view_feature [N][32] (overwritten;
 *   may be NULL).  The forward is recomputed into the workspace (edn_awp_bwd_workspace_floats floats) unless
 *   forward_in_workspace != 0: then `workspace` is the buffer edn_awp_fwd(keep_activations or EDN_BF16) just filled for the same
 *   inputs (allocated with the backward's size).  opt->precision: EDN_F32 = fp32 GEMMs, EDN_BF16 = TF32 tensor-core GEMMs;
 *   opt->phase 1 / 2 split the pass around the all-reduce of the BatchNorm gradient sums (needs forward_in_workspace). */
int64_t edn_awp_bwd_workspace_floats(int64_t n_rays, int32_t n_exposure, int32_t n_samples);
int64_t edn_awp_bwd_sums_offset_floats(int64_t n_rays, int32_t n_exposure, int32_t n_samples);
int edn_awp_bwd(const edn_awp_params* p, const float* depth_feature, const float* z_vals, const float* rays_d,
                int32_t rays_d_stride, const float* view_feature, int64_t n_rays, int32_t n_exposure, int32_t n_samples,
                float bn_eps, const edn_awp_options* opt, int32_t forward_in_workspace, const float* d_ccw, const edn_awp_grads* grads,
                float* d_depth_feature, float* d_rays_d, int32_t d_rays_d_stride, float* d_view_feature, float* workspace,
                void* stream);

/* ---- loss path ------------------------------------------------------------------------------------------------------ */

/* rbk_weighted_sum (dpnerf/blurmodel.py:112-127): x [N*E][C], w [N][E] -> out [N][C]. */
int edn_weighted_sum(const float* x, const float* w, float* out, int64_t n, int32_t n_exposure, int64_t channels,
                     void* stream);

/* EDN_CRF_LUMA alone = rec601; or-ed with EDN_CRF_LUMA_REC709 / EDN_CRF_LUMA_AVG for the other luma_standard values
 * (tonemapping.py:128-135). */
enum { EDN_CRF_GAMMA = 1, EDN_CRF_LEARN = 2, EDN_CRF_SKIP_LEARN = 4, EDN_CRF_LUMA = 8, EDN_CRF_LUMA_REC709 = 16, EDN_CRF_LUMA_AVG = 32 };

/* CRF residual MLP (tonemapping.py:7-57): linear.0 [16][1+extra], linear.2 [16][16], linear.4 [16][16], linear.6 [1][16]. */
typedef struct edn_crf_params {
  const float* w0; const float* b0;
  const float* w1; const float* b1;
  const float* w2; const float* b2;
  const float* w3; const float* b3;
  int32_t extra_features;
  float gamma;
} edn_crf_params;

/* CRF.forward + TonemappingTransform.encode_rgb / encode_luma (tonemapping.py:59-93, 111-139).
 *   x [M][3]; feat [M][F] (feat_per_channel = 0), [M][3][F] (= 1) or NULL (zero padded, tonemapping.py:83-86)
 *   flags: EDN_CRF_GAMMA (map_type contains 'gamma'), EDN_CRF_LEARN (map_type == 'learn'), EDN_CRF_SKIP_LEARN,
 *          EDN_CRF_LUMA (luma, out [M][1]; otherwise out [M][3]; rec601 unless EDN_CRF_LUMA_REC709 / _AVG is set too). */
int edn_crf_fwd(const edn_crf_params* p, const float* x, const float* feat, int32_t feat_per_channel, int32_t flags,
                int64_t m, float* out, void* stream);

/* Gradients of the CRF MLP, same layout as edn_crf_params. */
typedef struct edn_crf_grads {
  float* w0; float* b0;
  float* w1; float* b1;
  float* w2; float* b2;
  float* w3; float* b3;
} edn_crf_grads;

/* Backward of edn_crf_fwd: d_out [M][3] (or [M][1] with EDN_CRF_LUMA) -> d_x [M][3] (overwritten, may be NULL); the CRF
 * parameter gradients are ACCUMULATED into `grads` (may be NULL: no parameter gradients). */
int edn_crf_bwd(const edn_crf_params* p, const float* x, const float* feat, int32_t feat_per_channel, int32_t flags, int64_t m,
                const float* d_out, float* d_x, const edn_crf_grads* grads, void* stream);

/* egm_loss (utils/events.py:260-284): luma_* [M][channels], bii [M], color_mask [M][3] uint8 one-hot or NULL,
 * color_weight [3] or NULL -> out [1]. */
int edn_egm_loss_fwd(const float* luma_start, const float* luma_end, const float* bii, const uint8_t* color_mask,
                     const float* color_weight, int32_t channels, int64_t m, float log_eps, float* out, void* stream);

/* Backward of edn_egm_loss_fwd: d_loss [1] (device) -> d_luma_start / d_luma_end [M][channels] (overwritten). */
int edn_egm_loss_bwd(const float* luma_start, const float* luma_end, const float* bii, const uint8_t* color_mask,
                     const float* color_weight, int32_t channels, int64_t m, float log_eps, const float* d_loss,
                     float* d_luma_start, float* d_luma_end, void* stream);

/* img2mse (utils/metrics.py:7): mean((x - y)^2) over n elements -> out [1]. */
int edn_img2mse(const float* x, const float* y, int64_t n, float* out, void* stream);

/* Backward of edn_img2mse: d_loss [1] (device) -> d_x [n] (overwritten). */
int edn_img2mse_bwd(const float* x, const float* y, int64_t n, const float* d_loss, float* d_x, void* stream);

/* VoxelNeRFBase.TV_loss_app (voxnerf.py:126-130, 306-324) on the reference-layout ([1,C,H,W] fp32) planes / lines of one
 * field.  workspace: 12 doubles.  out [1] = sum_i 1e-2 TV(plane_i) + 1e-3 TV(line_i). */
int edn_tv_loss_app(const float* const planes_chw[3], const float* const lines_chw[3], const int32_t plane_h[3],
                    const int32_t plane_w[3], const int32_t line_len[3], const int32_t n_comp[3], double* workspace,
                    float* out, void* stream);

/* Backward of edn_tv_loss_app: d_loss [1] (device); the gradients are ACCUMULATED into grad_*_chw (reference layout). */
int edn_tv_loss_app_bwd(const float* const planes_chw[3], const float* const lines_chw[3], const int32_t plane_h[3],
                        const int32_t plane_w[3], const int32_t line_len[3], const int32_t n_comp[3], const float* d_loss,
                        float* const grad_planes_chw[3], float* const grad_lines_chw[3], void* stream);

/* Fused Adam sweep over one flat fp32 buffer (torch.optim.Adam semantics as constructed at run_nerf.py:272-274; weight_decay
 * is the L2 term of the color_net group, run_nerf.py:244-250).  step >= 1 is the update count used for bias correction.
 * All four buffers 16-byte aligned. */
int edn_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                  float beta2, float eps, float weight_decay, int64_t step, void* stream);

/* EDI prior for one frame (utils/edi.py:7-95, data/loader_events.py:99-131): events ev_x / ev_y / ev_p (p > 0 = positive),
 * n_seg (even) sub-interval index ranges [seg_start[j], seg_end[j]) (device int64), blurry [H][W][C].
 * workspace / output bii [n_seg][H][W] (brightness increment images), sharp [H][W][C]. */
int edn_edi_prior(const float* ev_x, const float* ev_y, const float* ev_p, const int64_t* seg_start, const int64_t* seg_end,
                  int32_t n_seg, int64_t max_seg_events, const float* blurry, int32_t H, int32_t W, int32_t C, float c_pos,
                  float c_neg, float* bii, float* sharp, void* stream);

/* ---- ray / event batch generation on the device -------------------------------------------------------------- */

/* get_rays_pix (utils/rays.py:25-36): coords [n][2] = (x, y), poses [n][3][4] (or one [3][4] when broadcast_pose != 0)
 * -> rays [n][3][2] (origin, direction).  No FMA contraction: bit-exact against the reference's torch arithmetic. */
int edn_rays_from_pixels(const float* coords, const float* poses, int32_t broadcast_pose, int64_t n, double fx, double fy, double cx,
                         double cy, int32_t add_halfpix, float* rays, void* stream);

/* LLFFDataset.__getitem__ (data/loader.py:325-356): ray_ids [n] int64 over [n_img][H][W] -> rays [n][3][2], rays_x / rays_y [n]
 * (pixel + 0.5), images_idx [n] int64, rgbsf [n][3] gathered from images [n_img][H][W][3], poses_out [n][3][4].  Every output
 * except rays may be NULL. */
int edn_make_rgb_batch(const int64_t* ray_ids, int64_t n, const float* images, const float* poses, int32_t n_img, int32_t H,
                       int32_t W, double fx, double fy, double cx, double cy, float* rays, float* rays_x, float* rays_y,
                       int64_t* images_idx, float* rgbsf, float* poses_out, void* stream);

/* gather_successor (utils/events.py:221-257): follow successor_map query_hops[i] + 1 times from query_idx[i], summing the
 * positive / negative polarities of the visited events; an out-of-range successor gives (-1, 0, 0). */
int edn_gather_successor(const int64_t* query_idx, const int64_t* query_hops, int64_t n, const int64_t* successor_map,
                         const int32_t* polarity, int64_t n_events, int64_t* succ_idx, int32_t* neg_cumsum, int32_t* pos_cumsum,
                         void* stream);

/* LLFFEventsDataset.interpolate_poses (data/loader_events.py:133-148, 175-183; utils/data.py:34-61, 167-183) without spherify:
 * t [n] (float64, clipped to the key range); rotations = SLERP between key quaternions key_quats [n_keys][4] (x, y, z, w);
 * translations = piecewise cubic trans_coef [n_breaks-1][4][3] (highest power first) in (t - trans_breaks[j]) -- the PPoly form
 * of scipy's interp1d(kind="cubic") spline, computed once on the host; then the column reorder [c1, -c0, c2, T], T *= bd_scale
 * and the left multiplication by recenter_inv [4][4] = inv(recenter c2w) (NULL = none) -> poses [n][3][4] fp32. */
int edn_interpolate_poses(const double* t, int64_t n, const double* key_times, const double* key_quats, int32_t n_keys,
                          const double* trans_breaks, const double* trans_coef, int32_t n_breaks, double bd_scale,
                          const double* recenter_inv, float* poses, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EVDEBLUR_B200_H_ */
