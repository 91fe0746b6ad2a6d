/* attnshift_b200.h -- C ABI of libattnshift_b200.so (sm_100a).
 *
 * Drop-in boundary for the AttentionShift hot path.  The reference has no native interface for this path (it is
 * Python / PyTorch over mmdetection); each entry point below names the reference code it replaces:
 *   VT  = models/vision_transformer.py
 *   VTD = mmdet/models/backbones/visual_transformer_det.py
 *   RH  = mmdet/models/roi_heads/stdroi_point_deform_attn_reppoints.py
 * Python host code (attentionshift_b200/*.py) binds these with ctypes and keeps the mmdet registry surface
 * (@BACKBONES VisionTransformerDet, @HEADS AttnShiftRoIHead); see INTEGRATION.md for the binding a maintainer adds.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless marked "host"; memory is allocated, owned and freed by the caller
 *     (PyTorch); the library never allocates persistent device memory and never retains a pointer across calls;
 *   - tensors are contiguous row-major unless a stride argument is given; "f16" = IEEE half stored in 2 bytes;
 *   - every function enqueues on `stream` and returns without synchronising;
 *   - return value 0 = success, a cudaError_t value, or AS_ERR_* (>10000); the only in-tree native precedent
 *     (mmdet/ops/chamfer_2d/src/chamfer_2d.cu:135-141) likewise returns an int status;
 *   - *_workspace() functions return the scratch size in bytes the matching call needs.
 */
#ifndef ATTNSHIFT_B200_H_
#define ATTNSHIFT_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* as_stream_t; /* == cudaStream_t */

#define AS_ERR_BAD_ARG 10001
#define AS_ERR_NO_DRIVER 10002
#define AS_ERR_TMAP 10003

/* ------------------------------------------------------------------ ViT block (VT:62-124, VTD:192-275) */

/* nn.Linear on tcgen05 tensor cores: y[M,N] = x[M,K] * w[N,K]^T + bias (VT:76/84 qkv / proj, VT:40-59 Mlp, VT:136 patch conv
 * as GEMM).  mode 0: f16 out; 1: GELU(erf) -> f16 out (VT:55-56); 2: f32 out = resid + y (VT:114-115 residual); 4: f32 out. */
int as_linear_f16(const void* x_f16, const void* w_f16, const float* bias, void* out, const float* resid, int M, int N,
                  int K, int mode, as_stream_t stream);

/* Weight gradient of a Linear without transposed copies: out [M, N] f32 = a^T b for a [R, M], b [R, N] fp16 row-major
 * (dW = dY^T X, R = tokens; autograd's addmm of VT:76 / 84 / 52 / 54 in the reference).  Same tcgen05 kernel as as_linear_f16
 * with both operands consumed as MN-major tiles; R arbitrary (TMA zero-fills), M and N multiples of 64. */
int as_linear_tn_f16(const void* a_f16, const void* b_f16, float* out, int R, int M, int N, as_stream_t stream);

/* VT:76: qkv Linear fused with the reshape/permute to heads: q,k [B,h,T,64] f16, vt [B,h,64,Tpad] f16 (V transposed). */
int as_qkv_proj_f16(const void* x_f16, const void* w_f16, const float* bias, void* q, void* k, void* vt, int B, int T,
                    int Tpad, int heads, as_stream_t stream);

/* VT:79-83: softmax(q k^T * 64^-0.5) v without materialising attn.  o [B,T,h*64] f16; m,l [B,h,T] f32 row statistics. */
int as_mhsa_fwd(const void* q, const void* k, const void* vt, void* o, float* m, float* l, int B, int T, int Tpad,
                int heads, as_stream_t stream);

/* Schedule of as_mhsa_fwd (no reference counterpart; benchmarking aid): 1 = two passes over S per tile, 2 = single pass
   with a window test per chunk, 3 = single pass with a row-sum test after the tile, 4 (default) = 3 with a quarter of the
   exponentials on the FMA pipe.  Env: AS_MHSA_VARIANT. */
int as_mhsa_set_variant(int variant);

/* VT:79-83 for the RoI decoders that reuse the ViT Block at T <= 197 with head_dim 32 (mae_bbox_head_rec.py:148-167,
 * mae_mask_head_pointSup.py:172-190): qkv [B, T, 3, heads, head_dim] f16 = the qkv Linear's output as it lies, T <= 256,
 * head_dim 32 or 64 -> o [B, T, heads*head_dim] f16.  One CTA per (head, batch item), K / V in shared memory, fp32 math. */
int as_mhsa_small(const void* qkv, void* o, int B, int T, int heads, int head_dim, as_stream_t stream);

/* Training path: batched transpose with zero padding, src [batch, R, C] f16 -> dst [batch, C, Rp] f16 (Rp >= R, C and Rp even).
 * Makes the K-major operands of the backward GEMMs (dW = dY^T X) and of as_mhsa_bwd (Q^T, K^T, dO^T) -- torch autograd's own
 * backward of VT:76-84 transposes implicitly inside cuBLAS. */
int as_transpose_pad_f16(const void* src, void* dst, int batch, int R, int C, int Rp, as_stream_t stream);

/* Backward of as_mhsa_fwd (what torch autograd derives from VT:79-83 in the reference; needed for DDP training of the
 * backbone, mmdet/apis/train.py:96-100).  Flash style on tcgen05: P is recomputed from (q, k, m, l), nothing of size T x T is
 * stored; two deterministic kernels (dK / dV with a resident key tile, dQ with a resident query tile).
 * q, k, v, d_o [B,heads,T,64] f16 (head-major rows); qt, kt, dot: the same tensors transposed and zero padded, [B,heads,64,Tpad]
 * (Tpad = T rounded up to 128); m, l [B,heads,T] as written by as_mhsa_fwd; delta [B,heads,T] = rowsum(dO o O).
 * dq, dk, dv [B,heads,T,64] f32, gradients w.r.t. the unscaled q, k (the head_dim^-0.5 of VT:79 is applied inside). */
int as_mhsa_bwd(const void* q, const void* k, const void* v, const void* d_o, const void* qt, const void* kt, const void* dot,
                const float* m, const float* l, const float* delta, float* dq, float* dk, float* dv, int B, int T, int Tpad,
                int heads, as_stream_t stream);

/* as_mhsa_bwd with one more output mode: dqkv16 non-null -> dQ | dK | dV are written as fp16 into ONE [B*T, 3*heads*64]
 * tensor laid out like the output of VT:76's qkv Linear (column = which*C + head*64 + d), the operand of that Linear's dX /
 * dW GEMMs; dq / dk / dv are then not written (may be null). */
int as_mhsa_bwd_ex(const void* q, const void* k, const void* v, const void* d_o, const void* qt, const void* kt,
                   const void* dot, const float* m, const float* l, const float* delta, float* dq, float* dk, float* dv,
                   void* dqkv16, int B, int T, int Tpad, int heads, as_stream_t stream);

/* ---- element-wise / reduction half of the block backward (VT:109-124 under autograd in the reference)
 * as_colsum: out [N] f32 = column sums of x [M, N] (bias gradient); x fp16 (x_is_f16) or fp32; cast16 (fp32 x only, may be
 *   null) receives half(x), the operand of the following GEMMs.  N % 4 == 0.
 * as_gelu_bwd_f16: d_pre [M, N] fp16 = d_hid * gelu'(pre) (erf GELU, VT:40) and d_bias [N] = column sums of d_pre.
 * as_layernorm_bwd: x [M, C] f32 = the LayerNorm input (statistics are recomputed), dy [M, C] fp16 = gradient of the fp16
 *   LayerNorm output; dx [M, C] f32 = LayerNorm backward (+ resid_grad [M, C] f32 when non-null: the residual branch of
 *   VT:113-114); dgb [2, C] = d gamma | d beta.  C a multiple of 128, <= 1024.
 * as_attn_bwd_prep: d_o, o [B, T, heads*64] fp16 -> d_oh [B, heads, T, 64] (head-major copy of d_o) and
 *   delta [B, heads, T] = rowsum(d_o o o), the inputs of as_mhsa_bwd.
 * Workspaces: as_colsum_workspace(N) bytes for as_colsum / as_gelu_bwd_f16, as_layernorm_bwd_workspace(C).  Column partials
 * are summed in a fixed order (no atomics). */
size_t as_colsum_workspace(int N);
int as_colsum(const void* x, int x_is_f16, int M, int N, void* cast16, float* out, void* workspace, size_t workspace_bytes,
              as_stream_t stream);
int as_gelu_bwd_f16(const void* d_hid, const void* pre, int M, int N, void* d_pre, float* d_bias, void* workspace,
                    size_t workspace_bytes, as_stream_t stream);
size_t as_layernorm_bwd_workspace(int C);
int as_layernorm_bwd(const float* x, const float* gamma, const void* dy, const float* resid_grad, int M, int C, float eps,
                     float* dx, float* dgb, void* workspace, size_t workspace_bytes, as_stream_t stream);
int as_attn_bwd_prep(const void* d_o, const void* o, int B, int T, int heads, void* d_oh, float* delta, as_stream_t stream);

/* VTD:236/242 attn.mean(1): out [B,T,ld] f32 (ld >= T), rowsum_part [B,T,rowsum_slices*ceil(T/128)] partial row sums in
 * column order (may be NULL).  rowsum_slices = 4: persistent schedule (needs ld = T rounded up to 128), one partial per
 * 32-column slice; rowsum_slices = 1: one CTA per tile, one partial per 128-column tile.
 * t_hi / t_lo (may be NULL): the TRANSPOSED map as a split-fp16 pair (x * t_scale = hi + lo), [B,ldt,ldt], ldt = T rounded
 * up to 128, fully written (zero padded) -- the K-major B operand of the tensor-core roll-out. */
int as_attn_headmean(const void* q, const void* k, const float* m, const float* l, float* out, int ld, float* rowsum_part,
                     int rowsum_slices, void* t_hi, void* t_lo, int ldt, float t_scale, int B, int T, int heads,
                     as_stream_t stream);
/* Same, producing only what the roll-out (RH:1257-1272 restricted to the rows RH:2272 reads) consumes of a layer
 * (persistent schedule, rowsum_slices = 4): out may be NULL -- transposed pair + row sums only, for every layer but the last;
 * q_row0 > 0 (needs t_hi = NULL) -- only the query tiles containing rows [q_row0, T) are computed, for the last layer, whose map
 * enters the roll-out through its point-token rows alone.  Rows of out / rowsum_part before the first computed tile are not
 * written. */
int as_attn_headmean_ex(const void* q, const void* k, const float* m, const float* l, float* out, int ld, float* rowsum_part,
                        int rowsum_slices, void* t_hi, void* t_lo, int ldt, float t_scale, int B, int T, int heads, int q_row0,
                        as_stream_t stream);

/* Batched f16 x f16 -> f32 GEMM on tcgen05: out[b] = resid[b] + alpha * x[b] w[b]^T (resid may alias out / be NULL). */
int as_bgemm_f16_f32(const void* x_f16, const void* w_f16, float* out, const float* resid, int batch, int M, int N, int K,
                     int x_rows, int w_rows, int ldo, long long out_bstride, float alpha, as_stream_t stream);

/* VT:110/114 LayerNorm (eps argument; 1e-6 at VT:146) with f16 output for the following GEMM. */
int as_layernorm_f16(const float* x, const float* gamma, const float* beta, void* y_f16, int M, int C, float eps,
                     as_stream_t stream);

/* VT:136 / VTD:195 patch embedding operand: img [B,3,H,W] f32 -> [B*(H/16)*(W/16), 768] f16 rows (c, ky, kx). */
int as_patch_im2col_f16(const float* img, void* cols_f16, int B, int H, int W, as_stream_t stream);

/* VTD:203-213: x [B, 1+N+Tp, C] = [cls + pos0 | emb + pos | point tokens]. */
int as_assemble_tokens(const float* emb, const float* cls, const float* pos, const float* ptok, float* x, int B, int N,
                       int Tp, int C, as_stream_t stream);

/* ------------------------------------------------------------------ attention roll-out (RH:1257-1272, RH:2272) */

size_t as_rollout_workspace(int B, int T, int n_rows);
/* attn / rowsum_part: HOST arrays of L device pointers (oldest layer first).  out [B,L,n_rows,T]. */
int as_rollout_rows(const float* const* attn, const float* const* rowsum_part, int L, int B, int T, int ld, int ntile,
                    int n_rows, float* out, void* workspace, size_t workspace_bytes, as_stream_t stream);

/* Same roll-out on the tensor cores: split-fp16 operands (3 MMAs per product, fp32-level accuracy).  t_hi / t_lo: HOST
 * arrays of L device pointers to the transposed maps written by as_attn_headmean.  out [B,L,n_rows,ldt] (row stride ldt). */
size_t as_rollout_tc_workspace(int B, int T, int ldt);
int as_rollout_rows_tc(const float* const* attn, const void* const* t_hi, const void* const* t_lo,
                       const float* const* rowsum_part, int L, int B, int T, int ld, int ldt, float t_scale, int ntile,
                       int n_rows, float* out, void* workspace, size_t workspace_bytes, as_stream_t stream);

/* ------------------------------------------------------------------ CAM -> pseudo box (RH:2272-2290, RH:60-116) */

int as_cam_gather(const float* rows, const int* obj_img, const int* obj_pt, int L, int n_rows, int row_stride, int N, int n_tot,
                  float* cams, as_stream_t stream);
int as_cam_minmax(const float* lows, int n_maps, int hp, int wp, float* minmax, void* scratch, as_stream_t stream);
size_t as_cam_bbox_workspace(int n_maps, int H, int W);
int as_cam_bbox(const float* lows, const float* minmax, const float* points, int n_maps, int n_tot, int hp, int wp,
                float cam_thr, float area_ratio, float* boxes, unsigned char* keep_mask, void* workspace,
                size_t workspace_bytes, as_stream_t stream);

/* ------------------------------------------------------------------ cosine maps (RH:339, RH:696, RH:297-301) */

size_t as_cosine_maps_workspace(int n_img, int G, int S, int N, int C);
int as_cosine_maps(const float* feats, long long feat_img_stride, int n_img, int N, int C, const int* grp_img,
                   const float* protos, int G, int S, float* sim, int clamp0, void* workspace, size_t workspace_bytes,
                   as_stream_t stream);

/* ------------------------------------------------------------------ refined instance maps (RH:1000-1046, RH:668-707) */

/* rowcnt [n_levels][n_items][H] (n_levels <= 4): level l = per-row candidate counts with the item's threshold doubled l
 * times (RH:360-364 doubles background thresholds until there are enough candidates; foreground levels repeat level 0). */
int as_norm_rowcount(const float* low, const float* minmax, const int* item_kind, const int* item_a, const int* item_b,
                     const float* item_thr, int n_items, int hp, int wp, int n_levels, int* rowcnt, as_stream_t stream);
int as_norm_select(const float* low, const float* minmax, const int* item_kind, const int* item_a, const int* item_b,
                   const float* item_thr, int hp, int wp, const int* rowcnt, const int* sel_item, const int* sel_k,
                   int n_sel, int* out_xy, as_stream_t stream);
int as_seed_proto(const float* feats, long long feat_img_stride, const int* row_img, const int* pts, int G, int P, int C,
                  int hp, int wp, float* proto, as_stream_t stream);
int as_refine_threshold(float* cur, int rows, int N, float tau, float* wsum, as_stream_t stream);
size_t as_weighted_centroid_workspace(int G, int S, int N, int C);
int as_weighted_centroid(const float* feats, long long feat_img_stride, const int* grp_img, const float* w,
                         const float* wsum, int G, int S, int N, int C, float* out, void* workspace,
                         size_t workspace_bytes, as_stream_t stream);
/* Rows of group g: [0,n) instances (multiplied in place by their patch-grid box mask), then n_extra further rows that take
 * part in the winner-take-all (first round, RH:668-707: 1 = the image-level bg supplement, followed by n bg rows copied to
 * bg_out; second round, RH:710-748 / RH:2812-2844: 2 = fg supplement + sampled bg supplement, bg_out = NULL). */
int as_refine_select(float* cur, int G, int S, int N, int wp, const int* grp_first, const int* grp_nobj,
                     const float* rois, int emit, int n_extra, float* fg_out, float* bg_out, as_stream_t stream);
/* RH:1010-1019 + RH:2356: full-resolution fg / bg maps and uint8 pseudo masks from the low-resolution affinities. */
int as_fuse_instance_maps(const float* fg_low, const float* bg_low, int n_tot, int hp, int wp, float mask_thr,
                          float* map_fg, float* map_bg, unsigned char* mask, void* stats_scratch, as_stream_t stream);

/* ------------------------------------------------------------------ mask-head point candidates (RH:433-461, RH:1980-1990) */

size_t as_mask_candidates_workspace(int n_tot, int H, int W);
int as_mask_candidates(const float* map_fg, const float* map_bg, const float* rois, int n_tot, int H, int W,
                       float pos_thr, float neg_thr, int corr_size, unsigned char* pos, int* rowcnt, void* workspace,
                       size_t workspace_bytes, as_stream_t stream);
int as_mask_select(const unsigned char* pos, const float* map_bg, const float* rois, const void* workspace,
                   float neg_thr, const int* rowcnt, const int* sel_obj, const int* sel_kind, const int* sel_k, int n_sel,
                   int H, int W, int* out_xy, as_stream_t stream);

/* ------------------------------------------------------------------ mean shift = the attention-shift loop
 * (RH:1778-1840 seeds, RH:830-854 cosine_shift_batch, RH:882-908 update_density_batch, RH:2011-2020 seed map) */

int as_erode_downsample(const float* map_fg, int n_tot, int H, int W, float thr, int corr_size, float* fg_low,
                        float* seed_map, as_stream_t stream);
int as_grid_seeds(const float* maps, float thr, const float* feats, long long feat_img_stride, const int* obj_img,
                  const float* rois, int n_tot, int N, int C, int wp, int S, int* seed_tok, float* proto,
                  as_stream_t stream);
size_t as_mean_shift_workspace(int n_img, int n_tot, int S, int N, int C);
int as_mean_shift(const float* feats, long long feat_img_stride, int n_img, int N, int C, int hp, int wp,
                  const int* obj_img, const float* rois, int n_tot, int S, float* proto, float* sim, int n_shift,
                  double tau0, double temp, int clamp0, int* trace, void* workspace, size_t workspace_bytes,
                  as_stream_t stream);

/* Tensor-core variant (C % 64 == 0): the affinity of all seeds of an image is one batched split-fp16 tcgen05 GEMM.
 * Instances grouped by image: img_first / img_nobj device arrays [n_img]; kmax = max_i img_nobj[i] * S (host value). */
size_t as_mean_shift_tc_workspace(int n_img, int n_tot, int S, int N, int C, int kmax);
int as_mean_shift_tc(const float* feats, long long feat_img_stride, int n_img, int N, int C, int hp, int wp,
                     const int* obj_img, const int* img_first, const int* img_nobj, int kmax, const float* rois, int n_tot,
                     int S, float* proto, float* sim, int n_shift, double tau0, double temp, int clamp0, int* trace,
                     void* workspace, size_t workspace_bytes, as_stream_t stream);

/* Same contract again as ONE persistent cooperative kernel (the attention-shift loop iterated on the device with no
 * launch and no host sync per step): an image's tokens are split over ceil(N/256) co-resident CTAs that synchronise
 * through a global counter; affinity on tcgen05 from split-fp16 operands, softmax / arg-max / prototype update from
 * shared memory.  Requires C % 128 == 0, C <= 768, kmax <= 64, at most 8 instances per image, ceil(N/256) <= #SMs;
 * returns AS_ERR_BAD_ARG otherwise (callers fall back to as_mean_shift_tc).  Replaces RH:830-854 + RH:882-908. */
size_t as_mean_shift_fused_workspace(int n_img, int N, int C);
/* profiling aid: device buffer [grid][16] of uint64 receiving accumulated ns per phase of the next calls; NULL = off */
void as_mean_shift_fused_debug(unsigned long long* buf);
/* diagnostics: thread-block clusters (one per image group) the driver reported as co-resident for the last cluster launch
 * of the fused kernel (env AS_MS_CLUSTER=1); -1 before any such launch */
int as_mean_shift_fused_occupancy(void);
int as_mean_shift_fused(const float* feats, long long feat_img_stride, int n_img, int N, int C, int hp, int wp,
                        const int* obj_img, const int* img_first, const int* img_nobj, int kmax, const float* rois,
                        int n_tot, int S, float* proto, float* sim_out, int n_shift, double tau0, double temp,
                        int clamp0, int* trace, void* workspace, size_t workspace_bytes, cudaStream_t stream);

/* Second-generation persistent kernel (round 2): 128 tokens per CTA so that two CTAs of different images share an SM (one
 * streams while the other sits in the latency-bound part of its iteration), similarities resident in TMEM, update
 * accumulators per 128-channel block.  Lifts the round-1 limits: up to 256 seed columns per image (kmax = max n_obj * S),
 * up to 16 instances per image (8 when kmax <= 64), any C % 128 == 0 up to 1024 (ViT-L).  max_obj = max_i img_nobj[i].
 * as_mean_shift_v2_supported() tells whether the kernel takes a problem; as_mean_shift_v2 returns AS_ERR_BAD_ARG when it
 * does not, or when one image's CTAs cannot be co-resident.  Replaces RH:830-854 + RH:882-908. */
int as_mean_shift_v2_supported(int N, int C, int kmax, int max_obj);
size_t as_mean_shift_v2_workspace(int n_img, int N, int C, int kmax, int max_obj);
void as_mean_shift_v2_debug(unsigned long long* buf);
/* diagnostics: resident CTAs per SM granted to the <= 64-column variant (2 expected), its registers / static / dynamic smem */
int as_mean_shift_v2_occupancy(int* regs, int* static_smem, int* dyn_smem);
int as_mean_shift_v2(const float* feats, long long feat_img_stride, int n_img, int N, int C, int hp, int wp,
                     const int* img_first, const int* img_nobj, int kmax, int max_obj, const float* rois, int n_tot, int S,
                     float* proto, float* sim_out, int n_shift, double tau0, double temp, int clamp0, int* trace,
                     void* workspace, size_t workspace_bytes, cudaStream_t stream);

/* RoIAlign of the MIL layer selection (RH:2953-2972 -> mmcv RoIAlign, CFG:64-68: 7 x 7, sampling_ratio 0, aligned) on the
 * token-major feature map: feats [n_img, hp*wp, C] f32, rois [n_roi,5] = (image, x1, y1, x2, y2) px -> out [n_roi, pooled^2, C]
 * f32, the row layout MAEBoxHeadMIL's LayerNorm / decoder_embed consume (MIL:146-150). */
int as_roi_align_tokens(const float* feats, long long feat_img_stride, const float* rois, int n_roi, int hp, int wp, int C,
                        int pooled, float spatial_scale, float* out, as_stream_t stream);

/* ------------------------------------------------------------------ point-token <-> GT matching (RH:2237-2257)
 * Replaces HungarianPointAssigner.assign's host hop (mmdet/core/bbox/assigners/hungarian_point_assigner.py:95-99: cost.cpu()
 * + scipy.optimize.linear_sum_assignment) followed by PointPseudoSampler (point_pseudo_sampler.py:34-37).
 * cost [sum_i G_i, P] f32 row-major: one row of proposal costs per GT (the transpose of assign()'s [P, G_i] matrix,
 * match_cost.py:56-58,90-106), GTs in image order (image i owns rows g_first[i] .. g_first[i] + g_count[i]); max_g >= every
 * g_count.  Writes, per image, the
 * min(P, G_i) matched proposals in ascending order to pos_inds[g_first[i] + k] and the GT of each to pos_gt[g_first[i] + k].
 * status [n_img] (may be null): 1 = the matrix held NaN / -inf or no finite matching exists (scipy raises ValueError there);
 * the outputs are then the identity pairing.  fp64 shortest-augmenting-path search, one warp per image; P, max_g <= 512. */
int as_hungarian_points(const float* cost, const int* g_first, const int* g_count, int n_img, int P, int max_g,
                        int* pos_inds, int* pos_gt, int* status, as_stream_t stream);

/* ------------------------------------------------------------------ part discovery (RH:265-301, RH:222-262) */

int as_filter_seeds(const float* sim, const float* fg_low, int n_tot, int S, int N, float pos_thr, int* keep,
                    float* score, as_stream_t stream);
int as_merge_prototypes(const float* proto, const int* keep, int n_tot, int S, int C, float thr, float* merged,
                        int* n_merged, as_stream_t stream);
int as_part_centers(const float* pmap, const int* n_parts, const float* rois, const float* feats,
                    long long feat_img_stride, const int* obj_img, int n_tot, int S, int N, int C, int wp, int KP,
                    float* centers, int* valid, int* part_id, float* cfeat, float* stat_scratch, as_stream_t stream);

/* ------------------------------------------------------------------ timing slots (measurement aid, no reference counterpart)
 * Event pairs owned by the library: as_timer_record(slot, 0 / 1, stream) marks the start / end of an interval on the stream;
 * inside a stream capture the records become external event nodes, so a replayed CUDA graph re-times its kernels on every
 * replay.  as_timer_elapsed reads the last interval of a slot (synchronise the stream first). */
int as_timer_slots(void);
int as_timer_record(int slot, int which, as_stream_t stream);
int as_timer_elapsed(int slot, float* ms);

/* ------------------------------------------------------------------ host-side RNG helper (no device work)
 * First k (<= 624) raw 32-bit outputs of at::mt19937 seeded like torch.Generator().manual_seed(seed), per key:
 * out [n_keys][k].  torch.randint(high) = out % high, torch.randperm = forward Fisher-Yates on out[i] % (n - i)
 * (RH:368, RH:447 draw from torch's CPU generator). */
int as_mt19937_draws(const unsigned* seeds, int n_keys, int k, unsigned* out);

#ifdef __cplusplus
}
#endif
#endif /* ATTNSHIFT_B200_H_ */
