/*
 * dvsg_warp.h -- C ABI of the B200-native (sm_100a) frame-warping hot path of DVSG.
 *
 * The reference (posgraph/coupe.DVSG) has no FFI: its boundary for this path is a set of
 * plain Python functions built from stock TensorFlow ops.  Each entry point below replaces
 * the group of TF ops named in its comment (file:line relative to the reference tree);
 * the Python modules under coupe/dvsg_b200/ keep the reference's function signatures and
 * bind these symbols with ctypes.
 *
 * Conventions
 *   - all tensors are fp32, contiguous, NHWC for images; indices are int32 internally;
 *   - every pointer is a DEVICE pointer owned by the caller (no hidden allocation, no
 *     hidden host<->device copy) except in the dvsg_host_* pipeline, which takes HOST
 *     pointers and a caller-created pipeline object that owns its staging buffers;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); all work
 *     is enqueued on it and the call returns without synchronising;
 *   - return value: 0 on success, negative DVSG_ERR_* otherwise; dvsg_last_error()
 *     returns a thread-local message for the last failing call on this thread;
 *   - the library is stateless and re-entrant (one host thread per GPU is the intended
 *     multi-GPU driving model; frames shard independently, no collective).
 */
#ifndef DVSG_WARP_H_
#define DVSG_WARP_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DVSG_OK               0
#define DVSG_ERR_INVALID     (-1)   /* bad shape / null pointer / misaligned argument   */
#define DVSG_ERR_CUDA        (-2)   /* a CUDA runtime call or kernel launch failed       */
#define DVSG_ERR_WORKSPACE   (-3)   /* caller-provided workspace too small               */
#define DVSG_ERR_UNSUPPORTED (-4)   /* valid request outside the implemented envelope    */

/* flags for dvsg_tps_warp_fwd / dvsg_bilinear_fwd / dvsg_flow_warp_fwd */
#define DVSG_FLAG_FORCE_DIRECT 1    /* use the direct-gather kernel even when the
                                       shared-memory-staged kernel is applicable          */

#define DVSG_FLAG_TPS_EXACT    2    /* TPS kernels: evaluate every radial term per pixel.
                                       Default (flag clear): the tile kernels evaluate the far
                                       field of the spline on 6 x 5 Chebyshev nodes per 32 x 8
                                       tile and interpolate (error <= 2e-7 normalised units, below
                                       the fp32 noise of the reference's own sum); control points
                                       near a tile are evaluated per pixel either way.  Pass the
                                       SAME flag to the forward and the backward call.           */

int         dvsg_version(void);
const char* dvsg_last_error(void);
/* number of kernels this library has launched on the calling thread (bench bookkeeping) */
long long   dvsg_launch_count(void);

/* ---- K1: TPS coefficient solve -------------------------------------------------------
 * Replaces _solve_system: ThinPlateSpline.py:143-166 (tf.matrix_inverse :159, tf.matmul
 * :163) and ThinPlateSpline2.py:142-165.
 *   coord  [B, pn, 2] control points (x, y); coord_batch_stride in floats between frames
 *          (pn*2 for a dense batch, 0 = one mesh shared by every frame);
 *   target [B, pn, 2] right-hand side: coord+vector (ThinPlateSpline.py:161) or the
 *          absolute targets (ThinPlateSpline2.py:160);
 *   T      [B, 2, pn+3] out; coefficient order (1, x, y, rbf_1..rbf_pn) as the grid rows of
 *          ThinPlateSpline.py:110.
 * The system matrix entries are formed in fp32 in the reference's op order; the
 * elimination itself runs in fp64 (partial pivoting) and T is rounded to fp32.
 * pn+3 <= 32: one warp per frame, no workspace.  Larger systems: one CTA per distinct
 * system inverts it into `workspace` (see dvsg_tps_solve_workspace_bytes).            */
size_t dvsg_tps_solve_workspace_bytes(int B, int pn, long long coord_batch_stride);
int dvsg_tps_solve(const float* coord, long long coord_batch_stride, const float* target,
                   float* T, int B, int pn, void* workspace, size_t workspace_bytes,
                   void* stream);
/* Backward of K1 w.r.t. the right-hand side (the only gradient the reference's callers
 * need, coord being a constant: model.py:62-68):  grad_target = (W^-T grad_T^T)[:pn].   */
int dvsg_tps_solve_bwd(const float* coord, long long coord_batch_stride, const float* grad_T,
                       float* grad_target, int B, int pn, void* workspace,
                       size_t workspace_bytes, void* stream);
/* Gradient w.r.t. the control-point positions themselves (no reference caller needs it: the
 * mesh is a constant, model.py:62-68; provided for completeness).  coord enters through the
 * radial terms of the dense grid (ThinPlateSpline.py:100-105) and through the system matrix of
 * the solve (:147-159):  grad_coord [B,pn,2] = both contributions, given T, grad_T [B,2,pn+3]
 * and the gradients grad_x / grad_y flat [B*oh*ow] w.r.t. the normalised sampling coordinates
 * (as dvsg_tps_warp_bwd writes them; null = solve part only).  `workspace` must hold W^-1 of
 * this mesh (dvsg_tps_prepare with the same coord / stride).  The right-hand side coord + vector
 * of ThinPlateSpline (:161) contributes grad_target, which the caller adds.  fp64, deterministic. */
int dvsg_tps_coord_bwd(const float* coord, long long coord_batch_stride, const float* T,
                       const float* grad_T, const float* grad_x, const float* grad_y,
                       float* grad_coord, int B, int oh, int ow, int pn, const void* workspace,
                       size_t workspace_bytes, void* stream);

/* Constant meshes (every reference call site, model.py:62-68): invert the system(s) into the
 * workspace once with dvsg_tps_prepare (fp64 Gauss-Jordan, any mesh size), then only apply
 * W^-1 per call (fp64 accumulation, rounded once).  The workspace
 * (dvsg_tps_prepare_workspace_bytes) must stay untouched between the calls.                  */
size_t dvsg_tps_prepare_workspace_bytes(int B, int pn, long long coord_batch_stride);
int dvsg_tps_prepare(const float* coord, long long coord_batch_stride, int B, int pn,
                     void* workspace, size_t workspace_bytes, void* stream);
/* *singular_out = 1 when a pivot of a system prepared in this workspace was exactly zero (duplicate control
 * points: tf.matrix_inverse, ThinPlateSpline.py:159, raises InvalidArgument there), else 0.  Synchronises. */
int dvsg_tps_prepare_status(const void* workspace, size_t workspace_bytes, int B, int pn,
                            long long coord_batch_stride, void* stream, int* singular_out);
int dvsg_tps_solve_prepared(const float* coord, long long coord_batch_stride, const float* target,
                            float* T, int B, int pn, void* workspace, size_t workspace_bytes,
                            void* stream);
/* dvsg_tps_solve_prepared with target = coord + vector (ThinPlateSpline.py:161) formed inside
 * the call: `vector` [B,pn,2] are the regressed offsets (networks.py:44).                    */
int dvsg_tps_solve_offsets_prepared(const float* coord, long long coord_batch_stride,
                                    const float* vector, float* T, int B, int pn, void* workspace,
                                    size_t workspace_bytes, void* stream);
int dvsg_tps_solve_bwd_prepared(const float* coord, long long coord_batch_stride,
                                const float* grad_T, float* grad_target, int B, int pn,
                                void* workspace, size_t workspace_bytes, void* stream);

/* ---- K2+K3: fused TPS grid generation + bilinear gather ------------------------------
 * Replaces _meshgrid (ThinPlateSpline.py:92-111), tf.matmul(T, grid) (:129) and
 * _interpolate (:30-90) -- the [B, pn+3, h*w] basis is never materialised.
 *   U [B,H,W,C] source; out [B,oh,ow,C];
 *   x_out, y_out: optional (may be NULL) flat [B*oh*ow] normalised sampling coordinates,
 *          the 2nd/3rd return values of ThinPlateSpline (:170);
 *   mask_out: optional [B,oh,ow] -- the warp of an all-ones image (model.py:82,85,121),
 *          i.e. the sum of the four bilinear weights in add_n order.                   */
/* 1 when dvsg_tps_warp_fwd / _bwd would use the tile-node evaluation for these shapes and flags
 * (assuming 16-byte aligned buffers), 0 when every radial term is evaluated per pixel.          */
int dvsg_tps_coords_mode(int H, int W, int C, int oh, int ow, int pn, int flags);
int dvsg_tps_warp_fwd(const float* U, const float* coord, long long coord_batch_stride,
                      const float* T, float* out, float* x_out, float* y_out, float* mask_out,
                      int B, int H, int W, int C, int oh, int ow, int pn, int flags,
                      void* stream);

/* Online loop (eval.py:106-110 runs the graph once per frame): dvsg_tps_solve_prepared followed
 * by dvsg_tps_warp_fwd in one call.  coord [pn,2] is the clip's constant mesh, `prepared` its
 * inverse from dvsg_tps_prepare, target = coord + vector [B,pn,2], T [B,2,pn+3] caller scratch. */
int dvsg_tps_warp_frames(const float* U, const float* coord, const float* target, void* prepared,
                         size_t prepared_bytes, float* T, float* out, float* x_out, float* y_out,
                         float* mask_out, int B, int H, int W, int C, int oh, int ow, int pn,
                         void* stream);
/* same, taking the regressed offsets `vector` [B,pn,2] (networks.py:44) instead of target      */
int dvsg_tps_warp_frames_offsets(const float* U, const float* coord, const float* vector,
                                 void* prepared, size_t prepared_bytes, float* T, float* out,
                                 float* x_out, float* y_out, float* mask_out, int B, int H, int W,
                                 int C, int oh, int ow, int pn, void* stream);

/* ---- K4: backward of K2+K3 (TF autodiff of ThinPlateSpline.py:48-89,129) -------------
 *   grad_out [B,oh,ow,C]; grad_x_in / grad_y_in: optional upstream gradients on the
 *   returned x, y (surf loss, trainer.py:363-386);
 *   grad_U [B,H,W,C]: accumulated into with atomics -- the caller zero-fills it (or
 *          passes NULL to skip the image gradient);
 *   grad_T [B,2,pn+3]: overwritten (may be NULL);
 *   grad_xs / grad_ys: optional flat [B*oh*ow] total gradient w.r.t. x_s, y_s.          */
int dvsg_tps_warp_bwd(const float* U, const float* coord, long long coord_batch_stride,
                      const float* T, const float* grad_out, const float* grad_x_in,
                      const float* grad_y_in, float* grad_U, float* grad_T, float* grad_xs,
                      float* grad_ys, int B, int H, int W, int C, int oh, int ow, int pn,
                      void* stream);

/* same with flags (DVSG_FLAG_TPS_EXACT: the flag the forward call was given)                   */
int dvsg_tps_warp_bwd_ex(const float* U, const float* coord, long long coord_batch_stride,
                         const float* T, const float* grad_out, const float* grad_x_in,
                         const float* grad_y_in, float* grad_U, float* grad_T, float* grad_xs,
                         float* grad_ys, int B, int H, int W, int C, int oh, int ow, int pn,
                         int flags, void* stream);

/* ---- K5: generic bilinear sampler -----------------------------------------------------
 * Replaces bilinear_interp (spatial_transformer.py:496-563): 1-px zero padding,
 * (W-1)/2 scaling, coordinate clip.  x, y flat [B*oh*ow] normalised; out [B,oh,ow,C].    */
int dvsg_bilinear_fwd(const float* im, const float* x, const float* y, float* out,
                      int B, int H, int W, int C, int oh, int ow, int flags, void* stream);
/* grad_im is accumulated into (caller zero-fills; NULL to skip); grad_x, grad_y
 * overwritten (NULL to skip).                                                          */
int dvsg_bilinear_bwd(const float* im, const float* x, const float* y, const float* grad_out,
                      float* grad_im, float* grad_x, float* grad_y,
                      int B, int H, int W, int C, int oh, int ow, void* stream);

/* ---- K6: dense optical-flow warp -------------------------------------------------------
 * Replaces tf_warp (warp_with_optical_flow.py:96-176).  flow [B,H,W,2], channel 0 = dx,
 * 1 = dy in pixels; out [B,H,W,C].                                                     */
int dvsg_flow_warp_fwd(const float* im, const float* flow, float* out,
                       int B, int H, int W, int C, int flags, void* stream);
int dvsg_flow_warp_bwd(const float* im, const float* flow, const float* grad_out,
                       float* grad_im, float* grad_flow, int B, int H, int W, int C,
                       void* stream);

/* ---- B1 / N2: spatial_transformer grids -------------------------------------------------
 * dvsg_st_meshgrid replaces _meshgrid (spatial_transformer.py:460-482): grid [3*oh*ow].
 * dvsg_homography_warp_fwd fuses ProjectiveTransformer._transform (:423-452, theta [B,8],
 * div_no_nan) or AffineTransformer._transform (:73-91, theta [B,6]) with bilinear_interp;
 * `projective` selects which.  x_out / y_out optional.                                  */
int dvsg_st_meshgrid(float* grid, int oh, int ow, void* stream);
int dvsg_homography_warp_fwd(const float* im, const float* theta, int projective, float* out,
                             float* x_out, float* y_out, int B, int H, int W, int C,
                             int oh, int ow, void* stream);
/* Backward of the two transformers' grid stage (no reference caller differentiates them; provided
 * for completeness): grad_x / grad_y flat [B*oh*ow] w.r.t. x_s / y_s (e.g. from dvsg_bilinear_bwd)
 * -> grad_theta [B,8] (projective: MatMul :437 and DivNoNan :446-447 gradients, zero where
 * z_s == 0) or [B,6].  Deterministic (fp64 partial sums, no atomics).                        */
int dvsg_homography_grid_bwd(const float* theta, const float* grad_x, const float* grad_y,
                             int projective, float* grad_theta, int B, int oh, int ow,
                             void* stream);

/* ---- ElasticTransformer (spatial_transformer.py:93-362) ---------------
 * The reference's second TPS formulation: regular mesh fixed at construction, U(r2) = r2 log r2
 * without epsilon, coefficients ordered (x, y, 1, w_1..w_pn), one L^-1 shared by every call.
 *   dvsg_elastic_prepare   = _initialize_tps (:324-362): L formed in fp32 as written (including the
 *                            L_mid quirk), inverted in fp64 into `workspace`;
 *   dvsg_elastic_solve     = theta @ L_inv (:283-285): theta_abs [B,2,pn] = source + offsets (all x,
 *                            then all y, :161) -> coef [B,2,pn+3]; _bwd: gradient w.r.t. theta_abs;
 *   dvsg_elastic_grid      = coefficients @ right_mat (:286-296) -> x_s, y_s flat [B*oh*ow]; the
 *                            sampling itself is dvsg_bilinear_fwd; _bwd: grad_coef [B,2,pn+3].
 * source_points [pn,2] (x, y) is the mesh of get_meshgrid(g, g) (:313-322).                        */
size_t dvsg_elastic_workspace_bytes(int pn);
int dvsg_elastic_prepare(const float* source_points, int pn, void* workspace, size_t workspace_bytes,
                         void* stream);
int dvsg_elastic_solve(const float* theta_abs, float* coef, int B, int pn, const void* workspace,
                       size_t workspace_bytes, void* stream);
int dvsg_elastic_solve_bwd(const float* grad_coef, float* grad_theta, int B, int pn,
                           const void* workspace, size_t workspace_bytes, void* stream);
int dvsg_elastic_grid(const float* source_points, const float* coef, float* x_out, float* y_out,
                      int B, int oh, int ow, int pn, void* stream);
int dvsg_elastic_grid_bwd(const float* source_points, const float* grad_x, const float* grad_y,
                          float* grad_coef, int B, int oh, int ow, int pn, void* stream);

/* ---- host-buffer pipeline (end-to-end path: H2D, solve, warp, D2H) ----------------------
 * The call a host-side user of ThinPlateSpline(U, coord, vector, out_size) makes when the
 * frames live in host memory (eval.py:106-110 feeds numpy through feed_dict every frame).
 * Frames are cut into chunks that move through `n_slots` device staging slots on separate
 * streams so that H2D, kernels and D2H overlap.  Host buffers should be pinned.        */
typedef struct dvsg_host_pipeline dvsg_host_pipeline;
int  dvsg_host_pipeline_create(dvsg_host_pipeline** out, int device, int H, int W, int C,
                               int pn, int frames_per_chunk, int n_slots);
void dvsg_host_pipeline_destroy(dvsg_host_pipeline* p);
/* U_host [B,H,W,C], coord_host [pn,2] (shared mesh) , vector_host [B,pn,2],
 * out_host [B,H,W,C]; blocks until out_host is complete (unless the pipeline is asynchronous).
 * The mesh's system is inverted when coord_host differs from the previous call's mesh.   */
int  dvsg_host_tps_warp(dvsg_host_pipeline* p, const float* U_host, const float* coord_host,
                        const float* vector_host, float* out_host, int B);
/* tf_warp on host buffers: im_host [B,H,W,C], flow_host [B,H,W,2] -> out_host [B,H,W,C]  */
int  dvsg_host_flow_warp(dvsg_host_pipeline* p, const float* im_host, const float* flow_host,
                         float* out_host, int B);
/* Streaming use (a clip is a long sequence of batches): with async != 0 the dvsg_host_* calls
 * return as soon as their copies and kernels are enqueued, so the upload of batch k+1 overlaps
 * the download of batch k; dvsg_host_pipeline_sync waits for everything submitted so far.  The
 * host buffers of a call must stay valid (and its outputs unread) until the sync.         */
int  dvsg_host_pipeline_set_async(dvsg_host_pipeline* p, int async);
int  dvsg_host_pipeline_sync(dvsg_host_pipeline* p);


/* ---- N4: frame ingest / egress (the steps on either side of the warp in eval.py) -------------
 * dvsg_frames_u8_to_f32 replaces read_frame's `cv2.cvtColor(frame, BGR2RGB); frame / 255.`
 * (eval.py:79-80; the resize of :80 is the identity when the video already has the working size,
 * see dvsg_frames_u8_resize_to_f32 otherwise): src [n_pixels,3] uint8 -> dst [n_pixels,3] fp32 = u/255 exactly as
 * numpy's fp64 division followed by the fp32 feed gives it; swap_rb != 0 reverses the channels.
 * dvsg_frames_f32_to_u8 replaces `np.uint8(frame * 255.)` + `cv2.cvtColor(.., RGB2BGR)`
 * (eval.py:112-113): truncation of the exact product, low byte of the int32 cast outside
 * [0, 256).  Both take device pointers.                                                     */
int dvsg_frames_u8_to_f32(const unsigned char* src, float* dst, long long n_pixels, int swap_rb,
                          void* stream);
int dvsg_frames_f32_to_u8(const float* src, unsigned char* dst, long long n_pixels, int swap_rb,
                          void* stream);
/* Ingest WITH the resize of eval.py:80, `cv2.resize(frame / 255., (out_w, out_h))` (INTER_LINEAR on
 * float64, then the fp32 feed): src [B,Hs,Ws,3] uint8 -> dst [B,h,w,3] fp32.  OpenCV's arithmetic
 * restated in double; equal to the cv2 result after the fp32 cast (tests compare against cv2).   */
int dvsg_frames_u8_resize_to_f32(const unsigned char* src, float* dst, int B, int Hs, int Ws, int h,
                                 int w, int swap_rb, void* stream);
/* Flow ingest (data_loader.py:239): `cv2.resize(np.load(flow_file), (w, h)) * [w, h]` -- a float32 [B,Hs,Ws,2] field of
 * normalised displacements -> [B,h,w,2] in pixels.  cv2's INTER_LINEAR on float32 restated in float (compared with cv2
 * itself in the tests: bit-identical), then the product with (w, h) rounded once.  Device pointers, 8-byte aligned.  */
int dvsg_flow_resize_scale(const float* src, float* dst, int B, int Hs, int Ws, int h, int w,
                           void* stream);
/* Host pipeline with uint8 frames on the host side: U_host, out_host [B,H,W,3] uint8; ingest and
 * egress run on the device, so a frame crosses PCIe as 3 B/pixel each way.  C must be 3.     */
int  dvsg_host_tps_warp_u8(dvsg_host_pipeline* p, const unsigned char* U_host,
                           const float* coord_host, const float* vector_host,
                           unsigned char* out_host, int B, int swap_rb);


/* ---- N3: consumers of the warp outputs in the training graph -------------------------------------
 * dvsg_tps_eval_points evaluates the spline of dvsg_tps_warp_fwd at the flat pixel indices
 * idx [B,P] (idx = col + row*ow; idx == oh*ow gives -1) instead of gathering the dense grid as
 * get_surf_loss does (trainer.py:363-386, tf.batch_gather at :379-380): x_out, y_out [B,P] equal
 * x[b*oh*ow + idx] of the dense kernel bit for bit.  _bwd OVERWRITES grad_T [B,2,pn+3].
 * dvsg_masked_mse_* replace Trainer.masked_MSE (trainer.py:232-243): per frame
 * sq = sum((pred*mask - gt*mask)^2), msum = sum(mask); loss = mean_b div_no_nan(sq, msum).
 * mask_channels is C or 1; the workspace holds the stage-1 partial sums (deterministic order).    */
int dvsg_tps_eval_points(const float* coord, long long coord_batch_stride, const float* T,
                         const int* idx, float* x_out, float* y_out, int B, int oh, int ow,
                         int pn, int P, void* stream);
int dvsg_tps_eval_points_bwd(const float* coord, long long coord_batch_stride, const int* idx,
                             const float* grad_x, const float* grad_y, float* grad_T, int B,
                             int oh, int ow, int pn, int P, void* stream);
size_t dvsg_masked_mse_workspace_bytes(int B, long long n_per_frame);
int dvsg_masked_mse_fwd(const float* pred, const float* gt, const float* mask, int mask_channels,
                        float* sq_out, float* msum_out, float* loss_out, void* workspace,
                        size_t workspace_bytes, int B, long long n_pixels, int C, void* stream);
int dvsg_masked_mse_bwd(const float* pred, const float* gt, const float* mask, int mask_channels,
                        const float* sq, const float* msum, float grad_loss, float* grad_pred,
                        float* grad_gt, float* grad_mask, int B, long long n_pixels, int C,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DVSG_WARP_H_ */
