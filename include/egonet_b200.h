/*
 * egonet_b200 -- C ABI of the B200-native EgoNet per-crop inference hot path.
 *
 * The upstream project (Nicholasli1995/EgoNet) is pure Python/PyTorch and has no
 * FFI of its own; the boundary it exposes for this path is the Python class
 * surface of libs/model.  This header is the C level
 * inserted one step below that surface: each entry point replaces the body of
 * one reference function, named in its comment (paths relative to the upstream
 * repository root).  INTEGRATION.md shows the ctypes binding a maintainer adds.
 *
 * Conventions
 *   - every function returns EGN_OK (0) or a negative egn_status; the message
 *     of the last failure on the calling thread is egn_last_error().
 *   - all device pointers are caller-owned (e.g. torch tensors); the library
 *     never frees or retains them past the call.  Handles own only their
 *     repacked weights.
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no
 *     entry point synchronises the device except egn_*_create/finalize/destroy
 *     and the *_host convenience calls.
 *   - there is no CPU fallback: on a machine without an sm_100 device every
 *     compute entry returns EGN_ERR_NO_DEVICE.
 */
#ifndef EGONET_B200_H_
#define EGONET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EGN_VERSION 100 /* 0.1.0 */

#if defined(__GNUC__)
#define EGN_API __attribute__((visibility("default")))
#else
#define EGN_API
#endif

typedef enum {
  EGN_OK = 0,
  EGN_ERR_INVALID = -1,   /* bad argument / unsupported configuration        */
  EGN_ERR_NO_DEVICE = -2, /* no CUDA device, or device is not sm_100         */
  EGN_ERR_CUDA = -3,      /* a CUDA runtime/driver call failed               */
  EGN_ERR_STATE = -4,     /* call order violated (e.g. forward before finalize) */
  EGN_ERR_MISSING = -5,   /* a required weight was never set                 */
  EGN_ERR_WORKSPACE = -6  /* workspace too small                              */
} egn_status;

EGN_API int egn_version(void);
EGN_API const char* egn_last_error(void);
/* 1 if cuda:current is an sm_100 (B200) device, 0 otherwise. Never fails. */
EGN_API int egn_device_ok(void);

/* ------------------------------------------------------------------------- */
/* HC: the HRNet heat-map / coordinate network                               */
/* replaces libs/model/heatmapModel/hrnet.py:563-614                         */
/*   (PoseHighResolutionNet.forward and every module it calls)               */
/* ------------------------------------------------------------------------- */

#define EGN_MAX_BRANCHES 4

typedef enum { EGN_HEAD_HEATMAP = 0, EGN_HEAD_COORDINATES = 1 } egn_head_type;

typedef enum {
  EGN_PREC_FP32 = 0,  /* fp32 storage + fp32 CUDA-core convs (exact comparator)              */
  EGN_PREC_FP16 = 1,  /* fp16 NHWC storage, fp32 accumulation, tcgen05 tensor cores        */
  EGN_PREC_FP16X2 = 2 /* split storage (every activation / weight an unevaluated sum of two
                         fp16 values, ~22 significant bits) and error-compensated tcgen05
                         convs: x*w = x_hi*w_hi + x_lo*w_hi + x_hi*w_lo in fp32 TMEM
                         accumulators.  The tensor-core mode that holds the reference's fp32
                         results to the 1e-4 parity bound.                                  */
} egn_precision;

typedef enum {
  EGN_CONV_AUTO = 0, /* tcgen05 kernels wherever the layer shape allows (fp16 / fp16x2) */
  EGN_CONV_SIMT = 1  /* force the CUDA-core kernels (debug comparator)            */
} egn_conv_impl;

/* Mirrors cfgs['heatmapModel'] of the reference YAML (configs/KITTI_inference:demo.yml:72-151). */
typedef struct {
  int in_channels;                 /* 3, or 5 with add_xy (hrnet.py:649-659)          */
  int input_w, input_h;            /* input_size = [w, h]                             */
  int heatmap_w, heatmap_h;        /* heatmap_size = [w, h]                           */
  int num_joints;                  /* 33                                              */
  int head_type;                   /* egn_head_type                                   */
  int final_conv_kernel;           /* extra.final_conv_kernel (heatmap head), 1 or 3  */
  int num_stages;                  /* 3 (stage2..stage4)                              */
  int stage_modules[3];            /* extra.stageN.num_modules                        */
  int stage_branches[3];           /* extra.stageN.num_branches                       */
  int stage_blocks[3][EGN_MAX_BRANCHES];   /* extra.stageN.num_blocks                */
  int stage_channels[3][EGN_MAX_BRANCHES]; /* extra.stageN.num_channels              */
  int precision;                   /* egn_precision                                   */
  int conv_impl;                   /* egn_conv_impl                                   */
  int keep_taps;                   /* 1: never recycle activation buffers so that     */
                                   /*    egn_hrnet_read_tap works after a forward     */
} egn_hrnet_cfg;

typedef struct egn_hrnet egn_hrnet;

EGN_API int egn_hrnet_create(const egn_hrnet_cfg* cfg, egn_hrnet** out);
EGN_API void egn_hrnet_destroy(egn_hrnet* h);

/* Number of state_dict entries the network expects and the i-th key / shape
 * (same names, shapes and order as PoseHighResolutionNet(cfgs).state_dict(),
 * hrnet.py:311-469).  shape receives up to 4 dims; returns ndim. */
EGN_API int egn_hrnet_num_weights(const egn_hrnet* h);
EGN_API const char* egn_hrnet_weight_key(const egn_hrnet* h, int i);
EGN_API int egn_hrnet_weight_shape(const egn_hrnet* h, int i, int64_t shape[4]);

/* Copy one fp32 state_dict tensor (HOST pointer, contiguous, torch OIHW for
 * convs) into the handle.  `*.num_batches_tracked` keys are accepted and
 * ignored.  Invalidates a previous finalize. */
EGN_API int egn_hrnet_set_weight(egn_hrnet* h, const char* key, const float* host_data,
                         const int64_t* shape, int ndim);

/* Fold eval-mode BatchNorm into the convolutions, repack for the kernels,
 * upload.  Fails with EGN_ERR_MISSING (message names the key) if a tensor was
 * never set. */
EGN_API int egn_hrnet_finalize(egn_hrnet* h);

/* Bytes of caller-provided device workspace a forward of `batch` crops needs. */
EGN_API size_t egn_hrnet_workspace_bytes(const egn_hrnet* h, int batch);

/* x: device fp32 [B, in_channels, input_h, input_w] (NCHW, as the reference feeds HC).
 * heatmap_out: device fp32 [B, num_joints, heatmap_h, heatmap_w] or NULL.
 * coords_out : device fp32 [B, num_joints, 2] (coordinate head only) or NULL.
 * logits_out : device fp32 [B, 2*num_joints] pre-sigmoid values or NULL. */
EGN_API int egn_hrnet_forward(egn_hrnet* h, const float* x, int batch, float* heatmap_out,
                      float* coords_out, float* logits_out, void* workspace,
                      size_t workspace_bytes, void* stream);

/* Debug/parity: copy a named intermediate activation of the LAST forward
 * (needs keep_taps=1) to device fp32 NCHW `out`.  Names: "stem1", "stem2",
 * "layer1", "stage<S>.<M>.out<B>", "head2.<K>".  dims receives [C,H,W]. */
EGN_API int egn_hrnet_read_tap(egn_hrnet* h, const char* name, int batch, const void* workspace,
                       float* out, int dims[3], void* stream);

/* One fused conv layer, out = act(conv(in, w) + bias [+ res]), for per-layer parity tests
 * (synchronous).  impl: 0 = CUDA-core kernel, 1 = tcgen05 kernel (fp16 / fp16x2).  dtype: 0 = fp32,
 * 1 = fp16, 2 = fp16x2 split storage ([B,H,W,2*Cp] fp16: hi plane then lo plane per pixel).
 * in/res/out: device NHWC with channels padded to a multiple of 16 (pad lanes zero); w: HOST fp32 torch OIHW; bias: HOST fp32 [Cout] or NULL.
 * ksize 1 (pad 0) or 3 (pad 1), stride 1 or 2. */
EGN_API int egn_conv2d_fused(int impl, int dtype, const void* in, const float* w_oihw_host,
                     const float* bias_host, const void* res, void* out, int B, int H, int W,
                     int Cin, int Cout, int ksize, int stride, int relu, void* stream);

/* Same layer launched `iters` times back to back (after one warm-up launch) with CUDA events around
 * the loop; avg_ms receives the mean device time per launch.  Kernel-tuning utility. */
EGN_API int egn_conv2d_bench(int impl, int dtype, const void* in, const float* w_oihw_host,
                     const float* bias_host, const void* res, void* out, int B, int H, int W,
                     int Cin, int Cout, int ksize, int stride, int relu, void* stream, int iters,
                     float* avg_ms);

/* Debug: one fp16 conv (no residual, no ReLU) that also writes the UN-ROUNDED fp32 result (accumulator
 * + bias) as NCHW [B,Cout,OH,OW] to acc_out -- measures the accumulation error of the tensor-core path
 * against an fp64 reference on fp16-exact operands (tools/probes/acc_precision.py). */
EGN_API int egn_debug_conv_acc(int impl, const void* in, const float* w_oihw_host, const float* bias_host,
                       void* out, float* acc_out, int B, int H, int W, int Cin, int Cout, int ksize,
                       int stride, void* stream);

/* Hardware probe (debug): SM cycles per tcgen05.mma (M=128, K=16, f16, operands in smem) when `ctas`
 * CTAs each issue iters*4*nacc MMAs of width n, rotating over nacc accumulators. */
EGN_API int egn_debug_umma_rate(int n, int nacc, int iters, int a_rows_shift, int ctas, double* cycles_per_mma);

/* Plan of one fused tcgen05 conv as text (debug / tests): which kernel runs the shape (v1-tap / v2-run / v3-persist /
 * v4-tapwin) and how it is tiled.  dtype 1 = fp16, 2 = fp16x2.  Geometry only: no GPU needed. */
EGN_API int egn_debug_conv_plan(int dtype, int Cin, int Cout, int H, int W, int ksize, int stride, char* out, int out_len);

/* Hardware probe (debug): SM cycles per K16 slice of the MMA sequences the fp16x2 conv kernels issue (pattern 0: one
 * N = n MMA; 1: full-width N = 2n + half-width N = n; 2: three N = n MMAs into H, L, L; 3: as 1, grouped per two
 * slices), operands resident in shared memory, A rows of a_sw bytes (128 / 64) with 8-row groups sbo_rows rows apart. */
EGN_API int egn_debug_umma_seq(int n, int pattern, int iters, int a_sw, int sbo_rows, int ctas, double* cycles_per_slice);

/* Hardware probe (debug): one 128 x KC x KC UMMA whose A descriptor starts `row_off` rows into a
 * TMA-written swizzled tile; bo_mode selects the descriptor base-offset encoding under test.
 * a: device fp16 [256][KC], b: device fp16 [KC][KC], out: device fp32 [128][KC], KC = swizzle/2. */
EGN_API int egn_debug_umma_probe(int swizzle_bytes, int row_off, int bo_mode, const void* a_f16,
                         const void* b_f16, float* out, void* stream);

/* Introspection for bench/roofline accounting. */
typedef struct {
  int kind;            /* 0 stem conv, 1 fused conv, 2 cross-resolution fuse, 3 head tail        */
  int use_tc;          /* 1 if this conv runs on the tcgen05 kernel                               */
  int Cin, Cout, H, W, OH, OW, ksize, stride, has_res;
  int64_t macs;        /* per crop                                                                */
  int64_t act_bytes;   /* algorithmic activation bytes per crop (inputs read once + output)       */
  int64_t weight_bytes;/* per launch                                                              */
  char name[96];       /* state_dict key of the conv / tap name of the fuse output                */
} egn_op_info;
EGN_API int egn_hrnet_op_info(const egn_hrnet* h, int i, egn_op_info* out);
/* One forward with a CUDA event between every launch; op_ms[egn_hrnet_num_launches] receives the
 * device time of each op (synchronises the stream). */
EGN_API int egn_hrnet_profile(egn_hrnet* h, const float* x, int batch, void* workspace,
                      size_t workspace_bytes, void* stream, float* op_ms);
EGN_API int64_t egn_hrnet_macs_per_crop(const egn_hrnet* h);
EGN_API int egn_hrnet_num_launches(const egn_hrnet* h);        /* kernels per forward          */
EGN_API int egn_hrnet_num_tc_launches(const egn_hrnet* h);     /* of which tcgen05 convs       */
EGN_API int64_t egn_hrnet_act_bytes_per_crop(const egn_hrnet* h); /* algorithmic activation bytes */
EGN_API int64_t egn_hrnet_weight_bytes(const egn_hrnet* h);

/* ------------------------------------------------------------------------- */
/* Heat-map decoders (device fp32 NCHW heat-maps [B,K,H,W])                  */
/* ------------------------------------------------------------------------- */

/* replaces libs/common/img_proc.py:608-637 get_max_preds.
 * idx [B*K] int32 (flat arg-max, first occurrence on ties) or NULL,
 * preds [B*K*2] fp32 (x = idx % W, y = floor(idx / W); zeroed where max <= 0),
 * maxvals [B*K] fp32. */
EGN_API int egn_argmax2d(const float* hm, int B, int K, int H, int W, int32_t* idx, float* preds,
                 float* maxvals, void* stream);

typedef enum {
  EGN_SOFTARGMAX_SOFTMAX = 0, /* img_proc.py:678-707 soft_arg_max (torch)        */
  EGN_SOFTARGMAX_SUM = 1      /* img_proc.py:639-676 soft_arg_max_np (sum-normalised, max>0 mask) */
} egn_softargmax_mode;

EGN_API int egn_soft_argmax2d(const float* hm, int B, int K, int H, int W, int mode, float* preds,
                      float* maxvals, void* stream);

/* ------------------------------------------------------------------------- */
/* Local -> screen coordinates                                               */
/* replaces the per-instance loop of EgoNet.get_keypoints, egonet.py:436-453 */
/*   (get_affine_transform(inv=1) img_proc.py:26-64 +                        */
/*    affine_transform_modified img_proc.py:71-78)                           */
/* coords [N,K,2] fp32 in (0,1); center/scale [N,2] fp64; rot [N] fp64       */
/* degrees or NULL (= 0); screen [N,K,2] fp64.                               */
/* ------------------------------------------------------------------------- */
EGN_API int egn_local_to_screen(const float* coords, const double* center, const double* scale,
                        const double* rot, int N, int K, int res_w, int res_h,
                        double* screen, void* stream);

/* ------------------------------------------------------------------------- */
/* L: the 2D->3D lifter                                                      */
/* replaces EgoNet.lift_2d_to_3d egonet.py:469-486 (normalize_1d,            */
/* FCModel.forward FCmodel.py:92-105, unnormalize_1d)                        */
/* ------------------------------------------------------------------------- */
typedef struct egn_lifter egn_lifter;

EGN_API int egn_lifter_create(int input_size, int output_size, int num_neurons, int num_blocks,
                      egn_lifter** out);
EGN_API void egn_lifter_destroy(egn_lifter* l);
/* state_dict keys of FCModel (FCmodel.py:72-83): w1.*, batch_norm1.*,
 * res_blocks.<i>.{w1,batch_norm1,w2,batch_norm2}.*, w2.*  (host fp32). */
EGN_API int egn_lifter_set_weight(egn_lifter* l, const char* key, const float* host_data,
                          const int64_t* shape, int ndim);
/* LS statistics (host fp64): mean_in/std_in [input_size], mean_out/std_out [output_size]. */
EGN_API int egn_lifter_set_stats(egn_lifter* l, const double* mean_in, const double* std_in,
                         const double* mean_out, const double* std_out);
EGN_API int egn_lifter_finalize(egn_lifter* l);
EGN_API size_t egn_lifter_workspace_bytes(const egn_lifter* l, int n);
/* kpts_2d device fp64 [n, input_size] -> kpts_3d device fp64 [n, output_size];
 * raw_out (device fp32 [n, output_size], network output before un-normalisation) or NULL. */
EGN_API int egn_lifter_forward(egn_lifter* l, const double* kpts_2d, int n, double* kpts_3d,
                       float* raw_out, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------- */
/* Per-instance pose solve                                                   */
/* replaces EgoNet.get_6d_rep egonet.py:279-295 (get_template :238-263,      */
/* compute_rigid_transform transformation.py:99-134, Rotation.as_euler('yxz')*/
/* egonet.py:274-276) and get_observation_angle_{trans,proj} egonet.py:203-236 */
/* ------------------------------------------------------------------------- */
typedef enum { EGN_ALPHA_TRANS = 0, EGN_ALPHA_PROJ = 1 } egn_alpha_mode;

/* kpts_3d device fp64 [N, P, 3], P = 8 or 32.
 * kpts_2d device fp64 [N, stride_2d] (element 0 of each row = first key-point's
 *   screen x) -- needed for EGN_ALPHA_PROJ only, else NULL.
 * fx, cx: K[0,0], K[0,2] (proj mode).
 * pose_out device fp64 [N, 7] = euler x,y,z | translation x,y,z | alpha.
 * rot_out  device fp64 [N, 9] row-major Kabsch rotation, or NULL. */
EGN_API int egn_pose_solve(const double* kpts_3d, int N, int P, const double* kpts_2d, int stride_2d,
                   double fx, double cx, int alpha_mode, double* pose_out, double* rot_out,
                   void* stream);

/* ------------------------------------------------------------------------- */
/* Crop front-end: decoded image(s) in HBM -> network input */
/* replaces, per box, EgoNet.crop_single_instance egonet.py:68-95 as called    */
/* by crop_instances egonet.py:105-155:                                        */
/*   get_affine_transform(c, s, 0, (h, w)) img_proc.py:26-64,                  */
/*   cv2.warpAffine(img, trans, (w, h), flags=INTER_LINEAR) egonet.py:85-89    */
/*     (OpenCV's 8-bit fixed-point bilinear path, BORDER_CONSTANT 0),          */
/*   transforms.ToTensor + Normalize(mean, std) car_instance.py:522-531.       */
/* center/scale come from modify_bbox (img_proc.py:453-459), as for            */
/* egn_local_to_screen; only scale[:,0] is used (img_proc.py:41-42).           */
/* ------------------------------------------------------------------------- */
typedef struct {
  const uint8_t* data; /* DEVICE pointer, interleaved 8-bit RGB rows (cv2 HWC layout) */
  int32_t height, width;
  int32_t pitch;       /* bytes between rows (>= 3*width)                             */
  int32_t channels;    /* must be 3                                                   */
} egn_image;

/* images_dev: DEVICE array of n_images descriptors; image_of_crop_dev: device int32 [N]
 * (index into images_dev) or NULL (every crop comes from image 0).
 * center_dev/scale_dev: device fp64 [N,2].  mean3_host/std3_host: HOST float[3] or NULL (0 / 1).
 * out_nchw: device fp32 [N,3,res_h,res_w] = Normalize(ToTensor(warpAffine(...))).
 * out_u8 : device uint8 [N,res_h,res_w,3], the warpAffine result itself, or NULL. */
EGN_API int egn_crop_instances(const egn_image* images_dev, int n_images, const int32_t* image_of_crop_dev,
                       const double* center_dev, const double* scale_dev, int N, int res_w, int res_h,
                       const float* mean3_host, const float* std3_host, float* out_nchw,
                       uint8_t* out_u8, void* stream);

/* ------------------------------------------------------------------------- */
/* General point-set alignment, batched       */
/* replaces compute_rigid_transform libs/common/transformation.py:99-134,      */
/* procrustes_transform :136-141 and compute_similarity_transform :48-97.      */
/* All pointers device fp64; point sets are [N, P, 3] (one instance = P points */
/* row-major; the reference's [3, P] arrays transposed).                       */
/* ------------------------------------------------------------------------- */
/* Y ~ R X + t.  W: NULL (w_mode 0), per-point weights [N,P] (w_mode 1) or full [N,P,P] matrices (w_mode 2);
 * centroids are unweighted means, as upstream.  R_out [N,9] row-major, t_out [N,3], aligned_out [N,P,3]
 * = R X + t (procrustes_transform); each may be NULL. */
EGN_API int egn_rigid_transform(const double* X, const double* Y, const double* W, int w_mode, int N, int P,
                        double* R_out, double* t_out, double* aligned_out, void* stream);
/* MATLAB-style procrustes: X targets, Y inputs.  d_out [N], b_out [N], Z_out [N,P,3] (transformed Y),
 * T_out [N,9] (Z = b * Y T + c), c_out [N,3]; each may be NULL. */
EGN_API int egn_similarity_transform(const double* X, const double* Y, int N, int P, int compute_optimal_scale,
                             double* d_out, double* b_out, double* Z_out, double* T_out, double* c_out,
                             void* stream);

/* ------------------------------------------------------------------------- */
/* Reprojection refinement                                */
/* replaces pnp_refine libs/common/transformation.py:143-157:                  */
/*   cv2.solvePnP(prediction, observation, K, dist, SOLVEPNP_ITERATIVE)        */
/*   (DLT start + Levenberg-Marquardt, OpenCV calib3d) then                    */
/*   Rodrigues(R) @ prediction.T + T, for N instances in one launch.           */
/* kpts_3d device fp64 [N,P,3] (camera frame), kpts_2d device fp64 [N,P,2]     */
/* (pixels), 6 <= P <= 64; zero lens distortion; max_iter 0 = OpenCV's 20.     */
/* refined device fp64 [N,P,3]; pose6 [N,6] = rvec | tvec or NULL; info [N,2]  */
/* = accepted LM iterations, final residual norm (px) or NULL; status int32    */
/* [N] (egn_pnp_status) or NULL.  Instances that are planar / degenerate are   */
/* returned unrefined (upstream keeps the prediction when solvePnP fails).     */
/* ------------------------------------------------------------------------- */
typedef enum { EGN_PNP_STATUS_OK = 0, EGN_PNP_STATUS_PLANAR = 1, EGN_PNP_STATUS_DEGENERATE = 2 } egn_pnp_status;

/* replaces refine_with_predicted_bbox tools/inference_legacy.py:518-547 for N instances: pred_rel [N,P,3] with
 * points 1.. relative to point 0; refined [N,P,3] (absolute); ok int32 [N] = 1 when the refined root stayed
 * within `threshold` of the predicted one (upstream returns (False, None) otherwise; refined then holds the
 * absolute unrefined box); status as egn_pnp_refine or NULL. */
EGN_API int egn_refine_with_bbox(const double* pred_rel, const double* kpts_2d, int N, int P, double fx, double fy,
                         double cx, double cy, double threshold, int max_iter, double* refined, int32_t* ok,
                         int32_t* status, void* stream);

EGN_API int egn_pnp_refine(const double* kpts_3d, const double* kpts_2d, int N, int P, double fx, double fy,
                   double cx, double cy, int max_iter, double* refined, double* pose6, double* info,
                   int32_t* status, void* stream);

/* ------------------------------------------------------------------------- */
/* Heat-map MSE loss (training config only; loss end of SURVEY.md 8a row a12)   */
/* replaces JointsMSELoss.forward libs/loss/function.py:28-46 and              */
/* JointsCompositeLoss.calc_hm_loss libs/loss/function.py:95-111               */
/* pred/target device fp32 [B,K,H,W]; target_weight device fp32 [B,K] or NULL;  */
/* loss_out device fp32 scalar; grad_out device fp32 [B,K,H,W] (d loss/d pred)  */
/* or NULL; workspace8: 8 bytes of device scratch.                              */
/* ------------------------------------------------------------------------- */
EGN_API int egn_mse_hm_fwd_bwd(const float* pred, const float* target, const float* target_weight, int B, int K,
                       int H, int W, float* loss_out, float* grad_out, void* workspace8, void* stream);

/* ------------------------------------------------------------------------- */
/* Training step of HC (BASELINE configs[3])             */
/* replaces the loop body of libs/trainer/trainer.py:183-198 for the heat-map  */
/* head: model.train() forward (hrnet.py:563-614 with nn.BatchNorm2d in        */
/* training mode, hrnet.py:63-133, 282-300), loss.backward() through it, and   */
/* the optimiser update (libs/optimizer/optimizer.py:9-41).  fp32 throughout   */
/* (the reference trains in fp32), BatchNorm sums in fp64.                     */
/*                                                                             */
/* Parameters and gradients are ONE flat fp32 device buffer each, owned by the */
/* caller, laid out in state_dict order: entry i of egn_hrnet_weight_key       */
/* starts at element egn_hrnet_train_param_offset(t, i) (-1 for the int64      */
/* num_batches_tracked entries, which the caller keeps).                       */
/* ------------------------------------------------------------------------- */
typedef struct egn_hrnet_train egn_hrnet_train;

/* cfg as for egn_hrnet_create (precision / conv_impl / keep_taps ignored); both head types. */
EGN_API int egn_hrnet_train_create(const egn_hrnet_cfg* cfg, egn_hrnet_train** out);
EGN_API void egn_hrnet_train_destroy(egn_hrnet_train* t);
EGN_API int64_t egn_hrnet_train_flat_size(const egn_hrnet_train* t);            /* floats in the flat buffer */
EGN_API int64_t egn_hrnet_train_param_offset(const egn_hrnet_train* t, int i);
EGN_API int egn_hrnet_train_param_trainable(const egn_hrnet_train* t, int i);  /* 1 parameter, 0 buffer    */
EGN_API size_t egn_hrnet_train_workspace_bytes(const egn_hrnet_train* t, int batch);
EGN_API int64_t egn_hrnet_train_flops_per_sample(const egn_hrnet_train* t);    /* 3 x 2 x conv MACs        */

/* Train-mode forward.  flat_params: device fp32 (running statistics inside it are updated in place with
 * `momentum` when update_running_stats != 0, as nn.BatchNorm2d does); x: device fp32 NCHW [B,C,H,W];
 * heatmap_out: device fp32 [B,K,hh,hw]; coords_out: device fp32 [B,K,2] (coordinate head) or NULL.
 * The workspace keeps every activation for the backward pass. */
EGN_API int egn_hrnet_forward_train(egn_hrnet_train* t, float* flat_params, const float* x, int batch,
                            float* heatmap_out, float* coords_out, float momentum, int update_running_stats,
                            void* workspace, size_t workspace_bytes, void* stream);
/* Backward of the last forward (same batch, same workspace).  grad_heatmap: device fp32 [B,K,hh,hw] =
 * d loss / d heatmap_out or NULL; grad_coords: device fp32 [B,K,2] = d loss / d coords_out or NULL (at least one);
 * flat_grads: device fp32, overwritten with d loss / d parameter (flat layout). */
EGN_API int egn_hrnet_backward(egn_hrnet_train* t, const float* flat_params, const float* grad_heatmap,
                       const float* grad_coords, int batch, float* flat_grads, void* workspace,
                       size_t workspace_bytes, void* stream);

/* Optimiser updates over flat device buffers with torch.optim semantics (optimizer.py:19-27): step counts from 1;
 * trainable_mask (device uint8 [n], NULL = all) skips frozen entries / buffers. */
EGN_API int egn_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                  const uint8_t* trainable_mask, int64_t n, int step, float lr, float beta1, float beta2, float eps,
                  float weight_decay, void* stream);
EGN_API int egn_sgd_step(float* params, const float* grads, float* momentum_buf, const uint8_t* trainable_mask,
                 int64_t n, int step, float lr, float momentum, float weight_decay, int nesterov, void* stream);

/* Coordinate (L_2d) and cross-ratio (L_cr) terms of JointsCompositeLoss, forward + gradient in one launch.
 * replaces calc_coor_loss libs/loss/function.py:159-168, calc_cross_ratio_loss :113-138 (appro_cr
 * libs/common/img_proc.py:709-720) and get_cr_mask :140-153.
 * coords_pred device fp32 [B,K,2] in (0,1); coords_gt_px device fp32 [B,K,2] in crop pixels (divided by img_w /
 * img_h inside, as upstream); criterion kinds 0 = mse, 1 = smooth L1, 2 = L1 (loss_dict :16-19), mean reduction;
 * cr_indices device int32 [L,4] point indices of each line (cr_indices_dict['bbox12'], car_instance.py:83-97) or
 * NULL; lines whose smallest non-zero pairwise distance is <= cr_threshold are masked out.
 * loss_out device fp32 [3] = {coor_weight * coor + cr_weight * cr, coor, cr}; grad_out device fp32 [B,K,2] =
 * d loss_out[0] / d coords_pred, or NULL. */
EGN_API int egn_coord_loss_fwd_bwd(const float* coords_pred, const float* coords_gt_px, int B, int K, float img_w,
                           float img_h, int coor_kind, float coor_weight, const int32_t* cr_indices, int L,
                           int cr_kind, float cr_weight, float target_cr, float cr_threshold, float* loss_out,
                           float* grad_out, void* stream);

/* ------------------------------------------------------------------------- */
/* KITTI object-evaluation overlaps, all pairs of a frame */
/* replaces groundBoxOverlap / box3DOverlap / imageBoxOverlap of               */
/* tools/kitti-eval/evaluate_object_3d_offline.cpp:224-344 (Boost.Geometry     */
/* polygon intersection of the oriented ground-plane rectangles).              */
/* det device fp64 [D,7], gt device fp64 [G,7], rows = ry, h, w, l, t1 (x),    */
/* t2 (y of the box bottom), t3 (z); criterion -1 union, 0 / detection area,   */
/* 1 / ground-truth area.  ground_out / box3d_out device fp64 [D,G] or NULL.   */
/* ------------------------------------------------------------------------- */
EGN_API int egn_box_overlaps(const double* det, const double* gt, int D, int G, int criterion, double* ground_out,
                     double* box3d_out, void* stream);
/* det / gt device fp64 [D,4] / [G,4] = x1, y1, x2, y2 -> out [D,G] */
EGN_API int egn_image_box_overlaps(const double* det, const double* gt, int D, int G, int criterion, double* out,
                           void* stream);

/* Gaussian heat-map targets of the training configuration (BASELINE configs[3]).
 * replaces generate_target libs/common/img_proc.py:347-409 (target_type 'gaussian'), one launch for N samples.
 * joints device fp64 [N,K,3] (crop pixels; column 2 = visibility when joints_vis is NULL), joints_vis device
 * fp32 [N,K] or NULL; input_size / heatmap_size as the reference's config lists ([0], [1]): the target is
 * allocated [K, heatmap_size[0], heatmap_size[1]] exactly as upstream does.
 * target device fp32 [N,K,heatmap_size0,heatmap_size1]; target_weight device fp32 [N,K] or NULL. */
EGN_API int egn_generate_target(const double* joints, const float* joints_vis, int N, int K, int input_size0,
                        int input_size1, int heatmap_size0, int heatmap_size1, double sigma, float* target,
                        float* target_weight, void* stream);

/* replaces the loops of EgoNet.get_observation_angle_trans / _proj (egonet.py:203-236):
 * alpha[n] = wrap(ry[n] - atan2(-z[n*stride_z], x[n*stride_x] - x_offset) - pi/2).
 * trans: x = translation[:,0], z = translation[:,2], x_offset = 0.
 * proj : x = first key-point's screen x, x_offset = K[0,2], z = &K[0,0] with stride_z = 0.
 * All pointers device fp64. */
EGN_API int egn_observation_angle(const double* ry, const double* x3d, int stride_x, const double* z3d,
                          int stride_z, double x_offset, int N, double* alpha, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EGONET_B200_H_ */
