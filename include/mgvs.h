/*
 * mgvs.h -- C ABI of the B200-native view-synthesis loss (libmgvs.so, built for sm_100a).
 *
 * The reference (uulm-mrm/MGNet) has no FFI for this path: the loss is a Python nn.Module that is
 * constructor-injected into the depth head (mgnet/modeling/mg_net.py:744,757,772-779) and called as
 * self.loss(predictions, targets) (mg_net.py:827-829).  The entry points below are what a binding
 * for that call binds; each cites the reference code it replaces.  INTEGRATION.md shows the
 * ctypes stub a maintainer adds on the reference side.
 *
 * Conventions: plain pointers and sizes only (no torch types); every pointer is DEVICE memory unless
 * the name ends in _host; all tensors are contiguous NCHW fp32 as guaranteed by
 * custom_fwd(cast_inputs=torch.float32) (mg_net.py:827); the library never allocates or frees device
 * memory, never synchronises the device and launches only on the stream it is given; every call is
 * CUDA-graph capturable.  Return value: 0 on success, a negative MGVS_E* code otherwise, with a
 * thread-local message retrievable through mgvs_last_error().
 */
#ifndef MGVS_H_
#define MGVS_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MGVS_ABI_VERSION 7   /* v2 image_dtype, v3 stash, v4 inv_height/inv_width, v5 padding modes, ssim_weight == 0, DGC, uncertainty, peer exchange,
                                v6 pose_mats, camera_lift of mgvs_view_synthesis_ex, exchange status word,
                                v7 mgvs_unpack_mask (bit-packed reprojection mask), mgvs_pose_tail_*, mgvs_workspace_bytes_ex2 (upsample pre-pass) */
#define MGVS_MAX_SCALES 8   /* n: number of inverse-depth maps (reference: 3, mg_net.py:760-764) */
#define MGVS_NUM_SOURCES 2  /* S: prev, next -- hard-coded in the reference (loss.py:116) */

enum { MGVS_IMAGE_F32 = 0, MGVS_IMAGE_U8 = 1 };

enum {
    MGVS_OK = 0,
    MGVS_EINVAL = -1,       /* bad dims / null pointer / misaligned pointer */
    MGVS_EUNSUPPORTED = -2, /* legal in the reference but not implemented by this build (e.g. fused upsample with W % 4 != 0) */
    MGVS_EWORKSPACE = -3,   /* workspace too small */
    MGVS_ECUDA = -4         /* launch failure */
};

/* One problem instance == one call of MultiViewPhotometricLoss.forward (loss.py:111-154). */
typedef struct MgvsProblem {
    int B, H, W, n;
    const void *target;                       /* targets["image_orig"]       [B,3,H,W]  (loss.py:125);
                                                 float, or uint8 when image_dtype == MGVS_IMAGE_U8 */
    const void *source[MGVS_NUM_SOURCES];     /* image_prev_orig, image_next_orig       (loss.py:116) */
    const float *inv_depth[MGVS_MAX_SCALES];  /* predictions["depth"][i]     [B,1,H,W]  (loss.py:112) */
    const float *camera;                      /* targets["camera_matrix"]; element (b,r,c) at
                                                 camera[b*cam_batch_stride + r*cam_row_stride + c];
                                                 only [:, :3, :3] is read       (loss.py:122-123) */
    long long cam_batch_stride, cam_row_stride;
    const float *poses;                       /* predictions["poses"]        [B,S,6] (tx,ty,tz,rx,ry,rz)
                                                 target->source, Euler        (loss.py:117-119); may be NULL
                                                 when pose_mats is given */
    const unsigned char *mask;                /* targets["reprojection_mask"] [B,1,H,W] bool, or NULL
                                                 (loss.py:147, 237-238) */
    /* MultiViewPhotometricLoss.__init__ arguments (loss.py:87-109, defaults config.py:108-117) */
    float ssim_weight;          /* float32(ssim_loss_weight); 0 selects the reference's raw 3-channel L1 branch (loss.py:195-196):
                                   the min runs over 3 channels per list entry, sel = entry * 3 + channel, no stash */
    float one_minus_ssim_weight;/* float32(1 - ssim_loss_weight) evaluated in double like Python does */
    float photometric_weight;
    float smoothing_weight;
    int automask;               /* automask_loss */
    int reduce_op;              /* 0 = "min".  "mean" (loss.py:242-243) is composed by the caller from two "min" evaluations that each see one
                                   source frame in both slots (mgnet_b200/loss.py _forward_mean); any other value is refused */
    int padding_mode;           /* grid_sample padding_mode (camera_utils.py:52-54): 0 = "zeros", 1 = "border", 2 = "reflection" */
    void *workspace;            /* >= mgvs_workspace_bytes(B,H,W,n) bytes, 256-byte aligned; must stay
                                   untouched between mgvs_forward and the matching mgvs_backward */
    size_t workspace_bytes;
    int image_dtype;            /* MGVS_IMAGE_F32: the three images are float in [0,1] (what the loss receives in the
                                   reference).  MGVS_IMAGE_U8: they are the uint8 images the data loader produced and
                                   the library applies the caller's own conversion `x.float() / 255.0`
                                   (mg_net.py:320-335) on the fly -- one correctly rounded division, so every
                                   result is bit-identical to the float path; host->device traffic drops 4x. */
    void *stash;                /* optional, >= mgvs_stash_bytes(B,H,W,n) bytes, 256-byte aligned, or NULL.
                                   Non-NULL selects the stash backward: mgvs_forward additionally writes the three
                                   coefficients of the closed-form SSIM adjoint of the selected source per
                                   (scale, channel, pixel) (48 B/px/scale) and mgvs_backward consumes them instead of
                                   recomputing the warps and SSIM statistics (~2.5x fewer instructions; the path is
                                   issue-bound, HBM is idle).  NULL keeps the recompute backward (nothing but `sel`
                                   and the sums carried over).  Must stay untouched between forward and backward. */
    size_t stash_bytes;
    int inv_height[MGVS_MAX_SCALES]; /* fused head-side upsample: all zero = inv_depth[i] are full resolution */
    int inv_width[MGVS_MAX_SCALES];  /* (the reference's contract).  Otherwise inv_depth[i] is the depth head's low-resolution
                                   map [B,1,inv_height[i],inv_width[i]] BEFORE its F.interpolate(scale_factor=stride,
                                   mode="bilinear", align_corners=True) (mg_net.py:803-806), with H = h*s and W = w*s for an
                                   integer s: mgvs_forward applies that upsample itself (an HBM-bound pre-pass into the workspace,
                                   bit-identical to ATen's CPU kernel; both big kernels then read the full-resolution maps by TMA) and
                                   mgvs_backward returns grad_inv[i] at the LOW resolution (the adjoint of the upsample is applied in
                                   fixed order: deterministic, unlike ATen's atomics on CUDA).  All maps must be low resolution or
                                   none; needs the stash (mgvs_stash_bytes_ex(..., 1)), a workspace sized by
                                   mgvs_workspace_bytes_ex2(..., 1) and W % 4 == 0. */
    const float *pose_mats;     /* optional [B,S,3,4] row-major (R|t): rows 0..2 of the 4x4 the reference builds with
                                   Pose.from_vec -> pose_vec2mat (pose.py:41-47, pose_utils.py:41-51).  Non-NULL replaces the in-kernel
                                   Euler evaluation of `poses`: a caller that forms R with torch gets the reference's own rotation bits
                                   (torch-CPU sin/cos are MKL-VML values, 1 ulp off the correctly rounded ones the kernel uses for ~5 %
                                   of arguments), so the selection mask is bit-exact for ANY angles.  mgvs_backward then writes
                                   grad_poses as [B,S,3,4] = dL/d(R|t) (no Euler chain; autograd of the caller's pose_vec2mat does it). */
} MgvsProblem;

int mgvs_abi_version(void);
const char *mgvs_last_error(void);

/* Scratch the caller must provide (per-tile partial sums, camera table, pose-gradient partials). */
size_t mgvs_workspace_bytes(int B, int H, int W, int n);
/* Same for a given image_dtype (uint8 ingestion keeps float copies of the three images in the workspace). */
size_t mgvs_workspace_bytes_ex(int B, int H, int W, int n, int image_dtype);
/* Same with fused_upsample != 0: room for the n full-resolution inverse-depth maps the upsample pre-pass of mgvs_forward writes. */
size_t mgvs_workspace_bytes_ex2(int B, int H, int W, int n, int image_dtype, int fused_upsample);

/* Size of the optional coefficient stash (MgvsProblem.stash). */
size_t mgvs_stash_bytes(int B, int H, int W, int n);
/* Same with fused_upsample != 0: room for the full-resolution depth gradients the upsample adjoint consumes. */
size_t mgvs_stash_bytes_ex(int B, int H, int W, int n, int fused_upsample);

/* Number of doubles in the partial-sum vector: 3n+3 =
 *   [0,n)     sum over masked pixels of the per-pixel minimum photometric loss, per scale (loss.py:245)
 *   [n]       N   = number of true mask entries                                      (loss.py:245)
 *   [n+1,2n+1)  sum of mask*|d/dx d_hat|*w_x per scale, already divided by the per-image mean (depth.py:48-51)
 *   [2n+1,3n+1) same in y
 *   [3n+1]    N_x, [3n+2] N_y                                                        (loss.py:285-286)
 * These are the only values that cross GPUs: the caller all-reduces (sum) this vector between
 * mgvs_forward and mgvs_finalize when the batch is sharded. */
int mgvs_num_sums(int n);

/* Fused forward: K^-1 back-projection, SE(3), projection, bilinear sampling of both sources, SSIM+L1,
 * identity automask, per-pixel min, smoothness.  Replaces loss.py:111-149 + geometry/*.
 *   sel  [n,B,H,W] uint8 out (may be NULL): argmin index in the reference's list order
 *        [warp_prev, id_prev, warp_next, id_next] (automask) or [warp_prev, warp_next]; ties -> lowest.
 *        With ssim_weight == 0 every list entry has 3 channels and the index is entry * 3 + channel (0..11).
 *   sums [3n+3] double out: this rank's partial sums (see above). */
int mgvs_forward(const MgvsProblem *p, unsigned char *sel, double *sums, void *cuda_stream);

/* Single-rank convenience: mgvs_forward followed by mgvs_finalize in the same launches (the last block of the
 * reduction also writes the two losses).  losses [2] float out; NULL behaves exactly like mgvs_forward. */
int mgvs_forward_losses(const MgvsProblem *p, unsigned char *sel, double *sums, float *losses, void *cuda_stream);

/* Turns (globally reduced) sums into the two weighted scalars of loss.py:151-154.
 *   losses [2] float out: loss_photometric, loss_smoothness. */
int mgvs_finalize(const MgvsProblem *p, const double *sums, float *losses, void *cuda_stream);

/* Fused backward.  Replaces the autograd replay of the loss graph.  With p->stash == NULL it recomputes the
 * forward from the same tiles; with the stash the forward filled it runs the box adjoint + per-output chain only.
 *   sel, sums   as produced by mgvs_forward (sums after the all-reduce if sharded)
 *   g_losses    [2] float, upstream gradients of (loss_photometric, loss_smoothness)
 *   grad_inv[i] [B,1,H,W] float out, fully overwritten
 *   grad_poses  [B,S,6] float out ([B,S,3,4] when p->pose_mats is given), fully overwritten; deterministic (no float atomics) */
int mgvs_backward(const MgvsProblem *p, const unsigned char *sel, const double *sums, const float *g_losses,
                  float *const *grad_inv, float *grad_poses, void *cuda_stream);

/* ---- mgnet.geometry primitives (forward only: the Python mirror raises when an input requires grad -- gradients of the
 * training path flow through mgvs_backward) ------------------------------------------------------------------------- */

/* view_synthesis (camera_utils.py:24-54): warped[B,3,H,W] = grid_sample(ref_image, project(reconstruct(depth))).
 * depth is METRIC depth (already 1/inv); pose34 is [B,3,4] (R|t) of ref_cam.Tcw, K as in MgvsProblem. */
int mgvs_view_synthesis(int B, int H, int W, const float *ref_image, const float *depth, const float *camera,
                        long long cam_batch_stride, long long cam_row_stride, const float *pose34,
                        float *warped, float *coords /* [B,H,W,2] or NULL */, void *cuda_stream);
/* Same with grid_sample's padding_mode (camera_utils.py:52-54): 0 zeros, 1 border, 2 reflection, and with the two cameras
 * of the reference's signature kept apart: `camera` = ref_cam.K projects (camera_utils.py:50), `camera_lift` = cam.K
 * back-projects (camera_utils.py:48; NULL = the same matrix).  They differ when one of the cameras was Camera.scaled(). */
int mgvs_view_synthesis_ex(int B, int H, int W, const float *ref_image, const float *depth, const float *camera,
                           long long cam_batch_stride, long long cam_row_stride, const float *camera_lift,
                           long long lift_batch_stride, long long lift_row_stride, const float *pose34, int padding_mode,
                           float *warped, float *coords /* [B,H,W,2] or NULL */, void *cuda_stream);

/* Camera.reconstruct(depth, frame="c") (camera.py:107-136): points[B,3,H,W] = (K^-1 grid) * depth. */
int mgvs_reconstruct(int B, int H, int W, const float *depth, const float *camera, long long cam_batch_stride,
                     long long cam_row_stride, float *points, void *cuda_stream);

/* Camera.project(X, frame) (camera.py:143-182): coords[B,H,W,2] normalised to [-1,1] (align_corners=True);
 * pose34 [B,3,4] = Tcw for frame "w", NULL for frame "c". */
int mgvs_project(int B, int H, int W, const float *points, const float *camera, long long cam_batch_stride,
                 long long cam_row_stride, const float *pose34, float *coords, void *cuda_stream);


/* ---- Sharded batch: peer-memory exchange of the partial sums ----------------------------------------
 * The only inter-GPU step of the path is the sum of 3n+3 doubles.  Instead of an NCCL all-reduce between mgvs_forward
 * and mgvs_finalize, ONE single-CTA kernel pushes this rank's sums into every peer's exchange buffer with NVLink P2P
 * stores, raises a flag there, waits for the other ranks' flags on its own buffer, adds the world's vectors in rank order
 * (deterministic, bit-identical on every rank) and writes the two losses -- reduction, exchange and finalize in one launch,
 * no host involvement, no NCCL kernel.  The exchange buffers are symmetric allocations the caller maps into every rank
 * (torch.distributed._symmetric_memory / cuMem* + fabric or POSIX handles); peer_base[r] is rank r's buffer as seen from
 * THIS rank.  A buffer holds a step counter, two generations of flags and two generations of world x 32 doubles; it must be
 * zero-filled once, before the first call on any rank (barrier in between), and every rank must make the same sequence of
 * calls.  The reference under DDP has no such exchange (each rank normalises by its local mask count). */
#define MGVS_MAX_RANKS 16
/* Byte offset, in the LOCAL exchange buffer, of a 64-bit status word: 0 while every exchange completed; otherwise the step
 * number of the first call whose peers did not all arrive within the wait bound (>= 40 s; unequal call sequences across
 * ranks).  That call writes NaN into `losses` and returns normally -- no device trap, the CUDA context stays usable. */
#define MGVS_EXCHANGE_STATUS_OFFSET 8
typedef struct MgvsPeerExchange {
    int rank, world;                    /* 1 <= world <= MGVS_MAX_RANKS */
    void *peer_base[MGVS_MAX_RANKS];    /* >= mgvs_exchange_bytes() each, 16-byte aligned; peer_base[rank] is the local one */
    unsigned long long max_spins;       /* wait bound in 20 ns polls of the peers' flags; 0 = the default 2^31 (>= 40 s).  Tests use a small
                                           value to exercise the timeout path. */
} MgvsPeerExchange;
size_t mgvs_exchange_bytes(void);
/*   sums [3n+3] double, in: this rank's partial sums (mgvs_forward); out: the global sums (what mgvs_backward takes)
 *   losses [2] float out (as mgvs_finalize) */
int mgvs_exchange_finalize(const MgvsProblem *p, const MgvsPeerExchange *x, double *sums, float *losses, void *cuda_stream);

/* ---- DGC depth rescaling at inference time -------------------------------------------------------
 * Replaces get_depth_prediction(depth_logits, use_dgc_scaling, camera_matrix, real_camera_height, panoptic_seg,
 * road_class_id, depth_filter_class_ids) (mgnet/postprocessing/depth_post_proc.py:11-71) together with its helpers
 * _get_scale_recovery (:74-104), _get_surface_normal (:107-151), _get_ground_mask (:154-185) and
 * Camera.reconstruct(frame="c") (mgnet/geometry/camera.py:107-136); exportable_post_proc.py:52-79 is the same
 * arithmetic with the inverse camera matrix handed in (camera_is_inverse).  Results are bit-identical to the
 * reference on CPU (points, normals, camera heights, ground mask, median, scale factor).  The reference handles one
 * image per call; B > 1 runs B independent images (one median each). */
#define MGVS_DGC_MAX_FILTER 16
enum { MGVS_PANOPTIC_NONE = 0, MGVS_PANOPTIC_I64 = 1, MGVS_PANOPTIC_I32 = 2 };

typedef struct MgvsDgcProblem {
    int B, H, W;                    /* H, W >= 3 (the normals use a 3x3 neighbourhood) */
    float *depth;                   /* depth_logits [B,1,H,W], rescaled IN PLACE like `depth_logits *= scale_factor`
                                       (depth_post_proc.py:58) */
    const float *camera;            /* camera_matrix; element (b,r,c) at camera[b*cam_batch_stride + r*cam_row_stride + c] */
    long long cam_batch_stride, cam_row_stride;
    int camera_is_inverse;          /* 0: K, the library forms Camera.Kinv (camera.py:72-81); 1: K^-1 given
                                       (exportable_post_proc.py:66-68) */
    const float *real_camera_height;/* device, element b at real_camera_height[b*height_stride] (stride 0 = shared) */
    long long height_stride;
    const void *panoptic;           /* panoptic_seg [B,H,W] int64 / int32, or NULL */
    int panoptic_dtype;             /* MGVS_PANOPTIC_* */
    int use_dgc;                    /* use_dgc_scaling; 0 = only the class filter runs (needs panoptic) */
    long long road_class_id;        /* ground mask = panoptic == road_class_id; with panoptic == NULL the ground mask
                                       comes from the surface normals (depth_post_proc.py:154-185) */
    long long filter_ids[MGVS_DGC_MAX_FILTER]; /* depth_filter_class_ids: depth -> 0, points -> NaN where panoptic == id */
    int n_filter;
    float *points;                  /* cam_xyz_points [B,3,H,W] out (already scaled), or NULL */
    float *scale;                   /* [B] out: scale_factor; NaN when the ground mask is empty or holds a NaN height
                                       (what torch.median yields in the reference) */
    long long *count;               /* [B] out or NULL: 0 iff the ground mask of image b is empty */
    void *workspace;                /* >= mgvs_dgc_workspace_bytes(B,H,W), 256-byte aligned */
    size_t workspace_bytes;
} MgvsDgcProblem;

size_t mgvs_dgc_workspace_bytes(int B, int H, int W);
/* 5 stream operations (1 memset + 4 kernels), no host synchronisation, graph capturable. */
int mgvs_dgc_rescale(const MgvsDgcProblem *p, void *cuda_stream);
/* Diagnostics for the parity tests: per-pixel camera heights [B,H,W] (|P.N|, depth_post_proc.py:96) and the ground
 * mask [B,H,W] uint8 that the median runs over.  Same inputs as mgvs_dgc_rescale; depth is not modified. */
int mgvs_dgc_heights(const MgvsDgcProblem *p, float *heights, unsigned char *ground, void *cuda_stream);

/* ---- Homoscedastic uncertainty weighting of the task losses ---------------------------------------
 * Replaces the epilogue of MGNet.forward (mgnet/modeling/mg_net.py:360-372):
 *     losses[key] = tau * exp(-log_vars[idx]) * value + 0.5 * log_vars[idx]      tau = 1.0 for "loss_sem_seg" else 0.5
 * and its two `.item()` host synchronisations per loss for the event storage (key + "_raw", key + "_uncertainty"):
 * the logging copies are written to device memory instead.  k <= 16 losses; tau_host is a HOST array. */
#define MGVS_MAX_LOSSES 16
/*   raw [k], log_vars [k] in;  weighted [k] out;  log_out [2k] out or NULL: raw values, then exp(log_vars) */
int mgvs_uncertainty_forward(int k, const float *raw, const float *log_vars, const float *tau_host, float *weighted,
                             float *log_out, void *cuda_stream);
/*   g_weighted [k] in;  g_raw [k] = g * tau * exp(-s);  g_log_vars [k] = g * (0.5 - tau * exp(-s) * raw) */
int mgvs_uncertainty_backward(int k, const float *raw, const float *log_vars, const float *tau_host,
                              const float *g_weighted, float *g_raw, float *g_log_vars, void *cuda_stream);

/* PoseCNN tail (reference layers.py:164-166): out[m] = scale * mean over rows of the row means of map m, for `maps` = B * 6 *
 * num_context_images contiguous [h, w] fp32 maps (the output of PoseCNN.conv4); scale = 0.01.  `out` viewed as [B, num_context, 6]
 * is predictions["poses"].  Deterministic (fixed-order fp64 accumulation); replaces three ATen launches forward and three backward
 * with one each.  backward: gx[m, :, :] = g[m] * scale / (h * w). */
int mgvs_pose_tail_forward(int maps, int h, int w, const float *x, float scale, float *out, void *cuda_stream);
int mgvs_pose_tail_backward(int maps, int h, int w, const float *g, float scale, float *gx, void *cuda_stream);

/* Bit-packed reprojection mask -> the byte-per-pixel bool tensor MgvsProblem.mask expects.  `bits` is numpy.packbits(mask, axis=-1):
 * most significant bit first, every image row padded to a whole number of bytes -- `rows` = B*H rows of ceil(W/8) bytes.  The
 * reference's mask is a torch.bool tensor (1 byte per pixel, loss.py:147); a data loader that ships it packed saves 7/8 of its
 * host-to-device bytes, which matters because the end-to-end path is bound by exactly those copies. */
int mgvs_unpack_mask(long long rows, int W, const unsigned char *bits, unsigned char *mask, void *cuda_stream);

/* Self-test hook used by the GPU tests: out[i] = a[i] / b[i] with the library's in-kernel exact division. */
int mgvs_test_div(const float *a, const float *b, float *out, long long count, void *cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* MGVS_H_ */
