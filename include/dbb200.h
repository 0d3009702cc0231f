/*
 * dbb200.h -- C ABI of libdbb200.so: the B200 (sm_100a) kernels behind the DB_text_minimal hot path.
 *
 * The reference (huyhoang17/DB_text_minimal) is pure Python/PyTorch and has no FFI layer; its
 * boundary for this path is three Python callables:
 *     DBTextModel.forward          src/models.py:34-48
 *     DBLoss.forward               src/losses.py:105-139
 *     SegDetectorRepresenter.__call__ / binarize / box_score_fast   src/postprocess.py:19-52,186-198
 * Each entry point below names the reference code it replaces.  The Python host side
 * (db_text_minimal_b200/*.py) mirrors those classes and reaches this library through ctypes;
 * INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions (all entries):
 *   - plain pointers + sizes, no torch types; every pointer is DEVICE memory unless named host_*;
 *   - the caller owns all memory (inputs, outputs, workspaces); the library never allocates
 *     device memory and never synchronises the stream;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); entries are CUDA-graph capturable;
 *   - return 0 on success or a negative DBB_E* code; dbb_strerror() names it;
 *   - float tensors are contiguous, 16-byte aligned.
 */
#ifndef DBB200_H_
#define DBB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DBB_VERSION 100

enum {
  DBB_OK = 0,
  DBB_EINVAL = -1,      /* bad shape / argument */
  DBB_EALIGN = -2,      /* pointer not 16-byte aligned */
  DBB_EWORKSPACE = -3,  /* workspace too small */
  DBB_ECUDA = -4,       /* CUDA runtime/driver error (see dbb_last_cuda_error) */
  DBB_EUNSUPPORTED = -5 /* configuration not supported by the sm_100a kernels */
};

int dbb_version(void);
const char* dbb_strerror(int code);
const char* dbb_last_cuda_error(void);
/* number of kernels this library has launched in this process (bench.py: "gpu_launches") */
uint64_t dbb_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * DBLoss  (replaces src/losses.py:18-40 OHEM BCE, :48-66 Dice, :75-82 masked L1, :105-139)
 *
 * preds (N, C, H, W) float32, C = 3 (train: P, T, B) or 2 (eval: P, T)
 * gts   (4, N, H, W) float32: prob_map, supervision_mask, thresh_map, text_area_map  (src/train.py:163-166)
 * reduction: 0 = 'mean' (reference default; degenerate OHEM), 1 = 'none' (true top-k OHEM)
 * losses5: device float[5] = prob, threshold, binary, prob+beta*thr, alpha*binary+prob+beta*thr
 *          (C == 2: binary = 0 and [4] == [3], the single tensor the reference returns)
 * state:   device DbbLossState, consumed by dbb_dbloss_bwd and readable by tests.
 * ------------------------------------------------------------------------------------------ */
typedef struct DbbLossState {
  double sums[12];      /* see db_loss.cu: S_* indices */
  long long n_pos;      /* int(sum(gt*mask))                       src/losses.py:25 */
  long long n_neg;      /* min(int(n_pos*ratio), int(sum((1-gt)*mask)))   :26-28 */
  long long n_above;    /* 'none': #negatives with loss strictly above tau */
  long long n_tie;      /* 'none': #picks taken among entries equal to tau */
  float tau;            /* 'none': k-th largest of (bce * negative); 'mean': mean bce */
  unsigned int tau_bits;
  int tie_ticket;       /* bwd: running ticket for tie picks */
  int reduction;
  float coef[8];        /* precomputed gradient coefficients, see db_loss.cu */
} DbbLossState;

size_t dbb_dbloss_workspace(int64_t n, int c, int64_t h, int64_t w, int reduction);
int dbb_dbloss_fwd(const float* preds, const float* gts, int64_t n, int c, int64_t h, int64_t w,
                   float alpha, float beta, int reduction, float negative_ratio, float eps,
                   float* losses5, DbbLossState* state, void* workspace, size_t workspace_bytes, void* stream);
/* grad_out5: device float[5] upstream gradients of the five returned scalars (autograd);
 * dpreds (N, C, H, W) float32 = d(sum_j grad_out[j]*loss[j]) / d preds. */
int dbb_dbloss_bwd(const float* preds, const float* gts, int64_t n, int c, int64_t h, int64_t w,
                   float alpha, float beta, int reduction, float eps, const float* grad_out5,
                   DbbLossState* state, float* dpreds, void* stream);

/* Ground-truth border map (replaces the distance field of draw_thresh_map, src/db_transforms.py:26-59 + compute_distance
 * :62-78): for every polygon, canvas[image] = fmax(canvas, 1 - min_edges clip(dist(pixel, edge) / D, 0, 1)) over the
 * bounding box (xmin, ymin, xmax, ymax) of the DILATED polygon (the Clipper dilation stays on the host and supplies bbox
 * and D).  canvas: (n_images, h, w) float32, non-negative; pts: float64 xy pairs of all polygons back to back;
 * poly_start[npoly + 1], poly_image[npoly], bbox[npoly][4], dist[npoly]; max_pts <= 64.  Device pointers.  float64,
 * bit-identical to the reference's numpy arithmetic. */
int dbb_thresh_map(float* canvas, int64_t n_images, int64_t h, int64_t w, const double* pts, const int* poly_start,
                   const int* poly_image, const long long* bbox, const double* dist, int npoly, int max_pts, void* stream);

/* Optimizer step (replaces torch.optim.Adam(...).step() of src/train.py:114-117,172; amsgrad off): Adam over flat float32
 * buffers of n elements (n % 4 == 0).  *step (device int64) is incremented, then used for the bias corrections, so the
 * call is CUDA-graph capturable.  grad_scale multiplies g first (1/world for a SUM all-reduce, 1 otherwise). */
int dbb_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                  float weight_decay, float grad_scale, long long* step, void* stream);

/* Per-step pixel metric (replaces src/text_metrics.py:63-82 cal_text_score + :14-23 _fast_hist): accumulates the 2x2
 * confusion matrix hist4[gt*2 + pred] of pred = (P*mask > thresh) vs gt = int(gt*mask) into a device uint64[4].
 * p may be a channel view of the (N,C,H,W) prediction: p_img_stride = C*H*W. */
int dbb_text_score_hist(const float* p, int64_t p_img_stride, const float* gt, const float* mask, int64_t n, int64_t h,
                        int64_t w, float thresh, unsigned long long* hist4, void* stream);

/* ------------------------------------------------------------------------------------------
 * Step function  B = 1/(1+exp(-k(P-T)))   (replaces src/modules/segmentation_head.py:106-108)
 * ------------------------------------------------------------------------------------------ */
int dbb_step_fwd(const float* p, const float* t, float* b, int64_t numel, float k, void* stream);
int dbb_step_bwd(const float* p, const float* t, const float* db, float* dp, float* dt, int64_t numel, float k,
                 void* stream);

/* ------------------------------------------------------------------------------------------
 * Post-processing front  (replaces src/postprocess.py:51-52 binarize, :116-118 findContours
 * candidate extraction, :186-198 box_score_fast, and the score/size filters at :124-130)
 *
 * pred (N, C, H, W) float32, channel 0 is used (postprocess.py:33).  Per image the kernel
 * binarises (P > thresh, strict), labels foreground 8-connected and background 4-connected,
 * builds the containment tree and emits one candidate per foreground component ("outer") and
 * one per enclosed background region ("hole"), i.e. exactly the cv2.RETR_LIST contour set,
 * with the float64 sum / count of P over the contour's fillPoly set.
 * ------------------------------------------------------------------------------------------ */
typedef struct DbbCandidate {
  int kind;             /* 0 = outer border of a foreground component, 1 = hole border */
  int first_y, first_x; /* raster-first pixel of the component/region (discovery order key) */
  int x0, y0, x1, y1;   /* bounding box of the fill set == bounding box of the contour */
  int count;            /* pixels in the fill set */
  double sum;           /* float64 sum of P over the fill set; score = sum / count (cv2.mean) */
  int keep;             /* 1 if not (box_thresh > score)  (postprocess.py:129) */
  int pad_;
} DbbCandidate;

size_t dbb_postprocess_workspace(int64_t n, int64_t h, int64_t w);
/* bitmap: (N, H, W) uint8 out; labels: (N, H, W) int32 out (fg: 1+root index, bg: -(1+root index));
 * cands: (N, max_cands) DbbCandidate out, sorted by (first_y, first_x) descending (cv2 order);
 * n_cands: (N) int32 out (total found, may exceed max_cands). */
int dbb_binarize_ccl_score(const float* pred, int64_t n, int c, int64_t h, int64_t w, float thresh, double box_thresh,
                           uint8_t* bitmap, int32_t* labels, DbbCandidate* cands, int32_t* n_cands, int max_cands,
                           void* workspace, size_t workspace_bytes, void* stream);


/* Border points of the KEPT candidates (replaces the contour handed to get_mini_boxes, src/postprocess.py:119-121: cv2.minAreaRect
 * only sees the contour's convex hull, which equals the hull of the run end pixels of the component).  Must follow
 * dbb_binarize_ccl_score on the same workspace and stream.  points: (N, cap, 2) int32 {candidate slot, (y << 16) | x} device out;
 * n_points: (N) int32 device out (> cap: buffer too small, call again with a larger one). */
int dbb_ccl_border_points(const void* workspace, size_t workspace_bytes, int64_t n, int64_t h, int64_t w, int32_t* points,
                          int32_t* n_points, int cap, void* stream);
/* HOST function: src/postprocess.py:119-147 (get_mini_boxes, sside filter, unclip, second get_mini_boxes, sside filter, rescale)
 * for every kept candidate of a batch, on host copies of the candidate records and border points.  dest_wh: (N, 2) int32
 * (dest_width, dest_height); boxes: (N, max_cands, 4, 2) int16 out, scores: (N, max_cands) float32 out (zero rows for dropped
 * candidates, as the reference); sside_out / mini_out (optional): first get_mini_boxes result per kept candidate;
 * threads: worker threads over images (0 = hardware concurrency). */
int dbb_boxes_from_border_points(const DbbCandidate* cands, const int32_t* n_cands, const int32_t* points, const int32_t* n_points,
                                 int64_t n, int max_cands, int cap_stride, int64_t h, int64_t w, const int32_t* dest_wh,
                                 float unclip_ratio, int min_size, int16_t* boxes, float* scores, float* sside_out,
                                 float* mini_out, int threads);
/* HOST function: get_mini_boxes (src/postprocess.py:158-184 = cv2.minAreaRect + cv2.boxPoints + corner ordering) of an integer
 * contour (npts, 2); box8: 4 x (x, y) float32 out. */
int dbb_mini_box(const int32_t* contour_xy, int npts, float* box8, float* sside);

/* HOST function: src/postprocess.py:54-104 (polygons_from_bitmap: contour, arcLength, approxPolyDP, < 4 points drop, unclip,
 * len(box) > 1 drop, sside filter, rescale) for every kept candidate of a batch.  bits: the packed bitmap (N, H, ceil(W/32))
 * uint32 = the first N*H*ceil(W/32) words of the dbb_binarize_ccl_score workspace, copied to the host.  counts (N, max_cands):
 * points of candidate s's polygon (0 = dropped); points (N, cap, 2) int32: the polygons back to back; scores (N, max_cands)
 * float64; totals (N): points produced per image (> cap: call again with a larger buffer). */
int dbb_polygons_from_bitmap(const uint32_t* bits, const DbbCandidate* cands, const int32_t* n_cands, int64_t n, int max_cands,
                             int64_t h, int64_t w, const int32_t* dest_wh, float unclip_ratio, int min_size, int32_t* counts,
                             int32_t* points, int cap, double* scores, int32_t* totals, int threads);

/* HOST functions (tests / single-contour use): one border of a byte bitmap traced as cv2.findContours(CHAIN_APPROX_SIMPLE) traces
 * it (start = the border's start pixel; is_hole: hole border, start = the foreground pixel left of the hole's raster-first
 * pixel), and cv2.approxPolyDP(contour, eps, closed=True) (+ cv2.arcLength; a negative eps is a ratio of the arc length:
 * src/postprocess.py:71-72 uses 0.005).  Both return the number of points written. */
int dbb_trace_contour(const uint8_t* bitmap, int64_t h, int64_t w, int start_x, int start_y, int is_hole, int32_t* out_xy, int cap);
int dbb_approx_poly_dp(const int32_t* contour_xy, int npts, double eps_or_negative_ratio, int32_t* out_xy, int cap, double* arc_length);

/* HOST function: pyclipper.PyclipperOffset().AddPath(path, JT_ROUND, ET_CLOSEDPOLYGON); Execute(delta) for one closed polygon
 * of any shape (src/postprocess.py:150-156 unclip; src/data_loaders.py:116-122 shrink with delta < 0; src/db_transforms.py:13-21).
 * Restates Clipper 6.4.2's ClipperOffset arithmetic (third-party, not in the reference tree: PARITY UNPINNED, see
 * csrc/clipper_offset.cu).  path_xy: (npts, 2) int64; result polygons are written back to back into out_xy (capacity
 * cap_points points), out_counts[i] = points of polygon i.  Returns the number of polygons, or a negative DBB_E* code. */
int dbb_clipper_offset(const int64_t* path_xy, int npts, double delta, double arc_tolerance, int64_t* out_xy, int cap_points,
                       int32_t* out_counts, int max_paths);
/* the raw offset path before the union step (tests) */
int dbb_clipper_offset_raw(const int64_t* path_xy, int npts, double delta, double arc_tolerance, int64_t* out_xy, int cap_points);

/* Loader -> device boundary (src/data_loaders.py:152-166, src/train.py:163-166): expands a batch shipped as uint8 image (N,3,H,W),
 * uint8 {0,1} prob / supervision-mask / text-area maps (N,H,W) and the float32 threshold map into the float32 tensors the
 * reference's loader produces: img_out (N,3,H,W) = uint8 - mean[c], gts_out (4,N,H,W).  H*W must be a multiple of 16. */
int dbb_unpack_batch(const uint8_t* img_u8, float mean0, float mean1, float mean2, const uint8_t* prob_u8, const uint8_t* mask_u8,
                     const float* thresh_f32, const uint8_t* area_u8, int64_t n, int64_t h, int64_t w, float* img_out,
                     float* gts_out, void* stream);

/* cv2.fillPoly on the device for the ground-truth canvases of a batch (shrink map, supervision mask, text-area map:
 * src/data_loaders.py:107,125,131,135, src/db_transforms.py:22), bit-exact with OpenCV for integer polygons.  verts (total, 2)
 * int32, start (npolys + 1), plane (npolys) = canvas index into maps (planes, H, W) float32, value (npolys): all device. */
int dbb_fill_polygons(const int32_t* verts, const int32_t* start, const int32_t* plane, const float* value, int npolys,
                      float* maps, int64_t h, int64_t w, void* stream);

/* bitmap = pred[:, 0] > thresh on its own (src/postprocess.py:51-52) */
int dbb_binarize(const float* pred, int64_t n, int c, int64_t h, int64_t w, float thresh, uint8_t* bitmap, void* stream);

/* ------------------------------------------------------------------------------------------
 * Whole-network executor: DBTextModel forward / backward  (replaces src/models.py:34-48 with
 * src/modules/resnet.py:231-242, segmentation_body.py:64-87, segmentation_head.py:35-45)
 *
 * The network is a fixed graph of sm_100a kernels (tcgen05 implicit-GEMM convolutions, fused
 * BN/ReLU/residual passes, FPN glue, fused head tail).  Parameters are passed as an array of
 * device pointers in the order returned by dbb_net_param_name(i) (the reference state_dict
 * keys), gradients likewise.
 * ------------------------------------------------------------------------------------------ */
typedef struct DbbNet DbbNet;

int dbb_net_num_params(void);                 /* trainable tensors incl. the unused fc / smooth */
const char* dbb_net_param_name(int i);        /* reference state_dict key */
int dbb_net_param_numel(int i);
int dbb_net_num_buffers(void);                /* BN running_mean / running_var, in pairs */
const char* dbb_net_buffer_name(int i);
int dbb_net_buffer_numel(int i);

/* Plans one (N, H, W, training) configuration.  Returns NULL on error (see dbb_last_cuda_error). */
DbbNet* dbb_net_create(int64_t n, int64_t h, int64_t w, int training);
/* precision: DBB_PRECISION_BF16 (the product path: bf16 activations, tcgen05 convolutions, fp32 accumulate) or
 * DBB_PRECISION_FP32 (parity mode: the same graph and the same elementwise / head kernels on float32 activations with
 * CUDA-core float32 convolutions; slow, used to check the network against the reference at north_star's fp32 tolerance:
 * P, T, B 1e-4, losses / gradients 1e-3). */
#define DBB_PRECISION_BF16 0
#define DBB_PRECISION_FP32 1
DbbNet* dbb_net_create_ex(int64_t n, int64_t h, int64_t w, int training, int precision);
int dbb_net_precision(const DbbNet* net);
void dbb_net_destroy(DbbNet* net);
size_t dbb_net_workspace_bytes(const DbbNet* net);
int64_t dbb_net_out_channels(const DbbNet* net);   /* 3 train, 2 eval */
uint64_t dbb_net_flops_fwd(const DbbNet* net);     /* 2*MACs of the convolutions, forward */

/* x (N,3,H,W) float32; params[i] float32 device pointers; buffers[i] BN running stats (updated in
 * training mode, momentum 0.1, src/modules/basic.py:34); out (N, 3|2, H, W) float32. */
int dbb_net_forward(DbbNet* net, const float* x, const float* const* params, float* const* buffers,
                    float* out, void* workspace, size_t workspace_bytes, void* stream);
/* out: the (N,3,H,W) tensor dbb_net_forward produced; dout (N,3,H,W) float32 -> grads[i] (float32, same shapes as params; OVERWRITTEN, not accumulated;
 * entries for the unused tensors are left untouched).  Must follow dbb_net_forward on the same workspace.
 * segment: -1 = whole backward; 0..dbb_net_num_segments()-1 = one slice (head+FPN first), so the host can
 * overlap the NCCL all-reduce of finished gradient buckets with the rest of the backward. */
int dbb_net_num_segments(void);
int dbb_net_backward(DbbNet* net, const float* out, const float* dout, const float* const* params, float* const* grads,
                     void* workspace, size_t workspace_bytes, int segment, void* stream);
/* Same, with the input image x of the matching forward call.  The bf16 path keeps a staged copy of the image in its
 * workspace and ignores x; the fp32 mode needs it for conv1's weight gradient (src/modules/resnet.py:171). */
int dbb_net_backward_ex(DbbNet* net, const float* x, const float* out, const float* dout, const float* const* params,
                        float* const* grads, void* workspace, size_t workspace_bytes, int segment, void* stream);

/* parity aid (tests): shape of / NCHW float32 copy of a named internal NHWC activation or gradient (bf16, or float32 in the fp32 mode), e.g. "af", "block3.out" */
int dbb_net_debug_shape(DbbNet* net, const char* name, int64_t* shape4);
int dbb_net_debug_read(DbbNet* net, const char* name, const void* workspace, float* out_nchw, void* stream);

/* ------------------------------------------------------------------------------------------
 * Single operators (used by the FPN / DBHead module drop-ins and by the parity tests)
 * ------------------------------------------------------------------------------------------ */
/* Implicit-GEMM convolution on tcgen05: x NHWC bf16 (N,H,W,Cin), w OIHW float32 -> y NHWC bf16 raw output
 * (+bias), fp32 accumulate.  kind: 0 = Conv2d fprop, 1 = Conv2d dgrad (x := dy, y := dx), 2 = ConvTranspose2d(k=2,s=2)
 * fprop, 3 = ConvTranspose2d dgrad. */
int dbb_conv2d(int kind, const void* x_nhwc_bf16, const float* w, const float* bias, void* y_nhwc_bf16,
               int64_t n, int64_t h, int64_t wdt, int cin, int cout, int ksize, int stride, int pad,
               void* workspace, size_t workspace_bytes, void* stream);
size_t dbb_conv2d_workspace(int kind, int64_t n, int64_t h, int64_t wdt, int cin, int cout, int ksize, int stride, int pad);
size_t dbb_conv2d_wgrad_workspace(void);
/* wgrad: dw OIHW float32 (overwritten; deterministic split-K through `workspace`) = sum_px dy[px, co] * x[px @ tap, ci] */
int dbb_conv2d_wgrad(int kind, const void* x_nhwc_bf16, const void* dy_nhwc_bf16, float* dw,
                     int64_t n, int64_t h, int64_t wdt, int cin, int cout, int ksize, int stride, int pad,
                     void* workspace, size_t workspace_bytes, void* stream);


/* Stem convolution (replaces self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False), src/modules/resnet.py:171,
 * 232): img is the NCHW float32 image, y the NHWC bf16 output (n, ceil(h/2), ceil(w/2), 64).  The workspace (1024-byte
 * aligned) holds the space-to-depth staging buffer and the packed weights; backward != 0 sizes it for dbb_conv1_wgrad
 * (dw: (64,3,7,7) float32; the image gets no gradient). */
size_t dbb_conv1_workspace(int64_t n, int64_t h, int64_t w, int backward);
int dbb_conv1_fwd(const float* img, const float* weight, void* y, int64_t n, int64_t h, int64_t w, void* workspace,
                  size_t workspace_bytes, void* stream);
int dbb_conv1_wgrad(const float* img, const void* dy, float* dw, int64_t n, int64_t h, int64_t w, void* workspace,
                    size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Memory-bound operators on NHWC bf16 activations (parity-tested in isolation; also the building blocks of the executor)
 * ------------------------------------------------------------------------------------------ */
size_t dbb_ops_workspace(void);
/* BatchNorm2d [+ residual] [+ ReLU]  (src/modules/basic.py:34-35, resnet.py:74-90).  stats4 out: [scale|shift|mean|invstd] */
int dbb_bn_fwd(const void* z, int64_t pixels, int c, const float* gamma, const float* beta, float* running_mean,
               float* running_var, int training, const void* residual, int relu, void* out, float* stats4,
               void* workspace, size_t workspace_bytes, void* stream);
int dbb_bn_bwd(const void* dout, const void* act, const void* z, int64_t pixels, int c, const float* gamma,
               const float* stats4, void* dz, void* dres, float* dgamma, float* dbeta, void* workspace,
               size_t workspace_bytes, void* stream);
/* MaxPool2d(3, 2, 1)  (src/modules/resnet.py:175) */
int dbb_maxpool_fwd(const void* x, int64_t n, int64_t h, int64_t w, int c, void* y, uint8_t* argmax, void* stream);
int dbb_maxpool_bwd(const void* dy, const uint8_t* argmax, int64_t n, int64_t h, int64_t w, int c, void* dx, void* stream);
/* FPN._upsample_add / _upsample_cat: F.interpolate(mode='nearest')  (src/modules/segmentation_body.py:79-87) */
int dbb_upsample_add_fwd(const void* xs, int64_t hs, int64_t ws, const void* y, int64_t n, int64_t h, int64_t w, int c,
                         void* out, void* stream);
int dbb_upsample_into(const void* xs, int64_t hs, int64_t ws, int64_t n, int64_t h, int64_t w, int c, void* dst,
                      int dst_ctotal, int dst_coff, void* stream);
int dbb_upsample_bwd(const void* d_big, int big_ctotal, int big_coff, int64_t n, int64_t h, int64_t w, int c, void* d_xs,
                     int64_t hs, int64_t ws, int accumulate, void* stream);
/* fused DBHead tail: BN+ReLU -> ConvTranspose2d(64,1,2,2) x2 -> Sigmoid -> step  (src/modules/segmentation_head.py:28-29,39-44,72-76,106-108) */
int dbb_head_tail_fwd(const void* zt, int64_t n, int64_t h2, int64_t w2, const float* gamma, const float* beta,
                      float* running_mean, float* running_var, int training, const float* w2b, const float* w2t,
                      const float* b2b, const float* b2t, float k, int out_c, float* out, float* stats4,
                      void* workspace, size_t workspace_bytes, void* stream);
int dbb_head_tail_bwd(const void* zt, int64_t n, int64_t h2, int64_t w2, const float* gamma, const float* stats4,
                      const float* w2b, const float* w2t, const float* out, const float* dout, float k, void* d_zt,
                      float* dgamma, float* dbeta, float* dw2b, float* dw2t, float* db2b, float* db2t,
                      void* workspace, size_t workspace_bytes, void* stream);

/* per-kernel CUDA-event timing on the launching stream (bench.py roofline); report is JSON text */
void dbb_profile_enable(int on);
size_t dbb_profile_report(char* buf, size_t cap);

/* layout helpers: NCHW float32 <-> NHWC bf16 */
int dbb_nchw_f32_to_nhwc_bf16(const float* x, void* y, int64_t n, int c, int64_t h, int64_t w, void* stream);
int dbb_nhwc_bf16_to_nchw_f32(const void* x, float* y, int64_t n, int c, int64_t h, int64_t w, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DBB200_H_ */
