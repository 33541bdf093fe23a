/* libydst -- B200-native (sm_100a) detect-and-track hot path behind a plain C ABI.
 *
 * The reference (GlassyWing/yolo_deepsort) is pure Python and has no FFI; the boundary it exposes is the
 * duck-typed Python surface used by video_deepsort.py.  Each entry point below names the reference
 * interface it replaces (paths relative to the reference checkout).  The Python mirror of that surface
 * (yolo_deepsort_b200/, dropin/) binds these symbols with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every function returns 0 on success; on failure a non-zero status and a message retrievable with
 *     ydst_last_error() (thread-local).  No C++ exception crosses this boundary.
 *   - pointers named *_dev are CUDA device pointers owned by the caller (e.g. torch tensors' data_ptr());
 *     pointers named *_host are host pointers.  The library only borrows them for the duration of a call.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it.  Functions documented as
 *     "synchronises" wait for that stream before returning (they hand results back to the host).
 *   - handles are not thread-safe; use one tracker handle per video stream (mirrors DeepSort.clone(),
 *     deep_sort/deep_sort.py:41-44).
 */
#ifndef YDST_H
#define YDST_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* ydst_last_error(void);
int ydst_version(void);
/* number of CUDA kernels this library has launched so far in this process (bench.py: gpu_launches) */
long long ydst_launch_count(void);
/* Per-op timing of the layer graphs (detector + ReID): between _begin and _end every op of a forward is bracketed by
 * CUDA events on the launching stream.  _end synchronises and returns, per op in launch order: kind (0 = tcgen05 conv,
 * 1 = first-layer conv, 2 maxpool, 3 upsample, 4 add, 5 copy, 6 yolo decode, 7 avgpool+L2), cfg layer index (-1 for
 * ReID), useful FLOPs (2*M*N*K, convs only), algorithmic bytes (convs only) and milliseconds.  Measurement aid for
 * bench.py's roofline object; no reference counterpart (the reference times with time.time(), img_detect.py:84-95). */
int ydst_profile_begin(void);
int ydst_profile_end(int cap, int* kind_host, int* layer_host, double* flops_host, double* bytes_host, float* ms_host, int* n_host);

/* ------------------------------------------------------------------------------------------------
 * Detector: Darknet graph + YOLO heads + NMS.
 * Replaces yolo3.models.Darknet.{__init__,load_darknet_weights,forward} (yolo3/models/models.py:277-366),
 * YOLOLayer.forward (:185-224) and soft_non_max_suppression (yolo3/utils/model_build.py:52-137).
 * ---------------------------------------------------------------------------------------------- */
enum { YDST_CONV = 0, YDST_MAXPOOL = 1, YDST_UPSAMPLE = 2, YDST_ROUTE = 3, YDST_SHORTCUT = 4, YDST_YOLO = 5 };
enum { YDST_ACT_LINEAR = 0, YDST_ACT_LEAKY = 1, YDST_ACT_MISH = 2, YDST_ACT_RELU = 3 };

typedef struct ydst_layer_desc {
    int type;            /* YDST_CONV ...                                                          */
    int filters;         /* conv: output channels                                                   */
    int size;            /* conv / maxpool: kernel size; upsample: factor                           */
    int stride;          /* conv / maxpool                                                          */
    int batch_normalize; /* conv: 1 -> weights carry [beta, gamma, mean, var], else [bias]          */
    int activation;      /* conv: YDST_ACT_*                                                        */
    int n_src;           /* route: number of sources; shortcut: 1                                   */
    int src[4];          /* route / shortcut: ABSOLUTE layer indices (negative cfg indices resolved) */
    int groups;          /* route: 0 or number of channel groups                                    */
    int group_id;        /* route: selected group                                                   */
    int classes;         /* yolo                                                                    */
    float anchors[6];    /* yolo: the three masked anchors (w0,h0,w1,h1,w2,h2) in pixels            */
} ydst_layer_desc;

typedef struct ydst_detector ydst_detector;

/* weights_host: the float32 payload of a darknet .weights file (everything after the 5 x int32 header),
 * i.e. per conv [bn.bias, bn.weight, running_mean, running_var | conv.bias] then conv.weight (Cout,Cin,k,k)
 * (yolo3/models/models.py:315-366).  height/width: network input size (Darknet.img_size).          */
int ydst_detector_create(const ydst_layer_desc* layers, int n_layers, const float* weights_host, size_t n_weights,
                         int height, int width, int batch, ydst_detector** out);
int ydst_detector_destroy(ydst_detector* d);
/* rows = sum over heads of 3*g*g; fields = 5 + classes */
int ydst_detector_shape(const ydst_detector* d, int* rows, int* fields);
/* Darknet.forward: x_dev is (batch,3,H,W) NCHW float32 (is_half=0) or float16 (is_half=1) in [0,1];
 * pred_dev receives (batch, rows, fields) float32.                                                  */
int ydst_detector_forward_nchw(ydst_detector* d, const void* x_dev, int is_half, float* pred_dev, void* stream);
/* ImageDetector.detect input path (yolo3/detect/img_detect.py:70-82) for a frame already at the network
 * size: frame_dev is HxWx3 uint8 RGB; the /255 happens on the device.  pred_dev may be NULL (the
 * prediction then stays in the handle for ydst_detector_nms).                                        */
int ydst_detector_forward_u8(ydst_detector* d, const uint8_t* frame_dev, float* pred_dev, void* stream);
/* soft_non_max_suppression(pred, conf_thres, iou_thres)[0] on the handle's last prediction (image 0):
 * dets_dev (max_det=300 x 6 float32: x1,y1,x2,y2,conf,cls, score-descending), n_dev (int32 count),
 * both on the device.                                                                               */
int ydst_detector_nms(ydst_detector* d, float conf_thres, float iou_thres, float* dets_dev, int* n_dev, void* stream);
/* Parity aid: the output of cfg layer `layer` from the last forward (what Darknet.forward keeps in layer_outputs,
 * yolo3/models/models.py:296-311), as dense NHWC float16 (N,H,W,C) -- float32 for the linear head convs.  For a
 * convolution whose following shortcut was fused into its epilogue this is the post-add tensor; a yolo layer aliases
 * its head conv. */
int ydst_detector_layer_shape(const ydst_detector* d, int layer, int* n, int* h, int* w, int* c, int* is_f32);
int ydst_detector_layer_output(const ydst_detector* d, int layer, void* dense_dev, void* stream);
/* total useful conv FLOPs of one forward (2*M*N*K over conv layers, logical shapes) */
double ydst_detector_flops(const ydst_detector* d);
/* number of kernel launches one forward issues */
int ydst_detector_launches(const ydst_detector* d);

/* stand-alone NMS on any (rows x fields) float32 prediction (same semantics as above); synchronises and
 * returns the count in *n_host.  Used by the parity tests to feed the oracle's exact predictions.   */
int ydst_nms(const float* pred_dev, int rows, int fields, float conf_thres, float iou_thres, float* dets_dev, int* n_host,
             void* stream);
/* soft_non_max_suppression with its keyword options (yolo3/utils/model_build.py:52-53): is_p1p2 (boxes are corners), merge (the
 * "Merge NMS" block :122-131 exactly as the reference executes it, see csrc/nms.cu), agnostic, classes (n_classes ids on the host, or
 * NULL / 0 for all).  `merge`, `is_p1p2` are what the sliding-window mode of ImageDetector.detect passes
 * (yolo3/detect/img_detect.py:140-143). */
int ydst_nms_ex(const float* pred_dev, int rows, int fields, float conf_thres, float iou_thres, int merge, int is_p1p2, int agnostic,
                const int* classes_host, int n_classes, float* dets_dev, int* n_host, void* stream);
/* Sliding-window mode (yolo3/detect/img_detect.py:97-137): tile t of the batch had its (x,y,w,h) boxes predicted in network pixels;
 * convert in place to corners (xywh2p1p2), scale by (ratio_w[t], ratio_h[t]) (resize_boxes) and shift by (off_x[t], off_y[t]).
 * pred_dev: [tiles][rows][fields] fp32; ratios_host / offsets_host: [tiles][2] = (w, h) / (x, y). */
int ydst_window_boxes(float* pred_dev, int tiles, int rows, int fields, const float* ratios_host, const float* offsets_host, void* stream);

/* One convolution through the same kernels the networks use (tcgen05 implicit GEMM, or the direct first-layer
 * kernel when cin == 3): nn.Conv2d(cin,cout,k,stride,(k-1)//2) [+ BatchNorm2d eval] [+ activation] [+ residual]
 * (yolo3/models/models.py:40-56, deep_sort/deep/model.py:5-37).  x_dev: dense NHWC float16 (float32 when cin == 3);
 * w_host: (cout,cin,k,k) float32; bn_host: NULL or [gamma,beta,mean,var] x cout; bias_host: NULL or (cout);
 * res_dev: NULL or dense NHWC float16 (N,Ho,Wo,cout), res_mode 1 = add after the activation, 2 = before;
 * y_dev: dense NHWC float16, or float32 when y_is_f32.  Synchronises.  Parity-test entry point.             */
int ydst_conv2d(const void* x_dev, int N, int H, int W, int cin, const float* w_host, int cout, int k, int stride,
                const float* bn_host, const float* bias_host, int act, const void* res_dev, int res_mode, void* y_dev,
                int y_is_f32, void* stream);

/* The tiling the planner picks for a stride-1 convolution (k = 1 or 3, cin % 64 == 0) on the halo kernel: N tile, K split,
 * resident CTAs per SM, CTAs launched and the cost model's estimate.  Pure host function (no CUDA call): lets the CPU test
 * tier pin the planner's choices for the BASELINE layer shapes.  No reference counterpart (cuDNN picks its own algorithms). */
int ydst_conv_tiling(int N, int H, int W, int cin, int cout, int k, int* block_n, int* ksplit, int* occupancy, int* ctas,
                     double* model_us);

/* ------------------------------------------------------------------------------------------------
 * ReID extractor: crop + cv2-exact resize + normalise + Net(reid=True).
 * Replaces DeepSort._get_features (deep_sort/deep_sort.py:133-146), Extractor.__call__
 * (deep_sort/deep/feature_extractor.py:34-58) and Net.forward (deep_sort/deep/model.py:81-92).
 * ---------------------------------------------------------------------------------------------- */
typedef struct ydst_reid ydst_reid;
/* weights_host: the checkpoint's 'net_dict' flattened in the order documented in
 * yolo_deepsort_b200/reid.py (stem conv w,b, stem BN g,b,m,v, then per block conv1 w, bn1, conv2 w, bn2,
 * [downsample conv w, bn]).                                                                          */
int ydst_reid_create(const float* weights_host, size_t n_weights, int max_batch, ydst_reid** out);
int ydst_reid_destroy(ydst_reid* r);
/* frame_dev: HxWx3 uint8 RGB; tlwh_dev: (m,4) float32 boxes (x,y,w,h); feat_dev: (m,512) float32 out.
 * Returns status 3 (and a message) if a box yields an empty crop -- the reference raises there.     */
int ydst_reid_extract(ydst_reid* r, const uint8_t* frame_dev, int height, int width, const float* tlwh_dev, int m,
                      float* feat_dev, void* stream);
/* Extractor on pre-cropped, pre-normalised input: x_dev (m,128,64,3) float32 NHWC */
int ydst_reid_forward(ydst_reid* r, const float* x_dev, int m, float* feat_dev, void* stream);
/* the crop/resize/normalise stage alone: out_dev (m,128,64,3) float32 NHWC; synchronises */
int ydst_crop_resize(const uint8_t* frame_dev, int height, int width, const float* tlwh_dev, int m, float* out_dev, void* stream);
double ydst_reid_flops_per_crop(void);

/* ------------------------------------------------------------------------------------------------
 * Association stage kernels (struct-of-arrays track state: mean (n,8), cov (n,8,8) float32).
 * Replace KalmanFilter.{initiate,predict,update,gating_distance} (deep_sort/sort/kalman_filter.py:54-256),
 * NearestNeighborDistanceMetric.distance (deep_sort/sort/nn_matching.py:158-187), gate_cost_matrix and the
 * clamp in min_cost_matching (deep_sort/sort/linear_assignment.py:52,147-203), iou_cost
 * (deep_sort/sort/iou_matching.py:44-91) and scipy.optimize.linear_sum_assignment (linear_assignment.py:56).
 * ---------------------------------------------------------------------------------------------- */
int ydst_kf_initiate(const float* det_tlwh_dev, int n, float* mean_dev, float* cov_dev, void* stream);
int ydst_kf_predict(float* mean_dev, float* cov_dev, int n, void* stream);
int ydst_kf_update(float* mean_dev, float* cov_dev, const float* det_tlwh_dev, int n, void* stream);
int ydst_gate_position(const float* mean_dev, const float* cov_dev, int n, const float* det_tlwh_dev, int m, float* maha_dev,
                       void* stream);
/* gallery_dev (G,512) raw features, seg_host[n+1] row offsets per track; det_feat_dev (m,512); cost_dev (n,m):
 * min cosine distance per track, gated by position Mahalanobis > 5.9915 -> 1e5, clamped > max_dist -> max_dist+1e-5.
 * Thresholds are doubles: the reference compares in float32 but forms max_distance + 1e-5 in double (linear_assignment.py:52). */
int ydst_appearance_cost(const float* gallery_dev, const int* seg_host, int n, const float* det_feat_dev, int m,
                         const float* mean_dev, const float* cov_dev, const float* det_tlwh_dev, double max_dist, float* cost_dev,
                         void* stream);
int ydst_iou_cost(const float* mean_dev, const int* tsu_dev, int n, const float* det_tlwh_dev, int m, double max_dist,
                  float* cost_dev, void* stream);
/* cost_dev (nr,nc) float32 row-major.  Writes min(nr,nc) pairs sorted by row, exactly as scipy returns them;
 * over_max_host[i] = cost[row,col] > max_dist.  Synchronises.                                        */
int ydst_lsap(const float* cost_dev, int nr, int nc, float max_dist, int* rows_host, int* cols_host, int* over_max_host,
              void* stream);

/* ----------------------------------------------------------------------------------------------
 * ActionIdentify: the rule-based action recognition on the (K,6) track rows.
 * Replaces action/action_Identify.py:15-47 (ActionIdentify.update), action/orbit.py:5-26 (Orbit) and the rules of
 * action/actions.py:23-150.  Rule kinds: 0 TakeOff(class_id, delta=(p0,p1)), 1 Landing(class_id, delta), 2 Glide(class_id,
 * delta), 3 FastCrossing(class_id, speed=p0), 4 BreakInto(class_id, timeout=p0).  The orbit cache lives on the device.
 * ---------------------------------------------------------------------------------------------- */
typedef struct ydst_action ydst_action;
int ydst_action_create(int max_age, int max_size, const int* kinds, const int* class_ids, const double* p0, const double* p1,
                       int n_rules, int capacity, ydst_action** out);
int ydst_action_destroy(ydst_action* a);
/* One ActionIdentify.update(detections): rows_host (k,6) int32 [x1,y1,x2,y2,track_id,class_id] (k may be 0: every orbit ages),
 * timestamp = time.time() of the call (orbit.py:26).  triples_host receives *n_host rows (track_id, class_id, rule index) in the
 * reference's order (cache insertion order, rules in list order); capacity k * n_rules.  Synchronises.                  */
int ydst_action_update(ydst_action* a, const int32_t* rows_host, int k, double timestamp, int32_t* triples_host, int* n_host,
                       void* stream);

/* ------------------------------------------------------------------------------------------------
 * Tracker: Tracker + Track + NearestNeighborDistanceMetric + the DeepSort.update output block.
 * Replaces deep_sort/sort/tracker.py:38-176, track.py:63-152, nn_matching.py:139-156,
 * deep_sort/deep_sort.py:60-88.  Track state lives on the device (struct of arrays); the integer
 * lifecycle bookkeeping runs on the host inside this call.
 * ---------------------------------------------------------------------------------------------- */
typedef struct ydst_tracker ydst_tracker;
int ydst_tracker_create(double max_dist, double max_iou_distance, int max_age, int n_init, int nn_budget, int cap_tracks,
                        int cap_dets, ydst_tracker** out);
int ydst_tracker_destroy(ydst_tracker* t);
/* One DeepSort.update step after feature extraction: tlwh_dev (m,4), feat_dev (m,512) float32 on the device,
 * payload_host (m) class ids.  out_host receives K rows [x1,y1,x2,y2,track_id,class_id] int32 (K <= cap_tracks),
 * *k_host = K (0 means the reference would return []).  Synchronises.                                */
int ydst_tracker_update(ydst_tracker* t, const float* tlwh_dev, const float* feat_dev, const int* payload_host, int m,
                        int32_t* out_host, int* k_host, void* stream);
/* same, but the class ids are still on the device as float32 (the detector's dets[:,5]) */
int ydst_tracker_update_dev(ydst_tracker* t, const float* tlwh_dev, const float* feat_dev, const float* cls_dev, int m,
                            int32_t* out_host, int* k_host, void* stream);
/* snapshot of the track table in list order: table_host (n,5) int32 [track_id, hits, age, time_since_update, state],
 * mean_host (n,8) float32; pass NULL to skip either; *n_host = number of tracks.  Synchronises.     */
int ydst_tracker_tracks(ydst_tracker* t, int32_t* table_host, float* mean_host, int cap, int* n_host, void* stream);
/* debug view of the last update: matches (k,2) [track_index, detection_index] of both stages, in order */
int ydst_tracker_last_matches(ydst_tracker* t, int32_t* pairs_host, int cap, int* n_host);

/* ------------------------------------------------------------------------------------------------
 * Fused per-frame pipeline: ImageDetector.detect + the tracker hand-off of VideoDetector.detect
 * (yolo3/detect/video_detect.py:134-149) + DeepSort.update, with one H2D (the frame) and one small D2H
 * (the (K,6) rows) per frame.
 * ---------------------------------------------------------------------------------------------- */
typedef struct ydst_pipeline ydst_pipeline;
int ydst_pipeline_create(ydst_detector* det, ydst_reid* reid, ydst_tracker* trk, float conf_thres, float iou_thres,
                         const int* class_mask_host, int n_mask, ydst_pipeline** out);
int ydst_pipeline_destroy(ydst_pipeline* p);
/* frame_host: HxWx3 uint8 RGB at the network size (pinned memory recommended).  dets_host (optional, 300x6 float32)
 * + *n_dets_host receive the post-NMS, box-rescaled detections; out_host/k_host as in ydst_tracker_update. */
int ydst_pipeline_step(ydst_pipeline* p, const uint8_t* frame_host, int32_t* out_host, int* k_host, float* dets_host,
                       int* n_dets_host, void* stream);
/* Software-pipelined form of the same step: _submit enqueues the DETECTOR half of a frame (copy, Darknet, NMS, hand-off) on an
 * internal stream and returns at once; _collect finishes the OLDEST submitted frame (crops, ReID, DeepSort.update with its host
 * lifecycle) on a second internal stream and returns its rows.  Up to two frames (2B with a micro-batch) may be in flight, so the steady state
 *     submit(f0); for t: { submit(f[t+1]); collect() -> rows of f[t]; }  collect()
 * runs the detector of frame t+1 under the association of frame t -- the look-ahead the reference's reader thread already has
 * (yolo3/detect/video_detect.py:86,112 queues decoded frames ahead of the loop).  Results are identical to _step's, frame by frame.
 * frame: HxWx3 uint8 RGB at the network size, host (pinned recommended) or device; a device frame is copied, so the caller
 * may reuse its buffer.  want_dets: also bring the post-NMS detections back for _collect's dets_host.                              */
int ydst_pipeline_submit(ydst_pipeline* p, const uint8_t* frame, int frame_is_host, int want_dets, void* stream);
int ydst_pipeline_collect(ydst_pipeline* p, int32_t* out_host, int* k_host, float* dets_host, int* n_dets_host);
/* _submit for a frame of ANY size and channel order (the first "next" row of the scope table): the reader thread's BGR->RGB
 * (yolo3/detect/video_detect.py:33-36) and ImageDetector's cv2.resize to the network size (yolo3/detect/img_detect.py:70) run on
 * the device with OpenCV's fixed-point INTER_LINEAR arithmetic; detections are scaled back to the captured size (resize_boxes,
 * yolo3/utils/model_build.py:12-19) and the ReID crops are cut from the captured frame, as DeepSort._get_features does.        */
int ydst_pipeline_submit_frame(ydst_pipeline* p, const uint8_t* frame, int height, int width, int frame_is_host, int is_bgr, int want_dets,
                               void* stream);
/* the ingest kernel alone: cv2.resize(src, (dst_w, dst_h), INTER_LINEAR) on uint8 HxWx3, optionally swapping R and B; synchronises */
int ydst_resize_u8(const uint8_t* src_dev, int src_h, int src_w, uint8_t* dst_dev, int dst_h, int dst_w, int swap_rb, void* stream);
/* The same resize from a sub-rectangle (x0, y0, roi_w, roi_h) of a (src_h, src_w) frame: the window crops of the sliding-window
 * mode, img[y:y+win_h+ov, x:x+win_w+ov] -> cv2.resize (yolo3/detect/img_detect.py:107-111). */
int ydst_resize_u8_roi(const uint8_t* src_dev, int src_h, int src_w, int x0, int y0, int roi_w, int roi_h, uint8_t* dst_dev, int dst_h,
                       int dst_w, int swap_rb, void* stream);
/* Tracker inputs of the frame returned by the LAST _collect / _step: its m rows of tlwh boxes (m x 4), ReID features (m x 512,
 * as handed to Tracker.update, deep_sort/deep_sort.py:55-60) and class ids, copied to the host.  Parity aid: lets a test feed
 * the reference association with exactly what the CUDA association saw.  Valid until the next _submit / _collect. */
int ydst_pipeline_last_inputs(ydst_pipeline* p, float* tlwh_host, float* feat_host, int32_t* cls_host, int cap_rows, int* m_host);
int ydst_pipeline_in_flight(const ydst_pipeline* p);
/* A detector handle created with batch B > 1 makes B the pipeline's MICRO-BATCH: B consecutive frames of the stream share one
 * Darknet forward (and one ReID forward), which amortises the per-layer launch latency that bounds a batch-1 frame; up to 2B
 * frames are then in flight.  _can_submit tells whether the next frame's slot is free.  Per-frame results are unchanged.      */
int ydst_pipeline_can_submit(const ydst_pipeline* p);
/* same with the frame already resident on the device (bench "value" leg) */
int ydst_pipeline_step_dev(ydst_pipeline* p, const uint8_t* frame_dev, int32_t* out_host, int* k_host, float* dets_host,
                           int* n_dets_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* YDST_H */
