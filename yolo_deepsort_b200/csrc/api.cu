// extern "C" surface of libydst (include/ydst.h): argument checking, handle ownership, error strings.
#include <cstring>
#include <mutex>

#include <algorithm>
#include "../../include/ydst.h"
#include "assoc.cuh"
#include "cosine_tc.cuh"
#include "net.cuh"
#include "tracker.cuh"
#include "action.cuh"

namespace ydst {
static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
static long long g_launches = 0;
void count_launch(int n) { g_launches += n; }
}  // namespace ydst

using namespace ydst;

struct ydst_detector { Detector* impl; };
struct ydst_reid { Reid* impl; };
struct ydst_tracker { Tracker* impl; };

struct ydst_pipeline {
    Detector* det; Reid* reid; Tracker* trk;
    float conf, iou;
    int* mask_dev = nullptr; int n_mask = 0;
    int B = 1;                        // detector micro-batch: frames per slot
    // Three slots of B frames form a three-stage pipeline over three streams: Darknet + NMS of batch k+2 (sA), crops + ReID of
    // batch k+1 (sC) and DeepSort.update of batch k (sB, with its host lifecycle).  Frame f lives in slot (f / B) % 3, position f % B.
    static constexpr int kSlots = 3;
    struct Slot {
        uint8_t* orig_dev = nullptr;  // [B][orig_cap] the frames as captured (RGB), when they do not have the network size
        uint8_t* raw_dev = nullptr;   // [orig_cap] staging for a BGR / host frame before the colour swap
        size_t orig_cap = 0;          // bytes per frame in orig_dev
        int fh[8] = {0}, fw[8] = {0}; // captured size of each frame
        bool resized[8] = {false};
        float* feat = nullptr;        // [B * max_det][512]
        cudaEvent_t ev_feat = nullptr;
        bool reid_launched = false;
        uint8_t* frame_dev = nullptr; // [B][H*W*3]
        float *tlwh = nullptr, *confd = nullptr, *cls = nullptr;   // [B][max_det](x4)
        int* h_counts = nullptr;      // pinned [B][8]: [0] candidates, [1] n_dets, [2] overflow, [3] m, [4] crop error flag
        float* h_dets = nullptr;      // pinned [B][300 x 6]
        float* h_cls = nullptr;       // pinned [B][max_det]: class ids of the tracker inputs (float, as the detector emits them)
        cudaEvent_t ev_det = nullptr; // detector half done (counters on the host)
        bool want_dets = false, launched = false, feat_ready = false;
        bool busy = false;            // holds frames that have not all been collected yet
        int n_frames = 0;             // frames stored
        int n_collected = 0;
        int feat_off[8] = {0};        // first feature row of each frame
    } slot[kSlots];
    cudaStream_t sA = nullptr, sB = nullptr, sC = nullptr, sH = nullptr;   // detector | association | crops+ReID | host-frame uploads
    cudaEvent_t ev_in = nullptr, ev_h2d = nullptr;
    bool h2d_pending = false;
    long long submitted = 0, collected = 0;
    int last_slot = -1, last_sub = 0, last_m = 0;   // frame returned by the last collect (ydst_pipeline_last_inputs)
    long long fill = 0, drain = 0;    // slot sequence numbers: slot[fill % kSlots] accepts frames, slot[drain % kSlots] is collected from
    int* h_payload = nullptr;
};

static cudaStream_t S(void* s) { return (cudaStream_t)s; }

extern "C" {

const char* ydst_last_error(void) { return g_err.c_str(); }
int ydst_version(void) { return 100; }
long long ydst_launch_count(void) { return g_launches; }

int ydst_profile_begin(void) {
    YDST_API_BEGIN
    profile_begin();
    YDST_API_END
}
int ydst_profile_end(int cap, int* kind_host, int* layer_host, double* flops_host, double* bytes_host, float* ms_host, int* n_host) {
    YDST_API_BEGIN
    conv_tc_trace_dump();
    YDST_CHECK(n_host, "null argument");
    YDST_CUDA(cudaDeviceSynchronize());
    std::vector<OpSample>& v = profile_samples();
    int n = 0;
    for (OpSample& s : v) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, s.e0, s.e1);
        cudaEventDestroy(s.e0); cudaEventDestroy(s.e1);
        if (n < cap) {
            if (kind_host) kind_host[n] = s.kind;
            if (layer_host) layer_host[n] = s.layer;
            if (flops_host) flops_host[n] = s.flops;
            if (bytes_host) bytes_host[n] = s.bytes;
            if (ms_host) ms_host[n] = ms;
            ++n;
        }
    }
    *n_host = n;
    v.clear();
    YDST_API_END
}

// ---------------- detector ----------------
int ydst_detector_create(const ydst_layer_desc* layers, int n_layers, const float* weights_host, size_t n_weights, int height,
                         int width, int batch, ydst_detector** out) {
    YDST_API_BEGIN
    YDST_CHECK(layers && n_layers > 0 && weights_host && out, "null argument");
    auto* h = new ydst_detector{nullptr};
    try { h->impl = new Detector(layers, n_layers, weights_host, n_weights, height, width, batch); }
    catch (...) { delete h; throw; }
    *out = h;
    YDST_API_END
}
int ydst_detector_destroy(ydst_detector* d) {
    YDST_API_BEGIN
    if (d) { d->impl->nms_.destroy(); delete d->impl; delete d; }
    YDST_API_END
}
int ydst_detector_shape(const ydst_detector* d, int* rows, int* fields) {
    YDST_API_BEGIN
    YDST_CHECK(d, "null handle");
    if (rows) *rows = d->impl->rows;
    if (fields) *fields = d->impl->fields;
    YDST_API_END
}
int ydst_detector_forward_nchw(ydst_detector* d, const void* x_dev, int is_half, float* pred_dev, void* stream) {
    YDST_API_BEGIN
    YDST_CHECK(d && x_dev, "null argument");
    d->impl->forward_nchw(x_dev, is_half, pred_dev, S(stream));
    YDST_API_END
}
int ydst_detector_forward_u8(ydst_detector* d, const uint8_t* frame_dev, float* pred_dev, void* stream) {
    YDST_API_BEGIN
    YDST_CHECK(d && frame_dev, "null argument");
    d->impl->forward_u8(frame_dev, pred_dev, S(stream));
    YDST_API_END
}
int ydst_detector_nms(ydst_detector* d, float conf_thres, float iou_thres, float* dets_dev, int* n_dev, void* stream) {
    YDST_API_BEGIN
    YDST_CHECK(d, "null handle");
    d->impl->nms(conf_thres, iou_thres, dets_dev, n_dev, S(stream));
    YDST_API_END
}
int ydst_detector_layer_shape(const ydst_detector* d, int layer, int* n, int* h, int* w, int* c, int* is_f32) {
    YDST_API_BEGIN
    YDST_CHECK(d, "null handle");
    d->impl->layer_shape(layer, n, h, w, c, is_f32);
    YDST_API_END
}
int ydst_detector_layer_output(const ydst_detector* d, int layer, void* dense_dev, void* stream) {
    YDST_API_BEGIN
    YDST_CHECK(d && dense_dev, "null argument");
    d->impl->layer_output(layer, dense_dev, S(stream));
    YDST_API_END
}
double ydst_detector_flops(const ydst_detector* d) { return d ? d->impl->plan.flops : 0.0; }
int ydst_detector_launches(const ydst_detector* d) { return d ? d->impl->plan.launches + 1 : 0; }

static int nms_standalone(const float* pred_dev, int rows, int fields, float conf_thres, float iou_thres, const NmsOptions* opt,
                          float* dets_dev, int* n_host, void* stream) {
    YDST_API_BEGIN
    YDST_CHECK(pred_dev && dets_dev && n_host && rows >= 0 && fields > 5, "bad argument");
    Nms nms;
    nms.init(4096, 300);
    try {
        nms.run(pred_dev, rows, fields, conf_thres, iou_thres, S(stream), 1, opt);
        int h[4];
        YDST_CUDA(cudaMemcpyAsync(dets_dev, nms.dets, sizeof(float) * 6 * 300, cudaMemcpyDeviceToDevice, S(stream)));
        YDST_CUDA(cudaMemcpyAsync(h, nms.counters, sizeof(int) * 4, cudaMemcpyDeviceToHost, S(stream)));
        YDST_CUDA(cudaStreamSynchronize(S(stream)));
        YDST_CHECK(h[2] == 0, "NMS candidate capacity exceeded (%d candidates > %d)", h[0], nms.cap);
        *n_host = h[1];
    } catch (...) { nms.destroy(); throw; }
    nms.destroy();
    YDST_API_END
}
int ydst_nms(const float* pred_dev, int rows, int fields, float conf_thres, float iou_thres, float* dets_dev, int* n_host, void* stream) {
    return nms_standalone(pred_dev, rows, fields, conf_thres, iou_thres, nullptr, dets_dev, n_host, stream);
}
int ydst_nms_ex(const float* pred_dev, int rows, int fields, float conf_thres, float iou_thres, int merge, int is_p1p2, int agnostic,
                const int* classes_host, int n_classes, float* dets_dev, int* n_host, void* stream) {
    NmsOptions opt{};
    opt.merge = merge != 0; opt.p1p2 = is_p1p2 != 0; opt.agnostic = agnostic != 0;
    if (classes_host && n_classes > 0) {
        opt.use_classes = 1;
        for (int i = 0; i < n_classes; ++i)
            if (classes_host[i] >= 0 && classes_host[i] < 256) opt.class_bits[classes_host[i] >> 6] |= 1ull << (classes_host[i] & 63);
    }
    return nms_standalone(pred_dev, rows, fields, conf_thres, iou_thres, &opt, dets_dev, n_host, stream);
}
int ydst_window_boxes(float* pred_dev, int tiles, int rows, int fields, const float* ratios_host, const float* offsets_host, void* stream) {
    YDST_API_BEGIN
    YDST_CHECK(pred_dev && ratios_host && offsets_host && tiles > 0 && rows > 0 && fields > 5, "bad argument");
    std::vector<float> geo((size_t)tiles * 4);
    for (int t = 0; t < tiles; ++t) {
        geo[t * 4 + 0] = ratios_host[t * 2]; geo[t * 4 + 1] = ratios_host[t * 2 + 1];
        geo[t * 4 + 2] = offsets_host[t * 2]; geo[t * 4 + 3] = offsets_host[t * 2 + 1];
    }
    float* geo_dev = nullptr;
    YDST_CUDA(cudaMalloc(&geo_dev, geo.size() * sizeof(float)));
    try {
        YDST_CUDA(cudaMemcpyAsync(geo_dev, geo.data(), geo.size() * sizeof(float), cudaMemcpyHostToDevice, S(stream)));
        launch_window_boxes(pred_dev, tiles, rows, fields, geo_dev, S(stream));
        count_launch();
        YDST_CUDA(cudaStreamSynchronize(S(stream)));
    } catch (...) { cudaFree(geo_dev); throw; }
    cudaFree(geo_dev);
    YDST_API_END
}

// Stand-alone convolution through the same kernels the networks use (conv_tc / conv_first), for parity tests.
int ydst_conv2d(const void* x_dev, int N, int H, int W, int cin, const float* w_host, int cout, int k, int stride, const float* bn_host,
                const float* bias_host, int act, const void* res_dev, int res_mode, void* y_dev, int y_is_f32, void* stream) {
    YDST_API_BEGIN
    YDST_CHECK(x_dev && w_host && y_dev && N > 0 && H > 0 && W > 0, "bad argument");
    DeviceArena arena;
    const int pad = (k - 1) / 2;
    const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
    const float *g = nullptr, *b = nullptr, *m = nullptr, *v = nullptr;
    if (bn_host) { g = bn_host; b = bn_host + cout; m = bn_host + 2 * cout; v = bn_host + 3 * cout; }
    auto cw = pack_conv(arena, w_host, cout, cin, k, g, b, m, v, bias_host, cin == 3);
    cudaStream_t st = S(stream);
    if (cin == 3) {
        YDST_CHECK(!y_is_f32 && !res_dev, "first-layer path: fp16 output, no residual");
        Act out = make_act(arena, N, Ho, Wo, cout);
        launch_conv_first((const float*)x_dev, N, H, W, cw->w32, cw->w_hilo, cw->scale, cw->bias, cout, stride, act, out, st);
        launch_unpack(out, (__half*)y_dev, st);
    } else {
        Act in = make_act(arena, N, H, W, cin);
        launch_pack((const __half*)x_dev, in, st);
        Act res;
        if (res_dev) { res = make_act(arena, N, Ho, Wo, cout); launch_pack((const __half*)res_dev, res, st); }
        ConvTcLaunch L;
        ConvWorkspace ws = make_conv_workspace(arena);
        if (y_is_f32) {
            float* f32 = (float*)arena.alloc((size_t)N * (Ho + 2) * (Wo + 2) * cw->cout16 * sizeof(float));
            Act geo; geo.N = N; geo.H = Ho; geo.W = Wo; geo.C = cout; geo.ctot = cw->cout16; geo.coff = 0;
            conv_tc_plan(L, in, geo, cw->w16, k, k, stride, cw->scale, cw->bias, act, 0, nullptr, f32, cout, &ws);
            conv_tc_run(L, st);
            launch_unpack_f32(f32, cw->cout16, N, Ho, Wo, cout, (float*)y_dev, st);
        } else {
            Act out = make_act(arena, N, Ho, Wo, cout);
            conv_tc_plan(L, in, out, cw->w16, k, k, stride, cw->scale, cw->bias, act, res_dev ? res_mode : 0, res_dev ? &res : nullptr, nullptr, cout, &ws);
            conv_tc_run(L, st);
            launch_unpack(out, (__half*)y_dev, st);
        }
    }
    YDST_CUDA(cudaStreamSynchronize(st));
    YDST_API_END
}

int ydst_conv_tiling(int N, int H, int W, int cin, int cout, int k, int* block_n, int* ksplit, int* occupancy, int* ctas, double* model_us) {
    YDST_API_BEGIN
    YDST_CHECK(N > 0 && H > 0 && W > 0 && cin > 0 && cin % 64 == 0 && cout > 0 && (k == 1 || k == 3), "bad argument");
    const long long P = (long long)N * (H + 2) * (W + 2);
    const int m_tiles = (int)((P + 127) / 128);
    const int halo = k == 3 ? W + 3 : 0;
    const int cout16 = (cout + 15) & ~15;
    const ConvTiling t = conv_tc_choose_tiling(m_tiles, cout16, k * k, cin / 64, halo, (size_t)48 << 20, 8192);
    if (block_n) *block_n = t.bn;
    if (ksplit) *ksplit = t.ksplit;
    if (occupancy) *occupancy = t.occupancy;
    if (ctas) *ctas = t.persistent ? std::min(148, ((m_tiles + t.mpair - 1) / t.mpair) * ((cout16 + t.bn - 1) / t.bn))
                            : ((m_tiles + t.mpair - 1) / t.mpair) * ((cout16 + t.bn - 1) / t.bn) * t.ksplit;
    if (model_us) *model_us = t.model_us;
    YDST_API_END
}

// ---------------- ReID ----------------
int ydst_reid_create(const float* weights_host, size_t n_weights, int max_batch, ydst_reid** out) {
    YDST_API_BEGIN
    YDST_CHECK(weights_host && out, "null argument");
    auto* h = new ydst_reid{nullptr};
    try { h->impl = new Reid(weights_host, n_weights, max_batch); }
    catch (...) { delete h; throw; }
    *out = h;
    YDST_API_END
}
int ydst_reid_destroy(ydst_reid* r) {
    YDST_API_BEGIN
    if (r) { delete r->impl; delete r; }
    YDST_API_END
}
int ydst_reid_extract(ydst_reid* r, const uint8_t* frame_dev, int height, int width, const float* tlwh_dev, int m, float* feat_dev,
                      void* stream) {
    YDST_API_BEGIN
    YDST_CHECK(r && frame_dev && (m == 0 || (tlwh_dev && feat_dev)), "null argument");
    YDST_CUDA(cudaMemsetAsync(r->impl->err_flag, 0, 8 * sizeof(int), S(stream)));
    r->impl->extract(frame_dev, height, width, tlwh_dev, m, feat_dev, S(stream));
    int flag = 0;
    YDST_CUDA(cudaMemcpyAsync(&flag, r->impl->err_flag, sizeof(int), cudaMemcpyDeviceToHost, S(stream)));
    YDST_CUDA(cudaStreamSynchronize(S(stream)));
    if (flag) { set_error("empty crop: a box has no pixels inside the frame (cv2.resize raises in the reference)"); return 3; }
    YDST_API_END
}
int ydst_reid_forward(ydst_reid* r, const float* x_dev, int m, float* feat_dev, void* stream) {
    YDST_API_BEGIN
    YDST_CHECK(r && (m == 0 || (x_dev && feat_dev)), "null argument");
    r->impl->forward(x_dev, m, feat_dev, S(stream));
    YDST_API_END
}
int ydst_crop_resize(const uint8_t* frame_dev, int height, int width, const float* tlwh_dev, int m, float* out_dev, void* stream) {
    YDST_API_BEGIN
    YDST_CHECK(frame_dev && (m == 0 || (tlwh_dev && out_dev)), "null argument");
    int* flag_dev = nullptr;
    YDST_CUDA(cudaMalloc(&flag_dev, sizeof(int)));
    int flag = 0;
    cudaMemsetAsync(flag_dev, 0, sizeof(int), S(stream));
    try {
        launch_crop_resize(frame_dev, height, width, tlwh_dev, m, out_dev, flag_dev, S(stream));
        YDST_CUDA(cudaMemcpyAsync(&flag, flag_dev, sizeof(int), cudaMemcpyDeviceToHost, S(stream)));
        YDST_CUDA(cudaStreamSynchronize(S(stream)));
    } catch (...) { cudaFree(flag_dev); throw; }
    cudaFree(flag_dev);
    if (flag) { set_error("empty crop: a box has no pixels inside the frame (cv2.resize raises in the reference)"); return 3; }
    YDST_API_END
}
double ydst_reid_flops_per_crop(void) {
    double f = 2.0 * 128 * 64 * 64 * 27;
    const int st[4][3] = {{64, 64, 0}, {64, 128, 1}, {128, 256, 1}, {256, 512, 1}};
    int h = 64, w = 32;
    for (int s = 0; s < 4; ++s) {
        if (s > 0) { h /= 2; w /= 2; }
        const double px = (double)h * w;
        f += 2.0 * px * st[s][1] * 9 * st[s][0];                // block 0 conv1
        f += 3 * 2.0 * px * st[s][1] * 9 * st[s][1];            // the other three 3x3 convs
        if (st[s][2]) f += 2.0 * px * st[s][1] * st[s][0];      // 1x1 downsample
    }
    return f;
}

// ---------------- association stage kernels ----------------
int ydst_kf_initiate(const float* det_tlwh_dev, int n, float* mean_dev, float* cov_dev, void* stream) {
    YDST_API_BEGIN
    launch_kf_initiate(det_tlwh_dev, nullptr, mean_dev, cov_dev, nullptr, n, S(stream));
    YDST_API_END
}
int ydst_kf_predict(float* mean_dev, float* cov_dev, int n, void* stream) {
    YDST_API_BEGIN
    launch_kf_predict(mean_dev, cov_dev, nullptr, n, S(stream));
    YDST_API_END
}
int ydst_kf_update(float* mean_dev, float* cov_dev, const float* det_tlwh_dev, int n, void* stream) {
    YDST_API_BEGIN
    launch_kf_update(mean_dev, cov_dev, nullptr, det_tlwh_dev, nullptr, n, S(stream));
    YDST_API_END
}
int ydst_gate_position(const float* mean_dev, const float* cov_dev, int n, const float* det_tlwh_dev, int m, float* maha_dev, void* stream) {
    YDST_API_BEGIN
    launch_gate_position(mean_dev, cov_dev, nullptr, n, det_tlwh_dev, m, maha_dev, S(stream));
    YDST_API_END
}
int ydst_appearance_cost(const float* gallery_dev, const int* seg_host, int n, const float* det_feat_dev, int m, const float* mean_dev,
                         const float* cov_dev, const float* det_tlwh_dev, double max_dist, float* cost_dev, void* stream) {
    YDST_API_BEGIN
    YDST_CHECK(seg_host && n >= 0 && m >= 0, "bad argument");
    if (n == 0 || m == 0) return 0;
    const int G = seg_host[n];
    std::vector<int> row_ptr(G);
    for (int g = 0; g < G; ++g) row_ptr[g] = g;
    float *gal_n = nullptr, *det_n = nullptr;
    int *d_rp = nullptr, *d_seg = nullptr;
    auto cleanup = [&]() { cudaFree(gal_n); cudaFree(det_n); cudaFree(d_rp); cudaFree(d_seg); };
    try {
        YDST_CUDA(cudaMalloc(&gal_n, (size_t)std::max(G, 1) * kFeat * sizeof(float)));
        YDST_CUDA(cudaMalloc(&det_n, (size_t)m * kFeat * sizeof(float)));
        YDST_CUDA(cudaMalloc(&d_rp, (size_t)std::max(G, 1) * sizeof(int)));
        YDST_CUDA(cudaMalloc(&d_seg, (size_t)(n + 1) * sizeof(int)));
        YDST_CUDA(cudaMemcpyAsync(d_rp, row_ptr.data(), G * sizeof(int), cudaMemcpyHostToDevice, S(stream)));
        YDST_CUDA(cudaMemcpyAsync(d_seg, seg_host, (n + 1) * sizeof(int), cudaMemcpyHostToDevice, S(stream)));
        launch_normalize_rows(gallery_dev, gal_n, G, S(stream));
        launch_normalize_rows(det_feat_dev, det_n, m, S(stream));
        CosineTc cos;
        cos.run(gal_n, d_rp, d_seg, G, n, det_n, m, mean_dev, cov_dev, nullptr, det_tlwh_dev, max_dist, cost_dev, S(stream));
        YDST_CUDA(cudaStreamSynchronize(S(stream)));
    } catch (...) { cleanup(); throw; }
    cleanup();
    YDST_API_END
}
int ydst_iou_cost(const float* mean_dev, const int* tsu_dev, int n, const float* det_tlwh_dev, int m, double max_dist, float* cost_dev,
                  void* stream) {
    YDST_API_BEGIN
    launch_iou_cost(mean_dev, nullptr, tsu_dev, n, det_tlwh_dev, nullptr, m, max_dist, cost_dev, S(stream));
    YDST_API_END
}
int ydst_lsap(const float* cost_dev, int nr, int nc, float max_dist, int* rows_host, int* cols_host, int* over_max_host, void* stream) {
    YDST_API_BEGIN
    YDST_CHECK(nr >= 0 && nc >= 0, "bad shape");
    if (nr == 0 || nc == 0) return 0;
    const bool tr = nc < nr;
    const int R = tr ? nc : nr, C = tr ? nr : nc;
    float* ct = nullptr; int* res = nullptr; void* work = nullptr;
    auto cleanup = [&]() { cudaFree(ct); cudaFree(res); cudaFree(work); };
    try {
        YDST_CUDA(cudaMalloc(&res, 2 * (size_t)R * sizeof(int)));
        YDST_CUDA(cudaMalloc(&work, lsap_work_bytes(R, C)));
        const float* c = cost_dev;
        if (tr) {
            YDST_CUDA(cudaMalloc(&ct, (size_t)nr * nc * sizeof(float)));
            launch_transpose(cost_dev, ct, nr, nc, S(stream));
            c = ct;
        }
        launch_lsap(c, R, C, max_dist, res, res + R, work, S(stream));
        std::vector<int> h(2 * (size_t)R);
        YDST_CUDA(cudaMemcpyAsync(h.data(), res, 2 * (size_t)R * sizeof(int), cudaMemcpyDeviceToHost, S(stream)));
        YDST_CUDA(cudaStreamSynchronize(S(stream)));
        if (!tr) {
            for (int r = 0; r < R; ++r) { rows_host[r] = r; cols_host[r] = h[r]; if (over_max_host) over_max_host[r] = h[R + r]; }
        } else {                                   // solver rows are original columns; emit sorted by original row
            std::vector<int> col_of_row(nr, -1), over(nr, 0);
            for (int r = 0; r < R; ++r) { col_of_row[h[r]] = r; over[h[r]] = h[R + r]; }
            int k = 0;
            for (int orow = 0; orow < nr; ++orow)
                if (col_of_row[orow] >= 0) { rows_host[k] = orow; cols_host[k] = col_of_row[orow]; if (over_max_host) over_max_host[k] = over[orow]; ++k; }
        }
        for (int r = 0; r < R; ++r) YDST_CHECK(h[r] >= 0, "cost matrix is infeasible");
    } catch (...) { cleanup(); throw; }
    cleanup();
    YDST_API_END
}

// ---------------- tracker ----------------
// ---- ActionIdentify (csrc/action.cu) ----
struct ydst_action { ydst::ActionIdentifyDev* impl; };
int ydst_action_create(int max_age, int max_size, const int* kinds, const int* class_ids, const double* p0, const double* p1, int n_rules,
                       int capacity, ydst_action** out) {
    YDST_API_BEGIN
    YDST_CHECK(out && (n_rules == 0 || (kinds && class_ids && p0 && p1)), "null argument");
    auto* h = new ydst_action{nullptr};
    try { h->impl = ydst::action_create(max_age, max_size, kinds, class_ids, p0, p1, n_rules, capacity); }
    catch (...) { delete h; throw; }
    *out = h;
    YDST_API_END
}
int ydst_action_destroy(ydst_action* a) {
    YDST_API_BEGIN
    if (a) { ydst::action_destroy(a->impl); delete a; }
    YDST_API_END
}
int ydst_action_update(ydst_action* a, const int32_t* rows_host, int k, double timestamp, int32_t* triples_host, int* n_host, void* stream) {
    YDST_API_BEGIN
    YDST_CHECK(a && a->impl && triples_host && n_host && (k == 0 || rows_host), "null argument");
    *n_host = ydst::action_update(a->impl, rows_host, k, timestamp, triples_host, S(stream));
    YDST_API_END
}

int ydst_tracker_create(double max_dist, double max_iou_distance, int max_age, int n_init, int nn_budget, int cap_tracks, int cap_dets,
                        ydst_tracker** out) {
    YDST_API_BEGIN
    YDST_CHECK(out, "null argument");
    auto* h = new ydst_tracker{nullptr};
    try { h->impl = new Tracker(max_dist, max_iou_distance, max_age, n_init, nn_budget, cap_tracks, cap_dets); }
    catch (...) { delete h; throw; }
    *out = h;
    YDST_API_END
}
int ydst_tracker_destroy(ydst_tracker* t) {
    YDST_API_BEGIN
    if (t) { delete t->impl; delete t; }
    YDST_API_END
}
int ydst_tracker_update(ydst_tracker* t, const float* tlwh_dev, const float* feat_dev, const int* payload_host, int m, int32_t* out_host,
                        int* k_host, void* stream) {
    YDST_API_BEGIN
    YDST_CHECK(t && out_host && k_host && (m == 0 || (tlwh_dev && feat_dev && payload_host)), "null argument");
    t->impl->update(tlwh_dev, feat_dev, payload_host, nullptr, m, out_host, k_host, S(stream));
    YDST_API_END
}
int ydst_tracker_update_dev(ydst_tracker* t, const float* tlwh_dev, const float* feat_dev, const float* cls_dev, int m, int32_t* out_host,
                            int* k_host, void* stream) {
    YDST_API_BEGIN
    YDST_CHECK(t && out_host && k_host && (m == 0 || (tlwh_dev && feat_dev && cls_dev)), "null argument");
    t->impl->update(tlwh_dev, feat_dev, nullptr, cls_dev, m, out_host, k_host, S(stream));
    YDST_API_END
}
int ydst_tracker_tracks(ydst_tracker* t, int32_t* table_host, float* mean_host, int cap, int* n_host, void* stream) {
    YDST_API_BEGIN
    YDST_CHECK(t && n_host, "null argument");
    t->impl->snapshot(table_host, mean_host, cap, n_host, S(stream));
    YDST_API_END
}
int ydst_tracker_last_matches(ydst_tracker* t, int32_t* pairs_host, int cap, int* n_host) {
    YDST_API_BEGIN
    YDST_CHECK(t && n_host, "null argument");
    const auto& m = t->impl->last_matches;
    *n_host = (int)m.size();
    YDST_CHECK((int)m.size() <= cap, "buffer too small");
    for (size_t i = 0; i < m.size(); ++i) { pairs_host[2 * i] = m[i].first; pairs_host[2 * i + 1] = m[i].second; }
    YDST_API_END
}

// ---------------- fused per-frame pipeline ----------------
int ydst_pipeline_create(ydst_detector* det, ydst_reid* reid, ydst_tracker* trk, float conf_thres, float iou_thres,
                         const int* class_mask_host, int n_mask, ydst_pipeline** out) {
    YDST_API_BEGIN
    YDST_CHECK(det && reid && trk && out, "null argument");
    YDST_CHECK(det->impl->batch >= 1 && det->impl->batch <= 8, "the per-frame pipeline takes a detector with batch 1..8 (its micro-batch)");
    auto* p = new ydst_pipeline();
    p->det = det->impl; p->reid = reid->impl; p->trk = trk->impl; p->conf = conf_thres; p->iou = iou_thres; p->n_mask = n_mask;
    p->B = det->impl->batch;
    const int md = p->det->nms_.max_det, B = p->B;
    for (auto& sl : p->slot) {
        YDST_CUDA(cudaMalloc(&sl.frame_dev, (size_t)B * p->det->H * p->det->W * 3));
        YDST_CUDA(cudaMalloc(&sl.tlwh, sizeof(float) * 4 * md * B));
        YDST_CUDA(cudaMalloc(&sl.confd, sizeof(float) * md * B));
        YDST_CUDA(cudaMalloc(&sl.cls, sizeof(float) * md * B));
        YDST_CUDA(cudaMallocHost(&sl.h_counts, sizeof(int) * 8 * B));
        YDST_CUDA(cudaMallocHost(&sl.h_dets, sizeof(float) * 6 * md * B));
        YDST_CUDA(cudaMallocHost(&sl.h_cls, sizeof(float) * md * B));
        YDST_CUDA(cudaEventCreateWithFlags(&sl.ev_det, cudaEventDisableTiming));
        YDST_CUDA(cudaEventCreateWithFlags(&sl.ev_feat, cudaEventDisableTiming));
        YDST_CUDA(cudaMalloc(&sl.feat, sizeof(float) * 512 * md * B));
    }
    YDST_CUDA(cudaMalloc(&p->mask_dev, sizeof(int) * (n_mask > 0 ? n_mask : 1)));
    if (n_mask > 0) YDST_CUDA(cudaMemcpy(p->mask_dev, class_mask_host, sizeof(int) * n_mask, cudaMemcpyHostToDevice));
    // oldest work first: the association of frame t (many small kernels, host in the loop) must not queue behind the CTAs of the
    // ReID forward of t+1 and the detector forward of t+2, which each fill the GPU on their own
    int prio_least = 0, prio_greatest = 0;
    YDST_CUDA(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
    const char* pe = getenv("YDST_STREAM_PRIO");
    const int prio_mode = pe ? atoi(pe) : 2;                           // 0: off, 1: association > ReID > detector, 2: association only (measured best)
    const bool use_prio = prio_mode != 0;
    const int prio_mid = (prio_mode == 1 && prio_greatest < prio_least) ? prio_greatest + 1 : prio_least;
    YDST_CUDA(cudaStreamCreateWithPriority(&p->sA, cudaStreamNonBlocking, prio_least));
    YDST_CUDA(cudaStreamCreateWithPriority(&p->sB, cudaStreamNonBlocking, use_prio ? prio_greatest : prio_least));
    YDST_CUDA(cudaStreamCreateWithPriority(&p->sC, cudaStreamNonBlocking, use_prio ? prio_mid : prio_least));
    YDST_CUDA(cudaStreamCreateWithPriority(&p->sH, cudaStreamNonBlocking, prio_least));
    YDST_CUDA(cudaEventCreateWithFlags(&p->ev_in, cudaEventDisableTiming));
    YDST_CUDA(cudaEventCreateWithFlags(&p->ev_h2d, cudaEventDisableTiming));
    p->h_payload = new int[md];
    *out = p;
    YDST_API_END
}
int ydst_pipeline_destroy(ydst_pipeline* p) {
    YDST_API_BEGIN
    if (p) {
        if (p->sA) cudaStreamSynchronize(p->sA);
        if (p->sB) cudaStreamSynchronize(p->sB);
        if (p->sC) cudaStreamSynchronize(p->sC);
        if (p->sH) cudaStreamSynchronize(p->sH);
        for (auto& sl : p->slot) {
            cudaFree(sl.frame_dev); cudaFree(sl.tlwh); cudaFree(sl.confd); cudaFree(sl.cls); cudaFree(sl.feat);
            cudaFree(sl.orig_dev); cudaFree(sl.raw_dev);
            cudaFreeHost(sl.h_counts); cudaFreeHost(sl.h_dets); cudaFreeHost(sl.h_cls);
            if (sl.ev_det) cudaEventDestroy(sl.ev_det);
            if (sl.ev_feat) cudaEventDestroy(sl.ev_feat);
        }
        cudaFree(p->mask_dev);
        if (p->sA) cudaStreamDestroy(p->sA);
        if (p->sB) cudaStreamDestroy(p->sB);
        if (p->sC) cudaStreamDestroy(p->sC);
        if (p->sH) cudaStreamDestroy(p->sH);
        if (p->ev_in) cudaEventDestroy(p->ev_in);
        if (p->ev_h2d) cudaEventDestroy(p->ev_h2d);
        delete[] p->h_payload;
        delete p;
    }
    YDST_API_END
}

// a frame can be stored if the slot being filled is free, or still filling (not yet handed to the detector)
static bool pipeline_can_submit(const ydst_pipeline* p) {
    const ydst_pipeline::Slot& sl = p->slot[p->fill % ydst_pipeline::kSlots];
    return !sl.busy || !sl.launched;
}

// detector half of one slot, enqueued on sA: Darknet forward over its frames, then per frame NMS, tracker hand-off, small D2H.
// A partially filled slot (end of stream / synchronous step) runs the same batch plan; the unused images are ignored.
static void pipeline_launch_detector(ydst_pipeline* p, ydst_pipeline::Slot& sl) {
    Detector& det = *p->det;
    const int md = det.nms_.max_det;
    if (p->h2d_pending) {                                              // host frames of this slot were uploaded on their own stream
        YDST_CUDA(cudaEventRecord(p->ev_h2d, p->sH));
        YDST_CUDA(cudaStreamWaitEvent(p->sA, p->ev_h2d, 0));
        p->h2d_pending = false;
    }
    det.forward_u8(sl.frame_dev, nullptr, p->sA);
    // one set of NMS launches for the whole micro-batch (image = blockIdx.y), then one D2H per result array
    det.nms_.run(det.pred, det.rows, det.fields, p->conf, p->iou, p->sA, sl.n_frames);
    NmsRatios ratios{};
    for (int b = 0; b < sl.n_frames; ++b) {
        // resize_boxes: x *= w / W, y *= h / H with python-float ratios applied to fp32 boxes (yolo3/utils/model_build.py:12-19)
        ratios.rw[b] = (float)((double)sl.fw[b] / (double)det.W);
        ratios.rh[b] = (float)((double)sl.fh[b] / (double)det.H);
    }
    det.nms_.to_tracker_inputs_batch(ratios, sl.n_frames, p->mask_dev, p->n_mask, sl.tlwh, sl.confd, sl.cls, p->sA);
    YDST_CUDA(cudaMemcpyAsync(sl.h_counts, det.nms_.counters, sizeof(int) * 8 * sl.n_frames, cudaMemcpyDeviceToHost, p->sA));
    if (sl.want_dets)
        YDST_CUDA(cudaMemcpyAsync(sl.h_dets, det.nms_.dets, sizeof(float) * 6 * md * sl.n_frames, cudaMemcpyDeviceToHost, p->sA));
    // class ids of the tracker inputs ride along with the counters (saves the tracker a kernel + a synchronisation)
    YDST_CUDA(cudaMemcpyAsync(sl.h_cls, sl.cls, sizeof(float) * md * sl.n_frames, cudaMemcpyDeviceToHost, p->sA));
    YDST_CUDA(cudaEventRecord(sl.ev_det, p->sA));
    sl.launched = true;
    ++p->fill;                                                         // the next frame starts a new slot
}

static void pipeline_submit(ydst_pipeline* p, const uint8_t* frame, bool frame_is_host, bool want_dets, cudaStream_t caller, int fh = 0,
                            int fw = 0, bool is_bgr = false) {
    YDST_CHECK(pipeline_can_submit(p), "the pipeline is full (%lld frames in flight): collect one first", p->submitted - p->collected);
    ydst_pipeline::Slot& sl = p->slot[p->fill % ydst_pipeline::kSlots];
    if (!sl.busy) {
        sl.busy = true; sl.n_frames = 0; sl.n_collected = 0;
        sl.launched = false; sl.feat_ready = false; sl.reid_launched = false; sl.want_dets = false;
    }
    const int sub = sl.n_frames;
    Detector& det = *p->det;
    const size_t bytes = (size_t)det.H * det.W * 3;
    if (fh <= 0 || fw <= 0) { fh = det.H; fw = det.W; }
    const bool plain = fh == det.H && fw == det.W && !is_bgr;            // already what the network eats
    sl.fh[sub] = fh; sl.fw[sub] = fw; sl.resized[sub] = !plain;
    const cudaMemcpyKind kind = frame_is_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    // A DEVICE frame belongs to the caller and may be freed or overwritten as soon as this call returns (a torch tensor going out
    // of scope hands its block back to the caching allocator, which reuses it on the caller's stream): its copy into the slot is
    // therefore enqueued on the CALLER's stream, ordered with whatever the caller does next, and sA picks up after it.  The slot
    // being filled has no reader left (its last collect waited for the ReID half), so writing it from another stream is safe.
    // HOST frames are read by the copy engine until collected (the Python side keeps them alive); at the network size they go
    // through the copy stream, under the previous slot's detector forward.
    const cudaStream_t cs = frame_is_host ? p->sA : caller;
    if (plain && frame_is_host) {
        YDST_CUDA(cudaMemcpyAsync(sl.frame_dev + sub * bytes, frame, bytes, kind, p->sH));
        p->h2d_pending = true;
    } else if (plain) {
        YDST_CUDA(cudaMemcpyAsync(sl.frame_dev + sub * bytes, frame, bytes, kind, cs));
    } else {
        // ingest on the device: (BGR ->) RGB copy of the captured frame for the ReID crops, cv2-exact resize to the network size
        const size_t fbytes = (size_t)fh * fw * 3;
        if (fbytes > sl.orig_cap) {
            YDST_CHECK(sub == 0, "frames of one micro-batch must not grow in size");
            YDST_CUDA(cudaStreamSynchronize(p->sA));
            YDST_CUDA(cudaStreamSynchronize(p->sC));
            cudaFree(sl.orig_dev); cudaFree(sl.raw_dev);
            sl.orig_cap = fbytes;
            YDST_CUDA(cudaMalloc(&sl.orig_dev, sl.orig_cap * p->B));
            YDST_CUDA(cudaMalloc(&sl.raw_dev, sl.orig_cap));
        }
        uint8_t* orig = sl.orig_dev + (size_t)sub * sl.orig_cap;
        YDST_CUDA(cudaMemcpyAsync(is_bgr ? sl.raw_dev : orig, frame, fbytes, kind, cs));
    }
    if (!frame_is_host) {
        YDST_CUDA(cudaEventRecord(p->ev_in, caller));
        YDST_CUDA(cudaStreamWaitEvent(p->sA, p->ev_in, 0));
    }
    if (!plain) {
        const size_t fbytes = (size_t)fh * fw * 3; (void)fbytes;
        uint8_t* orig = sl.orig_dev + (size_t)sub * sl.orig_cap;
        if (is_bgr) {
            launch_resize_u8(sl.raw_dev, fh, fw, orig, fh, fw, 1, p->sA);        // same size: a channel-swapping copy
            count_launch();
        }
        launch_resize_u8(orig, fh, fw, sl.frame_dev + sub * bytes, det.H, det.W, 0, p->sA);
        count_launch();
    }
    sl.want_dets = sl.want_dets || want_dets;
    sl.n_frames = sub + 1;
    ++p->submitted;
    if (sl.n_frames == p->B) pipeline_launch_detector(p, sl);
}

// ReID half of one slot on sC: crops of all its frames through one forward.  Needs the detector's counters on the host:
// blocks for them, or (block == false) only proceeds if the detector half has already finished.
static void pipeline_launch_reid(ydst_pipeline* p, ydst_pipeline::Slot& sl, bool block) {
    if (!sl.launched || sl.reid_launched) return;
    if (block) YDST_CUDA(cudaEventSynchronize(sl.ev_det));
    else if (cudaEventQuery(sl.ev_det) != cudaSuccess) return;
    Detector& det = *p->det;
    const int md = det.nms_.max_det;
    const size_t bytes = (size_t)det.H * det.W * 3;
    const uint8_t* frames[8]; const float* boxes[8]; int ms[8], hs[8], ws[8];
    int off = 0;
    for (int b = 0; b < sl.n_frames; ++b) {
        YDST_CHECK(sl.h_counts[b * 8 + 2] == 0, "NMS candidate capacity exceeded (%d candidates)", sl.h_counts[b * 8]);
        // crops come from the frame as captured (deep_sort/deep_sort.py:133-141 slices ori_img), boxes are already in its pixels
        frames[b] = sl.resized[b] ? sl.orig_dev + (size_t)b * sl.orig_cap : sl.frame_dev + b * bytes;
        hs[b] = sl.fh[b]; ws[b] = sl.fw[b];
        boxes[b] = sl.tlwh + (size_t)b * md * 4;
        ms[b] = sl.h_counts[b * 8 + 1] > 0 ? sl.h_counts[b * 8 + 3] : 0;
        sl.feat_off[b] = off; off += ms[b];
    }
    YDST_CUDA(cudaMemsetAsync(p->reid->err_flag, 0, 8 * sizeof(int), p->sC));
    p->reid->extract_multi(frames, hs, ws, boxes, ms, sl.n_frames, sl.feat, p->sC);
    // one empty-crop flag per frame, into that frame's counter row ([4])
    YDST_CUDA(cudaMemcpy2DAsync(sl.h_counts + 4, 8 * sizeof(int), p->reid->err_flag, sizeof(int), sizeof(int), sl.n_frames,
                                cudaMemcpyDeviceToHost, p->sC));
    YDST_CUDA(cudaEventRecord(sl.ev_feat, p->sC));
    sl.reid_launched = true;
}

// association half of the oldest submitted frame, on sB: DeepSort.update with its host lifecycle.  While it runs, the ReID half
// of the NEXT slot is started as soon as that slot's detector half is seen to be complete.
static int pipeline_collect(ydst_pipeline* p, int32_t* out_host, int* k_host, float* dets_host, int* n_dets_host) {
    YDST_CHECK(p->collected < p->submitted, "no frame in flight: submit one first");
    const int si = (int)(p->drain % ydst_pipeline::kSlots);
    ydst_pipeline::Slot& sl = p->slot[si];
    ydst_pipeline::Slot& nxt = p->slot[(si + 1) % ydst_pipeline::kSlots];
    YDST_CHECK(sl.busy && sl.n_collected < sl.n_frames, "pipeline slot bookkeeping is inconsistent");
    const int sub = sl.n_collected++;
    ++p->collected;
    struct Release {                                                   // the slot is free again once its last frame has been handed back
        ydst_pipeline* p; ydst_pipeline::Slot& sl;
        ~Release() { if (sl.n_collected == sl.n_frames && sl.launched) { sl.busy = false; ++p->drain; } }
    } release{p, sl};
    Detector& det = *p->det;
    const int md = det.nms_.max_det;
    if (!sl.launched) pipeline_launch_detector(p, sl);                 // partial batch
    pipeline_launch_reid(p, sl, true);
    if (!sl.feat_ready) {
        YDST_CUDA(cudaStreamWaitEvent(p->sB, sl.ev_feat, 0));          // the tracker's kernels read this slot's features
        sl.feat_ready = true;
    }
    if (nxt.busy) pipeline_launch_reid(p, nxt, false);
    const int n_dets = sl.h_counts[sub * 8 + 1], m = sl.h_counts[sub * 8 + 3];
    if (n_dets_host) *n_dets_host = n_dets;
    if (dets_host && sl.want_dets) {
        memcpy(dets_host, sl.h_dets + (size_t)sub * md * 6, sizeof(float) * 6 * n_dets);
        if (sl.resized[sub]) {                                       // resize_boxes (yolo3/utils/model_build.py:12-19), same fp32 products
            const float rw = (float)((double)sl.fw[sub] / (double)det.W), rh = (float)((double)sl.fh[sub] / (double)det.H);
            for (int i = 0; i < n_dets; ++i) {
                float* d = dets_host + (size_t)i * 6;
                d[0] *= rw; d[1] *= rh; d[2] *= rw; d[3] *= rh;
            }
        }
    }
    p->last_slot = si; p->last_sub = sub; p->last_m = n_dets > 0 ? m : 0;
    // the reference raises inside cv2.resize -- BEFORE tracker.update (deep_sort/deep/feature_extractor.py:45) -- when a box has
    // no pixels inside the frame: make this frame's flag host-visible and test it before the tracker consumes the features
    if (m > 0) {
        YDST_CUDA(cudaEventSynchronize(sl.ev_feat));
        if (sl.h_counts[sub * 8 + 4]) {
            set_error("empty crop: a detection has no pixels inside the frame (cv2.resize raises in the reference)");
            *k_host = -1;
            return 3;
        }
    }
    if (n_dets == 0) { *k_host = -1; return 0; }          // the reference skips tracker.update when nothing was detected
    const float* h_cls = sl.h_cls + (size_t)sub * md;
    for (int i = 0; i < m; ++i) p->h_payload[i] = (int)h_cls[i];
    p->trk->update(sl.tlwh + (size_t)sub * md * 4, sl.feat + (size_t)sl.feat_off[sub] * 512, p->h_payload, nullptr, m, out_host, k_host, p->sB);
    if (nxt.busy) pipeline_launch_reid(p, nxt, false);
    return 0;
}

int ydst_pipeline_submit(ydst_pipeline* p, const uint8_t* frame, int frame_is_host, int want_dets, void* stream) {
    YDST_API_BEGIN
    YDST_CHECK(p && frame, "null argument");
    pipeline_submit(p, frame, frame_is_host != 0, want_dets != 0, S(stream));
    YDST_API_END
}
int ydst_pipeline_submit_frame(ydst_pipeline* p, const uint8_t* frame, int height, int width, int frame_is_host, int is_bgr, int want_dets,
                               void* stream) {
    YDST_API_BEGIN
    YDST_CHECK(p && frame && height > 0 && width > 0, "bad argument");
    pipeline_submit(p, frame, frame_is_host != 0, want_dets != 0, S(stream), height, width, is_bgr != 0);
    YDST_API_END
}
int ydst_resize_u8(const uint8_t* src_dev, int src_h, int src_w, uint8_t* dst_dev, int dst_h, int dst_w, int swap_rb, void* stream) {
    YDST_API_BEGIN
    YDST_CHECK(src_dev && dst_dev, "null argument");
    launch_resize_u8(src_dev, src_h, src_w, dst_dev, dst_h, dst_w, swap_rb, S(stream));
    count_launch();
    YDST_CUDA(cudaStreamSynchronize(S(stream)));
    YDST_API_END
}
int ydst_resize_u8_roi(const uint8_t* src_dev, int src_h, int src_w, int x0, int y0, int roi_w, int roi_h, uint8_t* dst_dev, int dst_h,
                       int dst_w, int swap_rb, void* stream) {
    YDST_API_BEGIN
    YDST_CHECK(src_dev && dst_dev, "null argument");
    YDST_CHECK(x0 >= 0 && y0 >= 0 && roi_w > 0 && roi_h > 0 && x0 + roi_w <= src_w && y0 + roi_h <= src_h, "the window leaves the frame");
    launch_resize_u8(src_dev + ((size_t)y0 * src_w + x0) * 3, roi_h, roi_w, dst_dev, dst_h, dst_w, swap_rb, S(stream), (long long)src_w * 3);
    count_launch();
    YDST_API_END
}
int ydst_pipeline_collect(ydst_pipeline* p, int32_t* out_host, int* k_host, float* dets_host, int* n_dets_host) {
    YDST_API_BEGIN
    YDST_CHECK(p && out_host && k_host, "null argument");
    const int rc = pipeline_collect(p, out_host, k_host, dets_host, n_dets_host);
    if (rc) return rc;
    YDST_API_END
}
int ydst_pipeline_last_inputs(ydst_pipeline* p, float* tlwh_host, float* feat_host, int32_t* cls_host, int cap_rows, int* m_host) {
    YDST_API_BEGIN
    YDST_CHECK(p && m_host, "null argument");
    YDST_CHECK(p->last_slot >= 0, "no frame has been collected yet");
    const ydst_pipeline::Slot& sl = p->slot[p->last_slot];
    const int m = p->last_m, md = p->det->nms_.max_det;
    *m_host = m;
    YDST_CHECK(m <= cap_rows, "buffer too small (%d rows, capacity %d)", m, cap_rows);
    if (m > 0) {
        YDST_CUDA(cudaStreamSynchronize(p->sB));
        if (tlwh_host) YDST_CUDA(cudaMemcpy(tlwh_host, sl.tlwh + (size_t)p->last_sub * md * 4, sizeof(float) * 4 * m, cudaMemcpyDeviceToHost));
        if (feat_host) YDST_CUDA(cudaMemcpy(feat_host, sl.feat + (size_t)sl.feat_off[p->last_sub] * 512, sizeof(float) * 512 * m, cudaMemcpyDeviceToHost));
        if (cls_host) for (int i = 0; i < m; ++i) cls_host[i] = (int32_t)sl.h_cls[(size_t)p->last_sub * md + i];
    }
    YDST_API_END
}
int ydst_pipeline_in_flight(const ydst_pipeline* p) { return p ? (int)(p->submitted - p->collected) : 0; }
int ydst_pipeline_can_submit(const ydst_pipeline* p) { return p && pipeline_can_submit(p) ? 1 : 0; }

int ydst_pipeline_step(ydst_pipeline* p, const uint8_t* frame_host, int32_t* out_host, int* k_host, float* dets_host, int* n_dets_host,
                       void* stream) {
    YDST_API_BEGIN
    YDST_CHECK(p && frame_host && out_host && k_host, "null argument");
    YDST_CHECK(p->submitted == p->collected, "ydst_pipeline_step needs an empty pipeline (collect the frames in flight first)");
    pipeline_submit(p, frame_host, true, dets_host != nullptr, S(stream));
    const int rc = pipeline_collect(p, out_host, k_host, dets_host, n_dets_host);
    if (rc) return rc;
    YDST_API_END
}
int ydst_pipeline_step_dev(ydst_pipeline* p, const uint8_t* frame_dev, int32_t* out_host, int* k_host, float* dets_host, int* n_dets_host,
                           void* stream) {
    YDST_API_BEGIN
    YDST_CHECK(p && frame_dev && out_host && k_host, "null argument");
    YDST_CHECK(p->submitted == p->collected, "ydst_pipeline_step_dev needs an empty pipeline (collect the frames in flight first)");
    pipeline_submit(p, frame_dev, false, dets_host != nullptr, S(stream));
    const int rc = pipeline_collect(p, out_host, k_host, dets_host, n_dets_host);
    if (rc) return rc;
    YDST_API_END
}

}  // extern "C"
