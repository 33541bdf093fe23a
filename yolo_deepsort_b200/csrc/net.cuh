// Layer-graph planner/executor shared by the detector (Darknet cfg) and the ReID net.  See net.cu.
#pragma once
#include <map>
#include <memory>
#include <vector>

#include "../../include/ydst.h"
#include "conv_tc.cuh"
#include "layers.cuh"
#include "nms.cuh"

namespace ydst {

struct ConvWeights {
    int cin = 0, cout = 0, k = 0, cout16 = 0;
    __half* w16 = nullptr;    // [cout16][k*k*cin]  (tensor-core path)
    float* w32 = nullptr;     // [27][cout]         (first-layer path, CUDA cores)
    __half* w_hilo = nullptr; // [2][cout][32] fp16: hi = fp16(w), lo = fp16(w - hi), K = 27 padded to 32 (first-layer path, tensor cores)
    float* scale = nullptr;   // [cout rounded up to 256]
    float* bias = nullptr;
};

enum OpKind { OP_CONV_TC, OP_CONV_FIRST, OP_MAXPOOL, OP_UPSAMPLE, OP_ADD, OP_COPY, OP_YOLO, OP_AVGPOOL_L2, OP_CONV_FIRST_POOL };

struct Op {
    OpKind kind;
    ConvTcLaunch conv;           // OP_CONV_TC
    Act a, b, out;               // generic operands
    int i0 = 0, i1 = 0, i2 = 0, i3 = 0;
    const ConvWeights* w = nullptr;
    const float* fsrc = nullptr; // fp32 source (first-layer input / yolo head)
    float* fdst = nullptr;       // fp32 destination
    float anchors[6] = {0, 0, 0, 0, 0, 0};
    int layer = -1;              // cfg layer index (profiling / debugging)
};

class DeviceArena {
public:
    ~DeviceArena();
    void* alloc(size_t bytes, bool zero = true);
    size_t total = 0;
private:
    std::vector<void*> ptrs_;
};

// Owns activation buffers + op list for a fixed batch size.
struct Plan {
    std::vector<Op> ops;
    int launches = 0;
    double flops = 0;
    // The op list is static (fixed buffers, fixed shapes), so after one eager run it is captured into a CUDA graph:
    // one launch per forward, kernel->kernel edges (programmatic where the convs ask for it) instead of ~100 stream launches.
    mutable cudaGraphExec_t exec = nullptr;
    mutable int runs = 0;
    Plan() = default;
    Plan(const Plan&) = delete;
    Plan& operator=(const Plan&) = delete;
    Plan(Plan&& o) noexcept : ops(std::move(o.ops)), launches(o.launches), flops(o.flops), exec(o.exec), runs(o.runs) { o.exec = nullptr; }
    ~Plan() { if (exec) cudaGraphExecDestroy(exec); }
};
void run_plan(const Plan& plan, cudaStream_t st);

// Optional per-op timing (CUDA events recorded on the launching stream around every op of run_plan).
struct OpSample { int kind; int layer; double flops; double bytes; cudaEvent_t e0, e1; };
void profile_begin();
bool profile_active();
std::vector<OpSample>& profile_samples();
void profile_push(const OpSample& s);
// OpSample.kind of the association kernels (tracker.cu), next to the OpKind values of the layer graphs
enum TrackerOpKind { TOP_KF_PREDICT = 100, TOP_NORMALIZE, TOP_FILL, TOP_COSINE_MIN, TOP_COST_FINALIZE, TOP_LSAP, TOP_IOU_COST, TOP_KF_UPDATE,
                     TOP_KF_INITIATE, TOP_GALLERY_APPEND, TOP_GATHER, TOP_TRANSPOSE };

class Detector {
public:
    Detector(const ydst_layer_desc* layers, int n, const float* weights, size_t n_weights, int H, int W, int batch);
    void forward_u8(const uint8_t* frame_dev, float* pred_out, cudaStream_t st);
    void forward_nchw(const void* x_dev, int is_half, float* pred_out, cudaStream_t st);
    void nms(float conf, float iou, float* dets_out, int* n_out, cudaStream_t st);
    int H, W, batch, rows = 0, fields = 0;
    float* pred = nullptr;       // [batch][rows][fields]
    Nms nms_;
    Plan plan;
    // debug/parity view: output of cfg layer `l` of the last forward, unpacked to dense NHWC fp16 (N,H,W,C)
    void layer_shape(int l, int* n, int* h, int* w, int* c, int* is_f32) const;
    void layer_output(int l, void* dense_out, cudaStream_t st) const;
private:
    std::vector<Act> out_;            // per cfg layer (aliases for routes / fused shortcuts)
    std::vector<float*> head_f32_;    // fp32 buffers of the yolo head convs
    void build(const ydst_layer_desc* layers, int n, const float* weights, size_t n_weights);
    DeviceArena arena_;
    ConvWorkspace ws_;
    std::vector<std::unique_ptr<ConvWeights>> weights_;
    float* in_f32_ = nullptr;    // [batch][H][W][3]
};

class Reid {
public:
    Reid(const float* weights, size_t n_weights, int max_batch);
    void extract(const uint8_t* frame_dev, int H, int W, const float* tlwh_dev, int m, float* feat_out, cudaStream_t st);
    void forward(const float* x_dev, int m, float* feat_out, cudaStream_t st);
    // crops of several frames through ONE forward: frame b contributes m[b] boxes (tlwh[b]); features are written frame after frame
    void extract_multi(const uint8_t* const* frames_dev, const int* H, const int* W, const float* const* tlwh_dev, const int* m, int nb,
                       float* feat_out, cudaStream_t st);
    int max_batch;
    int* err_flag = nullptr;     // device, [8]: frame b of an extract_multi call raises err_flag[b] when one of its crops is empty
private:
    const Plan& plan_for(int m);
    DeviceArena arena_;
    ConvWorkspace ws_;
    std::vector<std::unique_ptr<ConvWeights>> weights_;
    std::map<int, Plan> plans_;
    float* in_f32_ = nullptr;    // [max_batch][128][64][3]
    float* feat_ = nullptr;      // [max_batch][512]
    std::vector<Act> bufs_;      // activation buffers allocated for max_batch
};

// Packs one conv's weights (fp32 OIHW + BN/bias) into device buffers.
std::unique_ptr<ConvWeights> pack_conv(DeviceArena& arena, const float* w_oihw, int cout, int cin, int k, const float* gamma,
                                       const float* beta, const float* mean, const float* var, const float* conv_bias, bool first_layer);
Act make_act(DeviceArena& arena, int N, int H, int W, int C);
// split-K scratch for one owner's convolutions (zeroed tickets; see conv_tc.cuh)
ConvWorkspace make_conv_workspace(DeviceArena& arena, size_t partial_bytes = (size_t)48 << 20, int n_tickets = 8192);

}  // namespace ydst
