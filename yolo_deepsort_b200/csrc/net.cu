// Layer-graph planner/executor: turns a Darknet layer list (or the fixed ReID architecture) into a flat
// list of kernel launches over flat-padded NHWC fp16 buffers.
//
// Reference behaviour reproduced here (yolo3/models/models.py): conv padding (k-1)//2 regardless of the cfg
// `pad` (:39), bias only without BN (:48), BN eps 1e-5 (:52), leaky 0.1 (:54), maxpool k2/s1 preceded by a
// ZERO pad right/bottom (:61-63), nearest upsample (:132), route = concat of absolute/relative layer
// outputs with optional channel-group select (:300-303), shortcut = out[-1] + out[from] with its own
// activation ignored (:304-306), heads concatenated in cfg order (:312).
//
// Fusions: BN folded into an fp32 scale/bias epilogue; activation in the conv epilogue; a conv whose only
// reader is the following shortcut adds the residual in its epilogue; head convs write fp32 for the decode.
#include "net.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace ydst {

DeviceArena::~DeviceArena() {
    for (void* p : ptrs_) cudaFree(p);
}
void* DeviceArena::alloc(size_t bytes, bool zero) {
    void* p = nullptr;
    bytes = (bytes + 255) & ~(size_t)255;
    YDST_CUDA(cudaMalloc(&p, bytes));
    if (zero) YDST_CUDA(cudaMemset(p, 0, bytes));
    ptrs_.push_back(p);
    total += bytes;
    return p;
}

Act make_act(DeviceArena& arena, int N, int H, int W, int C) {
    Act a;
    a.N = N; a.H = H; a.W = W; a.C = C; a.ctot = C; a.coff = 0;
    a.base = (__half*)arena.alloc((size_t)a.pixels() * C * sizeof(__half));
    return a;
}

ConvWorkspace make_conv_workspace(DeviceArena& arena, size_t partial_bytes, int n_tickets) {
    ConvWorkspace w;
    w.partial = (float*)arena.alloc(partial_bytes, false);
    w.partial_bytes = partial_bytes;
    w.tickets = (int*)arena.alloc(sizeof(int) * n_tickets, true);
    w.n_tickets = n_tickets;
    return w;
}

std::unique_ptr<ConvWeights> pack_conv(DeviceArena& arena, const float* w, int cout, int cin, int k, const float* gamma,
                                       const float* beta, const float* mean, const float* var, const float* conv_bias, bool first_layer) {
    auto cw = std::make_unique<ConvWeights>();
    cw->cin = cin; cw->cout = cout; cw->k = k; cw->cout16 = (cout + 15) & ~15;
    const int taps = k * k;
    const int npad = (cout + 255) & ~255;
    std::vector<float> sc(npad, 0.f), bi(npad, 0.f);
    for (int o = 0; o < cout; ++o) {
        if (gamma) {
            const double s = (double)gamma[o] / std::sqrt((double)var[o] + 1e-5);
            const double b0 = conv_bias ? (double)conv_bias[o] : 0.0;
            sc[o] = (float)s;
            bi[o] = (float)((b0 - (double)mean[o]) * s + (double)beta[o]);
        } else {
            sc[o] = 1.f;
            bi[o] = conv_bias ? conv_bias[o] : 0.f;
        }
    }
    cw->scale = (float*)arena.alloc(sizeof(float) * npad);
    cw->bias = (float*)arena.alloc(sizeof(float) * npad);
    YDST_CUDA(cudaMemcpy(cw->scale, sc.data(), sizeof(float) * npad, cudaMemcpyHostToDevice));
    YDST_CUDA(cudaMemcpy(cw->bias, bi.data(), sizeof(float) * npad, cudaMemcpyHostToDevice));
    if (first_layer) {
        YDST_CHECK(cin == 3 && k == 3, "first-layer path is 3x3 over 3 channels");
        std::vector<float> p((size_t)27 * cout);
        for (int o = 0; o < cout; ++o)
            for (int c = 0; c < 3; ++c)
                for (int r = 0; r < 3; ++r)
                    for (int s = 0; s < 3; ++s) p[(size_t)((r * 3 + s) * 3 + c) * cout + o] = w[((o * 3 + c) * 3 + r) * 3 + s];
        cw->w32 = (float*)arena.alloc(p.size() * sizeof(float));
        YDST_CUDA(cudaMemcpy(cw->w32, p.data(), p.size() * sizeof(float), cudaMemcpyHostToDevice));
        // tensor-core variant: the fp32 weight as a sum of two fp16 numbers (the products then carry ~22 significant bits)
        std::vector<__half> hl((size_t)2 * cout * 32, __float2half(0.f));
        for (int o = 0; o < cout; ++o)
            for (int k = 0; k < 27; ++k) {
                const float v = p[(size_t)k * cout + o];
                const __half hi = __float2half_rn(v);
                hl[(size_t)o * 32 + k] = hi;
                hl[(size_t)(cout + o) * 32 + k] = __float2half_rn(v - __half2float(hi));
            }
        cw->w_hilo = (__half*)arena.alloc(hl.size() * sizeof(__half));
        YDST_CUDA(cudaMemcpy(cw->w_hilo, hl.data(), hl.size() * sizeof(__half), cudaMemcpyHostToDevice));
    } else {
        const size_t K = (size_t)taps * cin;
        std::vector<__half> p((size_t)cw->cout16 * K, __float2half(0.f));
        for (int o = 0; o < cout; ++o)
            for (int c = 0; c < cin; ++c)
                for (int t = 0; t < taps; ++t) p[(size_t)o * K + (size_t)t * cin + c] = __float2half_rn(w[((size_t)o * cin + c) * taps + t]);
        cw->w16 = (__half*)arena.alloc(p.size() * sizeof(__half));
        YDST_CUDA(cudaMemcpy(cw->w16, p.data(), p.size() * sizeof(__half), cudaMemcpyHostToDevice));
    }
    return cw;
}

// every tensor-core conv prefetches the weights of the next one into L2 (YDST_WEIGHT_PREFETCH=0 disables)
static void link_weight_prefetch(Plan& plan) {
    const char* e = getenv("YDST_WEIGHT_PREFETCH");
    if (e && atoi(e) == 0) return;
    Op* prev = nullptr;
    for (Op& op : plan.ops) {
        if (op.kind != OP_CONV_TC) continue;
        if (prev) { prev->conv.p.pf_ptr = op.conv.w_ptr; prev->conv.p.pf_bytes = op.conv.w_bytes; }
        prev = &op;
    }
}

static bool g_profiling = false;
static std::vector<OpSample> g_samples;
void profile_begin() { g_profiling = true; g_samples.clear(); }
bool profile_active() { return g_profiling; }
std::vector<OpSample>& profile_samples() { g_profiling = false; return g_samples; }
void profile_push(const OpSample& s) { g_samples.push_back(s); }

// algorithmic bytes of one conv: activations in + out (fp16, logical dims) + weights
static double conv_bytes(const ConvTcLaunch& L) {
    const ConvTcParams& p = L.p;
    const double stride2 = p.mode == 1 ? 4.0 : 1.0;
    const double in_px = (double)p.N * p.Ho * p.Wo * stride2;
    return 2.0 * (in_px * p.cin + (double)p.N * p.Ho * p.Wo * p.cout * (p.out_f32 ? 2.0 : 1.0) + (double)p.cout * p.R * p.S * p.cin);
}

static void run_ops(const Plan& plan, cudaStream_t st);

static int graphs_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("YDST_GRAPH");
        v = e ? atoi(e) : 1;
        if (getenv("YDST_CONV_TRACE") && atoi(getenv("YDST_CONV_TRACE")) == 1) v = 0;   // that mode synchronises after every conv
    }
    return v;
}

void run_plan(const Plan& plan, cudaStream_t st) {
    if (g_profiling || !graphs_enabled()) { run_ops(plan, st); return; }
    if (!plan.exec) {
        if (plan.runs++ == 0) { run_ops(plan, st); return; }             // first run eager: one-time attribute / driver-entry setup
        // capture on a private stream (the caller's may be the legacy default stream, which cannot be captured)
        static thread_local cudaStream_t cap_dev[64] = {nullptr};          // one per device: a stream belongs to the device it was created on
        int dev = 0;
        YDST_CUDA(cudaGetDevice(&dev));
        cudaStream_t& cap = cap_dev[dev & 63];
        if (!cap) YDST_CUDA(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
        cudaGraph_t g = nullptr;
        YDST_CUDA(cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal));
        try { run_ops(plan, cap); }
        catch (...) { cudaStreamEndCapture(cap, &g); if (g) cudaGraphDestroy(g); throw; }
        YDST_CUDA(cudaStreamEndCapture(cap, &g));
        YDST_CUDA(cudaGraphInstantiate(&plan.exec, g, 0));
        YDST_CUDA(cudaGraphDestroy(g));
        count_launch(-plan.launches);                                      // the capture pass launched nothing
    }
    YDST_CUDA(cudaGraphLaunch(plan.exec, st));
    count_launch(plan.launches);
}

static void run_ops(const Plan& plan, cudaStream_t st) {
    for (const Op& op : plan.ops) {
        OpSample smp;
        if (g_profiling) {
            smp.kind = op.kind; smp.layer = op.layer;
            smp.flops = op.kind == OP_CONV_TC ? conv_tc_flops(op.conv) : 0.0;
            smp.bytes = op.kind == OP_CONV_TC ? conv_bytes(op.conv) : 0.0;
            cudaEventCreate(&smp.e0); cudaEventCreate(&smp.e1);
            cudaEventRecord(smp.e0, st);
        }
        count_launch();
        switch (op.kind) {
            case OP_CONV_TC: conv_tc_run(op.conv, st); break;
            case OP_CONV_FIRST:
                launch_conv_first(op.fsrc, op.out.N, op.i0, op.i1, op.w->w32, op.w->w_hilo, op.w->scale, op.w->bias, op.w->cout, op.i2, op.i3, op.out, st);
                break;
            case OP_CONV_FIRST_POOL:
                launch_conv_first_pool(op.fsrc, op.out.N, op.i0, op.i1, op.w->w_hilo, op.w->scale, op.w->bias, op.w->cout, op.i3, op.out, st);
                break;
            case OP_MAXPOOL: launch_maxpool(op.a, op.out, op.i0, op.i1, op.i2, st); break;
            case OP_UPSAMPLE: launch_upsample(op.a, op.out, op.i0, st); break;
            case OP_ADD: launch_add(op.a, op.b, op.out, st); break;
            case OP_COPY: launch_copy(op.a, op.out, st); break;
            case OP_YOLO:
                launch_yolo_decode(op.fsrc, op.i0, op.a.N, op.a.H, op.a.W, 3, op.anchors, op.i1, op.i2, op.i3, op.fdst, op.out.N /*rows_total*/,
                                   op.out.H /*row0*/, st);
                break;
            case OP_AVGPOOL_L2: launch_avgpool_l2(op.a, op.fdst, st); break;
        }
        if (g_profiling) { cudaEventRecord(smp.e1, st); g_samples.push_back(smp); }
    }
}

// ================================================================================================
// Detector
// ================================================================================================
Detector::Detector(const ydst_layer_desc* layers, int n, const float* weights, size_t n_weights, int H_, int W_, int batch_)
    : H(H_), W(W_), batch(batch_) {
    YDST_CHECK(batch >= 1 && H > 0 && W > 0, "bad detector geometry");
    build(layers, n, weights, n_weights);
    nms_.init(4096, 300, std::min(batch, 8));
}

void Detector::build(const ydst_layer_desc* L, int n, const float* weights, size_t n_weights) {
    in_f32_ = (float*)arena_.alloc((size_t)batch * H * W * 3 * sizeof(float));
    ws_ = make_conv_workspace(arena_);
    // readers of each layer's output
    std::vector<std::vector<int>> readers(n);
    for (int l = 0; l < n; ++l) {
        const int t = L[l].type;
        if (t == YDST_ROUTE) {
            for (int s = 0; s < L[l].n_src; ++s) {
                YDST_CHECK(L[l].src[s] >= 0 && L[l].src[s] < l, "route source out of range at layer %d", l);
                readers[L[l].src[s]].push_back(l);
            }
        } else {
            if (l > 0) readers[l - 1].push_back(l);
            if (t == YDST_SHORTCUT) {
                YDST_CHECK(L[l].src[0] >= 0 && L[l].src[0] < l, "shortcut source out of range at layer %d", l);
                readers[L[l].src[0]].push_back(l);
            }
        }
    }
    // first pass: YOLO geometry (row offsets) needs every head's grid size, known only after shapes propagate
    struct Shape { int C, H, W; };
    std::vector<Shape> shp(n);
    std::vector<int> yolo_layers;
    {
        Shape cur{3, H, W};
        for (int l = 0; l < n; ++l) {
            const ydst_layer_desc& d = L[l];
            switch (d.type) {
                case YDST_CONV: {
                    const int pad = (d.size - 1) / 2;
                    cur = Shape{d.filters, (cur.H + 2 * pad - d.size) / d.stride + 1, (cur.W + 2 * pad - d.size) / d.stride + 1};
                    break;
                }
                case YDST_MAXPOOL: {
                    if (d.size == 2 && d.stride == 1) break;                       // zero-pad right/bottom keeps the size
                    const int pad = (d.size - 1) / 2;
                    cur = Shape{cur.C, (cur.H + 2 * pad - d.size) / d.stride + 1, (cur.W + 2 * pad - d.size) / d.stride + 1};
                    break;
                }
                case YDST_UPSAMPLE: cur = Shape{cur.C, cur.H * d.size, cur.W * d.size}; break;
                case YDST_ROUTE: {
                    int c = 0;
                    for (int s = 0; s < d.n_src; ++s) {
                        c += shp[d.src[s]].C;
                        YDST_CHECK(shp[d.src[s]].H == shp[d.src[0]].H && shp[d.src[s]].W == shp[d.src[0]].W, "route spatial mismatch at layer %d", l);
                    }
                    if (d.groups > 0) c /= d.groups;
                    cur = Shape{c, shp[d.src[0]].H, shp[d.src[0]].W};
                    break;
                }
                case YDST_SHORTCUT: cur = shp[d.src[0]]; break;
                case YDST_YOLO: yolo_layers.push_back(l); break;
                default: YDST_CHECK(false, "unknown layer type %d at layer %d", d.type, l);
            }
            shp[l] = cur;
        }
    }
    YDST_CHECK(!yolo_layers.empty(), "network has no yolo layer");
    fields = 5 + L[yolo_layers[0]].classes;
    rows = 0;
    std::vector<int> row0(n, 0);
    for (int yl : yolo_layers) {
        YDST_CHECK(5 + L[yl].classes == fields, "yolo layers disagree on the class count");
        row0[yl] = rows;
        rows += 3 * shp[yl].H * shp[yl].W;
    }
    pred = (float*)arena_.alloc((size_t)batch * rows * fields * sizeof(float));

    // Zero-copy concatenation: a multi-source route reads a buffer that its producers wrote into directly -- each eligible source
    // layer's output IS a channel slice of the route's buffer (every kernel takes (ctot, coff) views), so no copy kernel runs for
    // it.  yolov4's CSP blocks concatenate up to 304x304x128 tensors: 22 copy launches, 0.59 ms of a 4.1 ms forward, before this.
    // Eligible: a convolution (not the first layer, not a YOLO head), max-pool, upsample or shortcut that is not yet part of
    // another concatenation; anything else (aliases of earlier layers, grouped routes) is copied into its slice as before.
    struct Slice { int route = -1, coff = 0; };
    std::vector<Slice> slice(n);
    static const bool zero_copy = !(getenv("YDST_ZERO_COPY_ROUTE") && atoi(getenv("YDST_ZERO_COPY_ROUTE")) == 0);
    if (zero_copy)
        for (int l = 0; l < n; ++l) {
            if (L[l].type != YDST_ROUTE || L[l].n_src < 2 || L[l].groups > 0) continue;
            int coff = 0;
            for (int k = 0; k < L[l].n_src; ++k) {
                const int sidx = L[l].src[k];
                const int t = L[sidx].type;
                const bool head = t == YDST_CONV && sidx + 1 < n && L[sidx + 1].type == YDST_YOLO;
                const bool ok = (t == YDST_CONV && sidx > 0 && !head) || t == YDST_MAXPOOL || t == YDST_UPSAMPLE || t == YDST_SHORTCUT;
                bool dup = false;
                for (int j = 0; j < k; ++j) dup = dup || L[l].src[j] == sidx;
                if (ok && !dup && slice[sidx].route < 0 && shp[sidx].C % 8 == 0 && coff % 8 == 0) { slice[sidx].route = l; slice[sidx].coff = coff; }
                coff += shp[sidx].C;
            }
        }
    std::vector<Act> concat(n);                              // the route buffers, allocated when their first producer needs them
    auto alloc_out = [&](int l) -> Act {
        if (slice[l].route < 0) return make_act(arena_, batch, shp[l].H, shp[l].W, shp[l].C);
        const int r = slice[l].route;
        if (!concat[r].base) concat[r] = make_act(arena_, batch, shp[r].H, shp[r].W, shp[r].C);
        Act v = concat[r];
        v.C = shp[l].C; v.coff = slice[l].coff;
        return v;
    };

    // second pass: buffers + ops
    std::vector<Act> out(n);
    std::vector<float*> head_f32(n, nullptr);
    std::vector<bool> fused_away(n, false);
    size_t wp = 0;
    auto take = [&](size_t cnt) {
        YDST_CHECK(wp + cnt <= n_weights, "weights payload too short: need %zu floats, have %zu", wp + cnt, n_weights);
        const float* p = weights + wp;
        wp += cnt;
        return p;
    };
    for (int l = 0; l < n; ++l) {
        const ydst_layer_desc& d = L[l];
        Op op;
        op.layer = l;
        switch (d.type) {
            case YDST_CONV: {
                const int cin = l == 0 ? 3 : shp[l - 1].C;
                const float *gamma = nullptr, *beta = nullptr, *mean = nullptr, *var = nullptr, *cb = nullptr;
                if (d.batch_normalize) { beta = take(d.filters); gamma = take(d.filters); mean = take(d.filters); var = take(d.filters); }
                else cb = take(d.filters);
                const float* w = take((size_t)d.filters * cin * d.size * d.size);
                const bool first = (l == 0);
                YDST_CHECK(first || cin % 16 == 0, "layer %d: Cin=%d is not a multiple of 16", l, cin);
                weights_.push_back(pack_conv(arena_, w, d.filters, cin, d.size, gamma, beta, mean, var, cb, first));
                const ConvWeights* cw = weights_.back().get();
                const bool to_yolo = l + 1 < n && L[l + 1].type == YDST_YOLO;
                const bool fuse_sc = !to_yolo && l + 1 < n && L[l + 1].type == YDST_SHORTCUT && readers[l].size() == 1 &&
                                     L[l + 1].src[0] != l && !first;
                if (to_yolo) {
                    YDST_CHECK(readers[l].size() == 1, "layer %d: a yolo head conv must only feed its yolo layer", l);
                    head_f32[l] = (float*)arena_.alloc((size_t)batch * (shp[l].H + 2) * (shp[l].W + 2) * cw->cout16 * sizeof(float));
                    Act geo;                                  // geometry only (fp32 destination)
                    geo.N = batch; geo.H = shp[l].H; geo.W = shp[l].W; geo.C = d.filters; geo.ctot = cw->cout16; geo.coff = 0;
                    op.kind = OP_CONV_TC;
                    conv_tc_plan(op.conv, out[l - 1], geo, cw->w16, d.size, d.size, d.stride, cw->scale, cw->bias, d.activation, 0, nullptr,
                                 head_f32[l], d.filters, &ws_);
                    out[l] = geo;
                } else {
                    // (a convolution with a fused shortcut writes the shortcut layer's tensor: that layer's slice, if it has one)
                    out[l] = (fuse_sc && slice[l + 1].route >= 0 && slice[l].route < 0) ? alloc_out(l + 1) : alloc_out(l);
                    if (first) {
                        op.kind = OP_CONV_FIRST;
                        op.fsrc = in_f32_; op.out = out[l]; op.w = cw;
                        op.i0 = H; op.i1 = W; op.i2 = d.stride; op.i3 = d.activation;
                        YDST_CHECK(d.size == 3, "first layer must be 3x3");
                    } else {
                        op.kind = OP_CONV_TC;
                        const Act* res = nullptr;
                        if (fuse_sc) { res = &out[L[l + 1].src[0]]; fused_away[l + 1] = true; }
                        conv_tc_plan(op.conv, out[l - 1], out[l], cw->w16, d.size, d.size, d.stride, cw->scale, cw->bias, d.activation,
                                     fuse_sc ? 1 : 0, res, nullptr, d.filters, &ws_);
                    }
                }
                if (op.kind == OP_CONV_TC) plan.flops += conv_tc_flops(op.conv) * ((double)d.filters / op.conv.p.cout);
                else plan.flops += 2.0 * batch * shp[l].H * shp[l].W * d.filters * 27.0;
                plan.ops.push_back(op);
                break;
            }
            case YDST_MAXPOOL: {
                out[l] = alloc_out(l);
                op.kind = OP_MAXPOOL; op.a = out[l - 1]; op.out = out[l];
                op.i0 = d.size; op.i1 = d.stride; op.i2 = (d.size == 2 && d.stride == 1) ? 1 : 0;
                plan.ops.push_back(op);
                break;
            }
            case YDST_UPSAMPLE: {
                out[l] = alloc_out(l);
                op.kind = OP_UPSAMPLE; op.a = out[l - 1]; op.out = out[l]; op.i0 = d.size;
                plan.ops.push_back(op);
                break;
            }
            case YDST_ROUTE: {
                if (d.n_src == 1) {
                    out[l] = out[d.src[0]];                                   // alias
                    YDST_CHECK(head_f32[d.src[0]] == nullptr, "route from a yolo head conv is not supported (layer %d)", l);
                    if (d.groups > 0) {
                        out[l].C = out[l].C / d.groups;
                        out[l].coff += d.group_id * out[l].C;
                    }
                } else {
                    YDST_CHECK(d.groups <= 0, "grouped multi-source route is not supported (layer %d)", l);
                    if (!concat[l].base) concat[l] = make_act(arena_, batch, shp[l].H, shp[l].W, shp[l].C);
                    out[l] = concat[l];
                    int coff = 0;
                    for (int s = 0; s < d.n_src; ++s) {
                        const Act& src = out[d.src[s]];
                        const bool in_place = src.base == out[l].base && src.coff == coff && src.ctot == out[l].ctot;   // written there by its producer
                        if (!in_place) {
                            Op cp;
                            cp.layer = l; cp.kind = OP_COPY; cp.a = src;
                            cp.out = out[l]; cp.out.C = cp.a.C; cp.out.coff = coff;
                            plan.ops.push_back(cp);
                        }
                        coff += src.C;
                    }
                }
                break;
            }
            case YDST_SHORTCUT: {
                if (fused_away[l]) { out[l] = out[l - 1]; break; }            // produced by the conv epilogue
                out[l] = alloc_out(l);
                op.kind = OP_ADD; op.a = out[l - 1]; op.b = out[d.src[0]]; op.out = out[l];
                plan.ops.push_back(op);
                break;
            }
            case YDST_YOLO: {
                YDST_CHECK(l > 0 && head_f32[l - 1] != nullptr, "yolo layer %d must follow a convolution", l);
                YDST_CHECK(shp[l - 1].C == 3 * fields, "yolo layer %d: head has %d channels, expected %d", l, shp[l - 1].C, 3 * fields);
                op.kind = OP_YOLO;
                op.fsrc = head_f32[l - 1];
                op.i0 = out[l - 1].ctot;                  // fp32 row stride (cout16)
                op.a.N = batch; op.a.H = shp[l].H; op.a.W = shp[l].W;
                op.i1 = d.classes; op.i2 = H; op.i3 = W;
                op.fdst = pred;
                op.out.N = rows; op.out.H = row0[l];      // rows_total / row0 (see run_plan)
                memcpy(op.anchors, d.anchors, sizeof(op.anchors));
                out[l] = out[l - 1];
                plan.ops.push_back(op);
                break;
            }
        }
    }
    YDST_CHECK(wp == n_weights, "weights payload has %zu floats, network consumes %zu", n_weights, wp);
    link_weight_prefetch(plan);
    plan.launches = (int)plan.ops.size();
    out_ = out;
    head_f32_ = head_f32;
}

void Detector::layer_shape(int l, int* n, int* h, int* w, int* c, int* is_f32) const {
    YDST_CHECK(l >= 0 && l < (int)out_.size(), "layer index %d out of range", l);
    const Act& a = out_[l];
    if (n) *n = a.N; if (h) *h = a.H; if (w) *w = a.W; if (c) *c = a.C;
    if (is_f32) *is_f32 = head_f32_[l] != nullptr;
}
void Detector::layer_output(int l, void* dense_out, cudaStream_t st) const {
    YDST_CHECK(l >= 0 && l < (int)out_.size(), "layer index %d out of range", l);
    const Act& a = out_[l];
    if (head_f32_[l]) launch_unpack_f32(head_f32_[l], a.ctot, a.N, a.H, a.W, a.C, (float*)dense_out, st);
    else {
        YDST_CHECK(a.base != nullptr, "layer %d has no materialised output", l);
        launch_unpack(a, (__half*)dense_out, st);
    }
}

void Detector::forward_u8(const uint8_t* frame_dev, float* pred_out, cudaStream_t st) {
    launch_u8_to_f32(frame_dev, in_f32_, (long long)batch * H * W * 3, st);
    count_launch();
    run_plan(plan, st);
    if (pred_out) YDST_CUDA(cudaMemcpyAsync(pred_out, pred, (size_t)batch * rows * fields * sizeof(float), cudaMemcpyDeviceToDevice, st));
}
void Detector::forward_nchw(const void* x_dev, int is_half, float* pred_out, cudaStream_t st) {
    launch_nchw_to_nhwc(x_dev, is_half, in_f32_, batch, 3, H, W, st);
    count_launch();
    run_plan(plan, st);
    if (pred_out) YDST_CUDA(cudaMemcpyAsync(pred_out, pred, (size_t)batch * rows * fields * sizeof(float), cudaMemcpyDeviceToDevice, st));
}
void Detector::nms(float conf, float iou, float* dets_out, int* n_out, cudaStream_t st) {
    nms_.run(pred, rows, fields, conf, iou, st);
    if (dets_out) YDST_CUDA(cudaMemcpyAsync(dets_out, nms_.dets, sizeof(float) * 6 * nms_.max_det, cudaMemcpyDeviceToDevice, st));
    if (n_out) YDST_CUDA(cudaMemcpyAsync(n_out, nms_.counters + 1, sizeof(int), cudaMemcpyDeviceToDevice, st));
}

// ================================================================================================
// ReID net (deep_sort/deep/model.py:48-95)
// ================================================================================================
static const int kStage[4][3] = {{64, 64, 0}, {64, 128, 1}, {128, 256, 1}, {256, 512, 1}};   // cin, cout, downsample

Reid::Reid(const float* weights, size_t n_weights, int max_batch_) : max_batch(max_batch_) {
    YDST_CHECK(max_batch >= 1, "max_batch must be >= 1");
    size_t wp = 0;
    auto take = [&](size_t cnt) {
        YDST_CHECK(wp + cnt <= n_weights, "ReID weights too short");
        const float* p = weights + wp;
        wp += cnt;
        return p;
    };
    auto conv_bn = [&](int cout, int cin, int k, bool has_bias, bool first) {
        const float* w = take((size_t)cout * cin * k * k);
        const float* cb = has_bias ? take(cout) : nullptr;
        const float* g = take(cout); const float* b = take(cout); const float* m = take(cout); const float* v = take(cout);
        weights_.push_back(pack_conv(arena_, w, cout, cin, k, g, b, m, v, cb, first));
    };
    conv_bn(64, 3, 3, true, true);
    for (int s = 0; s < 4; ++s)
        for (int blk = 0; blk < 2; ++blk) {
            const int cin = blk == 0 ? kStage[s][0] : kStage[s][1], cout = kStage[s][1];
            conv_bn(cout, cin, 3, false, false);
            conv_bn(cout, cout, 3, false, false);
            if (blk == 0 && kStage[s][2]) conv_bn(cout, cin, 1, false, false);
        }
    YDST_CHECK(wp == n_weights, "ReID weights: %zu floats given, %zu consumed", n_weights, wp);
    in_f32_ = (float*)arena_.alloc((size_t)max_batch * 128 * 64 * 3 * sizeof(float));
    feat_ = (float*)arena_.alloc((size_t)max_batch * 512 * sizeof(float));
    err_flag = (int*)arena_.alloc(8 * sizeof(int));           // one empty-crop flag per frame of a micro-batch
    ws_ = make_conv_workspace(arena_);
    bufs_.push_back(make_act(arena_, max_batch, 128, 64, 64));          // 0: stem out
    int h = 64, w = 32;
    for (int s = 0; s < 4; ++s) {
        if (s > 0) { h /= 2; w /= 2; }
        for (int k = 0; k < 4; ++k) bufs_.push_back(make_act(arena_, max_batch, h, w, kStage[s][1]));   // A, B, T, D
    }
}

const Plan& Reid::plan_for(int m) {
    auto it = plans_.find(m);
    if (it != plans_.end()) return it->second;
    if (plans_.size() > 256) plans_.clear();
    Plan plan;
    auto view = [&](int i) { Act a = bufs_[i]; a.N = m; return a; };
    size_t wi = 0;
    Act x = view(1);                                                    // stage-1 buffer A
    if (conv_first_pool_supported(128, 64, weights_[wi]->cout, 1, ACT_RELU, weights_[wi]->w_hilo)) {
        // conv 3->64 + BN + ReLU + MaxPool2d(3,2,1) in one kernel: the 128x64x64 activation stays on chip
        Op op; op.kind = OP_CONV_FIRST_POOL; op.fsrc = in_f32_; op.out = x; op.w = weights_[wi++].get();
        op.i0 = 128; op.i1 = 64; op.i2 = 1; op.i3 = ACT_RELU;
        plan.ops.push_back(op);
        plan.flops += 2.0 * m * 128 * 64 * 64 * 27;
    } else {
        // The stem output (1.1 MB per crop) is written and read back once by the max-pool.  YDST_STEM_CHUNK=n does both in chunks of n
        // crops that share ONE scratch region smaller than the L2, so that the pool reads hit the L2 and the scratch lines are
        // overwritten there before they are evicted.  Measured (r2): SLOWER -- 1.60 / 1.69 / 1.61 ms per 408 crops at n = 48 / 24 / 96
        // against 1.56 ms unchunked: both kernels are latency-bound, not HBM-bound, and the extra launches cost more than the
        // traffic saves.  Off by default; kept as a knob.
        static const int chunk_env = getenv("YDST_STEM_CHUNK") ? atoi(getenv("YDST_STEM_CHUNK")) : 0;
        const int chunk = chunk_env > 0 ? chunk_env : m;
        const ConvWeights* w0 = weights_[wi++].get();
        for (int n0 = 0; n0 < m; n0 += chunk) {
            const int cnt = std::min(chunk, m - n0);
            Act scratch = bufs_[0]; scratch.N = cnt;                    // the same rows for every chunk
            Act dst = x; dst.N = cnt; dst.base = x.base + (long long)n0 * x.Hp() * x.Wp() * x.ctot;
            Op op; op.kind = OP_CONV_FIRST; op.fsrc = in_f32_ + (size_t)n0 * 128 * 64 * 3; op.out = scratch; op.w = w0;
            op.i0 = 128; op.i1 = 64; op.i2 = 1; op.i3 = ACT_RELU;
            plan.ops.push_back(op);
            Op pl; pl.kind = OP_MAXPOOL; pl.a = scratch; pl.out = dst; pl.i0 = 3; pl.i1 = 2; pl.i2 = 0;
            plan.ops.push_back(pl);
        }
        plan.flops += 2.0 * m * 128 * 64 * 64 * 27;
    }
    for (int s = 0; s < 4; ++s) {
        const int base = 1 + 4 * s;
        Act A = view(base), B = view(base + 1), T = view(base + 2), D = view(base + 3);
        for (int blk = 0; blk < 2; ++blk) {
            const bool down = blk == 0 && kStage[s][2];
            const ConvWeights* w1 = weights_[wi++].get();
            const ConvWeights* w2 = weights_[wi++].get();
            const ConvWeights* wd = down ? weights_[wi++].get() : nullptr;
            // the block's output buffer: the stage buffer x is not currently in
            Act y = (x.base == A.base) ? B : A;
            if (s > 0 && blk == 0) y = A;                                // x lives in the previous stage's buffers
            Op c1; c1.kind = OP_CONV_TC;
            conv_tc_plan(c1.conv, x, T, w1->w16, 3, 3, down ? 2 : 1, w1->scale, w1->bias, ACT_RELU, 0, nullptr, nullptr, w1->cout, &ws_);
            plan.ops.push_back(c1); plan.flops += conv_tc_flops(c1.conv);
            Act res = x;
            if (down) {
                Op cd; cd.kind = OP_CONV_TC;
                conv_tc_plan(cd.conv, x, D, wd->w16, 1, 1, 2, wd->scale, wd->bias, ACT_LINEAR, 0, nullptr, nullptr, wd->cout, &ws_);
                plan.ops.push_back(cd); plan.flops += conv_tc_flops(cd.conv);
                res = D;
            }
            Op c2; c2.kind = OP_CONV_TC;
            conv_tc_plan(c2.conv, T, y, w2->w16, 3, 3, 1, w2->scale, w2->bias, ACT_RELU, 2, &res, nullptr, w2->cout, &ws_);
            plan.ops.push_back(c2); plan.flops += conv_tc_flops(c2.conv);
            x = y;
        }
    }
    {
        Op op; op.kind = OP_AVGPOOL_L2; op.a = x; op.fdst = feat_;
        plan.ops.push_back(op);
    }
    link_weight_prefetch(plan);
    plan.launches = (int)plan.ops.size();
    return plans_.emplace(m, std::move(plan)).first->second;
}

void Reid::forward(const float* x_dev, int m, float* feat_out, cudaStream_t st) {
    if (m == 0) return;
    YDST_CHECK(m <= max_batch, "ReID batch %d exceeds max_batch %d", m, max_batch);
    if (x_dev != in_f32_)
        YDST_CUDA(cudaMemcpyAsync(in_f32_, x_dev, (size_t)m * 128 * 64 * 3 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    // plans (tensor maps, tilings, CUDA graph) are cached per crop count; a stream's count changes by a few from one micro-batch
    // to the next, so large counts are rounded up to a multiple of 16: the padding rows hold zeros or stale crops, every crop is
    // computed independently of its neighbours, and only the first m feature rows are handed out
    const int m_plan = m <= 32 ? m : std::min(max_batch, (m + 15) & ~15);
    run_plan(plan_for(m_plan), st);
    if (feat_out && feat_out != feat_)
        YDST_CUDA(cudaMemcpyAsync(feat_out, feat_, (size_t)m * 512 * sizeof(float), cudaMemcpyDeviceToDevice, st));
}

void Reid::extract_multi(const uint8_t* const* frames_dev, const int* H, const int* W, const float* const* tlwh_dev, const int* m, int nb,
                         float* feat_out, cudaStream_t st) {
    // The crops of all frames go through the net in chunks of at most max_batch (the activation buffers' capacity): the reference
    // has no limit on the number of detections (deep_sort/deep_sort.py:133-146), so neither a busy frame nor a micro-batch of B
    // frames may overflow -- they just take more than one forward.  Stream order makes the reuse of in_f32_ / feat_ safe.
    int done = 0, fill = 0;                                 // feature rows written so far | crops staged for the current chunk
    CropBatch cb{};                                         // crop segments (one per frame) queued for the next launch
    int cb_first = 0;                                       // chunk position of the first queued crop
    auto launch_crops = [&]() {
        if (cb.n == 0) return;
        launch_crop_resize_multi(cb, in_f32_ + (size_t)cb_first * 128 * 64 * 3, st);
        count_launch();
        cb.n = 0;
        cb_first = fill;
    };
    auto flush = [&]() {
        launch_crops();
        if (fill == 0) return;
        forward(in_f32_, fill, feat_out ? feat_out + (size_t)done * 512 : nullptr, st);
        done += fill;
        fill = 0;
        cb_first = 0;
    };
    for (int b = 0; b < nb; ++b) {
        int first = 0;
        while (first < m[b]) {
            const int take = std::min(m[b] - first, max_batch - fill);
            if (cb.n == 0) cb.start[0] = 0;
            cb.frame[cb.n] = frames_dev[b]; cb.tlwh[cb.n] = tlwh_dev[b] + (size_t)first * 4; cb.err[cb.n] = err_flag + (b & 7);
            cb.H[cb.n] = H[b]; cb.W[cb.n] = W[b];
            cb.start[cb.n + 1] = cb.start[cb.n] + take;
            ++cb.n;
            first += take;
            fill += take;
            if (cb.n == 8) launch_crops();
            if (fill == max_batch) flush();
        }
    }
    flush();
}

void Reid::extract(const uint8_t* frame_dev, int H, int W, const float* tlwh_dev, int m, float* feat_out, cudaStream_t st) {
    extract_multi(&frame_dev, &H, &W, &tlwh_dev, &m, 1, feat_out, st);
}

}  // namespace ydst
