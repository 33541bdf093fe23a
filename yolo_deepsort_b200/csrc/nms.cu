// Detection post-processing for sm_100a: soft_non_max_suppression (yolo3/utils/model_build.py:52-137; multi_label,
// class-offset trick; the options of the sliding-window mode -- is_p1p2, merge -- and classes / agnostic included) plus the
// box hand-off to the tracker (resize_boxes :12-19, p1p2Toxywh :326-332, class mask
// yolo3/detect/video_detect.py:138-147).  Integer/index results are bit-exact with the reference
// given identical fp32 predictions: same fp32 operations, same strict comparisons, stable
// score-descending order with ties broken by candidate index (torchvision nms semantics).
#include "nms.cuh"

namespace ydst {

// ---- 1. candidates: obj > thr, then every class with obj*cls > thr (row-major (row, class) key) ----
__global__ void nms_collect_kernel(const float* __restrict__ pred, int rows, int nf, float conf, NmsCand* __restrict__ cand,
                                   int cap, int* __restrict__ count, NmsOptions opt) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    const int img = blockIdx.y;
    cand += (long long)img * cap; count += img * 8;
    const float* p = pred + ((long long)img * rows + row) * nf;
    const float obj = p[4];
    if (!(obj > conf)) return;
    const float cx = p[0], cy = p[1], w = p[2], h = p[3];
    // is_p1p2: the four fields already are corners (model_build.py:90-93)
    const float x1 = opt.p1p2 ? cx : cx - w / 2.f, y1 = opt.p1p2 ? cy : cy - h / 2.f;
    const float x2 = opt.p1p2 ? w : cx + w / 2.f, y2 = opt.p1p2 ? h : cy + h / 2.f;
    const int nc = nf - 5;
    for (int j = 0; j < nc; ++j) {
        const float s = p[5 + j] * obj;
        // `classes`: keep only the listed class ids (model_build.py:103-105)
        if (opt.use_classes && !((opt.class_bits[(j >> 6) & 3] >> (j & 63)) & 1ull)) continue;
        if (s > conf) {
            const int slot = atomicAdd(count, 1);
            if (slot < cap) {
                NmsCand c;
                c.x1 = x1; c.y1 = y1; c.x2 = x2; c.y2 = y2; c.score = s; c.cls = (float)j;
                c.key = row * nc + j;
                c.pad = 0;
                cand[slot] = c;
            }
        }
    }
}

// ---- 2. stable descending order by counting: rank(i) = #{j : s_j > s_i or (s_j == s_i and key_j < key_i)} ----
__global__ void nms_rank_kernel(const NmsCand* __restrict__ cand, const int* __restrict__ count, int cap, NmsCand* __restrict__ sorted) {
    const int img = blockIdx.y;
    cand += (long long)img * cap; sorted += (long long)img * cap; count += img * 8;
    const int n = min(*count, cap);
    __shared__ float ss[256];
    __shared__ int sk[256];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (blockIdx.x * blockDim.x >= n) return;
    NmsCand me;
    float si = 0.f;
    int ki = 0;
    if (i < n) { me = cand[i]; si = me.score; ki = me.key; }
    int rank = 0;
    for (int base = 0; base < n; base += 256) {
        const int j = base + threadIdx.x;
        if (j < n) { ss[threadIdx.x] = cand[j].score; sk[threadIdx.x] = cand[j].key; }
        __syncthreads();
        const int lim = min(256, n - base);
        for (int t = 0; t < lim; ++t) rank += (ss[t] > si || (ss[t] == si && sk[t] < ki)) ? 1 : 0;
        __syncthreads();
    }
    if (i < n) sorted[rank] = me;
}

// ---- 3. suppression bitmask on class-offset boxes (boxes + cls*4096 in fp32, as the reference does) ----
__global__ void nms_mask_kernel(const NmsCand* __restrict__ sorted, const int* __restrict__ count, int cap, float iou_thr,
                                unsigned long long* __restrict__ mask, int words, float off_mul) {
    const int img = blockIdx.y;
    sorted += (long long)img * cap; count += img * 8; mask += (long long)img * cap * words;
    const int n = min(*count, cap);
    const int i = blockIdx.x;                 // row
    if (i >= n) return;
    const NmsCand a = sorted[i];
    const float off_a = a.cls * off_mul;                 // 4096, or 0 when agnostic (model_build.py:117)
    const float ax1 = a.x1 + off_a, ay1 = a.y1 + off_a, ax2 = a.x2 + off_a, ay2 = a.y2 + off_a;
    const float area_a = (ax2 - ax1) * (ay2 - ay1);
    const int nw = (n + 63) >> 6;              // words beyond the candidate count are never read by the sweep
    for (int w = threadIdx.x; w < nw; w += blockDim.x) {
        unsigned long long bits = 0;
        for (int b = 0; b < 64; ++b) {
            const int j = w * 64 + b;
            if (j <= i || j >= n) continue;
            const NmsCand c = sorted[j];
            const float off_c = c.cls * off_mul;
            const float bx1 = c.x1 + off_c, by1 = c.y1 + off_c, bx2 = c.x2 + off_c, by2 = c.y2 + off_c;
            const float area_b = (bx2 - bx1) * (by2 - by1);
            const float iw = fmaxf(0.f, fminf(ax2, bx2) - fmaxf(ax1, bx1));
            const float ih = fmaxf(0.f, fminf(ay2, by2) - fmaxf(ay1, by1));
            const float inter = iw * ih;
            const float ovr = inter / (area_a + area_b - inter);
            if (ovr > iou_thr) bits |= 1ull << b;
        }
        mask[(long long)i * words + w] = bits;
    }
}

// ---- 4. sequential sweep (warp 0), output kept rows in score order, capped at max_det ----
// The sweep is a dependent chain (candidate i is kept only if no earlier kept candidate suppressed it), so the mask rows it
// ORs in are first staged in shared memory by the whole block (n x ceil(n/64) words, up to kSweepSmemRows candidates);
// beyond that it reads them from global memory.
static constexpr int kSweepSmemRows = 1024;
__global__ void __launch_bounds__(256) nms_sweep_kernel(const NmsCand* __restrict__ sorted, const int* __restrict__ count, int cap,
                                                        const unsigned long long* __restrict__ mask, int words, int max_det,
                                                        float* __restrict__ dets, int* __restrict__ n_out, int* __restrict__ overflow) {
    extern __shared__ unsigned long long smask[];
    __shared__ int s_kept[1024];               // max_det <= 1024 (Nms::init)
    const int img = blockIdx.x;
    sorted += (long long)img * cap; count += img * 8; mask += (long long)img * cap * words;
    dets += (long long)img * max_det * 6; n_out += img * 8; overflow += img * 8;
    const int total = *count;
    const int n = min(total, cap);
    const int nw = (n + 63) >> 6;
    const bool staged = n <= kSweepSmemRows;
    if (staged) {
        for (int e = threadIdx.x; e < n * nw; e += blockDim.x) {
            const int i = e / nw, w = e - i * nw;
            smask[e] = mask[(long long)i * words + w];
        }
    }
    __syncthreads();
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    if (lane == 0) *overflow = total > cap ? 1 : 0;
    // removed bits live in registers: word w is held by lane w % 32, slot w / 32 (words <= 128)
    unsigned long long removed[4] = {0, 0, 0, 0};
    int kept = 0;
    for (int i = 0; i < n; ++i) {
        const int w = i >> 6;
        const unsigned long long mine = removed[w >> 5];
        const unsigned long long word = __shfl_sync(0xffffffffu, mine, w & 31);
        if ((word >> (i & 63)) & 1ull) continue;
        if (kept < max_det && lane == 0) s_kept[kept] = i;       // rows are gathered after the sweep, off the dependent chain
        ++kept;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const int ww = s * 32 + lane;
            if (ww < nw) removed[s] |= staged ? smask[i * nw + ww] : mask[(long long)i * words + ww];
        }
    }
    const int nk = min(kept, max_det);
    if (lane == 0) *n_out = nk;
    __syncwarp();
    for (int e = lane; e < nk * 6; e += 32) {
        const int k = e / 6, f = e - k * 6;
        const NmsCand& c = sorted[s_kept[k]];
        dets[e] = f == 0 ? c.x1 : f == 1 ? c.y1 : f == 2 ? c.x2 : f == 3 ? c.y2 : f == 4 ? c.score : c.cls;
    }
}

// ---- 4b. "Merge NMS" exactly as the reference executes it (model_build.py:122-131), quirks included ----
// The reference calls its ELEMENTWISE bbox_iou(boxes[i], boxes) (model_build.py:354-381: (n,4) x (n,4) -> (n,), "+1" areas,
// eps 1e-16) where ultralytics has the pairwise box_iou.  With k kept rows out of n candidates that line
//   * raises inside the bare try/except unless the shapes broadcast, i.e. unless k == n or k == 1: nothing is merged and the
//     `redundant` filter never runs (the exception is swallowed, model_build.py:129-131);
//   * k == n: pairs kept row j (score order) with candidate j (row-major candidate order), weights_j = (iou_j > thr) * score_j is
//     a (1,n) row, torch.mm gives ONE (1,4) weighted mean box, which the broadcast assignment x[i, :4] = ... writes into EVERY
//     kept row; then iou.sum(1) raises (1-D tensor) and is swallowed;
//   * k == 1: the kept box against all n candidates -- the one case that does what "merge" promises -- same assignment, same
//     swallowed exception.
// Only applies for 1 < n < 3000 (:122).  fp32 sums here, torch.mm there: equal to rounding (summation order).
__global__ void __launch_bounds__(256) nms_merge_kernel(const NmsCand* __restrict__ sorted, const int* __restrict__ count, int cap,
                                                        float iou_thr, float off_mul, int max_det, float* __restrict__ dets,
                                                        const int* __restrict__ n_out) {
    __shared__ int order[1024];                // candidate (key) order -> index into sorted[], k == n case (n <= max_det <= 1024)
    __shared__ float red[5][8];
    const int img = blockIdx.x;
    sorted += (long long)img * cap; count += img * 8; dets += (long long)img * max_det * 6; n_out += img * 8;
    const int total = *count;
    const int n = min(total, cap), k = *n_out;
    if (!(total > 1 && total < 3000) || total > cap) return;
    if (k != n && k != 1) return;
    const bool all = k == n && k != 1;
    if (all) {
        for (int t = threadIdx.x; t < n; t += blockDim.x) {
            int rank = 0;
            const int key = sorted[t].key;
            for (int u = 0; u < n; ++u) rank += sorted[u].key < key ? 1 : 0;
            order[rank] = t;
        }
    }
    __syncthreads();
    float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const NmsCand a = all ? sorted[j] : sorted[0];            // kept row j | the single kept row
        const NmsCand b = all ? sorted[order[j]] : sorted[j];     // candidate j
        const float oa = a.cls * off_mul, ob = b.cls * off_mul;
        const float ax1 = a.x1 + oa, ay1 = a.y1 + oa, ax2 = a.x2 + oa, ay2 = a.y2 + oa;
        const float bx1 = b.x1 + ob, by1 = b.y1 + ob, bx2 = b.x2 + ob, by2 = b.y2 + ob;
        const float iw = fmaxf(fminf(ax2, bx2) - fmaxf(ax1, bx1) + 1.f, 0.f);
        const float ih = fmaxf(fminf(ay2, by2) - fmaxf(ay1, by1) + 1.f, 0.f);
        const float inter = iw * ih;
        const float area_a = (ax2 - ax1 + 1.f) * (ay2 - ay1 + 1.f), area_b = (bx2 - bx1 + 1.f) * (by2 - by1 + 1.f);
        const float iou = inter / (area_a + area_b - inter + 1e-16f);
        const float w = iou > iou_thr ? b.score : 0.f;
        acc[0] += w; acc[1] += w * b.x1; acc[2] += w * b.y1; acc[3] += w * b.x2; acc[4] += w * b.y2;
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int f = 0; f < 5; ++f) {
        float v = acc[f];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[f][wid] = v;
    }
    __syncthreads();
    float tot[5];
#pragma unroll
    for (int f = 0; f < 5; ++f) {
        float v = 0.f;
        for (int w = 0; w < 8; ++w) v += red[f][w];
        tot[f] = v;
    }
    for (int e = threadIdx.x; e < k * 4; e += blockDim.x) dets[(e >> 2) * 6 + (e & 3)] = tot[1 + (e & 3)] / tot[0];   // 0/0 = NaN, as torch
}

// ---- 5. hand-off: resize_boxes, xyxy -> tlwh, class mask; order preserved (one block, ballot scan) ----
__global__ void __launch_bounds__(1024) dets_to_tracks_kernel(const float* __restrict__ dets, const int* __restrict__ n_dets, NmsRatios ratios,
                                                              int max_det, const int* __restrict__ class_mask, int n_mask,
                                                              float* __restrict__ tlwh, float* __restrict__ conf,
                                                              float* __restrict__ cls, int* __restrict__ m_out) {
    __shared__ int warp_cnt[32];
    const int img = blockIdx.x;
    const float rw = ratios.rw[img], rh = ratios.rh[img];
    dets += (long long)img * max_det * 6; n_dets += img * 8; m_out += img * 8;
    tlwh += (long long)img * max_det * 4; conf += (long long)img * max_det; cls += (long long)img * max_det;
    const int n = *n_dets;
    const int i = threadIdx.x, lane = i & 31, wid = i >> 5;
    bool keep = false;
    float x1 = 0, y1 = 0, x2 = 0, y2 = 0, sc = 0, cl = 0;
    if (i < n) {
        x1 = dets[i * 6 + 0] * rw; y1 = dets[i * 6 + 1] * rh; x2 = dets[i * 6 + 2] * rw; y2 = dets[i * 6 + 3] * rh;
        sc = dets[i * 6 + 4]; cl = dets[i * 6 + 5];
        keep = n_mask == 0;
        for (int k = 0; k < n_mask; ++k) keep |= (cl == (float)class_mask[k]);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_cnt[wid] = __popc(bal);
    __syncthreads();
    int base = 0, total = 0;
    for (int w = 0; w < 32; ++w) { if (w < wid) base += warp_cnt[w]; total += warp_cnt[w]; }
    if (keep) {
        const int pos = base + __popc(bal & ((1u << lane) - 1u));
        tlwh[pos * 4 + 0] = x1; tlwh[pos * 4 + 1] = y1; tlwh[pos * 4 + 2] = x2 - x1; tlwh[pos * 4 + 3] = y2 - y1;
        conf[pos] = sc; cls[pos] = cl;
    }
    if (i == 0) *m_out = total;
}

void Nms::init(int cap_, int max_det_, int batch_) {
    cap = cap_; max_det = max_det_; batch = batch_;
    YDST_CHECK(batch >= 1 && batch <= 8, "nms batch 1..8");
    YDST_CHECK(cap % 64 == 0 && cap <= 8192, "nms candidate capacity must be a multiple of 64, <= 8192");
    YDST_CHECK(max_det <= 1024, "max_det <= 1024");
    words = cap / 64;
    YDST_CUDA(cudaMalloc(&cand, sizeof(NmsCand) * cap * batch));
    YDST_CUDA(cudaMalloc(&sorted, sizeof(NmsCand) * cap * batch));
    YDST_CUDA(cudaMalloc(&mask, sizeof(unsigned long long) * (size_t)cap * words * batch));
    YDST_CUDA(cudaMalloc(&counters, sizeof(int) * 8 * batch));
    YDST_CUDA(cudaMalloc(&dets, sizeof(float) * 6 * max_det * batch));
}
void Nms::destroy() {
    cudaFree(cand); cudaFree(sorted); cudaFree(mask); cudaFree(counters); cudaFree(dets);
    cand = sorted = nullptr; mask = nullptr; counters = nullptr; dets = nullptr;
}
// counters (per image, stride 8): [0] candidate count, [1] n_out, [2] overflow flag, [3] m (tracker inputs)
void Nms::run(const float* pred, int rows, int nf, float conf, float iou, cudaStream_t st, int nb, const NmsOptions* options) {
    YDST_CHECK(nb >= 1 && nb <= batch, "nms over %d images, capacity %d", nb, batch);
    NmsOptions opt{};
    if (options) opt = *options;
    const float off_mul = opt.agnostic ? 0.f : 4096.f;
    YDST_CUDA(cudaMemsetAsync(counters, 0, sizeof(int) * 8 * nb, st));
    nms_collect_kernel<<<dim3((rows + 255) / 256, nb), 256, 0, st>>>(pred, rows, nf, conf, cand, cap, counters, opt);
    nms_rank_kernel<<<dim3(cap / 256 > 0 ? cap / 256 : 1, nb), 256, 0, st>>>(cand, counters, cap, sorted);
    nms_mask_kernel<<<dim3(cap, nb), 64, 0, st>>>(sorted, counters, cap, iou, mask, words, off_mul);
    // the dynamic shared memory opt-in is a per-device attribute of the kernel
    static bool attr_set[64] = {false};
    int dev = 0;
    YDST_CUDA(cudaGetDevice(&dev));
    const int sweep_smem = kSweepSmemRows * (kSweepSmemRows / 64) * (int)sizeof(unsigned long long);
    if (!attr_set[dev & 63]) {
        YDST_CUDA(cudaFuncSetAttribute(nms_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sweep_smem));
        attr_set[dev & 63] = true;
    }
    nms_sweep_kernel<<<nb, 256, sweep_smem, st>>>(sorted, counters, cap, mask, words, max_det, dets, counters + 1, counters + 2);
    int launches = 4;
    if (opt.merge) {
        nms_merge_kernel<<<nb, 256, 0, st>>>(sorted, counters, cap, iou, off_mul, max_det, dets, counters + 1);
        ++launches;
    }
    YDST_CUDA(cudaGetLastError());
    count_launch(launches);
}
void Nms::to_tracker_inputs_batch(const NmsRatios& r, int nb, const int* class_mask_dev, int n_mask, float* tlwh, float* conf, float* cls,
                                  cudaStream_t st) {
    dets_to_tracks_kernel<<<nb, 1024, 0, st>>>(dets, counters + 1, r, max_det, class_mask_dev, n_mask, tlwh, conf, cls, counters + 3);
    YDST_CUDA(cudaGetLastError());
    count_launch();
}
void Nms::to_tracker_inputs(float rw, float rh, const int* class_mask_dev, int n_mask, float* tlwh, float* conf, float* cls, cudaStream_t st) {
    NmsRatios r{};
    r.rw[0] = rw; r.rh[0] = rh;
    to_tracker_inputs_batch(r, 1, class_mask_dev, n_mask, tlwh, conf, cls, st);
}

}  // namespace ydst
