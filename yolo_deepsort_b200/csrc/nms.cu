// Detection post-processing for sm_100a: the single-image video path of soft_non_max_suppression
// (yolo3/utils/model_build.py:52-137; merge=False, multi_label, class-offset trick) plus the
// box hand-off to the tracker (resize_boxes :12-19, p1p2Toxywh :326-332, class mask
// yolo3/detect/video_detect.py:138-147).  Integer/index results are bit-exact with the reference
// given identical fp32 predictions: same fp32 operations, same strict comparisons, stable
// score-descending order with ties broken by candidate index (torchvision nms semantics).
#include "nms.cuh"

namespace ydst {

// ---- 1. candidates: obj > thr, then every class with obj*cls > thr (row-major (row, class) key) ----
__global__ void nms_collect_kernel(const float* __restrict__ pred, int rows, int nf, float conf, NmsCand* __restrict__ cand,
                                   int cap, int* __restrict__ count) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    const int img = blockIdx.y;
    cand += (long long)img * cap; count += img * 8;
    const float* p = pred + ((long long)img * rows + row) * nf;
    const float obj = p[4];
    if (!(obj > conf)) return;
    const float cx = p[0], cy = p[1], w = p[2], h = p[3];
    const float x1 = cx - w / 2.f, y1 = cy - h / 2.f, x2 = cx + w / 2.f, y2 = cy + h / 2.f;
    const int nc = nf - 5;
    for (int j = 0; j < nc; ++j) {
        const float s = p[5 + j] * obj;
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
                                unsigned long long* __restrict__ mask, int words) {
    const int img = blockIdx.y;
    sorted += (long long)img * cap; count += img * 8; mask += (long long)img * cap * words;
    const int n = min(*count, cap);
    const int i = blockIdx.x;                 // row
    if (i >= n) return;
    const NmsCand a = sorted[i];
    const float off_a = a.cls * 4096.f;
    const float ax1 = a.x1 + off_a, ay1 = a.y1 + off_a, ax2 = a.x2 + off_a, ay2 = a.y2 + off_a;
    const float area_a = (ax2 - ax1) * (ay2 - ay1);
    const int nw = (n + 63) >> 6;              // words beyond the candidate count are never read by the sweep
    for (int w = threadIdx.x; w < nw; w += blockDim.x) {
        unsigned long long bits = 0;
        for (int b = 0; b < 64; ++b) {
            const int j = w * 64 + b;
            if (j <= i || j >= n) continue;
            const NmsCand c = sorted[j];
            const float off_c = c.cls * 4096.f;
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
void Nms::run(const float* pred, int rows, int nf, float conf, float iou, cudaStream_t st, int nb) {
    YDST_CHECK(nb >= 1 && nb <= batch, "nms over %d images, capacity %d", nb, batch);
    YDST_CUDA(cudaMemsetAsync(counters, 0, sizeof(int) * 8 * nb, st));
    nms_collect_kernel<<<dim3((rows + 255) / 256, nb), 256, 0, st>>>(pred, rows, nf, conf, cand, cap, counters);
    nms_rank_kernel<<<dim3(cap / 256 > 0 ? cap / 256 : 1, nb), 256, 0, st>>>(cand, counters, cap, sorted);
    nms_mask_kernel<<<dim3(cap, nb), 64, 0, st>>>(sorted, counters, cap, iou, mask, words);
    static bool attr_set = false;
    const int sweep_smem = kSweepSmemRows * (kSweepSmemRows / 64) * (int)sizeof(unsigned long long);
    if (!attr_set) {
        YDST_CUDA(cudaFuncSetAttribute(nms_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sweep_smem));
        attr_set = true;
    }
    nms_sweep_kernel<<<nb, 256, sweep_smem, st>>>(sorted, counters, cap, mask, words, max_det, dets, counters + 1, counters + 2);
    YDST_CUDA(cudaGetLastError());
    count_launch(4);
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
