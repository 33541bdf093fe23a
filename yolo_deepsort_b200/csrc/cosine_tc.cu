// Appearance cost on the tensor cores.
//
// The reference stacks every gallery row of every confirmed track (G x 512), normalises both sides and forms 1 - A B^T
// (deep_sort/sort/nn_matching.py:30-53,77-100,158-187): a G x m x 512 fp32 product -- 122.9 GFLOP at the stress size
// (60 000 gallery rows x 2000 detections), which CUDA-core FFMA needs >= 1.7 ms for even at its 72 TFLOP/s peak.  The costs are
// compared with 0.3 and fed to an exact assignment, so reduced precision is not an option; but fp32-grade products do not need
// fp32 multipliers: with x = hi + lo, hi = fp16(x), lo = fp16(x - hi) (|x| <= 1: unit rows), x y = hi hi' + lo hi' + hi lo' up to
// 2^-24 |x||y|, every partial product is exact in the fp32 accumulator, and the three terms are ONE GEMM over K = 3 * 512:
//     A' = [ hi | lo | hi ]   (G x 1536 fp16)        B' = [ hi' | hi' | lo' ]   (m x 1536 fp16)        A' B'^T = sum of the three.
// That GEMM runs on the tcgen05 kernel the convolutions use (conv_tc.cu, 1x1 mode with the plain-matrix flag): TMA-fed 128B-swizzled
// operand tiles, fp32 accumulators in TMEM.  Three kernels per call instead of five:
//   1. feat_split_kernel     gather + split of the gallery rows and of the detections          (HBM: 2 KB read, 3 KB written per row)
//   2. conv_tc2_kernel       dots[G][m]                                                          (tensor cores)
//   3. cost_segmin_kernel    per (track, detection): 1 - max over the track's rows, chi^2 gate, clamp   (HBM: reads dots once)
#include "cosine_tc.cuh"

#include <math_constants.h>

#include <algorithm>
#include <vector>

namespace ydst {

static inline int cdiv_ll(long long a, int b) { return (int)((a + b - 1) / b); }
static constexpr int kK3 = 3 * kFeat;

// one warp per row: x -> (hi, lo); dst row = [hi | lo | hi] (gallery side, order 0) or [hi | hi | lo] (detection side, order 1).
// Rows >= n_valid are written as zeros (padding up to the tile size).
__global__ void __launch_bounds__(256) feat_split_kernel(const float* __restrict__ src, const int* __restrict__ row_ptr, int n_valid, int n_rows,
                                                         int order, __half* __restrict__ dst) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= n_rows) return;
    __half2* d = reinterpret_cast<__half2*>(dst + (long long)row * kK3);
    if (row >= n_valid) {
        for (int i = lane; i < kK3 / 2; i += 32) d[i] = __floats2half2_rn(0.f, 0.f);
        return;
    }
    const float4* s = reinterpret_cast<const float4*>(src + (long long)(row_ptr ? row_ptr[row] : row) * kFeat);
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int q = it * 32 + lane;                       // float4 index 0..127
        const float4 v = __ldg(s + q);
        const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
        const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
        const __half2 l0 = __floats2half2_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2half2_rn(v.z - f1.x, v.w - f1.y);
        const int e = q * 2;                                // half2 index inside a 512-wide block
        d[e] = h0; d[e + 1] = h1;
        if (order == 0) { d[256 + e] = l0; d[256 + e + 1] = l1; d[512 + e] = h0; d[512 + e + 1] = h1; }
        else            { d[256 + e] = h0; d[256 + e + 1] = h1; d[512 + e] = l0; d[512 + e + 1] = l1; }
    }
}

// thread per (track r, detection j): running max of the dot products of the track's gallery rows (coalesced over j), then
// d = 1 - max (== min of 1 - dot: the subtraction is monotone and exact in the same direction), chi^2 gate (strict >), clamp
__global__ void __launch_bounds__(256) cost_segmin_kernel(const float* __restrict__ dots, int ld, const int* __restrict__ seg, int n, int m,
                                                          const float* __restrict__ mean, const float* __restrict__ cov,
                                                          const int* __restrict__ idx, const float* __restrict__ det_tlwh, float max_dist,
                                                          float clamp_val, float* __restrict__ cost) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (j >= m || r >= n) return;
    const int g0 = seg[r], g1 = seg[r + 1];
    float best = -CUDART_INF_F;
    for (int g = g0; g < g1; ++g) best = fmaxf(best, __ldg(dots + (long long)g * ld + j));
    float c = g1 > g0 ? 1.f - best : CUDART_INF_F;
    const int slot = idx ? idx[r] : r;
    const float* t = det_tlwh + j * 4;
    const float zx = t[0] + t[2] / 2.f, zy = t[1] + t[3] / 2.f;
    const float gd = maha_position(mean + (long long)slot * 8, cov + (long long)slot * 64, zx, zy);
    if (gd > kChi2inv95_2) c = kInftyCost;
    if (c > max_dist) c = clamp_val;
    cost[(long long)r * m + j] = c;
}

CosineTc::~CosineTc() {
    cudaFree(a_); cudaFree(b_); cudaFree(dots_); cudaFree(one_); cudaFree(zero_); cudaFree(ws_.partial); cudaFree(ws_.tickets);
}

void CosineTc::reserve(int g_pad, int m_pad) {
    if (g_pad <= g_cap_ && m_pad <= m_cap_) return;
    // grow geometrically; every plan holds tensor maps over the old buffers
    YDST_CUDA(cudaDeviceSynchronize());
    plans_.clear();
    const int g_new = std::max(g_pad, std::max(2048, g_cap_ + g_cap_ / 2)), m_new = std::max(m_pad, std::max(64, m_cap_));
    cudaFree(a_); cudaFree(b_); cudaFree(dots_);
    a_ = nullptr; b_ = nullptr; dots_ = nullptr;
    YDST_CUDA(cudaMalloc(&a_, (size_t)g_new * kK3 * sizeof(__half)));
    YDST_CUDA(cudaMalloc(&b_, (size_t)m_new * kK3 * sizeof(__half)));
    YDST_CUDA(cudaMalloc(&dots_, (size_t)g_new * m_new * sizeof(float)));
    g_cap_ = g_new; m_cap_ = m_new;
    if (!one_) {
        std::vector<float> ones(8192 + 256, 1.f);
        YDST_CUDA(cudaMalloc(&one_, ones.size() * sizeof(float)));
        YDST_CUDA(cudaMalloc(&zero_, ones.size() * sizeof(float)));
        YDST_CUDA(cudaMemcpy(one_, ones.data(), ones.size() * sizeof(float), cudaMemcpyHostToDevice));
        YDST_CUDA(cudaMemset(zero_, 0, ones.size() * sizeof(float)));
        ws_.partial_bytes = (size_t)48 << 20; ws_.n_tickets = 8192;
        YDST_CUDA(cudaMalloc(&ws_.partial, ws_.partial_bytes));
        YDST_CUDA(cudaMalloc(&ws_.tickets, ws_.n_tickets * sizeof(int)));
        YDST_CUDA(cudaMemset(ws_.tickets, 0, ws_.n_tickets * sizeof(int)));
    }
}

void CosineTc::run(const float* gallery, const int* row_ptr, const int* seg, int G, int n, const float* det_n, int m, const float* mean,
                   const float* cov, const int* idx, const float* det_tlwh, double max_dist, float* cost, cudaStream_t st) {
    launches_last = 0;
    if (n == 0 || m == 0) return;
    YDST_CHECK(m <= 8192, "appearance cost: %d detections exceed the GEMM's 8192 columns", m);
    const int g_pad = std::max(128, (G + 127) & ~127), m_pad = (m + 15) & ~15;
    reserve(g_pad, m_pad);
    // 1. operands
    if (prof_begin) prof_begin(114, 2048.0 * (G + m) + 3072.0 * (g_pad + m_pad), 0, st);
    feat_split_kernel<<<cdiv_ll(g_pad, 8), 256, 0, st>>>(gallery, row_ptr, G, g_pad, 0, a_);
    feat_split_kernel<<<cdiv_ll(m_pad, 8), 256, 0, st>>>(det_n, nullptr, m, m_pad, 1, b_);
    YDST_CUDA(cudaGetLastError());
    if (prof_end) prof_end(st);
    // 2. dots = A' B'^T on the tcgen05 GEMM (plans are cached per padded shape: tensor maps, tiling)
    const long long key = (long long)g_pad * 16384 + m_pad;
    auto it = plans_.find(key);
    if (it == plans_.end()) {
        if (plans_.size() > 64) plans_.clear();
        ConvTcLaunch L;
        conv_tc_plan_gemm(L, a_, g_pad, kK3, b_, m_pad, dots_, one_, zero_, &ws_);
        it = plans_.emplace(key, L).first;
    }
    if (prof_begin) prof_begin(112, 2.0 * kK3 * ((double)g_pad + m_pad) + 4.0 * g_pad * m_pad, 2.0 * g_pad * (double)m_pad * kK3, st);
    conv_tc_run(it->second, st);
    if (prof_end) prof_end(st);
    // 3. segmented max, gate, clamp
    const float clamp_val = (float)(max_dist + 1e-5);       // formed in double by the reference (linear_assignment.py:52)
    if (prof_begin) prof_begin(113, 4.0 * G * m + 4.0 * n * m + 288.0 * n + 16.0 * m, 0, st);
    cost_segmin_kernel<<<dim3(cdiv_ll(m, 256), n), 256, 0, st>>>(dots_, m_pad, seg, n, m, mean, cov, idx, det_tlwh, (float)max_dist, clamp_val, cost);
    YDST_CUDA(cudaGetLastError());
    if (prof_end) prof_end(st);
    launches_last = 4;
}

}  // namespace ydst
