// DeepSORT association kernels (declarations).  See assoc.cu.
#pragma once
#include "common.cuh"

namespace ydst {

static constexpr float kInftyCost = 1e+5f;        // deep_sort/sort/linear_assignment.py:3
static constexpr float kChi2inv95_2 = 5.9915f;    // deep_sort/sort/kalman_filter.py:10 (2 dof)
static constexpr int kFeat = 512;

// squared Mahalanobis distance in (x, y) only: project, 2x2 LU inverse (getrf + getri order), d S^-1 d^T
__device__ __forceinline__ float maha_position(const float* __restrict__ mean, const float* __restrict__ cov, float zx, float zy) {
    const float h = mean[3];
    const float sp = h * 0.05f;
    float a = cov[0] + sp * sp, b = cov[1], c = cov[8], d = cov[9] + sp * sp;
    const bool swap = fabsf(c) > fabsf(a);
    if (swap) { float t = a; a = c; c = t; t = b; b = d; d = t; }
    const float l = c * (1.f / a);
    const float u22 = d - l * b;
    const float iu00 = 1.f / a, iu11 = 1.f / u22;
    const float iu01 = (iu00 * b) * (-iu11);
    float i00 = iu00 - iu01 * l, i10 = 0.f - iu11 * l, i01 = iu01, i11 = iu11;
    if (swap) { float t = i00; i00 = i01; i01 = t; t = i10; i10 = i11; i11 = t; }
    const float d0 = -mean[0] + zx, d1 = -mean[1] + zy;
    const float t0 = fmaf(d1, i10, d0 * i00), t1 = fmaf(d1, i11, d0 * i01);
    return fmaf(t1, d1, t0 * d0);
}


// All Kalman entry points work on struct-of-arrays track state: mean [slots][8], cov [slots][64] (row-major 8x8).
// `idx` (device, may be null = identity) selects the slots of the n tracks being processed.
void launch_kf_initiate(const float* det_tlwh, const int* det_idx, float* mean, float* cov, const int* slot_idx, int n, cudaStream_t st);
void launch_kf_predict(float* mean, float* cov, const int* idx, int n, cudaStream_t st);
void launch_kf_update(float* mean, float* cov, const int* idx, const float* det_tlwh, const int* det_idx, int n, cudaStream_t st);
// squared Mahalanobis distance, position only: maha [n][m]
void launch_gate_position(const float* mean, const float* cov, const int* idx, int n, const float* det_tlwh, int m, float* maha,
                          cudaStream_t st);

// f / ||f||_2 for each of n rows of 512 (nn_matching.py:50-52); src and dst may alias
void launch_normalize_rows(const float* src, float* dst, int n, cudaStream_t st);

// (the appearance cost -- cosine GEMM, segmented max, gate, clamp -- lives in cosine_tc.cu)
// IoU cost (iou_matching.py:5-91) between tracks idx[0..n) and detections det_idx[0..m), clamped at max_dist (+1e-5)
void launch_iou_cost(const float* mean, const int* idx, const int* tsu, int n, const float* det_tlwh, const int* det_idx, int m,
                     double max_dist, float* cost, cudaStream_t st);
void launch_transpose(const float* src, float* dst, int rows, int cols, cudaStream_t st);

// Exact rectangular LSAP with scipy's tie-breaking.  cost: [R][C] row-major with R <= C.
// Outputs col4row [R] and over_max [R] (1 if cost[r][col4row[r]] > max_dist, linear_assignment.py:68).
// work: device scratch of lsap_work_bytes(R, C).
size_t lsap_work_bytes(int R, int C);
void launch_lsap(const float* cost, int R, int C, float max_dist, int* col4row, int* over_max, void* work, cudaStream_t st);

}  // namespace ydst
