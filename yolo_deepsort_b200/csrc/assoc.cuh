// DeepSORT association kernels (declarations).  See assoc.cu.
#pragma once
#include "common.cuh"

namespace ydst {

static constexpr float kInftyCost = 1e+5f;        // deep_sort/sort/linear_assignment.py:3
static constexpr float kChi2inv95_2 = 5.9915f;    // deep_sort/sort/kalman_filter.py:10 (2 dof)
static constexpr int kFeat = 512;

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

// Appearance cost.  gallery: normalised rows [G][512] addressed through row_ptr[g] (row index into `gallery`),
// row_track[g] = row of the cost matrix that gallery row g belongs to.  det_feat_n: normalised [m][512].
// cost_enc [n][m] must be pre-filled by launch_fill_inf; finalize applies 1 - max cosine -> gate -> clamp.
void launch_fill_i32(int* p, int v, long long n, cudaStream_t st);
void launch_cosine_min(const float* gallery, const int* row_ptr, const int* row_track, int G, const float* det_feat_n, int m,
                       int* cost_enc, cudaStream_t st);
void launch_cost_finalize(const int* cost_enc, const float* mean, const float* cov, const int* idx, int n, const float* det_tlwh,
                          int m, double max_dist, float* cost, cudaStream_t st);
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
