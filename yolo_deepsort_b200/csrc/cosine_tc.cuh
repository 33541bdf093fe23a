// Appearance cost on the tensor cores (declarations).  See cosine_tc.cu.
#pragma once
#include <map>

#include "assoc.cuh"
#include "conv_tc.cuh"

namespace ydst {

// cost[r][j] = clamp(gate(1 - max_{g in gallery of track r} <gallery_g, det_j>)) for n tracks x m detections
// (nn_matching.py:30-100,158-187 + linear_assignment.py:52,201-202), with the G x m x 512 product as ONE tcgen05 GEMM.
class CosineTc {
public:
    CosineTc() = default;
    ~CosineTc();
    CosineTc(const CosineTc&) = delete;
    CosineTc& operator=(const CosineTc&) = delete;
    // gallery: unit rows [..][512] fp32 addressed through row_ptr[g] (device, G entries); seg (device, n + 1 entries): the rows of
    // cost-matrix row r are g in [seg[r], seg[r+1]); det_n: unit rows [m][512] fp32; mean/cov/idx as in launch_cost_finalize.
    void run(const float* gallery, const int* row_ptr, const int* seg, int G, int n, const float* det_n, int m, const float* mean,
             const float* cov, const int* idx, const float* det_tlwh, double max_dist, float* cost, cudaStream_t st);
    int launches_last = 0;
    // event hooks of the caller's profiler (kind, algorithmic bytes, flops): called around the three phases when set
    void (*prof_begin)(int, double, double, cudaStream_t) = nullptr;
    void (*prof_end)(cudaStream_t) = nullptr;

private:
    void reserve(int g_pad, int m_pad);
    __half* a_ = nullptr;       // [g_cap][1536]  gallery rows split hi | lo | hi
    __half* b_ = nullptr;       // [m_cap][1536]  detection rows split hi | hi | lo
    float* dots_ = nullptr;     // [g_cap][m_cap]
    float *one_ = nullptr, *zero_ = nullptr;
    int g_cap_ = 0, m_cap_ = 0;
    ConvWorkspace ws_{};
    std::map<long long, ConvTcLaunch> plans_;
};

}  // namespace ydst
