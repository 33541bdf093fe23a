// Device-resident DeepSORT tracker (declarations).  See tracker.cu.
#pragma once
#include <vector>

#include "assoc.cuh"
#include "cosine_tc.cuh"

namespace ydst {

enum TrackState { TRACK_TENTATIVE = 1, TRACK_CONFIRMED = 2, TRACK_DELETED = 3 };   // deep_sort/sort/track.py:11-13

struct TrackHost {
    int id, hits, age, tsu, state, payload;
    int slot;        // row in the device struct-of-arrays state
    int gal_count;   // features stored in this track's gallery ring (<= budget)
    int gal_head;    // next ring position to overwrite
};

class Tracker {
public:
    Tracker(double max_dist, double max_iou, int max_age, int n_init, int budget, int cap_tracks, int cap_dets);
    ~Tracker();
    // payload_host may be null when cls_dev (float class ids on the device) is given
    void update(const float* tlwh_dev, const float* feat_dev, const int* payload_host, const float* cls_dev, int m, int32_t* out_host,
                int* k_host, cudaStream_t st);
    void snapshot(int32_t* table_host, float* mean_host, int cap, int* n_host, cudaStream_t st);
    std::vector<TrackHost> tracks;
    std::vector<std::pair<int, int>> last_matches;
    int next_id = 1;
    int launches_last = 0;

private:
    struct Assign { std::vector<std::pair<int, int>> matches; std::vector<int> um_t, um_d; };
    // solves the LSAP for the device cost matrix [nt][nd] and applies linear_assignment.py:58-72
    Assign solve(const float* cost, const std::vector<int>& track_indices, const std::vector<int>& det_indices, float max_dist, cudaStream_t st);
    int* upload(const std::vector<int>& v, cudaStream_t st);
    double max_dist_, max_iou_;    // kept in double: the clamp value max_distance + 1e-5 is formed in double by the reference
    int max_age_, n_init_, budget_, cap_t_, cap_d_;
    // device state
    float *mean_ = nullptr, *cov_ = nullptr, *gallery_ = nullptr, *det_n_ = nullptr;
    int* tsu_dev_ = nullptr;
    // device scratch
    float *cost_ = nullptr, *cost_t_ = nullptr, *out_mean_ = nullptr;
    int *col4row_ = nullptr, *over_ = nullptr, *ibuf_ = nullptr;
    CosineTc cos_;                  // appearance cost on the tensor cores (owns its operand / product scratch)
    void* lsap_work_ = nullptr;
    size_t ibuf_cap_ = 0, ibuf_used_ = 0;
    // pinned host staging
    int* h_ibuf_ = nullptr;
    int* h_res_ = nullptr;
    float* h_f_ = nullptr;
    std::vector<int> free_slots_;
};

}  // namespace ydst
