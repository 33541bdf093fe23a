// Device-resident DeepSORT tracker.
//
// Replaces Tracker/Track/NearestNeighborDistanceMetric (deep_sort/sort/tracker.py:38-176, track.py:63-152,
// nn_matching.py:139-187) and the output block of DeepSort.update (deep_sort/deep_sort.py:63-88).
//
// Data layout: every track owns a slot in struct-of-arrays device state -- mean[slot][8], cov[slot][64]
// and a ring of `budget` L2-normalised 512-d gallery rows -- so Kalman predict/update, both cost matrices
// and the assignment run as kernels over index lists; there are no per-track host objects holding tensors.
// The integer lifecycle (which list a track is in, ids, hit counters, the ORDER of unmatched detections that
// decides new ids) is a few hundred integer operations per frame and stays on the host, fed by the
// assignment result (two small D2H copies per frame).
#include "tracker.cuh"
#include "net.cuh"

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>

namespace ydst {

// copy `n` feature rows: dst_rows[i] (row index into gallery) <- det_n[src_rows[i]]
__global__ void __launch_bounds__(128) gallery_append_kernel(const float* __restrict__ det_n, const int* __restrict__ src_rows,
                                                            float* __restrict__ gallery, const int* __restrict__ dst_rows, int n) {
    const int i = blockIdx.x;
    if (i >= n) return;
    const float4 v = reinterpret_cast<const float4*>(det_n + (long long)src_rows[i] * kFeat)[threadIdx.x];
    reinterpret_cast<float4*>(gallery + (long long)dst_rows[i] * kFeat)[threadIdx.x] = v;
}
__global__ void gather_mean_kernel(const float* __restrict__ mean, const int* __restrict__ idx, int n, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * 8) return;
    out[i] = mean[(long long)idx[i >> 3] * 8 + (i & 7)];
}
__global__ void cls_to_int_kernel(const float* __restrict__ cls, int m, int* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) out[i] = (int)cls[i];
}

// per-kernel CUDA-event sample while ydst_profile_begin() is active (bench.py --config assoc: roofline objects per kernel)
struct TrkProf {
    OpSample s{};
    cudaStream_t st;
    bool on;
    TrkProf(int kind, double bytes, double flops, cudaStream_t st_) : st(st_), on(profile_active()) {
        if (!on) return;
        s.kind = kind; s.layer = -1; s.flops = flops; s.bytes = bytes;
        cudaEventCreate(&s.e0); cudaEventCreate(&s.e1);
        cudaEventRecord(s.e0, st);
    }
    ~TrkProf() {
        if (!on) return;
        cudaEventRecord(s.e1, st);
        profile_push(s);
    }
};

static TrkProf* g_cos_prof = nullptr;
static void cos_prof_begin(int kind, double bytes, double flops, cudaStream_t st) { g_cos_prof = new TrkProf(kind, bytes, flops, st); }
static void cos_prof_end(cudaStream_t) { delete g_cos_prof; g_cos_prof = nullptr; }

Tracker::Tracker(double max_dist, double max_iou, int max_age, int n_init, int budget, int cap_tracks, int cap_dets)
    : max_dist_(max_dist), max_iou_(max_iou), max_age_(max_age), n_init_(n_init), budget_(budget), cap_t_(cap_tracks), cap_d_(cap_dets) {
    YDST_CHECK(budget >= 1 && cap_tracks >= 1 && cap_dets >= 1, "bad tracker capacities");
    const size_t nt = cap_t_, nd = cap_d_;
    YDST_CUDA(cudaMalloc(&mean_, nt * 8 * sizeof(float)));
    YDST_CUDA(cudaMalloc(&cov_, nt * 64 * sizeof(float)));
    YDST_CUDA(cudaMalloc(&gallery_, nt * budget_ * kFeat * sizeof(float)));
    YDST_CUDA(cudaMalloc(&det_n_, nd * kFeat * sizeof(float)));
    YDST_CUDA(cudaMalloc(&cost_, nt * nd * sizeof(float)));
    YDST_CUDA(cudaMalloc(&cost_t_, nt * nd * sizeof(float)));
    YDST_CUDA(cudaMalloc(&col4row_, (nt + nd) * sizeof(int)));
    YDST_CUDA(cudaMalloc(&over_, (nt + nd) * sizeof(int)));
    YDST_CUDA(cudaMalloc(&out_mean_, nt * 8 * sizeof(float)));
    YDST_CUDA(cudaMalloc(&lsap_work_, lsap_work_bytes((int)std::max(nt, nd), (int)std::max(nt, nd))));
    ibuf_cap_ = nt * budget_ * 2 + 20 * (nt + nd) + 1024;
    YDST_CUDA(cudaMalloc(&ibuf_, ibuf_cap_ * sizeof(int)));
    YDST_CUDA(cudaMallocHost(&h_ibuf_, ibuf_cap_ * sizeof(int)));
    YDST_CUDA(cudaMallocHost(&h_res_, 2 * (nt + nd) * sizeof(int)));
    YDST_CUDA(cudaMallocHost(&h_f_, nt * 8 * sizeof(float)));
    cos_.prof_begin = cos_prof_begin; cos_.prof_end = cos_prof_end;
    free_slots_.reserve(nt);
    for (int s = cap_t_ - 1; s >= 0; --s) free_slots_.push_back(s);
}

Tracker::~Tracker() {
    cudaFree(mean_); cudaFree(cov_); cudaFree(gallery_); cudaFree(det_n_); cudaFree(cost_); cudaFree(cost_t_);
    cudaFree(col4row_); cudaFree(over_); cudaFree(out_mean_); cudaFree(lsap_work_); cudaFree(ibuf_);
    cudaFreeHost(h_ibuf_); cudaFreeHost(h_res_); cudaFreeHost(h_f_);
}

int* Tracker::upload(const std::vector<int>& v, cudaStream_t st) {
    YDST_CHECK(ibuf_used_ + v.size() <= ibuf_cap_, "tracker index staging overflow");
    int* h = h_ibuf_ + ibuf_used_;
    int* d = ibuf_ + ibuf_used_;
    if (!v.empty()) {
        memcpy(h, v.data(), v.size() * sizeof(int));
        YDST_CUDA(cudaMemcpyAsync(d, h, v.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    }
    ibuf_used_ += (v.size() + 3) & ~(size_t)3;
    return d;
}

Tracker::Assign Tracker::solve(const float* cost, const std::vector<int>& tis, const std::vector<int>& dis, float max_dist, cudaStream_t st) {
    // cost is [nt][nd] on the device, already clamped.  scipy transposes when nd < nt.
    const int nt = (int)tis.size(), nd = (int)dis.size();
    const bool transposed = nd < nt;
    const int R = transposed ? nd : nt, C = transposed ? nt : nd;
    const float* c = cost;
    if (transposed) { TrkProf pr(TOP_TRANSPOSE, 8.0 * nt * nd, 0, st); launch_transpose(cost, cost_t_, nt, nd, st); c = cost_t_; }
    { TrkProf pr(TOP_LSAP, 4.0 * R * C, 0, st); launch_lsap(c, R, C, max_dist, col4row_, over_, lsap_work_, st); }
    launches_last += transposed ? 2 : 1;
    YDST_CUDA(cudaMemcpyAsync(h_res_, col4row_, R * sizeof(int), cudaMemcpyDeviceToHost, st));
    YDST_CUDA(cudaMemcpyAsync(h_res_ + R, over_, R * sizeof(int), cudaMemcpyDeviceToHost, st));
    YDST_CUDA(cudaStreamSynchronize(st));
    // pairs (row, col) of the ORIGINAL matrix, sorted by row (what scipy returns)
    std::vector<int> col_of_row(nt, -1), over_of_row(nt, 0);
    std::vector<char> col_used(nd, 0);
    for (int r = 0; r < R; ++r) {
        const int cc = h_res_[r];
        YDST_CHECK(cc >= 0 && cc < C, "linear assignment failed (infeasible cost matrix)");
        const int orow = transposed ? cc : r, ocol = transposed ? r : cc;
        col_of_row[orow] = ocol;
        over_of_row[orow] = h_res_[R + r];
        col_used[ocol] = 1;
    }
    Assign a;
    // linear_assignment.py:58-72: unmatched detections (column order), unmatched tracks (row order), then the
    // assigned pairs in row order -- a pair whose cost exceeds max_distance goes to BOTH unmatched lists.
    for (int col = 0; col < nd; ++col)
        if (!col_used[col]) a.um_d.push_back(dis[col]);
    for (int row = 0; row < nt; ++row)
        if (col_of_row[row] < 0) a.um_t.push_back(tis[row]);
    for (int row = 0; row < nt; ++row) {
        if (col_of_row[row] < 0) continue;
        if (over_of_row[row]) { a.um_t.push_back(tis[row]); a.um_d.push_back(dis[col_of_row[row]]); }
        else a.matches.emplace_back(tis[row], dis[col_of_row[row]]);
    }
    return a;
}

void Tracker::update(const float* tlwh, const float* feat, const int* payload_host, const float* cls_dev, int m, int32_t* out_host,
                     int* k_host, cudaStream_t st) {
    YDST_CHECK(m <= cap_d_, "%d detections exceed the tracker's detection capacity %d", m, cap_d_);
    static const bool trace = getenv("YDST_TRACKER_TRACE") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = trace ? now() : 0;
    double t_a = 0, t_b = 0, t_c = 0;
    ibuf_used_ = 0;
    launches_last = 0;
    std::vector<int> payload(m, 0);
    int* h_cls = h_res_ + 2 * (cap_t_ + cap_d_) - cap_d_;          // tail of the pinned result buffer
    if (m > 0 && !payload_host) {
        YDST_CHECK(cls_dev != nullptr, "tracker update needs payload_host or cls_dev");
        int* d_cls = over_ + cap_t_;                                // scratch: over_ has cap_t+cap_d ints, LSAP uses <= min(cap_t,cap_d)
        cls_to_int_kernel<<<(m + 127) / 128, 128, 0, st>>>(cls_dev, m, d_cls);
        count_launch();
        YDST_CUDA(cudaMemcpyAsync(h_cls, d_cls, m * sizeof(int), cudaMemcpyDeviceToHost, st));
        YDST_CUDA(cudaStreamSynchronize(st));
        for (int i = 0; i < m; ++i) payload[i] = h_cls[i];
    } else if (m > 0) {
        for (int i = 0; i < m; ++i) payload[i] = payload_host[i];
    }

    // ---- predict (tracker.py:95-113) ----
    const int n = (int)tracks.size();
    if (n > 0) {
        std::vector<int> slots(n);
        for (int i = 0; i < n; ++i) slots[i] = tracks[i].slot;
        { int* d_s = upload(slots, st); TrkProf pr(TOP_KF_PREDICT, 576.0 * n, 0, st); launch_kf_predict(mean_, cov_, d_s, n, st); }
        ++launches_last;
        for (auto& t : tracks) { t.age += 1; t.tsu += 1; }
    }
    if (m > 0) { TrkProf pr(TOP_NORMALIZE, 4096.0 * m, 0, st); launch_normalize_rows(feat, det_n_, m, st); ++launches_last; }

    // ---- match (tracker.py:56-93) ----
    std::vector<int> confirmed, unconfirmed, all_dets(m);
    for (int i = 0; i < n; ++i) (tracks[i].state == TRACK_CONFIRMED ? confirmed : unconfirmed).push_back(i);
    for (int j = 0; j < m; ++j) all_dets[j] = j;
    Assign A;
    if (confirmed.empty() || m == 0) {
        A.um_t = confirmed; A.um_d = all_dets;
    } else {
        const int na = (int)confirmed.size();
        std::vector<int> slots(na), row_ptr, seg(na + 1, 0);
        for (int r = 0; r < na; ++r) {
            const TrackHost& t = tracks[confirmed[r]];
            slots[r] = t.slot;
            for (int k = 0; k < t.gal_count; ++k) row_ptr.push_back(t.slot * budget_ + k);
            seg[r + 1] = (int)row_ptr.size();
        }
        const int G = (int)row_ptr.size();
        int* d_slots = upload(slots, st);
        int* d_rp = upload(row_ptr, st);
        int* d_seg = upload(seg, st);
        // 1 - max cosine over each track's gallery rows, gate, clamp: split + tcgen05 GEMM + segmented max (cosine_tc.cu)
        cos_.run(gallery_, d_rp, d_seg, G, na, det_n_, m, mean_, cov_, d_slots, tlwh, max_dist_, cost_, st);
        launches_last += cos_.launches_last;
        A = solve(cost_, confirmed, all_dets, (float)max_dist_, st);
    }
    if (trace) t_a = now();
    std::vector<int> iou_cand = unconfirmed, um_t_a;
    for (int k : A.um_t) (tracks[k].tsu == 1 ? iou_cand : um_t_a).push_back(k);
    Assign B;
    if (iou_cand.empty() || A.um_d.empty()) {
        B.um_t = iou_cand; B.um_d = A.um_d;
    } else {
        const int nb = (int)iou_cand.size(), mb = (int)A.um_d.size();
        std::vector<int> slots(nb), tsu(nb);
        for (int r = 0; r < nb; ++r) { slots[r] = tracks[iou_cand[r]].slot; tsu[r] = tracks[iou_cand[r]].tsu; }
        int* d_slots = upload(slots, st);
        int* d_tsu = upload(tsu, st);
        int* d_dets = upload(A.um_d, st);
        { TrkProf pr(TOP_IOU_COST, 4.0 * nb * mb + 32.0 * nb + 16.0 * mb, 0, st); launch_iou_cost(mean_, d_slots, d_tsu, nb, tlwh, d_dets, mb, max_iou_, cost_, st); }
        ++launches_last;
        B = solve(cost_, iou_cand, A.um_d, (float)max_iou_, st);
    }
    if (trace) t_b = now();
    std::vector<std::pair<int, int>> matches = A.matches;
    matches.insert(matches.end(), B.matches.begin(), B.matches.end());
    // capacity is checked BEFORE any track state is touched, so a refused frame leaves the tracker exactly as predict() left it
    // (slots of tracks deleted in this very step only become reusable on the next frame)
    YDST_CHECK(B.um_d.size() <= free_slots_.size(), "tracker capacity (%d tracks) exhausted: %zu new detections, %zu free slots", cap_t_,
               B.um_d.size(), free_slots_.size());
    last_matches = matches;

    // ---- update matched tracks (tracker.py:129-156, track.py:125-144) ----
    std::vector<int> app_src, app_dst;
    if (!matches.empty()) {
        std::vector<int> slots(matches.size()), dets(matches.size());
        for (size_t i = 0; i < matches.size(); ++i) {
            TrackHost& t = tracks[matches[i].first];
            slots[i] = t.slot; dets[i] = matches[i].second;
            app_src.push_back(matches[i].second);
            app_dst.push_back(t.slot * budget_ + t.gal_head);
            t.gal_head = (t.gal_head + 1) % budget_;
            t.gal_count = std::min(t.gal_count + 1, budget_);
            t.hits += 1;
            t.tsu = 0;
            if (t.state == TRACK_TENTATIVE && t.hits >= n_init_) t.state = TRACK_CONFIRMED;
            t.payload = payload[matches[i].second];
        }
        { int* d_s = upload(slots, st); int* d_d = upload(dets, st); TrkProf pr(TOP_KF_UPDATE, 592.0 * matches.size(), 0, st);
          launch_kf_update(mean_, cov_, d_s, tlwh, d_d, (int)matches.size(), st); }
        ++launches_last;
    }
    // ---- mark missed (track.py:146-152) ----
    for (int k : um_t_a) if (tracks[k].state == TRACK_TENTATIVE || tracks[k].tsu > max_age_) tracks[k].state = TRACK_DELETED;
    for (int k : B.um_t) if (tracks[k].state == TRACK_TENTATIVE || tracks[k].tsu > max_age_) tracks[k].state = TRACK_DELETED;
    // ---- spawn, in the order of unmatched_detections (tracker.py:160-161) ----
    if (!B.um_d.empty()) {
        // slots of tracks deleted in this very step become reusable only after the spawn kernels are queued,
        // which is fine: stream order makes the new state overwrite the old one after all readers ran.
        std::vector<int> slots, dets;
        for (int d : B.um_d) {
            YDST_CHECK(!free_slots_.empty(), "tracker capacity (%d tracks) exhausted", cap_t_);
            const int slot = free_slots_.back();
            free_slots_.pop_back();
            TrackHost t;
            t.id = next_id++; t.hits = 1; t.age = 1; t.tsu = 0; t.state = TRACK_TENTATIVE; t.payload = payload[d];
            t.slot = slot; t.gal_count = 1; t.gal_head = 1 % budget_;
            tracks.push_back(t);
            slots.push_back(slot); dets.push_back(d);
            app_src.push_back(d); app_dst.push_back(slot * budget_ + 0);
        }
        { int* d_d = upload(dets, st); int* d_s = upload(slots, st); TrkProf pr(TOP_KF_INITIATE, 304.0 * slots.size(), 0, st);
          launch_kf_initiate(tlwh, d_d, mean_, cov_, d_s, (int)slots.size(), st); }
        ++launches_last;
    }
    if (!app_src.empty()) {
        int* d_src = upload(app_src, st); int* d_dst = upload(app_dst, st);
        TrkProf pr(TOP_GALLERY_APPEND, 4096.0 * app_src.size(), 0, st);
        gallery_append_kernel<<<(unsigned)app_src.size(), 128, 0, st>>>(det_n_, d_src, gallery_, d_dst, (int)app_src.size());
        YDST_CUDA(cudaGetLastError());
        ++launches_last;
    }
    // ---- drop deleted tracks (tracker.py:162); galleries of non-active ids die with their slot (nn_matching.py:156) ----
    {
        size_t w = 0;
        for (size_t i = 0; i < tracks.size(); ++i) {
            if (tracks[i].state == TRACK_DELETED) free_slots_.push_back(tracks[i].slot);
            else tracks[w++] = tracks[i];
        }
        tracks.resize(w);
    }
    // ---- outputs (deep_sort.py:63-88) ----
    std::vector<int> out_idx, out_slots;
    for (size_t i = 0; i < tracks.size(); ++i)
        if (tracks[i].state == TRACK_CONFIRMED && tracks[i].tsu <= 1) { out_idx.push_back((int)i); out_slots.push_back(tracks[i].slot); }
    const int K = (int)out_idx.size();
    if (K > 0) {
        gather_mean_kernel<<<(K * 8 + 127) / 128, 128, 0, st>>>(mean_, upload(out_slots, st), K, out_mean_);
        YDST_CUDA(cudaGetLastError());
        ++launches_last;
        YDST_CUDA(cudaMemcpyAsync(h_f_, out_mean_, (size_t)K * 8 * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    if (trace) t_c = now();
    YDST_CUDA(cudaStreamSynchronize(st));
    if (trace)
        fprintf(stderr, "tracker_trace n %d m %d conf %zu | appearance stage %.1f us, iou stage (%zu x %zu) %.1f us, update+spawn enqueue %.1f us, final sync %.1f us\n",
                n, m, confirmed.size(), t_a - t_begin, iou_cand.size(), A.um_d.size(), t_b - t_a, t_c - t_b, now() - t_c);
    for (int k = 0; k < K; ++k) {
        // fp32, same operation order as the reference: w = a*h; tl = c - wh/2; br = wh + tl; clamp tl >= 0; int32 truncation
        volatile float cx = h_f_[k * 8 + 0], cy = h_f_[k * 8 + 1], a = h_f_[k * 8 + 2], h = h_f_[k * 8 + 3];
        volatile float w = a * h;
        volatile float x1 = cx - w / 2.f, y1 = cy - h / 2.f;
        volatile float x2 = w + x1, y2 = h + y1;
        const float cx1 = x1 < 0.f ? 0.f : x1, cy1 = y1 < 0.f ? 0.f : y1;
        const TrackHost& t = tracks[out_idx[k]];
        int32_t* o = out_host + k * 6;
        o[0] = (int32_t)cx1; o[1] = (int32_t)cy1; o[2] = (int32_t)x2; o[3] = (int32_t)y2; o[4] = t.id; o[5] = t.payload;
    }
    *k_host = K;
    count_launch(launches_last);
}

void Tracker::snapshot(int32_t* table_host, float* mean_host, int cap, int* n_host, cudaStream_t st) {
    const int n = (int)tracks.size();
    *n_host = n;
    YDST_CHECK(n <= cap, "snapshot buffer too small (%d tracks, cap %d)", n, cap);
    if (table_host)
        for (int i = 0; i < n; ++i) {
            const TrackHost& t = tracks[i];
            int32_t* r = table_host + i * 5;
            r[0] = t.id; r[1] = t.hits; r[2] = t.age; r[3] = t.tsu; r[4] = t.state;
        }
    if (mean_host && n > 0) {
        ibuf_used_ = 0;
        std::vector<int> slots(n);
        for (int i = 0; i < n; ++i) slots[i] = tracks[i].slot;
        gather_mean_kernel<<<(n * 8 + 127) / 128, 128, 0, st>>>(mean_, upload(slots, st), n, out_mean_);
        YDST_CUDA(cudaGetLastError());
        YDST_CUDA(cudaMemcpyAsync(h_f_, out_mean_, (size_t)n * 8 * sizeof(float), cudaMemcpyDeviceToHost, st));
        YDST_CUDA(cudaStreamSynchronize(st));
        memcpy(mean_host, h_f_, (size_t)n * 8 * sizeof(float));
    }
}

}  // namespace ydst
