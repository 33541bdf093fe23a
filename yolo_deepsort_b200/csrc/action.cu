// Device-side ActionIdentify.
//
// Replaces action/action_Identify.py:15-47 (the per-frame update of the orbit cache), action/orbit.py:5-26 (a bounded deque of
// bottom-centre points + time stamps per track id) and the five rules of action/actions.py:23-150 (TakeOff, Landing, Glide,
// FastCrossing, BreakInto).  This is bookkeeping on a few dozen (K,6) int32 rows per frame -- latency-bound by construction, one
// CTA -- kept on the device only so that a fused video loop never has to bring the rows back before the overlay needs them.
// Everything is computed in float64 / int32 exactly as the Python does (centre x = x1 + (x2 - x1) / 2, centre y = y2), so the
// emitted (track id, class id, rule) triples are identical; the reference returns them in dict-insertion order, which the host
// restores by sorting on the per-entry insertion sequence number this kernel hands back.
#include "action.cuh"

#include <algorithm>
#include <cstring>
#include <vector>

namespace ydst {

enum ActionKind { ACTION_TAKEOFF = 0, ACTION_LANDING = 1, ACTION_GLIDE = 2, ACTION_FAST_CROSSING = 3, ACTION_BREAK_INTO = 4 };

struct ActionRule { int kind, class_id; double p0, p1; };      // delta (x, y) | speed | timeout
static constexpr int kMaxOrbit = 16, kMaxRules = 16;

struct ActionState {                                           // struct of arrays, `cap` cache entries
    int* track_id;            // -1: free
    int* class_id;
    int* age;
    int* count;               // points in the deque (<= max_size)
    int* head;                // index of the OLDEST point in the ring
    long long* seq;           // insertion sequence number (dict order)
    double* pts;              // [cap][kMaxOrbit][2]
    double* ts;               // [cap][kMaxOrbit]
};

struct ActionRules { ActionRule r[kMaxRules]; int n; };

// One CTA.  rows: K x 6 int32 [x1, y1, x2, y2, track id, class id] (device).  out: up to K * n_rules records of 4 long long
// (seq, track id, class id, rule index); *n_out counts them; *next_seq is the running insertion counter; err: 1 = cache full.
__global__ void __launch_bounds__(1024) action_update_kernel(ActionState st, int cap, int max_age, int max_size, ActionRules rules,
                                                             const int* __restrict__ rows, int K, double now, long long* next_seq,
                                                             long long* out, int* n_out, int* err) {
    extern __shared__ int sh[];                                // [K] entry of each row, [K] "is new" flags / ranks
    int* entry = sh;
    int* rank = sh + K;
    __shared__ int n_new_total;
    const int tid = threadIdx.x, nt = blockDim.x;
    if (tid == 0) { *n_out = 0; n_new_total = 0; }
    // 1. find the cache entry of every row (action_Identify.py:21-26)
    for (int i = tid; i < K; i += nt) { entry[i] = -1; rank[i] = 0; }
    __syncthreads();
    for (int e = tid; e < cap; e += nt) {
        const int id = st.track_id[e];
        if (id < 0) continue;
        for (int i = 0; i < K; ++i)
            if (rows[i * 6 + 4] == id) entry[i] = e;
    }
    __syncthreads();
    // 2. new ids take free entries in row order: rank among the new rows, then the rank-th free entry
    if (tid == 0) {
        int r = 0;
        for (int i = 0; i < K; ++i)
            if (entry[i] < 0) rank[i] = r++;
        n_new_total = r;
    }
    __syncthreads();
    if (n_new_total > 0 && tid == 0) {
        int e = 0, given = 0;
        for (int i = 0; i < K && given < n_new_total; ++i) {
            if (entry[i] >= 0) continue;
            while (e < cap && st.track_id[e] >= 0) ++e;
            if (e >= cap) { *err = 1; break; }
            entry[i] = e;
            st.track_id[e] = rows[i * 6 + 4];
            st.class_id[e] = rows[i * 6 + 5];
            st.age[e] = 0; st.count[e] = 0; st.head[e] = 0;
            st.seq[e] = *next_seq + rank[i];
            rank[i] = -1;                                      // marks "created this frame": Orbit() without a point (:24)
            ++given; ++e;
        }
        *next_seq += n_new_total;
    }
    __syncthreads();
    // 3. known ids: Orbit.update (orbit.py:22-26) -- age 0, append the bottom-centre point and the time stamp
    for (int i = tid; i < K; i += nt) {
        const int e = entry[i];
        if (e < 0 || rank[i] == -1) continue;
        st.age[e] = 0;
        const int x1 = rows[i * 6 + 0], x2 = rows[i * 6 + 2], y2 = rows[i * 6 + 3];
        const double cx = (double)x1 + (double)(x2 - x1) / 2.0, cy = (double)y2;
        int cnt = st.count[e], hd = st.head[e];
        int slot;
        if (cnt < max_size) { slot = (hd + cnt) % max_size; ++cnt; }
        else { slot = hd; hd = (hd + 1) % max_size; }         // deque(maxlen): the oldest point falls out
        st.pts[((size_t)e * kMaxOrbit + slot) * 2 + 0] = cx;
        st.pts[((size_t)e * kMaxOrbit + slot) * 2 + 1] = cy;
        st.ts[(size_t)e * kMaxOrbit + slot] = now;
        st.count[e] = cnt; st.head[e] = hd;
    }
    __syncthreads();
    // 4. everybody else ages; max_age frames without a row delete the orbit (:31-38).  A row's entry has age 0 by now.
    for (int e = tid; e < cap; e += nt) {
        if (st.track_id[e] < 0) continue;
        bool targeted = false;
        for (int i = 0; i < K; ++i) targeted = targeted || entry[i] == e;
        if (targeted) continue;
        const int a = st.age[e] + 1;
        st.age[e] = a;
        if (a >= max_age) st.track_id[e] = -1;
    }
    __syncthreads();
    // 5. rules on the orbits seen this frame (:40-45)
    for (int i = tid; i < K; i += nt) {
        const int e = entry[i];
        if (e < 0) continue;
        const int cnt = st.count[e], hd = st.head[e], cls = st.class_id[e];
        for (int r = 0; r < rules.n; ++r) {
            const ActionRule R = rules.r[r];
            if (cnt == 0 || cls != R.class_id) continue;       // `len(orbit.deque) == 0 or orbit.class_id != self.class_id`
            bool ok = false;
            if (R.kind == ACTION_BREAK_INTO) {
                ok = (double)cnt > R.p0;
            } else {
                // `is_x` ends up True iff the condition holds for EVERY consecutive pair and there is at least one pair
                ok = cnt >= 2;
                for (int k = 1; k < cnt && ok; ++k) {
                    const int a = (hd + k - 1) % max_size, b = (hd + k) % max_size;
                    const double ax = st.pts[((size_t)e * kMaxOrbit + a) * 2], ay = st.pts[((size_t)e * kMaxOrbit + a) * 2 + 1];
                    const double bx = st.pts[((size_t)e * kMaxOrbit + b) * 2], by = st.pts[((size_t)e * kMaxOrbit + b) * 2 + 1];
                    bool c;
                    if (R.kind == ACTION_TAKEOFF) c = ay - by > R.p1 && fabs(ax - bx) > R.p0;
                    else if (R.kind == ACTION_LANDING) c = by - ay > R.p1 && fabs(ax - bx) > R.p0;
                    else if (R.kind == ACTION_GLIDE) c = fabs(by - ay) < R.p1 && fabs(bx - ax) > R.p0;
                    else {
                        const double ta = st.ts[(size_t)e * kMaxOrbit + a], tb = st.ts[(size_t)e * kMaxOrbit + b];
                        c = fabs(bx - ax) / ((tb - ta) * 1000.0) > R.p0;   // inf / nan exactly as numpy's float64 division
                    }
                    ok = c;
                }
            }
            if (ok) {
                const int o = atomicAdd(n_out, 1);
                out[(size_t)o * 4 + 0] = st.seq[e]; out[(size_t)o * 4 + 1] = st.track_id[e];
                out[(size_t)o * 4 + 2] = cls;       out[(size_t)o * 4 + 3] = r;
            }
        }
    }
}

struct ActionIdentifyDev {
    ActionState st{};
    ActionRules rules{};
    int cap = 0, max_age = 30, max_size = 4, cap_rows = 0;
    int* d_rows = nullptr; long long* d_out = nullptr; int* d_cnt = nullptr; long long* d_seq = nullptr;
    int* h_rows = nullptr; long long* h_out = nullptr; int* h_cnt = nullptr;

    ActionIdentifyDev(int max_age_, int max_size_, const ActionRule* r, int n_rules, int cap_) : cap(cap_), max_age(max_age_), max_size(max_size_) {
        YDST_CHECK(max_size >= 1 && max_size <= kMaxOrbit, "ActionIdentify: max_size must be in 1..%d (got %d)", kMaxOrbit, max_size);
        YDST_CHECK(n_rules >= 0 && n_rules <= kMaxRules && cap >= 1, "ActionIdentify: at most %d rules, capacity >= 1", kMaxRules);
        rules.n = n_rules;
        for (int i = 0; i < n_rules; ++i) rules.r[i] = r[i];
        cap_rows = cap;
        YDST_CUDA(cudaMalloc(&st.track_id, cap * sizeof(int)));
        YDST_CUDA(cudaMemset(st.track_id, 0xFF, cap * sizeof(int)));
        YDST_CUDA(cudaMalloc(&st.class_id, cap * sizeof(int)));
        YDST_CUDA(cudaMalloc(&st.age, cap * sizeof(int)));
        YDST_CUDA(cudaMalloc(&st.count, cap * sizeof(int)));
        YDST_CUDA(cudaMalloc(&st.head, cap * sizeof(int)));
        YDST_CUDA(cudaMalloc(&st.seq, cap * sizeof(long long)));
        YDST_CUDA(cudaMalloc(&st.pts, (size_t)cap * kMaxOrbit * 2 * sizeof(double)));
        YDST_CUDA(cudaMalloc(&st.ts, (size_t)cap * kMaxOrbit * sizeof(double)));
        YDST_CUDA(cudaMalloc(&d_rows, (size_t)cap_rows * 6 * sizeof(int)));
        YDST_CUDA(cudaMalloc(&d_out, (size_t)cap_rows * kMaxRules * 4 * sizeof(long long)));
        YDST_CUDA(cudaMalloc(&d_cnt, 2 * sizeof(int)));
        YDST_CUDA(cudaMalloc(&d_seq, sizeof(long long)));
        YDST_CUDA(cudaMemset(d_cnt, 0, 2 * sizeof(int)));
        YDST_CUDA(cudaMemset(d_seq, 0, sizeof(long long)));
        YDST_CUDA(cudaMallocHost(&h_rows, (size_t)cap_rows * 6 * sizeof(int)));
        YDST_CUDA(cudaMallocHost(&h_out, (size_t)cap_rows * kMaxRules * 4 * sizeof(long long)));
        YDST_CUDA(cudaMallocHost(&h_cnt, 2 * sizeof(int)));
    }
    ~ActionIdentifyDev() {
        cudaFree(st.track_id); cudaFree(st.class_id); cudaFree(st.age); cudaFree(st.count); cudaFree(st.head); cudaFree(st.seq);
        cudaFree(st.pts); cudaFree(st.ts); cudaFree(d_rows); cudaFree(d_out); cudaFree(d_cnt); cudaFree(d_seq);
        cudaFreeHost(h_rows); cudaFreeHost(h_out); cudaFreeHost(h_cnt);
    }
    // rows: K x 6 int32 on the host (the (K,6) block DeepSort.update returned); triples_out: (track id, class id, rule index) in the
    // reference's order, at most K * n_rules of them.  Synchronises.
    int update(const int32_t* rows_host, int K, double now, int32_t* triples_out, cudaStream_t stream) {
        YDST_CHECK(K >= 0 && K <= cap_rows, "ActionIdentify: %d rows exceed the capacity %d", K, cap_rows);
        if (K > 0) {
            memcpy(h_rows, rows_host, (size_t)K * 6 * sizeof(int));
            YDST_CUDA(cudaMemcpyAsync(d_rows, h_rows, (size_t)K * 6 * sizeof(int), cudaMemcpyHostToDevice, stream));
        }
        action_update_kernel<<<1, 1024, (size_t)std::max(K, 1) * 2 * sizeof(int), stream>>>(st, cap, max_age, max_size, rules, d_rows, K, now, d_seq, d_out,
                                                                                         d_cnt, d_cnt + 1);
        YDST_CUDA(cudaGetLastError());
        count_launch();
        YDST_CUDA(cudaMemcpyAsync(h_cnt, d_cnt, 2 * sizeof(int), cudaMemcpyDeviceToHost, stream));
        YDST_CUDA(cudaStreamSynchronize(stream));
        YDST_CHECK(h_cnt[1] == 0, "ActionIdentify: more than %d live orbits", cap);
        const int n = h_cnt[0];
        if (n > 0) {
            YDST_CUDA(cudaMemcpyAsync(h_out, d_out, (size_t)n * 4 * sizeof(long long), cudaMemcpyDeviceToHost, stream));
            YDST_CUDA(cudaStreamSynchronize(stream));
        }
        // dict order: by insertion sequence, rules in list order
        std::vector<int> idx(n);
        for (int i = 0; i < n; ++i) idx[i] = i;
        std::sort(idx.begin(), idx.end(), [&](int a, int b) {
            if (h_out[a * 4] != h_out[b * 4]) return h_out[a * 4] < h_out[b * 4];
            return h_out[a * 4 + 3] < h_out[b * 4 + 3];
        });
        for (int i = 0; i < n; ++i) {
            const long long* o = h_out + (size_t)idx[i] * 4;
            triples_out[i * 3 + 0] = (int32_t)o[1]; triples_out[i * 3 + 1] = (int32_t)o[2]; triples_out[i * 3 + 2] = (int32_t)o[3];
        }
        return n;
    }
};

ActionIdentifyDev* action_create(int max_age, int max_size, const int* kinds, const int* class_ids, const double* p0, const double* p1, int n_rules, int cap) {
    std::vector<ActionRule> r(std::max(n_rules, 1));
    for (int i = 0; i < n_rules; ++i) {
        YDST_CHECK(kinds[i] >= ACTION_TAKEOFF && kinds[i] <= ACTION_BREAK_INTO, "ActionIdentify: unknown rule kind %d", kinds[i]);
        r[i] = ActionRule{kinds[i], class_ids[i], p0[i], p1[i]};
    }
    return new ActionIdentifyDev(max_age, max_size, r.data(), n_rules, cap);
}
void action_destroy(ActionIdentifyDev* a) { delete a; }
int action_update(ActionIdentifyDev* a, const int32_t* rows_host, int K, double now, int32_t* triples_out, cudaStream_t stream) {
    return a->update(rows_host, K, now, triples_out, stream);
}

}  // namespace ydst
