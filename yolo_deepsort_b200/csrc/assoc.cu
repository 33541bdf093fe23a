// DeepSORT association kernels for sm_100a: batched Kalman filter on struct-of-arrays track state,
// appearance / IoU cost matrices with fused gating, and an exact on-device linear assignment.
//
// Reference stages replaced (deep_sort/sort/): kalman_filter.py:54-256, nn_matching.py:30-100,158-187,
// linear_assignment.py:51-56,147-203, iou_matching.py:5-91, plus scipy.optimize.linear_sum_assignment.
// These are HBM/latency-bound fp32 kernels: coalesced 16/32-byte accesses, warp-shuffle data exchange
// (8 lanes own the 8 rows of one covariance), no tensor cores (costs are compared against 0.3 / 5.9915,
// so they are kept in full fp32; the LSAP duals are fp64 exactly like scipy's).
#include "assoc.cuh"

#include <math_constants.h>

namespace ydst {

static inline int cdiv(long long a, int b) { return (int)((a + b - 1) / b); }

// ================================================================================================
// Kalman filter.  Thread layout for predict/update: 8 consecutive lanes per track, lane r owns row r
// of the 8x8 covariance (two float4) and mean[r]; 4 tracks per warp => 1 KB contiguous per warp access.
// ================================================================================================
__device__ __forceinline__ void tlwh_to_xyah(const float* t, float& x, float& y, float& a, float& h) {
    // Detection.to_xyah / tracker.py:145-146: xy += wh/2 ; a = w/h
    x = t[0] + t[2] / 2.f;
    y = t[1] + t[3] / 2.f;
    a = t[2] / t[3];
    h = t[3];
}

__global__ void kf_initiate_kernel(const float* __restrict__ det_tlwh, const int* __restrict__ det_idx, float* __restrict__ mean,
                                   float* __restrict__ cov, const int* __restrict__ slot_idx, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int d = det_idx ? det_idx[i] : i;
    const int slot = slot_idx ? slot_idx[i] : i;
    float x, y, a, h;
    tlwh_to_xyah(det_tlwh + d * 4, x, y, a, h);
    float* m = mean + (long long)slot * 8;
    m[0] = x; m[1] = y; m[2] = a; m[3] = h; m[4] = 0.f; m[5] = 0.f; m[6] = 0.f; m[7] = 0.f;
    // std = [2*(1/20)*h, .., 1e-2, .., 10*(1/160)*h, .., 1e-5, ..]  (kalman_filter.py:75-84), squared
    const float sp = 0.1f * h, sv = 0.0625f * h;
    const float dg[8] = {sp * sp, sp * sp, 1e-2f * 1e-2f, sp * sp, sv * sv, sv * sv, 1e-5f * 1e-5f, sv * sv};
    float* c = cov + (long long)slot * 64;
    for (int r = 0; r < 8; ++r)
        for (int q = 0; q < 8; ++q) c[r * 8 + q] = r == q ? dg[r] : 0.f;
}
void launch_kf_initiate(const float* det_tlwh, const int* det_idx, float* mean, float* cov, const int* slot_idx, int n, cudaStream_t st) {
    if (n == 0) return;
    kf_initiate_kernel<<<cdiv(n, 128), 128, 0, st>>>(det_tlwh, det_idx, mean, cov, slot_idx, n);
    YDST_CUDA(cudaGetLastError());
}

__global__ void __launch_bounds__(256) kf_predict_kernel(float* __restrict__ mean, float* __restrict__ cov, const int* __restrict__ idx, int n) {
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;     // track
    const int r = threadIdx.x & 7;
    const int lane = threadIdx.x & 31, gl = lane & ~7;
    const bool on = g < n;
    const int slot = on ? (idx ? idx[g] : g) : 0;
    float4* crow = reinterpret_cast<float4*>(cov + (long long)slot * 64 + r * 8);
    float p[8];
    float m = 0.f;
    if (on) {
        const float4 a = crow[0], b = crow[1];
        p[0] = a.x; p[1] = a.y; p[2] = a.z; p[3] = a.w; p[4] = b.x; p[5] = b.y; p[6] = b.z; p[7] = b.w;
        m = mean[(long long)slot * 8 + r];
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) p[j] = 0.f;
    }
    const float h = __shfl_sync(0xffffffffu, m, gl + 3);            // height BEFORE the motion step (kalman_filter.py:109)
    // F P : rows 0..3 += rows 4..7
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float o = __shfl_sync(0xffffffffu, p[j], gl + ((r + 4) & 7));
        if (r < 4) p[j] = p[j] + o;
    }
    // (F P) F^T : cols 0..3 += cols 4..7
#pragma unroll
    for (int j = 0; j < 4; ++j) p[j] = p[j] + p[j + 4];
    // + Q = diag([.05h,.05h,1e-2,.05h, h/160,h/160,1e-5,h/160]^2)
    const float sp = h * 0.05f, sv = h * 0.00625f;
    const float q = r < 4 ? (r == 2 ? 1e-2f * 1e-2f : sp * sp) : (r == 6 ? 1e-5f * 1e-5f : sv * sv);
#pragma unroll
    for (int j = 0; j < 8; ++j)
        if (j == r) p[j] = p[j] + q;
    const float mo = __shfl_sync(0xffffffffu, m, gl + ((r + 4) & 7));
    if (r < 4) m = m + mo;
    if (on) {
        crow[0] = make_float4(p[0], p[1], p[2], p[3]);
        crow[1] = make_float4(p[4], p[5], p[6], p[7]);
        mean[(long long)slot * 8 + r] = m;
    }
}
void launch_kf_predict(float* mean, float* cov, const int* idx, int n, cudaStream_t st) {
    if (n == 0) return;
    kf_predict_kernel<<<cdiv((long long)n * 8, 256), 256, 0, st>>>(mean, cov, idx, n);
    YDST_CUDA(cudaGetLastError());
}

// 4x4 LU with partial pivoting + solve for one right-hand side (LAPACK sgetf2/sgetrs operation order:
// first-max pivot, multiply the sub-column by the reciprocal pivot, right-looking rank-1 updates).
__device__ __forceinline__ void lu4_solve(float (&A)[4][4], float (&b)[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int pv = j;
        float mx = fabsf(A[j][j]);
#pragma unroll
        for (int i = j + 1; i < 4; ++i) {
            const float v = fabsf(A[i][j]);
            if (v > mx) { mx = v; pv = i; }
        }
        if (pv != j) {
#pragma unroll
            for (int i = j + 1; i < 4; ++i)
                if (i == pv) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) { const float t = A[j][q]; A[j][q] = A[i][q]; A[i][q] = t; }
                    const float t = b[j]; b[j] = b[i]; b[i] = t;
                }
        }
        const float rcp = 1.f / A[j][j];
#pragma unroll
        for (int i = j + 1; i < 4; ++i) {
            A[i][j] = A[i][j] * rcp;
#pragma unroll
            for (int q = j + 1; q < 4; ++q) A[i][q] = A[i][q] - A[i][j] * A[j][q];
        }
    }
    // L y = Pb (unit lower), U x = y
#pragma unroll
    for (int i = 1; i < 4; ++i)
#pragma unroll
        for (int q = 0; q < i; ++q) b[i] = b[i] - A[i][q] * b[q];
#pragma unroll
    for (int i = 3; i >= 0; --i) {
#pragma unroll
        for (int q = i + 1; q < 4; ++q) b[i] = b[i] - A[i][q] * b[q];
        b[i] = b[i] / A[i][i];
    }
}

__global__ void __launch_bounds__(256) kf_update_kernel(float* __restrict__ mean, float* __restrict__ cov, const int* __restrict__ idx,
                                                        const float* __restrict__ det_tlwh, const int* __restrict__ det_idx, int n) {
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const int r = threadIdx.x & 7;
    const int lane = threadIdx.x & 31, gl = lane & ~7;
    const bool on = g < n;
    const int slot = on ? (idx ? idx[g] : g) : 0;
    float4* crow = reinterpret_cast<float4*>(cov + (long long)slot * 64 + r * 8);
    float p[8], m = 0.f, z[4] = {0.f, 0.f, 0.f, 1.f};
    if (on) {
        const float4 a = crow[0], b = crow[1];
        p[0] = a.x; p[1] = a.y; p[2] = a.z; p[3] = a.w; p[4] = b.x; p[5] = b.y; p[6] = b.z; p[7] = b.w;
        m = mean[(long long)slot * 8 + r];
        const int d = det_idx ? det_idx[g] : g;
        tlwh_to_xyah(det_tlwh + d * 4, z[0], z[1], z[2], z[3]);
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) p[j] = (j == r) ? 1.f : 0.f;
    }
    const float h = __shfl_sync(0xffffffffu, m, gl + 3);
    // projected covariance S = P[:4,:4] + diag([.05h,.05h,1e-1,.05h]^2)   (kalman_filter.py:143-159)
    float S[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) S[a][b] = __shfl_sync(0xffffffffu, p[b], gl + a);
    const float sp = h * 0.05f;
    S[0][0] = S[0][0] + sp * sp; S[1][1] = S[1][1] + sp * sp; S[2][2] = S[2][2] + 1e-1f * 1e-1f; S[3][3] = S[3][3] + sp * sp;
    if (!on) { S[0][0] = S[1][1] = S[2][2] = S[3][3] = 1.f; }
    // Kalman gain row r:  S x = (P H)^T[:, r] = P[r][:4]
    float LU[4][4], k[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        k[a] = p[a];
#pragma unroll
        for (int b = 0; b < 4; ++b) LU[a][b] = S[a][b];
    }
    lu4_solve(LU, k);
    // mean += innovation . K^T
    float innov[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) innov[a] = z[a] - __shfl_sync(0xffffffffu, m, gl + a);
    float dm = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) dm = fmaf(innov[a], k[a], dm);
    m = m + dm;
    // cov -= (K S) K^T
    float ks[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        float s = 0.f;
#pragma unroll
        for (int b = 0; b < 4; ++b) s = fmaf(k[b], S[b][a], s);
        ks[a] = s;
    }
#pragma unroll
    for (int d = 0; d < 8; ++d) {
        float s = 0.f;
#pragma unroll
        for (int a = 0; a < 4; ++a) s = fmaf(ks[a], __shfl_sync(0xffffffffu, k[a], gl + d), s);
        p[d] = p[d] - s;
    }
    if (on) {
        crow[0] = make_float4(p[0], p[1], p[2], p[3]);
        crow[1] = make_float4(p[4], p[5], p[6], p[7]);
        mean[(long long)slot * 8 + r] = m;
    }
}
void launch_kf_update(float* mean, float* cov, const int* idx, const float* det_tlwh, const int* det_idx, int n, cudaStream_t st) {
    if (n == 0) return;
    kf_update_kernel<<<cdiv((long long)n * 8, 256), 256, 0, st>>>(mean, cov, idx, det_tlwh, det_idx, n);
    YDST_CUDA(cudaGetLastError());
}

__global__ void gate_position_kernel(const float* __restrict__ mean, const float* __restrict__ cov, const int* __restrict__ idx, int n,
                                     const float* __restrict__ det_tlwh, int m, float* __restrict__ maha) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)n * m) return;
    const int i = (int)(e / m), j = (int)(e % m);
    const int slot = idx ? idx[i] : i;
    float zx, zy, za, zh;
    tlwh_to_xyah(det_tlwh + j * 4, zx, zy, za, zh);
    maha[e] = maha_position(mean + (long long)slot * 8, cov + (long long)slot * 64, zx, zy);
}
void launch_gate_position(const float* mean, const float* cov, const int* idx, int n, const float* det_tlwh, int m, float* maha,
                          cudaStream_t st) {
    if ((long long)n * m == 0) return;
    gate_position_kernel<<<cdiv((long long)n * m, 256), 256, 0, st>>>(mean, cov, idx, n, det_tlwh, m, maha);
    YDST_CUDA(cudaGetLastError());
}

// ================================================================================================
// Appearance cost
// ================================================================================================
__global__ void __launch_bounds__(128) normalize_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, int n) {
    const int row = blockIdx.x;
    const float4 v = reinterpret_cast<const float4*>(src + (long long)row * kFeat)[threadIdx.x];
    float q = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    __shared__ float red[4];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = q;
    __syncthreads();
    const float nrm = sqrtf(red[0] + red[1] + red[2] + red[3]);
    reinterpret_cast<float4*>(dst + (long long)row * kFeat)[threadIdx.x] = make_float4(v.x / nrm, v.y / nrm, v.z / nrm, v.w / nrm);
}
void launch_normalize_rows(const float* src, float* dst, int n, cudaStream_t st) {
    if (n == 0) return;
    normalize_rows_kernel<<<n, 128, 0, st>>>(src, dst, n);
    YDST_CUDA(cudaGetLastError());
}

__global__ void iou_cost_kernel(const float* __restrict__ mean, const int* __restrict__ idx, const int* __restrict__ tsu, int n,
                                const float* __restrict__ det_tlwh, const int* __restrict__ det_idx, int m, float max_dist,
                                float clamp_val, float* __restrict__ cost) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)n * m) return;
    const int i = (int)(e / m), j = (int)(e % m);
    const int slot = idx ? idx[i] : i;
    const float* mu = mean + (long long)slot * 8;
    // Track.to_tlwh (track.py:81-94): w = a*h ; tl = centre - wh/2
    // every product/sum below is a separately rounded fp32 operation in the reference (torch elementwise ops), so the
    // intrinsics keep nvcc from contracting them into FMAs: the IoU cost is bit-exact, not merely close
    const float th = mu[3], tw = __fmul_rn(mu[2], th);
    const float tx = __fsub_rn(mu[0], tw / 2.f), ty = __fsub_rn(mu[1], th / 2.f);
    const float* d = det_tlwh + (det_idx ? det_idx[j] : j) * 4;
    const float dx = d[0], dy = d[1], dw = d[2], dh = d[3];
    // iou (iou_matching.py:25-41): +1 on the intersection extent only
    const float iw = fmaxf(__fadd_rn(__fsub_rn(fminf(__fadd_rn(tx, tw), __fadd_rn(dw, dx)), fmaxf(tx, dx)), 1.f), 0.f);
    const float ih = fmaxf(__fadd_rn(__fsub_rn(fminf(__fadd_rn(ty, th), __fadd_rn(dh, dy)), fmaxf(ty, dy)), 1.f), 0.f);
    const float inter = __fmul_rn(iw, ih);
    float c = __fsub_rn(1.f, __fdiv_rn(inter, __fsub_rn(__fadd_rn(__fmul_rn(tw, th), __fmul_rn(dw, dh)), inter)));
    if (tsu && tsu[i] > 1) c = kInftyCost;      // tsu is indexed by cost-matrix row
    if (c > max_dist) c = clamp_val;
    cost[e] = c;
}
void launch_iou_cost(const float* mean, const int* idx, const int* tsu, int n, const float* det_tlwh, const int* det_idx, int m,
                     double max_dist, float* cost, cudaStream_t st) {
    if ((long long)n * m == 0) return;
    const float clamp_val = (float)(max_dist + 1e-5);
    iou_cost_kernel<<<cdiv((long long)n * m, 256), 256, 0, st>>>(mean, idx, tsu, n, det_tlwh, det_idx, m, (float)max_dist, clamp_val, cost);
    YDST_CUDA(cudaGetLastError());
}

__global__ void transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols) {
    __shared__ float tile[32][33];
    const int x = blockIdx.x * 32 + threadIdx.x, y0 = blockIdx.y * 32;
    for (int k = threadIdx.y; k < 32; k += 8)
        if (x < cols && y0 + k < rows) tile[k][threadIdx.x] = src[(long long)(y0 + k) * cols + x];
    __syncthreads();
    const int ox = blockIdx.y * 32 + threadIdx.x, oy0 = blockIdx.x * 32;
    for (int k = threadIdx.y; k < 32; k += 8)
        if (ox < rows && oy0 + k < cols) dst[(long long)(oy0 + k) * rows + ox] = tile[threadIdx.x][k];
}
void launch_transpose(const float* src, float* dst, int rows, int cols, cudaStream_t st) {
    if ((long long)rows * cols == 0) return;
    transpose_kernel<<<dim3(cdiv(cols, 32), cdiv(rows, 32)), dim3(32, 8), 0, st>>>(src, dst, rows, cols);
    YDST_CUDA(cudaGetLastError());
}

// ================================================================================================
// Linear assignment: Crouse's shortest augmenting path, one CTA, columns scanned in parallel.
// Same `remaining` permutation, fp64 duals and selection rule as scipy, so ties resolve identically:
//   among the scanned columns with the lowest path cost pick the LAST unassigned one in array order,
//   or, if none is unassigned, the FIRST one.
// ================================================================================================
// A scan candidate is (path cost, key): the winner is the lexicographic minimum, with
//   key = C - 1 - it  for an unassigned column (so the LAST unassigned column in array order wins a tie), and
//   key = C + it      for an assigned one      (FIRST assigned column, and only if no unassigned column ties) --
// scipy's selection rule as one integer, which keeps the warp reduction to three shuffles per level.
struct LsapBest {
    double val;
    int key;
};
__device__ __forceinline__ bool lsap_better(const LsapBest& a, const LsapBest& b) {
    return a.val < b.val || (a.val == b.val && a.key < b.key);
}
// Warp-wide lexicographic minimum of (val, key) with three redux.sync instead of fifteen shuffles: the double is mapped to an
// order-preserving unsigned 64-bit integer (sign bit flipped for positives, all bits for negatives), its high and low words and
// then the key are reduced in turn, each step among the lanes still tied on the previous ones.  Exact: no arithmetic is involved.
__device__ __forceinline__ LsapBest lsap_warp_min(const LsapBest& v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v.val);
    const unsigned long long o = b ^ ((b >> 63) ? ~0ull : 0x8000000000000000ull);
    const unsigned hi = (unsigned)(o >> 32), lo = (unsigned)o;
    const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
    const unsigned mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xffffffffu);
    const bool tied = hi == mhi && lo == mlo;
    const unsigned mkey = __reduce_min_sync(0xffffffffu, tied ? (unsigned)v.key : 0x7fffffffu);
    const unsigned long long mo = ((unsigned long long)mhi << 32) | mlo;
    const unsigned long long mb = mo ^ ((mo >> 63) ? 0x8000000000000000ull : ~0ull);
    LsapBest r;
    r.val = __longlong_as_double((long long)mb);
    r.key = (int)mkey;
    return r;
}

struct LsapWork {
    double *u, *v, *spc;
    int *path, *col4row, *row4col, *remaining;
    int *ep_col;          // shortestPathCosts[j] is valid for the augmentation whose number it holds (otherwise: +inf, as scipy fills it)
    int *vis_col, *vis_row, *vis_bit;   // per augmentation: columns removed (SC), rows visited (SR), positions of `remaining` overwritten
    float* rowbuf;        // [C]: cost row of the next augmentation's first scan (known in advance: row cur + 1)
};

static __host__ __device__ inline size_t lsap_align16(size_t x) { return (x + 15) & ~(size_t)15; }
static __host__ __device__ inline size_t lsap_state_bytes(int R, int C) {
    return lsap_align16(sizeof(double) * R) + 2 * lsap_align16(sizeof(double) * C) + 7 * lsap_align16(sizeof(int) * C) +
           lsap_align16(sizeof(int) * R) + lsap_align16(sizeof(float) * C);
}
static __host__ __device__ inline void lsap_carve(unsigned char* p, int R, int C, LsapWork& w) {
    w.u = (double*)p; p += lsap_align16(sizeof(double) * R);
    w.v = (double*)p; p += lsap_align16(sizeof(double) * C);
    w.spc = (double*)p; p += lsap_align16(sizeof(double) * C);
    w.path = (int*)p; p += lsap_align16(sizeof(int) * C);
    w.row4col = (int*)p; p += lsap_align16(sizeof(int) * C);
    w.remaining = (int*)p; p += lsap_align16(sizeof(int) * C);
    w.ep_col = (int*)p; p += lsap_align16(sizeof(int) * C);
    w.vis_col = (int*)p; p += lsap_align16(sizeof(int) * C);
    w.vis_row = (int*)p; p += lsap_align16(sizeof(int) * C);
    w.vis_bit = (int*)p; p += lsap_align16(sizeof(int) * C);
    w.col4row = (int*)p; p += lsap_align16(sizeof(int) * R);
    w.rowbuf = (float*)p;
}

// state_in_smem: the whole solver state lives in dynamic shared memory instead of the global workspace -- the algorithm is a
// chain of dependent scans, so every global round trip (~700 clk) sat on the critical path.  cost_in_smem additionally stages
// the cost matrix when it is small.
//
// What scipy does per augmentation -- fill shortestPathCosts with inf, clear SR / SC, reset `remaining` to C-1..0, then, after the
// path is found, sweep ALL rows and columns for the dual update -- costs O(R + C) work and three CTA-wide barriers around a
// search that usually ends after ONE scan.  The same state is kept lazily here, with identical results:
//   * shortestPathCosts[j] carries the number of the augmentation that wrote it (ep_col); anything older reads as +inf;
//   * the removed columns (SC), the visited rows (SR) and the overwritten positions of `remaining` are LOGGED as the search goes
//     (one of each per scan), so the dual update and the restoration of `remaining` touch only those k entries;
//   * the first scan of augmentation `cur` always reads cost row `cur`: it is prefetched into shared memory while the previous
//     augmentation finishes, which takes the L2 round trip off the chain for the (common) single-scan augmentations.
template <int T>
__global__ void __launch_bounds__(T) lsap_kernel(const float* __restrict__ cost_g, int R, int C, float max_dist, int* __restrict__ col4row_out,
                                                 int* __restrict__ over_max, LsapWork w, int state_in_smem, int cost_in_smem) {
    extern __shared__ __align__(16) unsigned char lsap_smem[];
    __shared__ LsapBest s_part[32];
    __shared__ double s_min;
    __shared__ int s_i, s_nrem, s_sink, s_k;
    const int tid = threadIdx.x;
    auto bar = [&]() { if (T == 32) __syncwarp(); else __syncthreads(); };
    const float* cost = cost_g;
    if (state_in_smem) {
        lsap_carve(lsap_smem, R, C, w);
        if (cost_in_smem) {
            float* cs = (float*)(lsap_smem + lsap_state_bytes(R, C));
            for (int i = tid; i < R * C; i += T) cs[i] = cost_g[i];
            cost = cs;
        }
    }
    for (int i = tid; i < R; i += T) { w.u[i] = 0.0; w.col4row[i] = -1; }
    for (int j = tid; j < C; j += T) {
        w.v[j] = 0.0; w.row4col[j] = -1; w.path[j] = -1; w.ep_col[j] = -1; w.remaining[j] = C - 1 - j;
        if (!cost_in_smem) w.rowbuf[j] = cost_g[j];          // row 0
    }
    if (tid == 0) { s_min = 0.0; s_i = 0; s_nrem = C; s_sink = -1; s_k = 0; }
    bar();
    for (int cur = 0; cur < R; ++cur) {
        int k = 0;                                             // scans done in this augmentation (uniform: read from shared state)
        // the next augmentation's first row goes to registers now and to shared memory once this augmentation's scans are over:
        // its L2 round trip hides under the search (T >= C / 4 threads: at most four columns each)
        float nxt[4] = {0.f, 0.f, 0.f, 0.f};
        const bool pf = !cost_in_smem && cur + 1 < R && C <= 4 * T;
        if (pf) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (tid + q * T < C) nxt[q] = __ldg(cost_g + (long long)(cur + 1) * C + tid + q * T);
        }
        while (true) {
            const int i = s_i, nrem = s_nrem;
            const double mv = s_min, ui = w.u[i];
            const float* crow = (k == 0 && !cost_in_smem) ? w.rowbuf : cost + (long long)i * C;
            LsapBest best;
            best.val = CUDART_INF; best.key = 0x7fffffff;
            for (int it = tid; it < nrem; it += T) {
                const int j = w.remaining[it];
                const double r = mv + (double)crow[j] - ui - w.v[j];
                double s = w.ep_col[j] == cur ? w.spc[j] : CUDART_INF;
                if (r < s) { w.path[j] = i; w.spc[j] = r; w.ep_col[j] = cur; s = r; }
                LsapBest c;
                c.val = s; c.key = w.row4col[j] == -1 ? C - 1 - it : C + it;
                if (lsap_better(c, best)) best = c;
            }
            best = lsap_warp_min(best);
            if (T > 32) {
                if ((tid & 31) == 0) s_part[tid >> 5] = best;
                __syncthreads();
                if (tid < 32) {
                    LsapBest b2;
                    if (tid < T / 32) b2 = s_part[tid];
                    else { b2.val = CUDART_INF; b2.key = 0x7fffffff; }
                    best = lsap_warp_min(b2);
                }
            }
            if (tid == 0) {
                w.vis_row[k] = i;                             // SR[i] = true
                s_min = best.val;
                if (!(best.val < CUDART_INF)) {
                    s_sink = -2;                              // infeasible (cannot happen with finite costs)
                } else {
                    const int bit = best.key < C ? C - 1 - best.key : best.key - C;      // position in `remaining`
                    const int j = w.remaining[bit];
                    if (w.row4col[j] == -1) s_sink = j; else s_i = w.row4col[j];
                    w.vis_col[k] = j;                         // SC[j] = true
                    w.vis_bit[k] = bit;
                    w.remaining[bit] = w.remaining[nrem - 1];
                    s_nrem = nrem - 1;
                }
                s_k = k + 1;
            }
            ++k;
            bar();
            if (s_sink != -1) break;
        }
        const int sink = s_sink;
        if (sink < 0) {                                        // infeasible: report and stop
            for (int i = tid; i < R; i += T) { col4row_out[i] = -1; over_max[i] = 1; }
            return;
        }
        const double mv = s_min;
        // dual update over the logged rows / columns only (each row and each column appears once in its log); col4row still holds
        // the assignment from before this augmentation, as in scipy (the path is applied afterwards)
        for (int e = tid; e < k; e += T) {
            const int i = w.vis_row[e];
            if (i != cur) w.u[i] += mv - w.spc[w.col4row[i]];
            const int j = w.vis_col[e];
            w.v[j] -= mv - w.spc[j];
        }
        if (pf) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (tid + q * T < C) w.rowbuf[tid + q * T] = nxt[q];
        } else if (!cost_in_smem && cur + 1 < R) {
            for (int j = tid; j < C; j += T) w.rowbuf[j] = __ldg(cost_g + (long long)(cur + 1) * C + j);
        }
        bar();
        if (tid == 0) {
            w.u[cur] += mv;
            int j = sink;
            while (true) {
                const int i = w.path[j];
                w.row4col[j] = i;
                const int t = w.col4row[i];
                w.col4row[i] = j;
                j = t;
                if (i == cur) break;
            }
            // `remaining` back to C-1..0: only the logged positions were overwritten (in reverse: a position may be logged twice)
            for (int e = k - 1; e >= 0; --e) { const int bit = w.vis_bit[e]; w.remaining[bit] = C - 1 - bit; }
            s_min = 0.0; s_i = cur + 1; s_nrem = C; s_sink = -1; s_k = 0;
        }
        bar();
    }
    for (int i = tid; i < R; i += T) {
        const int j = w.col4row[i];
        col4row_out[i] = j;
        over_max[i] = cost[(long long)i * C + j] > max_dist ? 1 : 0;
    }
}

size_t lsap_work_bytes(int R, int C) { return lsap_state_bytes(R, C) + 64; }
void launch_lsap(const float* cost, int R, int C, float max_dist, int* col4row, int* over_max, void* work, cudaStream_t st) {
    if (R == 0 || C == 0) return;
    YDST_CHECK(R <= C, "launch_lsap needs R <= C (transpose first)");
    LsapWork w;
    lsap_carve((unsigned char*)work, R, C, w);
    static bool attr_set_dev[64] = {false};            // the shared-memory opt-in is a per-device attribute
    int dev = 0;
    YDST_CUDA(cudaGetDevice(&dev));
    bool& attr_set = attr_set_dev[dev & 63];
    if (!attr_set) {
        YDST_CUDA(cudaFuncSetAttribute(lsap_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        YDST_CUDA(cudaFuncSetAttribute(lsap_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        YDST_CUDA(cudaFuncSetAttribute(lsap_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    const size_t state_bytes = lsap_state_bytes(R, C);
    const size_t cost_bytes = (size_t)R * C * sizeof(float);
    const int state_in_smem = state_bytes <= 190 * 1024;
    const int cost_in_smem = state_in_smem && state_bytes + cost_bytes <= 96 * 1024;
    const size_t smem = state_in_smem ? state_bytes + (cost_in_smem ? cost_bytes : 0) : 0;
    if (C <= 128) lsap_kernel<32><<<1, 32, smem, st>>>(cost, R, C, max_dist, col4row, over_max, w, state_in_smem, cost_in_smem);
    else if (C <= 1024) lsap_kernel<256><<<1, 256, smem, st>>>(cost, R, C, max_dist, col4row, over_max, w, state_in_smem, cost_in_smem);
    else lsap_kernel<1024><<<1, 1024, smem, st>>>(cost, R, C, max_dist, col4row, over_max, w, state_in_smem, cost_in_smem);
    YDST_CUDA(cudaGetLastError());
}

}  // namespace ydst
