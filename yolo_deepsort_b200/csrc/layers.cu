// HBM-bound layer kernels for sm_100a: everything in the detector / ReID graphs that is not a dense
// contraction.  All activations use the flat-padded NHWC fp16 layout (common.cuh: Act); every kernel
// moves 16 bytes (8 channels) per thread per access and writes interior pixels only.
#include "layers.cuh"

#include <cstdlib>

namespace ydst {

static inline int cdiv(long long a, int b) { return (int)((a + b - 1) / b); }

// ------------------------------------------------------------------------------------------------
__global__ void u8_to_f32_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, long long n) {
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < n) {
        const uchar4 v = *reinterpret_cast<const uchar4*>(src + i);
        *reinterpret_cast<float4*>(dst + i) = make_float4(v.x / 255.f, v.y / 255.f, v.z / 255.f, v.w / 255.f);
    } else {
        for (long long j = i; j < n; ++j) dst[j] = src[j] / 255.f;
    }
}
void launch_u8_to_f32(const uint8_t* src, float* dst, long long n, cudaStream_t st) {
    u8_to_f32_kernel<<<cdiv((n + 3) / 4, 256), 256, 0, st>>>(src, dst, n);
}

__global__ void nchw_to_nhwc_kernel(const void* __restrict__ src, int is_half, float* __restrict__ dst, int N, int C, int H, int W) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // over N*H*W*C (NHWC order)
    const long long total = (long long)N * C * H * W;
    if (i >= total) return;
    const int c = (int)(i % C);
    long long t = i / C;
    const int x = (int)(t % W); t /= W;
    const int y = (int)(t % H);
    const int n = (int)(t / H);
    const long long s = (((long long)n * C + c) * H + y) * W + x;
    dst[i] = is_half ? __half2float(reinterpret_cast<const __half*>(src)[s]) : reinterpret_cast<const float*>(src)[s];
}
void launch_nchw_to_nhwc(const void* src, int src_is_half, float* dst, int N, int C, int H, int W, cudaStream_t st) {
    const long long total = (long long)N * C * H * W;
    nchw_to_nhwc_kernel<<<cdiv(total, 256), 256, 0, st>>>(src, src_is_half, dst, N, C, H, W);
}

// ------------------------------------------------------------------------------------------------
// First layer: Cin = 3 makes K = 27, far too thin for a tensor-core tile.  One thread computes TWO horizontally adjacent
// output pixels: their 3 x (3 + stride) x 3 input window sits in registers and every (broadcast) shared-memory weight
// vector is used for both pixels, which halves the LDS traffic that bounds this kernel.
template <int STRIDE>
__global__ void __launch_bounds__(128) conv_first_kernel(const float* __restrict__ in, int N, int H, int W, const float* __restrict__ w,
                                                         const float* __restrict__ scale, const float* __restrict__ bias, int cout,
                                                         int act, Act out) {
    extern __shared__ float sw[];          // [27][cout] then scale[cout], bias[cout]
    for (int i = threadIdx.x; i < 27 * cout; i += blockDim.x) sw[i] = w[i];
    float* ssc = sw + 27 * cout;
    float* sbi = ssc + cout;
    for (int i = threadIdx.x; i < cout; i += blockDim.x) { ssc[i] = scale[i]; sbi[i] = bias[i]; }
    __syncthreads();
    constexpr int WC = 3 + STRIDE;         // window columns shared by the two pixels
    const int pairs_x = (out.W + 1) >> 1;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)N * out.H * pairs_x;
    if (idx >= total) return;
    const int xo = (int)(idx % pairs_x) * 2;
    long long t = idx / pairs_x;
    const int yo = (int)(t % out.H);
    const int n = (int)(t / out.H);
    float v[3][WC][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s = 0; s < WC; ++s) {
            const int y = yo * STRIDE - 1 + r, x = xo * STRIDE - 1 + s;
            const bool ok = y >= 0 && y < H && x >= 0 && x < W;
            const float* px = in + (((long long)n * H + (ok ? y : 0)) * W + (ok ? x : 0)) * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) v[r][s][c] = ok ? __ldg(px + c) : 0.f;
        }
    const bool two = xo + 1 < out.W;
    const long long pix = ((long long)n * out.Hp() + yo + 1) * out.Wp() + xo + 1;
    __half* op = out.base + pix * out.ctot + out.coff;
    for (int co = 0; co < cout; co += 8) {
        float acc[2][8];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[0][q] = acc[1][q] = 0.f;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int s = 0; s < 3; ++s)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const int k = (r * 3 + s) * 3 + c;
                    const float4 w0 = *reinterpret_cast<const float4*>(sw + k * cout + co);
                    const float4 w1 = *reinterpret_cast<const float4*>(sw + k * cout + co + 4);
                    const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
                    const float a0 = v[r][s][c], a1 = v[r][s + STRIDE][c];
#pragma unroll
                    for (int q = 0; q < 8; ++q) { acc[0][q] = fmaf(a0, wv[q], acc[0][q]); acc[1][q] = fmaf(a1, wv[q], acc[1][q]); }
                }
#pragma unroll
        for (int px = 0; px < 2; ++px) {
            if (px == 1 && !two) break;
            uint4 pk;
            __half2* h = reinterpret_cast<__half2*>(&pk);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float a = apply_act(fmaf(acc[px][2 * q], ssc[co + 2 * q], sbi[co + 2 * q]), act);
                const float b = apply_act(fmaf(acc[px][2 * q + 1], ssc[co + 2 * q + 1], sbi[co + 2 * q + 1]), act);
                h[q] = __floats2half2_rn(a, b);
            }
            *reinterpret_cast<uint4*>(op + (long long)px * out.ctot + co) = pk;
        }
    }
}
// Tensor-core first layer (stride 1).  K = 3*3*3 = 27 is padded to 32 and the im2col tile of 128 consecutive output pixels is
// built in shared memory by the CTA's 128 threads (one row each, 64-byte rows in the SWIZZLE_64B layout the UMMA descriptor
// names).  Inputs and weights are fp32 in the reference and this layer sets the scale of everything behind it, so both are
// split into hi + lo fp16 parts and three products (hi*hi + lo*hi + hi*lo) accumulate in fp32 TMEM: ~22 significant bits.
__global__ void __launch_bounds__(128) conv_first_tc_kernel(const float* __restrict__ in, int N, int H, int W, const __half* __restrict__ w_hilo,
                                                            const float* __restrict__ scale, const float* __restrict__ bias, int cout, int act,
                                                            Act out) {
    __shared__ __align__(1024) uint8_t sm[2 * 8192 + 2 * 4096];
    __shared__ __align__(8) unsigned long long bar_storage;
    __shared__ uint32_t tmem_slot;
    __shared__ long long rowpix[128];
    __shared__ __align__(16) float s_sb[128];
    const int t = threadIdx.x, warp = t >> 5;
    const uint32_t sA_hi = smem_u32(sm), sA_lo = sA_hi + 8192u, sB_hi = sA_lo + 8192u, sB_lo = sB_hi + 4096u;
    const uint32_t bar = smem_u32(&bar_storage);
    const uint32_t tmem_cols = cout <= 32 ? 32u : 64u;
    // ---- once per CTA: barrier, TMEM, weights, scale/bias (the CTA then strides over the 128-pixel tiles) ----
    if (t == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (warp == 0) { tmem_alloc(smem_u32(&tmem_slot), tmem_cols); tmem_relinquish(); }
    // weights: [2][cout][32] fp16 -> two swizzled [cout][64 B] tiles
    for (int e = t; e < 2 * cout * 4; e += 128) {
        const int part = e / (cout * 4), rem = e - part * cout * 4;
        const int row = rem >> 2, c = rem & 3;
        const uint4 w4 = __ldg(reinterpret_cast<const uint4*>(w_hilo + ((size_t)part * cout + row) * 32 + c * 8));
        *reinterpret_cast<uint4*>(sm + 16384 + part * 4096 + row * 64 + ((c ^ ((row >> 1) & 3)) << 4)) = w4;
    }
    if (t < cout) { s_sb[t] = __ldg(scale + t); s_sb[64 + t] = __ldg(bias + t); }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_slot;
    const uint32_t idesc = make_idesc_f16(128, cout);
    const long long total = (long long)N * H * W;
    const long long ntiles = (total + 127) / 128;
    const uint32_t sw = (uint32_t)((t >> 1) & 3);          // SWIZZLE_64B: 16-byte chunk index ^= address bits [7:8] = (row >> 1) & 3
    const int chunks = cout >> 3, chunk_sh = 31 - __clz(chunks);
    const bool chunk_p2 = (chunks & (chunks - 1)) == 0;
    uint32_t parity = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, parity ^= 1u) {
        // ---- this thread's im2col row ----
        const long long idx = tile * 128 + t;
        const bool live = idx < total;
        int xo = 0, yo = 0, n = 0;
        if (live) {                                        // (32-bit divisions: the launcher checks N*H*W < 2^31; the 64-bit ones cost ~400 instructions per pixel)
            const unsigned u = (unsigned)idx, q = u / (unsigned)W;
            xo = (int)(u - q * (unsigned)W);
            n = (int)(q / (unsigned)H);
            yo = (int)(q - (unsigned)n * (unsigned)H);
        }
        float v[32];
#pragma unroll
        for (int k = 27; k < 32; ++k) v[k] = 0.f;
        {
            const bool rok[3] = {live && yo >= 1, live, live && yo + 1 < H}, cok[3] = {xo >= 1, true, xo + 1 < W};
            const float* ctr = in + (((long long)n * H + yo) * W + xo) * 3;      // never dereferenced unless the tap is inside
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int s = 0; s < 3; ++s) {
                    const bool ok = rok[r] && cok[s];
                    const float* px = ctr + ((r - 1) * W + (s - 1)) * 3;
#pragma unroll
                    for (int c = 0; c < 3; ++c) v[(r * 3 + s) * 3 + c] = ok ? __ldg(px + c) : 0.f;
                }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            uint4 hi4, lo4;
            __half2* hh = reinterpret_cast<__half2*>(&hi4);
            __half2* ll = reinterpret_cast<__half2*>(&lo4);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float a = v[c * 8 + 2 * q], b = v[c * 8 + 2 * q + 1];
                const __half2 h2 = __floats2half2_rn(a, b);
                const float2 back = __half22float2(h2);
                hh[q] = h2;
                ll[q] = __floats2half2_rn(a - back.x, b - back.y);
            }
            const uint32_t off = (uint32_t)t * 64u + ((((uint32_t)c) ^ sw) << 4);
            *reinterpret_cast<uint4*>(sm + off) = hi4;
            *reinterpret_cast<uint4*>(sm + 8192 + off) = lo4;
        }
        rowpix[t] = live ? ((long long)n * out.Hp() + yo + 1) * out.Wp() + xo + 1 : -1;
        fence_proxy_async();                               // generic-proxy smem writes -> visible to the tensor core (async proxy)
        tcgen05_fence_before();
        __syncthreads();
        tcgen05_fence_after();
        if (warp == 0 && elect_one()) {
            const uint32_t a_of[3] = {sA_hi, sA_lo, sA_hi}, b_of[3] = {sB_hi, sB_hi, sB_lo};
#pragma unroll
            for (int pr = 0; pr < 3; ++pr)
#pragma unroll
                for (int k = 0; k < 2; ++k)
                    umma_f16(tmem_base, make_smem_desc(a_of[pr] + 32u * k, 64), make_smem_desc(b_of[pr] + 32u * k, 64), idesc, (uint32_t)((pr | k) != 0));
            umma_commit(bar);
        }
        mbar_wait(bar, parity);
        tcgen05_fence_after();
        // Epilogue through shared memory: a thread owns one output pixel (TMEM lane), but a thread-per-row global store touches 32
        // cache lines per warp instruction.  The fp16 rows go to a swizzled [128][128 B] tile over the dead im2col tiles, and the
        // CTA then writes them out 16 bytes per thread with 8 (cout = 64) or 4 (cout = 32) consecutive threads per pixel.
        {
            uint4* rbase = reinterpret_cast<uint4*>(sm + (size_t)t * 128u);
            const int x7 = t & 7;
            for (int ch = 0; ch < (cout >> 4); ++ch) {
                __syncwarp();
                uint32_t acc[16];
                tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(ch * 16), acc);
                tcgen05_wait_ld();
                uint4 w0, w1;
                __half2* g0 = reinterpret_cast<__half2*>(&w0);
                __half2* g1 = reinterpret_cast<__half2*>(&w1);
                // (the kernel is instruction-bound: the activation kind is tested once per 16 columns, not once per element)
                float o[16];
                const float4* sc4 = reinterpret_cast<const float4*>(s_sb + ch * 16);
                const float4* bi4 = reinterpret_cast<const float4*>(s_sb + 64 + ch * 16);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 sc = sc4[q], bi = bi4[q];
                    o[4 * q + 0] = fmaf(__uint_as_float(acc[4 * q + 0]), sc.x, bi.x);
                    o[4 * q + 1] = fmaf(__uint_as_float(acc[4 * q + 1]), sc.y, bi.y);
                    o[4 * q + 2] = fmaf(__uint_as_float(acc[4 * q + 2]), sc.z, bi.z);
                    o[4 * q + 3] = fmaf(__uint_as_float(acc[4 * q + 3]), sc.w, bi.w);
                }
                if (act == ACT_RELU) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) o[j] = fmaxf(o[j], 0.f);
                } else if (act == ACT_LEAKY) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) o[j] = fmaxf(o[j], 0.1f * o[j]);          // == (x > 0 ? x : 0.1x) for every finite x
                } else if (act != ACT_LINEAR) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) o[j] = apply_act(o[j], act);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    g0[q] = __floats2half2_rn(o[2 * q], o[2 * q + 1]);
                    g1[q] = __floats2half2_rn(o[8 + 2 * q], o[8 + 2 * q + 1]);
                }
                rbase[(2 * ch) ^ x7] = w0;
                rbase[(2 * ch + 1) ^ x7] = w1;
            }
        }
        tcgen05_fence_before();                            // the accumulator has been read: the next tile's MMAs may overwrite it
        __syncthreads();
        for (int item = t; item < 128 * chunks; item += 128) {
            const int row = chunk_p2 ? item >> chunk_sh : item / chunks, c = item - row * chunks;
            const long long pix = rowpix[row];
            if (pix < 0) continue;
            const uint4 v4 = *reinterpret_cast<const uint4*>(sm + (size_t)row * 128u + (size_t)((c ^ (row & 7)) << 4));
            *reinterpret_cast<uint4*>(out.base + pix * out.ctot + out.coff + c * 8) = v4;
        }
        __syncthreads();                                   // the staging tile is the next im2col tile
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) { __syncwarp(); tmem_dealloc(tmem_base, tmem_cols); }
}

__device__ __forceinline__ uint4* act_ptr_w(const Act& a, int n, int y, int x, int c);
// First layer + MaxPool2d(3, 2, 1) in one kernel (the ReID stem, deep_sort/deep/model.py:47-53): the full-resolution activation
// (128x64x64 fp16 per crop, 1 MB) never goes to HBM.  A CTA owns a 4 x 16 tile of POOLED pixels: it stages the 11 x 35 x 3 input
// patch, builds the im2col rows of the 9 x 33 convolution outputs the tile's windows cover (three 128-row MMA tiles, hi/lo split as
// above), rounds them to fp16 into a swizzled staging tile and takes the 3x3 / stride-2 maxima from there.  The results are
// bit-identical to conv_first_tc_kernel followed by maxpool_kernel (max commutes with the monotone fp16 rounding).
static constexpr int kPoolTH = 4, kPoolTW = 16, kConvRH = 2 * kPoolTH + 1, kConvRW = 2 * kPoolTW + 1;      // 9 x 33 conv outputs
static constexpr int kPatchH = kConvRH + 2, kPatchW = kConvRW + 2;                                       // 11 x 35 input pixels
static constexpr int kStemPoolSmem = 3 * 16384 + 8192 + ((kPatchH * kPatchW * 3 * 4 + 15) & ~15) + 128 * 4 + 64 + 1024;
__global__ void __launch_bounds__(128) conv_first_pool_kernel(const float* __restrict__ in, int N, int H, int W, const __half* __restrict__ w_hilo,
                                                              const float* __restrict__ scale, const float* __restrict__ bias, int cout, int act,
                                                              Act out) {
    extern __shared__ uint8_t sm_raw[];
    uint8_t* sm = sm_raw + ((1024u - (smem_u32(sm_raw) & 1023u)) & 1023u);
    const uint32_t sm_u = smem_u32(sm);
    uint8_t* sB = sm + 3 * 16384;                                          // hi tile, lo tile (4 KB each)
    float* patch = reinterpret_cast<float*>(sB + 8192);
    float* s_sb = patch + ((kPatchH * kPatchW * 3 + 3) & ~3);               // scale[64], bias[64]
    unsigned long long* bar_ptr = reinterpret_cast<unsigned long long*>(s_sb + 128);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_ptr + 1);
    const uint32_t bar = smem_u32(bar_ptr);
    const int t = threadIdx.x, warp = t >> 5;
    if (t == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (warp == 0) { tmem_alloc(smem_u32(tmem_slot), 256u); tmem_relinquish(); }
    const int tiles_x = out.W / kPoolTW, tiles_y = out.H / kPoolTH;
    int b = blockIdx.x;
    const int tx = b % tiles_x; b /= tiles_x;
    const int ty = b % tiles_y;
    const int n = b / tiles_y;
    const int py0 = ty * kPoolTH, px0 = tx * kPoolTW;
    const int cy0 = 2 * py0 - 1, cx0 = 2 * px0 - 1;                         // first conv output of the region (may be -1: pool padding)
    // ---- input patch (zero outside the image = the convolution's padding) ----
    for (int e = t; e < kPatchH * kPatchW * 3; e += 128) {
        const int r = e / (kPatchW * 3), rem = e - r * (kPatchW * 3);
        const int c = rem / 3, ch = rem - c * 3;
        const int y = cy0 - 1 + r, x = cx0 - 1 + c;
        patch[e] = (y >= 0 && y < H && x >= 0 && x < W) ? __ldg(in + (((long long)n * H + y) * W + x) * 3 + ch) : 0.f;
    }
    if (t < cout) { s_sb[t] = __ldg(scale + t); s_sb[64 + t] = __ldg(bias + t); }
    // ---- weights: [2][cout][32] fp16 -> two swizzled [cout][64 B] tiles ----
    for (int e = t; e < 2 * cout * 4; e += 128) {
        const int part = e / (cout * 4), rem = e - part * cout * 4;
        const int row = rem >> 2, c = rem & 3;
        const uint4 w4 = __ldg(reinterpret_cast<const uint4*>(w_hilo + ((size_t)part * cout + row) * 32 + c * 8));
        *reinterpret_cast<uint4*>(sB + part * 4096 + row * 64 + ((c ^ ((row >> 1) & 3)) << 4)) = w4;
    }
    __syncthreads();
    // ---- im2col rows of the three MMA tiles (row = conv output r = i*128 + t of the 9 x 33 region) ----
    const uint32_t sw = (uint32_t)((t >> 1) & 3);
    bool inside[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int r = i * 128 + t;
        const bool live = r < kConvRH * kConvRW;
        const int ry = live ? r / kConvRW : 0, rx = live ? r - ry * kConvRW : 0;
        inside[i] = live && cy0 + ry >= 0 && cy0 + ry < H && cx0 + rx >= 0 && cx0 + rx < W;
        float v[32];
#pragma unroll
        for (int k = 27; k < 32; ++k) v[k] = 0.f;
#pragma unroll
        for (int dr = 0; dr < 3; ++dr)
#pragma unroll
            for (int ds = 0; ds < 3; ++ds) {
                const float* px = patch + ((ry + dr) * kPatchW + rx + ds) * 3;
#pragma unroll
                for (int c = 0; c < 3; ++c) v[(dr * 3 + ds) * 3 + c] = live ? px[c] : 0.f;
            }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            uint4 hi4, lo4;
            __half2* hh = reinterpret_cast<__half2*>(&hi4);
            __half2* ll = reinterpret_cast<__half2*>(&lo4);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float a = v[c * 8 + 2 * q], bb = v[c * 8 + 2 * q + 1];
                const __half ah = __float2half_rn(a), bh = __float2half_rn(bb);
                hh[q] = __halves2half2(ah, bh);
                ll[q] = __halves2half2(__float2half_rn(a - __half2float(ah)), __float2half_rn(bb - __half2float(bh)));
            }
            const uint32_t off = (uint32_t)i * 16384u + (uint32_t)t * 64u + ((((uint32_t)c) ^ sw) << 4);
            *reinterpret_cast<uint4*>(sm + off) = hi4;
            *reinterpret_cast<uint4*>(sm + 8192 + off) = lo4;
        }
    }
    fence_proxy_async();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp == 0 && elect_one()) {
        const uint32_t idesc = make_idesc_f16(128, cout);
        const uint32_t bh = sm_u + 3 * 16384, bl = bh + 4096;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const uint32_t ah = sm_u + (uint32_t)i * 16384u, al = ah + 8192u;
            const uint32_t a_of[3] = {ah, al, ah}, b_of[3] = {bh, bh, bl};
#pragma unroll
            for (int pr = 0; pr < 3; ++pr)
#pragma unroll
                for (int k = 0; k < 2; ++k)
                    umma_f16(tmem_base + (uint32_t)(i * 64), make_smem_desc(a_of[pr] + 32u * k, 64), make_smem_desc(b_of[pr] + 32u * k, 64), idesc,
                             (uint32_t)((pr | k) != 0));
        }
        umma_commit(bar);
    }
    mbar_wait(bar, 0);
    tcgen05_fence_after();
    // ---- BN + activation, fp16, into the staging tile [384 rows][128 B] over the (now dead) im2col tiles ----
    const uint32_t ninf2 = 0xFC00FC00u;
#pragma unroll 1
    for (int i = 0; i < 3; ++i) {
        const int r = i * 128 + t;
        uint4* rbase = reinterpret_cast<uint4*>(sm + (size_t)r * 128u);
        const int x7 = r & 7;
        for (int ch = 0; ch < (cout >> 4); ++ch) {
            __syncwarp();
            uint32_t acc[16];
            tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(i * 64 + ch * 16), acc);
            tcgen05_wait_ld();
            uint4 w0 = make_uint4(ninf2, ninf2, ninf2, ninf2), w1 = w0;
            if (inside[i]) {
                __half2* g0 = reinterpret_cast<__half2*>(&w0);
                __half2* g1 = reinterpret_cast<__half2*>(&w1);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int c0 = ch * 16 + 2 * q;
                    const float a = apply_act(fmaf(__uint_as_float(acc[2 * q]), s_sb[c0], s_sb[64 + c0]), act);
                    const float bb = apply_act(fmaf(__uint_as_float(acc[2 * q + 1]), s_sb[c0 + 1], s_sb[64 + c0 + 1]), act);
                    (q < 4 ? g0[q] : g1[q - 4]) = __floats2half2_rn(a, bb);
                }
            }
            rbase[(2 * ch) ^ x7] = w0;
            rbase[(2 * ch + 1) ^ x7] = w1;
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) { __syncwarp(); tmem_dealloc(tmem_base, 256u); }
    // ---- 3x3 / stride-2 maxima: 64 pooled pixels x (cout / 8) 16-byte channel chunks ----
    const int chunks = cout >> 3;
    for (int item = t; item < kPoolTH * kPoolTW * chunks; item += 128) {
        const int pp = item / chunks, ch = item - pp * chunks;
        const int ply = pp / kPoolTW, plx = pp - ply * kPoolTW;
        const __half2 ninf = __float2half2_rn(-INFINITY);
        __half2 m[4] = {ninf, ninf, ninf, ninf};
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const int r = (2 * ply + dy) * kConvRW + 2 * plx + dx;
                const uint4 v = *reinterpret_cast<const uint4*>(sm + (size_t)r * 128u + (size_t)((ch ^ (r & 7)) << 4));
                const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
                for (int q = 0; q < 4; ++q) m[q] = __hmax2(m[q], h[q]);
            }
        uint4 o;
        __half2* ho = reinterpret_cast<__half2*>(&o);
#pragma unroll
        for (int q = 0; q < 4; ++q) ho[q] = m[q];
        const int py = py0 + ply, px = px0 + plx;
        if (py < out.H && px < out.W) *act_ptr_w(out, n, py, px, ch * 8) = o;
    }
}

bool conv_first_pool_supported(int H, int W, int cout, int stride, int act, const __half* w_hilo) {
    // (read per plan, not cached: the parity test builds one extractor with the fused stem and one without)
    // opt-in: measured SLOWER than the two kernels (ReID forward of 416 crops 2.48 ms against 2.35 ms) -- three 128-row tiles per
    // CTA need 192 TMEM columns, so only two 128-thread CTAs fit an SM and their serial phases are latency-bound
    const bool on = (getenv("YDST_STEM_FUSED") && atoi(getenv("YDST_STEM_FUSED")) != 0) &&
                    !(getenv("YDST_STEM_TC") && atoi(getenv("YDST_STEM_TC")) == 0);
    return on && w_hilo && stride == 1 && cout == 64 && act != ACT_MISH && H % (2 * kPoolTH) == 0 && W % (2 * kPoolTW) == 0;
}
void launch_conv_first_pool(const float* in, int N, int H, int W, const __half* w_hilo, const float* scale, const float* bias, int cout, int act,
                            const Act& out, cudaStream_t st) {
    YDST_CHECK(conv_first_pool_supported(H, W, cout, 1, act, w_hilo) && out.H == H / 2 && out.W == W / 2 && out.C == cout,
               "fused first layer + maxpool: unsupported shape");
    static bool attr = false;
    if (!attr) {
        YDST_CUDA(cudaFuncSetAttribute(conv_first_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kStemPoolSmem));
        attr = true;
    }
    const long long ctas = (long long)N * (out.H / kPoolTH) * (out.W / kPoolTW);
    conv_first_pool_kernel<<<(unsigned)ctas, 128, kStemPoolSmem, st>>>(in, N, H, W, w_hilo, scale, bias, cout, act, out);
    YDST_CUDA(cudaGetLastError());
}

void launch_conv_first(const float* in, int N, int H, int W, const float* w, const __half* w_hilo, const float* scale, const float* bias,
                       int cout, int stride, int act, const Act& out, cudaStream_t st) {
    YDST_CHECK(cout % 8 == 0 && cout <= 64, "first-layer conv supports cout in {8..64}, multiple of 8 (got %d)", cout);
    YDST_CHECK(stride == 1 || stride == 2, "first-layer conv supports stride 1 and 2");
    static const bool tc_ok = !(getenv("YDST_STEM_TC") && atoi(getenv("YDST_STEM_TC")) == 0);
    if (tc_ok && w_hilo && stride == 1 && cout % 16 == 0 && act != ACT_MISH) {
        const long long tot = (long long)N * H * W;
        YDST_CHECK(tot < (1LL << 31), "first-layer conv: %lld pixels per launch exceed the 32-bit index arithmetic", tot);
        // persistent: 8 CTAs per SM (26 KB of shared memory, <= 64 TMEM columns each) stride over the tiles
        static const int per_sm = getenv("YDST_STEM_CTAS_PER_SM") ? atoi(getenv("YDST_STEM_CTAS_PER_SM")) : 8;
        const long long ntiles = cdiv(tot, 128), cap = 148LL * (per_sm > 0 ? per_sm : 8);
        conv_first_tc_kernel<<<(unsigned)(per_sm > 0 ? std::min(ntiles, cap) : ntiles), 128, 0, st>>>(in, N, H, W, w_hilo, scale, bias, cout, act, out);
        YDST_CUDA(cudaGetLastError());
        return;
    }
    const long long total = (long long)N * out.H * ((out.W + 1) / 2);
    const int smem = (27 * cout + 2 * cout) * (int)sizeof(float);
    if (stride == 1) conv_first_kernel<1><<<cdiv(total, 128), 128, smem, st>>>(in, N, H, W, w, scale, bias, cout, act, out);
    else conv_first_kernel<2><<<cdiv(total, 128), 128, smem, st>>>(in, N, H, W, w, scale, bias, cout, act, out);
    YDST_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ const uint4* act_ptr(const Act& a, int n, int y, int x, int c) {
    return reinterpret_cast<const uint4*>(a.base + (((long long)n * a.Hp() + y + 1) * a.Wp() + x + 1) * a.ctot + a.coff + c);
}
__device__ __forceinline__ uint4* act_ptr_w(const Act& a, int n, int y, int x, int c) {
    return reinterpret_cast<uint4*>(a.base + (((long long)n * a.Hp() + y + 1) * a.Wp() + x + 1) * a.ctot + a.coff + c);
}
// decode a linear index over (n, y, x, c8) of `a`
__device__ __forceinline__ bool decode_idx(const Act& a, long long idx, int& n, int& y, int& x, int& c) {
    const int cg = a.C >> 3;
    const long long total = (long long)a.N * a.H * a.W * cg;
    if (idx >= total) return false;
    if (total < (1LL << 31)) {                             // (every tensor of the path: 32-bit divisions cost ~20 instructions, 64-bit ones ~100 each)
        const unsigned u = (unsigned)idx, t1 = u / (unsigned)cg, t2 = t1 / (unsigned)a.W, t3 = t2 / (unsigned)a.H;
        c = (int)(u - t1 * (unsigned)cg) * 8;
        x = (int)(t1 - t2 * (unsigned)a.W);
        y = (int)(t2 - t3 * (unsigned)a.H);
        n = (int)t3;
        return true;
    }
    c = (int)(idx % cg) * 8;
    long long t = idx / cg;
    x = (int)(t % a.W); t /= a.W;
    y = (int)(t % a.H);
    n = (int)(t / a.H);
    return true;
}

__global__ void maxpool_kernel(Act in, Act out, int k, int stride, int zero_pad_br) {
    int n, yo, xo, c;
    if (!decode_idx(out, (long long)blockIdx.x * blockDim.x + threadIdx.x, n, yo, xo, c)) return;
    const int pad = zero_pad_br ? 0 : (k - 1) / 2;
    const __half2 ninf = __float2half2_rn(-INFINITY);
    __half2 m[4] = {ninf, ninf, ninf, ninf};
    for (int dy = 0; dy < k; ++dy) {
        const int y = yo * stride - pad + dy;
        for (int dx = 0; dx < k; ++dx) {
            const int x = xo * stride - pad + dx;
            const bool inside = y >= 0 && y < in.H && x >= 0 && x < in.W;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (inside) v = __ldg(act_ptr(in, n, y, x, c));
            else if (!zero_pad_br) continue;                 // -inf padding: ignore; zero padding: a real 0
            const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
            for (int q = 0; q < 4; ++q) m[q] = __hmax2(m[q], h[q]);
        }
    }
    uint4 o;
    __half2* ho = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int q = 0; q < 4; ++q) ho[q] = m[q];
    *act_ptr_w(out, n, yo, xo, c) = o;
}
void launch_maxpool(const Act& in, const Act& out, int k, int stride, int zero_pad_br, cudaStream_t st) {
    YDST_CHECK(in.C == out.C && in.C % 8 == 0, "maxpool channel mismatch");
    const long long total = (long long)out.N * out.H * out.W * (out.C / 8);
    maxpool_kernel<<<cdiv(total, 256), 256, 0, st>>>(in, out, k, stride, zero_pad_br);
    YDST_CUDA(cudaGetLastError());
}

__global__ void upsample_kernel(Act in, Act out, int s) {
    int n, y, x, c;
    if (!decode_idx(out, (long long)blockIdx.x * blockDim.x + threadIdx.x, n, y, x, c)) return;
    *act_ptr_w(out, n, y, x, c) = __ldg(act_ptr(in, n, y / s, x / s, c));
}
void launch_upsample(const Act& in, const Act& out, int s, cudaStream_t st) {
    YDST_CHECK(in.C == out.C && out.H == in.H * s && out.W == in.W * s, "upsample shape mismatch");
    const long long total = (long long)out.N * out.H * out.W * (out.C / 8);
    upsample_kernel<<<cdiv(total, 256), 256, 0, st>>>(in, out, s);
    YDST_CUDA(cudaGetLastError());
}

__global__ void add_kernel(Act a, Act b, Act out) {
    int n, y, x, c;
    if (!decode_idx(out, (long long)blockIdx.x * blockDim.x + threadIdx.x, n, y, x, c)) return;
    const uint4 va = __ldg(act_ptr(a, n, y, x, c)), vb = __ldg(act_ptr(b, n, y, x, c));
    const __half2* ha = reinterpret_cast<const __half2*>(&va);
    const __half2* hb = reinterpret_cast<const __half2*>(&vb);
    uint4 o;
    __half2* ho = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float2 fa = __half22float2(ha[q]), fb = __half22float2(hb[q]);
        ho[q] = __floats2half2_rn(fa.x + fb.x, fa.y + fb.y);
    }
    *act_ptr_w(out, n, y, x, c) = o;
}
void launch_add(const Act& a, const Act& b, const Act& out, cudaStream_t st) {
    YDST_CHECK(a.C == out.C && b.C == out.C && a.H == out.H && b.H == out.H && a.W == out.W && b.W == out.W, "shortcut shape mismatch");
    const long long total = (long long)out.N * out.H * out.W * (out.C / 8);
    add_kernel<<<cdiv(total, 256), 256, 0, st>>>(a, b, out);
    YDST_CUDA(cudaGetLastError());
}

__global__ void copy_kernel(Act in, Act out) {
    int n, y, x, c;
    if (!decode_idx(out, (long long)blockIdx.x * blockDim.x + threadIdx.x, n, y, x, c)) return;
    *act_ptr_w(out, n, y, x, c) = __ldg(act_ptr(in, n, y, x, c));
}
void launch_copy(const Act& in, const Act& out, cudaStream_t st) {
    YDST_CHECK(in.C == out.C && in.H == out.H && in.W == out.W && in.N == out.N, "copy shape mismatch");
    const long long total = (long long)out.N * out.H * out.W * (out.C / 8);
    copy_kernel<<<cdiv(total, 256), 256, 0, st>>>(in, out);
    YDST_CUDA(cudaGetLastError());
}

// dense NHWC fp16 <-> flat-padded view (test / boundary helpers)
__global__ void pack_kernel(const __half* __restrict__ src, Act dst, int to_padded) {
    int n, y, x, c;
    if (!decode_idx(dst, (long long)blockIdx.x * blockDim.x + threadIdx.x, n, y, x, c)) return;
    uint4* dense = reinterpret_cast<uint4*>(const_cast<__half*>(src) + (((long long)n * dst.H + y) * dst.W + x) * dst.C + c);
    if (to_padded) *act_ptr_w(dst, n, y, x, c) = *dense;
    else *dense = *act_ptr(dst, n, y, x, c);
}
void launch_pack(const __half* dense, const Act& padded, cudaStream_t st) {
    const long long total = (long long)padded.N * padded.H * padded.W * (padded.C / 8);
    pack_kernel<<<cdiv(total, 256), 256, 0, st>>>(dense, padded, 1);
    YDST_CUDA(cudaGetLastError());
}
void launch_unpack(const Act& padded, __half* dense, cudaStream_t st) {
    const long long total = (long long)padded.N * padded.H * padded.W * (padded.C / 8);
    pack_kernel<<<cdiv(total, 256), 256, 0, st>>>(dense, padded, 0);
    YDST_CUDA(cudaGetLastError());
}
__global__ void unpack_f32_kernel(const float* __restrict__ src, int cstride, int N, int H, int W, int C, float* __restrict__ dst) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)N * H * W * C) return;
    const int c = (int)(i % C);
    long long t = i / C;
    const int x = (int)(t % W); t /= W;
    const int y = (int)(t % H);
    const int n = (int)(t / H);
    dst[i] = src[(((long long)n * (H + 2) + y + 1) * (W + 2) + x + 1) * cstride + c];
}
void launch_unpack_f32(const float* padded, int cstride, int N, int H, int W, int C, float* dense, cudaStream_t st) {
    unpack_f32_kernel<<<cdiv((long long)N * H * W * C, 256), 256, 0, st>>>(padded, cstride, N, H, W, C, dense);
    YDST_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// block = 32 lanes over the fields of a box x 8 grid cells; grid.x over the cells of one grid row, grid.y = (n * na + a) * gy + y.
// The kernel is instruction-bound (12 M outputs of a 76x76 head at micro-batch 8), so the index arithmetic is per block row,
// not per element: no division by the field count, three fields per thread, 32 consecutive floats per warp load and store.
__global__ void __launch_bounds__(256) yolo_decode_kernel(const float* __restrict__ head, int cstride, int N, int gy, int gx, int na, float aw0,
                                                          float ah0, float aw1, float ah1, float aw2, float ah2, int nc, float s0, float s1,
                                                          float* __restrict__ pred, int rows_total, int row0) {
    const int nf = nc + 5;
    const int x = blockIdx.x * 8 + threadIdx.y;
    if (x >= gx) return;
    int t = blockIdx.y;
    const int y = t % gy; t /= gy;                         // block-uniform
    const int a = t % na;
    const int n = t / na;
    const long long pix = ((long long)n * (gy + 2) + y + 1) * (gx + 2) + x + 1;
    const float* src = head + pix * cstride + a * nf;
    const long long row = row0 + ((long long)a * gy + y) * gx + x;
    float* dst = pred + ((long long)n * rows_total + row) * nf;
    const float aw = a == 0 ? aw0 : (a == 1 ? aw1 : aw2), ah = a == 0 ? ah0 : (a == 1 ? ah1 : ah2);
    for (int k = threadIdx.x; k < nf; k += 32) {
        const float v = __ldg(src + k);
        float o;
        // (sigmoid(t)+grid) * scale, exp(t) * (anchor/scale) * scale with scale = (H/gy, W/gx, H/gy, W/gx):
        // x and w use the HEIGHT stride, y and h the WIDTH stride -- the reference's own quirk.
        if (k == 0) o = (1.f / (1.f + expf(-v)) + (float)x) * s0;
        else if (k == 1) o = (1.f / (1.f + expf(-v)) + (float)y) * s1;
        else if (k == 2) o = (expf(v) * (aw / s0)) * s0;
        else if (k == 3) o = (expf(v) * (ah / s1)) * s1;
        else o = 1.f / (1.f + expf(-v));
        dst[k] = o;
    }
}
void launch_yolo_decode(const float* head, int cstride, int N, int gy, int gx, int na, const float* anchors_wh, int nc, int img_h,
                        int img_w, float* pred, int rows_total, int row0, cudaStream_t st) {
    YDST_CHECK(na == 3, "yolo layer with %d anchors (3 supported)", na);
    const float s0 = (float)((double)img_h / gy), s1 = (float)((double)img_w / gx);
    const dim3 grid((unsigned)cdiv(gx, 8), (unsigned)(N * na * gy));
    yolo_decode_kernel<<<grid, dim3(32, 8), 0, st>>>(head, cstride, N, gy, gx, na, anchors_wh[0], anchors_wh[1], anchors_wh[2], anchors_wh[3],
                                             anchors_wh[4], anchors_wh[5], nc, s0, s1, pred, rows_total, row0);
    YDST_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) avgpool_l2_kernel(Act in, float* __restrict__ out) {
    const int n = blockIdx.x, c = threadIdx.x;       // 512 channels
    float s = 0.f;
    for (int y = 0; y < in.H; ++y)
        for (int x = 0; x < in.W; ++x)
            s += __half2float(in.base[(((long long)n * in.Hp() + y + 1) * in.Wp() + x + 1) * in.ctot + in.coff + c]);
    const float v = s / (float)(in.H * in.W);
    __shared__ float red[16];
    float q = v * v;
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = q;
    __syncthreads();
    float tot = 0.f;
    for (int i = 0; i < 16; ++i) tot += red[i];
    out[(long long)n * 512 + c] = v / sqrtf(tot);
}
void launch_avgpool_l2(const Act& in, float* out, cudaStream_t st) {
    YDST_CHECK(in.C == 512, "ReID tail expects 512 channels");
    avgpool_l2_kernel<<<in.N, 512, 0, st>>>(in, out);
    YDST_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// cv2.resize(u8, INTER_LINEAR) fixed-point model, exact (oracle/cv_resize_ref.py).
__device__ __forceinline__ void axis_coeff(int d, int n_dst, int n_src, bool clamp_frac, int& i0, int& i1, int& w0, int& w1) {
    const float f = (float)(((double)d + 0.5) * ((double)n_src / (double)n_dst) - 0.5);
    int i = (int)floorf(f);
    float fr = f - (float)i;
    if (clamp_frac) {
        if (i < 0) { i = 0; fr = 0.f; }
        if (i >= n_src - 1) { i = n_src - 1; fr = 0.f; }
        i0 = i; i1 = min(i + 1, n_src - 1);
    } else {
        i1 = min(max(i + 1, 0), n_src - 1);
        i0 = min(max(i, 0), n_src - 1);
    }
    w1 = __float2int_rn(fr * 2048.f);
    w0 = __float2int_rn((1.f - fr) * 2048.f);
}

// one output pixel (dx, dy) of crop b of `frame`, written to o[0..2]
__device__ __forceinline__ void crop_resize_pixel(const uint8_t* __restrict__ frame, int H, int W, const float* __restrict__ tlwh, int b, int dx, int dy,
                                                  float* __restrict__ o, int* err_flag) {
    const int DW = 64, DH = 128;
    const float bx = tlwh[b * 4 + 0], by = tlwh[b * 4 + 1], bw = tlwh[b * 4 + 2], bh = tlwh[b * 4 + 3];
    // DeepSort._s_tlwh_to_xyxy: int() truncation toward zero; x+w and y+h are fp32 sums; note the -1
    const int x1 = max((int)bx, 0), x2 = min((int)(bx + bw), W - 1);
    const int y1 = max((int)by, 0), y2 = min((int)(by + bh), H - 1);
    const int sw = x2 - x1, sh = y2 - y1;
    if (sw <= 0 || sh <= 0) {
        if (dx == 0 && dy == 0) atomicExch(err_flag, 1);
        o[0] = o[1] = o[2] = 0.f;
        return;
    }
    int v[3];
    if (sw == DW && sh == DH) {                      // same size: cv2.resize is a copy
        const uint8_t* s = frame + ((long long)(y1 + dy) * W + x1 + dx) * 3;
        v[0] = s[0]; v[1] = s[1]; v[2] = s[2];
    } else {
        int xa, xb, a0, a1, ya, yb, b0, b1;
        axis_coeff(dx, DW, sw, true, xa, xb, a0, a1);
        axis_coeff(dy, DH, sh, false, ya, yb, b0, b1);
        const uint8_t* r0 = frame + ((long long)(y1 + ya) * W + x1) * 3;
        const uint8_t* r1 = frame + ((long long)(y1 + yb) * W + x1) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int h0 = (int)r0[xa * 3 + c] * a0 + (int)r0[xb * 3 + c] * a1;
            const int h1 = (int)r1[xa * 3 + c] * a0 + (int)r1[xb * 3 + c] * a1;
            int r = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
            v[c] = min(max(r, 0), 255);
        }
    }
    const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c] = ((float)v[c] / 255.f - mean[c]) / stdv[c];
}

__global__ void __launch_bounds__(256) crop_resize_kernel(const uint8_t* __restrict__ frame, int H, int W, const float* __restrict__ tlwh,
                                                          int m, float* __restrict__ out, int* err_flag) {
    const int DW = 64, DH = 128;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)m * DH * DW) return;
    crop_resize_pixel(frame, H, W, tlwh, (int)(idx / (DW * DH)), (int)(idx % DW), (int)((idx / DW) % DH), out + idx * 3, err_flag);
}
// the crops of up to 8 frames of a micro-batch in ONE launch (eight 9-us launches of ~50 crops each were latency, not bandwidth)
__global__ void __launch_bounds__(256) crop_resize_multi_kernel(const CropBatch cb, float* __restrict__ out) {
    const int DW = 64, DH = 128;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)cb.start[cb.n] * DH * DW) return;
    const int g = (int)(idx / (DW * DH));                    // crop index within the launch; a block of 256 threads never straddles crops
    int f = 0;
#pragma unroll
    for (int i = 1; i < 8; ++i) f += (i < cb.n && g >= cb.start[i]) ? 1 : 0;
    crop_resize_pixel(cb.frame[f], cb.H[f], cb.W[f], cb.tlwh[f], g - cb.start[f], (int)(idx % DW), (int)((idx / DW) % DH), out + idx * 3, cb.err[f]);
}
// Whole-frame ingest (next to the hot path): the reader thread's BGR->RGB (yolo3/detect/video_detect.py:33-36) and
// ImageDetector's cv2.resize(img, (W, H), INTER_LINEAR) (yolo3/detect/img_detect.py:70) on the device, with the same
// fixed-point arithmetic as the crop kernel above (bit-exact against cv2 on every size tried, tests/test_gpu_ingest.py).
__global__ void __launch_bounds__(256) resize_u8_kernel(const uint8_t* __restrict__ src, int sh, int sw, long long pitch,
                                                        uint8_t* __restrict__ dst, int dh, int dw, int swap_rb) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)dh * dw) return;
    const int dx = (int)(idx % dw), dy = (int)(idx / dw);
    int v[3];
    if (sh == dh && sw == dw) {
        const uint8_t* s = src + (long long)dy * pitch + dx * 3;
        v[0] = s[0]; v[1] = s[1]; v[2] = s[2];
    } else {
        int xa, xb, a0, a1, ya, yb, b0, b1;
        axis_coeff(dx, dw, sw, true, xa, xb, a0, a1);
        axis_coeff(dy, dh, sh, false, ya, yb, b0, b1);
        const uint8_t* r0 = src + (long long)ya * pitch;
        const uint8_t* r1 = src + (long long)yb * pitch;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int h0 = (int)r0[xa * 3 + c] * a0 + (int)r0[xb * 3 + c] * a1;
            const int h1 = (int)r1[xa * 3 + c] * a0 + (int)r1[xb * 3 + c] * a1;
            const int r = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
            v[c] = min(max(r, 0), 255);
        }
    }
    uint8_t* o = dst + idx * 3;
    o[0] = (uint8_t)(swap_rb ? v[2] : v[0]); o[1] = (uint8_t)v[1]; o[2] = (uint8_t)(swap_rb ? v[0] : v[2]);
}
void launch_resize_u8(const uint8_t* src, int sh, int sw, uint8_t* dst, int dh, int dw, int swap_rb, cudaStream_t st, long long pitch) {
    YDST_CHECK(sh > 0 && sw > 0 && dh > 0 && dw > 0, "bad resize geometry");
    resize_u8_kernel<<<cdiv((long long)dh * dw, 256), 256, 0, st>>>(src, sh, sw, pitch > 0 ? pitch : (long long)sw * 3, dst, dh, dw, swap_rb);
    YDST_CUDA(cudaGetLastError());
}

// Sliding-window mode (yolo3/detect/img_detect.py:126-134): per tile, xywh2p1p2 (model_build.py:317-323), resize_boxes (:12-19: the
// python-float ratio multiplies the fp32 box), then the window offset is added -- same fp32 operations in the same order.
__global__ void __launch_bounds__(256) window_boxes_kernel(float* __restrict__ pred, int tiles, int rows, int nf, const float* __restrict__ geo) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)tiles * rows) return;
    const int t = (int)(i / rows);
    float* p = pred + i * nf;
    const float rw = geo[t * 4 + 0], rh = geo[t * 4 + 1], ox = geo[t * 4 + 2], oy = geo[t * 4 + 3];
    const float cx = p[0], cy = p[1], w = p[2], h = p[3];
    const float x1 = cx - w / 2.f, y1 = cy - h / 2.f, x2 = cx + w / 2.f, y2 = cy + h / 2.f;
    p[0] = __fadd_rn(__fmul_rn(x1, rw), ox); p[1] = __fadd_rn(__fmul_rn(y1, rh), oy);
    p[2] = __fadd_rn(__fmul_rn(x2, rw), ox); p[3] = __fadd_rn(__fmul_rn(y2, rh), oy);
}
void launch_window_boxes(float* pred, int tiles, int rows, int nf, const float* geo_dev, cudaStream_t st) {
    window_boxes_kernel<<<cdiv((long long)tiles * rows, 256), 256, 0, st>>>(pred, tiles, rows, nf, geo_dev);
    YDST_CUDA(cudaGetLastError());
}

void launch_crop_resize_multi(const CropBatch& cb, float* out, cudaStream_t st) {
    if (cb.n == 0 || cb.start[cb.n] == 0) return;
    crop_resize_multi_kernel<<<cdiv((long long)cb.start[cb.n] * 128 * 64, 256), 256, 0, st>>>(cb, out);
    YDST_CUDA(cudaGetLastError());
}

void launch_crop_resize(const uint8_t* frame, int H, int W, const float* tlwh, int m, float* out, int* err_flag, cudaStream_t st) {
    if (m == 0) return;
    crop_resize_kernel<<<cdiv((long long)m * 128 * 64, 256), 256, 0, st>>>(frame, H, W, tlwh, m, out, err_flag);
    YDST_CUDA(cudaGetLastError());
}

}  // namespace ydst
