// tcgen05 implicit-GEMM convolution for sm_100a.
//
// Replaces nn.Conv2d + BatchNorm2d(eval) + LeakyReLU/Mish (yolo3/models/models.py:40-56, run at :299)
// and the ReID BasicBlock convs (deep_sort/deep/model.py:5-37) with one kernel:
//
//   D[M = output pixels, N = Cout] = sum over taps (r,s) and channel blocks of A_tap[M, Cin] * W_tap[Cout, Cin]^T
//
//   * A (activations, fp16 NHWC with a physical zero border) is fetched by TMA straight into
//     128B/64B/32B-swizzled shared memory.  Stride-1 convs use the "flat-padded" trick: output pixels
//     are 128 consecutive rows of the padded [pixels, C] matrix and filter tap (r,s) is the SAME matrix
//     shifted by (r-1)*Wp + (s-1) rows, so a 3x3 conv is nine 2-D TMA loads per channel block and no
//     im2col buffer ever exists.  Stride-2 convs read one of four parity sub-lattices of the input
//     through 3-D tensor maps (strides baked into the map), M tile = TH x TW output pixels.
//   * B (weights, fp16 [Cout][tap][Cin], K-major) is fetched by TMA as well.
//   * One thread issues tcgen05.mma (M=128, N=block_n, K=16, fp32 accumulate in TMEM); smem stages are
//     recycled through tcgen05.commit -> mbarrier; a TMA producer warp runs `stages` k-blocks ahead.
//   * Epilogue (4 warps): tcgen05.ld the accumulator rows, y = acc*scale + bias (BN folded in fp32),
//     activation, optional residual add (before or after the activation), fp16 store into a channel
//     slice of the destination buffer (zero-copy route/concat) or fp32 store for the YOLO heads.
//     Border / out-of-range rows are computed but never stored, which keeps the zero border intact.
#include "conv_tc.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

namespace ydst {

static constexpr int kBlockM = 128;
static constexpr int kThreads = 192;   // warps 0-3: epilogue, warp 4: TMA producer, warp 5: MMA issuer + TMEM owner

struct ConvTcMaps {
    CUtensorMap a[4];
    CUtensorMap b;
};

// ask L2 to fetch this CTA's slice of the NEXT layer's weights (they come from HBM: 124 MB of weights do not stay resident)
__device__ __forceinline__ void prefetch_next_weights(const ConvTcParams& p) {
    if (!p.pf_bytes) return;
    const unsigned nctas = gridDim.x * gridDim.y * gridDim.z;
    const unsigned cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    const unsigned chunk = ((p.pf_bytes + nctas - 1) / nctas + 127u) & ~127u;
    const unsigned long long off = (unsigned long long)cta * chunk;
    if (off >= p.pf_bytes) return;
    const unsigned sz = (unsigned)min((unsigned long long)chunk, (unsigned long long)p.pf_bytes - off) & ~15u;
    if (sz) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"((const char*)p.pf_ptr + off), "r"(sz) : "memory");
}
__device__ __forceinline__ void grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_dep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync(int id = 1) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

// The single MMA-issuing thread is the critical path of a batch-1 convolution (hundreds of short k-steps per CTA), so the
// descriptors are not rebuilt per instruction: the high word (SBO, version, swizzle mode) is loop-invariant and the low word
// (start address >> 4 | LBO) advances by plain integer adds.
__device__ __forceinline__ uint32_t desc_hi(uint32_t row_bytes) {
    const uint32_t layout = row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u);
    return ((8u * row_bytes) >> 4) | (1u << 14) | (layout << 29);
}
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ void umma_f16_lh(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}\n"
        ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// ---- cta_group::2 forms (a pair of CTAs on two SMs shares one 256-row MMA; cute/arch/mma_sm100_umma.hpp, copy_sm100_tma.hpp,
// cutlass/arch/barrier.h name the same instructions).  The leader is the even CTA of the cluster: a shared::cluster address with
// bit 24 cleared names the leader's copy of a barrier.
static constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ void umma_f16_lh_2cta(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}\n"
        ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the barrier at the same shared-memory offset in BOTH CTAs once all tcgen05 ops issued so far have completed
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((unsigned short)3)
                 : "memory");
}
// TMA loads into this CTA's shared memory whose bytes are counted on the LEADER's barrier
__device__ __forceinline__ void tma_load_2d_2cta(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d_2cta(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// arrive on the copy of a barrier that lives in CTA `rank` of the cluster (cutlass::arch::ClusterBarrier::arrive(cta_id))
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t rank) {
    asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\tmbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar), "r"(rank) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// lean wait for the hot loops (the one in common.cuh carries a printf and is kept for the epilogue)
__device__ __forceinline__ void mbar_wait_hot(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity))
        if (++spins > (1u << 28)) __trap();          // a protocol bug must abort the launch, never hang the GPU
}

__device__ __forceinline__ bool env_res_early(const ConvTcParams& p) { return !(p.bo_mode & 2); }   // YDST_BO_MODE=2: tuning switch

__device__ __forceinline__ void trace_mark_epi(const ConvTcParams& p, int slot) {      // thread 100: a row that is valid in CTA 0
    if (p.trace && threadIdx.x == 100 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) p.trace[slot] = (unsigned long long)clock64();
}
__device__ __forceinline__ void trace_mark(const ConvTcParams& p, int slot) {
    if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
        p.trace[slot] = (unsigned long long)clock64();
        if (slot == 0 || slot == 7) {
            unsigned long long g;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
            p.trace[8 + (slot == 7)] = g;
        }
    }
    if (p.trace && slot == 0 && blockIdx.x == gridDim.x - 1 && blockIdx.y == gridDim.y - 1 && blockIdx.z == gridDim.z - 1) {
        unsigned long long g;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
        p.trace[10] = g;                                  // when the LAST CTA of the grid started
    }
}

// ---------------------------------------------------------------------------------------------
// Epilogue shared by both kernels.  Warps 0..3 own TMEM lanes 32w..32w+31 = output rows; a thread handles one output pixel.
// Columns are processed in groups of up to 64: four tcgen05.ld issued back to back, one wait, then y = acc*scale + bias
// (BN folded in fp32; scale/bias staged in shared memory), residual add before or after the activation, and one contiguous
// 128-byte fp16 store per pixel and group (fp32 for the YOLO heads).  The residual of group g+1 is fetched while group g is
// computed, and group 0's before the accumulator is even complete.
// ---------------------------------------------------------------------------------------------
// activation over 16 values with the (warp-uniform) kind test hoisted out of the element loop
// kMish selects the kernel instantiation that carries the (long) Mish code: everything else stays small and keeps o[] in registers
template <bool kMish>
__device__ __forceinline__ void act16(float (&o)[16], int act) {
    if (act == ACT_LEAKY) {
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] = fmaxf(o[j], 0.1f * o[j]);      // == (x > 0 ? x : 0.1x) for every finite x
    } else if (act == ACT_RELU) {
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] = fmaxf(o[j], 0.f);
    } else if (kMish && act == ACT_MISH) {
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] = apply_act(o[j], ACT_MISH);
    }
}

template <bool kMish>
__device__ __forceinline__ void compute16(const ConvTcParams& p, const float (&acc)[16], const float* s_scale, const float* s_bias,
                                          const uint4& r0, const uint4& r1, float (&o)[16]) {
    const float4* sc4 = reinterpret_cast<const float4*>(s_scale);       // 16-byte aligned (s_sb and the 16-column offsets are)
    const float4* bi4 = reinterpret_cast<const float4*>(s_bias);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 sc = sc4[q], bi = bi4[q];
        o[4 * q + 0] = fmaf(acc[4 * q + 0], sc.x, bi.x);
        o[4 * q + 1] = fmaf(acc[4 * q + 1], sc.y, bi.y);
        o[4 * q + 2] = fmaf(acc[4 * q + 2], sc.z, bi.z);
        o[4 * q + 3] = fmaf(acc[4 * q + 3], sc.w, bi.w);
    }
    if (p.res_mode) {
        float rs[16];
        const __half2* h0 = reinterpret_cast<const __half2*>(&r0);
        const __half2* h1 = reinterpret_cast<const __half2*>(&r1);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float2 f0 = __half22float2(h0[q]), f1 = __half22float2(h1[q]);
            rs[2 * q] = f0.x; rs[2 * q + 1] = f0.y;
            rs[8 + 2 * q] = f1.x; rs[8 + 2 * q + 1] = f1.y;
        }
        if (p.res_mode == 2) {
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] += rs[j];
        }
        act16<kMish>(o, p.act);
        if (p.res_mode == 1) {
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] += rs[j];
        }
    } else {
        act16<kMish>(o, p.act);
    }
}
__device__ __forceinline__ void pack16(const float (&o)[16], uint4& w0, uint4& w1) {
    __half2* g0 = reinterpret_cast<__half2*>(&w0);
    __half2* g1 = reinterpret_cast<__half2*>(&w1);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        g0[q] = __floats2half2_rn(o[2 * q], o[2 * q + 1]);
        g1[q] = __floats2half2_rn(o[8 + 2 * q], o[8 + 2 * q + 1]);
    }
}
template <bool kMish>
__device__ __forceinline__ void finish16(const ConvTcParams& p, const float (&acc)[16], int c, const float* s_scale, const float* s_bias,
                                         const uint4& r0, const uint4& r1, long long pix) {
    float o[16];
    compute16<kMish>(p, acc, s_scale, s_bias, r0, r1, o);
    if (p.out_f32) {
        float4* op = reinterpret_cast<float4*>(p.out_f32 + pix * p.cout + c);
#pragma unroll
        for (int q = 0; q < 4; ++q) op[q] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
    } else {
        uint4 w0, w1;
        pack16(o, w0, w1);
        uint4* op = reinterpret_cast<uint4*>(p.out + pix * p.out_ctot + p.out_coff + c);
        op[0] = w0;
        op[1] = w1;
    }
}

__device__ __forceinline__ void load_res_group(const ConvTcParams& p, long long pix, int c0, int gw, bool valid, uint4 (&r)[8]) {
    if (!p.res_mode || !valid) return;
    const uint4* rp = reinterpret_cast<const uint4*>(p.res + pix * p.res_ctot + p.res_coff + c0);
#pragma unroll
    for (int q = 0; q < 8; ++q)
        if (q * 8 < gw && c0 + q * 8 < p.cout) r[q] = __ldg(rp + q);
}

// stage this CTA's scale/bias columns in shared memory: s_sb[0..bn) = scale, s_sb[bn..2bn) = bias (epilogue threads only)
__device__ __forceinline__ void stage_scale_bias(const ConvTcParams& p, int n0, float* s_sb, int tid = threadIdx.x, int bar = 1) {
    for (int i = tid; i < 2 * p.block_n; i += 128)
        s_sb[i] = i < p.block_n ? __ldg(p.scale + n0 + i) : __ldg(p.bias + n0 + i - p.block_n);
    epi_bar_sync(bar);
}

template <bool kMish>
__device__ __forceinline__ void epilogue_tile(const ConvTcParams& p, uint32_t tmem_base, int warp, int n0, long long pix, bool valid,
                                              const float* s_sb, uint32_t bar_tmem, uint32_t parity = 0, uint32_t bar_release = 0) {
    const int gw = p.block_n < 64 ? p.block_n : 64;
    const int ngroups = p.block_n / gw;
    uint4 rcur[8], rnext[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) rcur[q] = rnext[q] = make_uint4(0, 0, 0, 0);
    load_res_group(p, pix, n0, gw, valid, rcur);
    mbar_wait(bar_tmem, parity);
    tcgen05_fence_after();
    for (int g = 0; g < ngroups; ++g) {
        const int c0 = n0 + g * gw;
        if (c0 >= p.cout) break;                                   // warp-uniform
        __syncwarp();                                              // reconverge before the .sync.aligned loads
        uint32_t v[4][16];
#pragma unroll
        for (int sub = 0; sub < 4; ++sub)
            if (sub * 16 < gw && c0 + sub * 16 < p.cout)
                tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(g * gw + sub * 16), v[sub]);
        if (g + 1 < ngroups && c0 + gw < p.cout) load_res_group(p, pix, c0 + gw, gw, valid, rnext);
        tcgen05_wait_ld();
        if (valid) {
#pragma unroll
            for (int sub = 0; sub < 4; ++sub)
                if (sub * 16 < gw && c0 + sub * 16 < p.cout) {
                    float acc[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc[j] = __uint_as_float(v[sub][j]);
                    const int cl = g * gw + sub * 16;
                    finish16<kMish>(p, acc, c0 + sub * 16, s_sb + cl, s_sb + p.block_n + cl, rcur[2 * sub], rcur[2 * sub + 1], pix);
                }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) rcur[q] = rnext[q];
    }
    if (bar_release) {                                             // the accumulator has been read: hand it back to the MMA issuer
        tcgen05_fence_before();
        mbar_arrive(bar_release);
    }
}

// fp32 output (the three YOLO head convolutions; flat mode, no residual): the rows of a warp are consecutive pixels, so each
// 16-column block is transposed through a warp-private 2.5 KB patch of the (dead) operand stages and written as 16-byte stores
// that cover 8 rows x 64 contiguous bytes per instruction instead of 32 rows x 16 bytes (32 cache lines per instruction, which
// together with the spills of the generic path made the 76x76 head the slowest launch of the forward: 73 us for 1.5 GFLOP).
template <bool kMish>
__device__ __forceinline__ void epilogue_tile_f32(const ConvTcParams& p, uint32_t tmem_base, int wq, int n0, long long row0, bool valid,
                                                  const float* s_sb, uint32_t bar_tmem, uint32_t parity, float* patch) {
    const int lane = threadIdx.x & 31;
    constexpr int kPitch = 20;                                     // floats per patch row: 16 + 4 keeps float4 rows on distinct banks
    mbar_wait(bar_tmem, parity);
    tcgen05_fence_after();
    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
    const uint32_t tbase = tmem_base + ((uint32_t)(wq * 32) << 16);
    for (int cl = 0; cl < p.block_n; cl += 16) {
        const int c = n0 + cl;
        if (c >= p.cout) break;                                    // warp-uniform
        uint32_t v[16];
        __syncwarp();
        tmem_ld_32x32b_x16(tbase + (uint32_t)cl, v);
        tcgen05_wait_ld();
        float o[16];
        const float4* sc4 = reinterpret_cast<const float4*>(s_sb + cl);
        const float4* bi4 = reinterpret_cast<const float4*>(s_sb + p.block_n + cl);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 sc = sc4[q], bi = bi4[q];
            o[4 * q + 0] = fmaf(__uint_as_float(v[4 * q + 0]), sc.x, bi.x);
            o[4 * q + 1] = fmaf(__uint_as_float(v[4 * q + 1]), sc.y, bi.y);
            o[4 * q + 2] = fmaf(__uint_as_float(v[4 * q + 2]), sc.z, bi.z);
            o[4 * q + 3] = fmaf(__uint_as_float(v[4 * q + 3]), sc.w, bi.w);
        }
        act16<kMish>(o, p.act);
        float4* mine = reinterpret_cast<float4*>(patch + lane * kPitch);
#pragma unroll
        for (int q = 0; q < 4; ++q) mine[q] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int r = it * 8 + (lane >> 2), ch = lane & 3;
            if ((vmask >> r) & 1u) {
                const float4 t = *reinterpret_cast<const float4*>(patch + r * kPitch + ch * 4);
                *reinterpret_cast<float4*>(p.out_f32 + (row0 + r) * p.cout + c + ch * 4) = t;
            }
        }
    }
}

// TMA-store variant (halo kernel, cout % 64 == 0, fp16 output): every thread-per-row global store of the direct epilogue
// touches its own 128-byte line (32 lines per warp instruction), which costs ~2000 clocks per 64-column group in the LSU.
// Here each 64-column group is written to a 128B-swizzled [128 rows][64 cols] staging tile in shared memory (the operand
// stages are dead by then) and one elected thread hands it to the TMA store unit.  Invalid rows (border pixels) are written
// as zeros, which is what the border already holds; rows past the end of the tensor are clipped by TMA.
template <bool kMish>
__device__ __forceinline__ void epilogue_tile_tma(const ConvTcParams& p, const CUtensorMap* out_map, uint32_t stage_smem,
                                                  unsigned char* stage_ptr, uint32_t tmem_base, int warp, int n0, int p0, long long pix,
                                                  bool valid, const float* s_sb, uint32_t bar_tmem, uint32_t parity, uint32_t bar_release,
                                                  int& stores, int tid = threadIdx.x, int bar = 1) {
    const int ngroups = p.block_n >> 6;
    const int row = tid;                                           // 0..127 within the epilogue warpgroup
    uint4 rcur[8], rnext[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) rcur[q] = rnext[q] = make_uint4(0, 0, 0, 0);
    load_res_group(p, pix, n0, 64, valid, rcur);
    mbar_wait(bar_tmem, parity);
    tcgen05_fence_after();
    for (int g = 0; g < ngroups; ++g, ++stores) {
        const int c0 = n0 + g * 64;
        if (c0 >= p.cout) break;
        __syncwarp();
        uint32_t v[4][16];
#pragma unroll
        for (int sub = 0; sub < 4; ++sub)
            tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(g * 64 + sub * 16), v[sub]);
        if (g + 1 < ngroups && c0 + 64 < p.cout) load_res_group(p, pix, c0 + 64, 64, valid, rnext);
        if (stores >= 2) {                                         // the buffer about to be refilled was read two stores ago
            if (tid == 0) tma_store_wait_read<1>();
            epi_bar_sync(bar);
        }
        tcgen05_wait_ld();
        if (bar_release && (g + 1 == ngroups || c0 + 64 >= p.cout)) {   // last read of this accumulator: hand it back to the MMA issuer
            tcgen05_fence_before();
            mbar_arrive(bar_release);
        }
        if (g == 0) trace_mark_epi(p, 11);
        const uint32_t buf = stage_smem + (uint32_t)(stores & 1) * (kBlockM * 128u);
#pragma unroll
        for (int sub = 0; sub < 4; ++sub) {
            uint4 w0 = make_uint4(0, 0, 0, 0), w1 = w0;
            if (valid) {
                float acc[16], o[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[j] = __uint_as_float(v[sub][j]);
                const int cl = g * 64 + sub * 16;
                compute16<kMish>(p, acc, s_sb + cl, s_sb + p.block_n + cl, rcur[2 * sub], rcur[2 * sub + 1], o);
                pack16(o, w0, w1);
            }
            uint4* rbase = reinterpret_cast<uint4*>(stage_ptr + (size_t)(stores & 1) * (kBlockM * 128u) + (size_t)row * 128u);
            const int x = row & 7;
            rbase[(2 * sub) ^ x] = w0;
            rbase[(2 * sub + 1) ^ x] = w1;
            if (g == 0 && sub == 0) trace_mark_epi(p, 15);
        }
        if (g == 0) trace_mark_epi(p, 12);
        fence_proxy_async();                                       // generic-proxy smem writes -> visible to the TMA (async proxy)
        if (g == 0) trace_mark_epi(p, 13);
        epi_bar_sync(bar);
        if (g == 0) trace_mark_epi(p, 14);
        if (tid == 0) {
            tma_store_2d(out_map, buf, p.out_coff + c0, p0);
            tma_store_commit();
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) rcur[q] = rnext[q];
    }
}

// Eight-warp variant of the TMA-store epilogue (one-tile-per-CTA launches of the halo kernel).  Warps w and w+4 share TMEM lane
// quarter w & 3: for every 64-column group, warpgroup `half` handles columns [32*half, 32*half+32) of each row, so two warps per
// scheduler cover each other's latencies and the per-thread work halves.  Every group has its own 16 KB staging buffer (no reuse,
// no waits on earlier stores).  The residual tile is not fetched thread-per-row (32 cache lines per warp instruction) but by TMA
// into the same swizzled staging buffer the result is written to: a thread reads its residual chunks, then overwrites them.
// The arithmetic is the bound here (measured: tcgen05.ld delivers > 300 B/clk/SM, tools/micro/tmem_read.cu), so the activation
// and residual mode are compile-time (no per-block uniform branches / parameter loads) and the fp32 work uses the packed
// FFMA2 / FMUL2 / FADD2 forms, which round exactly like their scalar counterparts.
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1, float c0, float c1) {
    asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}
__device__ __forceinline__ void fmul2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
    asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmul.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void fadd2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
    asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tadd.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}

// y = acc*scale + bias, residual before (kRes 2) or after (kRes 1) the activation, fp16 pack: the compile-time twin of compute16 + pack16
template <int kAct, int kRes>
__device__ __forceinline__ void finish16_static(const uint32_t (&v)[16], const float* s_scale, const float* s_bias, const uint4& r0, const uint4& r1,
                                                uint4& w0, uint4& w1) {
    const float4* sc4 = reinterpret_cast<const float4*>(s_scale);
    const float4* bi4 = reinterpret_cast<const float4*>(s_bias);
    float o[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 sc = sc4[q], bi = bi4[q];
        ffma2(o[4 * q + 0], o[4 * q + 1], __uint_as_float(v[4 * q + 0]), __uint_as_float(v[4 * q + 1]), sc.x, sc.y, bi.x, bi.y);
        ffma2(o[4 * q + 2], o[4 * q + 3], __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]), sc.z, sc.w, bi.z, bi.w);
    }
    float rs[16];
    if (kRes) {
        const __half2* h0 = reinterpret_cast<const __half2*>(&r0);
        const __half2* h1 = reinterpret_cast<const __half2*>(&r1);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float2 f0 = __half22float2(h0[q]), f1 = __half22float2(h1[q]);
            rs[2 * q] = f0.x; rs[2 * q + 1] = f0.y;
            rs[8 + 2 * q] = f1.x; rs[8 + 2 * q + 1] = f1.y;
        }
    }
    if (kRes == 2) {
#pragma unroll
        for (int j = 0; j < 16; j += 2) fadd2(o[j], o[j + 1], o[j], o[j + 1], rs[j], rs[j + 1]);
    }
    if (kAct == ACT_LEAKY) {
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
            float t0, t1;
            fmul2(t0, t1, o[j], o[j + 1], 0.1f, 0.1f);
            o[j] = fmaxf(o[j], t0);                                 // == (x > 0 ? x : 0.1x) for every finite x
            o[j + 1] = fmaxf(o[j + 1], t1);
        }
    } else if (kAct == ACT_RELU) {
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] = fmaxf(o[j], 0.f);
    } else if (kAct == ACT_MISH) {
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] = apply_act(o[j], ACT_MISH);
    }
    if (kRes == 1) {
#pragma unroll
        for (int j = 0; j < 16; j += 2) fadd2(o[j], o[j + 1], o[j], o[j + 1], rs[j], rs[j + 1]);
    }
    pack16(o, w0, w1);
}

// kAct / kRes < 0: activation and residual mode read from the parameters at run time (the rare combinations)
template <bool kMish, int kAct, int kRes>
__device__ __forceinline__ void epilogue_tile_wide(const ConvTcParams& p, const CUtensorMap* out_map, const CUtensorMap* res_map,
                                                   const uint32_t (&gaddr)[4], unsigned char* smem_generic, uint32_t smem_generic_u32,
                                                   uint32_t tmem_base, int wq, int half, int row, int n0, int p0, bool valid,
                                                   const float* s_sb, uint32_t bar_tmem, uint32_t bar_res, bool res_issued) {
    const int ngroups = p.block_n >> 6, bn = p.block_n;
    const bool has_res = kRes < 0 ? p.res_mode != 0 : kRes != 0;
    mbar_wait(bar_tmem, 0);
    tcgen05_fence_after();
    if (has_res && !res_issued && threadIdx.x == 0) {             // the operand stages are dead: fetch every group's residual tile
        for (int g = 0; g < ngroups && n0 + g * 64 < p.cout; ++g) {
            mbar_arrive_expect_tx(bar_res + 8u * g, kBlockM * 128u);
            tma_load_2d(gaddr[g], res_map, bar_res + 8u * g, p.res_coff + n0 + g * 64, p0);
        }
    }
    const int x = row & 7;
    const uint32_t tbase = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(half * 32);
    const uint32_t keep = valid ? 0xFFFFFFFFu : 0u;               // border pixels are stored as zeros (branch-free)
    // one 16-column block: scale/bias/activation/residual, fp16 pack, and the two 16-byte chunks of this row in the swizzled tile
    auto process = [&](const uint32_t (&v)[16], int g, int sub, uint4* rbase) {
        const int ch = 4 * half + 2 * sub;
        const int cl = g * 64 + half * 32 + sub * 16;
        uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0, w0, w1;
        if (has_res) { r0 = rbase[ch ^ x]; r1 = rbase[(ch + 1) ^ x]; }
        if (kAct >= 0) {
            finish16_static<kAct, kRes>(v, s_sb + cl, s_sb + bn + cl, r0, r1, w0, w1);
        } else {
            float acc[16], o[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] = __uint_as_float(v[j]);
            compute16<kMish>(p, acc, s_sb + cl, s_sb + bn + cl, r0, r1, o);
            pack16(o, w0, w1);
        }
        w0.x &= keep; w0.y &= keep; w0.z &= keep; w0.w &= keep;
        w1.x &= keep; w1.y &= keep; w1.z &= keep; w1.w &= keep;
        rbase[ch ^ x] = w0;
        rbase[(ch + 1) ^ x] = w1;
    };
    // the TMEM load of block i+1 is in flight while block i is computed; tcgen05.wait::ld waits for every outstanding load, so the
    // next one is issued right after the wait
    uint32_t va[16], vb[16];
    __syncwarp();
    tmem_ld_32x32b_x16(tbase, va);
#pragma unroll 1
    for (int g = 0; g < ngroups; ++g) {
        const int c0 = n0 + g * 64;
        if (c0 >= p.cout) break;
        uint4* rbase = reinterpret_cast<uint4*>(smem_generic + (gaddr[g] - smem_generic_u32) + (size_t)row * 128u);
        tcgen05_wait_ld();
        __syncwarp();
        tmem_ld_32x32b_x16(tbase + (uint32_t)(g * 64 + 16), vb);
        if (has_res) mbar_wait(bar_res + 8u * g, 0);
        process(va, g, 0, rbase);
        tcgen05_wait_ld();
        if (g + 1 < ngroups && c0 + 64 < p.cout) {
            __syncwarp();
            tmem_ld_32x32b_x16(tbase + (uint32_t)((g + 1) * 64), va);
        }
        process(vb, g, 1, rbase);
        fence_proxy_async();                                       // generic-proxy smem writes -> visible to the TMA (async proxy)
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (threadIdx.x == 0) {
            tma_store_2d(out_map, gaddr[g], p.out_coff + c0, p0);
            tma_store_commit();
        }
    }
}

// Persistent-loop epilogue (conv_tc2_kernel<.., kPers = true, ..>; kMp accumulators per tile, kCta2: the CTA is half of a pair).
// Two teams of four warps alternate tiles (one team when the planner says the MMAs of a tile outlast its epilogue): team e
// drains accumulator buffer e of tiles e, e + 2, ... while the MMA issuer fills the other one.  A thread owns one output row and
// walks the tile's 64-column groups; the TMEM load of the next 16-column block is in flight while the current one is finished
// (scale/bias, activation, residual -- all compile-time -- fp16 pack) and written into a 128B-swizzled 16 KB staging buffer that
// one thread hands to the TMA store unit.  Staging buffers rotate per team (nbuf = 2 or 3):
//   nbuf = 3: after issuing store q the team's thread 0 waits until store q - 1 has been read, which frees buffer (q + 2) % 3;
//             the other threads learn it at the staging barrier of group q + 1, one group before they write that buffer.  With a
//             residual, thread 0 then TMA-loads the residual tile of group q + 2 into the freed buffer (two groups of lead), and
//             the result overwrites the residual in place;
//   nbuf = 2: (no residual, shared memory tight) thread 0 waits for store q - 2 at the top of group q, plus one more barrier.
// kAct / kRes < 0: activation and residual mode read from the parameters at run time (the rare combinations).
template <bool kMish, int kAct, int kRes, int kMp, bool kCta2>
__device__ __forceinline__ void epilogue_persistent(const ConvTcParams& p, const CUtensorMap* out_map, const CUtensorMap* res_map,
                                                    unsigned char* smem_generic, uint32_t smem_generic_u32, uint32_t stage_u32,
                                                    uint32_t tmem_base, int team, float* s_sbt, uint32_t bar_tfull, uint32_t bar_tempty,
                                                    uint32_t bar_res, int cta_rank) {
    const int tid = threadIdx.x & 127, wq = (threadIdx.x >> 5) & 3, row = tid, x = row & 7;
    const int bn = p.block_n, ngroups = bn >> 6, nbuf = p.nbuf, ebar = 1 + team;
    // scheduling units: tiles, or (kCta2) pairs of M tiles shared by the two CTAs of a cluster -- CTA `rank` takes tile 2 * um + rank
    const int nteams = p.nteams;
    const int unit0 = kCta2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, ustride = kCta2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int units_m = kCta2 ? (p.m_tiles + 1) >> 1 : p.m_tiles, total_units = units_m * p.n_tiles;
    const bool has_res = kRes < 0 ? p.res_mode != 0 : kRes != 0;
    const int Wp = p.Wo + 2, HpWp = (p.Ho + 2) * Wp;
    const bool tr = p.trace && p.trace_tiles && blockIdx.x == 0 && tid == 0;
    // team-local group index -> coordinates of its 128 x 64 output block (false: past this CTA's last tile)
    auto coords = [&](int q, int& p0, int& c0) -> bool {
        const int j = q / (ngroups * kMp), r = q - j * (ngroups * kMp);
        const int h = r / ngroups, g = r - h * ngroups;
        const int u = unit0 + (nteams * j + team) * ustride;
        if (u >= total_units) return false;
        const int tn = u / units_m, um = u - tn * units_m;
        const int tm = kCta2 ? 2 * um + cta_rank : um;
        p0 = (tm * kMp + h) * kBlockM; c0 = tn * bn + g * 64;
        return true;
    };
    auto fetch_res = [&](int q) {                                  // thread 0 of the team only
        int p0, c0;
        if (!coords(q, p0, c0)) return;
        const uint32_t b = (uint32_t)(q % nbuf);
        mbar_arrive_expect_tx(bar_res + 8u * b, kBlockM * 128u);
        tma_load_2d(stage_u32 + b * (kBlockM * 128u), res_map, bar_res + 8u * b, p.res_coff + c0, p0);
    };
    if (has_res && tid == 0) { fetch_res(0); fetch_res(1); }
    int q = 0, staged_tn = -1;
    for (int j = 0;; ++j) {
        const int it = nteams * j + team;
        const int u = unit0 + it * ustride;
        if (u >= total_units) break;
        const int tn = u / units_m, um = u - tn * units_m;
        const int tm = kCta2 ? 2 * um + cta_rank : um;
        const int n0 = tn * bn;
        if (tn != staged_tn) {                                      // M runs fastest: the column block (and its scale/bias) rarely changes
            if (staged_tn >= 0) epi_bar_sync(ebar);                 // everyone is done with the previous tile's scale/bias
            stage_scale_bias(p, n0, s_sbt, tid, ebar);
            staged_tn = tn;
        }
        const int ab = it & 1;
#pragma unroll 1
        for (int h = 0; h < kMp; ++h) {                             // the 128-row accumulators of this tile, one after the other
        // (the row's geometry first: it is ready by the time the accumulator is)
        const int p0 = (tm * kMp + h) * kBlockM;
        const long long pp = (long long)p0 + row;
        const int rem = (int)((unsigned)pp % (unsigned)HpWp);          // (P_total < 2^31, checked by the planner: a 64-bit modulo costs ~100 instructions)
        const int y = rem / Wp, xx = rem - y * Wp;
        const bool valid = pp < p.P_total && (p.gemm || (y >= 1 && y <= p.Ho && xx >= 1 && xx <= p.Wo));
        const uint32_t keep = valid ? 0xFFFFFFFFu : 0u;             // border pixels are stored as zeros (branch-free)
        if (h == 0) {
            const uint32_t bar = bar_tfull + 8u * ab, par = ((uint32_t)(it >> 1)) & 1u;
            uint32_t spins = 0;
            while (!mbar_try_wait(bar, par)) { __nanosleep(32); if (++spins > (1u << 26)) __trap(); }
            tcgen05_fence_after();
            if (tr && it < 16) p.trace[16 + it * 8 + 2] = (unsigned long long)clock64();
        }
        const uint32_t tbase = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)((ab * kMp + h) * bn);
#pragma unroll 1
        for (int g = 0; g < ngroups; ++g, ++q) {
            const uint32_t b = (uint32_t)(q % nbuf);
            // the whole 64-column group is read out of TMEM first (64 registers; one CTA per SM leaves room for them), so the
            // accumulator goes back to the MMA issuer ~300 clocks after it was complete instead of after three quarters of the
            // group's arithmetic: with short tiles (1x1, K = 64) that wait was the tile period
            uint32_t v0[16], v1[16], v2[16], v3[16];
            const uint32_t tg = tbase + (uint32_t)(g * 64);
            __syncwarp();
            tmem_ld_32x32b_x16(tg, v0);
            tmem_ld_32x32b_x16(tg + 16u, v1);
            tmem_ld_32x32b_x16(tg + 32u, v2);
            tmem_ld_32x32b_x16(tg + 48u, v3);
            if (nbuf == 2 && q >= 2) {                              // the buffer about to be refilled was handed to the TMA two groups ago
                if (tid == 0) tma_store_wait_read<1>();
                epi_bar_sync(ebar);
            }
            uint4* rbase = reinterpret_cast<uint4*>(smem_generic + (stage_u32 - smem_generic_u32) + (size_t)b * (kBlockM * 128u) + (size_t)row * 128u);
            tcgen05_wait_ld();
            if (g + 1 == ngroups && h == kMp - 1) {                 // last read of this tile's accumulators: hand them back to the MMA issuer
                tcgen05_fence_before();
                if constexpr (kCta2) mbar_arrive_remote(bar_tempty + 8u * ab, 0u);   // the pair's MMAs come from the leader: both CTAs release there
                else mbar_arrive(bar_tempty + 8u * ab);
                if (tr && it < 16) p.trace[16 + it * 8 + 5] = (unsigned long long)clock64();
            }
            if (has_res) {
                const uint32_t bar = bar_res + 8u * b, par = ((uint32_t)(q / nbuf)) & 1u;
                uint32_t spins = 0;
                while (!mbar_try_wait(bar, par)) { if (++spins > (1u << 28)) __trap(); }
            }
            auto process = [&](const uint32_t (&v)[16], int sub) {
                const int cl = g * 64 + sub * 16;
                uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0, w0, w1;
                if (has_res) { r0 = rbase[(2 * sub) ^ x]; r1 = rbase[(2 * sub + 1) ^ x]; }
                if (kAct >= 0) {
                    finish16_static<kAct, kRes>(v, s_sbt + cl, s_sbt + bn + cl, r0, r1, w0, w1);
                } else {
                    float acc[16], o[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) acc[i] = __uint_as_float(v[i]);
                    compute16<kMish>(p, acc, s_sbt + cl, s_sbt + bn + cl, r0, r1, o);
                    pack16(o, w0, w1);
                }
                w0.x &= keep; w0.y &= keep; w0.z &= keep; w0.w &= keep;
                w1.x &= keep; w1.y &= keep; w1.z &= keep; w1.w &= keep;
                rbase[(2 * sub) ^ x] = w0;
                rbase[(2 * sub + 1) ^ x] = w1;
            };
            process(v0, 0);
            if (tr && it < 16 && g == 0 && h == 0 && it >= 8) p.trace[16 + it * 8 + 4] = (unsigned long long)clock64();   // (debug: first block done; tiles 8+ only)
            process(v1, 1);
            process(v2, 2);
            process(v3, 3);
            fence_proxy_async();                                    // generic-proxy smem writes -> visible to the TMA (async proxy)
            if (tr && it < 16 && g == 0 && h == 0) p.trace[16 + it * 8 + 6] = (unsigned long long)clock64();
            epi_bar_sync(ebar);
            if (tr && it < 16 && g == 0 && h == 0) p.trace[16 + it * 8 + 7] = (unsigned long long)clock64();
            if (tid == 0) {
                tma_store_2d(out_map, stage_u32 + b * (kBlockM * 128u), p.out_coff + n0 + g * 64, p0);
                tma_store_commit();
                if (nbuf == 3) {
                    tma_store_wait_read<1>();                       // store q - 1 has been read: buffer (q + 2) % 3 is free
                    if (has_res) fetch_res(q + 2);
                }
            }
        }
        }   // h
        if (tr && it < 16) p.trace[16 + it * 8 + 3] = (unsigned long long)clock64();
    }
    if (tid == 0) tma_store_wait_read<0>();                         // smem must outlive the bulk reads; writes are complete at grid end
}

// ---------------------------------------------------------------------------------------------
// Tap-per-stage kernel: stride-2 convolutions (parity sub-lattice tensor maps) and channel blocks narrower than 64.
// ---------------------------------------------------------------------------------------------
template <bool kMish>
__global__ void __launch_bounds__(kThreads) conv_tc_kernel(const __grid_constant__ ConvTcMaps maps, const ConvTcParams p, const int stages) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

    const uint32_t row_bytes = (uint32_t)p.block_k * 2u;
    const uint32_t a_bytes = kBlockM * row_bytes;
    const uint32_t b_bytes = (uint32_t)p.block_n * row_bytes;
    const uint32_t stage_bytes = (a_bytes + b_bytes + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + (uint32_t)stages * stage_bytes;
    // barriers: full[s] at +8s, empty[s] at +8(stages+s), tmem_full at +16*stages, tmem ptr after it, then scale/bias
    const uint32_t bar_full = bar_base, bar_empty = bar_base + 8u * stages, bar_tmem = bar_base + 16u * stages;
    const uint32_t tmem_slot = bar_tmem + 8u;
    float* s_sb = reinterpret_cast<float*>(smem_raw + (((bar_tmem + 16u + 15u) & ~15u) - smem_u32(smem_raw)));   // 16-byte aligned
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < p.block_n) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(bar_full + 8u * s, 1);
            mbar_init(bar_empty + 8u * s, 1);
        }
        mbar_init(bar_tmem, 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&maps.a[0]);
        tma_prefetch_desc(&maps.b);
    }
    if (warp == 5) {
        tmem_alloc(tmem_slot, tmem_cols);
        tmem_relinquish();
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    grid_dep_launch();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    // ---- tile coordinates ----
    const int n0 = blockIdx.y * p.block_n;
    int p0 = 0, img = 0, yo0 = 0, xo0 = 0;
    if (p.mode == 0) {
        p0 = blockIdx.x * kBlockM;
    } else {
        int t = blockIdx.x;
        const int tx = t % p.tiles_x; t /= p.tiles_x;
        const int ty = t % p.tiles_y; img = (t / p.tiles_y) * p.TN;     // a tile covers TN whole images when they are small
        yo0 = ty * p.TH; xo0 = tx * p.TW;
    }
    const int num_kb = p.R * p.S * p.cin_blocks;

    // (both issuing warps run warp-uniform loops; only TMA / expect_tx / tcgen05.mma / tcgen05.commit come from the elected lane:
    //  see conv_tc2_kernel)
    if (warp == 4) {
        const bool leader = elect_one();
        {
            // ================= TMA producer =================
            grid_dep_wait();
            int s = 0, cb = 0, r = 0, sx = 0;
            uint32_t ph = 1;
            const int cin_blocks = p.cin_blocks, block_k = p.block_k, S = p.S, mode = p.mode;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait_hot(bar_empty + 8u * s, ph);
                const uint32_t full = bar_full + 8u * s;
                const int c0 = cb * block_k;
                const uint32_t a_dst = smem_base + (uint32_t)s * stage_bytes;
                if (leader) {
                    mbar_arrive_expect_tx(full, a_bytes + b_bytes);
                    if (mode == 0) {
                        const int row = p0 + (r - p.R / 2) * p.in_Wp + (sx - S / 2);
                        tma_load_2d(a_dst, &maps.a[0], full, c0, row);
                    } else {
                        const int Y = r + p.pad_shift, X = sx + p.pad_shift;
                        tma_load_4d(a_dst, &maps.a[(Y & 1) * 2 + (X & 1)], full, c0, xo0 + (X >> 1), yo0 + (Y >> 1), img);
                    }
                    tma_load_2d(a_dst + a_bytes, &maps.b, full, (r * S + sx) * p.cin + c0, n0);
                }
                if (++cb == cin_blocks) { cb = 0; if (++sx == S) { sx = 0; ++r; } }
                if (++s == stages) { s = 0; ph ^= 1u; }
            }
            if (leader) prefetch_next_weights(p);
        }
    } else if (warp == 5) {
        const bool leader = elect_one();
        {
            // ================= MMA issuer =================
            const uint32_t idesc = make_idesc_f16(kBlockM, p.block_n);
            const uint32_t hi = desc_hi(row_bytes);
            const int ksteps = p.block_k >> 4;
            int s = 0;
            uint32_t ph = 0, acc = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait_hot(bar_full + 8u * s, ph);
                tcgen05_fence_after();
                const uint32_t a_lo = desc_lo(smem_base + (uint32_t)s * stage_bytes);
                const uint32_t b_lo = a_lo + (a_bytes >> 4);
                if (leader) {
                    for (int k = 0; k < ksteps; ++k) {
                        umma_f16_lh(tmem_base, a_lo + 2u * k, b_lo + 2u * k, hi, idesc, acc);
                        acc = 1;
                    }
                    umma_commit(bar_empty + 8u * s);     // frees the smem stage once these MMAs retire
                }
                acc = 1;
                if (++s == stages) { s = 0; ph ^= 1u; }
            }
            if (leader) umma_commit(bar_tmem);           // accumulator complete
        }
    } else {
        // ================= epilogue =================
        stage_scale_bias(p, n0, s_sb);
        const int row = warp * 32 + lane;
        long long pix = 0;
        bool valid;
        if (p.mode == 0) {
            const long long pp = (long long)p0 + row;
            const int Wp = p.Wo + 2, HpWp = (p.Ho + 2) * Wp;
            const int rem = (int)((unsigned)pp % (unsigned)HpWp);          // (P_total < 2^31, checked by the planner: a 64-bit modulo costs ~100 instructions)
            const int y = rem / Wp, x = rem - y * Wp;
            valid = pp < p.P_total && y >= 1 && y <= p.Ho && x >= 1 && x <= p.Wo;
            pix = pp;
        } else {
            const int per = p.TH * p.TW;
            const int ln = row / per, rr = row - ln * per;
            const int ly = rr / p.TW, lx = rr - ly * p.TW;
            const int yo = yo0 + ly, xo = xo0 + lx;
            valid = img + ln < p.N && yo < p.Ho && xo < p.Wo;
            pix = ((long long)(img + ln) * (p.Ho + 2) + yo + 1) * (p.Wo + 2) + xo + 1;
        }
        grid_dep_wait();
        // (measured: staging the rows in shared memory for coalesced 16-byte stores is SLOWER here than the thread-per-row stores --
        //  +5..10 % per layer from the two extra barriers per group -- unlike the TMA bulk stores of the halo kernel)
        epilogue_tile<kMish>(p, tmem_base, warp, n0, pix, valid, s_sb, bar_tmem);
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 5) {
        __syncwarp();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

// ---------------------------------------------------------------------------------------------
// Persistent form of the tap-per-stage kernel, for layers with many short tiles (the 304x304 front of Darknet: 5 800 tiles of
// 18 MMAs each, where a CTA's ~10 000 clocks of setup, first-load latency and epilogue bought ~900 clocks of tensor work).  One
// CTA per SM strides over the tiles: the (A tap, B tap) stage ring runs across tiles, TMEM holds two accumulators, two epilogue
// teams of four warps alternate tiles (direct stores: the stride-2 tiles are TH x TW patches, not runs of rows).  When the N
// tile's whole weight slab fits beside the ring it is loaded once per column block (weight-stationary) and the stages only
// carry activations.
// ---------------------------------------------------------------------------------------------
static constexpr int kThreadsP = 320;   // warps 0-7: two epilogue teams, warp 8: TMA producer, warp 9: MMA issuer + TMEM owner
template <bool kMish>
__global__ void __launch_bounds__(kThreadsP, 1) conv_tc_pers_kernel(const __grid_constant__ ConvTcMaps maps, const ConvTcParams p, const int stages) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const uint32_t row_bytes = (uint32_t)p.block_k * 2u;
    const uint32_t a_bytes = kBlockM * row_bytes;
    const uint32_t b_bytes = (uint32_t)p.block_n * row_bytes;
    const int num_kb = p.R * p.S * p.cin_blocks;
    const bool resident = p.b_resident != 0;                   // the weight slab [num_kb][block_n rows] sits in front of the ring
    const uint32_t slab_bytes = resident ? (((uint32_t)num_kb * b_bytes + 1023u) & ~1023u) : 0u;
    const uint32_t stage_bytes = ((resident ? a_bytes : a_bytes + b_bytes) + 1023u) & ~1023u;
    const uint32_t ring_base = smem_base + slab_bytes;
    const uint32_t bar_base = ring_base + (uint32_t)stages * stage_bytes;
    const uint32_t bar_full = bar_base, bar_empty = bar_base + 8u * stages, bar_tfull = bar_base + 16u * stages, bar_tempty = bar_tfull + 16u;
    const uint32_t bar_slab = bar_tempty + 16u, bar_slab_free = bar_slab + 8u, tmem_slot = bar_slab_free + 8u;
    float* s_sb = reinterpret_cast<float*>(smem_raw + (((tmem_slot + 8u + 15u) & ~15u) - smem_u32(smem_raw)));   // [team][2 * block_n]
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < 2 * p.block_n) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(bar_full + 8u * s, 1); mbar_init(bar_empty + 8u * s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(bar_tfull + 8u * s, 1); mbar_init(bar_tempty + 8u * s, 128); }
        mbar_init(bar_slab, 1); mbar_init(bar_slab_free, 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 8 && lane == 0) { tma_prefetch_desc(&maps.a[0]); tma_prefetch_desc(&maps.b); }
    if (warp == 9) { tmem_alloc(tmem_slot, tmem_cols); tmem_relinquish(); }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    grid_dep_launch();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    const int m_tiles = p.m_tiles, total_tiles = p.m_tiles * p.n_tiles, G = (int)gridDim.x;
    // tile t -> (tm, tn), M fastest; flat mode: 128 padded-pixel rows; patch mode: TN images x TH x TW output pixels
    auto tile_geom = [&](int tm, int& p0, int& img, int& yo0, int& xo0) {
        p0 = 0; img = 0; yo0 = 0; xo0 = 0;
        if (p.mode == 0) { p0 = tm * kBlockM; return; }
        int t = tm;
        const int tx = t % p.tiles_x; t /= p.tiles_x;
        const int ty = t % p.tiles_y; img = (t / p.tiles_y) * p.TN;
        yo0 = ty * p.TH; xo0 = tx * p.TW;
    };

    if (warp == 8) {
        const bool leader = elect_one();
        // ================= TMA producer =================
        grid_dep_wait();
        int s = 0, loaded_tn = -1;
        uint32_t ph = 1, ph_slab = 1;
        const int cin_blocks = p.cin_blocks, block_k = p.block_k, S = p.S, mode = p.mode;
        int tm = (int)blockIdx.x % m_tiles, tn = (int)blockIdx.x / m_tiles;
        for (int t = (int)blockIdx.x; t < total_tiles; t += G) {
            int p0, img, yo0, xo0;
            tile_geom(tm, p0, img, yo0, xo0);
            const int n0 = tn * p.block_n;
            if (resident && tn != loaded_tn) {                 // new column block: wait until the MMAs of the old slab are done, reload
                mbar_wait_hot(bar_slab_free, ph_slab);
                ph_slab ^= 1u;
                if (leader) {
                    mbar_arrive_expect_tx(bar_slab, (uint32_t)num_kb * b_bytes);
                    int cb = 0, tap = 0;
                    for (int kb = 0; kb < num_kb; ++kb) {
                        tma_load_2d(smem_base + (uint32_t)kb * b_bytes, &maps.b, bar_slab, tap * p.cin + cb * block_k, n0);
                        if (++cb == cin_blocks) { cb = 0; ++tap; }
                    }
                }
                loaded_tn = tn;
            }
            int cb = 0, r = 0, sx = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait_hot(bar_empty + 8u * s, ph);
                const uint32_t full = bar_full + 8u * s;
                const int c0 = cb * block_k;
                const uint32_t a_dst = ring_base + (uint32_t)s * stage_bytes;
                if (leader) {
                    mbar_arrive_expect_tx(full, resident ? a_bytes : a_bytes + b_bytes);
                    if (mode == 0) {
                        const int row = p0 + (r - p.R / 2) * p.in_Wp + (sx - S / 2);
                        tma_load_2d(a_dst, &maps.a[0], full, c0, row);
                    } else {
                        const int Y = r + p.pad_shift, X = sx + p.pad_shift;
                        tma_load_4d(a_dst, &maps.a[(Y & 1) * 2 + (X & 1)], full, c0, xo0 + (X >> 1), yo0 + (Y >> 1), img);
                    }
                    if (!resident) tma_load_2d(a_dst + a_bytes, &maps.b, full, (r * S + sx) * p.cin + c0, n0);
                }
                if (++cb == cin_blocks) { cb = 0; if (++sx == S) { sx = 0; ++r; } }
                if (++s == stages) { s = 0; ph ^= 1u; }
            }
            tm += G; while (tm >= m_tiles) { tm -= m_tiles; ++tn; }
        }
        if (leader) prefetch_next_weights(p);
    } else if (warp == 9) {
        const bool leader = elect_one();
        // ================= MMA issuer =================
        const uint32_t idesc = make_idesc_f16(kBlockM, p.block_n);
        const uint32_t hi = desc_hi(row_bytes);
        const int ksteps = p.block_k >> 4, bn = p.block_n;
        int s = 0, it = 0, loaded_tn = -1;
        uint32_t ph = 0, ph_slab = 0;
        int tm = (int)blockIdx.x % m_tiles, tn = (int)blockIdx.x / m_tiles;
        for (int t = (int)blockIdx.x; t < total_tiles; t += G, ++it) {
            const int ab = it & 1;
            const uint32_t tmem_d = tmem_base + (uint32_t)(ab * bn);
            int tm_next = tm + G, tn_next = tn;
            while (tm_next >= m_tiles) { tm_next -= m_tiles; ++tn_next; }
            if (resident && tn != loaded_tn) { mbar_wait_hot(bar_slab, ph_slab); ph_slab ^= 1u; loaded_tn = tn; }
            mbar_wait_hot(bar_tempty + 8u * ab, (((uint32_t)(it >> 1)) & 1u) ^ 1u);
            tcgen05_fence_after();
            uint32_t acc = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait_hot(bar_full + 8u * s, ph);
                tcgen05_fence_after();
                const uint32_t a_lo = desc_lo(ring_base + (uint32_t)s * stage_bytes);
                const uint32_t b_lo = resident ? desc_lo(smem_base + (uint32_t)kb * b_bytes) : a_lo + (a_bytes >> 4);
                if (leader) {
                    for (int k = 0; k < ksteps; ++k) {
                        umma_f16_lh(tmem_d, a_lo + 2u * k, b_lo + 2u * k, hi, idesc, acc);
                        acc = 1;
                    }
                    umma_commit(bar_empty + 8u * s);
                }
                acc = 1;
                if (++s == stages) { s = 0; ph ^= 1u; }
            }
            if (leader) {
                umma_commit(bar_tfull + 8u * ab);
                // the slab may be overwritten once the MMAs of the LAST tile that uses it have completed
                if (resident && (t + G >= total_tiles || tn_next != tn)) umma_commit(bar_slab_free);
            }
            tm = tm_next; tn = tn_next;
        }
    } else {
        // ================= epilogue: two teams alternate tiles =================
        const int team = warp >> 2, wq = warp & 3, tid = threadIdx.x & 127, ebar = 1 + team;
        float* s_sbt = s_sb + team * 256;                      // block_n <= 128: 2 * block_n floats per team
        grid_dep_wait();
        int staged_tn = -1;
        for (int it = team;; it += 2) {
            const int t = (int)blockIdx.x + it * G;
            if (t >= total_tiles) break;
            const int tn = t / m_tiles, tm = t - tn * m_tiles;
            const int n0 = tn * p.block_n;
            int p0, img, yo0, xo0;
            tile_geom(tm, p0, img, yo0, xo0);
            if (tn != staged_tn) {
                if (staged_tn >= 0) epi_bar_sync(ebar);
                stage_scale_bias(p, n0, s_sbt, tid, ebar);
                staged_tn = tn;
            }
            const int row = tid;
            long long pix;
            bool valid;
            if (p.mode == 0) {
                const long long pp = (long long)p0 + row;
                const int Wp = p.Wo + 2, HpWp = (p.Ho + 2) * Wp;
                const int rem = (int)((unsigned)pp % (unsigned)HpWp);          // (P_total < 2^31, checked by the planner: a 64-bit modulo costs ~100 instructions)
                const int y = rem / Wp, x = rem - y * Wp;
                valid = pp < p.P_total && y >= 1 && y <= p.Ho && x >= 1 && x <= p.Wo;
                pix = pp;
            } else {
                const int per = p.TH * p.TW;
                const int ln = row / per, rr = row - ln * per;
                const int ly = rr / p.TW, lx = rr - ly * p.TW;
                const int yo = yo0 + ly, xo = xo0 + lx;
                valid = img + ln < p.N && yo < p.Ho && xo < p.Wo;
                pix = ((long long)(img + ln) * (p.Ho + 2) + yo + 1) * (p.Wo + 2) + xo + 1;
            }
            const int ab = it & 1;
            epilogue_tile<kMish>(p, tmem_base + (uint32_t)(ab * p.block_n), wq, n0, pix, valid, s_sbt, bar_tfull + 8u * ab, ((uint32_t)(it >> 1)) & 1u,
                                 bar_tempty + 8u * ab);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 9) { __syncwarp(); tmem_dealloc(tmem_base, tmem_cols); }
}

// ---------------------------------------------------------------------------------------------
// Halo kernel (stride 1, 64-channel blocks).
//   * the activation chunk of one channel block -- the tile's 128 padded-pixel rows plus (Wp + 1) halo rows on each
//     side -- is fetched ONCE and all nine taps read it through row-shifted UMMA descriptors (tap (r,s) starts
//     r*Wp + s rows into the chunk), so A traffic drops from 9x to (128 + 2Wp + 2)/128 x;
//   * weights stream through their own ring of [block_n x 64] tiles, one per (tap, channel block);
//   * small-M layers fill the SMs by splitting K over channel blocks (grid.z) instead of shrinking the N tile; the
//     fp32 partials meet in a workspace and the last-arriving CTA of each tile sums them IN SPLIT ORDER (deterministic)
//     and runs the epilogue;
//   * griddepcontrol lets the next conv's prologue (barrier init, TMEM alloc, descriptor prefetch) overlap this one's
//     tail when launched with programmatic stream serialization.
// Row-shifted operand starts: measured on B200 (tests/test_gpu_conv.py under YDST_BO_MODE=0/1), the 128B-swizzle XOR is taken
// from the absolute shared-memory address bits -- exactly what TMA used when it wrote the chunk -- so a descriptor may start on
// any 128-byte row of a 1024-byte-aligned buffer with base_offset = 0 (mode 0, the default).  Setting base_offset to
// (addr >> 7) & 7 (mode 1) double-applies the shift and produces garbage; the knob stays as a hardware-behaviour probe.
// ---------------------------------------------------------------------------------------------
// Persistent instantiations run TWO epilogue warpgroups (warps 0-3 and 4-7): tile i is drained by group i & 1, which also owns
// accumulator buffer i & 1, so the epilogues of consecutive tiles overlap each other as well as the main loop.
static constexpr int kThreads2 = 320;   // warps 0-7: epilogue (two per TMEM lane quarter), warp 8: TMA producer, warp 9: MMA issuer + TMEM owner
// kCta2: launched as clusters of two CTAs along M.  The pair computes a 256 x block_n tile with ONE tcgen05.mma.cta_group::2 stream
// issued by the leader (even) CTA: each CTA stages the activation chunk of its own 128 rows and only HALF of the weight tile
// (block_n / 2 output channels), so the weight bytes an SM has to pull out of L2 -- what bounds the 3x3 layers, whose 128-row
// tiles re-stream up to 2.3 MB of weights each against ~42 B/clk/SM of L2 bandwidth -- halve, and so do the tensor core's
// shared-memory reads of B.  Each CTA keeps its own 128 x block_n accumulator in its own TMEM and runs the usual epilogue.
//   * both producers fill their own stages; every load completes on the LEADER's "full" barrier (which expects both CTAs' bytes);
//   * the leader's commits arrive on the "empty" / "accumulator complete" barriers of BOTH CTAs (multicast);
//   * TMEM is allocated / freed with the cta_group::2 forms by the same warp of both CTAs; cluster barriers bracket the kernel
//     (barrier init visible to the peer before the first remote arrive; nobody exits while the peer can still touch its memory).
template <bool kMish, bool kPers, bool kPair, bool kCta2 = false>
__global__ void __launch_bounds__(kThreads2, kPers ? 1 : 2) conv_tc2_kernel(const __grid_constant__ ConvTcMaps maps, const ConvTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // (warp-uniform for the compiler too)
    constexpr int kEpiGroups = 2;                              // epilogue warpgroups of 4 warps
    constexpr int kProdWarp = 4 * kEpiGroups, kMmaWarp = kProdWarp + 1;
    if (threadIdx.x == 0) trace_mark(p, 0);

    const bool k3 = p.R == 3;
    // A stage: 3x3 -> one halo chunk (a_boxes x a_box_rows rows of 128 B); 1x1 -> tpb tiles of 128 rows
    constexpr int mp = kPair ? 2 : 1;                          // 128-row accumulators per tile (compile time: the issue loops unroll)
    const uint32_t a_tile_bytes = (uint32_t)(kBlockM * mp) * 128u;
    const uint32_t a_stage_bytes = k3 ? (((uint32_t)(p.a_box_rows * p.a_boxes) * 128u + 1023u) & ~1023u) : (uint32_t)p.tpb * a_tile_bytes;
    uint32_t cta_rank = 0;
    if constexpr (kCta2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
    const bool lead_cta = cta_rank == 0;
    const uint32_t b_tile_bytes = (uint32_t)(kCta2 ? p.block_n >> 1 : p.block_n) * 128u;   // rows of the weight tile THIS CTA stages
    const uint32_t b_stage_bytes = (uint32_t)p.tpb * b_tile_bytes;
    // persistent mode keeps a dedicated 2 x 16 KB staging area for the TMA-store epilogue in front of the operand stages
    // (they are being refilled for the next tile while the epilogue runs); otherwise the staging aliases the dead stages
    const uint32_t stage_area = kPers ? (uint32_t)(p.nteams * p.nbuf) * kBlockM * 128u : 0u;   // nbuf 16 KB buffers per epilogue team
    const uint32_t a_base = smem_base + stage_area;
    const uint32_t b_base = a_base + (uint32_t)p.a_stages * a_stage_bytes;
    // the wide epilogue stages one 16 KB buffer per 64-column group over the (then dead) operand stages: keep the barriers clear of it
    const bool wide = !kPers && !kPair && p.store_tma;         // (the planner only sets store_tma on unsplit launches)
    const uint32_t stage_need = wide ? (uint32_t)(p.block_n >> 6) * (kBlockM * 128u) : (p.store_tma ? 2u * kBlockM * 128u : 0u);
    const uint32_t bar_base = max(b_base + (uint32_t)p.b_stages * b_stage_bytes, smem_base + stage_need);
    const int nab = k3 ? p.a_boxes : 1;                        // "full" barriers per A stage: one per TMA box of the halo chunk
    const uint32_t bar_fullA = bar_base, bar_emptyA = bar_fullA + 8u * p.a_stages * nab;
    const uint32_t bar_fullB = bar_emptyA + 8u * p.a_stages, bar_emptyB = bar_fullB + 8u * p.b_stages;
    const uint32_t bar_tfull = bar_emptyB + 8u * p.b_stages;   // [2] accumulator complete
    const uint32_t bar_tempty = bar_tfull + 16u;               // [2] accumulator drained by the epilogue (128 arrivals)
    const uint32_t bar_res = bar_tempty + 16u;                 // [8] residual tile of a 64-column group has landed (wide: [group]; persistent: [team][buffer])
    // Wide epilogue with a residual: the tiles are fetched by the PRODUCER into the weight stages as the last MMAs release them
    // (up to b_stages - 1 taps before the accumulator is complete), and the results are staged in place.  Group g lives in the
    // (g / gps)-th stage to be released after the last weight load, gps = 16 KB tiles per stage.
    const uint32_t gps = b_stage_bytes / (kBlockM * 128u);
    const bool res_early = !kCta2 && wide && p.res_mode && gps >= 1u && (uint32_t)(p.block_n >> 6) <= gps * (uint32_t)p.b_stages && env_res_early(p);
    // (every role that needs the group addresses computes them itself: the divisions stay off the common prologue)
    auto group_addrs = [&](uint32_t (&ga)[4], int first_stage) {
        int st = first_stage;
        uint32_t k = 0;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            ga[g] = res_early ? b_base + (uint32_t)st * b_stage_bytes + k * (kBlockM * 128u) : smem_base + (uint32_t)g * (kBlockM * 128u);
            if (++k == gps) { k = 0; if (++st == p.b_stages) st = 0; }
        }
    };
    const uint32_t tmem_slot = bar_res + 64u, flag_slot = tmem_slot + 4u;
    float* s_sb = reinterpret_cast<float*>(smem_raw + (((tmem_slot + 8u + 15u) & ~15u) - smem_u32(smem_raw)));   // 16-byte aligned
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < p.block_n * mp * (kPers ? 2 : 1)) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.a_stages * nab; ++s) mbar_init(bar_fullA + 8u * s, 1);
        for (int s = 0; s < p.a_stages; ++s) mbar_init(bar_emptyA + 8u * s, 1);
        for (int s = 0; s < p.b_stages; ++s) { mbar_init(bar_fullB + 8u * s, 1); mbar_init(bar_emptyB + 8u * s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(bar_tfull + 8u * s, 1); mbar_init(bar_tempty + 8u * s, kCta2 ? 256 : 128); }
        for (int s = 0; s < 8; ++s) mbar_init(bar_res + 8u * s, 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == kProdWarp && elect_one()) {
        tma_prefetch_desc(&maps.a[0]);
        tma_prefetch_desc(&maps.b);
        if (p.store_tma) tma_prefetch_desc(&maps.a[1]);
        if ((wide || kPers) && p.res_mode) tma_prefetch_desc(&maps.a[2]);
    }
    if (warp == kMmaWarp) {
        if constexpr (kCta2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(tmem_cols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            tmem_alloc(tmem_slot, tmem_cols);
            tmem_relinquish();
        }
    }
    tcgen05_fence_before();
    if constexpr (kCta2) cluster_sync_all(); else __syncthreads();
    tcgen05_fence_after();
    if (threadIdx.x == 0) trace_mark(p, 1);
    grid_dep_launch();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    const int cb0 = blockIdx.z * p.cbs_per_split;
    const int ncb = min(p.cbs_per_split, p.cin_blocks - cb0);
    // macro step = one A stage: a channel block with all its taps (3x3) or a group of tpb channel blocks (1x1)
    const int nmacro = k3 ? ncb : (ncb + p.tpb - 1) / p.tpb;
    const int nbs = k3 ? 9 / p.tpb : 1;                        // weight stages per macro step
    // tile iteration: a normal launch owns tile (blockIdx.x, blockIdx.y); a persistent CTA strides over all tiles, M fastest
    const int total_tiles = p.m_tiles * p.n_tiles;
    // persistent scheduling units: tiles, or (kCta2) pairs of M tiles shared by the two CTAs of a cluster
    const int unit0 = kCta2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, ustride = kCta2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int units_m = kCta2 ? (p.m_tiles + 1) >> 1 : p.m_tiles, total_units = units_m * p.n_tiles;
    auto tile_at = [&](int it, int& tm, int& tn) -> bool {
        if (!kPers) { tm = blockIdx.x; tn = blockIdx.y; return it == 0; }
        const int t = blockIdx.x + it * gridDim.x;
        if (t >= total_tiles) return false;
        tm = t % p.m_tiles; tn = t / p.m_tiles;
        return true;
    };

    // The two issuing warps run their loops with ALL lanes (warp-uniform control flow and operands, so ring indices, phases and
    // descriptor words live in uniform registers and feed UTCHMMA / UTMALDG directly); only the instructions that must come from one
    // thread -- TMA, expect_tx, tcgen05.mma, tcgen05.commit -- are guarded by the elected lane.  With the whole loop inside
    // `if (elect_one())` every descriptor word went through R2UR and the issue loop cost ~70 clk per MMA and ~900 clk of scalar
    // bookkeeping per tile (measured, YDST_CONV_TRACE=1): more than an N = 64 MMA takes on the tensor pipe.
    if (warp == kProdWarp) {
        const bool leader = elect_one();
        {
            // ================= TMA producer =================
            int sa = 0, sb = 0, loaded_tn = -1;
            uint32_t pha = 1, phb = 1;
            const int a_stages = p.a_stages, b_stages = p.b_stages, tpb = p.tpb, a_boxes = p.a_boxes, a_box_rows = p.a_box_rows;
            const int total_b = nmacro * nbs;
            int um = kPers ? unit0 % units_m : (int)blockIdx.x, tn = kPers ? unit0 / units_m : (int)blockIdx.y;
            for (int it = 0, u = unit0; kPers ? u < total_units : it == 0; ++it, u += ustride) {
                const int tm = (kPers && kCta2) ? 2 * um + (int)cta_rank : um;
                const int p0 = tm * kBlockM * mp, n0 = tn * p.block_n;
                int jm = 0, js = 0;                            // next weight stage to issue: macro step jm, sub-stage js
                auto issue_b = [&]() {
                    mbar_wait_hot(bar_emptyB + 8u * sb, phb);
                    const uint32_t full = bar_fullB + 8u * sb;
                    if (leader) {
                        if constexpr (kCta2) {
                            // this CTA's half of the weight tile, counted on the leader's barrier (which expects both halves)
                            const int nh = n0 + (int)cta_rank * (p.block_n >> 1);
                            if (lead_cta) mbar_arrive_expect_tx(full, 2u * b_stage_bytes);
                            if (k3) tma_load_3d_2cta(b_base + (uint32_t)sb * b_stage_bytes, &maps.b, full, (cb0 + jm) * 64, nh, js * tpb);
                            else tma_load_3d_2cta(b_base + (uint32_t)sb * b_stage_bytes, &maps.b, full, 0, nh, cb0 + jm * tpb);
                        } else {
                            mbar_arrive_expect_tx(full, b_stage_bytes);
                            if (k3) tma_load_3d(b_base + (uint32_t)sb * b_stage_bytes, &maps.b, full, (cb0 + jm) * 64, n0, js * tpb);
                            else tma_load_3d(b_base + (uint32_t)sb * b_stage_bytes, &maps.b, full, 0, n0, cb0 + jm * tpb);
                        }
                    }
                    if (++js == nbs) { js = 0; ++jm; }
                    if (++sb == b_stages) { sb = 0; phb ^= 1u; }
                };
                auto load_a = [&](int i) {
                    mbar_wait_hot(bar_emptyA + 8u * sa, pha);
                    const uint32_t full = bar_fullA + 8u * (uint32_t)(sa * nab);
                    const uint32_t dst = a_base + (uint32_t)sa * a_stage_bytes;
                    if (leader) {
                        if constexpr (kCta2) {
                            if (k3) {
                                for (int b = 0; b < a_boxes; ++b) {
                                    if (lead_cta) mbar_arrive_expect_tx(full + 8u * b, 2u * (uint32_t)a_box_rows * 128u);
                                    tma_load_2d_2cta(dst + (uint32_t)(b * a_box_rows) * 128u, &maps.a[0], full + 8u * b, (cb0 + i) * 64,
                                                     p0 - p.halo + b * a_box_rows);
                                }
                            } else {
                                if (lead_cta) mbar_arrive_expect_tx(full, 2u * a_stage_bytes);
                                tma_load_3d_2cta(dst, &maps.a[0], full, 0, p0, cb0 + i * tpb);
                            }
                        } else if (k3) {
                            // one barrier per box: the first filter row only needs the first box, so the MMAs start a box earlier
                            for (int b = 0; b < a_boxes; ++b) {
                                mbar_arrive_expect_tx(full + 8u * b, (uint32_t)a_box_rows * 128u);
                                tma_load_2d(dst + (uint32_t)(b * a_box_rows) * 128u, &maps.a[0], full + 8u * b, (cb0 + i) * 64,
                                            p0 - p.halo + b * a_box_rows);
                            }
                        } else {
                            mbar_arrive_expect_tx(full, a_stage_bytes);
                            tma_load_3d(dst, &maps.a[0], full, 0, p0, cb0 + i * tpb);
                        }
                    }
                    if (++sa == a_stages) { sa = 0; pha ^= 1u; }
                };
                int issued = 0;
                if (kPers && p.b_resident) {
                    // weight-stationary: the N tile's whole weight slab (total_b stages == b_stages) is fetched when the column block
                    // changes and then serves every M tile this CTA processes; only the activation chunks stream
                    if (tn != loaded_tn) {
                        for (; issued < total_b; ++issued) issue_b();      // waits for the previous slab's last readers (emptyB)
                        loaded_tn = tn;
                    }
                    issued = total_b;
                }
                if (it == 0) {
                    // weights never depend on the previous kernel: queue them before waiting on the grid dependency, so that
                    // under programmatic dependent launch they stream in while the producer of our input is still draining
                    const int pre = p.pdl ? b_stages : 1;
                    for (; issued < total_b && issued < pre; ++issued) issue_b();
                    grid_dep_wait();                           // the activations are the previous kernel's output
                    if (leader) trace_mark(p, 2);
                }
                load_a(0);
                for (int i = 0; i < nmacro; ++i) {
                    if (a_stages > 1 && i + 1 < nmacro) load_a(i + 1);
                    for (; issued < (i + 1) * nbs; ++issued) issue_b();
                    if (a_stages == 1 && i + 1 < nmacro) load_a(i + 1);
                }
                if (kPers) { um += ustride; while (um >= units_m) { um -= units_m; ++tn; } }
            }
            if (res_early) {
                // sb / phb point at the next stage the ring would refill, i.e. the next one the MMAs release
                const int ngroups = p.block_n >> 6, n0 = (int)blockIdx.y * p.block_n, p0 = (int)blockIdx.x * kBlockM;
                uint32_t gaddr[4];
                group_addrs(gaddr, sb);
                for (int g = 0; g < ngroups && n0 + g * 64 < p.cout; ++g) {
                    if (g % (int)gps == 0) {
                        if (g) { if (++sb == b_stages) { sb = 0; phb ^= 1u; } }
                        mbar_wait_hot(bar_emptyB + 8u * sb, phb);
                    }
                    if (leader) {
                        mbar_arrive_expect_tx(bar_res + 8u * g, kBlockM * 128u);
                        tma_load_2d(gaddr[g], &maps.a[2], bar_res + 8u * g, p.res_coff + n0 + g * 64, p0);
                    }
                }
            }
            if (leader) prefetch_next_weights(p);
        }
    } else if (warp == kMmaWarp) {
        const bool leader = elect_one();
        if (!kCta2 || lead_cta) {                              // (a pair's MMAs all come from its leader CTA)
            // ================= MMA issuer =================
            const uint32_t idesc = make_idesc_f16(kCta2 ? 2 * kBlockM : kBlockM, p.block_n);
            const uint32_t hi = desc_hi(128);
            const uint32_t b_tile16 = b_tile_bytes >> 4;
            const int a_stages = p.a_stages, b_stages = p.b_stages, tpb = p.tpb, bn = p.block_n;
            const bool bo1 = p.bo_mode == 1;
            auto mma = [&](uint32_t td, uint32_t al, uint32_t bl, uint32_t h, uint32_t accu) {
                if constexpr (kCta2) umma_f16_lh_2cta(td, al, bl, h, idesc, accu); else umma_f16_lh(td, al, bl, h, idesc, accu);
            };
            auto commit = [&](uint32_t bar) { if constexpr (kCta2) umma_commit_2cta(bar); else umma_commit(bar); };
            const bool tr = kPers && p.trace && p.trace_tiles && blockIdx.x == 0;
            // 3x3: filter row r reads chunk rows up to r*Wp + 2 + 127, i.e. needs the chunk's TMA boxes up to index need[r]
            int need[3] = {0, 0, 0};
            if (k3)
                for (int r = 0; r < 3; ++r) need[r] = min(p.a_boxes - 1, (r * p.in_Wp + kBlockM * mp + 1) / p.a_box_rows);
            const uint32_t row_step = (uint32_t)p.in_Wp * 8u - 24u;            // 16-byte units: next filter row
            int sa = 0, sb = 0;
            uint32_t pha = 0, phb = 0;
            const bool resident = kPers && p.b_resident;
            int tn = kPers ? unit0 / units_m : (int)blockIdx.y, um = kPers ? unit0 % units_m : 0;
            int prev_tn = -1;
            for (int it = 0, u = unit0; kPers ? u < total_units : it == 0; ++it, u += ustride) {
                // weight-stationary tiles: wait for the slab only on its first use, release it only after its last use; the ring
                // index restarts at 0 every tile (the slab occupies the whole ring) and the phase flips once per slab, not per tile
                bool first_use = true, last_use = true;
                int um_next = um, tn_next = tn;
                if (kPers) { um_next += ustride; while (um_next >= units_m) { um_next -= units_m; ++tn_next; } }
                if (resident) {
                    first_use = tn != prev_tn;
                    last_use = u + ustride >= total_units || tn_next != tn;
                    prev_tn = tn;
                    sb = 0;
                }
                const int ab = it & 1;                         // accumulator buffer
                const uint32_t tmem_d = tmem_base + (uint32_t)(ab * bn * mp);
                if (kPers) mbar_wait_hot(bar_tempty + 8u * ab, (((uint32_t)(it >> 1)) & 1u) ^ 1u);
                tcgen05_fence_after();
                if (tr && leader && it < 8) p.trace[16 + it * 8 + 4] = (unsigned long long)clock64();
                uint32_t acc = 0;
                for (int i = 0; i < nmacro; ++i) {
                    const uint32_t fullA = bar_fullA + 8u * (uint32_t)(sa * nab);
                    mbar_wait_hot(fullA, pha);
                    if (tr && leader && i == 0 && it < 16) p.trace[16 + it * 8] = (unsigned long long)clock64();
                    const uint32_t a_lo0 = desc_lo(a_base + (uint32_t)sa * a_stage_bytes);
                    if (k3) {
                        uint32_t a_lo = a_lo0;
                        int t_in = 0, box_ready = 0;
                        uint32_t b_lo = 0;
#pragma unroll
                        for (int r = 0; r < 3; ++r, a_lo += row_step) {
                            while (box_ready < need[r]) mbar_wait_hot(fullA + 8u * (uint32_t)(++box_ready), pha);
#pragma unroll
                            for (int sx = 0; sx < 3; ++sx, a_lo += 8u) {
                                if (t_in == 0) {
                                    if (first_use) mbar_wait_hot(bar_fullB + 8u * sb, phb);
                                    tcgen05_fence_after();
                                    if (leader && acc == 0 && it == 0) trace_mark(p, 3);
                                    b_lo = desc_lo(b_base + (uint32_t)sb * b_stage_bytes);
                                }
                                uint32_t hi_a = hi;
                                if (bo1) hi_a |= ((a_lo >> 3) & 7u) << 17;  // base_offset probe (bits 49..51)
                                if (leader) {
#pragma unroll
                                    for (int h = 0; h < mp; ++h) {         // the M tiles of the pair share this weight tile
                                        const uint32_t td = tmem_d + (uint32_t)(h * bn), al = a_lo + (uint32_t)h * (kBlockM * 8u);
                                        mma(td, al, b_lo, hi_a, acc);
                                        mma(td, al + 2u, b_lo + 2u, hi_a, 1u);
                                        mma(td, al + 4u, b_lo + 4u, hi_a, 1u);
                                        mma(td, al + 6u, b_lo + 6u, hi_a, 1u);
                                    }
                                }
                                acc = 1u;
                                b_lo += b_tile16;
                                if (++t_in == tpb) {
                                    t_in = 0;
                                    if (last_use && leader) commit(bar_emptyB + 8u * sb);
                                    if (++sb == b_stages) { sb = 0; if (!resident) phb ^= 1u; }
                                }
                            }
                        }
                    } else {
                        if (first_use) mbar_wait_hot(bar_fullB + 8u * sb, phb);
                        tcgen05_fence_after();
                        if (leader && acc == 0 && it == 0) trace_mark(p, 3);
                        uint32_t a_lo = a_lo0, b_lo = desc_lo(b_base + (uint32_t)sb * b_stage_bytes);
                        const int nsl = min(tpb, ncb - i * tpb);                      // the last group of a split may be short
                        for (int tt = 0; tt < nsl; ++tt, a_lo += a_tile_bytes >> 4, b_lo += b_tile16) {
                            if (leader) {
#pragma unroll
                                for (int h = 0; h < mp; ++h) {
                                    const uint32_t td = tmem_d + (uint32_t)(h * bn), al = a_lo + (uint32_t)h * (kBlockM * 8u);
                                    mma(td, al, b_lo, hi, acc);
                                    mma(td, al + 2u, b_lo + 2u, hi, 1u);
                                    mma(td, al + 4u, b_lo + 4u, hi, 1u);
                                    mma(td, al + 6u, b_lo + 6u, hi, 1u);
                                }
                            }
                            acc = 1u;
                        }
                        if (last_use && leader) commit(bar_emptyB + 8u * sb);
                        if (++sb == b_stages) { sb = 0; if (!resident) phb ^= 1u; }
                    }
                    if (leader) commit(bar_emptyA + 8u * sa);
                    if (++sa == a_stages) { sa = 0; pha ^= 1u; }
                }
                if (leader) commit(bar_tfull + 8u * ab);
                if (tr && leader && it < 16) p.trace[16 + it * 8 + 1] = (unsigned long long)clock64();
                if (resident && last_use) phb ^= 1u;           // the next slab lands in the next phase of every weight barrier
                if (leader && it == 0) trace_mark(p, 4);
                um = um_next; tn = tn_next;
            }
        }
    } else {
        // ================= epilogue =================
        if constexpr (kPers) {
            // persistent tile loop with TMA stores (the planner only makes store_tma launches persistent): two teams alternate tiles
            const int team = warp >> 2;
            grid_dep_wait();                                   // residual / output buffers belong to earlier kernels
#define YDST_PERS(A, R) epilogue_persistent<kMish, A, R, mp, kCta2>(p, &maps.a[1], &maps.a[2], smem_raw, smem_u32(smem_raw),                \
                                                         smem_base + (uint32_t)team * (uint32_t)p.nbuf * (kBlockM * 128u), tmem_base, team, \
                                                         s_sb + team * 512, bar_tfull, bar_tempty, bar_res + 24u * team, (int)cta_rank)
            const int key = p.act * 4 + p.res_mode;            // warp-uniform: one specialised epilogue per common combination
            if (team >= p.nteams) {
                // (wide tiles: one team keeps up with the MMAs and the second team's staging buffers buy another weight stage)
            } else if (kMish) {
                if (key == ACT_MISH * 4 + 0) YDST_PERS(ACT_MISH, 0);
                else if (key == ACT_MISH * 4 + 1) YDST_PERS(ACT_MISH, 1);
                else YDST_PERS(-1, -1);
            } else {
                if (key == ACT_LEAKY * 4 + 0) YDST_PERS(ACT_LEAKY, 0);
                else if (key == ACT_LEAKY * 4 + 1) YDST_PERS(ACT_LEAKY, 1);
                else if (key == ACT_RELU * 4 + 0) YDST_PERS(ACT_RELU, 0);
                else if (key == ACT_RELU * 4 + 2) YDST_PERS(ACT_RELU, 2);
                else YDST_PERS(-1, -1);
            }
#undef YDST_PERS
        } else {
        // one tile per CTA: both warpgroups work on it (wide epilogue), or the second one idles (direct-store, split-K, M-pair tiles)
        constexpr int eg = 0;
        const int half = warp >> 2;
        const int tid = threadIdx.x & 127, ebar = 1 + eg;
        const int wq = warp & 3;                               // TMEM lane quarter of this warp
        const int row = tid;
        float* s_sbg = s_sb + eg * 512;
        const uint32_t stage_g = smem_base + (uint32_t)eg * (2u * kBlockM * 128u);
        const int Wp = p.Wo + 2, HpWp = (p.Ho + 2) * Wp;
        int stores = 0, staged_tn = -1;
        bool first = true;
        uint32_t gaddr[4];
        {
            const int total_b = nmacro * nbs;                  // weight stages issued per tile: the ring position after the last one
            group_addrs(gaddr, res_early ? total_b % p.b_stages : 0);
        }
        int tm, tn;
        for (int it = 0; tile_at(it, tm, tn); ++it) {
            if (!wide && half) break;                          // direct-store and split-K epilogues use one warpgroup
            const int n0 = tn * p.block_n;
            const int ab = it & 1;
            const uint32_t bar_full = bar_tfull + 8u * ab, par = ((uint32_t)(it >> 1)) & 1u;
            if (wide) {
                for (int i = threadIdx.x; i < 2 * p.block_n; i += 256)
                    s_sbg[i] = i < p.block_n ? __ldg(p.scale + n0 + i) : __ldg(p.bias + n0 + i - p.block_n);
                asm volatile("bar.sync 1, 256;" ::: "memory");
            } else if (tn != staged_tn) {                      // M runs fastest: the column block (and its scale/bias) rarely changes
                if (!first) epi_bar_sync(ebar);                // everyone is done with the previous tile's scale/bias
                stage_scale_bias(p, n0, s_sbg, tid, ebar);
                staged_tn = tn;
            }
            if (first) grid_dep_wait();                        // residual / workspace / output buffers belong to earlier kernels
            first = false;
            for (int h = 0; h < mp; ++h) {                     // the 128-row accumulators of this tile, one after the other
            const int p0 = (tm * mp + h) * kBlockM;
            const uint32_t tmem_d = tmem_base + (uint32_t)((ab * mp + h) * p.block_n);
            constexpr uint32_t bar_rel = 0u;                       // (no second tile: nobody waits for the accumulator)
            const long long pp = (long long)p0 + row;
            const int rem = (int)((unsigned)pp % (unsigned)HpWp);          // (P_total < 2^31, checked by the planner: a 64-bit modulo costs ~100 instructions)
            const int y = rem / Wp, x = rem - y * Wp;
            const bool valid = pp < p.P_total && (p.gemm || (y >= 1 && y <= p.Ho && x >= 1 && x <= p.Wo));
            if (p.ksplit == 1) {
                if (it == 0 && h == 0 && threadIdx.x == 0 && p.trace) { mbar_wait(bar_full, par); trace_mark(p, 5); }
                if constexpr (!kPair) {
                    if (wide) {
#define YDST_WIDE(A, R) epilogue_tile_wide<kMish, A, R>(p, &maps.a[1], &maps.a[2], gaddr, smem_raw, smem_u32(smem_raw), tmem_d, wq, half, row, n0, p0, \
                                                        valid, s_sbg, bar_full, bar_res, res_early)
                        const int key = p.act * 4 + p.res_mode;    // warp-uniform: one specialised epilogue per common combination
                        if (kMish) {
                            if (key == ACT_MISH * 4 + 0) YDST_WIDE(ACT_MISH, 0);
                            else if (key == ACT_MISH * 4 + 1) YDST_WIDE(ACT_MISH, 1);
                            else YDST_WIDE(-1, -1);
                        } else {
                            if (key == ACT_LEAKY * 4 + 0) YDST_WIDE(ACT_LEAKY, 0);
                            else if (key == ACT_LEAKY * 4 + 1) YDST_WIDE(ACT_LEAKY, 1);
                            else if (key == ACT_RELU * 4 + 0) YDST_WIDE(ACT_RELU, 0);
                            else if (key == ACT_RELU * 4 + 2) YDST_WIDE(ACT_RELU, 2);
                            else if (key == ACT_LINEAR * 4 + 0) YDST_WIDE(ACT_LINEAR, 0);
                            else YDST_WIDE(-1, -1);
                        }
#undef YDST_WIDE
                    }
                    else if (p.out_f32 && p.mode == 0)          // the operand stages are dead once the accumulator is complete: warp-private patches
                        epilogue_tile_f32<kMish>(p, tmem_d, wq, n0, (long long)p0 + wq * 32, valid, s_sbg, bar_full, par,
                                                 reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw))) + wq * 640);
                    else epilogue_tile<kMish>(p, tmem_d, wq, n0, pp, valid, s_sbg, bar_full, par, bar_rel);
                } else if (p.store_tma)
                    epilogue_tile_tma<kMish>(p, &maps.a[1], stage_g, smem_raw + (stage_g - smem_u32(smem_raw)), tmem_d, wq, n0, p0, pp, valid,
                                             s_sbg, bar_full, par, bar_rel, stores, tid, ebar);
                else epilogue_tile<kMish>(p, tmem_d, wq, n0, pp, valid, s_sbg, bar_full, par, bar_rel);
                if (it == 0 && h == 0 && threadIdx.x == 0) trace_mark(p, 6);
            } else {
                const int bn = p.block_n;
                const long long tile = (long long)blockIdx.y * gridDim.x + blockIdx.x;
                float* wtile = p.ws + tile * p.ksplit * (kBlockM * bn);            // [split][row][bn] fp32
                float* mine = wtile + ((long long)blockIdx.z * kBlockM + row) * bn;
                mbar_wait(bar_full, 0);
                tcgen05_fence_after();
                for (int ch = 0; ch < (bn >> 4); ++ch) {
                    if (n0 + ch * 16 >= p.cout) break;
                    __syncwarp();
                    uint32_t v[16];
                    tmem_ld_32x32b_x16(tmem_d + ((uint32_t)(wq * 32) << 16) + (uint32_t)(ch * 16), v);
                    tcgen05_wait_ld();
                    if (!valid) continue;
                    float4* dst = reinterpret_cast<float4*>(mine + ch * 16);
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        __stcg(dst + q, make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                                                    __uint_as_float(v[4 * q + 3])));
                }
                __threadfence();
                epi_bar_sync();
                if (threadIdx.x == 0) {
                    const int t = atomicAdd(p.tickets + tile, 1);
                    asm volatile("st.shared.b32 [%0], %1;" ::"r"(flag_slot), "r"(t) : "memory");
                }
                epi_bar_sync();
                int ticket;
                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(ticket) : "r"(flag_slot) : "memory");
                if (ticket == p.ksplit - 1) {                              // last arrival: every partial of this tile is visible
                    __threadfence();
                    if (threadIdx.x == 0) p.tickets[tile] = 0;             // self-reset for the next launch
                    if (valid) {
                        const float* rowp = wtile + (long long)row * bn;
                        for (int ch = 0; ch < (bn >> 4); ++ch) {
                            const int c = n0 + ch * 16;
                            if (c >= p.cout) break;
                            uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0;
                            if (p.res_mode) {
                                const uint4* rp = reinterpret_cast<const uint4*>(p.res + pp * p.res_ctot + p.res_coff + c);
                                r0 = __ldg(rp); r1 = __ldg(rp + 1);
                            }
                            float acc[16];
#pragma unroll
                            for (int j = 0; j < 16; ++j) acc[j] = 0.f;
                            for (int z = 0; z < p.ksplit; ++z) {           // fixed order: the sum does not depend on arrival order
                                const float4* src = reinterpret_cast<const float4*>(rowp + (long long)z * kBlockM * bn + ch * 16);
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    const float4 t = __ldcg(src + q);
                                    acc[4 * q] += t.x; acc[4 * q + 1] += t.y; acc[4 * q + 2] += t.z; acc[4 * q + 3] += t.w;
                                }
                            }
                            finish16<kMish>(p, acc, c, s_sbg + ch * 16, s_sbg + bn + ch * 16, r0, r1, pp);
                        }
                    }
                }
            }
            }   // h
        }
        if (p.store_tma && tid == 0) tma_store_wait_read<0>();    // smem must outlive the bulk reads; writes are complete at grid end
        }
    }
    tcgen05_fence_before();
    if constexpr (kCta2) cluster_sync_all(); else __syncthreads();
    if (warp == kMmaWarp) {
        __syncwarp();
        if constexpr (kCta2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
        else tmem_dealloc(tmem_base, tmem_cols);
    }
    if (threadIdx.x == 0) trace_mark(p, 7);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        YDST_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q));
        YDST_CHECK(ptr != nullptr && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available from the driver");
        fn = (EncodeTiledFn)ptr;
    }
    return fn;
}

static void encode(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                   const cuuint32_t* box, int row_bytes) {
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUtensorMapSwizzle sw = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    CUresult r = encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    YDST_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with %d (rank %d dims %llu,%llu box %u,%u)", (int)r, rank,
               (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
}

static void ensure_conv_tc2_attrs();
static int max_active_clusters(bool mish, int smem_bytes);
static int g_num_sms = 0;
static int num_sms() {
    if (!g_num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

// Tiling model for the halo kernel (clocks at ~1.9 GHz; constants measured with YDST_CONV_TRACE on B200).
// A CTA costs  setup (~900) + first-data latency (~2200) + main loop + epilogue,  where the main loop is the slowest of
//   * operand bytes / fill rate -- an SM ingests ~42 B/clk from L2, and no more than (bytes in flight) / (~2100 clk latency);
//   * tensor time: a 128 x bn x 16 MMA takes ~max(16, bn/2) clocks;
//   * barrier round trips of the single issuing thread (~150 clocks per stage).
// Small-M layers may split K over channel blocks (grid.z); the fp32 partials then take a round trip through the workspace.
ConvTiling conv_tc_choose_tiling(int m_tiles128, int cout16, int taps, int cin_blocks, int halo, size_t ws_bytes, int max_tickets, int res_mode,
                                 bool allow_pers) {
    const int kSms = 148;
    const double kFill = 42.0, kLat = 2100.0, kLatPers = 4000.0, kSetup = 900.0, kFirst = 2200.0, kStep = 600.0, kClkPerUs = 1900.0;
    // narrow layers (cout < 64) may use a 64-wide N tile: the weight rows past cout are zero-filled by TMA and the store is clipped
    // at the end of the layer's channel slice, so they too get the eight-warp TMA-store epilogue instead of thread-per-row stores
    const bool tma_store_ok = cout16 % 64 == 0 || cout16 < 64;
    ConvTiling best{};
    best.model_us = 1e30;
    int bn_cap = 32;
    while (bn_cap < cout16 && bn_cap < 256) bn_cap <<= 1;
    if (cout16 < 64 && env_int("YDST_NARROW_TMA", 1)) bn_cap = 64;
    const int force_bn = env_int("YDST_FORCE_BN", 0);          // tuning aid: restrict the N tile (when the layer allows it)
    // measured on B200: neither M pairs nor the persistent tile loop beat the plain one-tile-per-CTA launch yet
    // (1163 / 1198 / 1226 frames/s for pair+persistent / persistent / neither at micro-batch 4), so both are opt-in
    const int allow_pair = env_int("YDST_MPAIR", 0), force_mp = env_int("YDST_FORCE_MPAIR", 0);
    const int cta2_mode = env_int("YDST_CTA2", 1);            // 0 off, 1 3x3 layers where the clock model prefers it, 2 wherever legal (test hook)
    const int pers_mode = env_int("YDST_PERSISTENT", 1);      // 0 off, 1 weight-stationary many-tile layers only, 2 wherever the clock model prefers it
    for (int mp = 1; mp <= 2; ++mp)
    for (int bn = bn_cap; bn >= 32; bn >>= 1) {
        if (force_bn && bn != std::min(force_bn, bn_cap)) continue;
        const int n_tiles = (cout16 + bn - 1) / bn;
        // a pair of M tiles per CTA shares every weight stage (half the weight traffic per output); only worth it when there are
        // tiles to spare, i.e. at least two waves of single tiles
        if (mp == 2 && ((long long)m_tiles128 * n_tiles < 2 * kSms || bn > 128)) continue;
        if (force_mp && mp != force_mp && !(mp == 1 && ((long long)m_tiles128 * n_tiles < 2 * kSms || bn > 128))) continue;   // test hook
        const int m_tiles = (m_tiles128 + mp - 1) / mp;
        int a_rows = kBlockM * mp + 2 * halo;
        { const int boxes = (a_rows + 255) / 256; a_rows = ((a_rows + boxes - 1) / boxes) * boxes; }
        const int chunk_bytes = (a_rows * 128 + 1023) & ~1023;
        if (chunk_bytes > 96 * 1024) continue;
        const long long tiles = (long long)m_tiles * n_tiles;
        int last_ks = -1;
        for (int cps = cin_blocks; cps >= 1; --cps) {
            const int ks = (cin_blocks + cps - 1) / cps;
            if (ks == last_ks) continue;
            last_ks = ks;
            if (ks > 1 && ((size_t)ks * tiles * kBlockM * bn * 4 > ws_bytes || tiles > max_tickets || mp > 1)) continue;
            if (ks > 16) continue;
            const long long ctas = tiles * ks;
            // two CTAs per SM: 227 KB of shared memory less 1 KB of system use per CTA -> 113 KB each
            const int budget_max = env_int("YDST_SMEM_BUDGET_KB", 200) * 1024, budget_2 = env_int("YDST_SMEM_BUDGET2_KB", 113) * 1024;
            const int budget_pers = env_int("YDST_SMEM_BUDGET_PERS_KB", 225) * 1024;   // one CTA per SM: everything the SM has (227 KB less alignment slack)
            const int co_model = env_int("YDST_CO_MODEL", 2);
            // weight-stage granularity: one TMA instruction costs the producer thread ~200 clocks to issue, so boxes below
            // ~16 KB make the PRODUCER the bottleneck (measured: tpb = 1 everywhere cost 15 % end to end); the step term below
            // steers towards multi-slice stages, the first-stage term keeps them from growing without bound
            // N = 256 tiles: one tap is already 32 KB of weights and 512 tensor clocks, so single-tap stages pipeline best
            const bool fine = bn == 256;
            const int tpb_opts3[3] = {fine ? 1 : 9, fine ? 1 : 3, 1}, tpb_opts1[3] = {8, 4, 2};
            for (int oi = 0; oi < 3; ++oi) {
                const int tpb = taps == 9 ? tpb_opts3[oi] : std::min(tpb_opts1[oi], cps);
                if (taps == 1 && oi > 0 && tpb == std::min(tpb_opts1[oi - 1], cps)) continue;     // same as the previous option
                const int nmacro = taps == 9 ? cps : (cps + tpb - 1) / tpb;
                const int nbs = taps == 9 ? 9 / tpb : 1;
                const int total_b = nmacro * nbs;
                const int a_stage = taps == 9 ? chunk_bytes : tpb * kBlockM * mp * 128;
                // c2 = 1: clusters of two CTAs share one cta_group::2 MMA stream -- each stages half of the weight tile (conv_tc2_kernel)
                for (int c2 = 0; c2 <= (cta2_mode ? 1 : 0); ++c2) {
                if (c2 && (mp != 1 || ks != 1 || bn < 128 || !allow_pers || m_tiles128 < 2 || !(tma_store_ok && (cout16 % bn == 0)) ||
                           (cta2_mode == 1 && taps != 9)))
                    continue;
                if (!c2 && cta2_mode == 2 && mp == 1 && ks == 1 && bn >= 128 && allow_pers && m_tiles128 >= 2 && tma_store_ok && cout16 % bn == 0) continue;   // test hook: force pairs
                const int b_stage = (tpb * bn * 128) >> c2;
                const int a_stages_max = std::min(nmacro, taps == 9 ? 2 : 4);
                // pass 0: leave room for a second CTA on the SM; pass 1: whole SM; pass 2: persistent -- one CTA per SM loops over
                // the tiles with two TMEM accumulators, so the epilogue of tile i runs under the main loop of tile i+1
                // (persistent: staging buffers per epilogue team -- three let a residual tile be prefetched and save a barrier per group)
                // A persistent CTA's activation ring runs ACROSS tiles (the next tiles' chunks are in flight while this one computes), so its
                // depth is bound by the load latency, not by the tile's own k-steps: measured ~2500 clk per TMA round trip under load.
                const int a_stages_pers = taps == 9 ? 4 : 8;
                for (int pass = 0; pass < 4; ++pass)
                for (int a_stages = pass >= 2 ? a_stages_pers : a_stages_max; a_stages >= 1; --a_stages) {
                    const bool pers = pass >= 2;
                    if (!pers && a_stages < a_stages_max && (pass != 0 || !co_model)) break;   // fewer activation stages only to fit two CTAs per SM
                    if (pers && a_stages < 2) break;
                    // measured: pairs + the persistent loop make the 3x3 layers tensor-bound in steady state, which pays
                    // from two tiles per SM on (ReID layer4 at 348 tiles: 117 -> 83 us; 76x76 128->256 at 381: 44 -> 25 us; but 38x38
                    // 256->512 at 200 tiles: 23 -> 36 us); between one and two waves the pair helps plain launches
                    if (c2 && cta2_mode == 1 && (pers ? tiles < 2 * kSms : (tiles <= kSms || tiles >= 2 * kSms))) break;
                    // persistent: two epilogue teams alternate tiles; wide tiles whose MMAs outlast an epilogue get by with one, and its
                    // staging buffers buy weight stages instead
                    for (int nt = pers ? 2 : 1; nt >= 1; --nt) {
                    const int nbuf = pass == 2 ? 3 : 2;
                    const bool st_ok = tma_store_ok && bn >= 64 && (cout16 % bn == 0 || cout16 < 64);
                    if (pers && (ks > 1 || tiles <= kSms || bn * mp > 256 || !allow_pers || !pers_mode)) continue;
                    if (pers && (!st_ok || (nbuf == 2 && res_mode))) continue;
                    if (mp == 2 && !pers && !allow_pair) continue;   // (plain launches: pairs measured slower than two co-resident CTAs)
                    const int staging = pers ? nt * nbuf * 16 * 1024 : 0;   // 16 KB store buffers per epilogue team
                    const int budget = (pass == 0 ? budget_2 : pers ? budget_pers : budget_max) - staging;
                    if (pass == 0 && ctas <= kSms) continue;      // one CTA per SM anyway: use the whole shared memory
                    // a persistent CTA prefetches the next tile's operands while the current one computes: two stages of each at least
                    const int a_st = pers ? std::max(2, a_stages) : a_stages;
                    const int fixed_p = a_st * a_stage + 6144;
                    int b_stages = std::min(std::min(8, pers ? 8 : total_b), (budget - fixed_p) / b_stage);
                    // weight-stationary persistent tiles: the N tile's whole weight slab fits and is fetched once per column block
                    const bool resident = pers && (!c2 || env_int("YDST_CTA2_RESIDENT", 1)) && total_b <= 16 && total_b * b_stage <= budget - fixed_p && m_tiles >= 2 * kSms / std::max(1, n_tiles) &&
                                          env_int("YDST_B_RESIDENT", 1);
                    if (resident) b_stages = total_b;
                    if (b_stages < (pers && !resident ? 2 : 1)) continue;
                    if (pers && mp == 2 && resident) continue;    // pairs exist to halve streamed weights; stationary ones gain nothing
                    // Measured: the persistent loop pays where the weights stay in shared memory and a CTA walks many
                    // tiles (ReID layer1: 172 -> 90 us per conv); with streamed weights or few waves it only trades the PDL overlap of
                    // plain launches for its own tail, so the default mode keeps those layers on plain launches.
                    if (pers && pers_mode == 1 && !(resident && (long long)m_tiles128 * n_tiles >= 6 * kSms) && !(c2 && taps == 9)) continue;
                    int smem = fixed_p + b_stages * b_stage + staging;
                    // the TMA-store epilogue stages 16 KB per 64-column group (two buffers in persistent / pair mode) at the start of smem
                    smem = std::max(smem, std::max(2, bn / 64) * 16 * 1024 + 6144);
                    const int occ = (!pers && smem <= budget_2) ? 2 : 1;
                    const double inflight = (double)a_st * a_stage + (double)b_stages * b_stage;
                    const double tiles_per_cta = std::ceil((double)ctas / kSms);
                    // CTAs sharing an SM fill it together: the bytes in flight are those of all co-resident CTAs
                    const double co = co_model ? std::min((double)occ, tiles_per_cta) : 1.0;
                    const double lat = pers ? kLatPers : kLat;    // a persistent CTA's loads queue behind those of every other SM's ring
                    const double rate = std::min(kFill, co * inflight / lat);
                    // the two operand streams are pipelined separately: each is also bound by its own bytes in flight
                    const double rate_a = co * (double)a_st * a_stage / lat, rate_b = co * (double)b_stages * b_stage / lat;
                    const double b_share = resident ? 1.0 / std::max(1.0, std::min(tiles_per_cta, (double)m_tiles)) : 1.0;   // slab amortised over the CTA's M tiles
                    const double bytes = (double)cps * ((taps == 9 ? a_rows * 128.0 : kBlockM * mp * 128.0) + b_share * taps * (bn >> c2) * 128);
                    // one 128 x bn x 16 MMA takes bn/2 tensor clocks but also reads 4 KB of A and 32*bn B of B from shared memory
                    // (128 B/clk): narrow N tiles are shared-memory-bound (bn = 32: 40 clk instead of 16, bn = 64: 48 instead of 32)
                    const double mma_step = std::max(bn / 2.0, (4096.0 + 32.0 * (bn >> c2)) / 128.0);
                    const double mma = (double)cps * taps * 4 * mp * mma_step;
                    const double steps = taps == 9 ? (double)cps * (nbs + 1) : 2.0 * nmacro;
                    double main_clk = std::max(std::max(bytes / rate, mma), steps * kStep);
                    if (co_model >= 2) {
                        const double bytes_a = (double)cps * (taps == 9 ? a_rows * 128.0 : kBlockM * mp * 128.0);
                        main_clk = std::max(main_clk, std::max(bytes_a / rate_a, (bytes - bytes_a) / rate_b));
                    }
                    const bool st = tma_store_ok && bn >= 64 && ks == 1;
                    const double epi = mp * (700.0 + ((bn + 63) / 64) * (st ? 700.0 : 2200.0));
                    const double per_sm = std::ceil((double)ctas / kSms);
                    double t = kSetup + kFirst + per_sm * main_clk + std::ceil(per_sm / occ) * epi;
                    if (pers) t = kSetup + kFirst + per_sm * std::max(main_clk, epi / nt) + epi;   // front paid once, epilogues hidden (teams alternate)
                    else if (per_sm > 1) t += (per_sm - 1) / occ * (kSetup + kFirst);          // every further wave pays the front again
                    if (ks > 1) t += 1500.0 + per_sm * (double)(ks + 1) * kBlockM * bn * 4 / 30.0;
                    t /= kClkPerUs;
                    // measured: where an N = 256 tile still leaves >= 96 CTAs it beats the model's pick -- the 128 x 256 MMA is
                    // the only shape whose operand reads fit the shared-memory bandwidth -- so it gets a bonus the clock model lacks
                    if (env_int("YDST_PREFER_BN256", 1) && bn == 256 && ctas >= 96 && ks == 1) t *= 0.5;
                    if (pers && pers_mode == 3) t *= 0.05;        // tuning hook: take the persistent loop wherever it is feasible
                    // measured: a deep 3x3 layer whose 128 x 256 tiles do not even fill one wave (19x19 512->1024: 112 tiles, 36 900 tensor
                    // clocks each) runs faster as paired 128 x 128 tiles, two per SM (34.4 -> 29.4 us)
                    if (c2 && !pers && cta2_mode == 1 && bn == 128 && taps == 9 && cin_blocks >= 8 && (long long)m_tiles128 * ((cout16 + 255) / 256) <= kSms) t *= 0.4;
                    if (t < best.model_us * 0.98) {              // near-ties go to the earlier (larger bn, fewer splits) candidate
                        best.bn = bn; best.ksplit = ks; best.cbs_per_split = cps; best.tpb = tpb; best.a_stages = a_st;
                        best.b_stages = b_stages; best.occupancy = occ; best.smem_bytes = smem; best.model_us = t; best.persistent = pers;
                        best.mpair = mp; best.b_resident = resident ? 1 : 0; best.nbuf = pers ? nbuf : 0; best.cta2 = c2; best.nteams = pers ? nt : 2;
                    }
                    }   // nt
                }
                }   // c2
            }
        }
    }
    if (getenv("YDST_DEBUG_PLAN"))
        fprintf(stderr, "  tiling m_tiles %d cout %d taps %d cin_blocks %d -> bn %d ks %d cps %d tpb %d a_st %d b_st %d smem %d occ %d pers %d mpair %d cta2 %d model %.1f us\n", m_tiles128,
                cout16, taps, cin_blocks, best.bn, best.ksplit, best.cbs_per_split, best.tpb, best.a_stages, best.b_stages, best.smem_bytes,
                best.occupancy, best.persistent, best.mpair, best.cta2, best.model_us);
    return best;
}


void conv_tc_plan(ConvTcLaunch& L, const Act& in, const Act& out, const __half* w_packed, int R, int S, int stride,
                  const float* scale, const float* bias, int act, int res_mode, const Act* res, float* out_f32, int cout_real,
                  const ConvWorkspace* ws) {
    ConvTcParams& p = L.p;
    memset(&L, 0, sizeof(L));
    L.w_ptr = w_packed;
    L.w_bytes = (unsigned)((size_t)((cout_real + 15) & ~15) * R * S * in.C * sizeof(__half));
    YDST_CHECK(in.C % 16 == 0, "conv_tc needs Cin %% 16 == 0 (got %d)", in.C);
    YDST_CHECK(in.ctot % 8 == 0 && in.coff % 8 == 0 && out.ctot % 8 == 0 && out.coff % 8 == 0, "channel strides/offsets must be multiples of 8");
    YDST_CHECK((R == 1 && S == 1) || (R == 3 && S == 3), "conv_tc supports 1x1 and 3x3 filters");
    YDST_CHECK(stride == 1 || stride == 2, "conv_tc supports stride 1 and 2");
    p.R = R; p.S = S;
    p.cin = in.C;
    p.block_k = in.C % 64 == 0 ? 64 : (in.C % 32 == 0 ? 32 : 16);
    p.cin_blocks = in.C / p.block_k;
    p.cout = (cout_real + 15) & ~15;
    YDST_CHECK(out_f32 != nullptr || out.C == cout_real, "output view has %d channels, conv produces %d", out.C, cout_real);
    p.N = out.N; p.Ho = out.H; p.Wo = out.W;
    YDST_CHECK(out.pixels() < (1LL << 31) && in.pixels() < (1LL << 31), "conv_tc: more than 2^31 padded pixels per launch");
    p.in_Wp = in.W + 2;
    p.scale = scale; p.bias = bias; p.act = act;
    p.res_mode = res_mode;
    if (res_mode) {
        YDST_CHECK(res && res->N == out.N && res->H == out.H && res->W == out.W && res->C == cout_real, "residual shape mismatch");
        p.res = res->base; p.res_ctot = res->ctot; p.res_coff = res->coff;
    }
    p.out = out.base; p.out_ctot = out.ctot; p.out_coff = out.coff; p.out_f32 = out_f32;
    const int row_bytes = p.block_k * 2;
    int m_tiles;
    if (stride == 1) {
        YDST_CHECK(in.H == out.H && in.W == out.W && in.N == out.N, "stride-1 conv must preserve the spatial size");
        p.mode = 0;
        p.P_total = out.pixels();
        m_tiles = (int)((p.P_total + kBlockM - 1) / kBlockM);
        const int halo = R == 3 ? in.W + 2 + 1 : 0;
        if (p.block_k == 64 && (kBlockM + 2 * halo) * 128 <= 96 * 1024 && env_int("YDST_CONV_V2", 1)) {
            // ---- halo kernel ----
            p.v2 = 1;
            p.halo = halo;
            ConvTiling t = conv_tc_choose_tiling(m_tiles, p.cout, R * S, p.cin_blocks, halo, ws ? ws->partial_bytes : 0, ws ? ws->n_tickets : 0, res_mode,
                                                 !out_f32 && env_int("YDST_TMA_STORE", 1));
            YDST_CHECK(t.bn >= 32, "no feasible tiling for this convolution");
            p.mpair = t.mpair;
            const int a_rows = kBlockM * t.mpair + 2 * halo;
            p.a_boxes = (a_rows + 255) / 256;
            p.a_box_rows = (a_rows + p.a_boxes - 1) / p.a_boxes;
            const int force_cps = env_int("YDST_FORCE_CPS", 0);         // test hook: force a K split of the planner's tile
            if (force_cps > 0 && ws && t.mpair == 1 && !t.persistent) {
                const int cps = std::min(force_cps, p.cin_blocks);
                const int ks = (p.cin_blocks + cps - 1) / cps;
                const long long tiles = (long long)m_tiles * ((p.cout + t.bn - 1) / t.bn);
                if ((size_t)ks * tiles * kBlockM * t.bn * 4 <= ws->partial_bytes && tiles <= ws->n_tickets) {
                    t.cbs_per_split = cps; t.ksplit = ks;
                    if (R == 1) t.tpb = std::min(t.tpb, cps);
                    const int nmacro = R == 3 ? cps : (cps + t.tpb - 1) / t.tpb;
                    t.a_stages = std::min(t.a_stages, nmacro);
                    t.b_stages = std::min(t.b_stages, nmacro * (R == 3 ? 9 / t.tpb : 1));
                }
            }
            p.block_n = t.bn; p.ksplit = t.ksplit; p.cbs_per_split = t.cbs_per_split; p.a_stages = t.a_stages; p.b_stages = t.b_stages;
            p.tpb = t.tpb;
            p.m_tiles = (m_tiles + t.mpair - 1) / t.mpair; p.n_tiles = (p.cout + t.bn - 1) / t.bn;
            p.persistent = t.persistent;
            p.b_resident = t.b_resident;
            p.ws = ws ? ws->partial : nullptr; p.tickets = ws ? ws->tickets : nullptr;
            p.bo_mode = env_int("YDST_BO_MODE", 0);
            p.store_tma = (!out_f32 && (p.cout % 64 == 0 || p.cout < 64) && t.bn >= 64 && t.ksplit == 1 && env_int("YDST_TMA_STORE", 1)) ? 1 : 0;
            YDST_CHECK(!t.persistent || p.store_tma, "persistent tiles need the TMA-store epilogue");
            p.nbuf = t.nbuf;
            p.cta2 = t.cta2;
            p.nteams = t.nteams ? t.nteams : 2;
            const int K = R * S * in.C;
            if (R == 3) {
                cuuint64_t dims[2] = {(cuuint64_t)in.C, (cuuint64_t)p.P_total};
                cuuint64_t strides[1] = {(cuuint64_t)in.ctot * 2};
                cuuint32_t box[2] = {64u, (cuuint32_t)p.a_box_rows};
                encode(&L.tmA[0], in.base + in.coff, 2, dims, strides, box, 128);
                cuuint64_t bdims[3] = {(cuuint64_t)in.C, (cuuint64_t)p.cout, 9};
                cuuint64_t bstrides[2] = {(cuuint64_t)K * 2, (cuuint64_t)in.C * 2};
                cuuint32_t bbox[3] = {64u, (cuuint32_t)(t.bn >> t.cta2), (cuuint32_t)t.tpb};
                encode(&L.tmB, w_packed, 3, bdims, bstrides, bbox, 128);
            } else {
                cuuint64_t dims[3] = {64, (cuuint64_t)p.P_total, (cuuint64_t)p.cin_blocks};
                cuuint64_t strides[2] = {(cuuint64_t)in.ctot * 2, 128};
                cuuint32_t box[3] = {64u, (cuuint32_t)(kBlockM * t.mpair), (cuuint32_t)t.tpb};
                encode(&L.tmA[0], in.base + in.coff, 3, dims, strides, box, 128);
                cuuint64_t bdims[3] = {64, (cuuint64_t)p.cout, (cuuint64_t)p.cin_blocks};
                cuuint64_t bstrides[2] = {(cuuint64_t)K * 2, 128};
                cuuint32_t bbox[3] = {64u, (cuuint32_t)(t.bn >> t.cta2), (cuuint32_t)t.tpb};
                encode(&L.tmB, w_packed, 3, bdims, bstrides, bbox, 128);
            }
            if (p.store_tma) {
                // the logical width ends with this layer's channel slice: a 64-column box never spills into a neighbouring slice
                cuuint64_t dims[2] = {(cuuint64_t)(out.coff + cout_real), (cuuint64_t)p.P_total};
                cuuint64_t strides[1] = {(cuuint64_t)out.ctot * 2};
                cuuint32_t box[2] = {64u, (cuuint32_t)kBlockM};
                encode(&L.tmA[1], out.base, 2, dims, strides, box, 128);
                if (res_mode) {                                  // the wide epilogue fetches the residual tile by TMA
                    cuuint64_t rdims[2] = {(cuuint64_t)(res->coff + cout_real), (cuuint64_t)p.P_total};
                    cuuint64_t rstrides[1] = {(cuuint64_t)res->ctot * 2};
                    encode(&L.tmA[2], res->base, 2, rdims, rstrides, box, 128);
                }
            }
            L.stages = 0;
            L.smem_bytes = t.smem_bytes + 1024;
            L.grid = dim3((unsigned)p.m_tiles, (unsigned)p.n_tiles, (unsigned)t.ksplit);
            // A persistent CTA owns its SM's shared memory for the whole launch, so a kernel of another stream (the association's
            // single-CTA LSAP runs ~70 us) that lands on an SM between two persistent launches holds back ONE CTA of the next launch --
            // and with it the launch.  YDST_PERS_SPARE_SMS SMs are therefore left out of every persistent grid.
            const int pers_sms = std::max(2, num_sms() - env_int("YDST_PERS_SPARE_SMS", 0));
            if (p.persistent) L.grid = dim3((unsigned)std::min(p.m_tiles * p.n_tiles, pers_sms), 1, 1);
            if (p.cta2) {
                YDST_CHECK(p.store_tma && t.mpair == 1 && t.ksplit == 1, "CTA pairs need the TMA-store path");
                if (p.persistent) {
                    // one cluster per SM pair strides over the pairs of M tiles; no more clusters than can be resident at once
                    const int pairs = ((p.m_tiles + 1) / 2) * p.n_tiles;
                    const int clusters = std::min(std::min(pairs, pers_sms / 2), max_active_clusters(act == ACT_MISH, L.smem_bytes));
                    L.grid = dim3(2u * (unsigned)clusters, 1, 1);
                } else {
                    L.grid.x = (L.grid.x + 1u) & ~1u;            // clusters of two along M: an odd tail tile gets an idle partner (all rows out of range)
                }
            }
            if (getenv("YDST_DEBUG_PLAN"))
                fprintf(stderr, "conv_plan2 k%d cin %d cout %d out %dx%dx%d grid %ux%ux%u bn %d cps %d tpb %d a_st %d b_st %d a_rows %d smem %d tma_st %d pers %d mpair %d res %d cta2 %d model %.1fus\n",
                        R, in.C, p.cout, out.N, out.H, out.W, L.grid.x, L.grid.y, L.grid.z, t.bn, t.cbs_per_split, t.tpb, t.a_stages, t.b_stages,
                        p.a_box_rows * p.a_boxes, L.smem_bytes, p.store_tma, p.persistent, p.mpair, p.b_resident, p.cta2, t.model_us);
            return;
        }
        cuuint64_t dims[2] = {(cuuint64_t)in.C, (cuuint64_t)p.P_total};
        cuuint64_t strides[1] = {(cuuint64_t)in.ctot * 2};
        cuuint32_t box[2] = {(cuuint32_t)p.block_k, (cuuint32_t)kBlockM};
        encode(&L.tmA[0], in.base + in.coff, 2, dims, strides, box, row_bytes);
    } else {
        YDST_CHECK(in.H % 2 == 0 && in.W % 2 == 0 && out.H == in.H / 2 && out.W == in.W / 2, "stride-2 conv needs even input dims");
        p.mode = 1;
        p.pad_shift = 1 - (R - 1) / 2;
        p.in_Hp_half = (in.H + 2) / 2;
        const int cand[5][2] = {{128, 1}, {64, 2}, {32, 4}, {16, 8}, {8, 16}};
        long long best = -1;
        for (auto& c : cand) {
            const long long t = (long long)((out.W + c[0] - 1) / c[0]) * ((out.H + c[1] - 1) / c[1]);
            if (best < 0 || t < best) { best = t; p.TW = c[0]; p.TH = c[1]; }
        }
        // small images (ReID layer4's 8 x 4 outputs): a 128-row tile over ONE image would be 3/4 padding, so it takes TN whole images
        p.TN = 1;
        if (out.H * out.W <= 64 && (out.W & (out.W - 1)) == 0 && (out.H & (out.H - 1)) == 0) {
            p.TW = out.W; p.TH = out.H; p.TN = kBlockM / (out.W * out.H);
        }
        p.tiles_x = (out.W + p.TW - 1) / p.TW;
        p.tiles_y = (out.H + p.TH - 1) / p.TH;
        m_tiles = p.tiles_x * p.tiles_y * ((out.N + p.TN - 1) / p.TN);
        for (int py = 0; py < 2; ++py)
            for (int px = 0; px < 2; ++px) {
                // parity sub-lattice (py, px) of the padded input: (channel, half-column, half-row, image)
                cuuint64_t dims[4] = {(cuuint64_t)in.C, (cuuint64_t)((in.W + 2 - px + 1) / 2), (cuuint64_t)p.in_Hp_half, (cuuint64_t)in.N};
                cuuint64_t strides[3] = {(cuuint64_t)2 * in.ctot * 2, (cuuint64_t)2 * (in.W + 2) * in.ctot * 2,
                                         (cuuint64_t)(in.H + 2) * (in.W + 2) * in.ctot * 2};
                cuuint32_t box[4] = {(cuuint32_t)p.block_k, (cuuint32_t)p.TW, (cuuint32_t)p.TH, (cuuint32_t)p.TN};
                encode(&L.tmA[py * 2 + px], in.base + ((long long)py * (in.W + 2) + px) * in.ctot + in.coff, 4, dims, strides, box, row_bytes);
            }
    }
    // N tile: the largest of {128,64,32,16} that still yields >= one CTA per SM; otherwise the smallest useful one.
    int bn = 128;
    while (bn > 16 && (bn > p.cout || (long long)m_tiles * ((p.cout + bn - 1) / bn) < num_sms())) bn >>= 1;
    if (bn < 32 && p.cout >= 32) bn = 32;
    // many short tiles (tap_pers_mode 1: at least four per SM; 2: more tiles than SMs, for the tests): the persistent form, with the
    // widest N tile (fewer re-fetches of the activation taps) -- conv_tc_pers_kernel
    const int tap_pers_mode = env_int("YDST_TAP_PERSISTENT", 1);
    bool tap_pers = false;
    if (tap_pers_mode && !out_f32) {
        int bnp = 128;
        while (bnp > 32 && bnp > p.cout) bnp >>= 1;
        const long long tiles = (long long)m_tiles * ((p.cout + bnp - 1) / bnp);
        if (p.cout % bnp == 0 && tiles >= (tap_pers_mode >= 2 ? num_sms() + 1 : 4 * num_sms())) { tap_pers = true; bn = bnp; }
    }
    p.block_n = bn;
    {
        const int K = R * S * in.C;
        cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)p.cout};
        cuuint64_t strides[1] = {(cuuint64_t)K * 2};
        cuuint32_t box[2] = {(cuuint32_t)p.block_k, (cuuint32_t)bn};
        encode(&L.tmB, w_packed, 2, dims, strides, box, row_bytes);
    }
    const int num_kb = R * S * p.cin_blocks;
    if (tap_pers) {
        const int budget = 216 * 1024, fixed = 4096 + 1024;            // barriers + two teams' scale/bias, alignment slack
        const int slab = (num_kb * bn * row_bytes + 1023) & ~1023;
        // weight-stationary when the slab leaves room for at least six activation stages
        const bool res = slab + 6 * ((kBlockM * row_bytes + 1023) & ~1023) + fixed <= budget;
        const int st_bytes = ((res ? kBlockM * row_bytes : kBlockM * row_bytes + bn * row_bytes) + 1023) & ~1023;
        const int stages_p = std::max(2, std::min(16, (budget - fixed - (res ? slab : 0)) / st_bytes));
        p.persistent = 1; p.b_resident = res ? 1 : 0;
        p.m_tiles = m_tiles; p.n_tiles = (p.cout + bn - 1) / bn;
        L.stages = stages_p;
        L.smem_bytes = (res ? slab : 0) + stages_p * st_bytes + fixed;
        L.grid = dim3((unsigned)std::min((long long)num_sms(), (long long)p.m_tiles * p.n_tiles), 1, 1);
        if (getenv("YDST_DEBUG_PLAN"))
            fprintf(stderr, "conv_plan k%d s%d cin %d cout %d out %dx%dx%d grid %ux1 bn %d bk %d stages %d smem %d tile %dx%dx%d persistent resident %d tiles %dx%d\n",
                    R, stride, in.C, p.cout, out.N, out.H, out.W, L.grid.x, bn, p.block_k, stages_p, L.smem_bytes, p.TN, p.TH, p.TW, p.b_resident,
                    p.m_tiles, p.n_tiles);
        return;
    }
    const int stage_bytes = (kBlockM * row_bytes + bn * row_bytes + 1023) & ~1023;
    int stages = std::max(2, std::min(8, (100 * 1024) / stage_bytes));
    stages = std::min(stages, std::max(num_kb, 1));
    L.stages = stages;
    L.smem_bytes = stages * stage_bytes + 16 * stages + 16 + 2048 + 1024;   // + scale/bias staging, alignment slack
    L.grid = dim3((unsigned)m_tiles, (unsigned)((p.cout + bn - 1) / bn), 1);
    if (getenv("YDST_DEBUG_PLAN"))
        fprintf(stderr, "conv_plan k%d s%d cin %d cout %d out %dx%dx%d grid %ux%u bn %d bk %d stages %d smem %d tile %dx%dx%d\n", R, stride, in.C, p.cout,
                out.N, out.H, out.W, L.grid.x, L.grid.y, bn, p.block_k, stages, L.smem_bytes, p.TN, p.TH, p.TW);
}

static void ensure_conv_tc2_attrs() {
    static bool attr2_set_dev[64] = {false};                  // the shared-memory opt-in is a per-device attribute of the kernel
    int dev = 0;
    YDST_CUDA(cudaGetDevice(&dev));
    bool& attr2_set = attr2_set_dev[dev & 63];
    if (attr2_set) return;
    const void* fns[12] = {(const void*)conv_tc2_kernel<false, false, false, true>, (const void*)conv_tc2_kernel<true, false, false, true>,
                           (const void*)conv_tc2_kernel<false, true, false, true>,  (const void*)conv_tc2_kernel<true, true, false, true>,
                           (const void*)conv_tc2_kernel<false, false, false>, (const void*)conv_tc2_kernel<false, true, false>,
                           (const void*)conv_tc2_kernel<true, false, false>,  (const void*)conv_tc2_kernel<true, true, false>,
                           (const void*)conv_tc2_kernel<false, false, true>,  (const void*)conv_tc2_kernel<false, true, true>,
                           (const void*)conv_tc2_kernel<true, false, true>,   (const void*)conv_tc2_kernel<true, true, true>};
    for (const void* fn : fns) {
        YDST_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
        // without this the driver may size the L1/shared split for ONE resident CTA; multi-wave layers want two per SM
        YDST_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }
    attr2_set = true;
}

// how many clusters of two persistent CTAs the device can hold at once (GPC boundaries may leave an SM without a partner)
static int max_active_clusters(bool mish, int smem_bytes) {
    ensure_conv_tc2_attrs();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2u * (unsigned)num_sms(), 1, 1); cfg.blockDim = dim3(kThreads2); cfg.dynamicSmemBytes = (size_t)smem_bytes;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    const void* fn = mish ? (const void*)conv_tc2_kernel<true, true, false, true> : (const void*)conv_tc2_kernel<false, true, false, true>;
    YDST_CUDA(cudaOccupancyMaxActiveClusters(&n, fn, &cfg));
    YDST_CHECK(n > 0, "no cluster of two persistent CTAs fits (smem %d)", smem_bytes);
    return std::min(n, num_sms() / 2);
}

static constexpr int kTraceSlots = 1024;
static int g_trace_on = -1, g_trace_next = 0;
static unsigned long long* g_trace_dev = nullptr;
static int g_trace_desc[kTraceSlots], g_trace_shape[kTraceSlots];

void conv_tc_plan_gemm(ConvTcLaunch& L, const __half* A, int rows, int K, const __half* Bw, int ncols, float* out_f32, const float* scale,
                       const float* bias, const ConvWorkspace* ws) {
    YDST_CHECK(rows >= 128 && rows % 128 == 0 && K % 64 == 0 && ncols % 16 == 0 && ncols > 0, "bad GEMM shape %d x %d x %d", rows, ncols, K);
    // a (rows x K) matrix is the flat-padded view of an image with H + 2 = rows and W + 2 = 1: one "pixel" per row
    Act in, out;
    in.base = const_cast<__half*>(A); in.N = 1; in.H = rows - 2; in.W = -1; in.C = K; in.ctot = K; in.coff = 0;
    out = in; out.base = nullptr; out.C = ncols; out.ctot = ncols;
    conv_tc_plan(L, in, out, Bw, 1, 1, 1, scale, bias, ACT_LINEAR, 0, nullptr, out_f32, ncols, ws);
    YDST_CHECK(L.p.v2 == 1, "the GEMM needs the halo kernel's 1x1 mode");
    L.p.gemm = 1;
}

void conv_tc_trace_dump() {
    if (g_trace_on != 2 || !g_trace_dev) return;
    cudaDeviceSynchronize();
    std::vector<unsigned long long> h((size_t)kTraceSlots * 16);
    cudaMemcpy(h.data(), g_trace_dev, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    const int n = std::min(g_trace_next, kTraceSlots);
    unsigned long long t0 = 0;
    for (int i = 0; i < n; ++i) {
        const int slot = (g_trace_next - n + i) % kTraceSlots;
        const unsigned long long* r = &h[(size_t)slot * 16];
        if (!r[8]) continue;
        if (!t0) t0 = r[8];
        fprintf(stderr, "conv_timeline %4d k%d cin %4d bn %3d W %3d cout %4d start %10.2f us dur %7.2f us  (clk: setup %llu first_data %llu mma %llu acc %llu epi %llu exit %llu)\n",
                i, g_trace_desc[slot] / 1000000, (g_trace_desc[slot] / 100) % 10000, (g_trace_desc[slot] % 100) * 32, g_trace_shape[slot] / 10000,
                g_trace_shape[slot] % 10000, (r[8] - t0) * 1e-3, (r[9] - r[8]) * 1e-3, r[1] - r[0], r[3] ? r[3] - r[0] : 0, r[4] ? r[4] - r[0] : 0,
                r[5] ? r[5] - r[0] : 0, r[6] ? r[6] - r[0] : 0, r[7] - r[0]);
        if (r[11]) fprintf(stderr, "    epilogue thread 100, group 0: tmem loaded %llu, computed+staged %llu, proxy fence %llu, barrier %llu\n", r[11] - r[0],
                           r[12] - r[0], r[13] - r[0], r[14] - r[0]);
    }
}

void conv_tc_run(const ConvTcLaunch& L, cudaStream_t stream) {
    static bool attr_set_dev[64] = {false};
    int dev_id = 0;
    YDST_CUDA(cudaGetDevice(&dev_id));
    bool& attr_set = attr_set_dev[dev_id & 63];
    if (!attr_set) {
        YDST_CUDA(cudaFuncSetAttribute(conv_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        YDST_CUDA(cudaFuncSetAttribute(conv_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        YDST_CUDA(cudaFuncSetAttribute(conv_tc_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        YDST_CUDA(cudaFuncSetAttribute(conv_tc_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        attr_set = true;
    }
    ConvTcMaps maps;
    memcpy(maps.a, L.tmA, sizeof(maps.a));
    maps.b = L.tmB;
    if (L.p.v2) {
        static int use_pdl = -1;
        ensure_conv_tc2_attrs();
        if (use_pdl < 0) use_pdl = env_int("YDST_PDL", 1);
        if (g_trace_on < 0) {
            g_trace_on = env_int("YDST_CONV_TRACE", 0);
            if (g_trace_on) {
                YDST_CUDA(cudaMalloc(&g_trace_dev, kTraceSlots * 16 * sizeof(unsigned long long)));
                YDST_CUDA(cudaMemset(g_trace_dev, 0, kTraceSlots * 16 * sizeof(unsigned long long)));
            }
        }
        const int trace_on = g_trace_on;
        unsigned long long* trace_dev = g_trace_dev;
        ConvTcParams prm = L.p;
        prm.pdl = use_pdl;
        if (trace_on == 1) { YDST_CUDA(cudaMemsetAsync(trace_dev, 0, (16 + 4 * 32) * 8, stream)); prm.trace = trace_dev; prm.trace_tiles = 1; }
        if (trace_on == 2) {                                  // timeline mode: one slot per launch, no synchronisation
            const int slot = g_trace_next++ % kTraceSlots;
            prm.trace = trace_dev + (size_t)slot * 16;          // not re-zeroed: a memset node would break the PDL chain
            g_trace_desc[slot] = L.p.R * 1000000 + L.p.cin * 100 + (L.p.block_n >> 5);   // compact tag: k, cin, bn/32
            g_trace_shape[slot] = L.p.Wo * 10000 + L.p.cout;
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = L.grid; cfg.blockDim = dim3(kThreads2); cfg.dynamicSmemBytes = (size_t)L.smem_bytes; cfg.stream = stream;
        cudaLaunchAttribute at[2];
        int nat = 0;
        if (use_pdl) {
            at[nat].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[nat].val.programmaticStreamSerializationAllowed = 1;
            ++nat;
        }
        if (L.p.cta2) {
            at[nat].id = cudaLaunchAttributeClusterDimension;
            at[nat].val.clusterDim.x = 2; at[nat].val.clusterDim.y = 1; at[nat].val.clusterDim.z = 1;
            ++nat;
        }
        cfg.attrs = at; cfg.numAttrs = nat;
        const int variant = L.p.cta2 ? (L.p.act == ACT_MISH ? 9 : 8) + (L.p.persistent ? 2 : 0) : (L.p.act == ACT_MISH ? 1 : 0) | (L.p.persistent ? 2 : 0) | (L.p.mpair == 2 ? 4 : 0);
        switch (variant) {
            case 10: YDST_CUDA(cudaLaunchKernelEx(&cfg, conv_tc2_kernel<false, true, false, true>, maps, prm)); break;
            case 11: YDST_CUDA(cudaLaunchKernelEx(&cfg, conv_tc2_kernel<true, true, false, true>, maps, prm)); break;
            case 8: YDST_CUDA(cudaLaunchKernelEx(&cfg, conv_tc2_kernel<false, false, false, true>, maps, prm)); break;
            case 9: YDST_CUDA(cudaLaunchKernelEx(&cfg, conv_tc2_kernel<true, false, false, true>, maps, prm)); break;
            case 0: YDST_CUDA(cudaLaunchKernelEx(&cfg, conv_tc2_kernel<false, false, false>, maps, prm)); break;
            case 1: YDST_CUDA(cudaLaunchKernelEx(&cfg, conv_tc2_kernel<true, false, false>, maps, prm)); break;
            case 2: YDST_CUDA(cudaLaunchKernelEx(&cfg, conv_tc2_kernel<false, true, false>, maps, prm)); break;
            case 3: YDST_CUDA(cudaLaunchKernelEx(&cfg, conv_tc2_kernel<true, true, false>, maps, prm)); break;
            case 4: YDST_CUDA(cudaLaunchKernelEx(&cfg, conv_tc2_kernel<false, false, true>, maps, prm)); break;
            case 5: YDST_CUDA(cudaLaunchKernelEx(&cfg, conv_tc2_kernel<true, false, true>, maps, prm)); break;
            case 6: YDST_CUDA(cudaLaunchKernelEx(&cfg, conv_tc2_kernel<false, true, true>, maps, prm)); break;
            default: YDST_CUDA(cudaLaunchKernelEx(&cfg, conv_tc2_kernel<true, true, true>, maps, prm)); break;
        }
        if (trace_on == 1) {
            unsigned long long h[16];
            YDST_CUDA(cudaStreamSynchronize(stream));
            YDST_CUDA(cudaMemcpy(h, trace_dev, sizeof(h), cudaMemcpyDeviceToHost));
            auto d = [&](int a, int b) { return h[b] && h[a] ? (long long)(h[b] - h[a]) : -1LL; };
            fprintf(stderr, "conv_trace k%d cin %d cout %d %dx%dx%d grid %ux%ux%u bn %d | setup %lld depwait %lld first_data %lld mma_issued %lld "
                            "acc_ready %lld epilogue_done %lld exit %lld clk | cta0 %.2f us, last CTA started +%.2f us\n",
                    L.p.R, L.p.cin, L.p.cout, L.p.N, L.p.Ho, L.p.Wo, L.grid.x, L.grid.y, L.grid.z, L.p.block_n, d(0, 1), d(0, 2), d(0, 3), d(0, 4),
                    d(0, 5), d(0, 6), d(0, 7), (h[9] - h[8]) * 1e-3, ((long long)h[10] - (long long)h[8]) * 1e-3);
            if (L.p.persistent) {
                unsigned long long tt[8 * 16];
                YDST_CUDA(cudaMemcpy(tt, trace_dev + 16, sizeof(tt), cudaMemcpyDeviceToHost));
                fprintf(stderr, "    tiles of CTA 0 (clk since start): accumulator free (MMA; tiles 8+: first 16 columns finished) | operands landed | MMAs committed | accumulator seen by the "
                                "epilogue | accumulator released | group 0 staged | barrier passed | tile stored\n");
                for (int i = 0; i < 16 && tt[8 * i]; ++i) {
                    auto d = [&](int k) { return tt[8 * i + k] ? (long long)(tt[8 * i + k] - h[0]) : -1LL; };
                    fprintf(stderr, "      [%2d] %lld | %lld | %lld | %lld | %lld | %lld | %lld | %lld\n", i, d(4), d(0), d(1), d(2), d(5), d(6), d(7), d(3));
                }
            }
            if (h[11]) fprintf(stderr, "    epilogue thread 100, group 0: acc_ready %lld | 16 cols in registers +%lld, computed +%lld, staged +%lld, next 16 cols in registers +%lld, "
                               "group staged and barrier passed +%lld (res_mode %d)\n", d(0, 5), d(5, 11), d(11, 15), d(15, 12), d(12, 13), d(13, 14), L.p.res_mode);
        }
        return;
    }
    if (L.p.persistent) {
        static bool attrp_set_dev[64] = {false};
        bool& attrp_set = attrp_set_dev[dev_id & 63];
        if (!attrp_set) {
            YDST_CUDA(cudaFuncSetAttribute(conv_tc_pers_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
            YDST_CUDA(cudaFuncSetAttribute(conv_tc_pers_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
            attrp_set = true;
        }
        if (L.p.act == ACT_MISH) conv_tc_pers_kernel<true><<<L.grid, kThreadsP, L.smem_bytes, stream>>>(maps, L.p, L.stages);
        else conv_tc_pers_kernel<false><<<L.grid, kThreadsP, L.smem_bytes, stream>>>(maps, L.p, L.stages);
        YDST_CUDA(cudaGetLastError());
        return;
    }
    if (L.p.act == ACT_MISH) conv_tc_kernel<true><<<L.grid, kThreads, L.smem_bytes, stream>>>(maps, L.p, L.stages);
    else conv_tc_kernel<false><<<L.grid, kThreads, L.smem_bytes, stream>>>(maps, L.p, L.stages);
    YDST_CUDA(cudaGetLastError());
}

double conv_tc_flops(const ConvTcLaunch& L) {
    const ConvTcParams& p = L.p;
    return 2.0 * p.N * p.Ho * p.Wo * (double)p.cout * p.R * p.S * p.cin;
}

}  // namespace ydst
