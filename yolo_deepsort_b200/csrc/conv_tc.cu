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

namespace ydst {

static constexpr int kBlockM = 128;
static constexpr int kThreads = 192;   // warps 0-3: epilogue, warp 4: TMA producer, warp 5: MMA issuer + TMEM owner

struct ConvTcMaps {
    CUtensorMap a[4];
    CUtensorMap b;
};

__global__ void __launch_bounds__(kThreads) conv_tc_kernel(const __grid_constant__ ConvTcMaps maps, const ConvTcParams p, const int stages) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const uint32_t row_bytes = (uint32_t)p.block_k * 2u;
    const uint32_t a_bytes = kBlockM * row_bytes;
    const uint32_t b_bytes = (uint32_t)p.block_n * row_bytes;
    const uint32_t stage_bytes = (a_bytes + b_bytes + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + (uint32_t)stages * stage_bytes;
    // barriers: full[s] at +8s, empty[s] at +8(stages+s), tmem_full at +16*stages, tmem ptr after it
    const uint32_t bar_full = bar_base, bar_empty = bar_base + 8u * stages, bar_tmem = bar_base + 16u * stages;
    const uint32_t tmem_slot = bar_tmem + 8u;
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
        const int ty = t % p.tiles_y; img = t / p.tiles_y;
        yo0 = ty * p.TH; xo0 = tx * p.TW;
    }
    const int num_kb = p.R * p.S * p.cin_blocks;

    if (warp == 4) {
        if (lane == 0) {
            // ================= TMA producer =================
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % stages;
                const uint32_t ph = (uint32_t)(kb / stages) & 1u;
                mbar_wait(bar_empty + 8u * s, ph ^ 1u);
                const uint32_t full = bar_full + 8u * s;
                mbar_arrive_expect_tx(full, a_bytes + b_bytes);
                const int tap = kb / p.cin_blocks, cb = kb - tap * p.cin_blocks;
                const int r = tap / p.S, sx = tap - r * p.S;
                const int c0 = cb * p.block_k;
                const uint32_t a_dst = smem_base + (uint32_t)s * stage_bytes;
                if (p.mode == 0) {
                    const int row = p0 + (r - p.R / 2) * p.in_Wp + (sx - p.S / 2);
                    tma_load_2d(a_dst, &maps.a[0], full, c0, row);
                } else {
                    const int Y = r + p.pad_shift, X = sx + p.pad_shift;
                    tma_load_3d(a_dst, &maps.a[(Y & 1) * 2 + (X & 1)], full, c0, xo0 + (X >> 1),
                                img * p.in_Hp_half + yo0 + (Y >> 1));
                }
                tma_load_2d(a_dst + a_bytes, &maps.b, full, tap * p.cin + c0, n0);
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            // ================= MMA issuer =================
            const uint32_t idesc = make_idesc_f16(kBlockM, p.block_n);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % stages;
                const uint32_t ph = (uint32_t)(kb / stages) & 1u;
                mbar_wait(bar_full + 8u * s, ph);
                tcgen05_fence_after();
                const uint32_t a_addr = smem_base + (uint32_t)s * stage_bytes;
                const uint32_t b_addr = a_addr + a_bytes;
                const int ksteps = p.block_k >> 4;
                for (int k = 0; k < ksteps; ++k) {
                    const uint64_t da = make_smem_desc(a_addr + 32u * k, row_bytes);
                    const uint64_t db = make_smem_desc(b_addr + 32u * k, row_bytes);
                    umma_f16(tmem_base, da, db, idesc, (uint32_t)((kb | k) != 0));
                }
                umma_commit(bar_empty + 8u * s);     // frees the smem stage once these MMAs retire
            }
            umma_commit(bar_tmem);                   // accumulator complete
        }
    } else {
        // ================= epilogue (warps 0..3 <-> TMEM lanes 32w..32w+31) =================
        const int row = warp * 32 + lane;
        long long pix = 0;
        bool valid;
        if (p.mode == 0) {
            const long long pp = (long long)p0 + row;
            const int Wp = p.Wo + 2, HpWp = (p.Ho + 2) * Wp;
            const int rem = (int)(pp % HpWp);
            const int y = rem / Wp, x = rem - y * Wp;
            valid = pp < p.P_total && y >= 1 && y <= p.Ho && x >= 1 && x <= p.Wo;
            pix = pp;
        } else {
            const int ly = row / p.TW, lx = row - ly * p.TW;
            const int yo = yo0 + ly, xo = xo0 + lx;
            valid = yo < p.Ho && xo < p.Wo;
            pix = ((long long)img * (p.Ho + 2) + yo + 1) * (p.Wo + 2) + xo + 1;
        }
        mbar_wait(bar_tmem, 0);
        tcgen05_fence_after();
        const int nchunks = p.block_n >> 4;
        for (int ch = 0; ch < nchunks; ++ch) {
            const int c = n0 + ch * 16;
            if (c >= p.cout) break;                                   // warp-uniform
            __syncwarp();                                             // reconverge before the .sync.aligned load
            uint32_t v[16];
            tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(ch * 16), v);
            tcgen05_wait_ld();
            if (!valid) continue;
            float o[16];
            const float4* sc4 = reinterpret_cast<const float4*>(p.scale + c);
            const float4* bi4 = reinterpret_cast<const float4*>(p.bias + c);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 sc = __ldg(sc4 + q), bi = __ldg(bi4 + q);
                o[4 * q + 0] = fmaf(__uint_as_float(v[4 * q + 0]), sc.x, bi.x);
                o[4 * q + 1] = fmaf(__uint_as_float(v[4 * q + 1]), sc.y, bi.y);
                o[4 * q + 2] = fmaf(__uint_as_float(v[4 * q + 2]), sc.z, bi.z);
                o[4 * q + 3] = fmaf(__uint_as_float(v[4 * q + 3]), sc.w, bi.w);
            }
            float rs[16];
            if (p.res_mode) {
                const uint4* rp = reinterpret_cast<const uint4*>(p.res + pix * p.res_ctot + p.res_coff + c);
                const uint4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
                const __half2* h0 = reinterpret_cast<const __half2*>(&r0);
                const __half2* h1 = reinterpret_cast<const __half2*>(&r1);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float2 f0 = __half22float2(h0[q]), f1 = __half22float2(h1[q]);
                    rs[2 * q] = f0.x; rs[2 * q + 1] = f0.y;
                    rs[8 + 2 * q] = f1.x; rs[8 + 2 * q + 1] = f1.y;
                }
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                float t = o[j];
                if (p.res_mode == 2) t += rs[j];
                t = apply_act(t, p.act);
                if (p.res_mode == 1) t += rs[j];
                o[j] = t;
            }
            if (p.out_f32) {
                float4* op = reinterpret_cast<float4*>(p.out_f32 + pix * p.cout + c);
#pragma unroll
                for (int q = 0; q < 4; ++q) op[q] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
            } else {
                uint4 w0, w1;
                __half2* g0 = reinterpret_cast<__half2*>(&w0);
                __half2* g1 = reinterpret_cast<__half2*>(&w1);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    g0[q] = __floats2half2_rn(o[2 * q], o[2 * q + 1]);
                    g1[q] = __floats2half2_rn(o[8 + 2 * q], o[8 + 2 * q + 1]);
                }
                uint4* op = reinterpret_cast<uint4*>(p.out + pix * p.out_ctot + p.out_coff + c);
                op[0] = w0;
                op[1] = w1;
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 5) {
        __syncwarp();
        tmem_dealloc(tmem_base, tmem_cols);
    }
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

void conv_tc_plan(ConvTcLaunch& L, const Act& in, const Act& out, const __half* w_packed, int R, int S, int stride,
                  const float* scale, const float* bias, int act, int res_mode, const Act* res, float* out_f32, int cout_real) {
    ConvTcParams& p = L.p;
    memset(&L, 0, sizeof(L));
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
        p.tiles_x = (out.W + p.TW - 1) / p.TW;
        p.tiles_y = (out.H + p.TH - 1) / p.TH;
        m_tiles = p.tiles_x * p.tiles_y * out.N;
        for (int py = 0; py < 2; ++py)
            for (int px = 0; px < 2; ++px) {
                cuuint64_t dims[3] = {(cuuint64_t)in.C, (cuuint64_t)((in.W + 2 - px + 1) / 2), (cuuint64_t)in.N * p.in_Hp_half};
                cuuint64_t strides[2] = {(cuuint64_t)2 * in.ctot * 2, (cuuint64_t)2 * (in.W + 2) * in.ctot * 2};
                cuuint32_t box[3] = {(cuuint32_t)p.block_k, (cuuint32_t)p.TW, (cuuint32_t)p.TH};
                encode(&L.tmA[py * 2 + px], in.base + ((long long)py * (in.W + 2) + px) * in.ctot + in.coff, 3, dims, strides, box, row_bytes);
            }
    }
    // N tile: the largest of {128,64,32,16} that still yields >= one CTA per SM; otherwise the smallest useful one.
    int bn = 128;
    while (bn > 16 && (bn > p.cout || (long long)m_tiles * ((p.cout + bn - 1) / bn) < num_sms())) bn >>= 1;
    if (bn < 32 && p.cout >= 32) bn = 32;
    p.block_n = bn;
    {
        const int K = R * S * in.C;
        cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)p.cout};
        cuuint64_t strides[1] = {(cuuint64_t)K * 2};
        cuuint32_t box[2] = {(cuuint32_t)p.block_k, (cuuint32_t)bn};
        encode(&L.tmB, w_packed, 2, dims, strides, box, row_bytes);
    }
    const int stage_bytes = (kBlockM * row_bytes + bn * row_bytes + 1023) & ~1023;
    const int num_kb = R * S * p.cin_blocks;
    int stages = std::max(2, std::min(8, (100 * 1024) / stage_bytes));
    stages = std::min(stages, std::max(num_kb, 1));
    L.stages = stages;
    L.smem_bytes = stages * stage_bytes + 16 * stages + 16 + 1024;
    L.grid = dim3((unsigned)m_tiles, (unsigned)((p.cout + bn - 1) / bn), 1);
}

void conv_tc_run(const ConvTcLaunch& L, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        YDST_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    ConvTcMaps maps;
    memcpy(maps.a, L.tmA, sizeof(maps.a));
    maps.b = L.tmB;
    conv_tc_kernel<<<L.grid, kThreads, L.smem_bytes, stream>>>(maps, L.p, L.stages);
    YDST_CUDA(cudaGetLastError());
}

double conv_tc_flops(const ConvTcLaunch& L) {
    const ConvTcParams& p = L.p;
    return 2.0 * p.N * p.Ho * p.Wo * (double)p.cout * p.R * p.S * p.cin;
}

}  // namespace ydst
