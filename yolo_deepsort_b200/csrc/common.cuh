// Shared helpers for libydst (sm_100a only): error plumbing, PTX wrappers for mbarrier / TMA /
// tcgen05 / TMEM, the padded-NHWC activation view, and small device utilities.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

namespace ydst {

// ---------------------------------------------------------------------------------------------
// errors: no C++ exception crosses the C ABI; every entry point returns an int status and leaves a
// message retrievable through ydst_last_error().
// ---------------------------------------------------------------------------------------------
void set_error(const std::string& msg);
// every kernel launch of this library is counted here (bench.py reports it as gpu_launches)
void count_launch(int n = 1);
struct Error {
    std::string msg;
};
#define YDST_CHECK(cond, ...)                                             \
    do {                                                                  \
        if (!(cond)) {                                                    \
            char _b[512];                                                 \
            snprintf(_b, sizeof(_b), __VA_ARGS__);                        \
            throw ::ydst::Error{std::string(_b) + " [" #cond "] at " __FILE__ ":" + std::to_string(__LINE__)}; \
        }                                                                 \
    } while (0)
#define YDST_CUDA(expr)                                                   \
    do {                                                                  \
        cudaError_t _e = (expr);                                          \
        if (_e != cudaSuccess)                                            \
            throw ::ydst::Error{std::string("CUDA error: ") + cudaGetErrorString(_e) + " in " #expr " at " __FILE__ ":" + std::to_string(__LINE__)}; \
    } while (0)
#define YDST_API_BEGIN try {
#define YDST_API_END                                   \
    }                                                  \
    catch (const ::ydst::Error& e) {                   \
        ::ydst::set_error(e.msg);                      \
        return 1;                                      \
    }                                                  \
    catch (const std::exception& e) {                  \
        ::ydst::set_error(std::string("exception: ") + e.what()); \
        return 2;                                      \
    }                                                  \
    return 0;

// ---------------------------------------------------------------------------------------------
// Activation view: NHWC fp16 with a physical one-pixel zero border ("flat-padded" layout).
//   pixel (n, y, x), y in [-1, H], x in [-1, W]  ->  flat index p = (n*(H+2) + y+1)*(W+2) + x+1
//   element address = base + p*ctot + coff + c
// Kernels only ever write interior pixels, so the border stays zero from the allocation-time
// memset and a 3x3/pad-1 conv becomes nine row-shifted reads of the same 2-D [pixels, channels]
// matrix.
// ---------------------------------------------------------------------------------------------
struct Act {
    __half* base = nullptr;   // buffer start (channel 0 of pixel p=0)
    int N = 0, H = 0, W = 0;  // logical dims
    int C = 0;                // channels of this view
    int ctot = 0;             // channel stride of the underlying buffer
    int coff = 0;             // first channel of this view inside the buffer
    __host__ __device__ int Hp() const { return H + 2; }
    __host__ __device__ int Wp() const { return W + 2; }
    __host__ __device__ long long pixels() const { return (long long)N * Hp() * Wp(); }
};

enum ActKind { ACT_LINEAR = 0, ACT_LEAKY = 1, ACT_MISH = 2, ACT_RELU = 3 };

__device__ __forceinline__ float apply_act(float x, int act) {
    if (act == ACT_LEAKY) return x > 0.f ? x : 0.1f * x;
    if (act == ACT_RELU) return fmaxf(x, 0.f);
    if (act == ACT_MISH) {
        // x * tanh(softplus(x)), softplus with torch's threshold 20 (yolo3/models/models.py:21).  With n = e^x:
        //   tanh(ln(1 + n)) = ((1 + n)^2 - 1) / ((1 + n)^2 + 1) = w / (w + 2),  w = n (n + 2)
        // -- one MUFU.EX2 and one MUFU.RCP instead of expf + log1pf + tanhf (about 60 instructions); no cancellation on either
        // side (w -> n for x << 0, w / (w + 2) -> 1 for x >> 0), relative error ~1e-6, far below the fp16 rounding of the result.
        const float n = __expf(x);
        const float w = n * (n + 2.f);
        return x > 20.f ? x : x * __fdividef(w, w + 2.f);
    }
    return x;
}

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .b32 %%rx;\n\t.reg .pred %%px;\n\t"
        "elect.sync %%rx|%%px, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, %%px;\n\t}\n"
        : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must abort the kernel (trap -> launch error), never hang the GPU.
// Used by the epilogue warps, which wait for the whole main loop: they back off with nanosleep so that their polling does not
// take issue slots from the single MMA-issuing / TMA-issuing threads that share their scheduler.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(spins < 64 ? 100 : 400);
        if (++spins > (1u << 24)) {
            printf("ydst: mbarrier wait timed out (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
            __trap();
        }
    }
}

// TMA tiled loads, completion on an mbarrier (complete_tx::bytes)
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
// TMA tiled store smem -> global (bulk async group)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], fp16 inputs, fp32 accumulate, issued by ONE thread.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive on an mbarrier once all tcgen05 ops previously issued by this thread have completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 16 consecutive fp32 columns: thread i of the warp receives lane (base_lane + i).
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}

// UMMA shared-memory descriptor for a K-major tile whose rows are `row_bytes` (32/64/128) long and
// swizzled with the matching TMA swizzle mode; 8-row groups are `8*row_bytes` apart (SBO).
// Bit layout: cute/arch/mma_sm100_desc.hpp (SmemDescriptor): addr>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout_type [61,64) with 128B=2, 64B=4, 32B=6.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t row_bytes) {
    uint64_t layout = row_bytes == 128 ? 2ull : (row_bytes == 64 ? 4ull : 6ull);
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;                               // LBO (unused for swizzled K-major) = 1
    d |= (uint64_t)((8u * row_bytes) >> 4) << 32;         // SBO
    d |= (uint64_t)1 << 46;                               // descriptor version (Blackwell)
    d |= layout << 61;
    return d;
}
// Instruction descriptor, kind::f16: D=f32, A=B=f16, both K-major, M=128 (InstrDescriptor bit layout).
__host__ __device__ inline uint32_t make_idesc_f16(int M, int N) {
    return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct __align__(16) Half8 {
    __half2 a, b, c, d;
};

}  // namespace ydst
