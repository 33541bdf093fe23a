// Launch-latency floor on this GPU: time per kernel of a 100-kernel dependent chain (stream vs CUDA graph), for
// an empty kernel, one with 200 KB dynamic shared memory, one that also allocates TMEM, and with/without PDL.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__global__ void k_empty(int* p) { if (p && threadIdx.x == 9999) *p = 1; }
__global__ void k_smem(int* p) { extern __shared__ char s[]; if (p && threadIdx.x == 9999) *p = s[0]; }
__global__ void k_tmem(int* p, int pdl) {
    extern __shared__ char s[];
    __shared__ uint32_t slot;
    if (pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (threadIdx.x < 32) {
        uint32_t a = (uint32_t)__cvta_generic_to_shared(&slot);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(a), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
    if (p && threadIdx.x == 9999) *p = s[0];
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(128u) : "memory");
}
template <typename F> float timeit(F f, cudaStream_t st, int reps) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaStreamSynchronize(st);
    cudaEventRecord(a, st); for (int i = 0; i < reps; ++i) f(); cudaEventRecord(b, st); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms * 1000.f / reps;
}
int main() {
    cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    const int N = 100, grid = 148, smem = 200 * 1024;
    cudaFuncSetAttribute(k_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_tmem, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    auto launch = [&](int kind, int pdl, cudaStream_t s) {
        cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(192); cfg.stream = s;
        cfg.dynamicSmemBytes = kind == 0 ? 0 : smem;
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
        int* np = nullptr;
        if (kind == 0) cudaLaunchKernelEx(&cfg, k_empty, np);
        else if (kind == 1) cudaLaunchKernelEx(&cfg, k_smem, np);
        else cudaLaunchKernelEx(&cfg, k_tmem, np, pdl);
    };
    const char* names[3] = {"empty", "200KB smem", "200KB smem + TMEM alloc"};
    for (int kind = 0; kind < 3; ++kind)
        for (int pdl = 0; pdl < 2; ++pdl) {
            float us_stream = timeit([&] { for (int i = 0; i < N; ++i) launch(kind, pdl, st); }, st, 20) / N;
            cudaGraph_t g; cudaGraphExec_t ge;
            cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
            for (int i = 0; i < N; ++i) launch(kind, pdl, st);
            cudaStreamEndCapture(st, &g); cudaGraphInstantiate(&ge, g, 0);
            float us_graph = timeit([&] { cudaGraphLaunch(ge, st); }, st, 20) / N;
            printf("%-28s pdl %d : stream %.2f us/kernel, graph %.2f us/kernel  (%s)\n", names[kind], pdl, us_stream, us_graph, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
