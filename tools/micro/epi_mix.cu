// Microbenchmark: which part of the convolution epilogue's instruction mix is slow on sm_100a?
// Each variant runs the per-16-column body (LDS scale/bias, FFMA, leaky = FMUL+FMNMX, F2FP pack, STS.128) with parts removed.
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
template <int kMask>   // bit0 LDS, bit1 leaky, bit2 pack, bit3 STS, bit4 fmnmx->fmul replacement
__global__ void k(float* out, long long* clk, int iters) {
    __shared__ __align__(16) float sb[512];
    __shared__ __align__(16) uint4 stage[256 * 8];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) sb[i] = 1.0f + i * 1e-3f;
    __syncthreads();
    float acc[16];
    for (int j = 0; j < 16; ++j) acc[j] = threadIdx.x * 0.01f + j;
    float4 sc[4], bi[4];
    for (int q = 0; q < 4; ++q) { sc[q] = make_float4(1.01f, 0.99f, 1.02f, 0.98f); bi[q] = make_float4(0.1f, -0.1f, 0.2f, -0.2f); }
    const long long t0 = clock64();
    uint4 w0 = make_uint4(0, 0, 0, 0), w1 = w0;
    for (int it = 0; it < iters; ++it) {
        float o[16];
        const int cl = (it & 15) * 16;
        if (kMask & 1) {
            const float4* s4 = reinterpret_cast<const float4*>(sb + cl);
            const float4* b4 = reinterpret_cast<const float4*>(sb + 256 + cl);
            for (int q = 0; q < 4; ++q) { sc[q] = s4[q]; bi[q] = b4[q]; }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            o[4 * q + 0] = fmaf(acc[4 * q + 0], sc[q].x, bi[q].x);
            o[4 * q + 1] = fmaf(acc[4 * q + 1], sc[q].y, bi[q].y);
            o[4 * q + 2] = fmaf(acc[4 * q + 2], sc[q].z, bi[q].z);
            o[4 * q + 3] = fmaf(acc[4 * q + 3], sc[q].w, bi[q].w);
        }
        if (kMask & 2) {
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = (kMask & 16) ? (o[j] * 0.1f) * 1.5f : fmaxf(o[j], 0.1f * o[j]);
        }
        if (kMask & 4) {
            __half2* g0 = reinterpret_cast<__half2*>(&w0);
            __half2* g1 = reinterpret_cast<__half2*>(&w1);
#pragma unroll
            for (int q = 0; q < 4; ++q) { g0[q] = __floats2half2_rn(o[2 * q], o[2 * q + 1]); g1[q] = __floats2half2_rn(o[8 + 2 * q], o[8 + 2 * q + 1]); }
        } else {
            w0 = make_uint4(__float_as_uint(o[0] + o[1]), __float_as_uint(o[2] + o[3]), __float_as_uint(o[4] + o[5]), __float_as_uint(o[6] + o[7]));
            w1 = make_uint4(__float_as_uint(o[8] + o[9]), __float_as_uint(o[10] + o[11]), __float_as_uint(o[12] + o[13]), __float_as_uint(o[14] + o[15]));
        }
        if (kMask & 8) {
            const int x = threadIdx.x & 7, ch = (it & 3) * 2;
            stage[threadIdx.x * 8 + (ch ^ x)] = w0;
            stage[threadIdx.x * 8 + ((ch + 1) ^ x)] = w1;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = o[j] * 0.5f + (float)(w0.x & 1);     // loop-carried dependence keeps everything live
    }
    const long long t1 = clock64();
    float s = 0;
    for (int j = 0; j < 16; ++j) s += acc[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)w1.y + (float)stage[threadIdx.x].x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
template <int kMask>
void run(const char* name, int threads) {
    float* out; long long* clk; long long h;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 8);
    const int iters = 2000;
    k<kMask><<<148, threads>>>(out, clk, iters);
    k<kMask><<<148, threads>>>(out, clk, iters);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    printf("%-28s threads %3d : %6.1f clk per 16-column body (%s)\n", name, threads, (double)h / iters, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(clk);
}
int main() {
    for (int threads : {128, 256}) {
        run<15>("all", threads);
        run<14>("no LDS", threads);
        run<13>("no leaky", threads);
        run<31>("leaky as 2 FMUL", threads);
        run<11>("no F2FP pack", threads);
        run<7>("no STS", threads);
        run<0>("FFMA + carry only", threads);
    }
    return 0;
}
