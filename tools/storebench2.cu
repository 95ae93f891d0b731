// storebench2.cu -- does a latency-bound "logic" phase before the stores change the achievable write rate,
// and does a CTA-cooperative store phase fix it?  (experiment, not product code)
#include <cstdio>
#include <cuda_runtime.h>
constexpr int kRegionF4 = 1125;
constexpr int kEnvs = 65536;

__device__ __forceinline__ float spin(float x, int iters)
{
    for (int i = 0; i < iters; ++i) x = __fmaf_rn(x, 1.0000001f, 0.5f);   // dependent chain, ~4 cycles each
    return x;
}

// F: warp per region: 1 KB load, spin, then 36 stores per lane
__global__ void warp_logic_store(float4 *out, const int4 *rec, int n, int iters)
{
    int env = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (env >= n) return;
    int4 a = rec[(size_t)env * 64 + lane], b = rec[(size_t)env * 64 + 32 + lane];
    float v = spin((float)((a.x ^ b.y) & 1), iters);
    float4 *p = out + (size_t)env * kRegionF4 + lane;
    float4 x = make_float4(v, v, v, v);
#pragma unroll
    for (int k = 0; k < kRegionF4 / 32; ++k) p[32 * k] = x;
    if (lane < kRegionF4 % 32) p[32 * (kRegionF4 / 32)] = x;
}

// G: same work, but after the logic the CTA's warps sweep the CTA's regions together, one region at a time
__global__ void cta_logic_store(float4 *out, const int4 *rec, int n, int iters)
{
    __shared__ float vals[32];
    const int per = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int env = blockIdx.x * per + warp;
    float v = 0.f;
    if (env < n) {
        int4 a = rec[(size_t)env * 64 + lane], b = rec[(size_t)env * 64 + 32 + lane];
        v = spin((float)((a.x ^ b.y) & 1), iters);
    }
    if (lane == 0) vals[warp] = v;
    __syncthreads();
    for (int e = 0; e < per; ++e) {
        int g = blockIdx.x * per + e;
        if (g >= n) break;
        float x1 = vals[e];
        float4 x = make_float4(x1, x1, x1, x1);
        float4 *p = out + (size_t)g * kRegionF4;
        for (int q = threadIdx.x; q < kRegionF4; q += blockDim.x) p[q] = x;
    }
}

template <typename F> float timeit(F f, int iters = 20)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) f();
    cudaEventRecord(a);
    for (int i = 0; i < iters; ++i) f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / iters;
}

int main()
{
    size_t n4 = (size_t)kEnvs * kRegionF4;
    float4 *out; int4 *rec;
    cudaMalloc(&out, n4 * 16); cudaMalloc(&rec, (size_t)kEnvs * 1024); cudaMemset(rec, 1, (size_t)kEnvs * 1024);
    double gb = n4 * 16 / 1e9;
    for (int iters : {0, 500, 1000, 2000}) {
        for (int smem_kb : {0, 9, 13, 19, 28}) {          // occupancy throttle for 4-warp CTAs: 16 / ~6 / ~4 / ~3 / 2 CTAs... via dynamic smem
            float t = timeit([&] {
                cudaFuncSetAttribute(warp_logic_store, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
                warp_logic_store<<<kEnvs / 4, 128, smem_kb * 1024 * 4>>>(out, rec, kEnvs, iters); });
            printf("F warp  spin %4d  smem/CTA %3d KB : %.4f ms %.0f GB/s\n", iters, smem_kb * 4, t, gb / t * 1e3);
        }
        for (int wpc : {4, 8, 16}) {
            float t = timeit([&] { cta_logic_store<<<kEnvs / wpc, wpc * 32>>>(out, rec, kEnvs, iters); });
            printf("G cta%2d spin %4d                  : %.4f ms %.0f GB/s\n", wpc, iters, t, gb / t * 1e3);
        }
    }
    return 0;
}
