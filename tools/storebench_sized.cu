// storebench.cu -- which store pattern reaches the HBM write peak?  (experiment, not product code)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/storebench tools/storebench.cu && /tmp/storebench
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kRegionF4 = REGION_F4;
constexpr int kEnvs = N_ENVS;

// A: warp per region, lane-interleaved float4, ascending
__global__ void warp_region(float4 *out, int n, float v)
{
    int env = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (env >= n) return;
    float4 *p = out + (size_t)env * kRegionF4 + lane;
    float4 x = make_float4(v, v, v, v);
#pragma unroll
    for (int k = 0; k < kRegionF4 / 32; ++k) p[32 * k] = x;
    if (lane < kRegionF4 % 32) p[32 * (kRegionF4 / 32)] = x;
}

// B: same, after a dependent 1 KB load per warp
__global__ void warp_region_load(float4 *out, const int4 *rec, int n)
{
    int env = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (env >= n) return;
    int4 a = rec[(size_t)env * 64 + lane], b = rec[(size_t)env * 64 + 32 + lane];
    float v = (float)((a.x ^ b.y) & 1);
    float4 *p = out + (size_t)env * kRegionF4 + lane;
    float4 x = make_float4(v, v, v, v);
#pragma unroll
    for (int k = 0; k < kRegionF4 / 32; ++k) p[32 * k] = x;
    if (lane < kRegionF4 % 32) p[32 * (kRegionF4 / 32)] = x;
}

// F: A with the residency limited through dynamic shared memory (how many open streams per SM does HBM like?)
__global__ void warp_region_smem(float4 *out, int n, float v)
{
    extern __shared__ float pad[];
    int env = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (env >= n) return;
    if (v == 12345.f) pad[threadIdx.x] = v;
    float4 *p = out + (size_t)env * kRegionF4 + lane;
    float4 x = make_float4(v, v, v, v);
#pragma unroll
    for (int k = 0; k < kRegionF4 / 32; ++k) p[32 * k] = x;
    if (lane < kRegionF4 % 32) p[32 * (kRegionF4 / 32)] = x;
}

// G: 24 warps resident per SM, but at most K of them store at the same time (per-SM token counter in global memory)
__global__ void warp_region_token(float4 *out, int n, float v, int *tokens, int K)
{
    extern __shared__ float pad[];
    int env = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (env >= n) return;
    if (v == 12345.f) pad[threadIdx.x] = v;
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    int *tok = tokens + smid * 32;
    if (lane == 0) {
        while (atomicAdd(tok, 1) >= K) { atomicSub(tok, 1); __nanosleep(500); }
    }
    __syncwarp();
    float4 *p = out + (size_t)env * kRegionF4 + lane;
    float4 x = make_float4(v, v, v, v);
#pragma unroll
    for (int k = 0; k < kRegionF4 / 32; ++k) p[32 * k] = x;
    if (lane < kRegionF4 % 32) p[32 * (kRegionF4 / 32)] = x;
    __syncwarp();
    if (lane == 0) atomicSub(tok, 1);
}

// C: CTA per group of regions, all threads sweep the group's contiguous bytes
__global__ void cta_group(float4 *out, int n, float v)
{
    int per = blockDim.x >> 5;
    size_t base = (size_t)blockIdx.x * per * kRegionF4;
    size_t cnt = (size_t)per * kRegionF4;
    float4 x = make_float4(v, v, v, v);
    for (size_t q = threadIdx.x; q < cnt; q += blockDim.x) out[base + q] = x;
}

// D: flat grid-stride fill (what a library fill does)
__global__ void flat_fill(float4 *out, size_t n4, float v)
{
    float4 x = make_float4(v, v, v, v);
    for (size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x; q < n4; q += (size_t)gridDim.x * blockDim.x) out[q] = x;
}

// E: flat, each thread 4 consecutive float4 (64 B)
__global__ void flat_fill4(float4 *out, size_t n4, float v)
{
    float4 x = make_float4(v, v, v, v);
    size_t q = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * 4;
    if (q + 3 < n4) { out[q] = x; out[q + 1] = x; out[q + 2] = x; out[q + 3] = x; }
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
    for (int wpc : {1, 2, 4, 8, 16}) {
        float t = timeit([&] { warp_region<<<(kEnvs + wpc - 1) / wpc, wpc * 32>>>(out, kEnvs, 1.f); });
        printf("A warp/region  %2d warps/CTA : %.4f ms %.0f GB/s\n", wpc, t, gb / t * 1e3);
    }
    for (int wpc : {2, 4, 8}) {
        float t = timeit([&] { warp_region_load<<<(kEnvs + wpc - 1) / wpc, wpc * 32>>>(out, rec, kEnvs); });
        printf("B +1KB load    %2d warps/CTA : %.4f ms %.0f GB/s\n", wpc, t, (gb + kEnvs * 1024 / 1e9) / t * 1e3);
    }
    for (int wpc : {4, 8, 16, 32}) {
        float t = timeit([&] { cta_group<<<kEnvs / wpc, wpc * 32>>>(out, kEnvs, 1.f); });
        printf("C cta sweep    %2d regions/CTA: %.4f ms %.0f GB/s\n", wpc, t, gb / t * 1e3);
    }
    cudaFuncSetAttribute(warp_region_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaFuncSetAttribute(warp_region_token, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    for (int ctas : {1, 2, 3, 4, 6, 8, 12, 16}) {
        size_t smem = (size_t)(220 * 1024 / ctas) & ~(size_t)1023;
        float t = timeit([&] { warp_region_smem<<<(kEnvs + 3) / 4, 128, smem>>>(out, kEnvs, 1.f); });
        printf("F warp/region  %2d warps/SM    : %.4f ms %.0f GB/s\n", ctas * 4, t, gb / t * 1e3);
    }
    int *tokens; cudaMalloc(&tokens, 256 * 32 * 4); cudaMemset(tokens, 0, 256 * 32 * 4);
    for (int K : {2, 4, 6, 8, 12, 16, 24}) {
        size_t smem = (size_t)(220 * 1024 / 6) & ~(size_t)1023;
        float t = timeit([&] { warp_region_token<<<(kEnvs + 3) / 4, 128, smem>>>(out, kEnvs, 1.f, tokens, K); });
        printf("G 24 warps/SM, %2d tokens      : %.4f ms %.0f GB/s\n", K, t, gb / t * 1e3);
    }
    for (int blocks : {148 * 4, 148 * 8, 148 * 16, 148 * 64}) {
        float t = timeit([&] { flat_fill<<<blocks, 256>>>(out, n4, 1.f); });
        printf("D flat stride  %5d blocks   : %.4f ms %.0f GB/s\n", blocks, t, gb / t * 1e3);
    }
    {
        float t = timeit([&] { flat_fill4<<<(unsigned)((n4 / 4 + 127) / 128), 128>>>(out, n4, 1.f); });
        printf("E flat 4xf4/thread           : %.4f ms %.0f GB/s\n", t, gb / t * 1e3);
    }
    return 0;
}
