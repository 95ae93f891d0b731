// storebench_read.cu -- what does the dependent record read cost the observation store stream?
// (experiment, not product code; results in DESIGN.md section 7)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/sbr tools/storebench_read.cu && /tmp/sbr
// One warp per env writes 18,000 B with STG.128 after reading RB bytes of "record":
//   dram : every env reads its own RB bytes (array larger than L2 together with the output stream)
//   l2   : every env reads the same RB bytes (always an L2 hit): isolates the dependency from the DRAM traffic
//   wb   : dram + the env writes WB bytes of record back after the stores (the real kernel's write-back)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)
constexpr int kN4 = 1125;
constexpr int kStride = 2560;     // bytes between records (the real record is 2528 B)

template <int RB, bool SAME, int WB>
__global__ void __launch_bounds__(128) k(float4 *out, int4 *rec, int n)
{
    extern __shared__ unsigned char smem[];
    int env = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (env >= n) return;
    if (rec == nullptr) smem[threadIdx.x] = 1;
    int4 *r = rec + (SAME ? 0 : (size_t)env * (kStride / 16));
    int acc = 0;
#pragma unroll
    for (int k = 0; k < (RB + 511) / 512; ++k)
        if (lane + 32 * k < RB / 16) { int4 a = r[lane + 32 * k]; acc ^= a.x ^ a.y; }
    acc = __reduce_xor_sync(0xffffffffu, acc);
    float v = (float)(acc & 1);
    float4 *p = out + (size_t)env * kN4 + lane;
    float4 x = make_float4(v, v, v, v);
#pragma unroll 8
    for (int k = 0; k < kN4 / 32; ++k) p[32 * k] = x;
    if (lane < kN4 % 32) p[32 * (kN4 / 32)] = x;
    if (WB > 0) {
        int4 *w = rec + (size_t)env * (kStride / 16);
        if (lane < WB / 16) w[lane] = make_int4(acc, acc, 1, 1);
    }
}

template <typename F> float timeit(F f, int iters = 20)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) f();
    CK(cudaGetLastError());
    cudaEventRecord(a);
    for (int i = 0; i < iters; ++i) f();
    cudaEventRecord(b);
    CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / iters;
}

int main()
{
    const int n = 65536;
    float4 *out; int4 *rec;
    CK(cudaMalloc(&out, (size_t)n * kN4 * 16)); CK(cudaMalloc(&rec, (size_t)n * kStride)); CK(cudaMemset(rec, 1, (size_t)n * kStride));
    const size_t smem_bytes = (size_t)(227 * 1024 / 6 - 1024) & ~(size_t)127;      // 6 CTAs = 24 warps per SM
    const int grid = n / 4;
    auto run = [&](const char *name, auto kern) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
        float t = timeit([&] { kern<<<grid, 128, smem_bytes>>>(out, rec, n); });
        printf("%-28s %.4f ms\n", name, t);
    };
    run("read    0 B", k<0, false, 0>);
    run("read  256 B dram", k<256, false, 0>);
    run("read  512 B dram", k<512, false, 0>);
    run("read 1024 B dram", k<1024, false, 0>);
    run("read 1536 B dram", k<1536, false, 0>);
    run("read 2048 B dram", k<2048, false, 0>);
    run("read  512 B l2", k<512, true, 0>);
    run("read 1024 B l2", k<1024, true, 0>);
    run("read 2048 B l2", k<2048, true, 0>);
    run("read 1024 B dram + wb 256", k<1024, false, 256>);
    run("read 1024 B dram + wb 512", k<1024, false, 512>);
    run("read  512 B dram + wb 256", k<512, false, 256>);
    run("read  512 B dram + wb 512", k<512, false, 512>);
    return 0;
}
