// zerocopy_bench.cu -- can the step kernel read its actions from, and write its small per-env outputs to,
// page-locked HOST memory directly (no cudaMemcpy before / after the kernel)?  (experiment, not product code)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/zc tools/zerocopy_bench.cu && /tmp/zc
// Model: 65,536 envs, one warp each, 4 warps per CTA, 24 warps per SM; every warp "works" for ~W microseconds
// (a dependent HBM load + a spin on clock64), reads 8 B of action and writes OUT bytes of outputs.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

// MODE 0: actions + outputs in device memory; 1: outputs AoS 32 B to host; 2: 1 + action from host;
//      3: outputs SoA (6 small stores) to host; 4: action from host only
template <int MODE>
__global__ void __launch_bounds__(128) k(const long long *act_dev, const long long *act_host, int4 *out_dev, int4 *out_host,
                                         unsigned char *soa_host, const int4 *rec, int n, int spin)
{
    extern __shared__ unsigned char smem[];
    int env = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (env >= n) return;
    if (rec == nullptr) smem[threadIdx.x] = 1;
    long long a = 0;
    if (lane == 0) a = (MODE == 2 || MODE == 4) ? act_host[env] : act_dev[env];
    int4 r = rec[(size_t)env * 64 + lane];
    a = __shfl_sync(0xffffffffu, a, 0);
    long long t0 = clock64();
    int acc = r.x ^ (int)a;
    while (clock64() - t0 < spin) acc = acc * 3 + 1;
    if (MODE == 3) {
        if (lane == 0) {
            reinterpret_cast<double *>(soa_host)[env] = (double)acc;
            reinterpret_cast<long long *>(soa_host + (size_t)n * 8)[env] = a;
            reinterpret_cast<int *>(soa_host + (size_t)n * 16)[env] = acc;
            soa_host[(size_t)n * 20 + env] = 1;
            soa_host[(size_t)n * 21 + env] = 2;
            soa_host[(size_t)n * 22 + env] = 3;
        }
    } else if (lane < 2) {
        int4 v = make_int4(acc, (int)a, lane, env);
        int4 *o = (MODE == 1 || MODE == 2) ? out_host : out_dev;
        o[(size_t)env * 2 + lane] = v;
    }
}

int main()
{
    const int n = 65536;
    long long *act_dev, *act_host; int4 *out_dev, *out_host, *rec; unsigned char *soa_host;
    CK(cudaMalloc(&act_dev, n * 8)); CK(cudaMalloc(&out_dev, n * 32)); CK(cudaMalloc(&rec, (size_t)n * 1024));
    CK(cudaMemset(act_dev, 0, n * 8)); CK(cudaMemset(rec, 1, (size_t)n * 1024));
    CK(cudaHostAlloc(&act_host, n * 8, cudaHostAllocMapped)); CK(cudaHostAlloc(&out_host, n * 32, cudaHostAllocMapped));
    CK(cudaHostAlloc(&soa_host, (size_t)n * 24, cudaHostAllocMapped));
    for (int i = 0; i < n; ++i) act_host[i] = i;
    const size_t smem_bytes = (size_t)(227 * 1024 / 6 - 1024) & ~(size_t)127;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int spin_us : {2, 5, 10}) {
        const int spin = spin_us * 1965;
        auto run = [&](const char *name, auto kern) {
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
            for (int i = 0; i < 3; ++i) kern<<<n / 4, 128, smem_bytes>>>(act_dev, act_host, out_dev, out_host, soa_host, rec, n, spin);
            CK(cudaDeviceSynchronize());
            cudaEventRecord(a);
            for (int i = 0; i < 20; ++i) kern<<<n / 4, 128, smem_bytes>>>(act_dev, act_host, out_dev, out_host, soa_host, rec, n, spin);
            cudaEventRecord(b);
            CK(cudaEventSynchronize(b));
            float ms; cudaEventElapsedTime(&ms, a, b);
            printf("spin %2d us  %-34s %.4f ms per launch\n", spin_us, name, ms / 20);
        };
        run("all in device memory", k<0>);
        run("outputs 32 B AoS -> host", k<1>);
        run("outputs AoS -> host, action <- host", k<2>);
        run("outputs SoA 6 stores -> host", k<3>);
        run("action <- host only", k<4>);
    }
    // reference: the copies as cudaMemcpyAsync
    cudaEventRecord(a);
    for (int i = 0; i < 20; ++i) { CK(cudaMemcpyAsync(act_dev, act_host, n * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpyAsync(out_host, out_dev, n * 24, cudaMemcpyDeviceToHost)); }
    cudaEventRecord(b); CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("memcpy H2D 512 KB + D2H 1.5 MB: %.4f ms per pair\n", ms / 20);
    return 0;
}
