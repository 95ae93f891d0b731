// compressbench.cu -- does B200's compute data compression (compressible allocations, cuMemCreate +
// CU_MEM_ALLOCATION_COMP_GENERIC) cut the DRAM cost of the observation stream?  60 % of every observation is zeros and
// another 27 % is 12 planes that broadcast one scalar each: highly compressible, and compression happens in L2 on the
// way to DRAM, transparent to readers.  (experiment, not product code; results in DESIGN.md section 7)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/compressbench tools/compressbench.cu -lcuda
// model: one warp per env, RB bytes of record read from ordinary memory, the 45-plane observation of a 10x10 board
// written with the run structure of the real one (27 zero planes, 12 broadcast planes, 6 map planes of 0/1 patterns,
// a few sparse 4-byte fix-ups), 384 B of record written back.  Output buffer: cudaMalloc vs compressible.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)
#define CU(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char *s_; cuGetErrorString(r_, &s_); printf("%s: %s\n", #x, s_); exit(1); } } while (0)
constexpr int kN4 = 1125;      // float4 per env (45 planes x 25)

// MODE 0: all zeros; 1: realistic mix; 2: realistic mix + sparse 4-byte fix-ups on top of the dense stores;
// 3: the dense pass leaves out the float4 that hold a sparse value, those get a 16-byte zero store and then the 4-byte value;
// 4: the dense pass leaves them out and each gets ONE 16-byte store with the value in place (every byte written once)
// 5: as 2, but every fix-up follows the dense store of its own plane at once (the line cannot have left L2)
__device__ __forceinline__ bool flagged(int plane, int l)
{
    if (plane >= 15 && plane <= 18) return l < 8;                                   // cells = lane, lanes with (lane & 3) == plane - 15
    if (plane >= 25 && plane <= 32) { const int r = plane - 25; return l < 16 && ((2 * l) & 7) == (r & 6) ? true : false; }
    return false;
}
template <int RB, int MODE>
__global__ void __launch_bounds__(128) k(float4 *out, int4 *rec, int n)
{
    extern __shared__ unsigned char smem[];
    int env = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (rec == nullptr) smem[threadIdx.x] = 1;
    if (env >= n) return;
    int4 *r = reinterpret_cast<int4 *>(reinterpret_cast<char *>(rec) + (size_t)env * 2560);
    int acc = 0;
#pragma unroll
    for (int q = 0; q < (RB + 511) / 512; ++q)
        if (lane + 32 * q < RB / 16) { int4 a = r[lane + 32 * q]; acc ^= a.x ^ a.y; }
    acc = __reduce_xor_sync(0xffffffffu, acc);
    const float v = (float)(acc & 1);
    float4 *p = out + (size_t)env * kN4;
    // planes: 0-9 map-like (0/1 patterns, plane 9 a ramp), 10 zero, 11-14 broadcast, 15-20 zero, 21-24 broadcast, 25-40 zero, 41-44 broadcast
    for (int plane = 0; plane < 45; ++plane) {
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (MODE != 0) {
            const bool bc = (plane >= 11 && plane <= 14) || (plane >= 21 && plane <= 24) || plane >= 41 || plane == 5;
            if (bc) { const float s = 0.125f * (float)((env + plane) & 7) + v; x = make_float4(s, s, s, s); }
            else if (plane < 10 && plane != 5) {
                const unsigned hsh = (unsigned)(env * 2654435761u) >> (plane + 3);
                const float a = (float)((hsh >> (lane & 7)) & 1u), b = (float)((hsh >> ((lane + 3) & 7)) & 1u);
                x = plane == 9 ? make_float4(a * 0.03125f * lane, 0.f, b * 0.0625f, 0.f) : make_float4(a, 0.f, 0.f, b);
            }
        }
        if (lane < 25 && !((MODE == 3 || MODE == 4) && flagged(plane, lane))) p[plane * 25 + lane] = x;
        if (MODE == 5) {
            float *f = reinterpret_cast<float *>(p);
            if (plane >= 15 && plane <= 18) { __syncwarp(); if ((lane & 3) == plane - 15) f[plane * 100 + lane] = 1.f; }
            if (plane >= 25 && plane <= 32) { __syncwarp(); if ((lane & 7) == plane - 25) f[plane * 100 + lane * 2] = 0.5f + v; }
        }
    }
    if (MODE == 2) {
        __syncwarp();
        float *f = reinterpret_cast<float *>(p);
        f[(15 + (lane & 3)) * 100 + lane] = 1.f;
        f[(25 + (lane & 7)) * 100 + lane * 2] = 0.5f + v;
    }
    if (MODE == 3) {
        float *f = reinterpret_cast<float *>(p);
        const int i1 = (15 + (lane & 3)) * 100 + lane, i2 = (25 + (lane & 7)) * 100 + lane * 2;
        p[i1 >> 2] = make_float4(0.f, 0.f, 0.f, 0.f);
        p[i2 >> 2] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncwarp();
        f[i1] = 1.f;
        f[i2] = 0.5f + v;
    }
    if (MODE == 4) {
        const int i1 = (15 + (lane & 3)) * 100 + lane, i2 = (25 + (lane & 7)) * 100 + lane * 2;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        (&a.x)[i1 & 3] = 1.f;
        (&b.x)[i2 & 3] = 0.5f + v;
        p[i1 >> 2] = a;
        p[i2 >> 2] = b;
    }
    if (RB > 0 && lane < 24) r[lane] = make_int4(acc + 1, acc, 1, 1);
}

__global__ void reader(const float4 *in, size_t n4, float *sink)
{
    float s = 0.f;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 a = in[i];
        s += a.x + a.y + a.z + a.w;
    }
    if (s == 123.456f) *sink = s;
}

template <typename F> float timeit(F f, int iters = 30)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 5; ++i) f();
    CK(cudaGetLastError());
    cudaEventRecord(a);
    for (int i = 0; i < iters; ++i) f();
    cudaEventRecord(b);
    CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / iters;
}

static void *alloc_compressible(size_t bytes, int dev, bool *compressed)
{
    CUmemAllocationProp prop = {};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = dev;
    prop.allocFlags.compressionType = CU_MEM_ALLOCATION_COMP_GENERIC;
    size_t gran = 0;
    CU(cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
    size_t size = (bytes + gran - 1) / gran * gran;
    CUmemGenericAllocationHandle h;
    CU(cuMemCreate(&h, size, &prop, 0));
    CUmemAllocationProp got = {};
    CU(cuMemGetAllocationPropertiesFromHandle(&got, h));
    *compressed = got.allocFlags.compressionType == CU_MEM_ALLOCATION_COMP_GENERIC;
    CUdeviceptr p;
    CU(cuMemAddressReserve(&p, size, gran, 0, 0));
    CU(cuMemMap(p, size, 0, h, 0));
    CUmemAccessDesc acc = {};
    acc.location = prop.location;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    CU(cuMemSetAccess(p, size, &acc, 1));
    printf("compressible allocation: %zu bytes, granularity %zu, compression %s\n", size, gran, *compressed ? "GENERIC" : "dropped by the driver");
    return reinterpret_cast<void *>(p);
}

int main()
{
    const int n = 65536;
    CK(cudaSetDevice(0));
    CK(cudaFree(0));
    int sup = 0;
    CU(cuDeviceGetAttribute(&sup, CU_DEVICE_ATTRIBUTE_GENERIC_COMPRESSION_SUPPORTED, 0));
    printf("CU_DEVICE_ATTRIBUTE_GENERIC_COMPRESSION_SUPPORTED = %d\n", sup);
    if (!sup) return 0;
    const size_t bytes = (size_t)n * kN4 * 16;
    float4 *plain; int4 *rec; float *sink;
    CK(cudaMalloc(&plain, bytes)); CK(cudaMalloc(&rec, (size_t)n * 2560)); CK(cudaMemset(rec, 1, (size_t)n * 2560)); CK(cudaMalloc(&sink, 4));
    bool compressed = false;
    float4 *comp = static_cast<float4 *>(alloc_compressible(bytes, 0, &compressed));
    const size_t smem_bytes = (size_t)(227 * 1024 / 7 - 1024) & ~(size_t)127;      // 7 CTAs = 28 warps per SM
    auto run = [&](const char *name, auto kern, int4 *r) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
        float a = timeit([&] { kern<<<n / 4, 128, smem_bytes>>>(plain, r, n); });
        float ra = timeit([&] { reader<<<148 * 8, 512>>>(plain, bytes / 16, sink); }, 10);
        float b = timeit([&] { kern<<<n / 4, 128, smem_bytes>>>(comp, r, n); });
        float rb = timeit([&] { reader<<<148 * 8, 512>>>(comp, bytes / 16, sink); }, 10);
        printf("%-44s cudaMalloc %.4f ms (read back %.4f) | compressible %.4f ms (read back %.4f)\n", name, a, ra, b, rb);
    };
    run("zeros, no record", k<0, 0>, rec);
    run("realistic planes, no record", k<0, 1>, rec);
    run("realistic planes + fix-ups, no record", k<0, 2>, rec);
    run("zeros, 1 KB record", k<1024, 0>, rec);
    run("realistic planes, 1 KB record", k<1024, 1>, rec);
    run("realistic planes + fix-ups, 1 KB record", k<1024, 2>, rec);
    run("dense leaves holes, 16 B zero + 4 B value, 1 KB", k<1024, 3>, rec);
    run("dense leaves holes, one 16 B store each, 1 KB", k<1024, 4>, rec);
    run("dense leaves holes, one 16 B store, no record", k<0, 4>, rec);
    run("fix-ups right behind their plane, 1 KB record", k<1024, 5>, rec);
    // fill through the runtime
    float a = timeit([&] { cudaMemsetAsync(plain, 0, bytes); }), b = timeit([&] { cudaMemsetAsync(comp, 0, bytes); });
    printf("cudaMemset of the buffer: cudaMalloc %.4f ms | compressible %.4f ms\n", a, b);
    return 0;
}
