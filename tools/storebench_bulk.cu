// storebench_bulk.cu -- does a bulk-async (TMA, cp.async.bulk.global.shared::cta) store path beat the STG.128
// pattern of the observation writer?  (experiment, not product code; results in DESIGN.md section 7)
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/sbb tools/storebench_bulk.cu && /tmp/sbb
//
// Model of one env-step's output: 45 planes of PB bytes (PB = 400 / 1600 / 3600 for L = 10 / 20 / 30) in the
// observation's run structure: S4 Z1 S1 Z3 S1 Z1 S4 Z6 S4 Z16 S4 (S = planes whose content is computed per env:
// map planes, broadcast scalars; Z = zero planes).  One warp per env, 4 warps per CTA, a dependent 1 KB record
// load per env first (what the real kernel cannot avoid).
//   stg      : every plane written with STG.128 by the 32 lanes (today's observation writer)
//   bulk     : S planes staged in shared memory (STS.128, double-buffered) and written with one bulk store per
//              chunk by one elected lane; Z planes bulk-stored from a CTA-wide zero page
//   bulk+fix : the same, then wait_group 0 and two 4-byte STG per lane on top (the sparse one-hots)
//   bulkzero : everything bulk-stored from the zero page (upper bound of the bulk path, no staging work)
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(__cvta_generic_to_global(gdst)), "r"((unsigned)__cvta_generic_to_shared(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// `comp` on the command line: the output buffer comes from a compressible allocation (section 7.2h) instead of cudaMalloc
static bool g_compressible = false;
static unsigned char *alloc_out(size_t bytes)
{
    unsigned char *p = nullptr;
    if (!g_compressible) { CK(cudaMalloc(&p, bytes)); return p; }
    CUmemAllocationProp prop = {};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = 0;
    prop.allocFlags.compressionType = CU_MEM_ALLOCATION_COMP_GENERIC;
    size_t gran = 0;
    cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED);
    const size_t size = (bytes + gran - 1) / gran * gran;
    CUmemGenericAllocationHandle h;
    CUdeviceptr d = 0;
    CUmemAccessDesc acc = {};
    acc.location = prop.location;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    if (cuMemCreate(&h, size, &prop, 0) != CUDA_SUCCESS || cuMemAddressReserve(&d, size, gran, 0, 0) != CUDA_SUCCESS ||
        cuMemMap(d, size, 0, h, 0) != CUDA_SUCCESS || cuMemSetAccess(d, size, &acc, 1) != CUDA_SUCCESS) { printf("compressible allocation failed\n"); exit(1); }
    return reinterpret_cast<unsigned char *>(d);
}

constexpr int kRuns = 11;
__constant__ int kRunPlanes[kRuns] = {4, 1, 1, 3, 1, 1, 4, 6, 4, 16, 4};      // S Z S Z S Z S Z S Z S

extern __shared__ __align__(128) unsigned char smem[];

template <int PB>
__global__ void __launch_bounds__(128) k_stg(float4 *out, const int4 *rec, int n)
{
    int env = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (env >= n) return;
    if (rec == nullptr) smem[threadIdx.x] = 1;
    int4 a = rec[(size_t)env * 64 + lane], b = rec[(size_t)env * 64 + 32 + lane];
    float v = (float)((a.x ^ b.y) & 1);
    constexpr int N4 = 45 * PB / 16;
    float4 *p = out + (size_t)env * N4 + lane;
    float4 x = make_float4(v, v, v, v);
#pragma unroll 8
    for (int k = 0; k < N4 / 32; ++k) p[32 * k] = x;
    if (lane < N4 % 32) p[32 * (N4 / 32)] = x;
}

// SP = planes per staging chunk, ZP = planes in the zero page
template <int PB, int SP, int ZP, bool FIX, bool ZERO_ONLY>
__global__ void __launch_bounds__(128) k_bulk(unsigned char *out, const int4 *rec, int n)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int env = blockIdx.x * 4 + warp;
    unsigned char *zero = smem;                                       // ZP * PB bytes, CTA-wide
    unsigned char *stage = smem + ZP * PB + warp * (2 * SP * PB);     // two buffers of SP planes per warp
    for (int q = threadIdx.x; q < ZP * PB / 16; q += 128) reinterpret_cast<int4 *>(zero)[q] = make_int4(0, 0, 0, 0);
    fence_async();
    __syncthreads();
    if (env >= n) return;
    int4 a = rec[(size_t)env * 64 + lane], b = rec[(size_t)env * 64 + 32 + lane];
    float v = (float)((a.x ^ b.y) & 1);
    unsigned char *o = out + (size_t)env * 45 * PB;
    int plane = 0, buf = 0;
#pragma unroll 1
    for (int r = 0; r < kRuns; ++r) {
        const int np = kRunPlanes[r];
        if (ZERO_ONLY || (r & 1)) {
            if (lane == 0)
                for (int k = 0; k < np; k += ZP) bulk_store(o + (size_t)(plane + k) * PB, zero, (unsigned)(min(ZP, np - k) * PB));
        } else {
            for (int k = 0; k < np; k += SP) {
                const int cp = min(SP, np - k);
                unsigned char *sb = stage + buf * (SP * PB);
                if (lane == 0) bulk_wait_read<1>();                   // the bulk store that read this buffer two chunks ago
                __syncwarp();
                float4 x = make_float4(v + k, v, v, v);
                for (int q = lane; q < cp * PB / 16; q += 32) reinterpret_cast<float4 *>(sb)[q] = x;
                fence_async();
                __syncwarp();
                if (lane == 0) { bulk_store(o + (size_t)(plane + k) * PB, sb, (unsigned)(cp * PB)); bulk_commit(); }
                buf ^= 1;
            }
        }
        plane += np;
    }
    if (lane == 0) {
        bulk_commit();
        if (FIX) bulk_wait<0>(); else bulk_wait_read<0>();
    }
    __syncwarp();
    if (FIX) {
        float *f = reinterpret_cast<float *>(o);
        f[(size_t)(15 + (lane & 3)) * (PB / 4) + lane] = 1.f;
        f[(size_t)(25 + (lane & 7)) * (PB / 4) + lane * 2] = v;
    }
}


// hybrid: zero runs bulk-stored from the CTA zero page (no staging, no waits), computed planes with STG.128.
// EARLY: the zero runs are issued before the dependent record load (they do not depend on the env's state).
// BULK_REC: the 1 KB record arrives by one bulk load + mbarrier instead of two LDG.128 per lane.
template <int PB, int ZP, bool FIX, bool EARLY, bool BULK_REC>
__global__ void __launch_bounds__(128) k_hybrid(unsigned char *out, const int4 *rec, int n)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int env = blockIdx.x * 4 + warp;
    unsigned char *zero = smem;                                       // ZP * PB bytes, CTA-wide
    int4 *recbuf = reinterpret_cast<int4 *>(smem + ZP * PB + warp * 1024);
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem + ZP * PB + 4 * 1024) + warp;
    for (int q = threadIdx.x; q < ZP * PB / 16; q += 128) reinterpret_cast<int4 *>(zero)[q] = make_int4(0, 0, 0, 0);
    if (BULK_REC && lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)));
    }
    fence_async();
    __syncthreads();
    if (env >= n) return;
    unsigned char *o = out + (size_t)env * 45 * PB;
    auto zeros = [&]() {
        if (lane == 0) {
            int plane = 0;
#pragma unroll 1
            for (int r = 0; r < kRuns; ++r) {
                const int np = kRunPlanes[r];
                if (r & 1)
                    for (int k = 0; k < np; k += ZP) bulk_store(o + (size_t)(plane + k) * PB, zero, (unsigned)(min(ZP, np - k) * PB));
                plane += np;
            }
            bulk_commit();
        }
    };
    if (BULK_REC && lane == 0) {
        const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 1024;" ::"r"(b) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 1024, [%2];"
                     ::"r"((unsigned)__cvta_generic_to_shared(recbuf)), "l"(__cvta_generic_to_global(rec + (size_t)env * 64)), "r"(b) : "memory");
    }
    int4 a, b;
    if (!BULK_REC) { a = rec[(size_t)env * 64 + lane]; b = rec[(size_t)env * 64 + 32 + lane]; }
    if (EARLY) zeros();
    if (BULK_REC) {
        const unsigned bb = (unsigned)__cvta_generic_to_shared(bar);
        unsigned done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bb) : "memory");
        a = recbuf[lane]; b = recbuf[32 + lane];
    }
    float v = (float)((a.x ^ b.y) & 1);
    if (!EARLY) zeros();
    {
        float4 x = make_float4(v, v, v, v);
        int plane = 0;
#pragma unroll
        for (int r = 0; r < kRuns; ++r) {
            constexpr int kRP[kRuns] = {4, 1, 1, 3, 1, 1, 4, 6, 4, 16, 4};
            const int np = kRP[r];
            if (!(r & 1)) {
                float4 *p = reinterpret_cast<float4 *>(o + (size_t)plane * PB) + lane;
                constexpr int dummy = 0; (void)dummy;
                const int n4 = np * PB / 16;
#pragma unroll 4
                for (int k = 0; k < n4 / 32; ++k) p[32 * k] = x;
                if (lane < n4 % 32) p[32 * (n4 / 32)] = x;
            }
            plane += np;
        }
    }
    if (lane == 0) { if (FIX) bulk_wait<0>(); else bulk_wait_read<0>(); }
    __syncwarp();
    if (FIX) {
        float *f = reinterpret_cast<float *>(o);
        f[(size_t)(15 + (lane & 3)) * (PB / 4) + lane] = 1.f;
        f[(size_t)(25 + (lane & 7)) * (PB / 4) + lane * 2] = v;
    }
}


// bulkfull: every S plane of the env staged in one per-warp buffer (18 planes), ONE proxy fence, then all 11 runs
// issued back to back by one lane; nothing waits until the end.  STAGE_PLANES * PB bytes per warp.
template <int PB, int ZP, bool FIX>
__global__ void __launch_bounds__(128) k_bulkfull(unsigned char *out, const int4 *rec, int n)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int env = blockIdx.x * 4 + warp;
    unsigned char *zero = smem;
    unsigned char *stage = smem + ZP * PB + warp * (18 * PB);
    for (int q = threadIdx.x; q < ZP * PB / 16; q += 128) reinterpret_cast<int4 *>(zero)[q] = make_int4(0, 0, 0, 0);
    fence_async();
    __syncthreads();
    if (env >= n) return;
    int4 a = rec[(size_t)env * 64 + lane], b = rec[(size_t)env * 64 + 32 + lane];
    float v = (float)((a.x ^ b.y) & 1);
    unsigned char *o = out + (size_t)env * 45 * PB;
    {
        float4 x = make_float4(v, v, v, v);
        constexpr int n4 = 18 * PB / 16;
#pragma unroll 4
        for (int k = 0; k < n4 / 32; ++k) reinterpret_cast<float4 *>(stage)[lane + 32 * k] = x;
        if (lane < n4 % 32) reinterpret_cast<float4 *>(stage)[lane + 32 * (n4 / 32)] = x;
    }
    fence_async();
    __syncwarp();
    if (lane == 0) {
        int plane = 0, sp = 0;
#pragma unroll 1
        for (int r = 0; r < kRuns; ++r) {
            const int np = kRunPlanes[r];
            if (r & 1) {
                for (int k = 0; k < np; k += ZP) bulk_store(o + (size_t)(plane + k) * PB, zero, (unsigned)(min(ZP, np - k) * PB));
            } else {
                bulk_store(o + (size_t)plane * PB, stage + sp * PB, (unsigned)(np * PB));
                sp += np;
            }
            plane += np;
        }
        bulk_commit();
        if (FIX) bulk_wait<0>(); else bulk_wait_read<0>();
    }
    __syncwarp();
    if (FIX) {
        float *f = reinterpret_cast<float *>(o);
        f[(size_t)(15 + (lane & 3)) * (PB / 4) + lane] = 1.f;
        f[(size_t)(25 + (lane & 7)) * (PB / 4) + lane * 2] = v;
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

template <int PB, int SP, int ZP>
void run(int n_envs, int ctas_per_sm_list_n, const int *ctas_per_sm_list)
{
    size_t bytes = (size_t)n_envs * 45 * PB;
    unsigned char *out; int4 *rec;
    out = alloc_out(bytes); CK(cudaMalloc(&rec, (size_t)n_envs * 1024)); CK(cudaMemset(rec, 1, (size_t)n_envs * 1024));
    const double gb = bytes / 1e9;
    const int grid = (n_envs + 3) / 4;
    const size_t need = (size_t)ZP * PB + 4 * 2 * SP * PB;
    printf("PB=%d (%.0f KB/env) n=%d SP=%d ZP=%d staging+zero smem/CTA=%zu\n", PB, 45.0 * PB / 1024, n_envs, SP, ZP, need);
    for (int i = 0; i < ctas_per_sm_list_n; ++i) {
        const int cps = ctas_per_sm_list[i];
        size_t smem_bytes = (size_t)(227 * 1024 / cps - 1024) & ~(size_t)127;     // forces `cps` CTAs per SM
        if (smem_bytes < need) { printf("  %d CTAs/SM: staging does not fit\n", cps); continue; }
        CK(cudaFuncSetAttribute(k_stg<PB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
        CK(cudaFuncSetAttribute(k_bulk<PB, SP, ZP, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
        CK(cudaFuncSetAttribute(k_bulk<PB, SP, ZP, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
        CK(cudaFuncSetAttribute(k_bulk<PB, SP, ZP, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
        float t0 = timeit([&] { k_stg<PB><<<grid, 128, smem_bytes>>>((float4 *)out, rec, n_envs); });
        float t1 = timeit([&] { k_bulk<PB, SP, ZP, false, false><<<grid, 128, smem_bytes>>>(out, rec, n_envs); });
        float t2 = timeit([&] { k_bulk<PB, SP, ZP, true, false><<<grid, 128, smem_bytes>>>(out, rec, n_envs); });
        float t3 = timeit([&] { k_bulk<PB, SP, ZP, false, true><<<grid, 128, smem_bytes>>>(out, rec, n_envs); });
        printf("  %2d CTAs/SM (%2d warps): stg %.4f ms %5.0f GB/s | bulk %.4f ms %5.0f | bulk+fix %.4f ms %5.0f | bulkzero %.4f ms %5.0f\n",
               cps, cps * 4, t0, gb / t0 * 1e3, t1, gb / t1 * 1e3, t2, gb / t2 * 1e3, t3, gb / t3 * 1e3);
        if (smem_bytes >= (size_t)ZP * PB + 4 * 18 * PB) {
            auto hy = [&](auto kern) {
                CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
                return timeit([&] { kern<<<grid, 128, smem_bytes>>>(out, rec, n_envs); });
            };
            float f0 = hy(k_bulkfull<PB, ZP, false>), f1 = hy(k_bulkfull<PB, ZP, true>);
            printf("      bulkfull %.4f ms %5.0f | +fix %.4f %5.0f\n", f0, gb / f0 * 1e3, f1, gb / f1 * 1e3);
        }
        if (false && smem_bytes >= (size_t)ZP * PB + 4 * 1024 + 64) {
            auto hy = [&](auto kern) {
                CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
                return timeit([&] { kern<<<grid, 128, smem_bytes>>>(out, rec, n_envs); });
            };
            float h0 = hy(k_hybrid<PB, ZP, false, false, false>), h1 = hy(k_hybrid<PB, ZP, true, false, false>);
            float h2 = hy(k_hybrid<PB, ZP, true, true, false>), h3 = hy(k_hybrid<PB, ZP, true, false, true>);
            float h4 = hy(k_hybrid<PB, ZP, true, true, true>);
            printf("      hybrid %.4f ms %5.0f | +fix %.4f %5.0f | +fix early %.4f %5.0f | +fix bulkrec %.4f %5.0f | +fix early bulkrec %.4f %5.0f\n",
                   h0, gb / h0 * 1e3, h1, gb / h1 * 1e3, h2, gb / h2 * 1e3, h3, gb / h3 * 1e3, h4, gb / h4 * 1e3);
        }
    }
    // correctness spot check of the bulk path: plane 0 word 0 == v + 0, a zero plane is zero, the fix-ups landed
    CK(cudaMemset(out, 0xff, bytes));
    size_t smem_bytes = (size_t)(227 * 1024 / 6 - 1024) & ~(size_t)127;
    if (smem_bytes < need) smem_bytes = need;
    CK(cudaFuncSetAttribute(k_bulk<PB, SP, ZP, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    k_bulk<PB, SP, ZP, true, false><<<grid, 128, smem_bytes>>>(out, rec, n_envs);
    CK(cudaDeviceSynchronize());
    float *h = (float *)malloc(45 * PB);
    CK(cudaMemcpy(h, out + (size_t)(n_envs - 1) * 45 * PB, 45 * PB, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int p = 0; p < 45; ++p) {
        const bool zero_plane = p == 4 || (p >= 6 && p <= 8) || p == 10 || (p >= 15 && p <= 20) || (p >= 25 && p <= 40);
        for (int q = 0; q < PB / 4; ++q) {
            float x = h[p * (PB / 4) + q];
            bool fix = (p >= 15 && p < 19 && q < 32 && (q & 3) == p - 15) || (p >= 25 && p < 33 && q < 64 && !(q & 1) && ((q / 2) & 7) == p - 25);
            if (fix) continue;
            if (zero_plane ? x != 0.f : !(x == 0.f || x == 1.f || x == 2.f || x == 3.f || x == 4.f)) ++bad;
        }
    }
    printf("  check: %d bad words, fix-up [15][0] = %g\n", bad, h[15 * (PB / 4)]);
    free(h);
    if (!g_compressible) cudaFree(out);
    cudaFree(rec);
}

int main(int argc, char **argv)
{
    g_compressible = argc > 1 && argv[1][0] == 'c';
    CK(cudaFree(0));
    printf("output buffer: %s\n", g_compressible ? "compressible allocation" : "cudaMalloc");
    const int small[] = {6, 5, 4, 3, 2};
    run<400, 4, 16>(65536, 5, small);
    run<400, 4, 6>(65536, 5, small);
    const int mid[] = {6, 3, 2};
    run<1600, 1, 4>(32768, 3, mid);
    const int large[] = {3, 2, 1};
    run<3600, 1, 4>(16384, 3, large);
    run<3600, 1, 16>(16384, 3, large);
    return 0;
}
