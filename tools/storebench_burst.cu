// storebench_burst.cu -- does it pay to fetch the env records in BURSTS instead of one by one?
// (experiment, not product code; results in DESIGN.md section 7)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/sbb tools/storebench_burst.cu && /tmp/sbb
// storebench_read.cu: a dependent record read that hits L2 is free (0.168 ms), the same read from DRAM costs 0.015 ms
// for the first 256 B and 0.013 ms per further KB -- DRAM reads sprinkled into a write-saturated HBM.
// storebench_l2.cu: the whole record array cannot be kept in L2 (protected capacity 35-40 MB).
// Here the reads are made bursty: every G-th CTA prefetches the records of the CTAs [b + D, b + D + G) into L2 in one
// go (prefetch.global.L2, optionally evict_last), so that DRAM sees one long read burst per G CTAs and every env's own
// read is an L2 hit; only (D + G) CTAs' worth of records has to survive in L2 at a time.  The write-back can demote
// the line again (evict_first).
//   model: one warp per env, S bytes of record read (stride STRIDE), 18,000 B observation written, WB bytes written back.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)
constexpr int kN4 = 1125;

__device__ __forceinline__ unsigned long long policy_first()
{
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void st_hint(int4 *a, int4 v, unsigned long long pol)
{
    asm volatile("st.global.L2::cache_hint.v4.s32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
}

// PF: 0 none, 1 prefetch.global.L2, 2 prefetch.global.L2::evict_last;  WBP: 0 plain write-back, 1 evict_first
template <int S, int WB, int STRIDE, int PF, int WBP, bool FIX>
__global__ void __launch_bounds__(128) k(float4 *out, int4 *rec, int n, int G, int D)
{
    extern __shared__ unsigned char smem[];
    const int cta = blockIdx.x;
    int env = cta * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (rec == nullptr) smem[threadIdx.x] = 1;
    if (PF != 0 && cta % G == 0) {
        // the records of CTAs [cta + D, cta + D + G); CTA 0 also covers the cold start [0, D)
        const int first = cta == 0 ? 0 : (cta + D) * 4;
        const int last = min((cta + D + G) * 4, n);
        constexpr int kLines = (S + 127) / 128;
        const int total = (last - first) * kLines;
        for (int i = threadIdx.x; i < total; i += 128) {
            const int e = first + i / kLines, l = i % kLines;
            const char *a = reinterpret_cast<const char *>(rec) + (size_t)e * STRIDE + (size_t)l * 128;
            if (PF == 1) asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
            else asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(a));
        }
    }
    if (env >= n) return;
    int4 *r = reinterpret_cast<int4 *>(reinterpret_cast<char *>(rec) + (size_t)env * STRIDE);
    int acc = 0;
#pragma unroll
    for (int q = 0; q < (S + 511) / 512; ++q)
        if (lane + 32 * q < S / 16) {
            int4 a = r[lane + 32 * q];
            acc ^= a.x ^ a.y;
        }
    acc = __reduce_xor_sync(0xffffffffu, acc);
    float v = (float)(acc & 1);
    float4 *p = out + (size_t)env * kN4 + lane;
    float4 x = make_float4(v, v, v, v);
#pragma unroll 8
    for (int q = 0; q < kN4 / 32; ++q) p[32 * q] = x;
    if (lane < kN4 % 32) p[32 * (kN4 / 32)] = x;
    if (FIX) {          // the sparse one-hots: 4-byte stores on top of lines the dense pass has just written
        __syncwarp();
        float *f = reinterpret_cast<float *>(out + (size_t)env * kN4);
        f[(15 + (lane & 3)) * 100 + lane] = 1.f;
        f[(25 + (lane & 7)) * 100 + lane * 2] = v;
    }
    if (lane < WB / 16) {
        int4 w = make_int4(acc + 1, acc, 1, 1);
        if (WBP == 1) st_hint(r + lane, w, policy_first()); else r[lane] = w;
    }
}

template <typename F> float timeit(F f, cudaStream_t s, int iters = 30)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 5; ++i) f();
    CK(cudaGetLastError());
    cudaEventRecord(a, s);
    for (int i = 0; i < iters; ++i) f();
    cudaEventRecord(b, s);
    CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / iters;
}

template <int S, int WB, int STRIDE>
void sweep(cudaStream_t s, float4 *out, int4 *rec, int n)
{
    const size_t smem_bytes = (size_t)(227 * 1024 / 7 - 1024) & ~(size_t)127;      // 7 CTAs = 28 warps per SM
    const int grid = n / 4;
    auto run = [&](auto kern, int G, int D) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
        return timeit([&] { kern<<<grid, 128, smem_bytes, s>>>(out, rec, n, G, D); }, s);
    };
    printf("S=%d WB=%d stride=%d (%.1f MB of record lines touched)\n", S, WB, STRIDE, (double)n * ((S + 127) / 128 * 128) / 1e6);
    printf("  no prefetch: plain %.4f ms | write-back evict_first %.4f ms\n", run(k<S, WB, STRIDE, 0, 0, true>, 1, 0),
           run(k<S, WB, STRIDE, 0, 1, true>, 1, 0));
    for (int D : {1100, 2200, 4400})
        for (int G : {16, 64, 256, 1024, 4096}) {
            float a = run(k<S, WB, STRIDE, 1, 0, true>, G, D), b = run(k<S, WB, STRIDE, 1, 1, true>, G, D);
            float c = run(k<S, WB, STRIDE, 2, 0, true>, G, D), d = run(k<S, WB, STRIDE, 2, 1, true>, G, D);
            printf("  D=%5d G=%5d (burst %6.2f MB): prefetch %.4f | + wb first %.4f | prefetch evict_last %.4f | + wb first %.4f ms\n", D, G,
                   (double)G * 4 * ((S + 127) / 128 * 128) / 1e6, a, b, c, d);
        }
    // per-CTA prefetch (G = 1) at the same distances: is it the burst or just the distance?
    for (int D : {1100, 4400})
        printf("  D=%5d G=    1: prefetch %.4f | prefetch evict_last + wb first %.4f ms\n", D, run(k<S, WB, STRIDE, 1, 0, true>, 1, D),
               run(k<S, WB, STRIDE, 2, 1, true>, 1, D));
}

int main()
{
    const int n = 65536;
    cudaStream_t s;
    CK(cudaStreamCreate(&s));
    float4 *out; int4 *rec;
    CK(cudaMalloc(&out, (size_t)n * kN4 * 16)); CK(cudaMalloc(&rec, (size_t)n * 2560)); CK(cudaMemset(rec, 1, (size_t)n * 2560));
    sweep<1024, 384, 2560>(s, out, rec, n);
    sweep<512, 256, 2560>(s, out, rec, n);
    sweep<1024, 384, 1024>(s, out, rec, n);
    return 0;
}
