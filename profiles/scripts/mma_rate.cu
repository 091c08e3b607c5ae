// mma_rate.cu — microbenchmark (profiling aid, not product): issue rate of the legacy mma.sync TF32 tiles and of
// packed FFMA2 on one B200, alone and interleaved, to decide where the decimator's q-tap FIR half should run.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_k8(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_k4(float (&d)[4], const unsigned (&a)[2], const unsigned (&b)[1])
{
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(b[0]));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c)
{
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b),
                       rc = *reinterpret_cast<unsigned long long *>(&c), rd;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2 *>(&rd);
}

// MODE 0: k8 only, 1: k4 only, 2: FFMA2 only, 3: k8 + FFMA2 interleaved 1:2, 4: bf16 k16 only, 5: FFMA only
template <int MODE>
__global__ void __launch_bounds__(256) rate_kernel(float *out, int iters, unsigned seed)
{
    constexpr int NA = 8;                      // independent accumulator chains
    float d[NA][4];
    float2 f[NA];
    float s[NA];
    unsigned a[4] = {seed + threadIdx.x, seed * 3 + 1, seed * 5 + 2, seed * 7 + 3}, b[2] = {seed * 11 + 4, seed * 13 + 5};
    unsigned a2[2] = {a[0], a[1]}, b1[1] = {b[0]};
    for (int i = 0; i < NA; ++i) { d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0.f; f[i] = make_float2(i, -i); s[i] = (float)i; }
    const float2 m = make_float2(1.0001f, 0.9999f), c = make_float2(1e-3f, -1e-3f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NA; ++i) {
            if (MODE == 0 || MODE == 3) mma_k8(d[i], a, b);
            if (MODE == 1) mma_k4(d[i], a2, b1);
            if (MODE == 4) mma_bf16(d[i], a, b);
            if (MODE == 2 || MODE == 3) { f[i] = ffma2(f[i], m, c); f[(i + 4) % NA] = ffma2(f[(i + 4) % NA], m, c); }
            if (MODE == 5) { s[i] = fmaf(s[i], 1.0001f, 1e-3f); s[(i + 4) % NA] = fmaf(s[(i + 4) % NA], 0.9999f, 1e-3f); }
        }
    }
    float acc = 0.f;
    for (int i = 0; i < NA; ++i) acc += d[i][0] + d[i][1] + d[i][2] + d[i][3] + f[i].x + f[i].y + s[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE> static void run(const char *name, double ops_per_iter_per_warp, float *out)
{
    const int iters = 4096, blocks = 148 * 4, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    rate_kernel<MODE><<<blocks, threads>>>(out, 64, 1);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    rate_kernel<MODE><<<blocks, threads>>>(out, iters, 1);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double warps = (double)blocks * threads / 32, total = warps * iters * ops_per_iter_per_warp;
    const double per_smsp_per_cycle = total / (ms * 1e-3) / (148.0 * 4) / (clk_khz * 1e3);
    printf("%-28s %8.3f ms  %.4f warp-instr/cycle/SMSP at %d MHz nominal (%.2f cycles each)\n", name, ms, per_smsp_per_cycle,
           clk_khz / 1000, 1.0 / per_smsp_per_cycle);
}

int main()
{
    float *out;
    cudaMalloc(&out, 148 * 4 * 256 * sizeof(float));
    run<0>("mma m16n8k8 tf32", 8, out);
    run<1>("mma m16n8k4 tf32", 8, out);
    run<4>("mma m16n8k16 bf16", 8, out);
    run<2>("ffma2", 16, out);
    run<5>("ffma", 16, out);
    run<3>("k8 + 2 ffma2 (count all 3)", 24, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
