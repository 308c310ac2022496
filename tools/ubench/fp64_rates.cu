// Micro-benchmark: fp64 throughput of DFMA vs DMMA (mma.sync m8n8k4 / m16n8k8 / m16n8k16) on one B200.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_rates fp64_rates.cu
#include <cuda_runtime.h>
#include <cstdio>

__global__ void k_dfma(double *out, int iters)
{
    double a = threadIdx.x * 1e-3, b = 1.0000001, acc[8] = {0, 1, 2, 3, 4, 5, 6, 7};
    for (int i = 0; i < iters; i++)
#pragma unroll
        for (int r = 0; r < 8; r++) acc[r] = fma(a, b, acc[r]);
    double s = 0; for (int r = 0; r < 8; r++) s += acc[r];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma884(double *out, int iters)
{
    double a = threadIdx.x * 1e-3, b = 1.0000001;
    double c[4][2] = {};
    for (int i = 0; i < iters; i++)
#pragma unroll
        for (int r = 0; r < 4; r++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[r][0]), "+d"(c[r][1]) : "d"(a), "d"(b));
    double s = 0; for (int r = 0; r < 4; r++) s += c[r][0] + c[r][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma16816(double *out, int iters)
{
    double a[8], b[4];
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-3 + i;
    for (int i = 0; i < 4; i++) b[i] = 1.0000001 + i;
    double c[2][4] = {};
    for (int i = 0; i < iters; i++)
#pragma unroll
        for (int r = 0; r < 2; r++)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                         : "+d"(c[r][0]), "+d"(c[r][1]), "+d"(c[r][2]), "+d"(c[r][3])
                         : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
    double s = 0; for (int r = 0; r < 2; r++) s += c[r][0] + c[r][1] + c[r][2] + c[r][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class K>
static void run(const char *name, K kern, double fma_per_thread_iter, int iters)
{
    double *out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(double));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<<<148 * 8, 256>>>(out, iters);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    kern<<<148 * 8, 256>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fma = fma_per_thread_iter * iters * 148.0 * 8 * 256;
    printf("%-12s %8.3f ms  %7.2f TFLOP/s  (%s)\n", name, ms, 2.0 * fma / ms * 1e-9, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main()
{
    const int iters = 20000;
    run("dfma", k_dfma, 8.0, iters);
    run("dmma.m8n8k4", k_dmma884, 4.0 * 256 / 32, iters);
    run("dmma.m16n8k16", k_dmma16816, 2.0 * 2048 / 32, iters);
    return 0;
}
