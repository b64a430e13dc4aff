// FP64 tensor-core (DMMA, mma.sync m8n8k4 / m16n8k4 / m16n8k8 / m16n8k16 f64) throughput on sm_100a via the legacy mma.sync path.
#include <cstdio>
#include <cuda_runtime.h>

template <int SHAPE>
__global__ void __launch_bounds__(256) k(double* sink, int iters)
{
    double a[8], b[4];
    double c[8][4];
    for(int i = 0; i < 8; ++i) a[i] = 1.0 + threadIdx.x * 1e-6 + i;
    for(int i = 0; i < 4; ++i) b[i] = 0.5 + threadIdx.x * 1e-7 + i;
    for(int t = 0; t < 8; ++t) for(int i = 0; i < 4; ++i) c[t][i] = 0.0;
    for(int it = 0; it < iters; ++it)
    {
#pragma unroll
        for(int t = 0; t < 8; ++t)
        {
            if(SHAPE == 884)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[t][0]), "+d"(c[t][1]) : "d"(a[0]), "d"(b[0]));
            if(SHAPE == 1684)
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+d"(c[t][0]), "+d"(c[t][1]), "+d"(c[t][2]), "+d"(c[t][3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
            if(SHAPE == 1688)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+d"(c[t][0]), "+d"(c[t][1]), "+d"(c[t][2]), "+d"(c[t][3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
            if(SHAPE == 16816)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                             : "+d"(c[t][0]), "+d"(c[t][1]), "+d"(c[t][2]), "+d"(c[t][3])
                             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
        }
    }
    double s = 0;
    for(int t = 0; t < 8; ++t) for(int i = 0; i < 4; ++i) s += c[t][i];
    if(s == 123.456) sink[0] = s;
}

template <int SHAPE>
void run(const char* name, double flopPerMma)
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double* sink; cudaMalloc(&sink, 8);
    const int blocks = p.multiProcessorCount * 8, iters = 2000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for(int rep = 0; rep < 4; ++rep)
    {
        cudaEventRecord(e0);
        k<SHAPE><<<blocks, 256>>>(sink, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if(rep) best = ms < best ? ms : best;
    }
    const double flop = flopPerMma * 8.0 * iters * (blocks * 256.0 / 32.0);
    printf("%-12s %8.3f ms  %7.2f TFLOP/s (%s)\n", name, best, flop / (best * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    run<884>("m8n8k4", 2.0 * 8 * 8 * 4);
    run<1684>("m16n8k4", 2.0 * 16 * 8 * 4);
    run<1688>("m16n8k8", 2.0 * 16 * 8 * 8);
    run<16816>("m16n8k16", 2.0 * 16 * 8 * 16);
    return 0;
}
