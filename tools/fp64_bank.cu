// Microbenchmark: DFMA issue rate on sm_100a as a function of how many distinct vector-register
// operands each instruction reads (register-file bank bandwidth vs the 2-cycle FP64 pipe).
#include <cstdio>
#include <cuda_runtime.h>

constexpr int CH = 8;
constexpr int ITERS = 2048;

// variant 0: acc = fma(acc, U, Rshared)   (1 reg + uniform + reused reg)
// variant 1: acc_c = fma(acc_c, y_c, x_c) (3 distinct regs per instruction)
// variant 2: acc_c = fma(acc_c, y, x_c)   (y shared across consecutive instr -> reuse, 2 distinct)
// variant 3: acc_c = fma(acc_c, y_c, U)   (2 distinct regs + uniform)
// variant 4: acc_c = fma(acc_c, y, x)     (y,x shared non-uniform regs -> reuse both)
template <int V>
__global__ void __launch_bounds__(256) k(double* sink, double ux, double uy)
{
    double acc[CH], x[CH], y[CH];
    const double t = threadIdx.x * 1e-9;
#pragma unroll
    for(int c = 0; c < CH; ++c)
    {
        acc[c] = ux + c + t;
        x[c] = ux + t * (c + 1);
        y[c] = uy - t * (c + 2);
    }
#pragma unroll 1
    for(int it = 0; it < ITERS / 4; ++it)
    {
#pragma unroll
        for(int u = 0; u < 4; ++u)
#pragma unroll
            for(int c = 0; c < CH; ++c)
            {
                if(V == 0) acc[c] = fma(acc[c], uy, ux);
                if(V == 1) acc[c] = fma(acc[c], y[c], x[c]);
                if(V == 2) acc[c] = fma(acc[c], y[0], x[c]);
                if(V == 3) acc[c] = fma(acc[c], y[c], ux);
                if(V == 4) acc[c] = fma(acc[c], y[0], x[0]);
            }
    }
    double s = 0;
#pragma unroll
    for(int c = 0; c < CH; ++c) s += acc[c] + x[c] + y[c];
    if(s == 123.456) sink[0] = s;
}

template <int V>
void run(const char* name, int blocks)
{
    double* sink;
    cudaMalloc(&sink, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for(int rep = 0; rep < 5; ++rep)
    {
        cudaEventRecord(e0);
        k<V><<<blocks, 256>>>(sink, 1.0000001, 0.9999999);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if(rep) best = ms < best ? ms : best;
    }
    const double flop = 2.0 * blocks * 256.0 * CH * ITERS;
    printf("%-40s %8.3f ms  %7.2f TFLOP/s\n", name, best, flop / (best * 1e-3) / 1e12);
    cudaFree(sink);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    for(int occ : {2, 4, 8})
    {
        const int blocks = p.multiProcessorCount * occ * 8;
        printf("-- grid = %d SMs x %d x 8 blocks of 256\n", p.multiProcessorCount, occ);
        run<0>("v0 acc*U+Rreuse", blocks);
        run<1>("v1 acc*y_c+x_c (3 distinct regs)", blocks);
        run<2>("v2 acc*y0+x_c (2 distinct + reuse)", blocks);
        run<3>("v3 acc*y_c+U (2 distinct + uniform)", blocks);
        run<4>("v4 acc*y0+x0 (1 distinct + 2 reuse)", blocks);
    }
    return 0;
}
