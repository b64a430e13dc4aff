// Loop-only microbenchmark of the T,Q,U Clenshaw step (no pixel staging, no rotation, no stores):
// what fraction of the FP64 peak does the recurrence loop itself reach, as a function of columns per
// thread (R), coefficient delivery (static kernel-parameter table vs shared memory) and occupancy?
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include "../cosmopp_b200/csrc/kernels.cuh"

using namespace cmg;

template <int R, bool STATIC, int MINB>
__global__ void __launch_bounds__(256, MINB)
loopKernel(const __grid_constant__ TquStaticTable T, int entryChunk, const double* __restrict__ tabGlobal, int lmax,
           int passes, double* sink)
{
    extern __shared__ double4 smemTab[];
    if(!STATIC)
    {
        for(int k = threadIdx.x; k < 2 * (lmax + 1); k += 256)
            smemTab[k] = reinterpret_cast<const double4*>(tabGlobal)[k];
        __syncthreads();
    }
    double acc = 0;
    for(int p = 0; p < passes; ++p)
    {
        TquState<R> st;
#pragma unroll
        for(int r = 0; r < R; ++r)
        {
            st.x2[r] = 1e-3 * (threadIdx.x + 7 * r + p) - 0.9;
            st.tt1[r] = st.tt2[r] = st.te1[r] = st.te2[r] = st.pp1[r] = st.pp2[r] = st.mm1[r] = st.mm2[r] = 0.0;
        }
        if(STATIC)
            tquClenshawStatic<R>(st, T, entryChunk);
        else
            tquClenshawShared<R>(st, smemTab, lmax);
#pragma unroll
        for(int r = 0; r < R; ++r)
            acc += st.tt1[r] + st.te1[r] + st.pp1[r] + st.mm1[r];
    }
    if(acc == 123.456)
        sink[0] = acc;
}

static TquStaticTable hostT;

template <int R, bool STATIC, int MINB>
void run(const char* name, int lmax, int ctasPerSm, const double* dTab, double* sink, int padSmem)
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * ctasPerSm * 4;
    const int passes = 8 / R * 4;
    auto k = loopKernel<R, STATIC, MINB>;
    const size_t smem = (STATIC ? 0 : sizeof(double4) * 2 * (lmax + 1)) + padSmem;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, 256, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    const int entry = (PQ_STATIC_STEPS + 1 - lmax) / PQ_STATIC_CHUNK;
    for(int rep = 0; rep < 4; ++rep)
    {
        cudaEventRecord(e0);
        k<<<blocks, 256, smem>>>(hostT, entry, dTab, lmax, passes, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if(rep) best = ms < best ? ms : best;
    }
    cudaError_t e = cudaGetLastError();
    const double steps = STATIC ? (PQ_STATIC_STEPS - entry * PQ_STATIC_CHUNK) : (lmax - 1);
    const double flop = 2.0 * blocks * 256.0 * passes * R * (steps * 10 + 2);
    printf("%-28s lmax %3d occ %d CTAs/SM: %8.3f ms %7.2f TFLOP/s executed  (%s)\n", name, lmax, occ, best, flop / (best * 1e-3) / 1e12,
           cudaGetErrorString(e));
}

int main()
{
    for(int i = 0; i < 2 * PQ_STATIC_STEPS + 1; ++i)
        hostT.s[i] = make_double4(1e-3, 2e-3, 3e-3, 4e-3 * ((i & 1) ? 1 : -1) * 0.999);
    for(int i = 0; i < PQ_STATIC_STEPS; ++i)
        hostT.s[2 * i + 1] = make_double4(-0.99, -0.98, -0.97, 0.01);
    std::vector<double4> tab(2 * 1025);
    for(size_t i = 0; i < tab.size(); ++i)
        tab[i] = (i & 1) ? make_double4(-0.99, -0.98, -0.97, 0.01) : make_double4(1e-3, 2e-3, 3e-3, 4e-3);
    double* dTab; cudaMalloc(&dTab, sizeof(double4) * tab.size());
    cudaMemcpy(dTab, tab.data(), sizeof(double4) * tab.size(), cudaMemcpyHostToDevice);
    double* sink; cudaMalloc(&sink, 8);
    for(int lmax : {192, 47})
    {
        run<1, true, 3>("static R=1 (80 regs)", lmax, 3, dTab, sink, 0);
        run<1, true, 4>("static R=1 (64 regs)", lmax, 4, dTab, sink, 0);
        run<2, true, 2>("static R=2 (128 regs)", lmax, 2, dTab, sink, 0);
        run<2, true, 3>("static R=2 (80 regs)", lmax, 3, dTab, sink, 0);
        run<2, true, 4>("static R=2 (64 regs)", lmax, 4, dTab, sink, 0);
        run<4, true, 2>("static R=4 (128 regs)", lmax, 2, dTab, sink, 0);
        run<2, false, 2>("shared R=2 (128 regs)", lmax, 2, dTab, sink, 0);
        run<2, false, 3>("shared R=2 (80 regs)", lmax, 3, dTab, sink, 0);
        run<4, false, 2>("shared R=4 (128 regs)", lmax, 2, dTab, sink, 0);
        run<4, false, 1>("shared R=4 (255 regs)", lmax, 1, dTab, sink, 0);
        run<2, false, 2>("shared R=2 occ1 (pad smem)", lmax, 1, dTab, sink, 120 * 1024);
    }
    return 0;
}
