// Store-pattern microbenchmark: how long must a contiguous run be for scattered FP64 stores to stream to HBM at full rate?
// Emulates the batched generator's output: NB packed upper triangles (dimension n, column-major), a CTA owns a tile of
// R rows x 8 columns and writes it for every batch element (chunks of 16 elements, like the DMMA kernels), tiles walked
// 16 column tiles at a time.  Only stores, no arithmetic.   nvcc -O3 -arch=sm_100a -o store_pattern.bin store_pattern.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int R>
__global__ void __launch_bounds__(256) pattern(double* out, long long n, int nb, long long stride, int perCta)
{
    const long long nColTiles = (n + 7) / 8;
    const long long colTile = (long long)blockIdx.y * 16 + (blockIdx.x % 16);
    const long long rowTile = blockIdx.x / 16;
    if(colTile >= nColTiles || rowTile * R > colTile * 8 + 7) return;
    const int b0 = blockIdx.z * perCta;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int LR = R < 32 ? R : 32;            // lanes along the run
    constexpr int CPI = 32 / LR;                   // columns per instruction
    const int il = lane % LR, jq = lane / LR;
    for(int c = 0; c < perCta; c += 16)
    {
        // warp w writes elements c + 2w, c + 2w + 1
        for(int bb = 0; bb < 2; ++bb)
        {
            const int b = b0 + c + 2 * warp + bb;
            if(b >= nb) continue;
            double* base = out + (long long)b * stride;
            for(int jj = 0; jj < 8; jj += CPI)
            {
                const long long j = colTile * 8 + jj + jq;
                if(j >= n) continue;
                double* col = base + j * (j + 1) / 2;
#pragma unroll
                for(int r = 0; r < R; r += LR)
                {
                    const long long i = rowTile * R + r + il;
                    if(i <= j) col[i] = 1.0;
                }
            }
        }
        __syncthreads();
    }
}

int main(int argc, char** argv)
{
    const long long n = argc > 1 ? atoll(argv[1]) : 9216;
    const int nb = argc > 2 ? atoi(argv[2]) : 128;
    const long long stride = n * (n + 1) / 2;
    double* out;
    if(cudaMalloc(&out, sizeof(double) * stride * nb) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](int R, int perCta) {
        const long long nColTiles = (n + 7) / 8, nRowTiles = (n + R - 1) / R;
        dim3 grid((unsigned)(nRowTiles * 16), (unsigned)((nColTiles + 15) / 16), (unsigned)((nb + perCta - 1) / perCta));
        float best = 1e9f;
        for(int rep = 0; rep < 3; ++rep)
        {
            cudaEventRecord(e0);
            switch(R)
            {
                case 8: pattern<8><<<grid, 256>>>(out, n, nb, stride, perCta); break;
                case 16: pattern<16><<<grid, 256>>>(out, n, nb, stride, perCta); break;
                case 32: pattern<32><<<grid, 256>>>(out, n, nb, stride, perCta); break;
                case 64: pattern<64><<<grid, 256>>>(out, n, nb, stride, perCta); break;
                case 128: pattern<128><<<grid, 256>>>(out, n, nb, stride, perCta); break;
            }
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if(ms < best) best = ms;
        }
        printf("run %4d rows (%5d B), %3d elements per CTA: %8.2f ms  %7.0f GB/s  %s\n", R, R * 8, perCta, best,
               sizeof(double) * (double)stride * nb / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    };
    for(int perCta : {128, 16})
        for(int R : {8, 16, 32, 64, 128})
            run(R, perCta);
    return 0;
}
