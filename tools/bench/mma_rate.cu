// throughput of legacy mma.sync.m16n8k16 (f16 -> f32) per SM on this GPU: warps x independent chains sweep
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
template <int CH>
__global__ void k(float *out, int iters, long long *cyc)
{
    float d[CH][4];
    for (int c = 0; c < CH; c++) for (int i = 0; i < 4; i++) d[c][i] = 0.f;
    uint32_t a0 = 0x3c003c00u + threadIdx.x, a1 = 0x3c003c00u, a2 = 0x38003800u, a3 = 0x3c003c00u, b0 = 0x3c003c00u, b1 = 0x34003400u;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c = 0; c < CH; c++)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    long long t1 = clock64();
    float s = 0;
    for (int c = 0; c < CH; c++) for (int i = 0; i < 4; i++) s += d[c][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main()
{
    float *out; long long *cyc, h;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    for (int warps : {1, 4, 8, 12, 16}) {
        k<1><<<148, warps * 32>>>(out, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("warps %2d chains 1: %.2f cycles/mma/warp, %.2f cycles/mma/SM\n", warps, (double)h / iters, (double)h / iters / warps);
        k<4><<<148, warps * 32>>>(out, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("warps %2d chains 4: %.2f cycles/mma/warp, %.2f cycles/mma/SM\n", warps, (double)h / iters / 4, (double)h / iters / 4 / warps);
    }
    return 0;
}
