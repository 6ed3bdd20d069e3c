// Issue cost of cp.async.bulk (UBLKCP) from one producer thread, in SM cycles per copy.
//   mode 0: lane 0 under a divergent branch (the compiler wraps every copy in an ELECT / R2UR.BROADCAST loop)
//   mode 1: the whole warp runs the loop, the copy sits under elect.sync
//   mode 2: the whole warp runs the loop, lane i < nseg issues segment i (one instruction, divergent operands)
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/bulk_issue tools/ubench/bulk_issue.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t su32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t pol)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}
__device__ __forceinline__ bool elect_one()
{
    uint32_t p;
    asm volatile("{ .reg .pred P; elect.sync _|P, 0xffffffff; selp.u32 %0, 1, 0, P; }" : "=r"(p));
    return p != 0;
}
__global__ void k(const uint8_t *src, size_t stride, int seg, int nseg, int nstage, int mode, long long *out)
{
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ uint64_t bar;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    const uint8_t *base = src + (size_t)blockIdx.x * nstage * nseg * stride;
    long long t0 = 0, t1 = 0;
    if (warp == 1) {
        if (lane == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(&bar)), "r"(seg * nseg * nstage) : "memory");
        __syncwarp();
        t0 = clock64();
        if (mode == 0) {
            if (lane == 0) {
#pragma unroll 1
                for (int s = 0; s < nstage; s++) {
                    const uint8_t *p = base + (size_t)s * nseg * stride;
#pragma unroll 1
                    for (int i = 0; i < nseg; i++) g2s(su32(sm) + i * seg, p + (size_t)i * stride, seg, su32(&bar), pol);
                }
            }
        } else if (mode == 1) {
#pragma unroll 1
            for (int s = 0; s < nstage; s++) {
                const uint8_t *p = base + (size_t)s * nseg * stride;
#pragma unroll 1
                for (int i = 0; i < nseg; i++)
                    if (elect_one()) g2s(su32(sm) + i * seg, p + (size_t)i * stride, seg, su32(&bar), pol);
            }
        } else {
#pragma unroll 1
            for (int s = 0; s < nstage; s++) {
                const uint8_t *p = base + (size_t)s * nseg * stride;
                if (lane < nseg) g2s(su32(sm) + lane * seg, p + (size_t)lane * stride, seg, su32(&bar), pol);
                __syncwarp();
            }
        }
        t1 = clock64();
    }
    // everyone waits for the bytes
    uint32_t ok = 0;
    while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(su32(&bar)) : "memory");
    long long t2 = clock64();
    if (warp == 1 && lane == 0) { out[2 * blockIdx.x] = t1 - t0; out[2 * blockIdx.x + 1] = t2 - t0; }
}
int main()
{
    const size_t bytes = 1ull << 30;
    uint8_t *src; long long *out, h[296];
    cudaMalloc(&src, bytes); cudaMemset(src, 1, bytes); cudaMalloc(&out, sizeof(h));
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int seg : {256, 1024, 4096, 8192})
        for (int nseg : {1, 4, 8})
            for (int mode = 0; mode < 3; mode++) {
                if (seg * nseg > 64 * 1024) continue;
                const int nstage = 64;
                for (int rep = 0; rep < 2; rep++) k<<<148, 64, 200 * 1024>>>(src, seg, seg, nseg, nstage, mode, out);  // contiguous rows: stride = seg
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
                double iss = 0, tot = 0;
                for (int i = 0; i < 148; i++) { iss += h[2 * i]; tot += h[2 * i + 1]; }
                printf("seg %5d nseg %d mode %d: issue %.1f cycles/copy, %.1f cycles/stage; all bytes landed after %.0f cycles (%.1f B/clk/SM)\n", seg, nseg, mode,
                       iss / 148 / (nstage * nseg), iss / 148 / nstage, tot / 148, (double)seg * nseg * nstage / (tot / 148));
            }
    return 0;
}
