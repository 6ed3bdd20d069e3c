// common.cuh -- shared device helpers for the sm_100a decode kernels.
//
// Everything here is written for sm_100a only (B200): 32-wide warps, mbarrier +
// cp.async.bulk (the TMA engine's 1-D bulk-copy path, SASS UBLKCP) for the weight stream,
// 128-bit shared/global accesses, warp-shuffle reductions.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define WT_F32 0
#define WT_F16 1
#define WT_Q4_0 2

namespace llmf90 {

// ------------------------------------------------------------------ row geometry
// Device row format per weight type (cols = contraction length):
//   f32 : cols*4 bytes
//   f16 : cols*2 bytes
//   q4_0: [cols/2 bytes of nibbles: block j at bytes 16j..16j+15, byte i = elements i (low
//          nibble) and i+16 (high nibble)] then [cols/32 f16 scales], padded to 16 bytes.
//         (Same bytes as ggml's 18-byte blocks, split into two 16-byte-aligned planes at
//          upload so a lane fetches one block's 32 weights with a single 128-bit load.)
__host__ __device__ inline size_t row_stride_bytes(int wtype, int cols)
{
    if (wtype == WT_F32) return (size_t)cols * 4;
    if (wtype == WT_F16) return (size_t)cols * 2;
    size_t q = (size_t)cols / 2, s = ((size_t)cols / 32) * 2;
    return q + ((s + 15) & ~(size_t)15);
}
__host__ __device__ inline size_t host_row_bytes(int wtype, int cols)
{
    if (wtype == WT_F32) return (size_t)cols * 4;
    if (wtype == WT_F16) return (size_t)cols * 2;
    return ((size_t)cols / 32) * 18;
}

// q4_0 matrices of the fused streaming kernel use a TILED device format made for mma.sync
// (m16n8k16): 16 rows x 8 blocks (256 columns) form a "group" of 2304 bytes laid out in the order
// the 32 lanes of a warp fetch it (lane = 4 g + t; g = row within 8, t = 32-bit word of a block):
//   chunk 0 (512 B): lane -> 16 B = word t of blocks 0..3 of row g
//   chunk 1        : word t of blocks 4..7 of row g
//   chunk 2, 3     : the same for row g + 8
//   scales (256 B) : lane -> 4 halves d[g][2t], d[g][2t+1], d[g+8][2t], d[g+8][2t+1]
// A row group (16 rows) is its ceil(blocks / 8) groups back to back; rows and blocks past the
// matrix are zero (scale 0).  Same bytes as ggml's 18-byte blocks, only permuted.
constexpr int Q4T_GROUP_BYTES = 2304;
__host__ __device__ inline int q4t_groups(int cols) { return ((cols >> 5) + 7) >> 3; }
__host__ __device__ inline size_t q4t_matrix_bytes(int rows, int cols)
{
    return (size_t)((rows + 15) >> 4) * (size_t)q4t_groups(cols) * Q4T_GROUP_BYTES;
}

#ifdef __CUDACC__
// ------------------------------------------------------------------ small PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
// arrive with a count: a tile group of G warps hands a slot back with G arrivals of weight 12 / G each
__device__ __forceinline__ void mbar_arrive_n(uint64_t *bar, uint32_t n)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(n) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// non-blocking test of a phase (the producer polls with it so that it can keep prefetching meanwhile)
__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Debug watchdog (compile with -DLLMF90_WATCHDOG): a wait that spins for ~1 s prints where it is
// stuck and gives up, so that a protocol bug ends in a diagnosable wrong answer instead of a hang.
#ifdef LLMF90_WATCHDOG
#define LLMF90_WD_DECL long long wd_t0_ = clock64(); bool wd_fired_ = false
#define LLMF90_WD_CHECK(code, a, b)                                                                   \
    if (!wd_fired_ && clock64() - wd_t0_ > 2000000000ll) {                                            \
        wd_fired_ = true;                                                                             \
        printf("WATCHDOG cta %d tid %d code %d a %d b %d\n", (int)blockIdx.x, (int)threadIdx.x, (code), (int)(a), (int)(b)); \
        break;                                                                                        \
    }
#else
#define LLMF90_WD_DECL
#define LLMF90_WD_CHECK(code, a, b)
#endif
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity, int code = 0)
{
    LLMF90_WD_DECL;
    while (!mbar_try_wait(bar, parity)) {
        LLMF90_WD_CHECK(code, (smem_u32(bar) >> 3) & 31, parity)
    }
}
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// 1-D bulk copy global -> shared through the TMA engine, completion on an mbarrier.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                         uint64_t *bar, uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
// fire-and-forget L2 prefetch of a contiguous byte range through the TMA engine (SASS UBLKPF.L2)
__device__ __forceinline__ void bulk_prefetch_l2(const void *src_gmem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ------------------------------------------------------------------ activation vector in smem
// For q4_0 the activation vector is stored XOR-swizzled at float4 granularity inside each
// 32-float block so that lane l reading block (l + 32k) is bank-conflict free.
template <int WT>
__device__ __forceinline__ int xs_index(int e)
{
    if (WT == WT_Q4_0) {
        const int j = e >> 5, i = (e >> 2) & 7, c = e & 3;
        return (j << 5) + (((i ^ (j & 7)) << 2) | c);
    }
    return e;
}

// one element of a device-format row (used for the embedding-row gather only)
__device__ __forceinline__ float row_elem(const uint8_t *row, int wtype, int cols, int e)
{
    if (wtype == WT_F32) return reinterpret_cast<const float *>(row)[e];
    if (wtype == WT_F16) return __half2float(reinterpret_cast<const __half *>(row)[e]);
    const int j = e >> 5, i = e & 31;
    const uint8_t b = row[j * 16 + (i & 15)];
    const int q = (i < 16) ? (b & 0x0f) : (b >> 4);
    const float d = __half2float(*reinterpret_cast<const __half *>(row + (cols >> 1) + 2 * j));
    return d * (float)(q - 8);
}

// ------------------------------------------------------------------ NR-row dot products
// acc[i] += sum over this lane's share of columns of  W[row i][c] * x[c].
// w0 points at row 0 (shared or global memory, 16-byte aligned), rows are `rs` bytes apart,
// xs is the activation vector in shared memory (xs_index<WT> layout).  The 32 lanes of the
// calling warp split the columns; the caller finishes with warp_sum.
template <int WT, int NR>
__device__ __forceinline__ void dot_rows(const uint8_t *__restrict__ w0, size_t rs,
                                         const float *__restrict__ xs, int cols, int lane,
                                         float (&acc)[NR])
{
    const float4 *x4 = reinterpret_cast<const float4 *>(xs);
    if (WT == WT_F32) {
        const int n4 = cols >> 2;
#pragma unroll 2
        for (int j = lane; j < n4; j += 32) {
            const float4 xv = x4[j];
#pragma unroll
            for (int i = 0; i < NR; i++) {
                const float4 wv = *reinterpret_cast<const float4 *>(w0 + i * rs + (size_t)j * 16);
                acc[i] = fmaf(wv.x, xv.x, acc[i]);
                acc[i] = fmaf(wv.y, xv.y, acc[i]);
                acc[i] = fmaf(wv.z, xv.z, acc[i]);
                acc[i] = fmaf(wv.w, xv.w, acc[i]);
            }
        }
    } else if (WT == WT_F16) {
        const int n8 = cols >> 3;
#pragma unroll 2
        for (int j = lane; j < n8; j += 32) {
            const float4 xa = x4[2 * j], xb = x4[2 * j + 1];
#pragma unroll
            for (int i = 0; i < NR; i++) {
                const uint4 wv = *reinterpret_cast<const uint4 *>(w0 + i * rs + (size_t)j * 16);
                const float2 f0 = __half22float2(*reinterpret_cast<const __half2 *>(&wv.x));
                const float2 f1 = __half22float2(*reinterpret_cast<const __half2 *>(&wv.y));
                const float2 f2 = __half22float2(*reinterpret_cast<const __half2 *>(&wv.z));
                const float2 f3 = __half22float2(*reinterpret_cast<const __half2 *>(&wv.w));
                acc[i] = fmaf(f0.x, xa.x, acc[i]);
                acc[i] = fmaf(f0.y, xa.y, acc[i]);
                acc[i] = fmaf(f1.x, xa.z, acc[i]);
                acc[i] = fmaf(f1.y, xa.w, acc[i]);
                acc[i] = fmaf(f2.x, xb.x, acc[i]);
                acc[i] = fmaf(f2.y, xb.y, acc[i]);
                acc[i] = fmaf(f3.x, xb.z, acc[i]);
                acc[i] = fmaf(f3.y, xb.w, acc[i]);
            }
        }
    } else {
        // q4_0: lane <-> block.  The 32 activations of the block stay in registers across the
        // NR rows.  Nibble q is turned into the float 16+q with one byte-permute
        // (0x41800000 | q<<19), so  sum (q-8)*x = sum (16+q)*x - 24*sum x  needs one PRMT and
        // one FMA per weight; the scale d multiplies once per block.
        const int nblk = cols >> 5;
        const size_t qbytes = (size_t)cols >> 1;
        for (int j = lane; j < nblk; j += 32) {
            float xr[32];
            float xsum = 0.f;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const float4 v = x4[(j << 3) + (i ^ (j & 7))];
                xr[4 * i + 0] = v.x; xr[4 * i + 1] = v.y; xr[4 * i + 2] = v.z; xr[4 * i + 3] = v.w;
                xsum += (v.x + v.y) + (v.z + v.w);
            }
            const float corr = -24.f * xsum;
#pragma unroll
            for (int i = 0; i < NR; i++) {
                const uint8_t *row = w0 + i * rs;
                const uint4 qv = *reinterpret_cast<const uint4 *>(row + (size_t)j * 16);
                const float d = __half2float(*reinterpret_cast<const __half *>(row + qbytes + 2 * j));
                const uint32_t w[4] = {qv.x, qv.y, qv.z, qv.w};
                float s0 = corr, s1 = 0.f;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint32_t lo = ((w[k] << 3) & 0x78787878u) | 0x80808080u;
                    const uint32_t hi = ((w[k] >> 1) & 0x78787878u) | 0x80808080u;
                    s0 = fmaf(__uint_as_float(prmt(lo, 0x41000000u, 0x7044u)), xr[4 * k + 0], s0);
                    s1 = fmaf(__uint_as_float(prmt(lo, 0x41000000u, 0x7144u)), xr[4 * k + 1], s1);
                    s0 = fmaf(__uint_as_float(prmt(lo, 0x41000000u, 0x7244u)), xr[4 * k + 2], s0);
                    s1 = fmaf(__uint_as_float(prmt(lo, 0x41000000u, 0x7344u)), xr[4 * k + 3], s1);
                    s0 = fmaf(__uint_as_float(prmt(hi, 0x41000000u, 0x7044u)), xr[16 + 4 * k + 0], s0);
                    s1 = fmaf(__uint_as_float(prmt(hi, 0x41000000u, 0x7144u)), xr[16 + 4 * k + 1], s1);
                    s0 = fmaf(__uint_as_float(prmt(hi, 0x41000000u, 0x7244u)), xr[16 + 4 * k + 2], s0);
                    s1 = fmaf(__uint_as_float(prmt(hi, 0x41000000u, 0x7344u)), xr[16 + 4 * k + 3], s1);
                }
                acc[i] = fmaf(d, s0 + s1, acc[i]);
            }
        }
    }
}

#endif  // __CUDACC__

}  // namespace llmf90
