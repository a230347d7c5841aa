// cta_group::2 (CTA pair, M = 256) tcgen05.mma on B200: syntax / semantics check + dispatch rate. Two CTAs of a cluster each hold 128 rows
// of A ([128 x 64] fp16, K-major SW128) and HALF of B ([N/2 x 64]); the leader CTA issues one M256 x N x K16 instruction per K step, the
// accumulator rows 0-127 land in the leader's TMEM and rows 128-255 in the peer's. B = identity, so D must reproduce each CTA's own A.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench_umma2 tools/ubench_umma2.cu && build/ubench_umma2
#include <cstdint>
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
    uint64_t d = 0;
    d |= (uint64_t)((a & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float a_value(int cta, int r, int k) { return (float)((cta * 17 + r * 5 + k * 3) % 53 - 26); }

template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k_umma2(long long *out, int *wrong, int iters) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    uint8_t *sA = smem, *sB = smem + 16384;
    const int tid = threadIdx.x, warp = tid >> 5;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    for (int i = tid; i < 128 * 8; i += 128) {   // A: this CTA's 128 rows
        const int r = i >> 3, c8 = i & 7;
        __half v[8];
        for (int e = 0; e < 8; ++e) v[e] = __float2half(a_value((int)rank, r, c8 * 8 + e));
        *reinterpret_cast<uint4 *>(sA + r * 128 + ((c8 ^ (r & 7)) << 4)) = *reinterpret_cast<uint4 *>(v);
    }
    for (int i = tid; i < (N / 2) * 8; i += 128) {   // B: rows n = rank * N/2 + nl of the N x 64 "identity" (n < 64: e_n, else 0)
        const int nl = i >> 3, c8 = i & 7, n = (int)rank * (N / 2) + nl;
        __half v[8];
        for (int e = 0; e < 8; ++e) v[e] = __float2half((c8 * 8 + e) == n ? 1.f : 0.f);
        *reinterpret_cast<uint4 *>(sB + nl * 128 + ((c8 ^ (nl & 7)) << 4)) = *reinterpret_cast<uint4 *>(v);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = slot;
    long long t0 = 0, t1 = 0;
    if (rank == 0 && tid == 32) {
        constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint64_t da = desc_sw128(smem_u32(sA) + k * 32), db = desc_sw128(smem_u32(sB) + k * 32);
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tbase), "l"(da),
                             "l"(db), "r"(idesc), "r"((uint32_t)(k != 0))
                             : "memory");
            }
        }
        t1 = clock64();
        const uint16_t mask = 3;
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "h"(mask)
                     : "memory");
    }
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(smem_u32(&bar))
                 : "memory");
    const long long t2 = clock64();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t u[32];
    int bad = 0;
    for (int c0 = 0; c0 < 64 && c0 < N; c0 += 32) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, "
            "%22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]),
              "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]),
              "=r"(u[23]), "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
            : "r"(tbase + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int c = 0; c < 32; ++c) bad += (__uint_as_float(u[c]) != a_value((int)rank, tid, c0 + c));
    }
    if (bad) atomicAdd(wrong + rank, bad);
    if (blockIdx.x == 0 && tid == 32) {
        out[0] = t1 - t0;
        out[1] = t2 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 256;" ::"r"(tbase) : "memory");
}

template <int N>
void run(long long *out, int *wrong) {
    cudaFuncSetAttribute(k_umma2<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
    cudaMemset(wrong, 0, 8);
    k_umma2<N><<<148, 128, 40 * 1024>>>(out, wrong, 1);
    cudaError_t e = cudaDeviceSynchronize();
    int w[2] = {-1, -1};
    cudaMemcpy(w, wrong, 8, cudaMemcpyDeviceToHost);
    printf("M=256 N=%3d cta_group::2 : single pass wrong elements: leader %d, peer %d (of 74 clusters) %s\n", N, w[0], w[1], e == cudaSuccess ? "" : cudaGetErrorString(e));
    const int iters = 256;
    k_umma2<N><<<148, 128, 40 * 1024>>>(out, wrong, iters);
    e = cudaDeviceSynchronize();
    long long h[2];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("           issue %.1f clk/mma, complete %.1f clk/mma (pair floor 256*N/512 = %d; per-SM equivalent of two M128 MMAs: %d) %s\n", h[0] / (iters * 4.0),
           h[1] / (iters * 4.0), 256 * N / 512, 2 * 128 * N / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    long long *out;
    int *wrong;
    cudaMalloc(&out, 64);
    cudaMalloc(&wrong, 8);
    run<64>(out, wrong);
    run<128>(out, wrong);
    run<256>(out, wrong);
    return 0;
}
