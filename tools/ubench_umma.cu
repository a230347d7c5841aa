// tcgen05.mma dispatch micro-benchmark (B200, sm_100a): cycles per M128 x N x K16 kind::f16 instruction as a function of N, of the
// number of INDEPENDENT accumulators the issue order rotates over, and of where A comes from (shared memory / tensor memory).
// Question it answers for the attention kernel: are the 3 + 4 small MMAs of a key step (N = 64 / 48, each a dependent accumulation
// into the same TMEM columns) paced by the tensor-pipe floor 128 N / 256 cycles, or by a per-instruction / dependent-chain latency?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench_umma tools/ubench_umma.cu && build/ubench_umma
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a, uint32_t sbo = 1024) {
    uint64_t d = 0;
    d |= (uint64_t)((a & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b),
                 "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b),
                 "r"(idesc), "r"(acc)
                 : "memory");
}

// ORDER 0: a-major (finish the 4-step chain of accumulator a, then the next a) ; 1: k-major (rotate over accumulators every instruction)
template <int N, int NACC, bool TS, int ORDER>
__global__ void __launch_bounds__(128, 1) k_umma(long long *out, int iters, int a_off, int a_sbo) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 56 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = slot;
    if (threadIdx.x == 32) {
        constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t sa = smem_u32(smem) + (uint32_t)a_off, sb = smem_u32(smem) + 24576;
        const uint32_t tA = tbase + 480;
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (ORDER == 0) {
#pragma unroll
                for (int a = 0; a < NACC; ++a)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (TS) mma_ts(tbase + a * N, tA + k * 8, desc_sw128(sb + k * 32), idesc, (it | k) != 0);
                        else mma_ss(tbase + a * N, desc_sw128(sa + k * 32, (uint32_t)a_sbo), desc_sw128(sb + k * 32), idesc, (it | k) != 0);
                    }
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int a = 0; a < NACC; ++a) {
                        if (TS) mma_ts(tbase + a * N, tA + k * 8, desc_sw128(sb + k * 32), idesc, (it | k) != 0);
                        else mma_ss(tbase + a * N, desc_sw128(sa + k * 32, (uint32_t)a_sbo), desc_sw128(sb + k * 32), idesc, (it | k) != 0);
                    }
            }
        }
        long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile(
            "{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(smem_u32(&bar))
            : "memory");
        long long t2 = clock64();
        if (blockIdx.x == 0) {
            out[0] = t1 - t0;
            out[1] = t2 - t0;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
}

template <int N, int NACC, bool TS, int ORDER>
void run(long long *out, int a_off = 0, int a_sbo = 1024) {
    const int iters = 256;
    cudaFuncSetAttribute(k_umma<N, NACC, TS, ORDER>, cudaFuncAttributeMaxDynamicSharedMemorySize, 60 * 1024);
    k_umma<N, NACC, TS, ORDER><<<148, 128, 60 * 1024>>>(out, 8, a_off, a_sbo);
    cudaDeviceSynchronize();
    k_umma<N, NACC, TS, ORDER><<<148, 128, 60 * 1024>>>(out, iters, a_off, a_sbo);
    cudaDeviceSynchronize();
    long long h[2];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    const double n = (double)iters * NACC * 4;
    cudaError_t e = cudaGetLastError();
    if (a_off || a_sbo != 1024) printf("[A start +%d B, SBO %d B] ", a_off, a_sbo);
    printf("N=%3d accumulators=%d A=%s order=%s : issue %.1f clk/mma, complete %.1f clk/mma (floor %d) %s\n", N, NACC, TS ? "tmem" : "smem",
           ORDER ? "rotate" : "chain ", h[0] / n, h[1] / n, 128 * N / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    long long *out;
    cudaMalloc(&out, 64);
    run<48, 1, true, 0>(out);
    run<48, 2, true, 0>(out);
    run<48, 2, true, 1>(out);
    run<48, 4, true, 1>(out);
    run<48, 8, true, 1>(out);
    run<64, 1, true, 0>(out);
    run<64, 2, true, 1>(out);
    run<64, 4, true, 1>(out);
    run<128, 1, true, 0>(out);
    run<128, 2, true, 1>(out);
    run<256, 1, true, 0>(out);
    run<48, 1, false, 0>(out);
    run<48, 4, false, 1>(out);
    run<64, 1, false, 0>(out);
    run<64, 4, false, 1>(out);
    run<128, 1, false, 0>(out);
    run<128, 2, false, 1>(out);
    run<256, 1, false, 0>(out);
    // shifted / re-pitched A views (halo-tile convolution): does a start that is not 1024-byte aligned, or 8-row groups 1280 B apart, cost reads?
    run<128, 1, false, 0>(out, 128, 1024);
    run<128, 1, false, 0>(out, 0, 1280);
    run<128, 1, false, 0>(out, 128 * 11, 1280);
    run<128, 1, false, 0>(out, 128 * 22, 1280);
    run<256, 1, false, 0>(out, 128 * 11, 1280);
    run<64, 1, false, 0>(out, 128 * 11, 1280);
    return 0;
}
