// TMEM load / store throughput per SM on B200 (sm_100a): W warps (4, 8, 16; warp w owns lane quadrant w % 4) issue back-to-back
// tcgen05.ld.32x32b.x32 (4 KB per warp-instruction) or tcgen05.st.32x32b.x16 (2 KB). Answers: is the attention kernel's per-tile
// read of the fp32 scores (128 x 64 x 4 B = 32 KB per 128 x 64 tile) a throughput limit next to the MUFU pipe?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench_tmem tools/ubench_tmem.cu && build/ubench_tmem
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>   // 0: ld x32 + wait each; 1: two ld x32 in flight per wait; 2: st x16 + wait; 3: ld x32 then st x16 (softmax-like)
__global__ void __launch_bounds__(512, 1) k_tmem(long long *out, float *sink, int iters) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = slot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64);
    uint32_t u[32], v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) u[i] = v[i] = threadIdx.x + i;
    float acc = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 1 || MODE == 3) {
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
                "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]), "=r"(u[10]),
                  "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]),
                  "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]),
                  "=r"(u[31])
                : "r"(taddr));
        }
        if (MODE == 1) {
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
                "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                  "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                  "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
                  "=r"(v[31])
                : "r"(taddr + 32));
        }
        if (MODE != 2) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (MODE == 2 || MODE == 3) {
            asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
                         "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]), "r"(u[8]), "r"(u[9]), "r"(u[10]), "r"(u[11]),
                         "r"(u[12]), "r"(u[13]), "r"(u[14]), "r"(u[15])
                         : "memory");
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        acc += __uint_as_float(u[it & 31]) + __uint_as_float(v[(it + 7) & 31]);
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    if (acc == 123.456f) sink[0] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

template <int MODE>
void run(const char *name, long long *out, float *sink, double bytes_per_iter_per_warp) {
    printf("%-34s", name);
    for (int warps : {4, 8, 16}) {
        const int iters = 4096;
        k_tmem<MODE><<<148, warps * 32>>>(out, sink, 16);
        cudaDeviceSynchronize();
        k_tmem<MODE><<<148, warps * 32>>>(out, sink, iters);
        cudaDeviceSynchronize();
        long long h;
        cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
        printf("  %2d warps: %6.1f clk/iter/warp-slot %6.1f B/clk/SM", warps, (double)h / iters, bytes_per_iter_per_warp * warps * iters / (double)h);
    }
    cudaError_t e = cudaGetLastError();
    printf("%s\n", e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    long long *out;
    float *sink;
    cudaMalloc(&out, 64);
    cudaMalloc(&sink, 64);
    run<0>("ld x32 + wait", out, sink, 4096);
    run<1>("2 x ld x32 + wait", out, sink, 8192);
    run<2>("st x16 + wait", out, sink, 2048);
    run<3>("ld x32, wait, st x16, wait", out, sink, 6144);
    return 0;
}
