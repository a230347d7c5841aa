// Which shared-memory layouts can a tcgen05.mma A descriptor (K-major, SWIZZLE_128B) address? Experiment behind the halo-tile
// convolution (conv_halo.cu): the A operand of tap (ky, kx) is a SHIFTED VIEW of one activation halo tile — start address moved
// by whole 128-byte pixel rows (not 1024-byte aligned) and 8-row groups `pitch` pixels apart (SBO = pitch * 128 B).
//   halo tile: pixel (hy, hx) at byte (hy * pitch + hx) * 128, its eight 16-byte channel chunks XOR-swizzled with address bits
//   [7, 10) — what a TMA SWIZZLE_128B box load into a 1024-byte aligned buffer produces.
//   A row m = (py, px), px < 8, py < 16  ->  halo pixel (py + ky, px + kx).
// B = 64 x 64 identity, so D[m][n] must equal A[m][n]. Prints, per (pitch, ky, kx, base_offset policy), the number of wrong elements.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench_umma_layout tools/ubench_umma_layout.cu && build/ubench_umma_layout
#include <cstdint>
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t sbo_bytes, uint32_t base_offset) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_offset & 7) << 49;
    d |= (uint64_t)2 << 61;
    return d;
}

__device__ __forceinline__ float a_value(int hy, int hx, int c) { return (float)((hy * 31 + hx * 7 + c * 3) % 61 - 30); }

__global__ void __launch_bounds__(128, 1) k_layout(int pitch, int ky, int kx, int policy, int *wrong) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    uint8_t *sA = smem;                 // halo tile: 18 rows x pitch pixels x 128 B  (<= 36 KB)
    uint8_t *sB = smem + 40 * 1024;     // 64 x 64 identity, canonical K-major SW128 (8 KB)
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 18 * pitch * 8; i += 128) {   // one 16-byte chunk per iteration
        const int pix = i >> 3, c8 = i & 7, hy = pix / pitch, hx = pix % pitch;
        const uint32_t off = (uint32_t)pix * 128;
        __half v[8];
        for (int e = 0; e < 8; ++e) v[e] = __float2half(a_value(hy, hx, c8 * 8 + e));
        *reinterpret_cast<uint4 *>(sA + off + ((c8 ^ ((off >> 7) & 7)) << 4)) = *reinterpret_cast<uint4 *>(v);
    }
    for (int i = tid; i < 64 * 8; i += 128) {
        const int n = i >> 3, c8 = i & 7;
        __half v[8];
        for (int e = 0; e < 8; ++e) v[e] = __float2half((c8 * 8 + e) == n ? 1.f : 0.f);
        *reinterpret_cast<uint4 *>(sB + n * 128 + ((c8 ^ (n & 7)) << 4)) = *reinterpret_cast<uint4 *>(v);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = slot;
    if (tid == 0) {
        constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t startA = smem_u32(sA) + (uint32_t)(ky * pitch + kx) * 128;
        const uint32_t bo = policy == 0 ? 0u : ((startA >> 7) & 7u);
        for (int k = 0; k < 4; ++k) {
            const uint64_t da = desc_sw128(startA + k * 32, (uint32_t)pitch * 128, bo);
            const uint64_t db = desc_sw128(smem_u32(sB) + k * 32, 1024, 0);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tbase), "l"(da),
                         "l"(db), "r"(idesc), "r"((uint32_t)(k != 0))
                         : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(smem_u32(&bar))
                 : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t u[32];
    int bad = 0;
    const int m = tid, py = m >> 3, px = m & 7;
    for (int c0 = 0; c0 < 64; c0 += 32) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, "
            "%22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]),
              "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]),
              "=r"(u[23]), "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
            : "r"(tbase + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int c = 0; c < 32; ++c) bad += (__uint_as_float(u[c]) != a_value(py + ky, px + kx, c0 + c));
    }
    if (bad) atomicAdd(wrong, bad);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tbase) : "memory");
}

int main() {
    int *wrong;
    cudaMalloc(&wrong, 4);
    cudaFuncSetAttribute(k_layout, cudaFuncAttributeMaxDynamicSharedMemorySize, 50 * 1024);
    for (int pitch : {8, 16, 10, 12}) {
        for (int policy = 0; policy < 2; ++policy) {
            printf("pitch %2d px (SBO %4d B), base_offset %s :", pitch, pitch * 128, policy ? "= (start>>7)&7" : "= 0           ");
            for (int ky = 0; ky < 3; ++ky)
                for (int kx = 0; kx < 3; ++kx) {
                    if (pitch == 8 && kx) continue;   // no room for a horizontal shift in an 8-pixel pitch
                    cudaMemset(wrong, 0, 4);
                    k_layout<<<1, 128, 50 * 1024>>>(pitch, ky, kx, policy, wrong);
                    int h = -1;
                    cudaMemcpy(&h, wrong, 4, cudaMemcpyDeviceToHost);
                    printf(" (%d,%d):%d", ky, kx, h);
                }
            cudaError_t e = cudaDeviceSynchronize();
            printf("%s\n", e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    }
    return 0;
}
