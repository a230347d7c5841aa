// C2 — 3x3 convolution (stride 1, pad 1, NHWC fp16) on tcgen05 from ACTIVATION HALO TILES, with the GroupNorm affine + SiLU of the
// input applied in shared memory: the `norm -> SiLU -> conv` of every diffusers ResnetBlock2D / conv_norm_out reached from
// utils/adaptive_mask_inpainting.py:1001 (UNet) and :680 / :1086 / :1112 (VAE), in ONE kernel.
//
// The implicit-GEMM kernel (gemm.cu, CONV) fetches the A operand of every (tap, 64-channel block) K-slab with its own TMA load — each
// input pixel crosses L2 -> SMEM nine times — and needs the normalised + activated tensor materialised in HBM by a separate pass
// (affine_act_kernel: 12-20 % of a VAE decode). Here:
//   * an output tile is 16 rows x 8 columns of pixels; per 64-channel block its 18 x 10 pixel halo is staged ONCE in shared memory
//     (23 KB instead of nine shifted 16 KB tiles): twelve builder warps read it from global memory with coalesced 16-byte loads,
//     apply act(x * scale[b, c] + shift[b, c]) in registers (fp16, the same roundings as the tensor the unfused path stores; pixels
//     outside the image are the convolution's zero padding) and store it in the 128B-swizzled K-major layout, while the tensor core
//     works on the previous blocks (4-deep ring);
//   * the A operand of tap (ky, kx) is a SHIFTED VIEW of that halo tile: start address + (ky * 10 + kx) * 128 B, 8-row groups
//     1280 B apart (SBO). The tensor core applies the 128B swizzle to absolute shared-memory address bits, so any 128-byte
//     aligned start and any SBO address the tile correctly (tools/ubench_umma_layout.cu: every shift, pitch 10 / 12 / 16: exact);
//   * weights stream through their own ring ([BN x 64] slabs, K order (ky, kx, cin) as in prep_conv3x3);
//   * epilogue as in gemm.cu: TMEM -> registers -> bias / per-sample bias row (time embedding) / residual / SiLU -> 64B-swizzled
//     32 x 32 panels -> TMA store (box 32 channels x 8 px x 4 rows), GroupNorm partial sums of the rounded output for the next norm.
// Warps: 0 TMA producer (weights), 1 MMA issuer (both converged, one elected lane), 2-5 epilogue, 6-17 halo builders.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace coma {
namespace ch {
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity) {   // acquires what peer-CTA threads released with their remote arrive
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, uint64_t *bar, int x, int y, int z, int w) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z), "r"(w)
        : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap *map, const void *src, int x, int y, int z, int w) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map), "r"(smem_u32(src)),
                 "r"(x), "r"(y), "r"(z), "r"(w)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// K-major operand, 128-byte rows, 128B swizzle; sbo_bytes = distance between 8-row groups
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t *v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, "
        "%25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// explicit shared-space accesses: the dynamic window's base is rounded up through an integer cast, so plain pointer dereferences
// compile to GENERIC LD / ST (slower path, long-scoreboard latency) instead of LDS / STS
__device__ __forceinline__ void sts128(uint32_t addr, const uint4 &v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sts128f(uint32_t addr, const float4 &v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ float exp2f_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float silu_f(float v) { return __fdividef(v, 1.0f + __expf(-v)); }   // as affine_act_kernel (unet_ops.cu)
// SiLU for the halo builders: ONE MUFU op per element (ex2) instead of two — the builders are MUFU-bound at BN = 128 (23040 ops per
// 64-channel block against 2304 clk of MMAs). The reciprocal of 1 + e^-v runs on the FMA pipes: magic-constant seed (12 % off) + three
// Newton steps (relative error < 4e-6 over [-90, 30]: 4e-5 of the fp16 results differ in the last bit from the exact value); v is clamped
// at -80 so that 1 + e^-v stays finite.
__device__ __forceinline__ float2 silu_fma2(float2 v) {   // two elements at a time in packed FP32x2 instructions (half the issue slots)
    const float2 a = __fmul2_rn(make_float2(fmaxf(v.x, -80.0f), fmaxf(v.y, -80.0f)), make_float2(-1.4426950408889634f, -1.4426950408889634f));
    const float2 y = __fadd2_rn(make_float2(exp2f_approx(a.x), exp2f_approx(a.y)), make_float2(1.0f, 1.0f));
    const float2 ny = make_float2(-y.x, -y.y), two = make_float2(2.0f, 2.0f);
    float2 r = make_float2(__int_as_float(0x7EF311C7 - __float_as_int(y.x)), __int_as_float(0x7EF311C7 - __float_as_int(y.y)));
    r = __fmul2_rn(r, __ffma2_rn(ny, r, two));
    r = __fmul2_rn(r, __ffma2_rn(ny, r, two));
    r = __fmul2_rn(r, __ffma2_rn(ny, r, two));
    return __fmul2_rn(v, r);
}

// ---- CTA pair (cta_group::2): the leader CTA (cluster rank 0) issues M256 x N x K16 instructions over BOTH CTAs' shared memory
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_rank(uint32_t local, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion is signalled on the barrier at the same offset in the LEADER CTA of the pair (peer bit of the address cleared)
__device__ __forceinline__ void tma_load_4d_pair(void *dst, const CUtensorMap *map, uint64_t *bar, int x, int y, int z, int w) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(x), "r"(y), "r"(z), "r"(w)
        : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t *bar) {   // arrives on the barrier at this offset in BOTH CTAs
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}
}  // namespace ch

constexpr int CH_TW = 8, CH_TH = 16;                       // output tile: 16 rows x 8 columns = 128 pixels (TMEM lane = py * 8 + px)
constexpr int CH_HW = CH_TW + 2, CH_HH = CH_TH + 2;        // halo: 18 rows x 10 columns
constexpr int CH_HPIX = CH_HW * CH_HH;                     // 180 pixels x 128 B per 64-channel block
constexpr int CH_HALO_BYTES = CH_HPIX * 128;               // 23040
constexpr int CH_HALO_STRIDE = (CH_HALO_BYTES + 1023) / 1024 * 1024;
constexpr int CH_NH = 4;                                   // halo ring: one block under the MMAs, one or two in the transform (a pair waits for the slower CTA), one in flight
constexpr int CH_PANEL_BYTES = 32 * 32 * 2;
__host__ __device__ constexpr int ch_epi_warps(int BN) { return 4; }   // K >= 576: the epilogue (TMEM read rate bound) hides behind the next tile's main loop even at BN = 256
constexpr int CH_TWARPS = 12;                              // halo builder warps: three per scheduler hide each other's load / MUFU latencies
__host__ __device__ constexpr int ch_threads(int BN) { return 64 + 32 * ch_epi_warps(BN) + 32 * CH_TWARPS; }
__host__ __device__ constexpr int ch_b_stages(int BN, bool pair) { return pair ? (BN <= 160 ? 8 : 6) : (BN <= 64 ? 8 : (BN <= 128 ? 6 : (BN <= 160 ? 5 : 3))); }   // 64-128 KB of weight slabs in flight (TMA latency ~2000 clk)
__host__ __device__ constexpr int ch_b_rows(int BN, bool pair) { return pair ? BN / 2 : BN; }   // weight rows per CTA and slab: a pair splits the tile's N
__host__ __device__ constexpr size_t ch_smem(int BN, bool pair) {
    return (size_t)CH_NH * CH_HALO_STRIDE + (size_t)ch_b_stages(BN, pair) * ch_b_rows(BN, pair) * 128 + (size_t)ch_epi_warps(BN) * 2 * CH_PANEL_BYTES + 2048 + 256 + 1024;
}

struct HaloArgs {
    const void *x;                                    // [B, H, W, C] f16, pixel stride ldx; up: stored as [B, H/2, W/2, C], read as its nearest x2 upsampling
    long long ldx;
    int up;
    int B, H, W, C, N;
    int cblocks, tiles_x, per_img, n_tiles, total;   // total = m_tiles * n_tiles work items; id = mt * n_tiles + nt (N fastest)
                                                     // PAIR: total = (m_tiles / 2) * n_tiles pair items; CTA `rank` of the pair owns pixel tile 2 * mp + rank
    const float *scale, *shift;                       // [B, C] or null: the input is used as stored
    int act_in;                                       // 1: SiLU after the affine
    const float *bias, *bias_rows;                    // [N]; [B, bias_rows_ld] per-sample rows (time embedding) or null
    long long bias_rows_ld;
    const __half *residual;                           // [B*H*W, ldo] or null
    long long ldo;
    int act_out;
    float *stats;                                     // [B*H*W/128, N, 2] (one row per 16 x 8 pixel tile) or null
};

// PAIR: two CTAs of a cluster (one TPC) work on two pixel tiles of the same output-channel tile with ONE tcgen05.mma.cta_group::2 stream
// (M = 256) issued by the leader: each CTA stages its own halo (A rows) but only HALF of every weight slab, and the tensor core reads
// A 4 KB + B 2 KB per CTA and K step instead of 4 + 4 KB (BN = 128) — the single-CTA kernel runs 1.5-2.2x above the MMA floor on
// shared-memory bandwidth. Cross-CTA protocol: weight slabs complete on the LEADER's b_full (cta_group::2 TMA,
// expect_tx for both halves posted by the leader); halo_ready / tmem_empty of the leader count elected arrivals of both CTAs (remote
// mbarrier.arrive); b_empty / halo_empty / tmem_full are released in both CTAs by multicast tcgen05.commit.
template <int BN, bool PAIR>
__global__ void __launch_bounds__(ch_threads(BN), 1)
    conv_halo_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmO,
                     const HaloArgs a) {
    using namespace ch;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr int EPI = ch_epi_warps(BN), PW = EPI / 4, NB = ch_b_stages(BN, PAIR), B_ROWS = ch_b_rows(BN, PAIR), B_BYTES = B_ROWS * 128;
    const uint32_t rank = PAIR ? cluster_rank() : 0u;
    const int item0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, item_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    uint8_t *sH = smem;
    uint8_t *sB = sH + CH_NH * CH_HALO_STRIDE;
    uint8_t *sE = sB + NB * B_BYTES;
    float4 *sRed = reinterpret_cast<float4 *>(sE + EPI * 2 * CH_PANEL_BYTES);   // [2][4 warps][16 lanes]: GroupNorm partial sums of a panel, per quarter
    uint64_t *bar = reinterpret_cast<uint64_t *>(sE + EPI * 2 * CH_PANEL_BYTES + 2048);
    uint64_t *halo_ready = bar, *halo_empty = halo_ready + CH_NH;
    uint64_t *b_full = halo_empty + CH_NH, *b_empty = b_full + NB;
    uint64_t *tmem_full = b_empty + NB, *tmem_empty = tmem_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr uint32_t TMEM_COLS = (2 * BN <= 128) ? 128 : ((2 * BN <= 256) ? 256 : 512);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmO) : "memory");
        for (int i = 0; i < CH_NH; ++i) {
            mbar_init(halo_ready + i, PAIR ? 2 * CH_TWARPS : 32 * CH_TWARPS);   // PAIR: one elected arrival per transform warp of both CTAs
            mbar_init(halo_empty + i, 1);
        }
        for (int i = 0; i < NB; ++i) {
            mbar_init(b_full + i, 1);
            mbar_init(b_empty + i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(tmem_full + i, 1);
            mbar_init(tmem_empty + i, PAIR ? 2 * EPI : 32 * EPI);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (PAIR) cluster_sync_all();   // the peer's barriers must exist before anything signals them
    else __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    // work item -> output-channel offset, image, tile origin (PAIR: of THIS CTA's pixel tile of the pair item)
    auto coords = [&](int t, int &n0, int &img, int &x0, int &y0) {
        const int nt = t % a.n_tiles, mt = PAIR ? 2 * (t / a.n_tiles) + (int)rank : t / a.n_tiles;
        n0 = nt * BN;
        img = mt / a.per_img;
        const int rem = mt - img * a.per_img;
        const int ty = rem / a.tiles_x;
        y0 = ty * CH_TH;
        x0 = (rem - ty * a.tiles_x) * CH_TW;
    };

    if (warp == 0) {
        // ---- producer: weight slabs only (the halo tiles are built by the last eight warps)
        int bs = 0;
        uint32_t bph = 0;
        for (int t = item0; t < a.total; t += item_step) {
            int n0, img, x0, y0;
            coords(t, n0, img, x0, y0);
            for (int cb = 0; cb < a.cblocks; ++cb) {
                for (int tap = 0; tap < 9; ++tap) {
                    mbar_wait(b_empty + bs, bph ^ 1);
                    if (elect_one()) {
                        if (PAIR) {   // this CTA's half of the slab's rows; both halves complete on the leader's barrier
                            if (rank == 0) mbar_expect_tx(b_full + bs, 2 * B_BYTES);
                            tma_load_4d_pair(sB + bs * B_BYTES, &tmW, b_full + bs, (tap * a.cblocks + cb) * 64, n0 + (int)rank * B_ROWS, 0, 0);
                        } else {
                            mbar_expect_tx(b_full + bs, B_BYTES);
                            tma_load_4d(sB + bs * B_BYTES, &tmW, b_full + bs, (tap * a.cblocks + cb) * 64, n0, 0, 0);
                        }
                    }
                    __syncwarp();
                    if (++bs == NB) { bs = 0; bph ^= 1; }
                }
            }
        }
        if (lane == 0) pdl_trigger();
    } else if (warp == 1) {
        // ---- MMA issuer (PAIR: the leader CTA only; the peer's warp 1 has nothing to do)
        constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((PAIR ? 256 : 128) >> 4) << 24);
        const uint64_t hdesc0 = umma_desc_sw128(smem_u32(sH), CH_HW * 128), bdesc0 = umma_desc_sw128(smem_u32(sB), 1024);
        int hs = 0, bs = 0, i = 0;
        uint32_t hph = 0, bph = 0;
        for (int t = item0; t < a.total && rank == 0; t += item_step, ++i) {
            const int acc = i & 1;
            if (PAIR) mbar_wait_cluster(tmem_empty + acc, ((i >> 1) & 1) ^ 1);
            else mbar_wait(tmem_empty + acc, ((i >> 1) & 1) ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
            for (int cb = 0; cb < a.cblocks; ++cb) {
                if (PAIR) mbar_wait_cluster(halo_ready + hs, hph);
                else mbar_wait(halo_ready + hs, hph);
                const uint64_t hd = hdesc0 + (uint64_t)(hs * (CH_HALO_STRIDE >> 4));
                for (int tap = 0; tap < 9; ++tap) {
                    mbar_wait(b_full + bs, bph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (elect_one()) {
                        const int ky = (tap * 11) >> 5, kx = tap - 3 * ky;
                        const uint64_t da = hd + (uint64_t)((ky * CH_HW + kx) * 8);      // (ky * 10 + kx) pixel rows of 128 B, in 16-byte units
                        const uint64_t db = bdesc0 + (uint64_t)(bs * (B_BYTES >> 4));
                        if (PAIR) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) umma_f16_pair(tmem_d, da + 2 * k, db + 2 * k, idesc, (cb | tap | k) != 0);
                            umma_commit_pair(b_empty + bs);
                            if (tap == 8) {
                                umma_commit_pair(halo_empty + hs);
                                if (cb == a.cblocks - 1) umma_commit_pair(tmem_full + acc);
                            }
                        } else {
#pragma unroll
                            for (int k = 0; k < 4; ++k) umma_f16(tmem_d, da + 2 * k, db + 2 * k, idesc, (cb | tap | k) != 0);
                            umma_commit(b_empty + bs);
                            if (tap == 8) {
                                umma_commit(halo_empty + hs);
                                if (cb == a.cblocks - 1) umma_commit(tmem_full + acc);
                            }
                        }
                    }
                    __syncwarp();
                    if (++bs == NB) { bs = 0; bph ^= 1; }
                }
                if (++hs == CH_NH) { hs = 0; hph ^= 1; }
            }
        }
    } else if (warp >= 2 + EPI) {
        // ---- halo builders: global -> registers -> act(x * scale + shift) -> swizzled shared memory. Thread (prow, c8) owns the 16-byte
        // channel chunk c8 of halo pixels prow, prow + 32, ... (a warp reads 4 pixels x 128 B: coalesced). All of a block's loads are
        // issued before the first is consumed. (First version: TMA load + in-place LDS / STS transform — ncu showed the builders stalled
        // on those LDS behind the tensor core's own operand reads, and the MMA stream waiting for them: tensor pipe 49.6 % active.)
        const int tt = threadIdx.x - (2 + EPI) * 32, c8 = tt & 7, prow = tt >> 3;
        constexpr int NPX = (CH_HPIX + 4 * CH_TWARPS - 1) / (4 * CH_TWARPS);   // pixels per thread and block (6)
        int hs = 0;
        uint32_t hph = 0;
        const uint32_t ready0 = PAIR ? map_to_rank(smem_u32(halo_ready), 0) : 0u;   // the leader's halo_ready[0]
        const __half *xg = reinterpret_cast<const __half *>(a.x);
        for (int t = item0; t < a.total; t += item_step) {
            int n0, img, x0, y0;
            coords(t, n0, img, x0, y0);
            for (int cb = 0; cb < a.cblocks; ++cb) {
                uint4 raw[NPX];
                bool ok[NPX];
#pragma unroll
                for (int i = 0; i < NPX; ++i) {
                    const int pix = prow + i * 4 * CH_TWARPS;
                    const int hy = pix / CH_HW, hx = pix - hy * CH_HW;
                    const int gy = y0 - 1 + hy, gx = x0 - 1 + hx;
                    ok[i] = pix < CH_HPIX && (unsigned)gy < (unsigned)a.H && (unsigned)gx < (unsigned)a.W;
                    raw[i] = make_uint4(0u, 0u, 0u, 0u);   // outside the image: the convolution's zero padding (of the ACTIVATED tensor)
                    if (ok[i]) {
                        const size_t src = a.up ? ((size_t)img * (a.H >> 1) + (gy >> 1)) * (a.W >> 1) + (gx >> 1) : ((size_t)img * a.H + gy) * a.W + gx;
                        raw[i] = __ldg(reinterpret_cast<const uint4 *>(xg + src * a.ldx + cb * 64 + c8 * 8));
                    }
                }
                float2 sc[4], sh[4];
                if (a.scale) {
                    const size_t co = (size_t)img * a.C + cb * 64 + c8 * 8;
                    const float4 s0 = __ldg(reinterpret_cast<const float4 *>(a.scale + co)), s1 = __ldg(reinterpret_cast<const float4 *>(a.scale + co + 4));
                    const float4 t0 = __ldg(reinterpret_cast<const float4 *>(a.shift + co)), t1 = __ldg(reinterpret_cast<const float4 *>(a.shift + co + 4));
                    sc[0] = make_float2(s0.x, s0.y); sc[1] = make_float2(s0.z, s0.w); sc[2] = make_float2(s1.x, s1.y); sc[3] = make_float2(s1.z, s1.w);
                    sh[0] = make_float2(t0.x, t0.y); sh[1] = make_float2(t0.z, t0.w); sh[2] = make_float2(t1.x, t1.y); sh[3] = make_float2(t1.z, t1.w);
                }
                mbar_wait(halo_empty + hs, hph ^ 1);   // the MMAs that read this buffer NH blocks ago have retired
                const uint32_t buf = smem_u32(sH) + (uint32_t)(hs * CH_HALO_STRIDE);
#pragma unroll
                for (int i = 0; i < NPX; ++i) {
                    const int pix = prow + i * 4 * CH_TWARPS;
                    if (pix >= CH_HPIX) continue;
                    if (a.scale && ok[i]) {
                        __half2 *h = reinterpret_cast<__half2 *>(&raw[i]);
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            float2 v = __ffma2_rn(__half22float2(h[u]), sc[u], sh[u]);
                            if (a.act_in == 1) v = silu_fma2(v);
                            h[u] = __floats2half2_rn(v.x, v.y);
                        }
                    }
                    sts128(buf + (uint32_t)(pix * 128 + ((c8 ^ (pix & 7)) << 4)), raw[i]);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
                if (PAIR) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(ready0 + (uint32_t)hs * 8);
                } else {
                    mbar_arrive(halo_ready + hs);
                }
                if (++hs == CH_NH) { hs = 0; hph ^= 1; }
            }
        }
    } else {
        // ---- epilogue: warp owns TMEM lanes [32q, 32q + 32) = tile rows 4q .. 4q+3 (8 pixels each); panels of 32 output channels
        const int q = warp & 3, half = (warp - 2) >> 2;
        uint8_t *ebuf = sE + (warp - 2) * (2 * CH_PANEL_BYTES);
        const uint32_t my_row = smem_u32(ebuf) + (uint32_t)lane * 64;
        const int sw = (lane >> 1) & 3;   // 64B swizzle: 16-byte chunk index ^= (row / 2) % 4
        const int py = q * 4 + (lane >> 3), px = lane & 7;
        uint32_t g = 0;
        int i = 0;
        const uint32_t empty0 = PAIR ? map_to_rank(smem_u32(tmem_empty), 0) : 0u;   // the leader's tmem_empty[0]
        for (int t = item0; t < a.total; t += item_step, ++i) {
            int n0, img, x0, y0;
            coords(t, n0, img, x0, y0);
            const int P = (min(BN, a.N - n0) + 31) >> 5;
            const size_t pixel = ((size_t)img * a.H + (y0 + py)) * a.W + (x0 + px);
            const bool pix_ok = (y0 + py) < a.H && (x0 + px) < a.W;
            const float *brow = a.bias_rows ? a.bias_rows + (size_t)img * a.bias_rows_ld : nullptr;
            const int acc = i & 1;
            mbar_wait(tmem_full + acc, (i >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN) + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
            for (int p = half; p < P; p += PW, ++g) {
                const uint32_t buf = g & 1;
                const uint32_t prow = my_row + buf * CH_PANEL_BYTES;
                uint32_t v[32];
                tmem_ld32(tmem_d + (uint32_t)(p * 32), v);
                const int nb = n0 + p * 32;
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 b4 = a.bias ? __ldg(reinterpret_cast<const float4 *>(a.bias + nb + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    f[j] = __uint_as_float(v[j]) + b4.x;
                    f[j + 1] = __uint_as_float(v[j + 1]) + b4.y;
                    f[j + 2] = __uint_as_float(v[j + 2]) + b4.z;
                    f[j + 3] = __uint_as_float(v[j + 3]) + b4.w;
                }
                if (brow) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4 *>(brow + nb + j));
                        f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
                    }
                }
                if (a.residual && pix_ok) {
                    const uint4 *rp = reinterpret_cast<const uint4 *>(a.residual + pixel * a.ldo + nb);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const uint4 r = __ldg(rp + c);
                        const __half2 *h = reinterpret_cast<const __half2 *>(&r);
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float2 x = __half22float2(h[u]);
                            f[c * 8 + 2 * u] += x.x;
                            f[c * 8 + 2 * u + 1] += x.y;
                        }
                    }
                }
                if (a.act_out == 1) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = silu_f(f[j]);
                }
                if (lane == 0) bulk_wait_read<1>();   // the store that last read this staging buffer has drained it
                __syncwarp();
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint4 w;
                    __half2 *h = reinterpret_cast<__half2 *>(&w);
#pragma unroll
                    for (int u = 0; u < 4; ++u) h[u] = __floats2half2_rn(f[c * 8 + 2 * u], f[c * 8 + 2 * u + 1]);
                    sts128(prow + (uint32_t)((c ^ sw) << 4), w);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    tma_store_4d(&tmO, ebuf + buf * CH_PANEL_BYTES, nb, x0, y0 + q * 4, img);   // clips rows / columns outside the tensor
                    bulk_commit();
                }
                if (a.stats) {
                    // GroupNorm partial sums of the ROUNDED fp16 panel (see gemm.cu): lane (cp, par) adds the 16 rows of parity par of
                    // column pair cp, parities combined in a fixed order; block id = (pixel tile, quarter) -> 32 pixels of one image
                    const int cp = lane & 15, par = lane >> 4;
                    const uint32_t pan = smem_u32(ebuf) + buf * CH_PANEL_BYTES;
                    float2 s2 = make_float2(0.f, 0.f), q2 = make_float2(0.f, 0.f);
#pragma unroll
                    for (int r16 = 0; r16 < 16; ++r16) {
                        const int r = 2 * r16 + par;
                        const uint32_t hb = lds32(pan + (uint32_t)(r * 64 + ((((cp >> 2) ^ ((r >> 1) & 3))) << 4) + (cp & 3) * 4));
                        const float2 x = __half22float2(*reinterpret_cast<const __half2 *>(&hb));
                        s2 = __fadd2_rn(s2, x);
                        q2 = __ffma2_rn(x, x, q2);
                    }
                    s2.x += __shfl_down_sync(0xffffffffu, s2.x, 16);
                    s2.y += __shfl_down_sync(0xffffffffu, s2.y, 16);
                    q2.x += __shfl_down_sync(0xffffffffu, q2.x, 16);
                    q2.y += __shfl_down_sync(0xffffffffu, q2.y, 16);
                    // one (sum, sumsq) row per 128-pixel TILE: the four quarters meet in shared memory (double-buffered by panel parity,
                    // one named barrier per panel) and warp 0 adds them in quarter order — 4x less for groupnorm_from_stats to read
                    const uint32_t red = smem_u32(sRed) + (uint32_t)((g & 1) * 64) * 16;
                    if (par == 0) sts128f(red + (uint32_t)(q * 16 + cp) * 16, make_float4(s2.x, q2.x, s2.y, q2.y));
                    asm volatile("bar.sync 2, 128;" ::: "memory");
                    const int col = nb + 2 * cp;
                    if (q == 0 && par == 0 && col < a.N) {
                        float4 r = lds128f(red + (uint32_t)cp * 16);
#pragma unroll
                        for (int w = 1; w < 4; ++w) {
                            const float4 o = lds128f(red + (uint32_t)(w * 16 + cp) * 16);
                            r.x += o.x; r.y += o.y; r.z += o.z; r.w += o.w;
                        }
                        const size_t blk = (size_t)(PAIR ? 2 * (t / a.n_tiles) + (int)rank : t / a.n_tiles);
                        *reinterpret_cast<float4 *>(a.stats + (blk * a.N + col) * 2) = r;
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            if (PAIR) {
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(empty0 + (uint32_t)acc * 8);
            } else {
                mbar_arrive(tmem_empty + acc);
            }
        }
        if (lane == 0) bulk_wait_read<0>();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (PAIR) {
        cluster_sync_all();   // neither CTA may exit (or free tensor memory) while the pair's last instructions / remote arrivals are in flight
        if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    } else {
        __syncthreads();
        if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

typedef CUresult (*EncodeTiledFnH)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                   const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFnH halo_encode_fn() {
    static EncodeTiledFnH fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFnH>(p);
    }
    return fn;
}
static int halo_map(CUtensorMap *m, const void *ptr, const cuuint64_t dims[4], const cuuint64_t strides[3], const cuuint32_t box[4],
                    CUtensorMapSwizzle swz, const char *what) {
    EncodeTiledFnH fn = halo_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return COMA_E_NODEVICE;
    }
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (%s) failed with CUresult %d", what, (int)r);
        return COMA_E_BADARG;
    }
    return 0;
}

template <int BN, bool PAIR>
static int launch_halo(const CUtensorMap &tw, const CUtensorMap &to, const HaloArgs &a, cudaStream_t st) {
    constexpr size_t smem = ch_smem(BN, PAIR);
    static bool attr[16] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 16 && !attr[dev]) {
        cudaError_t e = cudaFuncSetAttribute(conv_halo_kernel<BN, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(conv_halo): %s", cudaGetErrorString(e));
            return (int)e;
        }
        attr[dev] = true;
    }
    if (!PAIR) {
        const unsigned grid = (unsigned)(a.total < kNumSM ? a.total : kNumSM);
        launch_pdl(conv_halo_kernel<BN, PAIR>, dim3(grid), dim3(ch_threads(BN)), smem, st, tw, to, a);
    } else {
        // clusters of two CTAs (one TPC each); programmatic dependent launch as everywhere else
        const unsigned pairs = (unsigned)(a.total < kNumSM / 2 ? a.total : kNumSM / 2);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * pairs);
        cfg.blockDim = dim3(ch_threads(BN));
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        static const bool no_pdl = getenv("COMA_NO_PDL") != nullptr;
        cfg.numAttrs = no_pdl ? 1 : 2;
        cudaLaunchKernelEx(&cfg, conv_halo_kernel<BN, PAIR>, tw, to, a);
    }
    return check_launch("conv_halo_kernel");
}
}  // namespace coma

extern "C" int coma_conv3x3_halo_supported(int64_t B, int64_t H, int64_t W, int64_t C, int64_t N) {
    return (B > 0 && H % coma::CH_TH == 0 && W % coma::CH_TW == 0 && C % 64 == 0 && N % 64 == 0 && N >= 64 &&
            B * (H / coma::CH_TH) * (W / coma::CH_TW) >= 2 * coma::kNumSM) ? 1 : 0;
}

extern "C" int coma_conv3x3_halo_f16(const void *x, int64_t B, int64_t H, int64_t W, int64_t C, int64_t ldx, int up, const float *scale,
                                     const float *shift, int act_in, const void *Wt, int64_t ldw, int64_t N, const float *bias,
                                     const float *bias_rows, int64_t bias_rows_ld, const void *residual, int act_out, void *out_f16,
                                     int64_t ldo, float *stats, coma_stream_t stream) {
    using namespace coma;
    COMA_REQUIRE(x && Wt && out_f16, "null pointer");
    COMA_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && N > 0, "bad sizes");
    COMA_REQUIRE(H % CH_TH == 0 && W % CH_TW == 0, "halo conv: H must be a multiple of 16 and W of 8");
    COMA_REQUIRE(C % 64 == 0 && N % 64 == 0, "halo conv: C and N must be multiples of 64");
    COMA_REQUIRE(ldx % 8 == 0 && ldx >= C && ldw % 8 == 0 && ldw >= 9 * C && ldo % 8 == 0 && ldo >= N, "bad leading dimensions");
    COMA_REQUIRE(((uintptr_t)x | (uintptr_t)Wt | (uintptr_t)out_f16 | (uintptr_t)residual | (uintptr_t)bias | (uintptr_t)bias_rows |
                  (uintptr_t)scale | (uintptr_t)shift | (uintptr_t)stats) % 16 == 0, "pointers must be 16-byte aligned");
    COMA_REQUIRE((scale == nullptr) == (shift == nullptr), "scale and shift come together");
    COMA_REQUIRE(act_in >= 0 && act_in <= 1 && act_out >= 0 && act_out <= 1, "act must be 0 or 1");
    COMA_REQUIRE(!bias_rows || bias_rows_ld % 4 == 0, "bias_rows_ld must be a multiple of 4");
    const int bn = N % 256 == 0 ? 256 : (N % 160 == 0 ? 160 : (N % 128 == 0 ? 128 : 64));
    HaloArgs a;
    a.x = x; a.ldx = ldx; a.up = up ? 1 : 0;
    a.B = (int)B; a.H = (int)H; a.W = (int)W; a.C = (int)C; a.N = (int)N;
    a.cblocks = (int)(C / 64);
    a.tiles_x = (int)(W / CH_TW);
    a.per_img = a.tiles_x * (int)(H / CH_TH);
    a.n_tiles = (int)(N / bn);
    const long long m_tiles = (long long)B * a.per_img;
    // CTA pairs (cta_group::2; need an even number of pixel tiles). Measured on B200 (tools/conv_halo_bench.py,
    // profiles/r02_conv_halo_bench_pair.log): WITHOUT the fused transform the pair kernel beats the single-CTA one by 5-18 % everywhere;
    // WITH it (every caller) it wins where a block's MMAs leave the builders time — 256-wide tiles: 256^2 x 512->256 419 -> 399 us,
    // 128^2 x 512 210 -> 200 us, 256^2 x 256 221 -> 214 us — and loses at 128 / 160-wide tiles (512^2 x 256->128: 552 -> 609 us), where the
    // leader's MMA stream waits for the slower of two builder groups. Default: pairs for 256-wide tiles; COMA_HALO_PAIR=0 / 1 forces.
    static const int pair_env = getenv("COMA_HALO_PAIR") ? atoi(getenv("COMA_HALO_PAIR")) : -1;
    const bool pair = (pair_env == 1 || (pair_env < 0 && bn == 256)) && m_tiles % 2 == 0;
    const long long total = (pair ? m_tiles / 2 : m_tiles) * a.n_tiles;
    COMA_REQUIRE(total < (1LL << 31), "too many output tiles");
    a.total = (int)total;
    a.scale = scale; a.shift = shift; a.act_in = act_in;
    a.bias = bias; a.bias_rows = bias_rows; a.bias_rows_ld = bias_rows_ld > 0 ? bias_rows_ld : N;
    a.residual = (const __half *)residual; a.ldo = ldo; a.act_out = act_out; a.stats = stats;
    CUtensorMap tw, to;
    {
        cuuint64_t dims[4] = {(cuuint64_t)(9 * C), (cuuint64_t)N, 1, 1};
        cuuint64_t str[3] = {(cuuint64_t)ldw * 2, (cuuint64_t)ldw * 2, (cuuint64_t)ldw * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)(pair ? bn / 2 : bn), 1, 1};
        if (int e = halo_map(&tw, Wt, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, "halo weights")) return e;
    }
    {
        cuuint64_t dims[4] = {(cuuint64_t)N, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t str[3] = {(cuuint64_t)ldo * 2, (cuuint64_t)(W * ldo) * 2, (cuuint64_t)(H * W * ldo) * 2};
        cuuint32_t box[4] = {32, (cuuint32_t)CH_TW, 4, 1};
        if (int e = halo_map(&to, out_f16, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B, "halo output")) return e;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (pair) {
        switch (bn) {
            case 256: return launch_halo<256, true>(tw, to, a, st);
            case 160: return launch_halo<160, true>(tw, to, a, st);
            case 128: return launch_halo<128, true>(tw, to, a, st);
            default: return launch_halo<64, true>(tw, to, a, st);
        }
    }
    switch (bn) {
        case 256: return launch_halo<256, false>(tw, to, a, st);
        case 160: return launch_halo<160, false>(tw, to, a, st);
        case 128: return launch_halo<128, false>(tw, to, a, st);
        default: return launch_halo<64, false>(tw, to, a, st);
    }
}
