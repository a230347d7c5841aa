// G1 — fp16 GEMM on 5th-gen tensor cores for the inpainting UNet / VAE (HP-A):  D[M,N] = A[M,K] · W[N,K]^T (+ epilogue)
//
// Every dense contraction of the SD-1.5 inpainting UNet and the VAE — 1x1 convs, linear layers, the im2col form of the
// 3x3 convs, attention projections — is this shape with A = activations (row-major [M,K], K contiguous = NHWC) and
// W = weights ([N,K], K contiguous). Reference call sites: utils/adaptive_mask_inpainting.py:1001-1007 (UNet),
// :680,:1086,:1112 (VAE) — the arithmetic itself lives in diffusers' UNet2DConditionModel / AutoencoderKL [ext].
//
// Blackwell structure (persistent CTAs walking 128 x BN output tiles, warp-specialised):
//   warp 0    : TMA producer — cp.async.bulk.tensor.4d of the A (128 x 64) and W (BN x 64) K-slabs into a 3-5 stage
//               128B-swizzled shared-memory ring, completion on `full` mbarriers; for CONV the A slab is the output tile shifted by
//               one filter tap (implicit GEMM, hardware zero fill = padding, element stride 2 for strided convolutions)
//   warp 1    : MMA issuer — one elected thread issues 4 x tcgen05.mma.kind::f16 (M128 x BN x K16) per slab from shared-memory
//               descriptors into one of TWO fp32 TMEM accumulators; tcgen05.commit releases the slab (`empty`) and finally signals
//               `tmem_full`, so the epilogue of tile i overlaps the main loop of tile i+1
//   warps 2.. : epilogue (4 warps, 8 on the one-CTA-per-SM tiles) — tcgen05.ld of 32-column panels, + bias, + per-sample bias rows,
//               + residual (fetched by TMA), optional SiLU or fused GEGLU, fp16 panel staged in 64B-swizzled shared memory and
//               written by a TMA store (which clips the row / column tails); optionally the GroupNorm partial sums of the panel.
//               fp32 outputs and split-K slabs take the direct-store path.
// Host side: a small cost model picks the tile width and a split-K factor (deterministic slab reduction) for problems with too
// few tiles for 148 SMs; every launch uses programmatic dependent launch with a late trigger.
// Operands are fp16, accumulation fp32 (the reference runs the models in fp16, src/generation/inpaint.py:64).
#include <cuda.h>
#include <cuda_fp16.h>
#include <math.h>
#include <string.h>

#include "common.cuh"

namespace coma {

constexpr int G_BM = 128, G_BK = 64;
// warps: 0 = TMA producer, 1 = MMA issuer, 2.. = epilogue. One-CTA-per-SM tiles (BN > 128) run TWO epilogue warps per TMEM lane
// quarter (they split the tile's 32-column panels): a lone warp per SM sub-partition is issue-latency bound (~150 dependent
// instructions per panel) and cannot keep up with a short-K main loop. The BN <= 128 tiles get the same 8 warps per SM from
// their two resident CTAs.
#ifndef COMA_GEMM_WIDE_EPI_WARPS
#define COMA_GEMM_WIDE_EPI_WARPS 8
#endif
#ifndef COMA_GEMM_EPI_WARPS_160
#define COMA_GEMM_EPI_WARPS_160 12
#endif
__host__ __device__ constexpr int gemm_epi_warps(int BN) { return BN <= 128 ? 4 : (BN == 160 ? COMA_GEMM_EPI_WARPS_160 : COMA_GEMM_WIDE_EPI_WARPS); }
__host__ __device__ constexpr int gemm_threads(int BN) { return 64 + 32 * gemm_epi_warps(BN); }

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
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, uint64_t *bar, int x, int y, int z, int w) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z), "r"(w)
        : "memory");
}
// K-major, 128B-swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4, LBO = 1 (ignored for
// swizzled K-major), SBO = 1024 B (8 rows x 128 B) >> 4, version 1 (Blackwell), layout SWIZZLE_128B (2).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
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
__device__ __forceinline__ float exp2f_fast(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// Exact-form GELU x * Phi(x) for the fused GEGLU epilogue: erf through Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7 absolute on Phi — the
// result is rounded to fp16 right after), 13 instructions with two MUFU ops instead of libdevice erff's ~30: the 64^2-level projection
// (K = 320) is epilogue-bound (85.5 us against 69.2 us for the plain GEMM writing twice the output).
__device__ __forceinline__ float gelu_erf_fast(float x) {
    const float z = fabsf(x) * 0.70710678118654752f;
    const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
    float p = fmaf(t, 0.5f * 1.061405429f, 0.5f * -1.453152027f);
    p = fmaf(p, t, 0.5f * 1.421413741f);
    p = fmaf(p, t, 0.5f * -0.284496736f);
    p = fmaf(p, t, 0.5f * 0.254829592f);
    const float q = p * t * exp2f_fast(x * x * -0.72134752044448170f);   // 0.5 * (1 - erf(z)) = 0.5 * poly(t) * exp(-z^2), in [0, 0.5]
    return x * (x >= 0.0f ? 1.0f - q : q);
}
__device__ __forceinline__ bool elect_one() {   // one lane of the (converged) warp; the same lane every time for a full mask
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// ---- CTA pair (cta_group::2): two CTAs of a cluster (one TPC) own two adjacent 128-row tiles of the same N tile; the leader (cluster
// rank 0) issues M256 x BN x K16 instructions that read each CTA's own A slab and its HALF of the W slab, so an SM's operand stream
// (L2 -> shared memory, and the tensor core's shared-memory reads) per flop drops by a third at BN = 256. Same protocol as conv_halo.cu.
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t local, uint32_t rank) {   // shared::cluster address of a shared::cta address in CTA `rank`
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
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

// warp-synchronous TMEM load whose completion is awaited separately (tmem_wait_ld ties the registers to the wait)
__device__ __forceinline__ void tmem_ld32_async(uint32_t taddr, uint32_t *v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, "
        "%25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld(uint32_t *v) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]),
                   "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]),
                   "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]),
                   "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
                 :
                 : "memory");
}
// TMA store of a shared-memory box (bulk async group) and its completion primitives
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

struct GemmEpilogue {
    const float *bias;       // [N] or null
    const float *bias_rows;  // [ceil(M/rows_per_bias), N] or null (e.g. the per-sample time-embedding term of a ResnetBlock)
    int rows_per_bias;
    long long bias_rows_ld;  // row stride of bias_rows (>= N)
    const __half *residual;  // same layout as the output, or null
    __half *out16;
    float *out32;
    int ldo;
    long long o_s1, o_s2;    // output element strides of the two batch dims
    float alpha;             // scales the accumulator (attention: 1/sqrt(d))
    int act;                 // 0 = identity, 1 = SiLU, 2 = quick-GELU x*sigmoid(1.702x)
    int nb1;                 // extent of batch dim 1 (blockIdx.z = b2 * nb1 + b1)
    int tma;                 // 1: fp16 output (and residual) move as 32x32 boxes through shared memory + TMA (tmO / tmR)
    const float2 *ln_rows;   // or null: per-row (rstd, -rstd * mean) of a LayerNorm folded into W (W' = W * gamma): out = rstd * acc + (-rstd * mean) * ln_c1[n] + bias[n]
    const float *ln_c1;      // [N] column sums of W' (with ln_rows / ln_part_in)
    const float2 *ln_part_in;  // or null: instead of ln_rows, per-(row, 32-column block) (sum, sumsq) of the input row left by its PRODUCER's
    int ln_parts;              //   epilogue (ln_part_out): [M, ln_parts] float2, ln_parts = K / 32; the row statistics are formed here
    float ln_eps;
    float2 *ln_part_out;     // or null: TMA form only — leave (sum, sumsq) of every 32-column panel of the rounded fp16 OUTPUT row: [M, N/32] float2
    float *stats;            // TMA form only, or null: [M/32, N, 2] per-(32-row block, column) sum / sum of squares of the fp16
                             // output, for the GroupNorm that consumes it (coma_groupnorm_from_stats_f32)
};

constexpr int E_PANEL_BYTES = 32 * 32 * 2;  // one epilogue panel: 32 rows x 32 fp16 columns, 64-byte rows, 64B-swizzled
#ifndef COMA_GEMM_EPI_BUFS_160
#define COMA_GEMM_EPI_BUFS_160 2
#endif
// staging panels per epilogue warp: the residual panel is prefetched NBUF - 1 panels ahead and NBUF - 1 stores may be in flight
__host__ __device__ constexpr int gemm_epi_bufs(int BN) { return BN == 160 ? COMA_GEMM_EPI_BUFS_160 : 2; }
// weight-stationary form: A ring depth next to WS resident W slabs
__host__ __device__ constexpr int gemm_ws_stages(int BN, int WS) { return BN == 160 ? (gemm_epi_bufs(160) > 2 ? 3 : 4) : (BN == 192 ? 4 : 3); }

// Implicit-GEMM 3x3 convolution (stride 1, pad 1) on an NHWC tensor: the A operand of K-slab (tap, channel block) is the
// 128-pixel output tile shifted by (ky-1, kx-1), fetched by ONE 4-D TMA load whose out-of-image coordinates are
// zero-filled by the hardware (= the convolution's zero padding). No im2col matrix is ever materialised.
struct ConvGeom {
    int H, W, B;         // image extent
    int TW, TH, TB;      // tile = TB images x TH rows x TW columns = 128 output pixels
    int tiles_x, tiles_y;
    int cblocks;         // input channels / 64
    int stride, pad;     // input pixel of output (y, x), tap (ky, kx): (y*stride + ky - pad, x*stride + kx - pad)
};

struct TileSched {
    int n_tiles, m_tiles, total;  // total = n_tiles * m_tiles * batches * ksplit
    // split-K: work item id = tile id * ksplit + ks; slice ks reduces K-slabs [ks*kb_per, (ks+1)*kb_per) into its own fp32
    // slab of the workspace (out32 + ks * split_stride) — a fixed-order finishing pass sums the slabs (deterministic)
    int ksplit, kb_per;
    long long split_stride;
};

// Persistent, warp-specialised kernel: gridDim.x CTAs walk the tile list with stride gridDim.x. The fp32 accumulator is
// double-buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps the main loop of tile i+1; consecutive tile
// ids share the A tile (N fastest), which keeps the activation slab L2-resident while its N-tiles are produced.
// BN in {64, 128, 160, 256}: wider tiles raise the flop/byte of the operand stream (L2 -> SMEM is what bounds a 128x128
// tile at ~0.75 PFLOP/s on this part: 32 KB per 2.1 MFLOP).
// GEGLU (BN = 256 plain GEMM only): the weight rows of the feed-forward projection are interleaved in blocks of 32 (32 value
// rows, then their 32 gate rows: prep_geglu in nn.py), so accumulator columns [64p, 64p+32) / [64p+32, 64p+64) of a tile are
// value / gate of the same 32 features and the epilogue writes value * gelu(gate) — half as many output columns, no
// [M, 8C] intermediate and no separate GEGLU pass (diffusers GEGLU inside BasicTransformerBlock.ff).
// PAIR: launched as clusters of two CTAs; ts.total counts PAIR items (two adjacent M tiles x one N tile), ts.m_tiles pairs of M tiles.
// WS > 0 ("weight-stationary", plain / GEGLU GEMMs with K <= 64 * WS): a CTA keeps ONE N tile for its whole life — its W tile (BN x K, all
// WS slabs) is loaded into shared memory once and only A slabs stream through the ring. Short-K projections are bound by the L2 -> SM
// operand stream (K = 320, BN = 160: 180 KB per tile, of which 100 KB is the same W tile again); CTA c owns N tile c % n_tiles and walks
// the M tiles c / n_tiles, + gridDim / n_tiles, ... (ts.total = M tiles).
template <int BN, bool CONV, int STAGES, bool GEGLU = false, bool PAIR = false, int WS = 0>
__global__ void __launch_bounds__(gemm_threads(BN), ((BN <= 128 && !PAIR && !WS) ? 2 : 1))
    gemm_f16_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmR, int M, int N, int K,
                       const GemmEpilogue ep, const ConvGeom cg, const TileSched ts) {
    // The two-CTA-per-SM configurations (BN <= 128) have no room for alignment slack: the dynamic window is requested
    // 1024-byte aligned (128B-swizzle atoms) and the kernel traps if the toolchain ever places it otherwise.
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = smem_raw;
    if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
    constexpr int A_BYTES = G_BM * G_BK * 2, B_BYTES = (PAIR ? BN / 2 : BN) * G_BK * 2;   // PAIR: this CTA's half of the W slab's rows
    constexpr int NBUF = gemm_epi_bufs(BN);
    constexpr int EPI_WARPS = gemm_epi_warps(BN);
    constexpr int PW = EPI_WARPS / 4;       // epilogue warps per TMEM lane quarter = panel stride of one warp
    constexpr int B_SLOTS = WS ? WS : STAGES;
    uint8_t *sA = smem, *sB = smem + STAGES * A_BYTES;
    uint8_t *sE = sB + B_SLOTS * B_BYTES;    // epilogue staging: EPI_WARPS warps x NBUF panels
    uint64_t *full = reinterpret_cast<uint64_t *>(sE + EPI_WARPS * NBUF * E_PANEL_BYTES);
    uint64_t *empty = full + STAGES;
    uint64_t *tmem_full = empty + STAGES;   // [2]
    uint64_t *tmem_empty = tmem_full + 2;   // [2]
    uint64_t *res_full = tmem_empty + 2;    // [EPI_WARPS][NBUF]: residual panel landed
    uint64_t *w_full = res_full + EPI_WARPS * NBUF;   // WS: the resident W tile has landed
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(w_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_k = (K + G_BK - 1) / G_BK;
    const uint32_t rank = PAIR ? cluster_rank() : 0u;
    const int t0 = WS ? (int)blockIdx.x / ts.n_tiles : (PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x);
    const int tstep = WS ? (int)gridDim.x / ts.n_tiles : (PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x);
    constexpr uint32_t TMEM_COLS = (2 * BN <= 128) ? 128 : ((2 * BN <= 256) ? 256 : 512);  // power of two >= 2*BN

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        if (ep.tma) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmO) : "memory");
            if (ep.residual) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmR) : "memory");
        }
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(full + i, 1);
            mbar_init(empty + i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(tmem_full + i, 1);
            mbar_init(tmem_empty + i, PAIR ? 2 * EPI_WARPS : 32 * EPI_WARPS);   // PAIR: one elected arrival per epilogue warp of both CTAs
        }
        for (int i = 0; i < EPI_WARPS * NBUF; ++i) mbar_init(res_full + i, 1);
        mbar_init(w_full, 1);
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
    __syncthreads();
    if (PAIR) cluster_sync_all();   // the peer's barriers must exist before anything signals them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    // PDL: everything above overlapped the previous kernel's tail; its results are visible after pdl_wait(). The trigger for
    // OUR dependents is raised late (producer warp, after its last load): a persistent kernel that triggers at its start
    // shares its SMs for its whole duration with dependent CTAs parked in griddepcontrol.wait, and those cost issue slots
    // (measured: VAE decode 12.0 -> 14.8 ms with light-weight kernels behind every conv).
    pdl_wait();
    const uint32_t empty0 = PAIR ? map_to_rank(smem_u32(tmem_empty), 0) : 0u;   // the leader's tmem_empty[0]

    // tile id -> coordinates
    auto decode = [&](int t, int &m0, int &n0, int &b1, int &b2, int &cx0, int &cy0, int &cb0) {
        t /= ts.ksplit;
        const int nt = WS ? (int)blockIdx.x % ts.n_tiles : t % ts.n_tiles;
        const int r = WS ? t : t / ts.n_tiles;
        const int mt = PAIR ? 2 * (r % ts.m_tiles) + (int)rank : r % ts.m_tiles, z = r / ts.m_tiles;
        n0 = nt * BN;
        m0 = mt * G_BM;
        b1 = z % ep.nb1;
        b2 = z / ep.nb1;
        cx0 = cy0 = cb0 = 0;
        if (CONV) {
            if (cg.TB > 1) {
                cb0 = mt * cg.TB;
            } else {
                const int per_img = cg.tiles_x * cg.tiles_y;
                cb0 = mt / per_img;
                const int rem = mt % per_img;
                cy0 = (rem / cg.tiles_x) * cg.TH;
                cx0 = (rem % cg.tiles_x) * cg.TW;
            }
        }
    };

    // Producer and MMA issuer run as CONVERGED warps with one elected lane per issue group. Inside an `if (lane == 0)` region ptxas
    // wraps every UTMALDG / UTCHMMA / UTCBAR in an ELECT retry loop and rebuilds the shared-memory descriptors per slab on the
    // uniform datapath; with elect.sync the issue sequence is straight-line code (attention.cu measured 110-240 -> ~50 clk per MMA).
    // Ring stage / phase and the convolution's (tap, channel block) are carried incrementally: no division per K-slab.
    if (warp == 0) {
        int s = 0;
        uint32_t ph = 0;
        if (WS && t0 < ts.total && elect_one()) {   // the W tile of this CTA's N tile: once
            mbar_expect_tx(w_full, (uint32_t)num_k * B_BYTES);
            for (int kb = 0; kb < num_k; ++kb) tma_load_4d(sB + kb * B_BYTES, &tmB, w_full, kb * G_BK, ((int)blockIdx.x % ts.n_tiles) * BN, 0, 0);
        }
        __syncwarp();
        for (int t = t0; t < ts.total; t += tstep) {
            int m0, n0, b1, b2, cx0, cy0, cb0;
            decode(t, m0, n0, b1, b2, cx0, cy0, cb0);
            const int kb0 = (t % ts.ksplit) * ts.kb_per, kb1 = min(num_k, kb0 + ts.kb_per);
            int tap = 0, cb = 0;
            if (CONV) {
                tap = kb0 / cg.cblocks;
                cb = kb0 - tap * cg.cblocks;
            }
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(empty + s, ph ^ 1);
                if (elect_one()) {
                    if (PAIR) {   // both CTAs' slabs complete on the LEADER's barrier (cta_group::2 loads)
                        if (rank == 0) mbar_expect_tx(full + s, 2 * (A_BYTES + B_BYTES));
                        if (CONV) {
                            const int ty = (tap * 11) >> 5;
                            tma_load_4d_pair(sA + s * A_BYTES, &tmA, full + s, cb * G_BK, cx0 * cg.stride + (tap - 3 * ty) - cg.pad,
                                             cy0 * cg.stride + ty - cg.pad, cb0);
                        } else {
                            tma_load_4d_pair(sA + s * A_BYTES, &tmA, full + s, kb * G_BK, m0, b1, b2);
                        }
                        tma_load_4d_pair(sB + s * B_BYTES, &tmB, full + s, kb * G_BK, n0 + (int)rank * (BN / 2), b1, b2);
                    } else if (WS) {
                        mbar_expect_tx(full + s, A_BYTES);
                        tma_load_4d(sA + s * A_BYTES, &tmA, full + s, kb * G_BK, m0, b1, b2);
                    } else {
                    mbar_expect_tx(full + s, A_BYTES + B_BYTES);
                    if (CONV) {
                        const int ty = (tap * 11) >> 5;   // tap / 3 for tap in [0, 9)
                        tma_load_4d(sA + s * A_BYTES, &tmA, full + s, cb * G_BK, cx0 * cg.stride + (tap - 3 * ty) - cg.pad,
                                    cy0 * cg.stride + ty - cg.pad, cb0);
                    } else {
                        tma_load_4d(sA + s * A_BYTES, &tmA, full + s, kb * G_BK, m0, b1, b2);
                    }
                    tma_load_4d(sB + s * B_BYTES, &tmB, full + s, kb * G_BK, n0, b1, b2);
                    }
                }
                __syncwarp();
                if (CONV && ++cb == cg.cblocks) {
                    cb = 0;
                    ++tap;
                }
                if (++s == STAGES) {
                    s = 0;
                    ph ^= 1;
                }
            }
        }
        if (lane == 0) pdl_trigger();  // all of this CTA's operands are on their way: dependents may start launching
    } else if (warp == 1) {
        // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = F16, both K-major, N>>3, M>>4
        constexpr uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((PAIR ? 2 * G_BM : G_BM) >> 4) << 24);
        const uint64_t da0 = umma_desc_sw128(smem_u32(sA)), db0 = umma_desc_sw128(smem_u32(sB));
        int s = 0, i = 0;
        uint32_t ph = 0;
        if (WS && t0 < ts.total) mbar_wait(w_full, 0);
        for (int t = t0; t < ts.total && rank == 0; t += tstep, ++i) {   // PAIR: the leader CTA issues for both
            const int acc = i & 1;
            if (PAIR) mbar_wait_cluster(tmem_empty + acc, ((i >> 1) & 1) ^ 1);
            else mbar_wait(tmem_empty + acc, ((i >> 1) & 1) ^ 1);  // the epilogue has drained this accumulator
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
            const int kb0 = (t % ts.ksplit) * ts.kb_per, kb1 = min(num_k, kb0 + ts.kb_per);
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(full + s, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (elect_one()) {
                    const uint64_t da = da0 + (uint64_t)(s * (A_BYTES >> 4)), db = db0 + (uint64_t)((WS ? kb : s) * (B_BYTES >> 4));
                    if (PAIR) {
#pragma unroll
                        for (int k = 0; k < G_BK / 16; ++k) umma_f16_pair(tmem_d, da + 2 * k, db + 2 * k, idesc, ((kb - kb0) | k) != 0);
                        umma_commit_pair(empty + s);   // multicast: frees the slab in both CTAs
                        if (kb + 1 == kb1) umma_commit_pair(tmem_full + acc);
                    } else {
#pragma unroll
                    for (int k = 0; k < G_BK / 16; ++k)  // advance 32 B (16 fp16) inside the 128 B swizzle atom: +2 in 16-B units
                        umma_f16(tmem_d, da + 2 * k, db + 2 * k, idesc, ((kb - kb0) | k) != 0);
                    umma_commit(empty + s);  // frees the slab once the MMAs that read it have retired
                    if (kb + 1 == kb1) umma_commit(tmem_full + acc);
                    }
                }
                __syncwarp();
                if (++s == STAGES) {
                    s = 0;
                    ph ^= 1;
                }
            }
        }
    } else if (ep.tma) {
        // ---- epilogue, TMA form: a warp owns TMEM lanes [32q, 32q+32) = 32 output rows (always contiguous in memory, also
        // for the conv tiles: TW == W whenever a tile spans several image rows). Per 32-column panel: the residual panel
        // arrives by TMA into a 64B-swizzled 2 KB buffer (prefetched NBUF-1 panels ahead, across tile boundaries), each
        // thread folds accumulator + bias (+ residual, SiLU) for its row IN PLACE, and one lane hands the buffer to a TMA
        // store, which also clips rows >= M / columns >= N. No per-thread global stores, no cross-warp synchronisation.
        const float *__restrict__ bias = ep.bias;
        const int act = ep.act;
        const float alpha = ep.alpha;
        const int q = warp & 3;            // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;  // which of the PW interleaved panel subsets this warp owns
        const bool has_res = ep.residual != nullptr;
        uint8_t *ebuf = sE + (warp - 2) * (NBUF * E_PANEL_BYTES);
        uint64_t *rb = res_full + (warp - 2) * NBUF;
        uint8_t *my_row = ebuf + lane * 64;
        const int sw = (lane >> 1) & 3;  // 64B swizzle: 16-byte chunk index ^= (row / 2) % 4
        // load cursor (lane 0 only): walks the same (tile, panel) stream as the consumer loop, NBUF-1 panels ahead
        int lt = t0, lp = 0, lP = 0, lrow0 = 0, ln0 = 0, lb1 = 0, lb2 = 0;
        uint32_t lg = 0;
        auto cursor_tile = [&]() {  // position the cursor on this warp's first panel of tile lt (skipping tiles without one)
            while (lt < ts.total) {
                int m0, n0, b1, b2, cx0, cy0, cb0;
                decode(lt, m0, n0, b1, b2, cx0, cy0, cb0);
                lrow0 = CONV ? ((cb0 * cg.H + cy0) * cg.W + cx0) : m0;
                ln0 = n0; lb1 = b1; lb2 = b2; lp = half;
                lP = (min(BN, N - n0) + 31) >> 5;
                if (lp < lP) break;
                lt += tstep;
            }
        };
        auto cursor_issue = [&]() {  // issue the residual load of the cursor's panel, then advance
            if (lt >= ts.total) return;
            const uint32_t buf = lg % NBUF;
            mbar_expect_tx(rb + buf, E_PANEL_BYTES);
            tma_load_4d(ebuf + buf * E_PANEL_BYTES, &tmR, rb + buf, ln0 + lp * 32, lrow0 + q * 32, lb1, lb2);
            ++lg;
            lp += PW;
            if (lp >= lP) {
                lt += tstep;
                cursor_tile();
            }
        };
        if (has_res && lane == 0) {
            cursor_tile();
#pragma unroll 1
            for (int i = 0; i < NBUF - 1; ++i) cursor_issue();
        }
        uint32_t g = 0;
        int i = 0;
        for (int t = t0; t < ts.total; t += tstep, ++i) {
            int m0, n0, b1, b2, cx0, cy0, cb0;
            decode(t, m0, n0, b1, b2, cx0, cy0, cb0);
            const int row0 = (CONV ? ((cb0 * cg.H + cy0) * cg.W + cx0) : m0) + q * 32;
            constexpr int PCOLS = GEGLU ? 64 : 32;  // accumulator columns behind one 32-column output panel
            const int P = (min(BN, N - n0) + PCOLS - 1) / PCOLS;
            const int brow_row = min(row0 + lane, M - 1) / ep.rows_per_bias;
            const float *brow = ep.bias_rows ? ep.bias_rows + (size_t)brow_row * ep.bias_rows_ld : nullptr;
            float2 ln = ep.ln_rows ? __ldg(ep.ln_rows + min(row0 + lane, M - 1)) : make_float2(alpha, 0.0f);   // folded LayerNorm: row scale, row offset
            if (ep.ln_part_in) {   // statistics of this thread's input row from its producer's per-panel sums, added in panel order
                const float2 *pp = ep.ln_part_in + (size_t)min(row0 + lane, M - 1) * ep.ln_parts;
                float sx = 0.f, sq = 0.f;
                for (int i2 = 0; i2 < ep.ln_parts; ++i2) {
                    const float2 v2 = __ldg(pp + i2);
                    sx += v2.x;
                    sq += v2.y;
                }
                const float inv_c = 1.0f / (float)(32 * ep.ln_parts);
                const float mean = sx * inv_c, var = fmaxf(fmaf(-mean, mean, sq * inv_c), 0.0f);
                const float rstd = rsqrtf(var + ep.ln_eps);
                ln = make_float2(rstd, -rstd * mean);
            }
            const int acc = i & 1;
            mbar_wait(tmem_full + acc, (i >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN) + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
            for (int p = half; p < P; p += PW, ++g) {
                const uint32_t buf = g % NBUF;
                uint8_t *prow = my_row + buf * E_PANEL_BYTES;
                uint32_t v[32];
                tmem_ld32_async(tmem_d + (uint32_t)(p * PCOLS), v);
                const int nb = n0 + p * PCOLS;
                float f[32];
                if constexpr (GEGLU) {
                    uint32_t gt[32];
                    tmem_ld32_async(tmem_d + (uint32_t)(p * PCOLS + 32), gt);
                    if (lane == 0) bulk_wait_read<NBUF - 1>();
                    __syncwarp();
                    tmem_wait_ld(v);
                    tmem_wait_ld(gt);
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 bv = bias ? __ldg(reinterpret_cast<const float4 *>(bias + nb + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
                        const float4 bg = bias ? __ldg(reinterpret_cast<const float4 *>(bias + nb + 32 + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
                        float bvs[4] = {bv.x, bv.y, bv.z, bv.w}, bgs[4] = {bg.x, bg.y, bg.z, bg.w};
                        if (ep.ln_c1) {   // folded LayerNorm: + (-rstd * mean) * column sum of W'
                            const float4 cv = __ldg(reinterpret_cast<const float4 *>(ep.ln_c1 + nb + j)), cg2 = __ldg(reinterpret_cast<const float4 *>(ep.ln_c1 + nb + 32 + j));
                            bvs[0] = fmaf(ln.y, cv.x, bvs[0]); bvs[1] = fmaf(ln.y, cv.y, bvs[1]); bvs[2] = fmaf(ln.y, cv.z, bvs[2]); bvs[3] = fmaf(ln.y, cv.w, bvs[3]);
                            bgs[0] = fmaf(ln.y, cg2.x, bgs[0]); bgs[1] = fmaf(ln.y, cg2.y, bgs[1]); bgs[2] = fmaf(ln.y, cg2.z, bgs[2]); bgs[3] = fmaf(ln.y, cg2.w, bgs[3]);
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            // value and gate pass through fp16 like the unfused projection output did (same roundings)
                            const float a = __half2float(__float2half_rn(fmaf(__uint_as_float(v[j + u]), ln.x, bvs[u])));
                            const float x = __half2float(__float2half_rn(fmaf(__uint_as_float(gt[j + u]), ln.x, bgs[u])));
                            f[j + u] = a * gelu_erf_fast(x);
                        }
                    }
                } else {
                if (nb + 32 <= N) {
                    if (bias) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4 *>(bias + nb + j));
                            f[j] = b4.x; f[j + 1] = b4.y; f[j + 2] = b4.z; f[j + 3] = b4.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = 0.0f;
                    }
                    if (brow) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4 *>(brow + nb + j));
                            f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
                        }
                    }
                } else {  // last, partial panel of the matrix
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const bool ok = nb + j < N;
                        f[j] = (ok && bias) ? __ldg(bias + nb + j) : 0.0f;
                        if (ok && brow) f[j] += __ldg(brow + nb + j);
                    }
                }
                if (ep.ln_c1) {   // folded LayerNorm: + (-rstd * mean) * column sum of W' (N % 32 == 0 is required with ln)
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 c4 = __ldg(reinterpret_cast<const float4 *>(ep.ln_c1 + nb + j));
                        f[j] = fmaf(ln.y, c4.x, f[j]); f[j + 1] = fmaf(ln.y, c4.y, f[j + 1]); f[j + 2] = fmaf(ln.y, c4.z, f[j + 2]); f[j + 3] = fmaf(ln.y, c4.w, f[j + 3]);
                    }
                }
                if (has_res) {
                    mbar_wait(rb + buf, (g / NBUF) & 1);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const uint4 r = *reinterpret_cast<const uint4 *>(prow + ((c ^ sw) << 4));
                        const __half2 *h = reinterpret_cast<const __half2 *>(&r);
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float2 x = __half22float2(h[u]);
                            f[c * 8 + 2 * u] += x.x;
                            f[c * 8 + 2 * u + 1] += x.y;
                        }
                    }
                } else {
                    if (lane == 0) bulk_wait_read<NBUF - 1>();  // the store that last read this buffer has drained it
                    __syncwarp();
                }
                tmem_wait_ld(v);
                const float2 alpha2 = make_float2(ln.x, ln.x);   // alpha, or the row's rstd when a LayerNorm is folded into W
#pragma unroll
                for (int j = 0; j < 32; j += 2) {  // packed FP32x2 FMAs
                    const float2 r = __ffma2_rn(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), alpha2,
                                                make_float2(f[j], f[j + 1]));
                    f[j] = r.x;
                    f[j + 1] = r.y;
                }
                if (act == 1) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = __fdividef(f[j], 1.0f + __expf(-f[j]));
                } else if (act == 2) {   // quick-GELU (CLIP text encoder): x * sigmoid(1.702 x)
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = __fdividef(f[j], 1.0f + __expf(-1.702f * f[j]));
                }
                }  // !GEGLU
                float2 ls = make_float2(0.f, 0.f), lq = make_float2(0.f, 0.f);   // LayerNorm partial sums of the ROUNDED row (ln_part_out)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint4 w;
                    __half2 *h = reinterpret_cast<__half2 *>(&w);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        h[u] = __floats2half2_rn(f[c * 8 + 2 * u], f[c * 8 + 2 * u + 1]);
                        if (!GEGLU && ep.ln_part_out) {
                            const float2 r2 = __half22float2(h[u]);
                            ls = __fadd2_rn(ls, r2);
                            lq = __ffma2_rn(r2, r2, lq);
                        }
                    }
                    *reinterpret_cast<uint4 *>(prow + ((c ^ sw) << 4)) = w;
                }
                if (!GEGLU && ep.ln_part_out && row0 + lane < M && nb + 32 <= N)
                    ep.ln_part_out[(size_t)(row0 + lane) * (N >> 5) + (nb >> 5)] = make_float2(ls.x + ls.y, lq.x + lq.y);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the TMA engine
                __syncwarp();
                if (lane == 0) {
                    tma_store_4d(&tmO, ebuf + buf * E_PANEL_BYTES, GEGLU ? (n0 >> 1) + p * 32 : nb, row0, b1, b2);
                    bulk_commit();
                    if (has_res) {
                        bulk_wait_read<1>();  // the previous panel's store has finished reading its buffer: refill it
                        cursor_issue();
                    }
                }
                if (!GEGLU && ep.stats && row0 < M) {
                    // GroupNorm statistics of the tensor being written, for free: the panel sits in shared memory as the
                    // ROUNDED fp16 values the next GroupNorm will see. Lane (cp, par) adds the 16 rows of parity `par` of
                    // column pair cp (conflict-free: a 64-byte row spans 16 banks, odd rows the other 16), the two parities
                    // are combined in a fixed order and lanes 0-15 write (sum, sumsq) x 2 columns for this 32-row block.
                    const int cp = lane & 15, par = lane >> 4;
                    const uint8_t *pan = ebuf + buf * E_PANEL_BYTES;
                    float2 s2 = make_float2(0.f, 0.f), q2 = make_float2(0.f, 0.f);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int r = 2 * i + par;
                        const __half2 h = *reinterpret_cast<const __half2 *>(pan + r * 64 + ((((cp >> 2) ^ ((r >> 1) & 3))) << 4) + (cp & 3) * 4);
                        const float2 x = __half22float2(h);
                        s2 = __fadd2_rn(s2, x);
                        q2 = __ffma2_rn(x, x, q2);
                    }
                    s2.x += __shfl_down_sync(0xffffffffu, s2.x, 16);
                    s2.y += __shfl_down_sync(0xffffffffu, s2.y, 16);
                    q2.x += __shfl_down_sync(0xffffffffu, q2.x, 16);
                    q2.y += __shfl_down_sync(0xffffffffu, q2.y, 16);
                    const int col = nb + 2 * cp;
                    if (par == 0 && col < N) {
                        float *dst = ep.stats + ((size_t)(row0 >> 5) * N + col) * 2;
                        if (col + 1 < N) *reinterpret_cast<float4 *>(dst) = make_float4(s2.x, q2.x, s2.y, q2.y);
                        else *reinterpret_cast<float2 *>(dst) = make_float2(s2.x, q2.x);
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
        if (lane == 0) bulk_wait_read<0>();  // shared memory must outlive the last stores' reads
    } else {
        const float *__restrict__ bias = ep.bias;
        const int act = ep.act, ldo = ep.ldo;
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        int i = 0;
        for (int t = t0; t < ts.total; t += tstep, ++i) {
            int m0, n0, b1, b2, cx0, cy0, cb0;
            decode(t, m0, n0, b1, b2, cx0, cy0, cb0);
            const size_t obase = (size_t)b1 * ep.o_s1 + (size_t)b2 * ep.o_s2 + (size_t)(t % ts.ksplit) * ts.split_stride;
            const __half *__restrict__ residual = ep.residual ? ep.residual + obase : nullptr;
            __half *__restrict__ out16 = ep.out16 ? ep.out16 + obase : nullptr;
            float *__restrict__ out32 = ep.out32 ? ep.out32 + obase : nullptr;
            const int acc = i & 1;
            mbar_wait(tmem_full + acc, (i >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
            int row = m0 + q * 32 + lane;
            if (CONV) {  // TMEM lane -> pixel of the tile (x fastest, then y, then image)
                const int r = q * 32 + lane;
                const int tx = r % cg.TW, ty = (r / cg.TW) % cg.TH, tb = r / (cg.TW * cg.TH);
                row = (cb0 + tb < cg.B) ? ((cb0 + tb) * cg.H + cy0 + ty) * cg.W + cx0 + tx : M;
            }
#pragma unroll 1
            for (int c0 = ((warp - 2) >> 2) * 32; c0 < BN; c0 += 32 * PW) {
                uint32_t v[32];
                tmem_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
                if (row < M && n0 + c0 < N) {
                    const int nb = n0 + c0;
                    const size_t off = (size_t)row * ldo + nb;
                    const float *brow = ep.bias_rows ? ep.bias_rows + (size_t)(row / ep.rows_per_bias) * ep.bias_rows_ld : nullptr;
                    if (nb + 32 <= N && (ldo % 8) == 0 && (obase % 8) == 0) {
                        float f[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            f[j] = __uint_as_float(v[j]) * ep.alpha + (bias ? __ldg(bias + nb + j) : 0.0f);
                            if (brow) f[j] += __ldg(brow + nb + j);
                        }
                        if (residual) {
#pragma unroll
                            for (int j = 0; j < 32; j += 8) {
                                const uint4 r = *reinterpret_cast<const uint4 *>(residual + off + j);
                                const __half2 *h = reinterpret_cast<const __half2 *>(&r);
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    const float2 x = __half22float2(h[u]);
                                    f[j + 2 * u] += x.x;
                                    f[j + 2 * u + 1] += x.y;
                                }
                            }
                        }
                        if (act == 1) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) f[j] = __fdividef(f[j], 1.0f + __expf(-f[j]));
                        } else if (act == 2) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) f[j] = __fdividef(f[j], 1.0f + __expf(-1.702f * f[j]));
                        }
                        if (out16) {
#pragma unroll
                            for (int j = 0; j < 32; j += 8) {
                                uint4 w;
                                __half2 *h = reinterpret_cast<__half2 *>(&w);
#pragma unroll
                                for (int u = 0; u < 4; ++u) h[u] = __floats2half2_rn(f[j + 2 * u], f[j + 2 * u + 1]);
                                *reinterpret_cast<uint4 *>(out16 + off + j) = w;
                            }
                        }
                        if (out32) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4)
                                *reinterpret_cast<float4 *>(out32 + off + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                        }
                    } else {
                        for (int j = 0; j < 32; ++j) {
                            if (nb + j < N) {
                                float x = __uint_as_float(v[j]) * ep.alpha + (bias ? bias[nb + j] : 0.0f);
                                if (brow) x += brow[nb + j];
                                if (residual) x += __half2float(residual[off + j]);
                                if (act == 1) x = __fdividef(x, 1.0f + __expf(-x));
                                else if (act == 2) x = __fdividef(x, 1.0f + __expf(-1.702f * x));
                                if (out16) out16[off + j] = __float2half_rn(x);
                                if (out32) out32[off + j] = x;
                            }
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            if (PAIR) {
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(empty0 + (uint32_t)acc * 8);
            } else {
                mbar_arrive(tmem_empty + acc);  // 128 arrivals hand the accumulator back to the MMA warp
            }
        }
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

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 4-D fp16 tensor map {cols, rows, nb1, nb2} over row-major [rows, cols] matrices with row stride `ld` and batch strides
// s1, s2 (elements); box = 64 columns x box_rows rows x 1 x 1.
static int make_map(CUtensorMap *m, const void *ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows, int64_t nb1, int64_t s1,
                    int64_t nb2, int64_t s2) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return COMA_E_NODEVICE;
    }
    cuuint64_t dims[4] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)nb1, (cuuint64_t)nb2};
    // a batch dim of extent 1 may carry any (valid) stride
    cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)(nb1 > 1 ? s1 : ld) * 2, (cuuint64_t)(nb2 > 1 ? s2 : ld) * 2};
    cuuint32_t box[4] = {(cuuint32_t)G_BK, (cuuint32_t)box_rows, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void *>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%lld ld=%lld)", (int)r, (long long)rows,
                  (long long)cols, (long long)ld);
        return COMA_E_BADARG;
    }
    return 0;
}

// fp16 [rows, cols] matrix (row stride ld, batch strides s1 / s2) as 32 x 32 boxes with 64-byte rows, 64B-swizzled: the
// epilogue's staging panels (TMA store of the output, TMA load of the residual).
static int make_panel_map(CUtensorMap *m, const void *ptr, int64_t rows, int64_t cols, int64_t ld, int64_t nb1, int64_t s1, int64_t nb2,
                          int64_t s2) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return COMA_E_NODEVICE;
    }
    cuuint64_t dims[4] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)nb1, (cuuint64_t)nb2};
    cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)(nb1 > 1 ? s1 : ld) * 2, (cuuint64_t)(nb2 > 1 ? s2 : ld) * 2};
    cuuint32_t box[4] = {32, 32, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void *>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (epilogue panel) failed with CUresult %d (rows=%lld cols=%lld ld=%lld)", (int)r,
                  (long long)rows, (long long)cols, (long long)ld);
        return COMA_E_BADARG;
    }
    return 0;
}

// Decides whether the epilogue can run in its TMA form and builds the output / residual panel maps.
static int setup_epilogue_maps(GemmEpilogue &ep, CUtensorMap *to, CUtensorMap *tr, int64_t M, int64_t N, int64_t nb1, int64_t nb2) {
    memset(to, 0, sizeof(*to));
    memset(tr, 0, sizeof(*tr));
    const bool ok = ep.out16 && !ep.out32 && ep.ldo % 8 == 0 && (nb1 == 1 || ep.o_s1 % 8 == 0) && (nb2 == 1 || ep.o_s2 % 8 == 0) &&
                    (uintptr_t)ep.out16 % 16 == 0 && (uintptr_t)ep.residual % 16 == 0 && (uintptr_t)ep.bias % 16 == 0 &&
                    (uintptr_t)ep.bias_rows % 16 == 0 && ep.bias_rows_ld % 4 == 0;
    ep.tma = ok ? 1 : 0;
    if (!ok) return 0;
    if (int e = make_panel_map(to, ep.out16, M, N, ep.ldo, nb1, ep.o_s1, nb2, ep.o_s2)) return e;
    if (ep.residual)
        if (int e = make_panel_map(tr, ep.residual, M, N, ep.ldo, nb1, ep.o_s1, nb2, ep.o_s2)) return e;
    return 0;
}

template <int BN, bool CONV, bool GEGLU = false, bool PAIR = false, int WS = 0>
static int launch_gemm(const CUtensorMap &ta, const CUtensorMap &tb, const CUtensorMap &to, const CUtensorMap &tr, int M, int N, int K,
                       const GemmEpilogue &ep, int nbatch, cudaStream_t st, int ksplit, long long split_stride,
                       const ConvGeom &cg = ConvGeom{}, int m_tiles_conv = 0) {
    // short-K problems are TMA-latency bound: as many slabs in flight as shared memory allows (PAIR: half a W slab per CTA)
    constexpr int STAGES = WS ? gemm_ws_stages(BN, WS) : (PAIR ? (BN == 256 ? 6 : ((gemm_epi_bufs(160) > 2 || gemm_epi_warps(160) > 8) ? 6 : 7)) : (BN <= 128 ? 3 : (BN == 160 ? ((gemm_epi_bufs(160) > 2 || gemm_epi_warps(160) > 8) ? 4 : 5) : 4)));
    // operand ring + epilogue panels + mbarriers (full / empty per stage, 2 + 2 accumulator barriers, one per residual panel, w_full) + the TMEM slot
    constexpr size_t bars = (size_t)(2 * STAGES + 5 + gemm_epi_warps(BN) * gemm_epi_bufs(BN)) * 8 + 16;
    constexpr size_t smem = (size_t)STAGES * G_BM * G_BK * 2 + (size_t)(WS ? WS : STAGES) * ((PAIR ? BN / 2 : BN) * G_BK * 2) +
                            gemm_epi_warps(BN) * gemm_epi_bufs(BN) * E_PANEL_BYTES + (bars > 256 ? 512 : 256);
    static_assert(bars <= 512 && smem <= 232448, "shared-memory budget");
    static bool attr[16] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 16 && !attr[dev]) {
        cudaError_t e = cudaFuncSetAttribute(gemm_f16_tn_kernel<BN, CONV, STAGES, GEGLU, PAIR, WS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(gemm): %s", cudaGetErrorString(e));
            return (int)e;
        }
        attr[dev] = true;
    }
    TileSched ts;
    ts.n_tiles = (N + BN - 1) / BN;
    ts.m_tiles = CONV ? m_tiles_conv : (M + G_BM - 1) / G_BM;
    if (PAIR) {
        if (ts.m_tiles % 2) {
            set_error("gemm: CTA pairs need an even number of M tiles");
            return COMA_E_BADARG;
        }
        ts.m_tiles /= 2;   // pairs of adjacent M tiles
    }
    const int num_k = (K + G_BK - 1) / G_BK;
    ts.kb_per = (num_k + ksplit - 1) / ksplit;
    ts.ksplit = (num_k + ts.kb_per - 1) / ts.kb_per;  // every slice owns at least one K-slab
    ts.split_stride = split_stride;
    const long long total = (long long)ts.n_tiles * ts.m_tiles * nbatch * ts.ksplit;
    if (total >= (1LL << 31)) {
        set_error("too many output tiles");
        return COMA_E_BADARG;
    }
    ts.total = (int)total;
    if (WS) {
        // one N tile per CTA: gridDim = n_tiles x (CTAs per N tile); the kernel walks ts.total = M tiles per N tile
        if (nbatch != 1 || ts.ksplit != 1 || num_k > WS || ts.n_tiles > kNumSM) {
            set_error("gemm: weight-stationary form needs one batch, no K split, K <= %d and at most %d N tiles", 64 * WS, kNumSM);
            return COMA_E_BADARG;
        }
        int per = kNumSM / ts.n_tiles;
        if (per > ts.m_tiles) per = ts.m_tiles;
        ts.total = ts.m_tiles;
        launch_pdl(gemm_f16_tn_kernel<BN, CONV, STAGES, GEGLU, PAIR, WS>, dim3((unsigned)(per * ts.n_tiles)), dim3(gemm_threads(BN)), smem, st, ta, tb, to, tr, M, N, K, ep, cg, ts);
    } else if (!PAIR) {
        const int slots = kNumSM * (BN <= 128 ? 2 : 1);
        const unsigned grid = (unsigned)(ts.total < slots ? ts.total : slots);
        launch_pdl(gemm_f16_tn_kernel<BN, CONV, STAGES, GEGLU, PAIR, WS>, dim3(grid), dim3(gemm_threads(BN)), smem, st, ta, tb, to, tr, M, N, K, ep, cg, ts);
    } else {
        // clusters of two CTAs (one TPC each); programmatic dependent launch as everywhere else
        const unsigned pairs = (unsigned)(ts.total < kNumSM / 2 ? ts.total : kNumSM / 2);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * pairs);
        cfg.blockDim = dim3(gemm_threads(BN));
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
        cudaLaunchKernelEx(&cfg, gemm_f16_tn_kernel<BN, CONV, STAGES, GEGLU, PAIR, WS>, ta, tb, to, tr, M, N, K, ep, cg, ts);
    }
    return check_launch("gemm_f16_tn_kernel");
}

// CTA pairs: 160- / 256-wide tiles with an even number of M tiles, a deep K and no K split. Measured on B200 (tools/gemm_bench.py,
// interleaved A/B, forced on against forced off): 8192^3 894 -> 786 us (1230 -> 1399 TFLOP/s); K >= 1280 plain GEMMs at N = 320 / 640 /
// 1280 / 10240 +10..13 %; K = 960 +3 %; K = 640 -1..6 %, K = 320 -10..12 % (a short main loop is epilogue-paced, and the pair's MMA
// stream waits for the slower of two epilogues); split-K items (M = 512 convs) -17 %. COMA_GEMM_PAIR=0 / 1 forces (1: whenever the
// tile shape allows).
static bool gemm_use_pair(int bn, int64_t m_tiles, int64_t K, int ksplit) {
    static const int env = getenv("COMA_GEMM_PAIR") ? atoi(getenv("COMA_GEMM_PAIR")) : -1;
    if (env == 0 || (bn != 256 && bn != 160) || m_tiles % 2) return false;
    return env == 1 || (K >= 16 * G_BK && ksplit == 1);
}

// Weight-stationary form: plain GEMMs with K <= 320, at least two 160-wide... see the kernel comment. COMA_GEMM_WS=0 disables.
static bool gemm_use_ws(int64_t M, int64_t N, int64_t K, int64_t nbatch) {
    static const int env = getenv("COMA_GEMM_WS") ? atoi(getenv("COMA_GEMM_WS")) : -1;
    if (env == 0 || nbatch != 1 || K > 5 * G_BK || N < 160 || (N % 160 != 0 && N % 160 < 96)) return false;
    const int64_t n_tiles = (N + 159) / 160, m_tiles = (M + G_BM - 1) / G_BM;
    if (n_tiles > kNumSM) return false;
    const int64_t per = kNumSM / n_tiles;
    return m_tiles >= 2 * per;   // every CTA amortises its W tile over at least two M tiles
}

// ---- tile width and split-K factor from a small cost model (microseconds, calibrated on B200 with tools/gemm_bench.py) --
// A work item = one (128 x BN tile, K slice). The SM count quantises the schedule: 80 tiles of a deep-K problem leave 68 SMs
// idle, which a narrower tile or a K split repairs. Split slices write fp32 slabs that a finishing pass reduces.
struct GemmPlan {
    int bn, ksplit;
};

static GemmPlan plan_gemm(int64_t m_tiles, int64_t N, int64_t K, int64_t nbatch, int64_t M, bool can_split, int64_t ws_elems) {
    const int cands[4] = {256, 160, 128, 64};
    const double t_kb[4] = {0.41, 0.27, 0.22, 0.14};   // one 64-deep K slab of a 128 x BN tile on one SM
    const double t_epi[4] = {0.9, 0.6, 0.5, 0.35};     // epilogue of one tile (hidden behind the next tile's main loop)
    const double t_fixed = 3.0, t_finish = 2.5;
    const int64_t num_k = (K + G_BK - 1) / G_BK;
    GemmPlan best = {128, 1};
    double best_t = 1e30;
    for (int c = 0; c < 4; ++c) {
        const int bn = cands[c];
        if (bn > 64 && N <= bn / 2 && N <= 128) continue;  // do not pad a narrow problem into a wide tile
        const int64_t tiles = m_tiles * ((N + bn - 1) / bn) * nbatch;
        const int max_split = (can_split && bn >= 128) ? 16 : 1;
        for (int ks = 1; ks <= max_split; ++ks) {
            const int64_t kb = (num_k + ks - 1) / ks;
            if (ks > 1 && (kb < 8 || (int64_t)ks * M * N > ws_elems)) break;
            const int64_t items = tiles * ((num_k + kb - 1) / kb);
            const int64_t waves = (items + kNumSM - 1) / kNumSM;
            const double body = (double)kb * t_kb[c];
            double t = t_fixed + (double)waves * (body > t_epi[c] ? body : t_epi[c]) + t_epi[c];
            if (ks > 1) t += t_finish + 8.0 * (double)ks * (double)M * (double)N / 5.0e6;  // slab write + read at ~5 TB/s
            if (t < best_t - 1e-9) {
                best_t = t;
                best = {bn, ks};
            }
        }
    }
    return best;
}

// Split-K finishing pass: out = act(alpha * sum_ks slab[ks] + bias + bias_rows[row / rows_per_bias] + residual), slabs summed
// in index order (bit-reproducible). One thread per 8 consecutive columns.
__global__ void splitk_finish_kernel(const float *__restrict__ ws, int ksplit, long long slab, long long M, int N, float alpha,
                                     const float *__restrict__ bias, const float *__restrict__ bias_rows, int rows_per_bias,
                                     long long bias_rows_ld, const __half *__restrict__ residual, int act, __half *__restrict__ out16,
                                     float *__restrict__ out32, int ldo) {
    pdl_trigger();
    pdl_wait();
    const int n8 = N / 8;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * n8) return;
    const long long row = i / n8;
    const int col = (int)(i % n8) * 8;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = 0.0f;
    const float *p = ws + row * N + col;
    for (int k = 0; k < ksplit; ++k, p += slab) {
        const float4 a = *reinterpret_cast<const float4 *>(p), b = *reinterpret_cast<const float4 *>(p + 4);
        f[0] += a.x; f[1] += a.y; f[2] += a.z; f[3] += a.w; f[4] += b.x; f[5] += b.y; f[6] += b.z; f[7] += b.w;
    }
    const float *brow = bias_rows ? bias_rows + (row / rows_per_bias) * bias_rows_ld : nullptr;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        f[j] = f[j] * alpha + (bias ? __ldg(bias + col + j) : 0.0f);
        if (brow) f[j] += __ldg(brow + col + j);
    }
    const long long off = row * ldo + col;
    if (residual) {
        const uint4 r = *reinterpret_cast<const uint4 *>(residual + off);
        const __half2 *h = reinterpret_cast<const __half2 *>(&r);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float2 x = __half22float2(h[u]);
            f[2 * u] += x.x;
            f[2 * u + 1] += x.y;
        }
    }
    if (act == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = __fdividef(f[j], 1.0f + __expf(-f[j]));
    } else if (act == 2) {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = __fdividef(f[j], 1.0f + __expf(-1.702f * f[j]));
    }
    if (out16) {
        uint4 w;
        __half2 *h = reinterpret_cast<__half2 *>(&w);
#pragma unroll
        for (int u = 0; u < 4; ++u) h[u] = __floats2half2_rn(f[2 * u], f[2 * u + 1]);
        *reinterpret_cast<uint4 *>(out16 + off) = w;
    }
    if (out32) {
        *reinterpret_cast<float4 *>(out32 + off) = make_float4(f[0], f[1], f[2], f[3]);
        *reinterpret_cast<float4 *>(out32 + off + 4) = make_float4(f[4], f[5], f[6], f[7]);
    }
}

static bool split_eligible(const GemmEpilogue &ep, int64_t N, int64_t nbatch, const float *ws) {
    return ws && nbatch == 1 && N % 8 == 0 && ep.ldo % 8 == 0 && (uintptr_t)ws % 16 == 0 && (uintptr_t)ep.residual % 16 == 0 &&
           (uintptr_t)ep.out16 % 16 == 0 && (uintptr_t)ep.out32 % 16 == 0;
}

// Rewrites the epilogue of a split launch (raw fp32 partial slabs) and returns the one the finishing pass applies.
static GemmEpilogue split_epilogue(GemmEpilogue &ep, float *ws, int64_t N) {
    const GemmEpilogue fin = ep;
    ep.bias = nullptr; ep.bias_rows = nullptr; ep.residual = nullptr; ep.out16 = nullptr; ep.out32 = ws; ep.ldo = (int)N;
    ep.o_s1 = 0; ep.o_s2 = 0; ep.alpha = 1.0f; ep.act = 0; ep.tma = 0; ep.rows_per_bias = 1; ep.stats = nullptr; ep.ln_rows = nullptr; ep.ln_c1 = nullptr; ep.ln_part_in = nullptr; ep.ln_part_out = nullptr; ep.ln_parts = 0; ep.ln_eps = 0.f;
    return fin;
}

static int launch_split_finish(const GemmEpilogue &fin, const float *ws, int ksplit, int64_t M, int64_t N, cudaStream_t st) {
    const long long n = M * (N / 8);
    launch_pdl(splitk_finish_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, ws, ksplit, (long long)(M * N), (long long)M, (int)N,
               fin.alpha, fin.bias, fin.bias_rows, fin.rows_per_bias, fin.bias_rows_ld, fin.residual, fin.act, fin.out16, fin.out32, fin.ldo);
    return check_launch("splitk_finish_kernel");
}

#define COMA_DISPATCH_BN(rc, bn, CONVF, ...)                           \
    switch (bn) {                                                      \
        case 64: rc = launch_gemm<64, CONVF>(__VA_ARGS__); break;      \
        case 128: rc = launch_gemm<128, CONVF>(__VA_ARGS__); break;    \
        case 160:                                                      \
            if (pair) rc = launch_gemm<160, CONVF, false, true>(__VA_ARGS__); \
            else rc = launch_gemm<160, CONVF>(__VA_ARGS__);            \
            break;                                                     \
        default:                                                       \
            if (pair) rc = launch_gemm<256, CONVF, false, true>(__VA_ARGS__); \
            else rc = launch_gemm<256, CONVF>(__VA_ARGS__);            \
            break;                                                     \
    }

}  // namespace coma

// Host-only query of the schedule the planner would pick (no device work): tests and tuning tools read it.
extern "C" int coma_gemm_plan(int64_t M, int64_t N, int64_t K, int64_t workspace_elems, int conv_m_tiles, int *tile_n, int *ksplit) {
    using namespace coma;
    COMA_REQUIRE(M > 0 && N > 0 && K > 0 && tile_n && ksplit, "bad arguments");
    const int64_t m_tiles = conv_m_tiles > 0 ? conv_m_tiles : (M + G_BM - 1) / G_BM;
    const GemmPlan p = plan_gemm(m_tiles, N, K, 1, M, workspace_elems > 0 && N % 8 == 0, workspace_elems);
    *tile_n = p.bn;
    *ksplit = p.ksplit;
    return 0;
}

extern "C" int coma_gemm_f16_ex(const coma_gemm_args *g, coma_stream_t stream) {
    using namespace coma;
    COMA_REQUIRE(g && g->A && g->W && (g->out_f16 || g->out_f32), "null pointer");
    const int64_t M = g->M, N = g->N, K = g->K, nb1 = g->nb1 > 0 ? g->nb1 : 1, nb2 = g->nb2 > 0 ? g->nb2 : 1;
    COMA_REQUIRE(M > 0 && N > 0 && K > 0 && M < (1LL << 31) && N < (1LL << 31) && K < (1LL << 31), "bad sizes");
    COMA_REQUIRE(g->lda >= K && g->ldw >= K && g->ldo >= (g->geglu ? N / 2 : N), "leading dimensions smaller than the row length");
    COMA_REQUIRE(g->lda % 8 == 0 && g->ldw % 8 == 0, "lda / ldw must be multiples of 8 elements (16-byte TMA strides)");
    COMA_REQUIRE((nb1 == 1 || (g->a_s1 % 8 == 0 && g->w_s1 % 8 == 0)) && (nb2 == 1 || (g->a_s2 % 8 == 0 && g->w_s2 % 8 == 0)),
                 "batch strides of A / W must be multiples of 8 elements");
    COMA_REQUIRE(((uintptr_t)g->A | (uintptr_t)g->W) % 16 == 0, "A and W must be 16-byte aligned");
    COMA_REQUIRE(g->act >= 0 && g->act <= 2, "act must be 0 (identity), 1 (SiLU) or 2 (quick-GELU)");
    COMA_REQUIRE(!g->out_f16 || (uintptr_t)g->out_f16 % 16 == 0, "out_f16 must be 16-byte aligned");
    COMA_REQUIRE(!g->out_f32 || (uintptr_t)g->out_f32 % 16 == 0, "out_f32 must be 16-byte aligned");
    COMA_REQUIRE(!g->residual || (uintptr_t)g->residual % 16 == 0, "residual must be 16-byte aligned");
    COMA_REQUIRE(!g->bias_rows || g->rows_per_bias > 0, "rows_per_bias must be positive");
    GemmEpilogue ep;
    ep.bias = g->bias;
    ep.bias_rows = g->bias_rows;
    ep.rows_per_bias = (int)(g->rows_per_bias > 0 ? g->rows_per_bias : 1);
    ep.bias_rows_ld = g->bias_rows_ld > 0 ? g->bias_rows_ld : N;
    ep.residual = (const __half *)g->residual;
    ep.out16 = (__half *)g->out_f16;
    ep.out32 = g->out_f32;
    ep.ldo = (int)g->ldo;
    ep.o_s1 = g->o_s1;
    ep.o_s2 = g->o_s2;
    ep.alpha = g->alpha;
    ep.act = g->act;
    ep.nb1 = (int)nb1;
    ep.stats = nullptr;
    ep.ln_rows = reinterpret_cast<const float2 *>(g->ln_row_stats);
    ep.ln_c1 = g->ln_c1;
    ep.ln_part_in = reinterpret_cast<const float2 *>(g->ln_partials_in);
    ep.ln_parts = (int)(K / 32);
    ep.ln_eps = g->ln_eps;
    ep.ln_part_out = reinterpret_cast<float2 *>(g->ln_partials_out);
    COMA_REQUIRE(!(g->ln_row_stats && g->ln_partials_in), "ln_row_stats and ln_partials_in are alternatives");
    COMA_REQUIRE(!g->ln_partials_in || (K % 32 == 0 && (uintptr_t)g->ln_partials_in % 8 == 0 && g->ln_eps > 0.f), "ln_partials_in: needs K % 32 == 0 and ln_eps > 0");
    COMA_REQUIRE(!g->ln_partials_out || (N % 32 == 0 && !g->geglu && nb1 * nb2 == 1 && g->out_f16 && !g->out_f32 && (uintptr_t)g->ln_partials_out % 8 == 0),
                 "ln_partials_out: needs N % 32 == 0, fp16 output, no batching / GEGLU");
    COMA_REQUIRE(!(g->ln_row_stats || g->ln_partials_in) == !g->ln_c1, "ln_row_stats / ln_partials_in and ln_c1 come together");
    if (ep.ln_part_in) ep.ln_rows = nullptr;
    COMA_REQUIRE(!(g->ln_row_stats || g->ln_partials_in) || (N % 32 == 0 && g->alpha == 1.0f && nb1 * nb2 == 1 && g->out_f16 && !g->out_f32 && (uintptr_t)g->ln_row_stats % 8 == 0 &&
                                                              (uintptr_t)g->ln_c1 % 16 == 0),
                 "folded LayerNorm: needs N % 32 == 0, alpha = 1, no batching, fp16 output, aligned vectors");
    if (g->geglu) {
        // fused GEGLU: W / bias rows interleaved in blocks of 32 (value, gate); output [M, N/2] fp16
        COMA_REQUIRE(N % 256 == 0 && g->out_f16 && !g->out_f32 && !g->residual && !g->bias_rows && g->act == 0 && nb1 * nb2 == 1,
                     "geglu: needs N % 256 == 0, fp16 output only, no residual / bias rows / activation / batching");
        COMA_REQUIRE(g->ldo >= N / 2 && g->ldo % 8 == 0 && (uintptr_t)g->bias % 16 == 0, "geglu: bad output stride or bias alignment");
        CUtensorMap ta, tb, to, tr;
        if (int e = make_map(&ta, g->A, M, K, g->lda, G_BM, 1, 0, 1, 0)) return e;
        const bool pair = gemm_use_pair(256, (M + G_BM - 1) / G_BM, K, 1);
        if (int e = make_map(&tb, g->W, N, K, g->ldw, pair ? 128 : 256, 1, 0, 1, 0)) return e;
        if (int e = setup_epilogue_maps(ep, &to, &tr, M, N / 2, 1, 1)) return e;
        COMA_REQUIRE(ep.tma == 1, "geglu: output not eligible for the TMA epilogue");
        if (pair) return launch_gemm<256, false, true, true>(ta, tb, to, tr, (int)M, (int)N, (int)K, ep, 1, (cudaStream_t)stream, 1, 0);
        return launch_gemm<256, false, true>(ta, tb, to, tr, (int)M, (int)N, (int)K, ep, 1, (cudaStream_t)stream, 1, 0);
    }
    const bool can_split = split_eligible(ep, N, nb1 * nb2, g->workspace) && !ep.ln_c1 && !ep.ln_part_out;   // the finishing pass does not know the folded LayerNorm
    const GemmPlan plan = plan_gemm((M + G_BM - 1) / G_BM, N, K, nb1 * nb2, M, can_split, g->workspace_elems);
    const bool ws = gemm_use_ws(M, N, K, nb1 * nb2);
    const int bn = ws ? 160 : plan.bn;
    const bool pair = !ws && gemm_use_pair(bn, (M + G_BM - 1) / G_BM, K, plan.ksplit);
    CUtensorMap ta, tb, to, tr;
    if (int e = make_map(&ta, g->A, M, K, g->lda, G_BM, nb1, g->a_s1, nb2, g->a_s2)) return e;
    if (int e = make_map(&tb, g->W, N, K, g->ldw, pair ? bn / 2 : bn, nb1, g->w_s1, nb2, g->w_s2)) return e;
    GemmEpilogue fin = ep;
    if (plan.ksplit > 1 && !ws) fin = split_epilogue(ep, g->workspace, N);
    if (int e = setup_epilogue_maps(ep, &to, &tr, M, N, nb1, nb2)) return e;
    COMA_REQUIRE(!(ep.ln_c1 || ep.ln_part_out) || ep.tma == 1, "folded LayerNorm: output not eligible for the TMA epilogue");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = 0;
    if (ws) return launch_gemm<160, false, false, false, 5>(ta, tb, to, tr, (int)M, (int)N, (int)K, ep, 1, st, 1, (long long)(M * N));
    COMA_DISPATCH_BN(rc, bn, false, ta, tb, to, tr, (int)M, (int)N, (int)K, ep, (int)(nb1 * nb2), st, plan.ksplit, (long long)(M * N))
    if (rc == 0 && plan.ksplit > 1) rc = launch_split_finish(fin, g->workspace, plan.ksplit, M, N, st);
    return rc;
}

namespace coma {
// NHWC tensor map {C, W, H, B} with box {64, TW, TH, TB}
static int make_conv_map(CUtensorMap *m, const void *ptr, int64_t B, int64_t H, int64_t W, int64_t C, int64_t ldx, int TW, int TH,
                         int TB, int stride) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return COMA_E_NODEVICE;
    }
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)ldx * 2, (cuuint64_t)(W * ldx) * 2, (cuuint64_t)(H * W * ldx) * 2};
    // strided convolution: the box TRAVERSES stride*T pixels with element stride `stride`, i.e. lands T pixels in shared memory
    cuuint32_t box[4] = {(cuuint32_t)G_BK, (cuuint32_t)(TW * stride), (cuuint32_t)(TH * stride), (cuuint32_t)TB};
    cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void *>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (conv) failed with CUresult %d", (int)r);
        return COMA_E_BADARG;
    }
    return 0;
}
}  // namespace coma

extern "C" int coma_conv3x3_f16(const void *x, int64_t B, int64_t H, int64_t W, int64_t C, int64_t ldx, const void *Wt, int64_t ldw,
                                int64_t N, const float *bias, const float *bias_rows, int64_t bias_rows_ld, const void *residual, int act,
                                void *out_f16, float *out_f32, int64_t ldo, coma_stream_t stream) {
    return coma_conv3x3_f16_ws(x, B, H, W, C, ldx, Wt, ldw, N, bias, bias_rows, bias_rows_ld, residual, act, out_f16, out_f32, ldo, nullptr, 0,
                               stream);
}

extern "C" int coma_conv3x3_f16_ws(const void *x, int64_t B, int64_t H, int64_t W, int64_t C, int64_t ldx, const void *Wt, int64_t ldw,
                                   int64_t N, const float *bias, const float *bias_rows, int64_t bias_rows_ld, const void *residual, int act,
                                   void *out_f16, float *out_f32, int64_t ldo, float *workspace, int64_t workspace_elems,
                                   coma_stream_t stream) {
    return coma_conv3x3_strided_f16(x, B, H, W, C, ldx, 1, 1, Wt, ldw, N, bias, bias_rows, bias_rows_ld, residual, act, out_f16, out_f32, ldo,
                                    workspace, workspace_elems, nullptr, nullptr, stream);
}

extern "C" int coma_conv3x3_strided_f16(const void *x, int64_t B, int64_t Hin, int64_t Win, int64_t C, int64_t ldx, int stride, int pad,
                                        const void *Wt, int64_t ldw, int64_t N, const float *bias, const float *bias_rows,
                                        int64_t bias_rows_ld, const void *residual, int act, void *out_f16, float *out_f32, int64_t ldo,
                                        float *workspace, int64_t workspace_elems, float *stats, int *stats_written,
                                        coma_stream_t stream) {
    using namespace coma;
    COMA_REQUIRE(x && Wt && (out_f16 || out_f32), "null pointer");
    COMA_REQUIRE(B > 0 && Hin > 0 && Win > 0 && C > 0 && N > 0, "bad sizes");
    COMA_REQUIRE((stride == 1 && pad == 1) || (stride == 2 && (pad == 0 || pad == 1)), "stride 1 / pad 1, or stride 2 with pad 1 (both sides) or pad 0 (one zero row / column after the image)");
    // output extent: stride 1: same; stride 2, pad 1: floor((n - 1) / 2) + 1; stride 2, pad 0 with (0, 1) padding: floor((n - 2) / 2) + 1
    const int64_t H = stride == 1 ? Hin : (pad ? (Hin - 1) / 2 + 1 : (Hin - 2) / 2 + 1);
    const int64_t W = stride == 1 ? Win : (pad ? (Win - 1) / 2 + 1 : (Win - 2) / 2 + 1);
    COMA_REQUIRE(H > 0 && W > 0, "image too small");
    COMA_REQUIRE(C % 64 == 0 && ldx % 8 == 0 && ldx >= C, "implicit-GEMM conv needs C % 64 == 0 (use im2col otherwise)");
    COMA_REQUIRE(ldw >= 9 * C && ldw % 8 == 0 && ldo >= N, "bad leading dimensions");
    COMA_REQUIRE(((uintptr_t)x | (uintptr_t)Wt) % 16 == 0, "x and W must be 16-byte aligned");
    COMA_REQUIRE(act == 0 || act == 1, "act must be 0 or 1");
    // tile geometry: 128 output pixels = TB images x TH rows x TW columns
    int TW = (int)(W < 128 ? W : 128), TH = (int)(128 / TW < H ? 128 / TW : H), TB = 128 / (TW * TH);
    COMA_REQUIRE(TW * TH * TB == 128 && W % TW == 0 && H % TH == 0, "image extent does not tile into 128-pixel blocks (use im2col)");
    ConvGeom cg;
    cg.H = (int)H; cg.W = (int)W; cg.B = (int)B; cg.TW = TW; cg.TH = TH; cg.TB = TB;
    cg.tiles_x = (int)(W / TW); cg.tiles_y = (int)(H / TH); cg.cblocks = (int)(C / 64); cg.stride = stride; cg.pad = pad;
    const int64_t m_tiles = TB > 1 ? (B + TB - 1) / TB : B * cg.tiles_x * cg.tiles_y;
    const int64_t M = B * H * W, K = 9 * C;
    GemmEpilogue ep;
    ep.bias = bias; ep.bias_rows = bias_rows; ep.rows_per_bias = (int)(H * W); ep.bias_rows_ld = bias_rows_ld > 0 ? bias_rows_ld : N; ep.residual = (const __half *)residual;
    ep.out16 = (__half *)out_f16; ep.out32 = out_f32; ep.ldo = (int)ldo; ep.o_s1 = 0; ep.o_s2 = 0; ep.alpha = 1.0f; ep.act = act;
    ep.nb1 = 1;
    ep.stats = nullptr;
    ep.ln_rows = nullptr;
    ep.ln_c1 = nullptr;
    ep.ln_part_in = nullptr;
    ep.ln_part_out = nullptr;
    ep.ln_parts = 0;
    ep.ln_eps = 0.f;
    // split-K needs every slab row < M to be written: tiles never straddle M except with an odd image count at TB > 1
    const bool can_split = split_eligible(ep, N, 1, workspace) && (TB == 1 || B % TB == 0);
    const GemmPlan plan = plan_gemm(m_tiles, N, K, 1, M, can_split, workspace_elems);
    const int bn = plan.bn;
    const bool pair = gemm_use_pair(bn, m_tiles, K, plan.ksplit);
    CUtensorMap ta, tb, to, tr;
    COMA_REQUIRE(stride == 1 || TB == 1, "strided implicit conv: the output must have at least 128 pixels per image");
    if (int e = make_conv_map(&ta, x, B, Hin, Win, C, ldx, TW, TH, TB, stride)) return e;
    if (int e = make_map(&tb, Wt, N, K, ldw, pair ? bn / 2 : bn, 1, 0, 1, 0)) return e;
    GemmEpilogue fin = ep;
    if (plan.ksplit > 1) fin = split_epilogue(ep, workspace, N);
    if (int e = setup_epilogue_maps(ep, &to, &tr, M, N, 1, 1)) return e;
    // fused GroupNorm statistics: only from the TMA epilogue of an unsplit launch, whole 32-row blocks, 16-byte aligned rows
    const bool do_stats = stats && ep.tma && plan.ksplit == 1 && M % 32 == 0 && N % 2 == 0 && (uintptr_t)stats % 16 == 0;
    ep.stats = do_stats ? stats : nullptr;
    if (stats_written) *stats_written = do_stats ? 1 : 0;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = 0;
    COMA_DISPATCH_BN(rc, bn, true, ta, tb, to, tr, (int)M, (int)N, (int)K, ep, 1, st, plan.ksplit, (long long)(M * N), cg, (int)m_tiles)
    if (rc == 0 && plan.ksplit > 1) rc = launch_split_finish(fin, workspace, plan.ksplit, M, N, st);
    return rc;
}

extern "C" int coma_gemm_f16_tn(const void *A, int64_t lda, const void *W, int64_t ldw, int64_t M, int64_t N, int64_t K,
                                const float *bias, const void *residual, int act, void *out_f16, float *out_f32, int64_t ldo,
                                coma_stream_t stream) {
    coma_gemm_args g = {};
    g.A = A; g.lda = lda; g.W = W; g.ldw = ldw; g.M = M; g.N = N; g.K = K; g.nb1 = 1; g.nb2 = 1;
    g.bias = bias; g.residual = residual; g.act = act; g.out_f16 = out_f16; g.out_f32 = out_f32; g.ldo = ldo; g.alpha = 1.0f;
    return coma_gemm_f16_ex(&g, stream);
}
