// K2 — pair distance -> contact count + proximity expectation (reference: utils/coma.py:284-291, :116-119).
//
// Layout: one CTA owns a 16(h) x 128(o) tile of the [H,O] accumulators, each thread 8 pairs (8 h-rows x 1 o-column)
// held in registers across ALL samples of the call; vertices of the current sample chunk are staged in shared memory
// (object vertices as SoA so the per-lane reads are conflict-free, human vertices as float4 so one broadcast LDS.128
// serves the warp).  The accumulators are read once at kernel start and written once at the end, coalesced along o.
//   bytes per launch  = 12*S*(H+O) (vertices) + 2 accumulators * (4 R + 4 W) * H*O  = 16 B / vertex-pair at S = 1
//   => HBM-bound at S = 1 (the reference's per-sample streaming form), ALU/SFU-bound once S >~ 8.
// Bit-exactness of `count`: the squared distance is built from explicitly rounded __fsub_rn/__fmul_rn/__fadd_rn in the
// reference's order ((x+y)+z). torch then takes the correctly rounded fp32 sqrt and compares with fp32(thres); because
// IEEE sqrt is monotone, {x : sqrt_rn(x) < thres} == {x : x < T} with T the smallest float whose rounded root is >=
// thres (found on the host), so the verdict `sq < T` is identical and the hot loop needs no IEEE square root.
// The proximity term exp(-d/size) only has to hold 1e-4: d = sq*rsqrt(sq) (MUFU) and exp via one MUFU.EX2.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace coma {

// Squared distance in the association of the reference's `torch.sum(torch.square(h - o), dim=-1)`:
//   ORD = 0  ((x^2 + y^2) + z^2)  ATen's CPU reduction (and numpy);
//   ORD = 1  ((x^2 + z^2) + y^2)  ATen's CUDA reduction over a contiguous last dimension of 3 — measured on B200 with torch 2.11
//            for fp32 and fp64 and every shape tried (tools/probe_torch_cuda_semantics.py, profiles/r02_torch_cuda_semantics.json).
// The reference's production runs are device="cuda" (src/coma/extract_coma.py:329), so ORD = 1 is what `ComA` uses by default;
// ORD = 0 reproduces the CPU reference (the committed CPU-generated golden vectors) bit for bit.
template <int ORD>
__device__ __forceinline__ float sq_dist(float dx, float dy, float dz) {
    const float xx = __fmul_rn(dx, dx), yy = __fmul_rn(dy, dy), zz = __fmul_rn(dz, dz);
    return ORD == 0 ? __fadd_rn(__fadd_rn(xx, yy), zz) : __fadd_rn(__fadd_rn(xx, zz), yy);
}

constexpr int K2_TO = 128;  // object columns per CTA (= blockDim.x)
constexpr int K2_TY = 2;    // blockDim.y
constexpr int K2_RH = 8;    // human rows per thread
constexpr int K2_TH = K2_TY * K2_RH;
constexpr int K2_CS = 16;   // samples staged per chunk

template <int ORD>
__global__ void __launch_bounds__(K2_TO *K2_TY, 4)
    pair_accumulate_kernel(const float *__restrict__ hv, const float *__restrict__ ov, int S, int H, int O, float sq_thres,
                           float neg_log2e_over_size, float *__restrict__ count, float *__restrict__ nom) {
    __shared__ float4 sh[K2_CS][K2_TH];
    __shared__ float so[K2_CS][3][K2_TO];

    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * K2_TO + tx;
    const int o0 = blockIdx.x * K2_TO, h0 = blockIdx.y * K2_TH;
    const int o = o0 + tx;

    // The running accumulators are fetched FIRST (their HBM latency hides behind the vertex staging and the first
    // distances) and every sample is then added on top in sample order, exactly like the reference's in-place `+=`.
    float cnt[K2_RH], acc[K2_RH];
#pragma unroll
    for (int r = 0; r < K2_RH; ++r) {
        const int h = h0 + ty * K2_RH + r;
        const bool ok = (h < H) && (o < O);
        const size_t q = (size_t)h * O + o;
        cnt[r] = ok ? __ldcs(count + q) : 0.0f;
        acc[r] = ok ? __ldcs(nom + q) : 0.0f;
    }

    for (int s0 = 0; s0 < S; s0 += K2_CS) {
        const int ns = min(K2_CS, S - s0);
        // stage human rows: ns x 16 vertices
        for (int i = tid; i < ns * K2_TH; i += K2_TO * K2_TY) {
            int cs = i / K2_TH, r = i % K2_TH, h = h0 + r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (h < H) {
                const float *p = hv + ((size_t)(s0 + cs) * H + h) * 3;
                v.x = p[0]; v.y = p[1]; v.z = p[2];
            }
            sh[cs][r] = v;
        }
        // stage object columns: ns x 128 vertices, coalesced over the 384 contiguous floats of a sample's tile
        for (int i = tid; i < ns * K2_TO * 3; i += K2_TO * K2_TY) {
            int cs = i / (K2_TO * 3), e = i % (K2_TO * 3);
            int oo = e / 3, k = e % 3;
            float v = 0.f;
            if (o0 + oo < O) v = ov[((size_t)(s0 + cs) * O + o0) * 3 + e];
            so[cs][k][oo] = v;
        }
        __syncthreads();
        for (int cs = 0; cs < ns; ++cs) {
            const float ox = so[cs][0][tx], oy = so[cs][1][tx], oz = so[cs][2][tx];
#pragma unroll
            for (int r = 0; r < K2_RH; ++r) {
                const float4 hvv = sh[cs][ty * K2_RH + r];
                const float dx = __fsub_rn(hvv.x, ox), dy = __fsub_rn(hvv.y, oy), dz = __fsub_rn(hvv.z, oz);
                const float sq = sq_dist<ORD>(dx, dy, dz);
                cnt[r] += (sq < sq_thres) ? 1.0f : 0.0f;
                float d, e;
                asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(sq));
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(d * neg_log2e_over_size));
                acc[r] += e;
            }
        }
        __syncthreads();
    }
    if (o < O) {
#pragma unroll
        for (int r = 0; r < K2_RH; ++r) {
            const int h = h0 + ty * K2_RH + r;
            if (h < H) {
                const size_t q = (size_t)h * O + o;
                __stcs(count + q, cnt[r]);
                __stcs(nom + q, acc[r]);
            }
        }
    }
}

// ---- 16-byte variant: each thread owns 4 consecutive object columns x 4 human rows (needs O % 4 == 0 and 16-byte aligned
// accumulators). Same arithmetic; the accumulator traffic moves as LDG.128 / STG.128 with streaming cache hints, which
// is what lets the S = 1 form run close to the HBM copy rate.
constexpr int K2V_TX = 128;            // threads along o, 4 columns each -> 512 columns per CTA
constexpr int K2V_TY = 2;
constexpr int K2V_RH = 4;              // human rows per thread
constexpr int K2V_TH = K2V_TY * K2V_RH;
constexpr int K2V_TO = K2V_TX * 4;
constexpr int K2V_CS = 4;              // samples staged per chunk

template <int ORD>
__device__ __forceinline__ void pair_update(float hx, float hy, float hz, float ox, float oy, float oz, float sq_thres,
                                            float nl2e, float &cnt, float &acc) {
    const float dx = __fsub_rn(hx, ox), dy = __fsub_rn(hy, oy), dz = __fsub_rn(hz, oz);
    const float sq = sq_dist<ORD>(dx, dy, dz);
    cnt += (sq < sq_thres) ? 1.0f : 0.0f;
    float d, e;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(sq));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(d * nl2e));
    acc += e;
}

template <int ORD>
__global__ void __launch_bounds__(K2V_TX *K2V_TY, 4)
    pair_accumulate_vec4_kernel(const float *__restrict__ hv, const float *__restrict__ ov, int S, int H, int O, float sq_thres,
                                float nl2e, float *__restrict__ count, float *__restrict__ nom) {
    __shared__ float4 sh[K2V_CS][K2V_TH];
    __shared__ __align__(16) float so[K2V_CS][3][K2V_TO];
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * K2V_TX + tx;
    const int o0 = blockIdx.x * K2V_TO, h0 = blockIdx.y * K2V_TH;
    const int o = o0 + 4 * tx;
    const bool col_ok = o < O;  // O % 4 == 0: a float4 is either fully inside or fully outside

    float4 cnt[K2V_RH], acc[K2V_RH];
#pragma unroll
    for (int r = 0; r < K2V_RH; ++r) {
        const int h = h0 + ty * K2V_RH + r;
        const bool ok = col_ok && h < H;
        const size_t q = (size_t)h * O + o;
        cnt[r] = ok ? __ldcs(reinterpret_cast<const float4 *>(count + q)) : make_float4(0.f, 0.f, 0.f, 0.f);
        acc[r] = ok ? __ldcs(reinterpret_cast<const float4 *>(nom + q)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int s0 = 0; s0 < S; s0 += K2V_CS) {
        const int ns = min(K2V_CS, S - s0);
        for (int i = tid; i < ns * K2V_TH; i += K2V_TX * K2V_TY) {
            const int cs = i / K2V_TH, r = i % K2V_TH, h = h0 + r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (h < H) {
                const float *p = hv + ((size_t)(s0 + cs) * H + h) * 3;
                v.x = p[0]; v.y = p[1]; v.z = p[2];
            }
            sh[cs][r] = v;
        }
        for (int i = tid; i < ns * K2V_TO * 3; i += K2V_TX * K2V_TY) {
            const int cs = i / (K2V_TO * 3), e = i % (K2V_TO * 3);
            const int oo = e / 3, k = e % 3;
            float v = 0.f;
            if (o0 + oo < O) v = ov[((size_t)(s0 + cs) * O + o0) * 3 + e];
            so[cs][k][oo] = v;
        }
        __syncthreads();
        for (int cs = 0; cs < ns; ++cs) {
            const float4 ox = *reinterpret_cast<const float4 *>(&so[cs][0][4 * tx]);
            const float4 oy = *reinterpret_cast<const float4 *>(&so[cs][1][4 * tx]);
            const float4 oz = *reinterpret_cast<const float4 *>(&so[cs][2][4 * tx]);
#pragma unroll
            for (int r = 0; r < K2V_RH; ++r) {
                const float4 hvv = sh[cs][ty * K2V_RH + r];
                pair_update<ORD>(hvv.x, hvv.y, hvv.z, ox.x, oy.x, oz.x, sq_thres, nl2e, cnt[r].x, acc[r].x);
                pair_update<ORD>(hvv.x, hvv.y, hvv.z, ox.y, oy.y, oz.y, sq_thres, nl2e, cnt[r].y, acc[r].y);
                pair_update<ORD>(hvv.x, hvv.y, hvv.z, ox.z, oy.z, oz.z, sq_thres, nl2e, cnt[r].z, acc[r].z);
                pair_update<ORD>(hvv.x, hvv.y, hvv.z, ox.w, oy.w, oz.w, sq_thres, nl2e, cnt[r].w, acc[r].w);
            }
        }
        __syncthreads();
    }
    if (col_ok) {
#pragma unroll
        for (int r = 0; r < K2V_RH; ++r) {
            const int h = h0 + ty * K2V_RH + r;
            if (h < H) {
                const size_t q = (size_t)h * O + o;
                __stcs(reinterpret_cast<float4 *>(count + q), cnt[r]);
                __stcs(reinterpret_cast<float4 *>(nom + q), acc[r]);
            }
        }
    }
}

// ---- streaming variant for S <= 4 samples per launch (the reference's one-sample-per-call form): no shared memory, no
// barriers. A thread owns 4 object columns x 4 human rows; per sample it reads its 4 object vertices as three float4 (48
// contiguous bytes) and the 4 human vertices as warp-uniform broadcast loads, everything else is the accumulator stream
// (LDG.128 / STG.128, evict-first). With nothing to synchronise on, CTAs are pure load -> math -> store pipelines and the
// SM keeps >100 KB of accumulator traffic in flight.
template <int RH, int MINB, int ORD>
__global__ void __launch_bounds__(256, MINB)
    pair_accumulate_stream_kernel(const float *__restrict__ hv, const float *__restrict__ ov, int S, int H, int O, float sq_thres,
                                  float nl2e, float *__restrict__ count, float *__restrict__ nom) {
    constexpr int K2V_RH = RH;
    const int tx = threadIdx.x & 127, ty = threadIdx.x >> 7;
    const int o = blockIdx.x * K2V_TO + 4 * tx, h0 = (blockIdx.y * 2 + ty) * RH;
    if (o >= O) return;
    float4 cnt[K2V_RH], acc[K2V_RH];
#pragma unroll
    for (int r = 0; r < K2V_RH; ++r) {
        const bool ok = h0 + r < H;
        const size_t q = (size_t)(h0 + r) * O + o;
        cnt[r] = ok ? __ldcs(reinterpret_cast<const float4 *>(count + q)) : make_float4(0.f, 0.f, 0.f, 0.f);
        acc[r] = ok ? __ldcs(reinterpret_cast<const float4 *>(nom + q)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int s = 0; s < S; ++s) {
        const float4 *op = reinterpret_cast<const float4 *>(ov + ((size_t)s * O + o) * 3);  // 12 floats, 16-byte aligned (o % 4 == 0)
        const float4 a = __ldg(op), b = __ldg(op + 1), c = __ldg(op + 2);  // x0 y0 z0 x1 | y1 z1 x2 y2 | z2 x3 y3 z3
#pragma unroll
        for (int r = 0; r < K2V_RH; ++r) {
            const int h = min(h0 + r, H - 1);
            const float *hp = hv + ((size_t)s * H + h) * 3;
            const float hx = __ldg(hp), hy = __ldg(hp + 1), hz = __ldg(hp + 2);
            pair_update<ORD>(hx, hy, hz, a.x, a.y, a.z, sq_thres, nl2e, cnt[r].x, acc[r].x);
            pair_update<ORD>(hx, hy, hz, a.w, b.x, b.y, sq_thres, nl2e, cnt[r].y, acc[r].y);
            pair_update<ORD>(hx, hy, hz, b.z, b.w, c.x, sq_thres, nl2e, cnt[r].z, acc[r].z);
            pair_update<ORD>(hx, hy, hz, c.y, c.z, c.w, sq_thres, nl2e, cnt[r].w, acc[r].w);
        }
    }
#pragma unroll
    for (int r = 0; r < K2V_RH; ++r) {
        if (h0 + r < H) {
            const size_t q = (size_t)(h0 + r) * O + o;
            __stcs(reinterpret_cast<float4 *>(count + q), cnt[r]);
            __stcs(reinterpret_cast<float4 *>(nom + q), acc[r]);
        }
    }
}

// Smallest float T with sqrtf_rn(T) >= thres, so that  sqrtf_rn(x) < thres  <=>  x < T  for every x >= 0.
static float squared_threshold_f32(float thres) {
    if (!(thres > 0.0f)) return 0.0f;  // d < thres never holds for thres <= 0 or NaN (d >= 0)
    if (isinf(thres)) return INFINITY;
    volatile float t = thres * thres;
    if (isinf(t)) t = 3.402823466e+38f;
    while (t > 0.0f && sqrtf(t) >= thres) t = nextafterf(t, 0.0f);
    while (sqrtf(t) < thres) t = nextafterf(t, INFINITY);
    return t;
}

}  // namespace coma

extern "C" int coma_pair_accumulate_order_f32(const float *hv, const float *ov, int64_t S, int64_t H, int64_t O, float thres,
                                              float grid_size, int sum_order, float *count, float *nom, coma_stream_t stream) {
    using namespace coma;
    COMA_REQUIRE(hv && ov && count && nom, "null pointer");
    COMA_REQUIRE(S >= 0 && H > 0 && O > 0, "bad sizes");
    COMA_REQUIRE(H * O < (int64_t)1 << 40 && S < (int64_t)1 << 30, "sizes out of range");
    COMA_REQUIRE(sum_order == COMA_SUM_ORDER_TORCH_CPU || sum_order == COMA_SUM_ORDER_TORCH_CUDA, "sum_order must be 0 (torch CPU) or 1 (torch CUDA)");
    if (S == 0) return 0;
    dim3 block(K2_TO, K2_TY);
    dim3 grid((unsigned)((O + K2_TO - 1) / K2_TO), (unsigned)((H + K2_TH - 1) / K2_TH));
    COMA_REQUIRE(grid.y <= 65535u, "H too large for one launch (max 1048560)");
    COMA_REQUIRE(grid_size > 0.0f, "spatial_grid_size must be positive");
    const float nl2e = (float)(-1.4426950408889634 / (double)grid_size);
    const float sq_thres = squared_threshold_f32(thres);
    const bool vec4 = (O % 4 == 0) && (((uintptr_t)count | (uintptr_t)nom) % 16 == 0);
    const bool cu = sum_order == COMA_SUM_ORDER_TORCH_CUDA;
    cudaStream_t st = (cudaStream_t)stream;
    if (vec4 && S <= 4 && ((uintptr_t)ov % 16 == 0)) {
        // measured (tools/k2_stream_bench.py): 4 rows/thread at 3 CTAs/SM (80 registers, no spills) 5.79-5.97 TB/s;
        // 4 CTAs/SM (64 registers, spills) 5.30; 2 rows/thread at 6 CTAs/SM 5.28.
        dim3 vgrid((unsigned)((O + K2V_TO - 1) / K2V_TO), (unsigned)((H + 2 * 4 - 1) / (2 * 4)));
        COMA_REQUIRE(vgrid.y <= 65535u, "H too large for one launch");
        if (cu) pair_accumulate_stream_kernel<4, 3, 1><<<vgrid, 256, 0, st>>>(hv, ov, (int)S, (int)H, (int)O, sq_thres, nl2e, count, nom);
        else pair_accumulate_stream_kernel<4, 3, 0><<<vgrid, 256, 0, st>>>(hv, ov, (int)S, (int)H, (int)O, sq_thres, nl2e, count, nom);
        return check_launch("pair_accumulate_stream_kernel");
    }
    if (vec4) {
        dim3 vblock(K2V_TX, K2V_TY);
        dim3 vgrid((unsigned)((O + K2V_TO - 1) / K2V_TO), (unsigned)((H + K2V_TH - 1) / K2V_TH));
        COMA_REQUIRE(vgrid.y <= 65535u, "H too large for one launch (max 524280)");
        if (cu) pair_accumulate_vec4_kernel<1><<<vgrid, vblock, 0, st>>>(hv, ov, (int)S, (int)H, (int)O, sq_thres, nl2e, count, nom);
        else pair_accumulate_vec4_kernel<0><<<vgrid, vblock, 0, st>>>(hv, ov, (int)S, (int)H, (int)O, sq_thres, nl2e, count, nom);
        return check_launch("pair_accumulate_vec4_kernel");
    }
    if (cu) pair_accumulate_kernel<1><<<grid, block, 0, st>>>(hv, ov, (int)S, (int)H, (int)O, sq_thres, nl2e, count, nom);
    else pair_accumulate_kernel<0><<<grid, block, 0, st>>>(hv, ov, (int)S, (int)H, (int)O, sq_thres, nl2e, count, nom);
    return check_launch("pair_accumulate_kernel");
}

extern "C" int coma_pair_accumulate_f32(const float *hv, const float *ov, int64_t S, int64_t H, int64_t O, float thres,
                                        float grid_size, float *count, float *nom, coma_stream_t stream) {
    return coma_pair_accumulate_order_f32(hv, ov, S, H, O, thres, grid_size, COMA_SUM_ORDER_TORCH_CPU, count, nom, stream);
}
