// K3 — relative-orientation soft histograms (reference: utils/coma.py:123-172, :102-112, :295-323).
//
// Work decomposition.  One warp owns one (human vertex h, object vertex o) pair for ALL samples of the call; lane l
// owns the bins n = l + 32 j (j < NB = ceil(N/32)), so both histograms of the pair (2 x N fp32) live in registers and
// the [H,O,N] grids are read-modify-written exactly once per launch, fully coalesced (32 consecutive floats per j).
// Samples are processed in chunks of 32: lane l canonicalises the pair's two normals for sample s0+l (the prologue is
// thereby amortised over the 32 lanes), parks the six components in shared memory, and then every lane walks the 32
// samples reading them back as warp-broadcast LDS.128/LDS.64 and evaluates its 2 x NB bins.
//
// Per bin evaluation:  c = G[n] . cn   (3 FP32 ops)  ->  score = 2^(-(acos(c) * sqrt(log2 e)/sigma)^2)
// with a branch-free acos(|c|) = sqrt(1-|c|) * P6(|c|) (Abramowitz-Stegun 4.4.45 form, one MUFU.SQRT, |err| <= 4.3e-7
// rad; fit: tools/fit_acos.py) and one MUFU.EX2.  Two bins are evaluated per instruction with Blackwell's packed
// FP32x2 pipe (FFMA2 / FMUL2 / FADD2, sm_100+): the FMA-type work costs half the issue slots, which moves the kernel
// from issue-bound to the SFU/FMA-pipe balance point (2 MUFU and ~15 FMA-pipe cycles per bin evaluation).
// The kernel is SFU / FP32-pipe bound (2 x N = 500 evaluations per pair-sample at N = 250), not HBM bound:
//   bytes per launch = 24*S*(H+O) (normals, L2-resident) + 2 grids * (4 R + 4 W) * H*O*N.
//
// Numerics.  The canonicalisation prologue uses explicitly rounded __fmul_rn/__fadd_rn/__fdiv_rn/__fsqrt_rn in the
// reference's operation order (including its literal, non-skew "cross-product matrix" and the antipodal reflect
// branch), so the canonical normals are bit-identical to torch's fp32 result.  The reference then promotes to fp64 for
// dot/acos/exp and rounds each per-sample sum to fp32; here those run in fp32, which keeps every grid entry within
// ~4e-5 relative of the reference for any sigma (tolerance 1e-4).
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace coma {

struct Vec3 {
    float x, y, z;
};

// The reference's 3-term `torch.sum(., dim=-1)`: (x+y)+z on the CPU, (x+z)+y on CUDA (ord = 1; measured on B200, see pair.cu).
// A runtime flag: the prologue runs once per (pair, sample) against hundreds of bin evaluations, two selects per sum are free.
__device__ __forceinline__ float sum3_ref(float x, float y, float z, int ord) {
    const float m = ord ? z : y, l = ord ? y : z;
    return __fadd_rn(__fadd_rn(x, m), l);
}

// v / (sqrt(sum3(x^2, y^2, z^2)) + eps)   — utils/transformations.py:14-17
__device__ __forceinline__ Vec3 normalize_ref(Vec3 v, float eps, int ord = 0) {
    float n = __fadd_rn(__fsqrt_rn(sum3_ref(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y), __fmul_rn(v.z, v.z), ord)), eps);
    return Vec3{__fdiv_rn(v.x, n), __fdiv_rn(v.y, n), __fdiv_rn(v.z, n)};
}

__device__ __forceinline__ float dot_ref(Vec3 a, Vec3 b, int ord = 0) {
    return sum3_ref(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y), __fmul_rn(a.z, b.z), ord);
}

// Three quotients by ONE divisor with a shared reciprocal: r = rcp(d) refined by one Newton step, then per numerator q = x r,
// q += fma(-q, d, x) r (Markstein's correction). With a faithful r the result is the correctly rounded x / d except for rare
// 1-ulp cases — good for K3 (1e-4 contract), NOT for the bit-exact canonicalisation entry point. 12 instructions instead of three
// div.rn sequences (~33, each with its own MUFU.RCP, FCHK range check and slow-path branch). NaN / zero divisors give NaN like
// the IEEE division; an infinite divisor gives NaN instead of 0 (a non-finite normal: the pair is poisoned either way).
#ifndef K3C_FAST_DIV
#define K3C_FAST_DIV 1
#endif
__device__ __forceinline__ void div3_shared(float &x, float &y, float &z, float d) {
    float r;
    asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(d));
    r = fmaf(r, fmaf(-d, r, 1.0f), r);
    float q = x * r; x = fmaf(fmaf(-q, d, x), r, q);
    q = y * r;       y = fmaf(fmaf(-q, d, y), r, q);
    q = z * r;       z = fmaf(fmaf(-q, d, z), r, q);
}

// canonicalize_a_wrt_b_to_p for one (a, b) pair; a, b, p, sp already normalised — utils/coma.py:135-170
// FASTDIV (K3's cone kernel only): the two triples of divisions use div3_shared.
template <bool FASTDIV = false>
__device__ __forceinline__ Vec3 canonicalize_ref(Vec3 a, Vec3 b, Vec3 p, Vec3 sp, float eps, int ord = 0) {
    const float b_dot_p = dot_ref(b, p, ord), a_dot_b = dot_ref(a, b, ord), a_dot_p = dot_ref(a, p, ord), a_dot_sp = dot_ref(a, sp, ord);
    const float one_plus = __fadd_rn(1.0f, b_dot_p);
    const bool replace = one_plus < eps;  // :143
    const bool pos_div = one_plus > 0.0f && one_plus < INFINITY;
    // rows of the reference's "cross product matrix" (:149-155): [b0,-b2,b1], [b2,0,-b0], [-b1,0,0]
    Vec3 bxp;
    bxp.x = __fadd_rn(__fadd_rn(__fmul_rn(b.x, p.x), __fmul_rn(-b.z, p.y)), __fmul_rn(b.y, p.z));
    bxp.y = __fadd_rn(__fadd_rn(__fmul_rn(b.z, p.x), __fmul_rn(0.0f, p.y)), __fmul_rn(-b.x, p.z));
    bxp.z = __fadd_rn(__fadd_rn(__fmul_rn(-b.y, p.x), __fmul_rn(0.0f, p.y)), __fmul_rn(0.0f, p.z));
    const float a_dot_bxp = dot_ref(a, bxp, ord);  // :159
    const float av[3] = {a.x, a.y, a.z}, bv[3] = {b.x, b.y, b.z}, pv[3] = {p.x, p.y, p.z}, sv[3] = {sp.x, sp.y, sp.z};
    const float xv[3] = {bxp.x, bxp.y, bxp.z};
    float f[3];
    if (FASTDIV) {
        float v0 = __fmul_rn(xv[0], a_dot_bxp), v1 = __fmul_rn(xv[1], a_dot_bxp), v2 = __fmul_rn(xv[2], a_dot_bxp);   // :162
        div3_shared(v0, v1, v2, one_plus);                                                                         // :163
        f[0] = v0; f[1] = v1; f[2] = v2;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float v;
        if (FASTDIV) {
            v = replace ? 0.0f : f[k];
        } else {
            v = __fmul_rn(xv[k], a_dot_bxp);                   // :162
            // :163. With p = (0,0,1) the third component of bxp is an exact zero for EVERY pair, and div.rn's operand check (FCHK) sends a
            // zero numerator down its ~60-instruction slow path — the whole warp, four times per chunk. +-0 / (positive finite) is the
            // zero itself, so only non-zero numerators are divided (bit-identical; a NaN divisor still takes the generic division).
            const bool zero_num = v == 0.0f && pos_div;
            const float q = __fdiv_rn(zero_num ? 1.0f : v, one_plus);
            v = replace ? 0.0f : (zero_num ? v : q);
        }
        v = __fadd_rn(v, __fmul_rn(b_dot_p, av[k]));           // :164
        v = __fadd_rn(v, __fmul_rn(a_dot_b, pv[k]));           // :165
        v = __fsub_rn(v, __fmul_rn(a_dot_p, bv[k]));           // :166
        if (replace) v = __fsub_rn(__fmul_rn(__fmul_rn(2.0f, a_dot_sp), sv[k]), av[k]);  // :145,:169
        f[k] = v;
    }
    const float n = __fsqrt_rn(sum3_ref(__fmul_rn(f[0], f[0]), __fmul_rn(f[1], f[1]), __fmul_rn(f[2], f[2]), ord));
    if (FASTDIV) {
        div3_shared(f[0], f[1], f[2], n);                      // :170
        return Vec3{f[0], f[1], f[2]};
    }
    return Vec3{__fdiv_rn(f[0], n), __fdiv_rn(f[1], n), __fdiv_rn(f[2], n)};  // :170
}

__device__ __forceinline__ float mufu_sqrt(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float mufu_ex2(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// score(c) = exp(-acos(c)^2 / sigma^2) = 2^(-(sk*acos(c))^2),  sk = sqrt(log2 e)/sigma,  hp_sk = (pi/2)*sk.
// acos(c) = pi/2 - sgn(c)*u,  u = asin(|c|) for |c| <= 1/2,  u = pi/2 - 2 asin(sqrt((1-|c|)/2)) otherwise;
// asin(s) = s + s*z*R(z) with z = s^2 in [0, 1/4] and R a degree-4 minimax polynomial (fit: tools/fit_asin.py).
__device__ __forceinline__ float orient_score(float c, float sk, float hp_sk) {
    const float a = fminf(fabsf(c), 1.0f);
    const bool big = a > 0.5f;
    const float z = big ? fmaf(a, -0.5f, 0.5f) : a * a;
    const float s = big ? mufu_sqrt(z) : a;
    float r = 0.038206443190574646f;
    r = fmaf(r, z, 0.026494283229112625f);
    r = fmaf(r, z, 0.045010700821876526f);
    r = fmaf(r, z, 0.07498808950185776f);
    r = fmaf(r, z, 0.16666673123836517f);
    const float asn = fmaf(s * z, r, s);
    const float u = big ? fmaf(asn, -2.0f, 1.5707963267948966f) : asn;
    const float gs = fmaf(copysignf(u, c), -sk, hp_sk);  // sk * acos(c)
    return mufu_ex2(-gs * gs);
}

constexpr int K3_WARPS = 8;

template <int NB>
__global__ void __launch_bounds__(K3_WARPS * 32)
    orient_accumulate_kernel(const float *__restrict__ hn, const float *__restrict__ on, int S, int H, int O,
                             const double *__restrict__ grid, int N, int n_base, float sk, float hp_sk, float eps, int ord, Vec3 p,
                             Vec3 sp, float *__restrict__ PH, float *__restrict__ PO) {
    __shared__ __align__(16) float cnbuf[K3_WARPS][32][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long pair = (long long)blockIdx.x * K3_WARPS + warp;
    if (pair >= (long long)H * O) return;  // whole warp exits together; only __syncwarp below
    const int h = (int)(pair / O), o = (int)(pair % O);

    float gx[NB], gy[NB], gz[NB], ah[NB], ao[NB];
    float *ph = PH + (size_t)pair * N, *po = PO + (size_t)pair * N;
#pragma unroll
    for (int j = 0; j < NB; ++j) {
        const int n = n_base + lane + 32 * j;
        const bool ok = n < N;
        gx[j] = ok ? (float)grid[3 * n + 0] : 0.f;  // fp64 bin centres (ComA.canon_normal_grid), rounded once
        gy[j] = ok ? (float)grid[3 * n + 1] : 0.f;
        gz[j] = ok ? (float)grid[3 * n + 2] : 0.f;
        ah[j] = ok ? ph[n] : 0.f;
        ao[j] = ok ? po[n] : 0.f;
    }

    for (int s0 = 0; s0 < S; s0 += 32) {
        const int ns = min(32, S - s0);
        if (lane < ns) {
            const float *ph3 = hn + ((size_t)(s0 + lane) * H + h) * 3;
            const float *po3 = on + ((size_t)(s0 + lane) * O + o) * 3;
            const Vec3 a = normalize_ref(Vec3{ph3[0], ph3[1], ph3[2]}, eps, ord);
            const Vec3 b = normalize_ref(Vec3{po3[0], po3[1], po3[2]}, eps, ord);
            const Vec3 ch = canonicalize_ref(a, b, p, sp, eps, ord);  // human normal w.r.t. object normal (:295-301)
            const Vec3 co = canonicalize_ref(b, a, p, sp, eps, ord);  // object normal w.r.t. human normal (:302-309)
            float4 *dst = reinterpret_cast<float4 *>(&cnbuf[warp][lane][0]);
            dst[0] = make_float4(ch.x, ch.y, ch.z, co.x);
            dst[1] = make_float4(co.y, co.z, 0.f, 0.f);
        }
        __syncwarp();
        for (int i = 0; i < ns; ++i) {
            const float4 v0 = *reinterpret_cast<const float4 *>(&cnbuf[warp][i][0]);
            const float2 v1 = *reinterpret_cast<const float2 *>(&cnbuf[warp][i][4]);
#pragma unroll
            for (int j = 0; j < NB; ++j) {
                const float ch_c = fmaf(gx[j], v0.x, fmaf(gy[j], v0.y, gz[j] * v0.z));
                const float co_c = fmaf(gx[j], v0.w, fmaf(gy[j], v1.x, gz[j] * v1.y));
                ah[j] += orient_score(ch_c, sk, hp_sk);
                ao[j] += orient_score(co_c, sk, hp_sk);
            }
        }
        __syncwarp();
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) {
        const int n = n_base + lane + 32 * j;
        if (n < N) {
            ph[n] = ah[j];
            po[n] = ao[j];
        }
    }
}


// ---- packed FP32x2 variant (default) -----------------------------------------------------------------------------------
__device__ __forceinline__ float2 f2(float v) { return make_float2(v, v); }

// Two scores at once. sk2 = (sk, sk), hp2 = (hp_sk, hp_sk), nsk2 = (-sk, -sk).
__device__ __forceinline__ float2 orient_score2(float2 c, float2 nsk2, float2 hp2) {
    const float2 a = make_float2(fminf(fabsf(c.x), 1.0f), fminf(fabsf(c.y), 1.0f));
    const float2 w = __ffma2_rn(a, f2(-1.0f), f2(1.0f));
    const float2 s = make_float2(mufu_sqrt(w.x), mufu_sqrt(w.y));
    float2 r = f2(2.251368249e-03f);
    r = __ffma2_rn(r, a, f2(-1.101238653e-02f));
    r = __ffma2_rn(r, a, f2(2.674933150e-02f));
    r = __ffma2_rn(r, a, f2(-4.872440174e-02f));
    r = __ffma2_rn(r, a, f2(8.873733133e-02f));
    r = __ffma2_rn(r, a, f2(-2.145836949e-01f));
    r = __ffma2_rn(r, a, f2(1.570796132e+00f));
    const float2 ga = __fmul2_rn(s, r);                 // acos(|c|)
    const float2 u = __ffma2_rn(ga, nsk2, hp2);         // sk*(pi/2 - acos|c|) >= 0
    const float2 us = make_float2(copysignf(u.x, c.x), copysignf(u.y, c.y));
    const float2 gs = __ffma2_rn(us, f2(-1.0f), hp2);   // sk*acos(c)
    const float2 q = __fmul2_rn(gs, gs);
    return make_float2(mufu_ex2(-q.x), mufu_ex2(-q.y));
}

// K3_NP = bin PAIRS per lane -> 2*NP bins per lane, 64*NP bins per launch
template <int K3_NP, int MINB>
__global__ void __launch_bounds__(K3_WARPS * 32, MINB)
    orient_accumulate_kernel_x2(const float *__restrict__ hn, const float *__restrict__ on, int S, int H, int O,
                                const double *__restrict__ grid, int N, int n_base, float sk, float hp_sk, float eps, int ord, Vec3 p,
                                Vec3 sp, float *__restrict__ PH, float *__restrict__ PO) {
    // per warp: 32 samples x {chx,chx,chy,chy | chz,chz,cox,cox | coy,coy,coz,coz}: broadcast LDS.128 yields (v,v) pairs
    __shared__ __align__(16) float4 cnbuf[K3_WARPS][32][3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long pair = (long long)blockIdx.x * K3_WARPS + warp;
    if (pair >= (long long)H * O) return;
    const int h = (int)(pair / O), o = (int)(pair % O);

    // lane owns bins n = n_base + lane + 32*j, j < 8; packed as (j, j+4) so both halves are valid together more often
    float2 gx[K3_NP], gy[K3_NP], gz[K3_NP], ah[K3_NP], ao[K3_NP];
    float *ph = PH + (size_t)pair * N, *po = PO + (size_t)pair * N;
#pragma unroll
    for (int j = 0; j < K3_NP; ++j) {
        const int n0 = n_base + lane + 32 * j, n1 = n0 + 32 * K3_NP;
        const bool ok0 = n0 < N, ok1 = n1 < N;
        gx[j] = make_float2(ok0 ? (float)grid[3 * n0 + 0] : 0.f, ok1 ? (float)grid[3 * n1 + 0] : 0.f);
        gy[j] = make_float2(ok0 ? (float)grid[3 * n0 + 1] : 0.f, ok1 ? (float)grid[3 * n1 + 1] : 0.f);
        gz[j] = make_float2(ok0 ? (float)grid[3 * n0 + 2] : 0.f, ok1 ? (float)grid[3 * n1 + 2] : 0.f);
        ah[j] = make_float2(ok0 ? ph[n0] : 0.f, ok1 ? ph[n1] : 0.f);
        ao[j] = make_float2(ok0 ? po[n0] : 0.f, ok1 ? po[n1] : 0.f);
    }
    const float2 nsk2 = f2(-sk), hp2 = f2(hp_sk);

    for (int s0 = 0; s0 < S; s0 += 32) {
        const int ns = min(32, S - s0);
        if (lane < ns) {
            const float *ph3 = hn + ((size_t)(s0 + lane) * H + h) * 3;
            const float *po3 = on + ((size_t)(s0 + lane) * O + o) * 3;
            const Vec3 a = normalize_ref(Vec3{ph3[0], ph3[1], ph3[2]}, eps, ord);
            const Vec3 b = normalize_ref(Vec3{po3[0], po3[1], po3[2]}, eps, ord);
            const Vec3 ch = canonicalize_ref(a, b, p, sp, eps, ord);  // human normal w.r.t. object normal (:295-301)
            const Vec3 co = canonicalize_ref(b, a, p, sp, eps, ord);  // object normal w.r.t. human normal (:302-309)
            cnbuf[warp][lane][0] = make_float4(ch.x, ch.x, ch.y, ch.y);
            cnbuf[warp][lane][1] = make_float4(ch.z, ch.z, co.x, co.x);
            cnbuf[warp][lane][2] = make_float4(co.y, co.y, co.z, co.z);
        }
        __syncwarp();
#pragma unroll 2
        for (int i = 0; i < ns; ++i) {
            const float4 v0 = cnbuf[warp][i][0], v1 = cnbuf[warp][i][1], v2 = cnbuf[warp][i][2];
            const float2 hx = make_float2(v0.x, v0.y), hy = make_float2(v0.z, v0.w), hz = make_float2(v1.x, v1.y);
            const float2 ox = make_float2(v1.z, v1.w), oy = make_float2(v2.x, v2.y), oz = make_float2(v2.z, v2.w);
#pragma unroll
            for (int j = 0; j < K3_NP; ++j) {
                const float2 ch_c = __ffma2_rn(gx[j], hx, __ffma2_rn(gy[j], hy, __fmul2_rn(gz[j], hz)));
                const float2 co_c = __ffma2_rn(gx[j], ox, __ffma2_rn(gy[j], oy, __fmul2_rn(gz[j], oz)));
                ah[j] = __fadd2_rn(ah[j], orient_score2(ch_c, nsk2, hp2));
                ao[j] = __fadd2_rn(ao[j], orient_score2(co_c, nsk2, hp2));
            }
        }
        __syncwarp();
    }
#pragma unroll
    for (int j = 0; j < K3_NP; ++j) {
        const int n0 = n_base + lane + 32 * j, n1 = n0 + 32 * K3_NP;
        if (n0 < N) {
            ph[n0] = ah[j].x;
            po[n0] = ao[j].x;
        }
        if (n1 < N) {
            ph[n1] = ah[j].y;
            po[n1] = ao[j].y;
        }
    }
}

// ---- cone-limited variant (default whenever sigma is small enough) ------------------------------------------------------
// exp(-g^2/sigma^2) with sigma = 0.25 rad is below 2^-32 for every bin further than gamma = sigma*sqrt(32 ln 2) = 1.18 rad from
// the canonical normal: ~70 % of the 250 bins of a (pair, sample, histogram). This kernel never evaluates them:
//   * bins are grouped into <= 8 PATCHES of <= 32 spatially compact bins (coma_orient_bin_patches, host; any partition is
//     correct, compactness only buys speed); lane l owns bin l of every patch, so G and both accumulators stay in registers;
//   * per chunk of 32 samples, lane l canonicalises sample l (as before) and tests its two directions against every patch's
//     bounding cap: patch j is needed iff dir . centre_j > cos(gamma + r_j). A ballot turns that into the list of samples that
//     need patch j, and the warp walks the list TWO SAMPLES PER STEP (packed FP32x2 math), reading the directions back as
//     warp-broadcast LDS.128;
//   * inside the cone acos(c)^2 is an analytic function of t = 1 - c: acos(1-t)^2 = t R(t), R a degree 3-6 minimax polynomial
//     on [0, tmax] (tools/fit_acos2.py) — no square root: ONE MUFU (ex2) and 3 + DEG + 2 FMA-pipe ops per evaluation instead of
//     two MUFU and ~15. Beyond tmax = 1 - cos(gamma) the polynomial keeps growing, so a bin of a needed patch that lies outside
//     the cone receives some positive value below 2^-drop_bits instead of its true (smaller) score.
// Contract: every term the dense kernel would add and this one drops or clamps is < 2^-drop_bits (default 2^-32), i.e. after S
// samples each bin is within S * 2^-32 ABSOLUTE of the dense result — < 2e-7 of the pair's largest bin in the worst case (some
// bin of a pair always holds >= 0.002 S) — on top of the same ~5e-6 relative evaluation error as before.
struct ConeParams {
    float c[7];      // R's coefficients times -log2(e)/sigma^2: score = 2^(t * sum c[k] t^k)
    float tmax;      // 1 - cos(gamma)
    float gamma;     // cone half-angle (rad)
    int npatch;
};

// One patch, one histogram, one chunk: `cnt` (even) directions were compacted into the warp's list (structure of arrays x[], y[],
// z[], conflict-free to write), so three broadcast LDS.64 deliver a PAIR of samples as the packed operands of three FFMA2.
constexpr int K3C_LIST = 72;   // floats per component array: 64 samples of a chunk + the null entry, 16-byte multiple

template <int DEG>
__device__ __forceinline__ float cone_steps(int cnt, const float *__restrict__ list, float ngx, float ngy, float ngz, float acc,
                                            const ConeParams &cp) {
    const float2 c0 = f2(cp.c[0]), c1 = f2(cp.c[1]), c2 = f2(cp.c[2]), c3 = f2(cp.c[3]), c4 = f2(cp.c[DEG >= 4 ? 4 : 0]),
                 c5 = f2(cp.c[DEG >= 5 ? 5 : 0]), c6 = f2(cp.c[DEG >= 6 ? 6 : 0]);
    const float2 *lx = reinterpret_cast<const float2 *>(list), *ly = lx + K3C_LIST / 2, *lz = lx + K3C_LIST;
    const float2 gx2 = f2(ngx), gy2 = f2(ngy), gz2 = f2(ngz), one2 = f2(1.0f);
    // the pair's two samples accumulate into the two halves of one register pair (one FADD2 per step); within each half the
    // samples are added in ascending order, the halves meet once at the end
    float2 acc2 = make_float2(acc, 0.0f);
    const int pairs = cnt >> 1;
#pragma unroll 4
    for (int k = 0; k < pairs; ++k) {   // cnt is warp-uniform: no divergence
#ifdef K3C_SCALAR_DOT   // A/B (measured 2 % SLOWER: 692 vs 678 ms at cfg 4, S = 256): scalar FFMAs (lite-pipe eligible) for the dot product
        const float2 ax = lx[k], ay = ly[k], az = lz[k];
        const float2 t = make_float2(fmaf(ax.x, ngx, fmaf(ay.x, ngy, fmaf(az.x, ngz, 1.0f))), fmaf(ax.y, ngx, fmaf(ay.y, ngy, fmaf(az.y, ngz, 1.0f))));
#else
        const float2 t = __ffma2_rn(lx[k], gx2, __ffma2_rn(ly[k], gy2, __ffma2_rn(lz[k], gz2, one2)));   // 1 - G.n
#endif
        // No clamp at tmax: every fitted t*R(t) keeps growing on (tmax, 2] (tools/fit_acos2.py, "monotone beyond"), so a bin of a
        // needed patch that lies outside the cone receives a positive value BELOW 2^-drop_bits instead of its true, smaller score.
        float2 r = DEG == 6 ? c6 : DEG == 5 ? c5 : DEG == 4 ? c4 : c3;
        if (DEG >= 6) r = __ffma2_rn(r, t, c5);
        if (DEG >= 5) r = __ffma2_rn(r, t, c4);
        if (DEG >= 4) r = __ffma2_rn(r, t, c3);
        r = __ffma2_rn(r, t, c2);
        r = __ffma2_rn(r, t, c1);
        r = __ffma2_rn(r, t, c0);
        const float2 a = __fmul2_rn(t, r);
        acc2 = __fadd2_rn(acc2, make_float2(mufu_ex2(a.x), mufu_ex2(a.y)));
    }
    return acc2.x + acc2.y;
}

// Compacts the directions of the lanes' two samples (slot A: sample s0+lane, slot B: s0+32+lane) that need the patch into the
// warp's list, A entries first, both in ascending sample order; an odd count is padded with the NULL direction (0,0,0): t = 1,
// i.e. 90 degrees off every bin — beyond every cone this kernel accepts (gamma < 1.5), so the pad adds a term below 2^-drop_bits,
// like any other out-of-cone evaluation. Returns the (even, warp-uniform) entry count.
__device__ __forceinline__ int cone_compact(bool needA, Vec3 dA, bool needB, Vec3 dB, float *__restrict__ list, int lane) {
    const unsigned mA = __ballot_sync(0xffffffffu, needA), mB = __ballot_sync(0xffffffffu, needB);
    const unsigned lt = (1u << lane) - 1u;
    const int nA = __popc(mA), cnt = nA + __popc(mB);
    if (needA) {
        const int pos = __popc(mA & lt);
        list[pos] = dA.x;
        list[K3C_LIST + pos] = dA.y;
        list[2 * K3C_LIST + pos] = dA.z;
    }
    if (needB) {
        const int pos = nA + __popc(mB & lt);
        list[pos] = dB.x;
        list[K3C_LIST + pos] = dB.y;
        list[2 * K3C_LIST + pos] = dB.z;
    }
    if ((cnt & 1) && lane == 0) {
        list[cnt] = 0.f;
        list[K3C_LIST + cnt] = 0.f;
        list[2 * K3C_LIST + cnt] = 0.f;
    }
    __syncwarp();
    return (cnt + 1) & ~1;
}

constexpr int K3C_PATCHES = 8;

// The per-patch state (the lane's bin centre, its two accumulators) lives in SHARED memory and the patch loop is ROLLED: the
// first version kept it in registers with the loop unrolled 8x — 93 KB of SASS, `no_instruction` (instruction-cache misses) the
// second-largest stall and 80 registers (3 CTAs/SM). Rolled: ~12 KB of code, ~50 registers; the price is 7 conflict-free
// shared-memory accesses per (patch, chunk) against ~900 instructions of work. Chunks are 64 samples (two per lane), which
// halves the per-(patch, histogram) bookkeeping — ballots, compaction, loop set-up and remainder — per sample.
// PRENORM: hn / on already hold normalize_ref() of the normals (normalize_vectors_kernel, once per launch) — the normalisation depends on
// (sample, vertex) only, and recomputing it for every PAIR costs two IEEE square roots and six IEEE divisions per (pair, sample).
template <int DEG, int MINB, bool PRENORM>
__global__ void __launch_bounds__(K3_WARPS * 32, MINB)
    orient_accumulate_cone_kernel(const float *__restrict__ hn, const float *__restrict__ on, int S, int H, int O,
                                  const double *__restrict__ grid, int N, const int *__restrict__ perm, ConeParams cp, float eps, int ord,
                                  Vec3 p, Vec3 sp, float *__restrict__ PH, float *__restrict__ PO) {
    __shared__ float4 caps[K3C_PATCHES];                               // (centre, cos(min(gamma + radius, pi))) per patch
    __shared__ float sG[K3C_PATCHES][3][32];                           // negated bin centres: [patch][component][lane]
    __shared__ float sAcc[K3C_PATCHES][2][K3_WARPS * 32];              // accumulators: [patch][histogram][thread]
    __shared__ __align__(16) float lists[K3_WARPS][2][3 * K3C_LIST];   // per warp, per histogram: x[] y[] z[] of the needed samples
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, tid = threadIdx.x;
    const long long pair = (long long)blockIdx.x * K3_WARPS + warp;

    auto bin_of = [&](int j) -> int {   // bin owned by this lane in patch j (-1: none)
        const int slot = 32 * j + lane;
        return j < cp.npatch ? (perm ? __ldg(perm + slot) : (slot < N ? slot : -1)) : -1;
    };
    // bin centres + bounding cap of every patch (all warps of the CTA own the same bins: warp w prepares patch w)
    for (int j = warp; j < cp.npatch; j += K3_WARPS) {
        const int n = bin_of(j);
        const bool ok = n >= 0;
        const float gx = ok ? (float)grid[3 * n + 0] : 0.f, gy = ok ? (float)grid[3 * n + 1] : 0.f, gz = ok ? (float)grid[3 * n + 2] : 0.f;
        sG[j][0][lane] = -gx;   // fp64 bin centres, rounded once; stored negated (t = 1 - G.n)
        sG[j][1][lane] = -gy;
        sG[j][2][lane] = -gz;
        float sx = warp_sum(gx), sy = warp_sum(gy), sz = warp_sum(gz);
        const float nrm = sqrtf(sx * sx + sy * sy + sz * sz);
        if (nrm > 1e-6f) { sx /= nrm; sy /= nrm; sz /= nrm; } else { sx = 0.f; sy = 0.f; sz = 1.f; }
        float cmin = ok ? gx * sx + gy * sy + gz * sz : 1.0f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cmin = fminf(cmin, __shfl_xor_sync(0xffffffffu, cmin, o));
        const float reach = cp.gamma + acosf(fminf(fmaxf(cmin, -1.0f), 1.0f)) + 4e-3f;   // + margin for fp32 rounding of the test
        if (lane == 0) caps[j] = make_float4(sx, sy, sz, reach >= 3.14159f ? -2.0f : cosf(reach));
    }
    const bool live = pair < (long long)H * O;
    const int h = live ? (int)(pair / O) : 0, o = live ? (int)(pair % O) : 0;
    float *ph = PH + (size_t)(live ? pair : 0) * N, *po = PO + (size_t)(live ? pair : 0) * N;
    for (int j = 0; j < cp.npatch; ++j) {
        const int n = live ? bin_of(j) : -1;
        sAcc[j][0][tid] = n >= 0 ? ph[n] : 0.f;
        sAcc[j][1][tid] = n >= 0 ? po[n] : 0.f;
    }
    __syncthreads();
    if (!live) return;   // after the only block-wide barrier; below only __syncwarp (sAcc columns are thread-private)

    float *listH = lists[warp][0], *listO = lists[warp][1];
    for (int s0 = 0; s0 < S; s0 += 64) {
        Vec3 ch[2], co[2];
        bool have[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int s = s0 + 32 * u + lane;
            have[u] = s < S;
            ch[u] = co[u] = Vec3{0.f, 0.f, 1.f};
            if (have[u]) {
                const float *ph3 = hn + ((size_t)s * H + h) * 3;
                const float *po3 = on + ((size_t)s * O + o) * 3;
                const Vec3 a = PRENORM ? Vec3{ph3[0], ph3[1], ph3[2]} : normalize_ref(Vec3{ph3[0], ph3[1], ph3[2]}, eps, ord);
                const Vec3 b = PRENORM ? Vec3{po3[0], po3[1], po3[2]} : normalize_ref(Vec3{po3[0], po3[1], po3[2]}, eps, ord);
                ch[u] = canonicalize_ref<K3C_FAST_DIV != 0>(a, b, p, sp, eps, ord);  // human normal w.r.t. object normal (:295-301)
                co[u] = canonicalize_ref<K3C_FAST_DIV != 0>(b, a, p, sp, eps, ord);  // object normal w.r.t. human normal (:302-309)
            }
        }
#pragma unroll 1
        for (int j = 0; j < cp.npatch; ++j) {
            const float4 cap = caps[j];
            // a NaN direction (degenerate normal) must poison its bins like in the reference: !(x <= w) keeps it "needed"
            auto need = [&](const Vec3 &d) { return !(fmaf(d.x, cap.x, fmaf(d.y, cap.y, d.z * cap.z)) <= cap.w); };
            const int cntH = cone_compact(have[0] && need(ch[0]), ch[0], have[1] && need(ch[1]), ch[1], listH, lane);
            const int cntO = cone_compact(have[0] && need(co[0]), co[0], have[1] && need(co[1]), co[1], listO, lane);
            const float ngx = sG[j][0][lane], ngy = sG[j][1][lane], ngz = sG[j][2][lane];
            sAcc[j][0][tid] = cone_steps<DEG>(cntH, listH, ngx, ngy, ngz, sAcc[j][0][tid], cp);
            sAcc[j][1][tid] = cone_steps<DEG>(cntO, listO, ngx, ngy, ngz, sAcc[j][1][tid], cp);
            __syncwarp();   // the lists are rewritten for the next patch
        }
    }
    for (int j = 0; j < cp.npatch; ++j) {
        const int n = bin_of(j);
        if (n >= 0) {
            ph[n] = sAcc[j][0][tid];
            po[n] = sAcc[j][1][tid];
        }
    }
}

// out[i] = normalize_ref(in[i]) — utils/transformations.py:14-17 once per (sample, vertex), bit-identical to the in-kernel evaluation
__global__ void normalize_vectors_kernel(const float *__restrict__ in, long long n, float eps, int ord, float *__restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Vec3 v = normalize_ref(Vec3{in[3 * i], in[3 * i + 1], in[3 * i + 2]}, eps, ord);
    out[3 * i + 0] = v.x;
    out[3 * i + 1] = v.y;
    out[3 * i + 2] = v.z;
}

__global__ void canonicalize_kernel(const float *__restrict__ a, int A, const float *__restrict__ b, int B, Vec3 p, Vec3 sp,
                                    float eps, int ord, float *__restrict__ out) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= (long long)A * B) return;
    const int i = (int)(q / B), j = (int)(q % B);
    const Vec3 an = normalize_ref(Vec3{a[3 * i], a[3 * i + 1], a[3 * i + 2]}, eps, ord);
    const Vec3 bn = normalize_ref(Vec3{b[3 * j], b[3 * j + 1], b[3 * j + 2]}, eps, ord);
    const Vec3 f = canonicalize_ref(an, bn, p, sp, eps, ord);
    out[3 * q + 0] = f.x;
    out[3 * q + 1] = f.y;
    out[3 * q + 2] = f.z;
}

// host-side mirror of normalize_vectors_torch for the two principle vectors (3 floats: exact fp32 host arithmetic)
static Vec3 normalize_host(const float *v, float eps) {
    volatile float xx = v[0] * v[0], yy = v[1] * v[1], zz = v[2] * v[2];
    volatile float s = xx + yy;
    s = s + zz;
    volatile float n = sqrtf(s);
    n = n + eps;
    return Vec3{v[0] / n, v[1] / n, v[2] / n};
}

}  // namespace coma

namespace coma {
// acos(1-t)^2 = t R(t): generated by tools/fit_acos2.py --emit (R minimax on [0, tmax], max |dR| listed)
struct Acos2Fit {
    float tmax;
    int deg;
    float c[7];
};
static const Acos2Fit kAcos2Fits[] = {
    {0.10f, 3, {1.999999991e+00f, 3.333361032e-01f, 8.875197085e-02f, 3.071955545e-02f, 0.000000000e+00f, 0.000000000e+00f, 0.000000000e+00f}},   // |dR| <= 8.8e-09
    {0.20f, 4, {2.000000003e+00f, 3.333325774e-01f, 8.891860234e-02f, 2.816586494e-02f, 1.237028604e-02f, 0.000000000e+00f, 0.000000000e+00f}},   // |dR| <= 3.1e-09
    {0.32f, 4, {2.000000038e+00f, 3.333275810e-01f, 8.902846103e-02f, 2.740019559e-02f, 1.403493065e-02f, 0.000000000e+00f, 0.000000000e+00f}},   // |dR| <= 3.8e-08
    {0.46f, 5, {1.999999985e+00f, 3.333355296e-01f, 8.883518296e-02f, 2.904723379e-02f, 8.295234272e-03f, 7.016475926e-03f, 0.000000000e+00f}},   // |dR| <= 1.5e-08
    {0.64f, 5, {1.999999854e+00f, 3.333489836e-01f, 8.861818956e-02f, 3.026155092e-02f, 5.534580991e-03f, 9.209118341e-03f, 0.000000000e+00f}},   // |dR| <= 1.5e-07
    {0.86f, 6, {2.000000152e+00f, 3.333169853e-01f, 8.917436458e-02f, 2.672181738e-02f, 1.573746627e-02f, -4.379195104e-03f, 6.805618570e-03f}},   // |dR| <= 1.5e-07
};

static int launch_dense(const float *hn, const float *on, int64_t S, int64_t H, int64_t O, const double *grid, int64_t N, float skf,
                        float hp, float epsf, int ord, Vec3 p, Vec3 sp, float *PH, float *PO, cudaStream_t st) {
    static const char *const variant = getenv("COMA_B200_K3");  // experiments only (read once): "v1" selects the scalar-FP32 kernel
    const bool use_x2 = !(variant && variant[0] == 'v' && variant[1] == '1');
    // x3 (default: 8 bins/lane, 80 regs, 3 CTAs/SM) | x2 (128 regs, 2 CTAs/SM) | x4 (4 bins/lane, 4 CTAs/SM, two passes)
    const int x2_kind = (variant && variant[0] == 'x') ? atoi(variant + 1) : 3;
    const long long pairs = (long long)H * O;
    const unsigned blocks = (unsigned)((pairs + K3_WARPS - 1) / K3_WARPS);
    const int64_t bins_per_launch = (use_x2 && x2_kind == 4) ? 128 : 256;
    for (int64_t n_base = 0; n_base < N; n_base += bins_per_launch) {
        const int64_t rem = N - n_base;
        const int nb = (int)((rem > 256 ? 256 : rem) + 31) / 32;
#define LAUNCH(NBV)                                                                                                   \
    orient_accumulate_kernel<NBV><<<blocks, K3_WARPS * 32, 0, st>>>(hn, on, (int)S, (int)H, (int)O, grid, (int)N,     \
                                                                     (int)n_base, skf, hp, epsf, ord, p, sp, PH, PO)
#define LAUNCH_X2(NP, MINB)                                                                                          \
    orient_accumulate_kernel_x2<NP, MINB><<<blocks, K3_WARPS * 32, 0, st>>>(hn, on, (int)S, (int)H, (int)O, grid, (int)N, \
                                                                            (int)n_base, skf, hp, epsf, ord, p, sp, PH, PO)
        if (use_x2) {
            if (x2_kind == 3) LAUNCH_X2(4, 3);
            else if (x2_kind == 4) LAUNCH_X2(2, 4);
            else LAUNCH_X2(4, 2);
        } else if (nb <= 1) LAUNCH(1);
        else if (nb <= 2) LAUNCH(2);
        else if (nb <= 4) LAUNCH(4);
        else LAUNCH(8);
#undef LAUNCH
#undef LAUNCH_X2
        if (int e = check_launch(use_x2 ? "orient_accumulate_kernel_x2" : "orient_accumulate_kernel")) return e;
    }
    return 0;
}
}  // namespace coma

extern "C" int coma_orient_accumulate_cone_f32(const float *hn, const float *on, int64_t S, int64_t H, int64_t O,
                                               const double *grid, int64_t N, double sigma, double eps, const float *p_host,
                                               const float *sub_p_host, const int32_t *bin_perm, int drop_bits, int sum_order, float *PH,
                                               float *PO, coma_stream_t stream) {
    return coma_orient_accumulate_cone_ws_f32(hn, on, S, H, O, grid, N, sigma, eps, p_host, sub_p_host, bin_perm, drop_bits, sum_order, PH, PO,
                                              nullptr, stream);
}

extern "C" int coma_orient_accumulate_cone_ws_f32(const float *hn, const float *on, int64_t S, int64_t H, int64_t O,
                                                  const double *grid, int64_t N, double sigma, double eps, const float *p_host,
                                                  const float *sub_p_host, const int32_t *bin_perm, int drop_bits, int sum_order, float *PH,
                                                  float *PO, float *workspace, coma_stream_t stream) {
    using namespace coma;
    COMA_REQUIRE(hn && on && grid && p_host && sub_p_host && PH && PO, "null pointer");
    COMA_REQUIRE(S >= 0 && H > 0 && O > 0 && N > 0, "bad sizes");
    COMA_REQUIRE(sigma > 0.0, "normal_gaussian_sigma must be positive");
    COMA_REQUIRE(H * O < (int64_t)1 << 34, "H*O out of range");
    COMA_REQUIRE(drop_bits == 0 || (drop_bits >= 24 && drop_bits <= 120), "drop_bits must be 0 (dense) or in [24, 120]");
    COMA_REQUIRE(sum_order == COMA_SUM_ORDER_TORCH_CPU || sum_order == COMA_SUM_ORDER_TORCH_CUDA, "sum_order must be 0 (torch CPU) or 1 (torch CUDA)");
    const int ord = sum_order;
    if (S == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const float epsf = (float)eps;
    const Vec3 p = normalize_host(p_host, epsf), sp = normalize_host(sub_p_host, epsf);
    const double sk = sqrt(1.4426950408889634) / sigma;
    // cone form: needs the whole histogram in one launch (N <= 256) and a cone inside the polynomial's domain
    const double gamma = sigma * sqrt(drop_bits * 0.6931471805599453);
    const double tneed = 1.0 - cos(gamma);
    const Acos2Fit *fit = nullptr;
    if (drop_bits > 0 && N <= 32 * K3C_PATCHES && gamma < 1.5)
        for (const Acos2Fit &f : kAcos2Fits)
            if (tneed <= f.tmax) {
                fit = &f;
                break;
            }
    static const bool force_dense = getenv("COMA_B200_K3_DENSE") != nullptr;   // A/B switch for benchmarks (read once)
    if (!fit || force_dense)
        return launch_dense(hn, on, S, H, O, grid, N, (float)sk, (float)(sk * 1.5707963267948966), epsf, ord, p, sp, PH, PO, st);
    ConeParams cp;
    for (int k = 0; k < 7; ++k) cp.c[k] = (float)(-sk * sk * (double)fit->c[k]);
    cp.tmax = (float)tneed;
    cp.gamma = (float)gamma;
    cp.npatch = (int)((N + 31) / 32);
    const long long pairs = (long long)H * O;
    const unsigned blocks = (unsigned)((pairs + K3_WARPS - 1) / K3_WARPS);
    static const int ctas = getenv("COMA_B200_K3C_CTAS") ? atoi(getenv("COMA_B200_K3C_CTAS")) : 4;   // A/B: CTAs per SM (read once; 4 = 64 registers, no spills)
    // workspace (>= 3 S (H + O) floats): the normals are normalised ONCE per (sample, vertex) instead of once per (pair, sample)
    if (workspace) {
        float *hn_n = workspace, *on_n = workspace + 3 * S * H;
        const long long nh = (long long)S * H, no = (long long)S * O;
        normalize_vectors_kernel<<<(unsigned)((nh + 255) / 256), 256, 0, st>>>(hn, nh, epsf, ord, hn_n);
        normalize_vectors_kernel<<<(unsigned)((no + 255) / 256), 256, 0, st>>>(on, no, epsf, ord, on_n);
        if (int e = check_launch("normalize_vectors_kernel")) return e;
        hn = hn_n;
        on = on_n;
    }
#define LAUNCH_CONE(DEG)                                                                                                                \
    if (workspace)                                                                                                                      \
        orient_accumulate_cone_kernel<DEG, 4, true><<<blocks, K3_WARPS * 32, 0, st>>>(hn, on, (int)S, (int)H, (int)O, grid, (int)N, bin_perm, \
                                                                                      cp, epsf, ord, p, sp, PH, PO);                    \
    else if (ctas == 4)                                                                                                                 \
        orient_accumulate_cone_kernel<DEG, 4, false><<<blocks, K3_WARPS * 32, 0, st>>>(hn, on, (int)S, (int)H, (int)O, grid, (int)N, bin_perm, \
                                                                                       cp, epsf, ord, p, sp, PH, PO);                   \
    else if (ctas == 6)                                                                                                                 \
        orient_accumulate_cone_kernel<DEG, 6, false><<<blocks, K3_WARPS * 32, 0, st>>>(hn, on, (int)S, (int)H, (int)O, grid, (int)N, bin_perm, \
                                                                                       cp, epsf, ord, p, sp, PH, PO);                   \
    else                                                                                                                                \
        orient_accumulate_cone_kernel<DEG, 5, false><<<blocks, K3_WARPS * 32, 0, st>>>(hn, on, (int)S, (int)H, (int)O, grid, (int)N, bin_perm, \
                                                                                       cp, epsf, ord, p, sp, PH, PO)
    switch (fit->deg) {
        case 3: LAUNCH_CONE(3); break;
        case 4: LAUNCH_CONE(4); break;
        case 5: LAUNCH_CONE(5); break;
        default: LAUNCH_CONE(6); break;
    }
#undef LAUNCH_CONE
    return check_launch("orient_accumulate_cone_kernel");
}

extern "C" int coma_orient_accumulate_f32(const float *hn, const float *on, int64_t S, int64_t H, int64_t O,
                                          const double *grid, int64_t N, double sigma, double eps, const float *p_host,
                                          const float *sub_p_host, float *PH, float *PO, coma_stream_t stream) {
    return coma_orient_accumulate_cone_f32(hn, on, S, H, O, grid, N, sigma, eps, p_host, sub_p_host, nullptr, 0, COMA_SUM_ORDER_TORCH_CPU,
                                           PH, PO, stream);
}

namespace coma {
// Recursive bisection along the principal axis of the point set: `leaves` groups of <= 32 points each, spatially contiguous.
static void bisect_bins(const std::vector<double> &g, std::vector<int> idx, int leaves, std::vector<std::vector<int>> &out) {
    if (leaves == 1) {
        out.push_back(idx);
        return;
    }
    const int n = (int)idx.size();
    double c[3] = {0, 0, 0}, M[3][3] = {{0}};
    for (int i : idx) for (int d = 0; d < 3; ++d) c[d] += g[3 * i + d] / n;
    for (int i : idx)
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) M[a][b] += (g[3 * i + a] - c[a]) * (g[3 * i + b] - c[b]);
    double v[3] = {0.5773, 0.5774, 0.5775};   // power iteration: dominant eigenvector of the 3x3 scatter matrix
    for (int it = 0; it < 200; ++it) {
        double w[3] = {0, 0, 0};
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) w[a] += M[a][b] * v[b];
        const double nrm = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
        if (nrm < 1e-300) break;
        for (int a = 0; a < 3; ++a) v[a] = w[a] / nrm;
    }
    std::vector<std::pair<double, int>> proj;
    for (int i : idx) proj.push_back({(g[3 * i] - c[0]) * v[0] + (g[3 * i + 1] - c[1]) * v[1] + (g[3 * i + 2] - c[2]) * v[2], i});
    std::stable_sort(proj.begin(), proj.end());
    const int ll = leaves / 2, lr = leaves - ll;
    int nl = (int)llround((double)n * ll / leaves);
    nl = std::max(nl, n - 32 * lr);   // the right part must fit its leaves
    nl = std::min(nl, 32 * ll);       // and so must the left
    std::vector<int> left, right;
    for (int q = 0; q < n; ++q) (q < nl ? left : right).push_back(proj[q].second);
    bisect_bins(g, left, ll, out);
    bisect_bins(g, right, lr, out);
}
}  // namespace coma

// Host-only: group the N bin centres into ceil(N/32) patches of <= 32 spatially compact bins (recursive principal-axis bisection,
// deterministic; bounding-cap radius 0.80-0.88 rad for the 250-point Fibonacci sphere, the 8-cap covering bound being 0.84).
// perm_host[32*j + l] = bin index of lane l in patch j, -1 where a patch has fewer than 32 bins.
extern "C" int coma_orient_bin_patches(const double *grid_host, int64_t N, int32_t *perm_host) {
    using namespace coma;
    COMA_REQUIRE(grid_host && perm_host, "null pointer");
    COMA_REQUIRE(N > 0 && N <= 32 * K3C_PATCHES, "N must be in [1, 256]");
    const int n = (int)N, k = (n + 31) / 32;
    std::vector<double> g(grid_host, grid_host + 3 * n);
    std::vector<int> all(n);
    for (int i = 0; i < n; ++i) all[i] = i;
    std::vector<std::vector<int>> parts;
    bisect_bins(g, all, k, parts);
    for (int q = 0; q < 32 * k; ++q) perm_host[q] = -1;
    for (int j = 0; j < k; ++j) {
        std::sort(parts[j].begin(), parts[j].end());   // ascending bin index inside a patch: neighbouring lanes, neighbouring addresses
        for (size_t l = 0; l < parts[j].size(); ++l) perm_host[32 * j + (int)l] = parts[j][l];
    }
    return 0;
}

extern "C" int coma_canonicalize_order_f32(const float *a, int64_t A, const float *b, int64_t B, const float *p_host,
                                           const float *sub_p_host, float eps, int sum_order, float *out, coma_stream_t stream);
extern "C" int coma_canonicalize_f32(const float *a, int64_t A, const float *b, int64_t B, const float *p_host,
                                     const float *sub_p_host, float eps, float *out, coma_stream_t stream) {
    return coma_canonicalize_order_f32(a, A, b, B, p_host, sub_p_host, eps, COMA_SUM_ORDER_TORCH_CPU, out, stream);
}

extern "C" int coma_canonicalize_order_f32(const float *a, int64_t A, const float *b, int64_t B, const float *p_host,
                                           const float *sub_p_host, float eps, int sum_order, float *out, coma_stream_t stream) {
    using namespace coma;
    COMA_REQUIRE(a && b && p_host && sub_p_host && out, "null pointer");
    COMA_REQUIRE(sum_order == 0 || sum_order == 1, "sum_order must be 0 (torch CPU) or 1 (torch CUDA)");
    COMA_REQUIRE(A > 0 && B > 0 && A * B < (int64_t)1 << 38, "bad sizes");
    const Vec3 p = normalize_host(p_host, eps), sp = normalize_host(sub_p_host, eps);
    const long long q = (long long)A * B;
    canonicalize_kernel<<<(unsigned)((q + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a, (int)A, b, (int)B, p, sp, eps, sum_order, out);
    return check_launch("canonicalize_kernel");
}
