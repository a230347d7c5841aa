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
#include <stdlib.h>

#include "common.cuh"

namespace coma {

struct Vec3 {
    float x, y, z;
};

// v / (sqrt((x^2+y^2)+z^2) + eps)   — utils/transformations.py:14-17
__device__ __forceinline__ Vec3 normalize_ref(Vec3 v, float eps) {
    float n = __fadd_rn(__fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y)), __fmul_rn(v.z, v.z))), eps);
    return Vec3{__fdiv_rn(v.x, n), __fdiv_rn(v.y, n), __fdiv_rn(v.z, n)};
}

__device__ __forceinline__ float dot_ref(Vec3 a, Vec3 b) {
    return __fadd_rn(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fmul_rn(a.z, b.z));
}

// canonicalize_a_wrt_b_to_p for one (a, b) pair; a, b, p, sp already normalised — utils/coma.py:135-170
__device__ __forceinline__ Vec3 canonicalize_ref(Vec3 a, Vec3 b, Vec3 p, Vec3 sp, float eps) {
    const float b_dot_p = dot_ref(b, p), a_dot_b = dot_ref(a, b), a_dot_p = dot_ref(a, p), a_dot_sp = dot_ref(a, sp);
    const float one_plus = __fadd_rn(1.0f, b_dot_p);
    const bool replace = one_plus < eps;  // :143
    // rows of the reference's "cross product matrix" (:149-155): [b0,-b2,b1], [b2,0,-b0], [-b1,0,0]
    Vec3 bxp;
    bxp.x = __fadd_rn(__fadd_rn(__fmul_rn(b.x, p.x), __fmul_rn(-b.z, p.y)), __fmul_rn(b.y, p.z));
    bxp.y = __fadd_rn(__fadd_rn(__fmul_rn(b.z, p.x), __fmul_rn(0.0f, p.y)), __fmul_rn(-b.x, p.z));
    bxp.z = __fadd_rn(__fadd_rn(__fmul_rn(-b.y, p.x), __fmul_rn(0.0f, p.y)), __fmul_rn(0.0f, p.z));
    const float a_dot_bxp = dot_ref(a, bxp);  // :159
    const float av[3] = {a.x, a.y, a.z}, bv[3] = {b.x, b.y, b.z}, pv[3] = {p.x, p.y, p.z}, sv[3] = {sp.x, sp.y, sp.z};
    const float xv[3] = {bxp.x, bxp.y, bxp.z};
    float f[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float v = __fmul_rn(xv[k], a_dot_bxp);                 // :162
        v = replace ? 0.0f : __fdiv_rn(v, one_plus);           // :163
        v = __fadd_rn(v, __fmul_rn(b_dot_p, av[k]));           // :164
        v = __fadd_rn(v, __fmul_rn(a_dot_b, pv[k]));           // :165
        v = __fsub_rn(v, __fmul_rn(a_dot_p, bv[k]));           // :166
        if (replace) v = __fsub_rn(__fmul_rn(__fmul_rn(2.0f, a_dot_sp), sv[k]), av[k]);  // :145,:169
        f[k] = v;
    }
    const float n = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(f[0], f[0]), __fmul_rn(f[1], f[1])), __fmul_rn(f[2], f[2])));
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
                             const double *__restrict__ grid, int N, int n_base, float sk, float hp_sk, float eps, Vec3 p,
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
            const Vec3 a = normalize_ref(Vec3{ph3[0], ph3[1], ph3[2]}, eps);
            const Vec3 b = normalize_ref(Vec3{po3[0], po3[1], po3[2]}, eps);
            const Vec3 ch = canonicalize_ref(a, b, p, sp, eps);  // human normal w.r.t. object normal (:295-301)
            const Vec3 co = canonicalize_ref(b, a, p, sp, eps);  // object normal w.r.t. human normal (:302-309)
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
                                const double *__restrict__ grid, int N, int n_base, float sk, float hp_sk, float eps, Vec3 p,
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
            const Vec3 a = normalize_ref(Vec3{ph3[0], ph3[1], ph3[2]}, eps);
            const Vec3 b = normalize_ref(Vec3{po3[0], po3[1], po3[2]}, eps);
            const Vec3 ch = canonicalize_ref(a, b, p, sp, eps);  // human normal w.r.t. object normal (:295-301)
            const Vec3 co = canonicalize_ref(b, a, p, sp, eps);  // object normal w.r.t. human normal (:302-309)
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

__global__ void canonicalize_kernel(const float *__restrict__ a, int A, const float *__restrict__ b, int B, Vec3 p, Vec3 sp,
                                    float eps, float *__restrict__ out) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= (long long)A * B) return;
    const int i = (int)(q / B), j = (int)(q % B);
    const Vec3 an = normalize_ref(Vec3{a[3 * i], a[3 * i + 1], a[3 * i + 2]}, eps);
    const Vec3 bn = normalize_ref(Vec3{b[3 * j], b[3 * j + 1], b[3 * j + 2]}, eps);
    const Vec3 f = canonicalize_ref(an, bn, p, sp, eps);
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

extern "C" int coma_orient_accumulate_f32(const float *hn, const float *on, int64_t S, int64_t H, int64_t O,
                                          const double *grid, int64_t N, double sigma, double eps, const float *p_host,
                                          const float *sub_p_host, float *PH, float *PO, coma_stream_t stream) {
    using namespace coma;
    COMA_REQUIRE(hn && on && grid && p_host && sub_p_host && PH && PO, "null pointer");
    COMA_REQUIRE(S >= 0 && H > 0 && O > 0 && N > 0, "bad sizes");
    COMA_REQUIRE(sigma > 0.0, "normal_gaussian_sigma must be positive");
    COMA_REQUIRE(H * O < (int64_t)1 << 34, "H*O out of range");
    if (S == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const float epsf = (float)eps;
    const Vec3 p = normalize_host(p_host, epsf), sp = normalize_host(sub_p_host, epsf);
    const double sk = sqrt(1.4426950408889634) / sigma;
    const float skf = (float)sk, hp = (float)(sk * 1.5707963267948966);
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
                                                                     (int)n_base, skf, hp, epsf, p, sp, PH, PO)
#define LAUNCH_X2(NP, MINB)                                                                                          \
    orient_accumulate_kernel_x2<NP, MINB><<<blocks, K3_WARPS * 32, 0, st>>>(hn, on, (int)S, (int)H, (int)O, grid, (int)N, \
                                                                            (int)n_base, skf, hp, epsf, p, sp, PH, PO)
        if (use_x2) {
            if (x2_kind == 3) LAUNCH_X2(4, 3);
            else if (x2_kind == 4) LAUNCH_X2(2, 4);
            else LAUNCH_X2(4, 2);
        } else if (nb <= 1) LAUNCH(1);
        else if (nb <= 2) LAUNCH(2);
        else if (nb <= 4) LAUNCH(4);
        else LAUNCH(8);
#undef LAUNCH
        if (int e = check_launch("orient_accumulate_kernel")) return e;
    }
    return 0;
}

extern "C" int coma_canonicalize_f32(const float *a, int64_t A, const float *b, int64_t B, const float *p_host,
                                     const float *sub_p_host, float eps, float *out, coma_stream_t stream) {
    using namespace coma;
    COMA_REQUIRE(a && b && p_host && sub_p_host && out, "null pointer");
    COMA_REQUIRE(A > 0 && B > 0 && A * B < (int64_t)1 << 38, "bad sizes");
    const Vec3 p = normalize_host(p_host, eps), sp = normalize_host(sub_p_host, eps);
    const long long q = (long long)A * B;
    canonicalize_kernel<<<(unsigned)((q + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a, (int)A, b, (int)B, p, sp, eps, out);
    return check_launch("canonicalize_kernel");
}
