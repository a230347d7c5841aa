// K5 — read-out reductions (reference: utils/coma.py:328-476, utils/coma_occupancy.py:297-312).
// All of them are single streaming passes over the accumulators (HBM-bound): one warp per (h,o) row of the [H*O, N]
// grids with coalesced 128-byte row segments and shuffle reductions, persistent grid of kNumSM * 8 CTAs.
#include <math.h>

#include "common.cuh"

namespace coma {

constexpr int K5_WARPS = 8;

// x / d, correctly rounded. div.rn's operand range check (FCHK) sends a ZERO numerator down its ~60-instruction slow path, and the
// occupancy grids K5c normalises are full of exact zeros (voxels a vertex never touched). When the
// divisor is a positive, finite, normal number — `plain`, uniform over a row — 0 / d is the zero itself, sign included, so only the
// non-zero numerators are divided; any other divisor (0 -> NaN, negative -> -0.0, NaN, inf) takes the generic division.
__device__ __forceinline__ bool plain_divisor(float d) { return d >= 1.17549435e-38f && d < INFINITY; }
__device__ __forceinline__ float div_rn_zero_fast(float x, float d, bool plain) {
    if (!plain) return __fdiv_rn(x, d);
    const float q = __fdiv_rn(x == 0.f ? 1.0f : x, d);
    return x == 0.f ? x : q;
}

// K5a: P[q,:] /= (sum P[q,:] + eps) in place; cmap[q] = (sum_n P[q,n] w[n]) * nom[q]/denom[q]
__global__ void __launch_bounds__(K5_WARPS * 32)
    normalize_contact_kernel(float *__restrict__ P, long long HO, int N, float eps, const float *__restrict__ w,
                             const float *__restrict__ nom, const float *__restrict__ denom, float *__restrict__ cmap) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * K5_WARPS + (threadIdx.x >> 5), nwarps = (long long)gridDim.x * K5_WARPS;
    for (long long q = warp0; q < HO; q += nwarps) {
        float *row = P + (size_t)q * N;
        float s = 0.f;
        for (int n = lane; n < N; n += 32) s += row[n];
        s = warp_sum(s);
        const float d = __fadd_rn(s, eps);
        float acc = 0.f;
        for (int n = lane; n < N; n += 32) {
            const float v = __fdiv_rn(row[n], d);
            row[n] = v;
            if (cmap) acc = fmaf(v, w[n], acc);
        }
        if (cmap) {
            acc = warp_sum(acc);
            if (lane == 0) cmap[q] = acc * __fdiv_rn(nom[q], denom[q]);
        }
    }
}

// K5a, N <= 256 (the reference's 250 normal bins): a lane keeps its <= 8 row entries in registers, so a row is read ONCE and
// written once (the generic kernel re-reads it for the second pass); two rows per warp iteration keep 16 independent
// 4-byte loads per lane in flight. Same per-lane accumulation order as the generic kernel -> identical results.
__global__ void __launch_bounds__(K5_WARPS * 32)
    normalize_contact_reg_kernel(float *__restrict__ P, long long HO, int N, float eps, const float *__restrict__ w,
                                 const float *__restrict__ nom, const float *__restrict__ denom, float *__restrict__ cmap) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * K5_WARPS + (threadIdx.x >> 5), nwarps = (long long)gridDim.x * K5_WARPS;
    float wv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) wv[j] = (cmap && lane + 32 * j < N) ? w[lane + 32 * j] : 0.f;
    for (long long q0 = 2 * warp0; q0 < HO; q0 += 2 * nwarps) {
        float v[2][8];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const float *row = P + (size_t)(q0 + r) * N;
#pragma unroll
            for (int j = 0; j < 8; ++j) v[r][j] = (q0 + r < HO && lane + 32 * j < N) ? __ldcs(row + lane + 32 * j) : 0.f;
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            if (q0 + r >= HO) break;
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (lane + 32 * j < N) s += v[r][j];
            s = warp_sum(s);
            const float d = __fadd_rn(s, eps);
            float acc = 0.f;
            float *row = P + (size_t)(q0 + r) * N;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (lane + 32 * j < N) {
                    const float x = __fdiv_rn(v[r][j], d);   // (the zero-fast form below costs this register-bound kernel 12 %: 5051 -> 4487 GB/s on dense data)
                    __stcs(row + lane + 32 * j, x);
                    acc = fmaf(x, wv[j], acc);
                }
            }
            if (cmap) {
                acc = warp_sum(acc);
                if (lane == 0) cmap[q0 + r] = acc * __fdiv_rn(nom[q0 + r], denom[q0 + r]);
            }
        }
    }
}

// K5b: entropy read-out of a normalised grid (utils/coma.py:455-463): q = round(P n_bin) / n_bin, score = 1 + sum q ln q / ln n_bin.
// With k = rint(P n_bin) (an exact integer, torch.round's half-to-even):  sum q ln q / ln n_bin = sum k (log2 k - log2 n_bin) /
// (n_bin log2 n_bin) — one MUFU.LG2 and ~5 FP32 ops per element instead of an IEEE division + logf (~35 instructions), which is
// what held round 1's kernel at 20 % of the HBM rate. Two rows per warp in flight, 8-byte loads when N is even. Error vs the
// literal expression: <= 3e-7 absolute on the score (lg2.approx carries 2^-22; the reference's own q = k / n_bin division is
// replaced by the mathematically identical k-form), far inside the 2e-5 the round-half flips of P already impose.
__device__ __forceinline__ float entropy_term(float pv, float n_bin, float lg_nb) {
    const float k = rintf(__fmul_rn(pv, n_bin));
    float lg;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(k));
    return k == 0.0f ? 0.0f : k * (lg - lg_nb);   // k = 0: 0 * log 0 -> 0 (:459); NaN / negative inputs give NaN like torch.log
}

__global__ void __launch_bounds__(K5_WARPS * 32)
    entropy_kernel(const float *__restrict__ P, long long HO, int N, float n_bin, float lg_nb, float inv_norm, const float *__restrict__ wgt,
                   float base, float *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * K5_WARPS + (threadIdx.x >> 5), nwarps = (long long)gridDim.x * K5_WARPS;
    // wgt != NULL: the `_v2` read-out (utils/coma.py:529-579), every term weighted by the bin's alignment with the principle vector
    const bool vec2 = (N & 1) == 0 && (reinterpret_cast<uintptr_t>(P) & 7) == 0 && !wgt;
    for (long long q = 2 * warp0; q < HO; q += 2 * nwarps) {
        const bool two = q + 1 < HO;
        const float *r0 = P + (size_t)q * N, *r1 = P + (size_t)(two ? q + 1 : q) * N;
        float a0 = 0.f, a1 = 0.f;
        if (vec2) {
            const float2 *v0 = reinterpret_cast<const float2 *>(r0), *v1 = reinterpret_cast<const float2 *>(r1);
            for (int n = lane; n < N / 2; n += 32) {
                const float2 x = __ldcs(v0 + n), y = __ldcs(v1 + n);
                a0 += entropy_term(x.x, n_bin, lg_nb) + entropy_term(x.y, n_bin, lg_nb);
                a1 += entropy_term(y.x, n_bin, lg_nb) + entropy_term(y.y, n_bin, lg_nb);
            }
        } else {
            for (int n = lane; n < N; n += 32) {
                const float x = r0[n], y = r1[n], w = wgt ? wgt[n] : 1.0f;
                a0 += entropy_term(x, n_bin, lg_nb) * w;
                a1 += entropy_term(y, n_bin, lg_nb) * w;
            }
        }
        a0 = warp_sum(a0);
        a1 = warp_sum(a1);
        if (lane == 0) {
            out[q] = fmaf(a0, inv_norm, base);
            if (two) out[q + 1] = fmaf(a1, inv_norm, base);
        }
    }
}

__global__ void significant_pairs_kernel(const float *__restrict__ count, int H, int O, float num, uint8_t *__restrict__ sig,
                                         uint8_t *__restrict__ any_o, uint8_t *__restrict__ any_h) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= (long long)H * O) return;
    const bool s = count[q] >= num;
    if (sig) sig[q] = s ? 1 : 0;
    if (s) {  // benign races: every writer stores the same value
        if (any_o) any_o[q / O] = 1;
        if (any_h) any_h[q % O] = 1;
    }
}

// axis = 1: out[h] = max over masked o (one warp per row)
__global__ void __launch_bounds__(256)
    masked_rowmax_kernel(const float *__restrict__ cmap, int H, int O, const uint8_t *__restrict__ mask, float *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int h = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (h >= H) return;
    float m = -INFINITY;
    bool nan = false, any = false;
    for (int o = lane; o < O; o += 32) {
        if (mask[o]) {
            const float v = cmap[(size_t)h * O + o];
            any = true;
            nan |= (v != v);
            m = fmaxf(m, v);
        }
    }
    any = __any_sync(0xffffffffu, any);
    nan = __any_sync(0xffffffffu, nan);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
    if (lane == 0) out[h] = !any ? 0.0f : (nan ? __int_as_float(0x7fc00000) : m);
}

// axis = 0: out[o] = max over masked h (one thread per column, coalesced across o)
__global__ void __launch_bounds__(256)
    masked_colmax_kernel(const float *__restrict__ cmap, int H, int O, const uint8_t *__restrict__ mask, float *__restrict__ out) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= O) return;
    float m = -INFINITY;
    bool nan = false, any = false;
    for (int h = 0; h < H; ++h) {
        if (mask[h]) {
            const float v = cmap[(size_t)h * O + o];
            any = true;
            nan |= (v != v);
            m = fmaxf(m, v);
        }
    }
    out[o] = !any ? 0.0f : (nan ? __int_as_float(0x7fc00000) : m);
}

// K5c pass 1: per-vertex sums of the occupancy grid (one CTA per vertex)
__global__ void __launch_bounds__(256) occupancy_rowsum_kernel(const float *__restrict__ grids, long long V, float *__restrict__ sums,
                                                                unsigned *__restrict__ flags_t, int HW, unsigned *__restrict__ dense) {
    // flags_t (optional, 16-byte path only): [granule][HW] bit matrix, bit h of word (g, h / 32) = "granule g (32 consecutive float4 =
    // 512 B) of row h holds a non-zero bit pattern"; dense: bit h = "row h must be rewritten in full" (its sum is not a positive
    // finite number: 0 -> every voxel becomes NaN, negative -> -0.0). Pass 2 then touches only the flagged granules.
    const int h = blockIdx.x;
    const float *row = grids + (size_t)h * V;
    float s = 0.f;
    if ((V & 3) == 0 && (reinterpret_cast<uintptr_t>(grids) & 15) == 0) {  // 16-byte loads, four in flight per thread
        const float4 *r4 = reinterpret_cast<const float4 *>(row);
        const long long V4 = V >> 2;
        const int lane = threadIdx.x & 31;
        const unsigned hbit = 1u << (h & 31);
        unsigned *fl = flags_t ? flags_t + (h >> 5) : nullptr;
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
        long long ib = threadIdx.x - lane;  // the warp's first float4 of this step (warp-uniform): one granule per load
        for (; ib + 3 * 256 + 32 <= V4; ib += 4 * 256) {
            float4 a[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) a[u] = __ldcs(r4 + ib + u * 256 + lane);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                s4[u] += (a[u].x + a[u].y) + (a[u].z + a[u].w);
                if (fl) {
                    const unsigned nz = (__float_as_uint(a[u].x) | __float_as_uint(a[u].y)) | (__float_as_uint(a[u].z) | __float_as_uint(a[u].w));
                    if (__any_sync(0xffffffffu, nz != 0u) && lane == 0) atomicOr(fl + (size_t)((ib + u * 256) >> 5) * HW, hbit);
                }
            }
        }
        for (; ib < V4; ib += 256) {  // tail: the last granule of a row may be partial
            const bool in = ib + lane < V4;
            const float4 a = in ? __ldcs(r4 + ib + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
            s4[0] += (a.x + a.y) + (a.z + a.w);
            if (fl) {
                const unsigned nz = (__float_as_uint(a.x) | __float_as_uint(a.y)) | (__float_as_uint(a.z) | __float_as_uint(a.w));
                if (__any_sync(0xffffffffu, nz != 0u) && lane == 0) atomicOr(fl + (size_t)(ib >> 5) * HW, hbit);
            }
        }
        s = (s4[0] + s4[1]) + (s4[2] + s4[3]);
    } else
    for (long long i = threadIdx.x; i < V; i += blockDim.x) s += row[i];
    s = warp_sum(s);
    __shared__ float part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += part[w];
        sums[h] = t;
        if (dense && !(t > 0.f && t < INFINITY)) atomicOr(dense + (h >> 5), 1u << (h & 31));
    }
}

// K5c pass 2: normalise in place and take the NaN-propagating max over the selected vertices.
// grid = (ceil(V/256), HSPLIT): each CTA covers a slab of vertices so that small grids (Sg = 30 -> 27 000 voxels) still
// fill the machine; slabs are merged with an integer atomicMax on the float bits — the normalised values are >= 0, for
// which the int order equals the float order and the canonical NaN (0/0 of a vertex that never hit) compares highest,
// i.e. exactly torch.max's NaN propagation. `field` must be zero-filled before the launch.
__global__ void __launch_bounds__(256)
    occupancy_norm_max_kernel(float *__restrict__ grids, int H, long long V, const float *__restrict__ sums,
                              const uint8_t *__restrict__ sel, float *__restrict__ field) {
    const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const int per = (H + gridDim.y - 1) / gridDim.y;
    const int h_lo = blockIdx.y * per, h_hi = min(H, h_lo + per);
    int m = 0;
    for (int h = h_lo; h < h_hi; ++h) {
        const size_t i = (size_t)h * V + v;
        const float x = __fdiv_rn(grids[i], sums[h]);
        grids[i] = x;
        if (!sel || sel[h]) m = max(m, __float_as_int(x) & 0x7fffffff);
    }
    if (m != 0) atomicMax(reinterpret_cast<int *>(field) + v, m);
}

// float4 form (V % 4 == 0): a thread owns four consecutive voxels, two vertex rows in flight
__global__ void __launch_bounds__(256)
    occupancy_norm_max4_kernel(float *__restrict__ grids, int H, long long V, const float *__restrict__ sums,
                               const uint8_t *__restrict__ sel, float *__restrict__ field) {
    const long long v4 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long V4 = V >> 2;
    if (v4 >= V4) return;
    const int per = (H + gridDim.y - 1) / gridDim.y;
    const int h_lo = blockIdx.y * per, h_hi = min(H, h_lo + per);
    int m[4] = {0, 0, 0, 0};
    float4 *g4 = reinterpret_cast<float4 *>(grids);
    auto fold = [&](float4 &a, float sm, bool use) {
        a.x = __fdiv_rn(a.x, sm); a.y = __fdiv_rn(a.y, sm); a.z = __fdiv_rn(a.z, sm); a.w = __fdiv_rn(a.w, sm);
        if (use) {
            m[0] = max(m[0], __float_as_int(a.x) & 0x7fffffff); m[1] = max(m[1], __float_as_int(a.y) & 0x7fffffff);
            m[2] = max(m[2], __float_as_int(a.z) & 0x7fffffff); m[3] = max(m[3], __float_as_int(a.w) & 0x7fffffff);
        }
    };
    int h = h_lo;
    for (; h + 1 < h_hi; h += 2) {
        const size_t i0 = (size_t)h * V4 + v4, i1 = i0 + V4;
        float4 a = __ldcs(g4 + i0), b = __ldcs(g4 + i1);
        fold(a, sums[h], !sel || sel[h]);
        fold(b, sums[h + 1], !sel || sel[h + 1]);
        __stcs(g4 + i0, a);
        __stcs(g4 + i1, b);
    }
    if (h < h_hi) {
        const size_t i0 = (size_t)h * V4 + v4;
        float4 a = __ldcs(g4 + i0);
        fold(a, sums[h], !sel || sel[h]);
        __stcs(g4 + i0, a);
    }
    int *f = reinterpret_cast<int *>(field) + 4 * v4;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (m[k] != 0) atomicMax(f + k, m[k]);
}

// K5c pass 2, sparse form (16-byte path): a vertex's hits cluster where the vertex moves, so most 512-byte granules of its row are
// untouched zeros, and 0 / sum = +0 rewrites what is already there and adds nothing to the max.  A warp owns granule g for a slab of
// rows and visits only the rows whose bit is set in pass 1's flag matrix (or in `dense`): read, normalise, write, fold into the
// max — bit-identical grids and field, HBM traffic = one full read (pass 1) + the occupied granules instead of two reads + a write.
constexpr int K5C_WARPS = 8;
__global__ void __launch_bounds__(K5C_WARPS * 32)
    occupancy_norm_max_sparse_kernel(float *__restrict__ grids, int H, long long V, const float *__restrict__ sums,
                                     const uint8_t *__restrict__ sel, const unsigned *__restrict__ flags_t, int HW,
                                     const unsigned *__restrict__ dense, long long G, float *__restrict__ field) {
    const int lane = threadIdx.x & 31;
    const long long g = (long long)blockIdx.x * K5C_WARPS + (threadIdx.x >> 5);
    if (g >= G) return;  // warp-uniform
    const long long V4 = V >> 2, v4 = g * 32 + lane;
    const bool in = v4 < V4;
    const int per = (HW + gridDim.y - 1) / gridDim.y;
    const int w_lo = blockIdx.y * per, w_hi = min(HW, w_lo + per);
    float4 *g4 = reinterpret_cast<float4 *>(grids);
    const unsigned *fl = flags_t + (size_t)g * HW;
    int m[4] = {0, 0, 0, 0};
    auto fold = [&](float4 &a, float sm, bool use) {
        const bool plain = plain_divisor(sm);   // warp-uniform (one row); ncu before: 237 instructions per 512-byte row visit, issue 74 %
        a.x = div_rn_zero_fast(a.x, sm, plain); a.y = div_rn_zero_fast(a.y, sm, plain);
        a.z = div_rn_zero_fast(a.z, sm, plain); a.w = div_rn_zero_fast(a.w, sm, plain);
        if (use) {
            m[0] = max(m[0], __float_as_int(a.x) & 0x7fffffff); m[1] = max(m[1], __float_as_int(a.y) & 0x7fffffff);
            m[2] = max(m[2], __float_as_int(a.z) & 0x7fffffff); m[3] = max(m[3], __float_as_int(a.w) & 0x7fffffff);
        }
    };
    for (int hw = w_lo; hw < w_hi; ++hw) {
        unsigned w = fl[hw] | dense[hw];  // warp-uniform
        while (w) {
            int hs[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {  // up to four rows in flight
                hs[u] = w ? hw * 32 + (__ffs(w) - 1) : -1;
                w &= w - 1;   // 0 stays 0
            }
            float4 a[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (hs[u] >= 0 && hs[u] < H && in) a[u] = __ldcs(g4 + (size_t)hs[u] * V4 + v4);
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (hs[u] >= 0 && hs[u] < H && in) {
                    fold(a[u], sums[hs[u]], !sel || sel[hs[u]]);
                    __stcs(g4 + (size_t)hs[u] * V4 + v4, a[u]);
                }
        }
    }
    if (in) {
        int *f = reinterpret_cast<int *>(field) + 4 * v4;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (m[k] != 0) atomicMax(f + k, m[k]);
    }
}

__global__ void mark_selected_kernel(const long long *__restrict__ idx, long long n, int H, uint8_t *__restrict__ sel) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        long long h = idx[i];
        if (h < 0) h += H;  // python-style negative index
        if (h >= 0 && h < H) sel[h] = 1;
    }
}

}  // namespace coma

extern "C" int coma_normalize_contact_readout_f32(float *P, int64_t HO, int64_t N, float eps, const float *w, const float *nom,
                                                  const float *denom, float *cmap, coma_stream_t stream) {
    using namespace coma;
    COMA_REQUIRE(P, "null pointer");
    COMA_REQUIRE(!cmap || (w && nom && denom), "w/nom/denom are required when cmap is requested");
    COMA_REQUIRE(HO > 0 && N > 0 && N < (int64_t)1 << 30, "bad sizes");
    const long long blocks = (HO + K5_WARPS - 1) / K5_WARPS;
    const unsigned grid = (unsigned)(blocks < kNumSM * 8 ? blocks : kNumSM * 8);
    if (N <= 256)
        normalize_contact_reg_kernel<<<grid, K5_WARPS * 32, 0, (cudaStream_t)stream>>>(P, HO, (int)N, eps, w, nom, denom, cmap);
    else
        normalize_contact_kernel<<<grid, K5_WARPS * 32, 0, (cudaStream_t)stream>>>(P, HO, (int)N, eps, w, nom, denom, cmap);
    return check_launch("normalize_contact_kernel");
}

extern "C" int coma_entropy_readout_weighted_f32(const float *P, int64_t HO, int64_t N, float n_bin, const float *weights,
                                                 float weight_sum, float *out, coma_stream_t stream) {
    using namespace coma;
    COMA_REQUIRE(P && out, "null pointer");
    COMA_REQUIRE(HO > 0 && N > 0 && N < (int64_t)1 << 30 && n_bin > 1.0f, "bad sizes");
    const long long blocks = ((HO + 1) / 2 + K5_WARPS - 1) / K5_WARPS;
    const unsigned grid = (unsigned)(blocks < kNumSM * 8 ? blocks : kNumSM * 8);
    const double lg = log2((double)n_bin);
    entropy_kernel<<<grid, K5_WARPS * 32, 0, (cudaStream_t)stream>>>(P, HO, (int)N, n_bin, (float)lg, (float)(1.0 / ((double)n_bin * lg)), weights,
                                                                     weights ? weight_sum : 1.0f, out);
    return check_launch("entropy_kernel");
}

extern "C" int coma_entropy_readout_f32(const float *P, int64_t HO, int64_t N, float n_bin, float *out, coma_stream_t stream) {
    return coma_entropy_readout_weighted_f32(P, HO, N, n_bin, nullptr, 1.0f, out, stream);
}

extern "C" int coma_significant_pairs(const float *count, int64_t H, int64_t O, float num, uint8_t *sig, uint8_t *any_o,
                                      uint8_t *any_h, coma_stream_t stream) {
    using namespace coma;
    COMA_REQUIRE(count, "null pointer");
    COMA_REQUIRE(H > 0 && O > 0 && H < (int64_t)1 << 31 && O < (int64_t)1 << 31 && H * O < (int64_t)1 << 38, "bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    if (any_o) cudaMemsetAsync(any_o, 0, (size_t)H, st);
    if (any_h) cudaMemsetAsync(any_h, 0, (size_t)O, st);
    significant_pairs_kernel<<<(unsigned)((H * O + 255) / 256), 256, 0, st>>>(count, (int)H, (int)O, num, sig, any_o, any_h);
    return check_launch("significant_pairs_kernel");
}

extern "C" int coma_masked_max_f32(const float *cmap, int64_t H, int64_t O, const uint8_t *mask, int axis, float *out,
                                   coma_stream_t stream) {
    using namespace coma;
    COMA_REQUIRE(cmap && mask && out, "null pointer");
    COMA_REQUIRE(H > 0 && O > 0 && H < (int64_t)1 << 31 && O < (int64_t)1 << 31, "bad sizes");
    COMA_REQUIRE(axis == 0 || axis == 1, "axis must be 0 or 1");
    cudaStream_t st = (cudaStream_t)stream;
    if (axis == 1) {
        masked_rowmax_kernel<<<(unsigned)((H + 7) / 8), 256, 0, st>>>(cmap, (int)H, (int)O, mask, out);
        return check_launch("masked_rowmax_kernel");
    }
    masked_colmax_kernel<<<(unsigned)((O + 255) / 256), 256, 0, st>>>(cmap, (int)H, (int)O, mask, out);
    return check_launch("masked_colmax_kernel");
}

extern "C" int coma_occupancy_readout_f32(float *grids, int64_t H, int64_t V, const int64_t *sel_idx, int64_t nsel, float *field,
                                          coma_stream_t stream) {
    using namespace coma;
    COMA_REQUIRE(grids && field, "null pointer");
    COMA_REQUIRE(H > 0 && V > 0 && H < (int64_t)1 << 24, "bad sizes");
    COMA_REQUIRE(!sel_idx || nsel >= 0, "bad selection");
    cudaStream_t st = (cudaStream_t)stream;
    static const char *const force = getenv("COMA_B200_K5C");  // A/B runs only ("dense" = round 1's two-read pass 2), read once per process
    const bool vec = (V & 3) == 0 && (reinterpret_cast<uintptr_t>(grids) & 15) == 0 && (reinterpret_cast<uintptr_t>(field) & 15) == 0;
    const bool sparse = vec && !(force && force[0] == 'd');
    // scratch: H row sums, H selection flags, and for the sparse form the [granules][HW] flag matrix + HW dense-row words —
    // stream-ordered allocation, freed on the same stream
    const int64_t HW = (H + 31) / 32, G = (V / 4 + 31) / 32;
    const size_t off_sel = sizeof(float) * (size_t)H;
    const size_t off_flags = (off_sel + (size_t)H + 15) / 16 * 16;
    const size_t flag_bytes = sparse ? sizeof(unsigned) * (size_t)HW * (size_t)(G + 1) : 0;
    char *scratch = nullptr;
    {   // keep freed scratch cached in the device's default pool: with the default release threshold (0) every synchronisation hands
        // the memory back to the driver and the next read-out pays for a fresh mapping inside its critical path
        static bool pool_set[64] = {false};
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64 && !pool_set[dev]) {
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                uint64_t keep = 256ull << 20;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
            pool_set[dev] = true;
            cudaGetLastError();
        }
    }
    cudaError_t e = cudaMallocAsync(&scratch, off_flags + flag_bytes, st);
    if (e != cudaSuccess) {
        set_error("coma_occupancy_readout_f32: scratch allocation failed: %s", cudaGetErrorString(e));
        return (int)e;
    }
    float *sums = reinterpret_cast<float *>(scratch);
    uint8_t *sel = nullptr;
    unsigned *dense = sparse ? reinterpret_cast<unsigned *>(scratch + off_flags) : nullptr;   // [HW]
    unsigned *flags_t = sparse ? dense + HW : nullptr;                                         // [G][HW]
    if (sparse) cudaMemsetAsync(dense, 0, flag_bytes, st);
    occupancy_rowsum_kernel<<<(unsigned)H, 256, 0, st>>>(grids, V, sums, flags_t, (int)HW, dense);
    int rc = check_launch("occupancy_rowsum_kernel");
    if (!rc && sel_idx) {
        sel = reinterpret_cast<uint8_t *>(scratch + off_sel);
        cudaMemsetAsync(sel, 0, (size_t)H, st);
        if (nsel > 0) {
            mark_selected_kernel<<<(unsigned)((nsel + 255) / 256), 256, 0, st>>>((const long long *)sel_idx, nsel, (int)H, sel);
            rc = check_launch("mark_selected_kernel");
        }
    }
    if (!rc) {
        cudaMemsetAsync(field, 0, sizeof(float) * (size_t)V, st);
        if (sparse) {
            const long long gb = (G + K5C_WARPS - 1) / K5C_WARPS;
            // one 32-row word per slab (blockIdx.y), slabs in launch order: every resident CTA then works on the same few rows, like
            // the lock-step row loop of the dense kernel. Measured on B200 (128^3, 1310 rows, tools/occ_ab.py): with 9 words per slab the
            // hot warps drift apart over the whole 11 GB (rows are 8 MB apart) and the pass takes 32 ms instead of < 1 ms.
            const long long hsplit = HW > 65535 ? 65535 : HW;
            occupancy_norm_max_sparse_kernel<<<dim3((unsigned)gb, (unsigned)hsplit), K5C_WARPS * 32, 0, st>>>(
                grids, (int)H, V, sums, sel, flags_t, (int)HW, dense, G, field);
        } else {
            const long long vb = ((vec ? V / 4 : V) + 255) / 256;
            long long hsplit = (8LL * kNumSM + vb - 1) / vb;
            hsplit = hsplit < 1 ? 1 : (hsplit > H ? H : (hsplit > 65535 ? 65535 : hsplit));
            if (vec)
                occupancy_norm_max4_kernel<<<dim3((unsigned)vb, (unsigned)hsplit), 256, 0, st>>>(grids, (int)H, V, sums, sel, field);
            else
                occupancy_norm_max_kernel<<<dim3((unsigned)vb, (unsigned)hsplit), 256, 0, st>>>(grids, (int)H, V, sums, sel, field);
        }
        rc = check_launch("occupancy_norm_max_kernel");
    }
    cudaFreeAsync(scratch, st);
    return rc;
}
