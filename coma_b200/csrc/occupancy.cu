// K4 — per-vertex occupancy voxel counts (reference: utils/coma_occupancy.py:160-183, :272-295).
//
// The reference tests every one of the Sg^3 voxel centres against every vertex (99.6 % misses).  Here one warp owns one
// (vertex h, sample s) task, derives the index box that can contain hits (|centre - v| < thr per axis) and only tests
// those candidates, with the reference's exact fp64 arithmetic, then adds 1.0f to each hit voxel with RED.ADD.F32 straight
// into the vertex's grid (L2); CTAs that share a vertex are adjacent in launch order so only a few vertices' grids are hot
// at a time.
//
// Round-2 form (`occupancy_kernel`): a lane owns two (j, k) COLUMNS of the box and keeps their (cy[j]-v1)^2, (cz[k]-v2)^2
// and cell offset in registers; the warp then walks the <= 2*tol+2 x-planes, where a candidate costs two DADDs, one
// compare and the RED — the reference's sum ((x+y)+z) needs nothing else once y and z are known.  The box is tight
// (ceil / floor of the linear index estimate with 1/64 voxel of margin instead of a whole voxel of slack on both sides:
// 8 x 8 instead of 10 x 10 columns at scale_tolerance 3), and the per-axis reciprocal spacing is computed once per CTA.
// Lanes 0..2 derive one axis range each.
// ~300 warp-instructions per task against ~1300 for round 1's candidate-per-lane loop (`occupancy_kernel_v1`, kept behind
// COMA_B200_OCC_PATH=v1 for A/B runs): the kernel moves from issue-bound towards the RED rate (1.29 clk per lane-RED per SM).
//
// Bit-exactness: d < thr is decided as  ((dx*dx + dy*dy) + dz*dz) < T  in fp64 with explicitly rounded ops, where T is
// the smallest double whose correctly rounded sqrt is >= thr (computed on the host).  Because IEEE sqrt is monotone,
// {x : sqrt_rn(x) < thr} == {x : x < T}, so the verdict equals the reference's sqrt-then-compare exactly, without
// paying for an fp64 square root per candidate.  Counts are integers in fp32 (exact below 2^24 samples).  A column with
// dy^2 >= T or dz^2 >= T can never hit (adding non-negative terms and rounding are monotone), so it is dropped up front.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace coma {

constexpr int K4_WARPS = 8;
constexpr int K4_SPLIT = 16;  // CTAs per vertex
#ifndef K4_MIN_CTAS
#define K4_MIN_CTAS 5  // <= 51 registers: 40 resident warps per SM for an issue-bound kernel with ~3 stall cycles per instruction
#endif

// ---- round-1 kernel (A/B reference): one candidate per lane and step, one voxel of slack around the box ----------------
__device__ __forceinline__ void axis_range_v1(const double *c, int Sg, double v, double thr, int &lo, int &n) {
    const double c0 = c[0];
    const double inv = (Sg > 1) ? (double)(Sg - 1) / (c[Sg - 1] - c0) : 0.0;
    double flo = floor((v - thr - c0) * inv) - 1.0, fhi = ceil((v + thr - c0) * inv) + 1.0;
    // NaN / inf vertices produce an empty or clamped box; the fp64 test rejects them anyway
    int ilo = (flo > -1e9) ? ((flo < 1e9) ? (int)flo : Sg) : 0;
    int ihi = (fhi > -1e9) ? ((fhi < 1e9) ? (int)fhi : Sg - 1) : -1;
    ilo = max(ilo, 0);
    ihi = min(ihi, Sg - 1);
    lo = ilo;
    n = max(ihi - ilo + 1, 0);
}

__global__ void __launch_bounds__(K4_WARPS * 32)
    occupancy_kernel_v1(const float *__restrict__ hvc, int S, int H, const double *__restrict__ centers, int Sg, double thr,
                        double T, float *__restrict__ grids, int split) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sc = reinterpret_cast<double *>(smem_raw);  // [3][Sg] centres
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int h = blockIdx.x / split, part = blockIdx.x % split;
    const size_t V = (size_t)Sg * Sg * Sg;
    for (int i = threadIdx.x; i < 3 * Sg; i += blockDim.x) sc[i] = centers[i];
    __syncthreads();
    float *dst = grids + (size_t)h * V;
    const double *cx = sc, *cy = sc + Sg, *cz = sc + 2 * Sg;
    for (int s = part * K4_WARPS + warp; s < S; s += split * K4_WARPS) {
        const float *vp = hvc + ((size_t)s * H + h) * 3;
        const double v0 = (double)vp[0], v1 = (double)vp[1], v2 = (double)vp[2];
        int ilo, ni, jlo, nj, klo, nk;
        axis_range_v1(cx, Sg, v0, thr, ilo, ni);
        axis_range_v1(cy, Sg, v1, thr, jlo, nj);
        axis_range_v1(cz, Sg, v2, thr, klo, nk);
        const int plane = nj * nk;
        const float rnk = 1.0f / (float)max(nk, 1);
        for (int ii = 0; ii < ni; ++ii) {
            const int i = ilo + ii;
            const double dx = __dsub_rn(cx[i], v0), xx = __dmul_rn(dx, dx);
            if (!(xx < T)) continue;  // warp-uniform: the whole plane is out of range
            for (int c = lane; c < plane; c += 32) {
                const int jj = __float2int_rz(((float)c + 0.5f) * rnk);
                const int j = jlo + jj, k = klo + (c - jj * nk);
                const double dy = __dsub_rn(cy[j], v1), dz = __dsub_rn(cz[k], v2);
                const double sum = __dadd_rn(__dadd_rn(xx, __dmul_rn(dy, dy)), __dmul_rn(dz, dz));  // (x+y)+z
                if (sum < T) atomicAdd(dst + ((size_t)i * Sg + j) * Sg + k, 1.0f);  // RED.E.ADD.F32
            }
        }
    }
}

// ---- round-2 kernel ----------------------------------------------------------------------------------------------------
// Index range [lo, lo+n) per axis that can contain hits: every hit has v-thr < c[i] < v+thr.  Centres are strictly
// increasing and uniform to within 1/64 of their spacing (load_voxelgrid's fp32 middle term is off by < 4e-4 of a voxel
// at Sg = 2048), so with x = (v -+ thr - c0) * inv the hits lie in [ceil(x_lo - 1/64), floor(x_hi + 1/64)].
__device__ __forceinline__ void axis_range(double c0, double inv, int Sg, double v, double thr, int &lo, int &n) {
    const double m = 1.0 / 64.0;
    const double flo = ceil((v - thr - c0) * inv - m), fhi = floor((v + thr - c0) * inv + m);
    // NaN / inf vertices produce an empty or clamped box; the fp64 test rejects them anyway
    int ilo = (flo > -1e9) ? ((flo < 1e9) ? (int)flo : Sg) : 0;
    int ihi = (fhi > -1e9) ? ((fhi < 1e9) ? (int)fhi : Sg - 1) : -1;
    ilo = max(ilo, 0);
    ihi = min(ihi, Sg - 1);
    lo = ilo;
    n = max(ihi - ilo + 1, 0);
}

// c -> (c / nk, c % nk) for 0 <= c < 2^24 (a plane has <= 2040^2 columns): float reciprocal estimate, corrected by at most one
__device__ __forceinline__ void split_column(int c, int nk, float rnk, int &j, int &k) {
    j = __float2int_rz(((float)c + 0.5f) * rnk);
    k = c - j * nk;
    if (k < 0) { j -= 1; k += nk; }
    else if (k >= nk) { j += 1; k -= nk; }
}

// grids[cell] += 1.0f if `hit`: RED.E.ADD.F32 with the constant as an immediate. (Written as a predicated `red`; ptxas still emits a
// short forward branch around every RED — SASS checked — so the per-plane cost that is left is this control flow.)
__device__ __forceinline__ void red_add_one_if(float *p, bool hit) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t@q red.global.add.f32 [%0], 0f3F800000;\n\t}" ::"l"(p), "r"((int)hit) : "memory");
}

__global__ void __launch_bounds__(K4_WARPS * 32, K4_MIN_CTAS)
    occupancy_kernel(const float *__restrict__ hvc, int S, int H, const double *__restrict__ centers, int Sg, double thr,
                     double T, float *__restrict__ grids, int split) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sc = reinterpret_cast<double *>(smem_raw);  // [3][Sg] centres
    __shared__ double s_c0[3], s_inv[3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int h = blockIdx.x / split, part = blockIdx.x % split;
    const size_t V = (size_t)Sg * Sg * Sg;
    const size_t plane_stride = (size_t)Sg * Sg;

    for (int i = threadIdx.x; i < 3 * Sg; i += blockDim.x) sc[i] = centers[i];
    if (threadIdx.x < 3) {
        const double c0 = centers[threadIdx.x * Sg];
        s_c0[threadIdx.x] = c0;
        s_inv[threadIdx.x] = (Sg > 1) ? (double)(Sg - 1) / (centers[threadIdx.x * Sg + Sg - 1] - c0) : 0.0;
    }
    __syncthreads();

    float *dst = grids + (size_t)h * V;
    const double *cx = sc, *cy = sc + Sg, *cz = sc + 2 * Sg;
    const int ax = lane < 3 ? lane : 0;  // lanes 0..2 derive the index range of one axis each (the others repeat axis 0)
    const double c0_ax = s_c0[ax], inv_ax = s_inv[ax];
    for (int s = part * K4_WARPS + warp; s < S; s += split * K4_WARPS) {
        const float *vp = hvc + ((size_t)s * H + h) * 3;
        const double v0 = (double)vp[0], v1 = (double)vp[1], v2 = (double)vp[2];
        int lo_ax, n_ax;
        axis_range(c0_ax, inv_ax, Sg, ax == 0 ? v0 : (ax == 1 ? v1 : v2), thr, lo_ax, n_ax);
        const int ilo = __shfl_sync(0xffffffffu, lo_ax, 0), ni = __shfl_sync(0xffffffffu, n_ax, 0);
        const int jlo = __shfl_sync(0xffffffffu, lo_ax, 1), nj = __shfl_sync(0xffffffffu, n_ax, 1);
        const int klo = __shfl_sync(0xffffffffu, lo_ax, 2), nk = __shfl_sync(0xffffffffu, n_ax, 2);
        const int plane = nj * nk;
        if (ni <= 0 || plane <= 0) continue;  // warp-uniform
        const float rnk = 1.0f / (float)nk;
        for (int cb = 0; cb < plane; cb += 64) {
            // two (j, k) columns per lane: consecutive lanes take consecutive k (contiguous cells)
            const int ca = cb + lane, cc = ca + 32;
            bool ok_a = ca < plane, ok_b = cc < plane;
            int ja, ka, jb, kb;
            split_column(ok_a ? ca : 0, nk, rnk, ja, ka);
            split_column(ok_b ? cc : 0, nk, rnk, jb, kb);
            const double dya = __dsub_rn(cy[jlo + ja], v1), dza = __dsub_rn(cz[klo + ka], v2);
            const double dyb = __dsub_rn(cy[jlo + jb], v1), dzb = __dsub_rn(cz[klo + kb], v2);
            const double yya = __dmul_rn(dya, dya), zza = __dmul_rn(dza, dza);
            const double yyb = __dmul_rn(dyb, dyb), zzb = __dmul_rn(dzb, dzb);
            ok_a = ok_a && (yya < T) && (zza < T);
            ok_b = ok_b && (yyb < T) && (zzb < T);
            if (!__any_sync(0xffffffffu, ok_a || ok_b)) continue;
            float *pa = dst + (size_t)ilo * plane_stride + (size_t)(jlo + ja) * Sg + (klo + ka);
            float *pb = dst + (size_t)ilo * plane_stride + (size_t)(jlo + jb) * Sg + (klo + kb);
            for (int ii = 0; ii < ni; ++ii, pa += plane_stride, pb += plane_stride) {
                const double dx = __dsub_rn(cx[ilo + ii], v0), xx = __dmul_rn(dx, dx);
                if (!(xx < T)) continue;  // warp-uniform: the whole plane is out of range
                red_add_one_if(pa, ok_a && __dadd_rn(__dadd_rn(xx, yya), zza) < T);  // (x+y)+z
                red_add_one_if(pb, ok_b && __dadd_rn(__dadd_rn(xx, yyb), zzb) < T);
            }
        }
    }
}

// Smallest double T with sqrt_rn(T) >= thr  (so that  sqrt_rn(x) < thr  <=>  x < T).
static double squared_threshold(double thr) {
    if (!(thr > 0.0)) return 0.0;  // d < thr is never true for thr <= 0 (d >= 0) or NaN
    if (isinf(thr)) return INFINITY;
    volatile double t = thr * thr;
    while (sqrt(t) >= thr) t = nextafter(t, 0.0);
    while (sqrt(t) < thr) t = nextafter(t, INFINITY);
    return t;
}

}  // namespace coma

extern "C" int coma_occupancy_accumulate(const float *hvc, int64_t S, int64_t H, const double *centers, int64_t Sg,
                                         double thr, float *grids, coma_stream_t stream) {
    using namespace coma;
    COMA_REQUIRE(hvc && centers && grids, "null pointer");
    COMA_REQUIRE(S >= 0 && H > 0 && Sg > 0 && Sg <= 2040, "bad sizes (Sg <= 2040: the per-axis centres live in 48 KB of shared memory)");
    if (S == 0) return 0;
    const double T = squared_threshold(thr);
    cudaStream_t st = (cudaStream_t)stream;
    static const char *const force = getenv("COMA_B200_OCC_PATH");  // A/B runs only ("v1" = round 1's kernel), read once per process
    COMA_REQUIRE(H * K4_SPLIT < (int64_t)1 << 31, "H too large");
    const unsigned grid = (unsigned)(H * K4_SPLIT);
    const size_t smem = 3 * sizeof(double) * Sg;
    if (force && force[0] == 'v' && force[1] == '1')
        occupancy_kernel_v1<<<grid, K4_WARPS * 32, smem, st>>>(hvc, (int)S, (int)H, centers, (int)Sg, thr, T, grids, K4_SPLIT);
    else
        occupancy_kernel<<<grid, K4_WARPS * 32, smem, st>>>(hvc, (int)S, (int)H, centers, (int)Sg, thr, T, grids, K4_SPLIT);
    return check_launch("occupancy_kernel");
}
