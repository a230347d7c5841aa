// K4 — per-vertex occupancy voxel counts (reference: utils/coma_occupancy.py:160-183, :272-295).
//
// The reference tests every one of the Sg^3 voxel centres against every vertex (99.6 % misses).  Here one warp owns one
// (vertex h, sample s) task, derives the index box that can contain hits (|centre - v| < thr per axis) and only tests
// those <= (2*tol+1)^3 candidates, with the reference's exact fp64 arithmetic, then adds 1.0f to each hit voxel.
//   * small grids (4*Sg^3 <= 200 KB, e.g. the preset Sg = 30): the CTA that owns vertex h keeps the vertex's whole grid
//     in shared memory (integer ATOMS), loops over all samples and flushes it to HBM once  -> bytes = 12*S*H + 8*H*Sg^3.
//   * large grids (e.g. 128^3): RED.ADD.F32 straight into the L2-resident slice of the vertex's grid; CTAs that share a
//     vertex are adjacent in launch order so only a few vertices' grids are hot at a time.
// Bit-exactness: d < thr is decided as  ((dx*dx + dy*dy) + dz*dz) < T  in fp64 with explicitly rounded ops, where T is
// the smallest double whose correctly rounded sqrt is >= thr (computed on the host).  Because IEEE sqrt is monotone,
// {x : sqrt_rn(x) < thr} == {x : x < T}, so the verdict equals the reference's sqrt-then-compare exactly, without
// paying for an fp64 square root per candidate.  Counts are integers in fp32 (exact below 2^24 samples).
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace coma {

constexpr int K4_WARPS = 8;
constexpr int K4_SPLIT = 16;  // CTAs per vertex on the global-atomic path

struct Box {
    int lo[3], n[3];
};

// Index range [lo, lo+n) per axis that can contain hits. Centres are assumed strictly increasing and uniformly spaced
// (load_voxelgrid); one voxel of slack on both sides absorbs every rounding effect.
__device__ __forceinline__ void axis_range(const double *c, int Sg, double v, double thr, int &lo, int &n) {
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

template <bool SMEM>
__global__ void __launch_bounds__(K4_WARPS * 32)
    occupancy_kernel(const float *__restrict__ hvc, int S, int H, const double *__restrict__ centers, int Sg, double thr,
                     double T, float *__restrict__ grids, int split) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sc = reinterpret_cast<double *>(smem_raw);           // [3][Sg] centres
    unsigned *sg = reinterpret_cast<unsigned *>(sc + 3 * (size_t)Sg);  // [Sg^3] integer hit counts (SMEM path only)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int h = SMEM ? blockIdx.x : blockIdx.x / split;
    const int part = SMEM ? 0 : blockIdx.x % split;
    const int nparts = SMEM ? 1 : split;
    const size_t V = (size_t)Sg * Sg * Sg;

    for (int i = threadIdx.x; i < 3 * Sg; i += blockDim.x) sc[i] = centers[i];
    if (SMEM)
        for (size_t i = threadIdx.x; i < V; i += blockDim.x) sg[i] = 0u;
    __syncthreads();

    float *dst = grids + (size_t)h * V;
    const double *cx = sc, *cy = sc + Sg, *cz = sc + 2 * Sg;
    for (int s = part * K4_WARPS + warp; s < S; s += nparts * K4_WARPS) {
        const float *vp = hvc + ((size_t)s * H + h) * 3;
        const double v0 = (double)vp[0], v1 = (double)vp[1], v2 = (double)vp[2];
        int ilo, ni, jlo, nj, klo, nk;
        axis_range(cx, Sg, v0, thr, ilo, ni);
        axis_range(cy, Sg, v1, thr, jlo, nj);
        axis_range(cz, Sg, v2, thr, klo, nk);
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
                if (sum < T) {
                    const size_t cell = ((size_t)i * Sg + j) * Sg + k;
                    if (SMEM) atomicAdd(sg + cell, 1u);  // native ATOMS.ADD (an fp32 shared atomic would be a CAS loop)
                    else atomicAdd(dst + cell, 1.0f);    // RED.E.ADD.F32
                }
            }
        }
    }
    if (SMEM) {
        __syncthreads();
        float *g = grids + (size_t)h * V;
        for (size_t i = threadIdx.x; i < V; i += blockDim.x) {
            const unsigned a = sg[i];
            if (a != 0u) g[i] += (float)a;
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
    COMA_REQUIRE(S >= 0 && H > 0 && Sg > 0 && Sg <= 2048, "bad sizes");
    if (S == 0) return 0;
    const double T = squared_threshold(thr);
    const size_t V = (size_t)Sg * Sg * Sg;
    const size_t smem_small = 3 * sizeof(double) * Sg + sizeof(float) * V;
    cudaStream_t st = (cudaStream_t)stream;
    static const char *const force = getenv("COMA_B200_OCC_PATH");  // experiments only ("smem" | "global"), read once per process
    // measured on B200 (tools/microbench.py, Sg = 30): RED.ADD into the L2-resident grid beats the shared-memory
    // histogram (5.0 vs 7.4 ms per 256 x 10475 vertex-samples), so the global path is the default for every size
    const bool use_smem = smem_small <= 200 * 1024 && (force && force[0] == 's');
    if (use_smem) {
        static bool attr_set[16] = {false};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev >= 0 && dev < 16 && !attr_set[dev]) {
            cudaError_t e = cudaFuncSetAttribute(occupancy_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            if (e != cudaSuccess) {
                set_error("cudaFuncSetAttribute(occupancy_kernel): %s", cudaGetErrorString(e));
                return (int)e;
            }
            attr_set[dev] = true;
        }
        occupancy_kernel<true><<<(unsigned)H, K4_WARPS * 32, smem_small, st>>>(hvc, (int)S, (int)H, centers, (int)Sg, thr, T,
                                                                             grids, 1);
    } else {
        COMA_REQUIRE(H * K4_SPLIT < (int64_t)1 << 31, "H too large");
        occupancy_kernel<false><<<(unsigned)(H * K4_SPLIT), K4_WARPS * 32, 3 * sizeof(double) * Sg, st>>>(
            hvc, (int)S, (int)H, centers, (int)Sg, thr, T, grids, K4_SPLIT);
    }
    return check_launch("occupancy_kernel");
}
