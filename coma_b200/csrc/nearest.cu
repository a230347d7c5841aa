// K1 — nearest mesh vertex of each sampled point (reference: utils/coma.py:88-91, utils/coma_occupancy.py:70-74).
//
// One CTA per sampled point; the 256 threads stride over the V mesh vertices keeping (min squared distance, index) and
// the CTA reduces with warp shuffles. fp64 with explicitly rounded products and sums in the reference order
// ((x+y)+z, diff = point - vertex); ties resolve to the lowest vertex index exactly like np.argmin.
// fp64-ALU bound: 8 flops per (vertex, point) pair, 24*(N+V) bytes.
#include <math.h>

#include "common.cuh"

namespace coma {

__device__ __forceinline__ void argmin_combine(double &d, long long &i, double d2, long long i2) {
    if (d2 < d || (d2 == d && i2 < i)) {
        d = d2;
        i = i2;
    }
}

__global__ void __launch_bounds__(256)
    nearest_vertex_kernel(const double *__restrict__ pts, const double *__restrict__ verts, long long V,
                          long long *__restrict__ out) {
    const long long n = blockIdx.x;
    const double p0 = pts[3 * n], p1 = pts[3 * n + 1], p2 = pts[3 * n + 2];
    double best = INFINITY;
    long long bi = 0x7fffffffffffffffLL;
    for (long long v = threadIdx.x; v < V; v += blockDim.x) {
        const double dx = __dsub_rn(p0, verts[3 * v]), dy = __dsub_rn(p1, verts[3 * v + 1]), dz = __dsub_rn(p2, verts[3 * v + 2]);
        const double d = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
        if (d < best) {  // v increases per thread, so strict '<' keeps this thread's first minimum
            best = d;
            bi = v;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double d2 = __shfl_xor_sync(0xffffffffu, best, o);
        const long long i2 = __shfl_xor_sync(0xffffffffu, bi, o);
        argmin_combine(best, bi, d2, i2);
    }
    __shared__ double sd[8];
    __shared__ long long si[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        sd[warp] = best;
        si[warp] = bi;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) argmin_combine(best, bi, sd[w], si[w]);
        // all-NaN / all-inf rows: np.argmin returns 0 for an all-inf row
        out[n] = (bi == 0x7fffffffffffffffLL) ? 0 : bi;
    }
}

}  // namespace coma

extern "C" int coma_nearest_vertex_f64(const double *pts, int64_t N, const double *verts, int64_t V, int64_t *out_idx,
                                       coma_stream_t stream) {
    using namespace coma;
    COMA_REQUIRE(pts && verts && out_idx, "null pointer");
    COMA_REQUIRE(N >= 0 && V > 0 && N < (int64_t)1 << 31, "bad sizes");
    if (N == 0) return 0;
    nearest_vertex_kernel<<<(unsigned)N, 256, 0, (cudaStream_t)stream>>>(pts, verts, (long long)V, (long long *)out_idx);
    return check_launch("nearest_vertex_kernel");
}
