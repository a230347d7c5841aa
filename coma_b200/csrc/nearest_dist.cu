// K7 — nearest-neighbour distances between two point sets, forward and backward.
//
// Reference: `chamfer_distance` (src/application/optimize.py:155-165: torch.cdist both ways, row minima, two means; it sits inside
// the SMPL-X pose optimiser's Adam loop and is differentiated through) and `minimum_distance` (src/generation/optimize_depth.py:29-44:
// cdist, row minima, sort, mean of the k smallest). Both materialise the [NA, NB] distance matrix; all that is ever used is
//   dist[i] = min_j ||a_i - b_j||,   idx[i] = argmin_j,
// whose gradient touches one b per a:  d dist_i / d a_i = (a_i - b_idx) / dist_i = - d dist_i / d b_idx   (0 where dist_i = 0,
// like torch.cdist's backward).
// Forward: thread per a-point, the b-points stream through shared memory in tiles (broadcast LDS.128); the B axis is split
// over blockIdx.y so that small A sets still fill 148 SMs, partial results meet in a 64-bit atomicMin on the key
// (float bits of dist^2) << 32 | j  (dist^2 >= 0: the unsigned order of the bits is the numeric order; equal distances resolve
// to the lowest j, torch.min's first-occurrence rule). A finishing kernel unpacks and takes the IEEE square root.
// Squared distances are sums of exactly rounded differences ((dx^2+dy^2)+dz^2, no FMA), i.e. the values of
// torch.cdist(compute_mode="donot_use_mm_for_euclid_dist"); cdist's default matmul path (|a|^2+|b|^2-2ab) carries cancellation
// errors of ~1e-4 relative for nearby points, so this is the more accurate of the two.
// Backward: thread per a-point; grad_a written, grad_b accumulated with RED.ADD.F32 (a few thousand scattered atomics).
#include <math.h>

#include "common.cuh"

namespace coma {

constexpr int K7_THREADS = 256, K7_TILE = 1024;

__global__ void __launch_bounds__(K7_THREADS)
    nearest_dist_partial_kernel(const float *__restrict__ a, int NA, const float *__restrict__ b, int NB, int b_per_block,
                                unsigned long long *__restrict__ key) {
    __shared__ float4 sb[K7_TILE];
    const int i = blockIdx.x * K7_THREADS + threadIdx.x;
    const bool live = i < NA;
    const float ax = live ? a[3 * i] : 0.f, ay = live ? a[3 * i + 1] : 0.f, az = live ? a[3 * i + 2] : 0.f;
    const int j_begin = blockIdx.y * b_per_block, j_end = min(NB, j_begin + b_per_block);
    float best = INFINITY;
    int bj = 0x7fffffff;
    for (int j0 = j_begin; j0 < j_end; j0 += K7_TILE) {
        const int n = min(K7_TILE, j_end - j0);
        __syncthreads();
        for (int t = threadIdx.x; t < n; t += K7_THREADS) sb[t] = make_float4(b[3 * (j0 + t)], b[3 * (j0 + t) + 1], b[3 * (j0 + t) + 2], 0.f);
        __syncthreads();
#pragma unroll 4
        for (int t = 0; t < n; ++t) {
            const float4 q = sb[t];
            const float dx = __fsub_rn(ax, q.x), dy = __fsub_rn(ay, q.y), dz = __fsub_rn(az, q.z);
            const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            if (d < best) {   // strict: the first minimum of this (ascending) range wins
                best = d;
                bj = j0 + t;
            }
        }
    }
    if (live && bj != 0x7fffffff)
        atomicMin(key + i, ((unsigned long long)__float_as_uint(best) << 32) | (unsigned)bj);
}

__global__ void nearest_dist_finish_kernel(const unsigned long long *__restrict__ key, int NA, float *__restrict__ dist,
                                           int *__restrict__ idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NA) return;
    const unsigned long long k = key[i];
    const bool none = k == 0xffffffffffffffffULL;   // every distance was NaN (or NB == 0)
    dist[i] = none ? __int_as_float(0x7fc00000) : __fsqrt_rn(__uint_as_float((unsigned)(k >> 32)));
    idx[i] = none ? 0 : (int)(unsigned)k;
}

__global__ void nearest_dist_backward_kernel(const float *__restrict__ a, int NA, const float *__restrict__ b, const int *__restrict__ idx,
                                             const float *__restrict__ dist, const float *__restrict__ grad_dist,
                                             float *__restrict__ grad_a, float *__restrict__ grad_b) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NA) return;
    const int j = idx[i];
    const float d = dist[i], g = grad_dist[i];
    const float s = d > 0.f ? g / d : 0.f;   // torch.cdist backward: zero gradient at zero distance
    const float gx = s * (a[3 * i] - b[3 * j]), gy = s * (a[3 * i + 1] - b[3 * j + 1]), gz = s * (a[3 * i + 2] - b[3 * j + 2]);
    if (grad_a) {
        grad_a[3 * i] = gx;
        grad_a[3 * i + 1] = gy;
        grad_a[3 * i + 2] = gz;
    }
    if (grad_b) {
        atomicAdd(grad_b + 3 * j, -gx);
        atomicAdd(grad_b + 3 * j + 1, -gy);
        atomicAdd(grad_b + 3 * j + 2, -gz);
    }
}

}  // namespace coma

extern "C" int coma_nearest_distance_f32(const float *a, int64_t NA, const float *b, int64_t NB, float *dist, int32_t *idx,
                                         uint64_t *scratch, coma_stream_t stream) {
    using namespace coma;
    COMA_REQUIRE(a && b && dist && idx && scratch, "null pointer");
    COMA_REQUIRE(NA >= 0 && NB > 0 && NA < (int64_t)1 << 30 && NB < (int64_t)1 << 31, "bad sizes");
    if (NA == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(scratch, 0xff, sizeof(uint64_t) * (size_t)NA, st);
    if (e != cudaSuccess) {
        set_error("coma_nearest_distance_f32: memset failed: %s", cudaGetErrorString(e));
        return (int)e;
    }
    const unsigned gx = (unsigned)((NA + K7_THREADS - 1) / K7_THREADS);
    // split B so that ~4 CTAs per SM exist, in whole tiles
    int64_t want = (4LL * kNumSM + gx - 1) / gx;
    const int64_t tiles = (NB + K7_TILE - 1) / K7_TILE;
    if (want > tiles) want = tiles;
    if (want < 1) want = 1;
    if (want > 65535) want = 65535;
    const int64_t per = ((tiles + want - 1) / want) * K7_TILE;
    const unsigned gy = (unsigned)((NB + per - 1) / per);
    nearest_dist_partial_kernel<<<dim3(gx, gy), K7_THREADS, 0, st>>>(a, (int)NA, b, (int)NB, (int)per, (unsigned long long *)scratch);
    if (int rc = check_launch("nearest_dist_partial_kernel")) return rc;
    nearest_dist_finish_kernel<<<gx, K7_THREADS, 0, st>>>((const unsigned long long *)scratch, (int)NA, dist, idx);
    return check_launch("nearest_dist_finish_kernel");
}

extern "C" int coma_nearest_distance_backward_f32(const float *a, int64_t NA, const float *b, int64_t NB, const int32_t *idx,
                                                  const float *dist, const float *grad_dist, float *grad_a, float *grad_b,
                                                  coma_stream_t stream) {
    using namespace coma;
    COMA_REQUIRE(a && b && idx && dist && grad_dist, "null pointer");
    COMA_REQUIRE(grad_a || grad_b, "at least one of grad_a / grad_b is needed");
    COMA_REQUIRE(NA >= 0 && NB > 0 && NA < (int64_t)1 << 30, "bad sizes");
    if (NA == 0) return 0;
    nearest_dist_backward_kernel<<<(unsigned)((NA + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a, (int)NA, b, idx, dist, grad_dist, grad_a, grad_b);
    return check_launch("nearest_dist_backward_kernel");
}
