// K6 — per-vertex normals of a fixed-topology triangle mesh, batched over samples (the sample ingest of the ComA
// extraction, reference utils/coma.py:665-686: open3d `TriangleMesh.compute_vertex_normals()` on every fitted SMPL-X mesh,
// then `normalize_vectors_np(., eps)`). open3d sums the UN-normalised face normals (v1-v0) x (v2-v0) (area weighting) over
// the faces incident to a vertex and normalises; vertices without a finite direction get (0, 0, 1).
//
// The SMPL-X topology (10 475 vertices, 20 908 faces) is shared by all samples, so the host builds ONE corner list (CSR by
// vertex, ordered exactly like the numpy restatement accumulates: all faces having the vertex as corner 0 in face order,
// then corner 1, then corner 2) and a thread per (sample, vertex) gathers its faces in that order: no atomics, fp64 with
// explicitly rounded products / sums -> bit-identical to the oracle (oracle/oracle.py: vertex_normals).
// Bytes: 24*V read (L2-resident gathers) + 24*V written per sample; latency bound, ~6 incident faces per vertex.
#include <math.h>

#include "common.cuh"

namespace coma {

__global__ void __launch_bounds__(256)
    vertex_normals_kernel(const double *__restrict__ verts, long long S, int V, const int *__restrict__ faces,
                          const int *__restrict__ corner_off, const int *__restrict__ corner_face, double eps, double *__restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S * V) return;
    const int v = (int)(i % V);
    const double *vb = verts + (i / V) * (long long)V * 3;
    double nx = 0.0, ny = 0.0, nz = 0.0;
    for (int k = corner_off[v]; k < corner_off[v + 1]; ++k) {
        const int f = corner_face[k];
        const double *p0 = vb + 3LL * faces[3 * f], *p1 = vb + 3LL * faces[3 * f + 1], *p2 = vb + 3LL * faces[3 * f + 2];
        const double a0 = __dsub_rn(p1[0], p0[0]), a1 = __dsub_rn(p1[1], p0[1]), a2 = __dsub_rn(p1[2], p0[2]);
        const double b0 = __dsub_rn(p2[0], p0[0]), b1 = __dsub_rn(p2[1], p0[1]), b2 = __dsub_rn(p2[2], p0[2]);
        // np.cross: cp0 = a1*b2 - a2*b1, cp1 = a2*b0 - a0*b2, cp2 = a0*b1 - a1*b0, every product rounded on its own
        nx = __dadd_rn(nx, __dsub_rn(__dmul_rn(a1, b2), __dmul_rn(a2, b1)));
        ny = __dadd_rn(ny, __dsub_rn(__dmul_rn(a2, b0), __dmul_rn(a0, b2)));
        nz = __dadd_rn(nz, __dsub_rn(__dmul_rn(a0, b1), __dmul_rn(a1, b0)));
    }
    const double n = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(nx, nx), __dmul_rn(ny, ny)), __dmul_rn(nz, nz)));
    double ox, oy, oz;
    if (n > 0.0) {
        ox = __ddiv_rn(nx, n); oy = __ddiv_rn(ny, n); oz = __ddiv_rn(nz, n);
    } else {
        ox = nx; oy = ny; oz = nz;   // 0/1 or nan/1: decided by the finiteness test below
    }
    if (n == 0.0 || !(isfinite(ox) && isfinite(oy) && isfinite(oz))) {
        ox = 0.0; oy = 0.0; oz = 1.0;
    }
    if (eps >= 0.0) {  // normalize_vectors_np(v, eps): v / (||v|| + eps)   (utils/transformations.py:14-17)
        const double m = __dadd_rn(__dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(ox, ox), __dmul_rn(oy, oy)), __dmul_rn(oz, oz))), eps);
        ox = __ddiv_rn(ox, m); oy = __ddiv_rn(oy, m); oz = __ddiv_rn(oz, m);
    }
    out[3 * i] = ox;
    out[3 * i + 1] = oy;
    out[3 * i + 2] = oz;
}

}  // namespace coma

extern "C" int coma_vertex_normals_f64(const double *verts, int64_t S, int64_t V, const int32_t *faces, int64_t F,
                                       const int32_t *corner_off, const int32_t *corner_face, double eps, double *out,
                                       coma_stream_t stream) {
    using namespace coma;
    COMA_REQUIRE(verts && faces && corner_off && corner_face && out, "null pointer");
    COMA_REQUIRE(S > 0 && V > 0 && F > 0 && V < (1LL << 31) && F < (1LL << 29) && S * V < (1LL << 40), "bad sizes");
    const long long n = S * V;
    vertex_normals_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(verts, S, (int)V, faces, corner_off, corner_face, eps, out);
    return check_launch("vertex_normals_kernel");
}
