/*
 * coma_b200.h — C ABI of libcoma_b200.so (hand-written sm_100a CUDA kernels for snuvclab/coma's hot paths).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`; buffers are caller-owned, row-major,
 *     contiguous; nothing is allocated, retained or freed by the library.
 *   - `stream` is a cudaStream_t (as void*); every call only enqueues work on it and returns.
 *   - return value: 0 on success, otherwise a cudaError_t value (launch/config failure) or a negative COMA_E_* code;
 *     coma_b200_last_error() gives a thread-local human-readable message.
 *   - accumulating entry points ADD into their output buffers (they mirror the reference's in-place `+=`), so
 *     repeated calls over successive sample batches are equivalent to one call over the concatenation.
 *
 * Each entry point names the reference code (snuvclab/coma @ c89e2d1, paths relative to its root) it replaces.
 */
#ifndef COMA_B200_H
#define COMA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define COMA_E_BADARG (-1)   /* null pointer / non-positive size / unsupported parameter */
#define COMA_E_NODEVICE (-2) /* no sm_100 device / driver */

typedef void *coma_stream_t; /* cudaStream_t */

#if defined(__GNUC__)
#define COMA_API __attribute__((visibility("default")))
#else
#define COMA_API
#endif

COMA_API int coma_b200_version(void);
COMA_API const char *coma_b200_last_error(void);
/* Number of kernel launches enqueued by this library in the calling process (bench.py's `gpu_launches`). */
COMA_API int64_t coma_b200_launch_count(void);

/* Host-side staging helper (no device work, callable without a GPU): dst[i,r,c] = (float)(src[i][(row0+r)*3+c] - (sub ? sub[i][c] : 0))
 * for n samples x `rows` rows of 3 doubles — the fp64 -> fp32 rounding of the reference's to_np_torch_recursive (utils/misc.py:47-54)
 * and the fp64 subtraction of utils/coma_occupancy.py:287, written straight into a (pinned) fp32 staging buffer in one call per chunk.
 * equal_to (or NULL): 3 doubles every sub[i] is compared with (the reference's "same object in every sample" invariant,
 * utils/coma_occupancy.py:277-284); *first_mismatch = first differing sample or -1. */
COMA_API int coma_host_rows_equal_f64(const double *const *rows, int64_t n, const double *ref, int64_t count, int64_t *first_mismatch);
COMA_API int coma_host_stage_rows_f64_f32(const double *const *src, const double *const *sub, int64_t n, int64_t row0, int64_t rows,
                                          float *dst, const double *equal_to, int64_t *first_mismatch);
/* Name of the kernel (variant) the calling thread's most recent successful entry point enqueued — lets the parity tests
 * assert WHICH kernel a shape was routed to (e.g. the S <= 4 streaming form of K2). Static storage, never NULL. */
COMA_API const char *coma_b200_last_kernel(void);

/* ---- K6: per-vertex normals of a fixed-topology mesh, batched over samples (sample ingest) -------------------
 * Replaces open3d TriangleMesh.compute_vertex_normals() + normalize_vectors_np(., eps) on every fitted SMPL-X mesh
 * (utils/coma.py:665-686): area-weighted sum of the face normals (v1-v0)x(v2-v0) over incident faces, normalised, (0,0,1) where
 * no finite direction exists; eps >= 0 additionally applies v/(||v||+eps), eps < 0 skips it.
 * verts [S,V,3] f64, faces [F,3] i32; corner_off [V+1] / corner_face [3F] = CSR of the faces incident to each vertex ordered
 * (corner index, face index) — the accumulation order of the numpy restatement, which makes the result bit-identical. */
COMA_API int coma_vertex_normals_f64(const double *verts, int64_t S, int64_t V, const int32_t *faces, int64_t F,
                                     const int32_t *corner_off, const int32_t *corner_face, double eps, double *out,
                                     coma_stream_t stream);

/* ---- K1: nearest mesh vertex of each sampled point --------------------------------------------------------------
 * Replaces utils/coma.py:88-91 (dup. utils/coma_occupancy.py:70-74): idx[n] = argmin_v ((p0-v0)^2+(p1-v1)^2)+(p2-v2)^2
 * evaluated in fp64 with separately rounded products/sums; first minimum wins (np.argmin). Bit-exact.
 * pts [N,3] f64, verts [V,3] f64 -> out_idx [N] i64. */
COMA_API int coma_nearest_vertex_f64(const double *pts, int64_t N, const double *verts, int64_t V, int64_t *out_idx,
                            coma_stream_t stream);

/* ---- K7: nearest-neighbour distances between two point sets, forward + backward ------------------------
 * Replaces the torch.cdist + row-min of `chamfer_distance` (src/application/optimize.py:155-165, differentiated inside the pose
 * optimiser) and `minimum_distance` (src/generation/optimize_depth.py:29-44) without the [NA,NB] matrix:
 *   dist[i] = min_j ||a_i - b_j||  (IEEE sqrt of the exactly accumulated (dx^2+dy^2)+dz^2),  idx[i] = the first arg-min.
 * a [NA,3], b [NB,3] f32; dist [NA] f32, idx [NA] i32; scratch: NA uint64 (contents irrelevant). */
COMA_API int coma_nearest_distance_f32(const float *a, int64_t NA, const float *b, int64_t NB, float *dist, int32_t *idx,
                                       uint64_t *scratch, coma_stream_t stream);
/* Backward of the above: grad_a[i] = grad_dist[i] * (a_i - b_idx[i]) / dist[i] (0 where dist = 0, like torch.cdist), written;
 * grad_b[idx[i]] -= the same, ACCUMULATED with atomics into a caller-zeroed [NB,3] buffer. Either gradient may be NULL. */
COMA_API int coma_nearest_distance_backward_f32(const float *a, int64_t NA, const float *b, int64_t NB, const int32_t *idx,
                                                const float *dist, const float *grad_dist, float *grad_a, float *grad_b,
                                                coma_stream_t stream);

/* ---- K2: pair distance -> contact count + proximity expectation -------------------------------------------------
 * Replaces ComA.aggregate_single_sample_for_contact, utils/coma.py:284-291 (+ negative_exp :116-119), for S samples:
 *   d = sqrt(((hx-ox)^2+(hy-oy)^2)+(hz-oz)^2) (fp32, separately rounded);  count[h,o] += (d < thres);
 *   nom[h,o] += exp(-d / grid_size).          (`denom` is the scalar number of samples: the caller keeps it.)
 * hv [S,H,3] f32, ov [S,O,3] f32; count, nom [H,O] f32 (accumulated in place). count is bit-exact. */
COMA_API int coma_pair_accumulate_f32(const float *hv, const float *ov, int64_t S, int64_t H, int64_t O, float thres,
                             float grid_size, float *count, float *nom, coma_stream_t stream);
/* Same, with the association of the 3-term sum chosen explicitly. The reference's `torch.sum(torch.square(.), dim=-1)` adds
 * ((x^2+y^2)+z^2) on the CPU but ((x^2+z^2)+y^2) on CUDA (measured on B200, torch 2.11: tools/probe_torch_cuda_semantics.py);
 * the two differ in the last bit of the squared distance for ~23 % of random pairs and hence in `count` for pairs within one
 * ulp of the threshold. TORCH_CUDA is bit-exact against the reference run with device="cuda" (its production setting,
 * src/coma/extract_coma.py:329), TORCH_CPU against device="cpu" (= coma_pair_accumulate_f32, the CPU-generated goldens). */
#define COMA_SUM_ORDER_TORCH_CPU 0
#define COMA_SUM_ORDER_TORCH_CUDA 1
COMA_API int coma_pair_accumulate_order_f32(const float *hv, const float *ov, int64_t S, int64_t H, int64_t O, float thres,
                                            float grid_size, int sum_order, float *count, float *nom, coma_stream_t stream);

/* ---- K3: relative-orientation soft histograms --------------------------------------------------------------------
 * Replaces canonicalize_a_wrt_b_to_p (utils/coma.py:123-172, both calls :295-309) + geodesic_gaussian_scores
 * (:102-112) + the two in-place adds (:312-323), for S samples:
 *   PH[h,o,n] += exp(-acos(clip(G[n] . canon(hn[h] | on[o])))^2 / sigma^2), PO likewise with canon(on[o] | hn[h]).
 * hn [S,H,3], on [S,O,3] f32 (un-normalised is fine); grid [N,3] f64 (ComA.canon_normal_grid);
 * p_host, sub_p_host: 3 floats each on the HOST (principle / sub-principle vectors); PH, PO [H,O,N] f32. */
COMA_API int coma_orient_accumulate_f32(const float *hn, const float *on, int64_t S, int64_t H, int64_t O, const double *grid,
                               int64_t N, double sigma, double eps, const float *p_host, const float *sub_p_host,
                               float *PH, float *PO, coma_stream_t stream);

/* Same accumulation, CONE-LIMITED: scores below 2^-drop_bits (bins further than sigma*sqrt(drop_bits ln 2) from the canonical
 * normal: ~70 % of the bins at sigma = 0.25, drop_bits = 32) are not evaluated. Every term this entry point drops or clamps is
 * < 2^-drop_bits, so after S samples each bin is within S * 2^-drop_bits ABSOLUTE of coma_orient_accumulate_f32's result; terms
 * inside the cone carry the same ~5e-6 relative evaluation error. drop_bits = 0 selects the dense kernel (bit-identical to
 * coma_orient_accumulate_f32); when the cone does not fit the kernel's polynomial domain (sigma > ~0.31 at 32 bits) or N > 256
 * the dense kernel runs as well. bin_perm: DEVICE int32 [32 * ceil(N/32)] from coma_orient_bin_patches (compact patches = fewer
 * evaluations) or NULL (bins grouped in index order; same results, less culling). sum_order (COMA_SUM_ORDER_*): the association of
 * the reference's 3-term torch.sum(dim=-1) inside normalisation / canonicalisation — only the last bits of the canonical normals
 * depend on it, except next to the antipodal singularity (1 + b.p -> 0) where the reference itself differs by ~1e-3 between its
 * CPU and CUDA runs. */
COMA_API int coma_orient_accumulate_cone_f32(const float *hn, const float *on, int64_t S, int64_t H, int64_t O, const double *grid,
                                             int64_t N, double sigma, double eps, const float *p_host, const float *sub_p_host,
                                             const int32_t *bin_perm, int drop_bits, int sum_order, float *PH, float *PO,
                                             coma_stream_t stream);
/* Same with a device workspace of >= 3*S*(H+O) floats (or NULL): the normals are normalised once per (sample, vertex) by a pre-pass into
 * the workspace instead of once per (pair, sample) inside the kernel — bit-identical results (the same correctly rounded operations). */
COMA_API int coma_orient_accumulate_cone_ws_f32(const float *hn, const float *on, int64_t S, int64_t H, int64_t O, const double *grid,
                                                int64_t N, double sigma, double eps, const float *p_host, const float *sub_p_host,
                                                const int32_t *bin_perm, int drop_bits, int sum_order, float *PH, float *PO,
                                                float *workspace, coma_stream_t stream);
/* HOST-only helper (no GPU work): groups the N <= 256 bin centres grid_host [N,3] f64 into ceil(N/32) patches of <= 32 compact
 * bins; perm_host [32 * ceil(N/32)] int32 receives the bin index of each (patch, lane) slot, -1 for empty slots. */
COMA_API int coma_orient_bin_patches(const double *grid_host, int64_t N, int32_t *perm_host);

/* Canonicalised normals only (utils/coma.py:123-172): out[i,j,:] = canon(a[i] | b[j]); a [A,3], b [B,3] f32. */
COMA_API int coma_canonicalize_f32(const float *a, int64_t A, const float *b, int64_t B, const float *p_host,
                          const float *sub_p_host, float eps, float *out, coma_stream_t stream);
COMA_API int coma_canonicalize_order_f32(const float *a, int64_t A, const float *b, int64_t B, const float *p_host,
                                         const float *sub_p_host, float eps, int sum_order, float *out, coma_stream_t stream);

/* ---- K4: per-vertex occupancy voxel counts ----------------------------------------------------------------------
 * Replaces ComA_Occupancy.aggregate_single_sample_for_occupancy, utils/coma_occupancy.py:289-295, for S samples:
 *   grids[h,i,j,k] += ( sqrt(((cx[i]-v0)^2+(cy[j]-v1)^2)+(cz[k]-v2)^2) < thr )  in fp64, v = (double)hvc[s,h,:].
 * hvc [S,H,3] f32 = fp32(human_verts - obj_verts[0]) (the host-side subtraction of :287-288);
 * centers [3,Sg] f64 per-axis voxel centres (load_voxelgrid :160-171): strictly increasing and uniformly spaced to within 1/64 of
 * the spacing (the kernel derives its candidate box from a linear index estimate; load_voxelgrid's centres are uniform to < 4e-4);
 * thr = voxel_size*scale_tolerance; grids [H,Sg,Sg,Sg] f32 accumulated in place, Sg <= 2040. Bit-exact (integer counts). */
COMA_API int coma_occupancy_accumulate(const float *hvc, int64_t S, int64_t H, const double *centers, int64_t Sg, double thr,
                              float *grids, coma_stream_t stream);

/* ---- K5: read-outs ---------------------------------------------------------------------------------------------------
 * K5a  normalize_prob_grid_for_normals (utils/coma.py:328-330) fused with compute_contact_map (:342-356):
 *      P[q,:] /= (sum_n P[q,:] + eps)  IN PLACE, then cmap[q] = (sum_n P[q,n]*w[n]) * nom[q]/denom[q].
 *      P [HO,N] f32, w [N] f32 ((1 - p.G[n])/2), nom/denom [HO] f32, cmap [HO] f32 (may be NULL: normalise only). */
COMA_API int coma_normalize_contact_readout_f32(float *P, int64_t HO, int64_t N, float eps, const float *w, const float *nom,
                                       const float *denom, float *cmap, coma_stream_t stream);

/* K5a' significant_contact_pairs (:369-382) + the two `any` reductions (:407,:421):
 *      sig[h,o] = count[h,o] >= num; any_o[h] = OR_o sig[h,o]; any_h[o] = OR_h sig[h,o]. u8 outputs. */
COMA_API int coma_significant_pairs(const float *count, int64_t H, int64_t O, float num, uint8_t *sig, uint8_t *any_o,
                           uint8_t *any_h, coma_stream_t stream);

/* K5a'' aggregate_contact_for_significant_pairs (:398-427): masked max of cmap [H,O].
 *      axis=1: out[h] = max_{o: mask[o]} cmap[h,o] (mask [O]);  axis=0: out[o] = max_{h: mask[h]} cmap[h,o] (mask [H]).
 *      If no mask bit is set the output is all zeros (the reference's fallback). NaN propagates like torch.max. */
COMA_API int coma_masked_max_f32(const float *cmap, int64_t H, int64_t O, const uint8_t *mask, int axis, float *out,
                        coma_stream_t stream);

/* K5b  compute_nonphysical_response_sphere (:455-463) on an already-normalised grid:
 *      q = rint(P*n_bin)/n_bin; out[q] = 1 + sum_n (q==0 ? 0 : q*log q) / log(n_bin). P [HO,N] f32 -> out [HO] f32. */
COMA_API int coma_entropy_readout_f32(const float *P, int64_t HO, int64_t N, float n_bin, float *out, coma_stream_t stream);
/* `_v2` of the same read-out (utils/coma.py:529-579): every term (q ln q / ln n_bin + 1) is weighted by weights[n] (the bin's
 * alignment G[n] . p, [N] f32 on the device); weight_sum = sum_n weights[n] (host). weights = NULL gives the unweighted form. */
COMA_API int coma_entropy_readout_weighted_f32(const float *P, int64_t HO, int64_t N, float n_bin, const float *weights,
                                               float weight_sum, float *out, coma_stream_t stream);

/* K5c  normalize_prob_grid_for_spatials + max over vertices (utils/coma_occupancy.py:297-312):
 *      grids[h,:] /= sum(grids[h,:]) IN PLACE (NaN where a vertex never hit, as in the reference), then
 *      field[v] = max_{h in sel} grids[h,v] (NaN-propagating). sel_idx [nsel] i64 vertex indices, or NULL for all H. */
COMA_API int coma_occupancy_readout_f32(float *grids, int64_t H, int64_t V, const int64_t *sel_idx, int64_t nsel, float *field,
                               coma_stream_t stream);

/* ==== HP-A: adaptive-mask inpainting loop (utils/adaptive_mask_inpainting.py:984-1157) ================================= */

/* ---- G1: fp16 tensor-core GEMM with fused epilogue (tcgen05 / TMEM / TMA) ------------------------------------------------
 * out[m,n] = act( sum_k A[m,k]*W[n,k] + bias[n] + residual[m,n] ),  fp16 operands, fp32 accumulation.
 * The dense contractions of the UNet / VAE the pipeline calls at utils/adaptive_mask_inpainting.py:1001-1007 (unet),
 * :680 (vae.encode), :1086,:1112 (vae.decode): linear layers, 1x1 convs, attention projections and the im2col form of
 * the 3x3 convs (diffusers UNet2DConditionModel / AutoencoderKL, not vendored in the reference).
 * A [M,K] fp16 row stride lda; W [N,K] fp16 row stride ldw (lda, ldw multiples of 8); bias [N] f32 or NULL;
 * residual [M,N] fp16 with row stride ldo or NULL; act: 0 identity, 1 SiLU, 2 quick-GELU x*sigmoid(1.702x) (CLIP MLP); out_f16 / out_f32 [M,N] row stride ldo
 * (either may be NULL, not both). */
/* General form: two batch dimensions (nb1 fastest) with independent element strides for A, W and the output, an
 * accumulator scale `alpha` (applied before the bias terms) and a second, per-row-group bias
 * `bias_rows[(m / rows_per_bias), n]` (the time-embedding term of a ResnetBlock). Batched attention products
 * (Q K^T per head, P V per head) are expressed with the strides; `residual` shares the output's layout. */
typedef struct coma_gemm_args {
    const void *A; int64_t lda, a_s1, a_s2;
    const void *W; int64_t ldw, w_s1, w_s2;
    void *out_f16; float *out_f32; int64_t ldo, o_s1, o_s2;
    const void *residual;
    const float *bias;
    const float *bias_rows; int64_t rows_per_bias, bias_rows_ld; /* row stride of bias_rows, 0 = N */
    int64_t M, N, K, nb1, nb2;
    float alpha;
    int act;
    float *workspace; int64_t workspace_elems; /* optional fp32 scratch (>= ksplit*M*N floats) enabling split-K on problems with
                                                * too few output tiles for 148 SMs; NULL = never split */
    int geglu; /* 1: W / bias rows are interleaved in blocks of 32 (32 value rows, their 32 gate rows, ...) and the epilogue writes
                * out[M, N/2] = value * gelu(gate) (diffusers GEGLU fused into its projection); needs N % 256 == 0 */
    /* LayerNorm folded into the projection that consumes it (diffusers BasicTransformerBlock: norm1 -> to_q/k/v, norm2 -> to_q, norm3 ->
     * ff): A is the UN-normalised input, W = W0 * gamma (column-wise), bias = W0 beta + b0, ln_c1[n] = sum_k W[n,k], ln_row_stats[m] =
     * (rstd_m, -rstd_m * mean_m) from coma_layernorm_stats_f16; the epilogue computes rstd * acc - rstd * mean * c1 + bias. Both NULL: off. */
    const float *ln_row_stats;
    const float *ln_c1;
    /* ... or without any statistics kernel: the GEMM that PRODUCES the LayerNorm's input leaves, per row and 32-column panel of its rounded
     * fp16 output, (sum, sumsq) in ln_partials_out [M, N/32] float2; the consumer passes that buffer as ln_partials_in (+ ln_c1, ln_eps;
     * K % 32 == 0) and forms (rstd, -rstd * mean) of its rows in the epilogue, adding the K/32 partials in panel order. */
    const float *ln_partials_in;
    float ln_eps;
    float *ln_partials_out;
} coma_gemm_args;
COMA_API int coma_gemm_f16_ex(const coma_gemm_args *args, coma_stream_t stream);
/* Host-only: the output-tile width (64 / 128 / 160 / 256) and split-K factor the scheduler picks for an M x N x K problem with
 * `workspace_elems` floats of split-K scratch (0 = splitting not allowed); conv_m_tiles > 0 overrides ceil(M / 128). */
COMA_API int coma_gemm_plan(int64_t M, int64_t N, int64_t K, int64_t workspace_elems, int conv_m_tiles, int *tile_n, int *ksplit);

/* 3x3 convolution, stride 1, zero padding 1, as an IMPLICIT GEMM: the A operand of every (tap, 64-channel) K-slab is a
 * shifted 128-pixel tile fetched by one 4-D TMA load (out-of-image rows/columns are hardware zero-filled), so no im2col
 * matrix exists. x [B,H,W,C] NHWC f16 (row stride ldx, C % 64 == 0, H and W must tile into 128-pixel blocks);
 * W [N, ldw >= 9C] f16 with K order (ky, kx, c); epilogue as in coma_gemm_f16_ex with rows_per_bias = H*W. */
COMA_API int coma_conv3x3_f16(const void *x, int64_t B, int64_t H, int64_t W, int64_t C, int64_t ldx, const void *Wt, int64_t ldw,
                              int64_t N, const float *bias, const float *bias_rows, int64_t bias_rows_ld, const void *residual, int act,
                              void *out_f16, float *out_f32, int64_t ldo, coma_stream_t stream);

/* Same, with an fp32 workspace that allows a deterministic split-K schedule (deep-K convolutions at 8x8 / 16x16). */
COMA_API int coma_conv3x3_f16_ws(const void *x, int64_t B, int64_t H, int64_t W, int64_t C, int64_t ldx, const void *Wt, int64_t ldw,
                                 int64_t N, const float *bias, const float *bias_rows, int64_t bias_rows_ld, const void *residual, int act,
                                 void *out_f16, float *out_f32, int64_t ldo, float *workspace, int64_t workspace_elems,
                                 coma_stream_t stream);

/* Strided form: stride 2 with pad 1 (UNet downsamplers) or pad 0 + one zero row / column after the image (VAE encoder
 * downsamplers, diffusers F.pad (0,1,0,1)); the shifted TMA tiles use an element stride of 2, still no im2col matrix.
 * Hin / Win are the INPUT extent; the output [B, Ho, Wo, N] must tile into 128-pixel blocks with >= 128 pixels per image. */
COMA_API int coma_conv3x3_strided_f16(const void *x, int64_t B, int64_t Hin, int64_t Win, int64_t C, int64_t ldx, int stride, int pad,
                                      const void *Wt, int64_t ldw, int64_t N, const float *bias, const float *bias_rows,
                                      int64_t bias_rows_ld, const void *residual, int act, void *out_f16, float *out_f32, int64_t ldo,
                                      float *workspace, int64_t workspace_elems, float *stats, int *stats_written,
                                      coma_stream_t stream);
/* stats (optional, [B*Ho*Wo/32, N, 2] f32): the epilogue also leaves, per 32-row block and output channel, the sum and the sum
 * of squares of the fp16 values it wrote, so the GroupNorm that consumes this tensor needs no pass over it
 * (coma_groupnorm_from_stats_f32). *stats_written tells whether this launch could provide them (fp16 TMA epilogue, no
 * split-K, M % 32 == 0); otherwise the caller falls back to coma_groupnorm_affine_f16. */
COMA_API int coma_groupnorm_from_stats_f32(const float *stats, int64_t B, int64_t HW, int64_t C, int G, float eps, const float *gamma,
                                           const float *beta, float *mean, float *rstd, float *scale, float *shift,
                                           coma_stream_t stream);
/* Same with partial sums per `rows_per_block` rows ([B*HW/rows_per_block, C, 2]): coma_conv3x3_halo_f16 leaves one row per 128-pixel tile. */
COMA_API int coma_groupnorm_from_stats_rb_f32(const float *stats, int64_t B, int64_t HW, int64_t rows_per_block, int64_t C, int G, float eps,
                                              const float *gamma, const float *beta, float *mean, float *rstd, float *scale, float *shift,
                                              coma_stream_t stream);

COMA_API int coma_gemm_f16_tn(const void *A, int64_t lda, const void *W, int64_t ldw, int64_t M, int64_t N, int64_t K,
                              const float *bias, const void *residual, int act, void *out_f16, float *out_f32, int64_t ldo,
                              coma_stream_t stream);

/* ---- C1: 3x3 convolution (stride 1, zero padding 1) with 1..4 output channels, fused with the GroupNorm affine + SiLU of its
 * input: out[b,y,x,n] = bias[n] + sum_{ky,kx,c} W[n,(ky,kx,c)] * act(x[b,y+ky-1,x+kx-1,c]*scale[b,c] + shift[b,c]) (zero outside the image).
 * Replaces decoder.conv_norm_out -> SiLU -> conv_out (128 -> 3) of AutoencoderKL.decode and conv_norm_out -> SiLU -> conv_out (320 -> 4) of
 * the UNet (utils/adaptive_mask_inpainting.py:1086,:1112,:1001) without the normalised intermediate tensor. CUDA cores, halo tiles in shared
 * memory. x [B,H,W,C] NHWC f16 (row stride ldx), C % 8 == 0; scale / shift [B,C] f32 or both NULL; Wt [Cout, ldw >= 9C] f16, K order
 * (ky,kx,c); out_f32 / out_f16 [B*H*W, ldo] (either may be NULL). */
COMA_API int coma_conv3x3_small_n_f16(const void *x, int64_t B, int64_t H, int64_t W, int64_t C, int64_t ldx, const float *scale,
                                      const float *shift, int act, const void *Wt, int64_t ldw, int64_t Cout, const float *bias,
                                      float *out_f32, void *out_f16, int64_t ldo, coma_stream_t stream);

/* ---- C2: 3x3 convolution (stride 1, pad 1) from activation HALO tiles with the GroupNorm affine + SiLU of the input fused in: the
 * `norm -> SiLU -> conv` of diffusers' ResnetBlock2D (UNet utils/adaptive_mask_inpainting.py:1001, VAE :680 / :1086 / :1112) in one
 * tcgen05 kernel. Per 16 x 8 pixel output tile and 64-channel block ONE TMA load brings the 18 x 10 halo; the nine taps are shifted
 * views of it (no nine-fold L2 -> SMEM traffic); transform warps apply act_in(x * scale[b,c] + shift[b,c]) in shared memory, so the
 * normalised tensor is never written to HBM. x [B,H,W,C] NHWC f16 (row stride ldx) — with up != 0 x is stored as [B,H/2,W/2,C] and read
 * as its nearest-neighbour x2 upsampling (diffusers Upsample2D: interpolate + conv, no 4x larger intermediate); scale / shift [B,C]
 * f32 or both NULL (input used as stored); Wt [N, ldw >= 9C] f16, K order (ky,kx,c); bias [N]; bias_rows [B, bias_rows_ld] per-sample rows (time embedding) or
 * NULL; residual [B*H*W, ldo] f16 or NULL; out [B*H*W, ldo] f16; stats [B*H*W/128, N, 2] f32 or NULL (GroupNorm partial sums of the
 * output, one row per 16 x 8 pixel tile: coma_groupnorm_from_stats_rb_f32 with rows_per_block = 128). Needs H % 16 == 0, W % 8 == 0, C % 64 == 0, N % 64 == 0 (coma_conv3x3_halo_supported). */
COMA_API int coma_conv3x3_halo_supported(int64_t B, int64_t H, int64_t W, int64_t C, int64_t N);
COMA_API int coma_conv3x3_halo_f16(const void *x, int64_t B, int64_t H, int64_t W, int64_t C, int64_t ldx, int up, const float *scale,
                                   const float *shift, int act_in, const void *Wt, int64_t ldw, int64_t N, const float *bias,
                                   const float *bias_rows, int64_t bias_rows_ld, const void *residual, int act_out, void *out_f16,
                                   int64_t ldo, float *stats, coma_stream_t stream);

/* ---- A1: fused multi-head attention forward (tcgen05: S = Q K^T and O = P V on the tensor cores, online softmax between
 * them, scores never leave the SM). out[b,s,h*d:(h+1)*d] = softmax(Q_h K_h^T * scale) V_h.
 * q [B,S,heads*d] (row stride ldq), k [B,L,heads*d] (ldk), vt = V^T [B,heads,d,Lp] (coma_transpose_heads_f16), all f16;
 * d % 8 == 0, d <= 192; out [B,S,heads*d] f16 (row stride ldo). Replaces the baddbmm + softmax + bmm of the reference's
 * attention (diffusers Attention under torch 1.13, reached from utils/adaptive_mask_inpainting.py:1001). */
COMA_API int coma_attention_fwd_f16(const void *q, const void *k, const void *vt, int64_t B, int64_t heads, int64_t S, int64_t L,
                                    int64_t d, int64_t ldq, int64_t ldk, int64_t Lp, float scale, void *out, int64_t ldo,
                                    coma_stream_t stream);
/* Same with an optional fp32 output: out_f32 [B,S,heads*d] f32 (row stride ldo, in elements) is written INSTEAD of the fp16 `out`
 * when non-NULL — the normalised TMEM accumulator without the final fp16 rounding (parity tests). */
COMA_API int coma_attention_fwd_ex_f16(const void *q, const void *k, const void *vt, int64_t B, int64_t heads, int64_t S, int64_t L,
                                       int64_t d, int64_t ldq, int64_t ldk, int64_t Lp, float scale, void *out, float *out_f32,
                                       int64_t ldo, coma_stream_t stream);

/* Same with V UNtransposed: v [B,L,heads*d] f16 (row stride ldv), e.g. a column slice of the fused q|k|v projection — the P V
 * product reads the [keys x d] tile as an MN-major tensor-core operand, so no V^T tensor and no coma_transpose_heads_f16 launch in
 * front of the attention. 64-key steps for every shape; out / out_f32 as in coma_attention_fwd_ex_f16. */
COMA_API int coma_attention_fwd_nt_f16(const void *q, const void *k, const void *v, int64_t B, int64_t heads, int64_t S, int64_t L,
                                       int64_t d, int64_t ldq, int64_t ldk, int64_t ldv, float scale, void *out, float *out_f32,
                                       int64_t ldo, coma_stream_t stream);

/* ---- U*: the non-contraction layers of the UNet / VAE, NHWC fp16 activations ([B, H*W, C], row stride ld*) -------------
 * (diffusers UNet2DConditionModel / AutoencoderKL layers reached from utils/adaptive_mask_inpainting.py:1001, :680, :1086) */

/* GroupNorm statistics + folded affine: scale[b,c] = rstd*gamma[c], shift[b,c] = beta[c] - mean*rstd*gamma[c] so that
 * gn(x) = x*scale + shift. ONE launch: per-chunk partial sums, the last CTA of each sample reduces them in index order
 * (fp64, bit-reproducible). workspace: coma_groupnorm_workspace_doubles(B, G) doubles (contents irrelevant); counters: B
 * unsigned ints that must be ZERO before the first call and are left zero by every call (per-stream persistent scratch).
 * mean / rstd ([B,G] f32) may be NULL. */
COMA_API int64_t coma_groupnorm_workspace_doubles(int64_t B, int G);
COMA_API int coma_groupnorm_affine_f16(const void *x, int64_t B, int64_t HW, int64_t C, int64_t ldx, int G, float eps,
                                       const float *gamma, const float *beta, double *workspace, unsigned *counters, float *mean,
                                       float *rstd, float *scale, float *shift, coma_stream_t stream);
COMA_API int coma_affine_act_f16(const void *x, int64_t B, int64_t HW, int64_t C, int64_t ldx, const float *scale,
                                 const float *shift, int act, void *y, int64_t ldy, coma_stream_t stream);
/* Nearest x2 upsampling fused with the affine + activation (input of the Upsample2D convolutions): y [B,2H,2W,C]. */
COMA_API int coma_upsample2x_affine_act_f16(const void *x, int64_t B, int64_t H, int64_t W, int64_t C, int64_t ldx,
                                            const float *scale, const float *shift, int act, void *y, int64_t ldy,
                                            coma_stream_t stream);
/* im2col for 3x3 convolutions: out[(b,oy,ox), (ky*3+kx)*C + c] = act(x[b,iy,ix,c]*scale+shift) (0 outside the image),
 * iy = oy*stride + ky - pad. stride 1|2; pad 1 (symmetric) | 0 (VAE encoder's bottom/right-only padding);
 * upsample 1 folds a nearest x2 upsampling of x in front of the convolution. scale/shift NULL = raw x.
 * out [B*Ho*Wo, ldo] f16, ldo >= 9C and a multiple of 8 (extra columns are zeroed). */
COMA_API int coma_im2col3x3_f16(const void *x, int64_t B, int64_t H, int64_t W, int64_t C, int64_t ldx, int stride, int pad,
                                int upsample, const float *scale, const float *shift, int act, void *out, int64_t ldo,
                                coma_stream_t stream);
COMA_API int coma_layernorm_f16(const void *x, int64_t M, int64_t C, int64_t ldx, const float *gamma, const float *beta, float eps,
                                void *y, int64_t ldy, coma_stream_t stream);
/* LayerNorm statistics only: out[m] = (rstd_m, -rstd_m * mean_m) as float2 — the per-row inputs of a LayerNorm folded into the consuming
 * GEMM (coma_gemm_args.ln_row_stats). C % 8 == 0, C <= 2048. */
COMA_API int coma_layernorm_stats_f16(const void *x, int64_t M, int64_t C, int64_t ldx, float eps, float *out, coma_stream_t stream);
/* In-place softmax over the first L columns of each of R rows (row stride ld); columns [L, ld) are set to 0. */
COMA_API int coma_softmax_rows_f16(void *s, int64_t R, int64_t L, int64_t ld, coma_stream_t stream);
/* Same with a CAUSAL mask: the R rows form [S x L] matrices and row q of each only attends to columns <= q (the text-encoder
 * attention of transformers' CLIPTextModel, called at utils/adaptive_mask_inpainting.py:478,534). ld <= 1024. */
COMA_API int coma_softmax_rows_causal_f16(void *s, int64_t R, int64_t S, int64_t L, int64_t ld, coma_stream_t stream);
/* GEGLU: y[m,c] = h[m,c] * gelu(h[m,C+c]) (exact erf GELU), h [M,2C] -> y [M,C]. */
COMA_API int coma_geglu_f16(const void *h, int64_t M, int64_t C, int64_t ldh, void *y, int64_t ldy, coma_stream_t stream);
/* v [B,L,heads*d] (row stride ldv) -> vt [B,heads,d,Lpad] zero padded: the K-major operand of P.V. */
COMA_API int coma_transpose_heads_f16(const void *v, int64_t B, int64_t L, int64_t heads, int64_t d, int64_t ldv, void *vt,
                                      int64_t Lpad, coma_stream_t stream);
/* Sinusoidal timestep embedding [cos | sin] (flip_sin_to_cos, freq shift 0): t [B] f32 -> out [B,dim] f16. */
COMA_API int coma_timestep_embedding_f16(const float *t, int64_t B, int64_t dim, void *out, coma_stream_t stream);
COMA_API int coma_silu_f16(const void *x, int64_t n, void *y, coma_stream_t stream);

/* ---- L*: the denoising-loop glue of AdaptiveMaskInpaintPipeline.__call__ ---------------------------------------------- */

/* Classifier-free guidance + DDIMScheduler.step (eta 0, epsilon prediction, no clipping):
 * utils/adaptive_mask_inpainting.py:1009-1017. eps [2*half_rows, ld] f32 (uncond rows first, then text rows), C latent
 * channels; x, x_prev, x0 [half_rows, C] f32; alpha_t / alpha_prev = alphas_cumprod at t and t - 1000/steps. */
COMA_API int coma_cfg_ddim_step_f32(const float *eps, int64_t half_rows, int64_t ld, int64_t C, float guidance, const float *x,
                                    double alpha_t, double alpha_prev, float *x_prev, float *x0, coma_stream_t stream);
/* :990-996 — 9-channel UNet input (latents | mask | masked-image latents), duplicated for CFG: out [2*rows, ldo] f16. */
COMA_API int coma_assemble_unet_input_f16(const float *latents, const float *mask64, const float *masked_latents, int64_t rows,
                                          void *out, int64_t ldo, coma_stream_t stream);
/* adapt_mask (:1123-1141) + prepare_mask_and_masked_image tensor branch (:166-206, :239) + nearest /8 (:690), B images:
 *   area[b] = sum(seg[b]);  use_default = force_default || area[b] < area_thres  (area_thres = 512*512*human_detection_thres)
 *   mask = use_default ? (default >= 128) : (dilate(seg, 3x3 ones, iterations) != 0 && default != 0)
 *   masked_image = image * (1 - mask) (f16, row stride ldm);  mask_small[b, i, j] = mask[b, 8i, 8j]
 * seg [B,H,W] u8, default_mask [H,W] u8 (0..255), image [B,H,W,3] f32; scratch 2*B*H*W bytes; area_ws B uint64. Bit-exact. */
COMA_API int coma_adaptive_mask_u8(const uint8_t *seg, const uint8_t *default_mask, int64_t B, int64_t H, int64_t W, int dilate_iters,
                                   float area_thres, int force_default, const float *image, uint8_t *scratch, uint8_t *mask_out,
                                   void *masked_image, int64_t ldm, float *mask_small, int *used_default,
                                   unsigned long long *area_ws, coma_stream_t stream);
/* decode_to_npuint8_image (:1111-1115): (x/2+0.5).clamp(0,1)*255 truncated; img [rows, ld>=3] f32 -> out [rows,3] u8. */
COMA_API int coma_image_to_u8(const float *img, int64_t rows, int64_t ld, uint8_t *out, coma_stream_t stream);
/* DiagonalGaussianDistribution.sample * scaling_factor (:675-684): out = (mean + exp(0.5*logvar)*noise) * scaling. */
COMA_API int coma_sample_latents_f32(const float *mean, const float *logvar, const float *noise, int64_t n, float scaling,
                                     float *out, coma_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* COMA_B200_H */
