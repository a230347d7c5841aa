// HP-A support kernels around the tensor-core GEMM: everything in the UNet / VAE that is not a contraction.
// Activations are NHWC fp16 ([B, H*W, C] matrices, the GEMM's A operand layout); statistics and epilogues are fp32.
// Reference call sites: utils/adaptive_mask_inpainting.py:1001-1007 (unet), :680/:1086/:1112 (vae) — the layer
// definitions themselves are diffusers' (UNet2DConditionModel / AutoencoderKL 0.20.2, not vendored).
//   groupnorm_partial           GroupNorm(32) statistics in ONE launch (fp32 partials per channel, last CTA combines in fp64, fixed order)
//   groupnorm_from_stats        the same affine from the partial sums a convolution epilogue left behind (no pass over x)
//   groupnorm_apply             y = act(gn(x))                        (Transformer2D / VAE attention inputs, conv_norm_out)
//   im2col3x3                   [B,H,W,C] -> [B*Ho*Wo, 9C] with the GroupNorm affine + SiLU applied on the fly, stride 1/2,
//                               nearest x2 upsampling and the VAE encoder's asymmetric (0,1,0,1) padding folded in
//   layernorm                   per-token LayerNorm
//   softmax_rows                in-place row softmax of fp16 attention scores (fp32 math), padding columns zeroed
//   geglu                       hidden * gelu(gate)  (exact erf GELU)
//   transpose_heads             V [B,L,heads*d] -> V^T [B,heads,d,Lpad]   (K-major B operand for P·V)
//   timestep_embedding          sinusoidal embedding (flip_sin_to_cos, shift 0)
#include <cuda_fp16.h>
#include <math.h>

#include "common.cuh"

namespace coma {

// ---------------------------------------------------------------------------------------------- GroupNorm statistics
// Tail of the statistics kernels: CTA (chunk, b) has written its per-group partial sums; the LAST CTA of sample b to
// arrive (ticket counter, self-resetting) adds the chunks in index order (bit-reproducible, fp64) and writes the folded
// affine  scale[b,c] = rstd*gamma[c],  shift[b,c] = beta[c] - mean*rstd*gamma[c]  — no memset, no finalize launch.
__device__ __forceinline__ void groupnorm_finish(double *__restrict__ partial, unsigned *__restrict__ counter, int b, int C, int G,
                                                 long long count, float eps, const float *__restrict__ gamma,
                                                 const float *__restrict__ beta, float *__restrict__ mean, float *__restrict__ rstd,
                                                 float *__restrict__ scale, float *__restrict__ shift, float *sm /* >= 2*G floats */) {
    __shared__ int is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned ticket = atomicAdd(counter + b, 1u);
        is_last = (ticket == gridDim.x - 1);
        if (is_last) counter[b] = 0;  // ready for the next launch on this stream
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // all 256 threads: thread (slice, g) adds every NS-th chunk of group g, then the NS slices are combined in index order —
    // fixed association, independent of scheduling
    const int chunks = gridDim.x;
    double *dsm = reinterpret_cast<double *>(sm);
    const int NS = G <= 128 ? (int)blockDim.x / G : 1;  // slices per group
    for (int g0 = 0; g0 < G; g0 += (int)blockDim.x / NS) {
        const int g = g0 + (int)threadIdx.x % ((int)blockDim.x / NS), sl = (int)threadIdx.x / ((int)blockDim.x / NS);
        double s = 0.0, q = 0.0;
        if (g < G) {
            const double *p = partial + ((size_t)b * chunks * G + g) * 2;
#pragma unroll 4
            for (int k = sl; k < chunks; k += NS) {
                s += __ldcg(p + (size_t)k * G * 2);
                q += __ldcg(p + (size_t)k * G * 2 + 1);
            }
        }
        __syncthreads();
        if (g < G) {
            dsm[(sl * G + g) * 2] = s;       // [NS][G][2] doubles <= 2 * 256 doubles = 4 KB (host sizes smem accordingly)
            dsm[(sl * G + g) * 2 + 1] = q;
        }
        __syncthreads();
        if (sl == 0 && g < G) {
            for (int k = 1; k < NS; ++k) {
                s += dsm[(k * G + g) * 2];
                q += dsm[(k * G + g) * 2 + 1];
            }
            const double m = s / (double)count;
            double var = q / (double)count - m * m;
            var = var < 0.0 ? 0.0 : var;
            const double r = 1.0 / sqrt(var + (double)eps);
            if (mean) mean[b * G + g] = (float)m;
            if (rstd) rstd[b * G + g] = (float)r;
            s = m;
            q = r;
        }
        __syncthreads();
        if (sl == 0 && g < G) {
            dsm[2 * g] = s;      // slot (0, g): mean, rstd
            dsm[2 * g + 1] = q;
        }
    }
    __syncthreads();
    const int cpg = C / G;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const double m = reinterpret_cast<double *>(sm)[2 * (c / cpg)], r = reinterpret_cast<double *>(sm)[2 * (c / cpg) + 1];
        const double ga = gamma ? (double)gamma[c] : 1.0, be = beta ? (double)beta[c] : 0.0;
        scale[(size_t)b * C + c] = (float)(r * ga);
        shift[(size_t)b * C + c] = (float)(be - m * r * ga);
    }
}


// grid (chunks, B), block 256. 16-byte loads: a thread owns 8 consecutive channels; when C/8 < 256 the spare threads
// split the chunk's pixel rows. Partial sums go through shared memory, then one fp64 atomicAdd pair per (b, group).
__global__ void __launch_bounds__(256)
    groupnorm_partial_kernel(const __half *__restrict__ x, int HW, int C, long long ldx, int G, int rows_per_chunk,
                             double *__restrict__ partial /* [B,chunks,G,2] */, unsigned *__restrict__ counter, long long count, float eps,
                             const float *__restrict__ gamma, const float *__restrict__ beta, float *__restrict__ mean,
                             float *__restrict__ rstd, float *__restrict__ scale, float *__restrict__ shift) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float sm[];  // [rg][2][C]
    const int b = blockIdx.y, p0 = blockIdx.x * rows_per_chunk, p1 = min(HW, p0 + rows_per_chunk);
    const __half *xb = x + (size_t)b * HW * ldx;
    const int cgn = C / 8;                              // column groups of 8 channels
    const int rg = cgn >= 256 ? 1 : 256 / cgn;          // row lanes
    const int ry = threadIdx.x / cgn;                   // this thread's row lane (threads beyond rg*cgn idle)
    if (ry < rg) {
        for (int cg0 = threadIdx.x % cgn; cg0 < cgn; cg0 += (cgn >= 256 ? 256 : cgn)) {
            float s[8], q[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) s[k] = q[k] = 0.f;
            // 8 independent 16-byte loads in flight per thread (a one-load-per-iteration loop is pure L2 latency)
            for (int p = p0 + ry; p < p1; p += 8 * rg) {
                uint4 raw[8];
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    raw[u] = (p + u * rg < p1) ? *reinterpret_cast<const uint4 *>(xb + (size_t)(p + u * rg) * ldx + cg0 * 8) : make_uint4(0, 0, 0, 0);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const __half2 *h = reinterpret_cast<const __half2 *>(&raw[u]);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float2 v = __half22float2(h[k]);
                        s[2 * k] += v.x;
                        s[2 * k + 1] += v.y;
                        q[2 * k] = fmaf(v.x, v.x, q[2 * k]);
                        q[2 * k + 1] = fmaf(v.y, v.y, q[2 * k + 1]);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                sm[(ry * 2 + 0) * C + cg0 * 8 + k] = s[k];
                sm[(ry * 2 + 1) * C + cg0 * 8 + k] = q[k];
            }
        }
    }
    __syncthreads();
    const int cpg = C / G;
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        double s = 0.0, q = 0.0;
        for (int r = 0; r < rg; ++r)
            for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
                s += (double)sm[(r * 2 + 0) * C + c];
                q += (double)sm[(r * 2 + 1) * C + c];
            }
        double *dst = partial + (((size_t)b * gridDim.x + blockIdx.x) * G + g) * 2;
        dst[0] = s;
        dst[1] = q;
    }
    __syncthreads();  // sm is reused by the finishing block
    groupnorm_finish(partial, counter, b, C, G, count, eps, gamma, beta, mean, rstd, scale, shift, sm);
}

// scalar fallback (C % 8 != 0 or unaligned rows): thread t owns channels t, t+256, ...
__global__ void __launch_bounds__(256)
    groupnorm_partial_scalar_kernel(const __half *__restrict__ x, int HW, int C, long long ldx, int G, int rows_per_chunk,
                                    double *__restrict__ partial, unsigned *__restrict__ counter, long long count, float eps,
                                    const float *__restrict__ gamma, const float *__restrict__ beta, float *__restrict__ mean,
                                    float *__restrict__ rstd, float *__restrict__ scale, float *__restrict__ shift) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float sm[];  // [2][C]
    const int b = blockIdx.y, p0 = blockIdx.x * rows_per_chunk, p1 = min(HW, p0 + rows_per_chunk);
    const __half *xb = x + (size_t)b * HW * ldx;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f, q = 0.f;
        for (int p = p0; p < p1; ++p) {
            const float v = __half2float(xb[(size_t)p * ldx + c]);
            s += v;
            q = fmaf(v, v, q);
        }
        sm[c] = s;
        sm[C + c] = q;
    }
    __syncthreads();
    const int cpg = C / G;
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        double s = 0.0, q = 0.0;
        for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
            s += (double)sm[c];
            q += (double)sm[C + c];
        }
        double *dst = partial + (((size_t)b * gridDim.x + blockIdx.x) * G + g) * 2;
        dst[0] = s;
        dst[1] = q;
    }
    __syncthreads();
    groupnorm_finish(partial, counter, b, C, G, count, eps, gamma, beta, mean, rstd, scale, shift, sm);
}

// GroupNorm affine from the per-(32-row block, channel) partial sums a convolution epilogue left behind (gemm.cu, `stats`):
// grid (G, B), 256 threads. Thread t adds blocks t, t+256, ... of every channel of the group (fp64, fixed assignment), the 256
// partials are combined in index order -> bit-reproducible; no pass over the activation tensor at all.
__global__ void __launch_bounds__(256)
    groupnorm_from_stats_kernel(const float *__restrict__ stats, int HW, int rows_per_block, int C, int G, float eps, const float *__restrict__ gamma,
                                const float *__restrict__ beta, float *__restrict__ mean, float *__restrict__ rstd,
                                float *__restrict__ scale, float *__restrict__ shift) {
    pdl_trigger();
    pdl_wait();
    const int g = blockIdx.x, b = blockIdx.y, cpg = C / G, nblk = HW / rows_per_block;
    const float *base = stats + ((size_t)b * nblk * C + (size_t)g * cpg) * 2;
    double s = 0.0, q = 0.0;
    // four row blocks per iteration: their loads are independent (a one-row loop is pure L2 latency)
    for (int r = threadIdx.x; r < nblk; r += 4 * 256) {
        float fs[4] = {0.f, 0.f, 0.f, 0.f}, fq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (r + u * 256 < nblk) {
                const float2 *row = reinterpret_cast<const float2 *>(base + (size_t)(r + u * 256) * C * 2);
                for (int c = 0; c < cpg; ++c) {  // <= 80 channels per group: fp32 is exact enough within one 32-row block row
                    const float2 v = __ldcg(row + c);
                    fs[u] += v.x;
                    fq[u] += v.y;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            s += (double)fs[u];
            q += (double)fq[u];
        }
    }
    __shared__ double ss[256], qq[256];
    ss[threadIdx.x] = s;
    qq[threadIdx.x] = q;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {  // fixed-shape tree
        if (threadIdx.x < w) {
            ss[threadIdx.x] += ss[threadIdx.x + w];
            qq[threadIdx.x] += qq[threadIdx.x + w];
        }
        __syncthreads();
    }
    const double count = (double)HW * cpg;
    const double m = ss[0] / count;
    double var = qq[0] / count - m * m;
    var = var < 0.0 ? 0.0 : var;
    const double r = 1.0 / sqrt(var + (double)eps);
    if (threadIdx.x == 0) {
        if (mean) mean[b * G + g] = (float)m;
        if (rstd) rstd[b * G + g] = (float)r;
    }
    for (int c = threadIdx.x; c < cpg; c += 256) {
        const int ch = g * cpg + c;
        const double ga = gamma ? (double)gamma[ch] : 1.0, be = beta ? (double)beta[ch] : 0.0;
        scale[(size_t)b * C + ch] = (float)(r * ga);
        shift[(size_t)b * C + ch] = (float)(be - m * r * ga);
    }
}

// SiLU with the approximate reciprocal (MUFU.RCP, <= 2 ulp): the result is stored as fp16, and an IEEE division costs ~10 instructions
__device__ __forceinline__ float silu_f(float v) { return __fdividef(v, 1.0f + __expf(-v)); }

// y[b,p,c] = act(x*scale[b,c] + shift[b,c]); act: 0 none, 1 SiLU. grid (row blocks, B): a thread owns ONE 16-byte channel chunk
// (its 8 scale / shift values live in registers for the whole CTA: no per-element index division, no reloads) and walks
// the sample's pixel rows with a stride of 256 / (C/8) rows, four independent 16-byte loads in flight.
__global__ void __launch_bounds__(256)
    affine_act_kernel(const __half *__restrict__ x, int HW, int C, long long ldx, const float *__restrict__ scale,
                      const float *__restrict__ shift, int act, __half *__restrict__ y, long long ldy, int rows_per_cta) {
    pdl_trigger();
    pdl_wait();
    const int c8n = C / 8;
    const int b = blockIdx.y;
    const int lanes = c8n >= 256 ? 1 : 256 / c8n;       // row lanes per pass
    const int ry = threadIdx.x / c8n;
    if (ry >= lanes) return;
    const int p0 = blockIdx.x * rows_per_cta, p1 = min(HW, p0 + rows_per_cta);
    const __half *xb = x + (size_t)b * HW * ldx;
    __half *yb = y + (size_t)b * HW * ldy;
    for (int ch = threadIdx.x % c8n; ch < c8n; ch += (c8n >= 256 ? 256 : c8n)) {
        const int c = ch * 8;
        const float4 *sp = reinterpret_cast<const float4 *>(scale + (size_t)b * C + c), *tp = reinterpret_cast<const float4 *>(shift + (size_t)b * C + c);
        const float4 s0 = __ldg(sp), s1 = __ldg(sp + 1), t0 = __ldg(tp), t1 = __ldg(tp + 1);
        const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w}, sh[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
        for (int p = p0 + ry; p < p1; p += 4 * lanes) {
            uint4 raw[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (p + u * lanes < p1) raw[u] = *reinterpret_cast<const uint4 *>(xb + (size_t)(p + u * lanes) * ldx + c);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (p + u * lanes >= p1) break;
                const __half2 *h = reinterpret_cast<const __half2 *>(&raw[u]);
                uint4 o;
                __half2 *oh = reinterpret_cast<__half2 *>(&o);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    float2 v = __half22float2(h[t]);
                    v.x = fmaf(v.x, sc[2 * t], sh[2 * t]);
                    v.y = fmaf(v.y, sc[2 * t + 1], sh[2 * t + 1]);
                    if (act == 1) {
                        v.x = silu_f(v.x);
                        v.y = silu_f(v.y);
                    }
                    oh[t] = __floats2half2_rn(v.x, v.y);
                }
                *reinterpret_cast<uint4 *>(yb + (size_t)(p + u * lanes) * ldy + c) = o;
            }
        }
    }
}

// nearest x2 upsampling fused with the affine + activation: y[b, Y, X, c] = act(x[b, Y/2, X/2, c]*scale + shift)
__global__ void __launch_bounds__(256)
    upsample2x_affine_act_kernel(const __half *__restrict__ x, int B, int H, int W, int C, long long ldx,
                                 const float *__restrict__ scale, const float *__restrict__ shift, int act, __half *__restrict__ y,
                                 long long ldy) {
    pdl_trigger();
    pdl_wait();
    const int c8n = C / 8;
    const long long total = (long long)B * 4 * H * W * c8n;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % c8n) * 8;
        long long r = i / c8n;
        const int X = (int)(r % (2 * W)), Y = (int)((r / (2 * W)) % (2 * H)), b = (int)(r / ((long long)4 * H * W));
        uint4 o = *reinterpret_cast<const uint4 *>(x + ((size_t)(b * H + (Y >> 1)) * W + (X >> 1)) * ldx + c);
        if (scale) {
            __half2 *h = reinterpret_cast<__half2 *>(&o);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                float2 v = __half22float2(h[t]);
                const int cc = b * C + c + 2 * t;
                v.x = fmaf(v.x, scale[cc], shift[cc]);
                v.y = fmaf(v.y, scale[cc + 1], shift[cc + 1]);
                if (act == 1) {
                    v.x = silu_f(v.x);
                    v.y = silu_f(v.y);
                }
                h[t] = __floats2half2_rn(v.x, v.y);
            }
        }
        *reinterpret_cast<uint4 *>(y + r * ldy + c) = o;
    }
}

// ---------------------------------------------------------------------------------------------- im2col 3x3
// out[(b,oy,ox), tap*C + c] = act(x[b, iy, ix, c]*scale + shift), 0 outside the (upsampled) image.
// iy = oy*stride + ky - pad, ix likewise; with `up` the source pixel is (iy>>1, ix>>1) of the stored tensor.
template <bool VEC8>
__global__ void __launch_bounds__(256)
    im2col3x3_kernel(const __half *__restrict__ x, int B, int H, int W, int C, long long ldx, int Ho, int Wo, int stride, int pad,
                     int up, const float *__restrict__ scale, const float *__restrict__ shift, int act, __half *__restrict__ out,
                     long long ldo) {
    pdl_trigger();
    pdl_wait();
    const int Hin = up ? 2 * H : H, Win = up ? 2 * W : W;  // logical input extent
    const int cv = VEC8 ? C / 8 : C;
    const long long total = (long long)B * Ho * Wo * 9 * cv;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % cv) * (VEC8 ? 8 : 1);
        long long r = i / cv;
        const int tap = (int)(r % 9);
        r /= 9;
        const int ox = (int)(r % Wo), oy = (int)((r / Wo) % Ho), b = (int)(r / ((long long)Wo * Ho));
        const int iy = oy * stride + tap / 3 - pad, ix = ox * stride + tap % 3 - pad;
        const bool in = iy >= 0 && iy < Hin && ix >= 0 && ix < Win;
        __half *dst = out + r * ldo + (long long)tap * C + c;
        if (VEC8) {
            uint4 o = make_uint4(0u, 0u, 0u, 0u);
            if (in) {
                const int sy = up ? iy >> 1 : iy, sx = up ? ix >> 1 : ix;
                o = *reinterpret_cast<const uint4 *>(x + ((size_t)(b * H + sy) * W + sx) * ldx + c);
                if (scale) {
                    __half2 *h = reinterpret_cast<__half2 *>(&o);
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        float2 v = __half22float2(h[t]);
                        const int cc = b * C + c + 2 * t;
                        v.x = fmaf(v.x, scale[cc], shift[cc]);
                        v.y = fmaf(v.y, scale[cc + 1], shift[cc + 1]);
                        if (act == 1) {
                            v.x = silu_f(v.x);
                            v.y = silu_f(v.y);
                        }
                        h[t] = __floats2half2_rn(v.x, v.y);
                    }
                }
            }
            *reinterpret_cast<uint4 *>(dst) = o;
        } else {
            float v = 0.f;
            if (in) {
                const int sy = up ? iy >> 1 : iy, sx = up ? ix >> 1 : ix;
                v = __half2float(x[((size_t)(b * H + sy) * W + sx) * ldx + c]);
                if (scale) {
                    v = fmaf(v, scale[b * C + c], shift[b * C + c]);
                    if (act == 1) v = silu_f(v);
                }
            }
            *dst = __float2half_rn(v);
        }
    }
}

// zero the K-padding columns [K, ldo) of an im2col matrix (only when 9C is not a multiple of 8)
__global__ void zero_cols_kernel(__half *__restrict__ out, long long rows, int K, long long ldo) {
    pdl_trigger();
    pdl_wait();
    const int padc = (int)(ldo - K);
    const long long total = rows * padc;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
        out[(i / padc) * ldo + K + (i % padc)] = __float2half_rn(0.f);
}

// ---------------------------------------------------------------------------------------------- LayerNorm (warp per row)
__global__ void __launch_bounds__(256)
    layernorm_kernel(const __half *__restrict__ x, long long M, int C, long long ldx, const float *__restrict__ gamma,
                     const float *__restrict__ beta, float eps, __half *__restrict__ y, long long ldy) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= M) return;
    const __half *xr = x + row * ldx;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += __half2float(xr[c]);
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
    for (int c = lane; c < C; c += 32) {
        const float d = __half2float(xr[c]) - mean;
        q = fmaf(d, d, q);
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
    __half *yr = y + row * ldy;
    for (int c = lane; c < C; c += 32) yr[c] = __float2half_rn((__half2float(xr[c]) - mean) * rstd * gamma[c] + beta[c]);
}

// vectorised form (C % 8 == 0, C <= 256*NCH, 16-byte aligned rows): the row is read ONCE as 16-byte chunks into registers
// (lane owns chunks lane, lane+32, ...), mean and centred variance come from the registers, one 16-byte store per chunk.
template <int NCH>
__global__ void __launch_bounds__(256)
    layernorm_vec_kernel(const __half *__restrict__ x, long long M, int C, long long ldx, const float *__restrict__ gamma,
                         const float *__restrict__ beta, float eps, __half *__restrict__ y, long long ldy) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= M) return;
    const int nch = C / 8;
    const __half *xr = x + row * ldx;
    float v[NCH][8];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
        const int ch = lane + 32 * j;
        uint4 raw = make_uint4(0, 0, 0, 0);
        if (ch < nch) raw = *reinterpret_cast<const uint4 *>(xr + ch * 8);
        const __half2 *h = reinterpret_cast<const __half2 *>(&raw);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const float2 f = __half22float2(h[t]);
            v[j][2 * t] = f.x;
            v[j][2 * t + 1] = f.y;
            s += f.x + f.y;
        }
    }
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
        if (lane + 32 * j < nch) {
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const float d = v[j][t] - mean;
                q = fmaf(d, d, q);
            }
        }
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
    __half *yr = y + row * ldy;
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
        const int ch = lane + 32 * j;
        if (ch < nch) {
            const float4 g0 = __ldg(reinterpret_cast<const float4 *>(gamma + ch * 8)), g1 = __ldg(reinterpret_cast<const float4 *>(gamma + ch * 8 + 4));
            const float4 b0 = __ldg(reinterpret_cast<const float4 *>(beta + ch * 8)), b1 = __ldg(reinterpret_cast<const float4 *>(beta + ch * 8 + 4));
            const float ga[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, be[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            uint4 o;
            __half2 *oh = reinterpret_cast<__half2 *>(&o);
#pragma unroll
            for (int t = 0; t < 4; ++t)
                oh[t] = __floats2half2_rn((v[j][2 * t] - mean) * rstd * ga[2 * t] + be[2 * t],
                                          (v[j][2 * t + 1] - mean) * rstd * ga[2 * t + 1] + be[2 * t + 1]);
            *reinterpret_cast<uint4 *>(yr + ch * 8) = o;
        }
    }
}

// LayerNorm statistics only: out[row] = (rstd, -rstd * mean) — what the GEMM epilogue needs to apply a LayerNorm that was folded into
// its weights (gemm.cu, `ln_row_stats`): y = rstd * (x W'^T) - rstd * mean * c1 + c2. Same arithmetic as layernorm_vec_kernel (row held in
// registers, centred variance), no output tensor.
template <int NCH>
__global__ void __launch_bounds__(256)
    layernorm_stats_kernel(const __half *__restrict__ x, long long M, int C, long long ldx, float eps, float2 *__restrict__ out) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= M) return;
    const int nch = C / 8;
    const __half *xr = x + row * ldx;
    float v[NCH][8];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
        const int ch = lane + 32 * j;
        uint4 raw = make_uint4(0, 0, 0, 0);
        if (ch < nch) raw = *reinterpret_cast<const uint4 *>(xr + ch * 8);
        const __half2 *h = reinterpret_cast<const __half2 *>(&raw);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const float2 f = __half22float2(h[t]);
            v[j][2 * t] = f.x;
            v[j][2 * t + 1] = f.y;
            s += f.x + f.y;
        }
    }
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
        if (lane + 32 * j < nch) {
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const float d = v[j][t] - mean;
                q = fmaf(d, d, q);
            }
        }
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
    if (lane == 0) out[row] = make_float2(rstd, -rstd * mean);
}

// ---------------------------------------------------------------------------------------------- row softmax (in place)
// Rows are read ONCE into registers, reduced with shuffles, and written once.
//   L <= 1024 : one warp per row, up to 32 values per lane (cross-attention rows, L = 77, and the 16x16 / 32x32 levels)
//   L <= 4096 and 16-byte aligned rows: one 256-thread CTA per row, 16 values per thread as two 16-byte loads
//   otherwise : three-pass fallback
// causal_S > 0: the rows form [causal_S x L] matrices and row q of a matrix only sees columns <= q (CLIP's causal text mask)
__global__ void __launch_bounds__(256) softmax_rows_warp_kernel(__half *__restrict__ s, long long R, int L_full, long long ld, int causal_S) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= R) return;
    const int L = causal_S > 0 ? min(L_full, (int)(r % causal_S) + 1) : L_full;
    __half *row = s + r * ld;
    float v[32];
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const int c = lane + 32 * j;
        v[j] = (c < L) ? __half2float(row[c]) : -INFINITY;
        m = fmaxf(m, v[j]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        v[j] = (lane + 32 * j < L) ? __expf(v[j] - m) : 0.f;
        sum += v[j];
    }
    const float inv = 1.0f / warp_sum(sum);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const int c = lane + 32 * j;
        if (c < ld && c < 1024) row[c] = __float2half_rn(v[j] * inv);  // columns [L, ld) receive 0
    }
}

__global__ void __launch_bounds__(256) softmax_rows_block_kernel(__half *__restrict__ s, int L, long long ld) {
    pdl_trigger();
    pdl_wait();
    __half *row = s + (long long)blockIdx.x * ld;
    __shared__ float red[8];
    __shared__ float bc;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float v[16];
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int c = (threadIdx.x + 256 * j) * 8;
        uint4 raw = make_uint4(0xfc00fc00u, 0xfc00fc00u, 0xfc00fc00u, 0xfc00fc00u);  // -inf halves
        if (c < L) raw = *reinterpret_cast<const uint4 *>(row + c);
        const __half2 *h = reinterpret_cast<const __half2 *>(&raw);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const float2 f = __half22float2(h[t]);
            v[8 * j + 2 * t] = (c + 2 * t < L) ? f.x : -INFINITY;
            v[8 * j + 2 * t + 1] = (c + 2 * t + 1 < L) ? f.y : -INFINITY;
            m = fmaxf(m, fmaxf(v[8 * j + 2 * t], v[8 * j + 2 * t + 1]));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) red[warp] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = red[0];
        for (int w = 1; w < 8; ++w) t = fmaxf(t, red[w]);
        bc = t;
    }
    __syncthreads();
    m = bc;
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        v[j] = __expf(v[j] - m);  // exp(-inf) = 0 for the masked tail
        sum += v[j];
    }
    sum = warp_sum(sum);
    __syncthreads();
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[w];
        bc = 1.0f / t;
    }
    __syncthreads();
    const float inv = bc;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int c = (threadIdx.x + 256 * j) * 8;
        if (c < ld) {
            uint4 o;
            __half2 *h = reinterpret_cast<__half2 *>(&o);
#pragma unroll
            for (int t = 0; t < 4; ++t) h[t] = __floats2half2_rn(v[8 * j + 2 * t] * inv, v[8 * j + 2 * t + 1] * inv);
            *reinterpret_cast<uint4 *>(row + c) = o;
        }
    }
}

__global__ void __launch_bounds__(256) softmax_rows_kernel(__half *__restrict__ s, int L, long long ld) {
    pdl_trigger();
    pdl_wait();
    __half *row = s + (long long)blockIdx.x * ld;
    __shared__ float red[8];
    __shared__ float bc;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float m = -INFINITY;
    for (int c = threadIdx.x; c < L; c += blockDim.x) m = fmaxf(m, __half2float(row[c]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) red[warp] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = red[0];
        for (int w = 1; w < 8; ++w) t = fmaxf(t, red[w]);
        bc = t;
    }
    __syncthreads();
    m = bc;
    float sum = 0.f;
    for (int c = threadIdx.x; c < L; c += blockDim.x) sum += __expf(__half2float(row[c]) - m);
    sum = warp_sum(sum);
    __syncthreads();
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[w];
        bc = 1.0f / t;
    }
    __syncthreads();
    const float inv = bc;
    for (int c = threadIdx.x; c < (int)ld; c += blockDim.x)
        row[c] = __float2half_rn(c < L ? __expf(__half2float(row[c]) - m) * inv : 0.f);
}

// ---------------------------------------------------------------------------------------------- GEGLU
__global__ void __launch_bounds__(256)
    geglu_kernel(const __half *__restrict__ h, long long M, int C, long long ldh, __half *__restrict__ y, long long ldy) {
    pdl_trigger();
    pdl_wait();
    const int c2n = C / 2;
    const long long total = M * c2n;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / c2n;
        const int c = (int)(i % c2n) * 2;
        const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(h + r * ldh + c));
        const float2 g = __half22float2(*reinterpret_cast<const __half2 *>(h + r * ldh + C + c));
        const float gx = 0.5f * g.x * (1.0f + erff(g.x * 0.70710678118654752f));
        const float gy = 0.5f * g.y * (1.0f + erff(g.y * 0.70710678118654752f));
        *reinterpret_cast<__half2 *>(y + r * ldy + c) = __floats2half2_rn(a.x * gx, a.y * gy);
    }
}

// ---------------------------------------------------------------------------------------------- V -> V^T per head
// v [B, L, heads*d] (row stride ldv) -> vt [B, heads, d, Lpad], zero padded. 32x32 smem tiles.
__global__ void __launch_bounds__(256)
    transpose_heads_kernel(const __half *__restrict__ v, int L, int heads, int d, long long ldv, __half *__restrict__ vt, int Lpad) {
    pdl_trigger();
    pdl_wait();
    __shared__ __half tile[32][33];
    const int bh = blockIdx.z, b = bh / heads, hd = bh % heads;
    const int l0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int j = ty; j < 32; j += 8) {
        const int l = l0 + j, dd = d0 + tx;
        tile[j][tx] = (l < L && dd < d) ? v[((size_t)b * L + l) * ldv + hd * d + dd] : __float2half_rn(0.f);
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int dd = d0 + j, l = l0 + tx;
        if (dd < d && l < Lpad) vt[(((size_t)b * heads + hd) * d + dd) * Lpad + l] = tile[tx][j];
    }
}

// ---------------------------------------------------------------------------------------------- timestep embedding
// diffusers get_timestep_embedding(t, dim, flip_sin_to_cos=True, downscale_freq_shift=0): [cos(t f_i) | sin(t f_i)]
__global__ void timestep_embedding_kernel(const float *__restrict__ t, int B, int dim, __half *__restrict__ out) {
    pdl_trigger();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int half = dim / 2;
    if (i >= B * half) return;
    const int b = i / half, k = i % half;
    const float f = expf(-logf(10000.0f) * (float)k / (float)half);
    const float a = t[b] * f;
    out[b * dim + k] = __float2half_rn(cosf(a));
    out[b * dim + half + k] = __float2half_rn(sinf(a));
}

// elementwise SiLU on fp16 (time-embedding MLP input of every ResnetBlock)
__global__ void silu_kernel(const __half *__restrict__ x, long long n, __half *__restrict__ y) {
    pdl_trigger();
    pdl_wait();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        y[i] = __float2half_rn(silu_f(__half2float(x[i])));
}

static inline unsigned blocks_for(long long total, int threads = 256) {
    long long b = (total + threads - 1) / threads;
    const long long cap = (long long)kNumSM * 16;
    return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace coma

using namespace coma;

extern "C" int64_t coma_groupnorm_workspace_doubles(int64_t B, int G) { return 2LL * G * (4LL * coma::kNumSM + B); }

extern "C" int coma_groupnorm_affine_f16(const void *x, int64_t B, int64_t HW, int64_t C, int64_t ldx, int G, float eps,
                                         const float *gamma, const float *beta, double *workspace, unsigned *counters, float *mean,
                                         float *rstd, float *scale, float *shift, coma_stream_t stream) {
    COMA_REQUIRE(x && workspace && counters && scale && shift, "null pointer");
    COMA_REQUIRE(B > 0 && HW > 0 && C > 0 && G > 0 && C % G == 0 && ldx >= C, "bad sizes");
    COMA_REQUIRE(C <= 8192 && B <= 65535 && G <= 1024, "C, B or G too large");
    cudaStream_t st = (cudaStream_t)stream;
    // enough chunks to fill the machine, each at least 16 pixel rows; B * chunks <= 4 * 148 + B (the workspace bound)
    long long chunks = (4LL * kNumSM + B - 1) / B;
    long long rows = (HW + chunks - 1) / chunks;
    rows = rows < 8 ? 8 : rows;
    chunks = (HW + rows - 1) / rows;
    const bool vec = (C % 8 == 0) && (ldx % 8 == 0) && ((uintptr_t)x % 16 == 0);
    const long long count = HW * (C / G);
    if (vec) {
        const int cgn = (int)(C / 8), rg = cgn >= 256 ? 1 : 256 / cgn;
        size_t smem = sizeof(float) * 2 * C * rg;
        if (smem < sizeof(double) * 2 * (256 + G)) smem = sizeof(double) * 2 * (256 + G);
        launch_pdl(groupnorm_partial_kernel, dim3((unsigned)chunks, (unsigned)B), dim3(256), smem, st, (const __half *)x, (int)HW, (int)C, ldx, G,
                   (int)rows, workspace, counters, count, eps, gamma, beta, mean, rstd, scale, shift);
    } else {
        size_t smem = sizeof(float) * 2 * C;
        if (smem < sizeof(double) * 2 * (256 + G)) smem = sizeof(double) * 2 * (256 + G);
        launch_pdl(groupnorm_partial_scalar_kernel, dim3((unsigned)chunks, (unsigned)B), dim3(256), smem, st, (const __half *)x, (int)HW, (int)C,
                   ldx, G, (int)rows, workspace, counters, count, eps, gamma, beta, mean, rstd, scale, shift);
    }
    return check_launch("groupnorm_partial_kernel");
}

extern "C" int coma_groupnorm_from_stats_rb_f32(const float *stats, int64_t B, int64_t HW, int64_t rows_per_block, int64_t C, int G, float eps,
                                                const float *gamma, const float *beta, float *mean, float *rstd, float *scale, float *shift,
                                                coma_stream_t stream) {
    COMA_REQUIRE(stats && scale && shift, "null pointer");
    COMA_REQUIRE(B > 0 && HW > 0 && rows_per_block > 0 && HW % rows_per_block == 0 && C > 0 && G > 0 && C % G == 0 && B <= 65535 && G <= 65535, "bad sizes");
    COMA_REQUIRE((uintptr_t)stats % 8 == 0, "stats must be 8-byte aligned");
    launch_pdl(groupnorm_from_stats_kernel, dim3((unsigned)G, (unsigned)B), dim3(256), 0, (cudaStream_t)stream, stats, (int)HW, (int)rows_per_block, (int)C,
               G, eps, gamma, beta, mean, rstd, scale, shift);
    return check_launch("groupnorm_from_stats_kernel");
}

extern "C" int coma_groupnorm_from_stats_f32(const float *stats, int64_t B, int64_t HW, int64_t C, int G, float eps, const float *gamma,
                                             const float *beta, float *mean, float *rstd, float *scale, float *shift,
                                             coma_stream_t stream) {
    return coma_groupnorm_from_stats_rb_f32(stats, B, HW, 32, C, G, eps, gamma, beta, mean, rstd, scale, shift, stream);
}

extern "C" int coma_affine_act_f16(const void *x, int64_t B, int64_t HW, int64_t C, int64_t ldx, const float *scale,
                                   const float *shift, int act, void *y, int64_t ldy, coma_stream_t stream) {
    COMA_REQUIRE(x && y && scale && shift, "null pointer");
    COMA_REQUIRE(B > 0 && HW > 0 && C > 0 && C % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0, "C, ldx, ldy must be multiples of 8");
    COMA_REQUIRE(((uintptr_t)x | (uintptr_t)y | (uintptr_t)scale | (uintptr_t)shift) % 16 == 0, "x / y / scale / shift must be 16-byte aligned");
    COMA_REQUIRE(B <= 65535 && HW < (1LL << 31), "B or HW too large");
    // ~8 CTAs per SM in flight, each at least one unrolled pass (4 rows per row lane)
    const long long c8n = C / 8, lanes = c8n >= 256 ? 1 : 256 / c8n;
    long long chunks = (8LL * kNumSM + B - 1) / B;
    long long rows = (HW + chunks - 1) / chunks;
    rows = rows < 4 * lanes ? 4 * lanes : rows;
    chunks = (HW + rows - 1) / rows;
    launch_pdl(affine_act_kernel, dim3((unsigned)chunks, (unsigned)B), dim3(256), 0, (cudaStream_t)stream, (const __half *)x, (int)HW, (int)C, ldx,
               scale, shift, act, (__half *)y, ldy, (int)rows);
    return check_launch("affine_act_kernel");
}

extern "C" int coma_upsample2x_affine_act_f16(const void *x, int64_t B, int64_t H, int64_t W, int64_t C, int64_t ldx,
                                              const float *scale, const float *shift, int act, void *y, int64_t ldy,
                                              coma_stream_t stream) {
    COMA_REQUIRE(x && y, "null pointer");
    COMA_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0, "C, ldx, ldy must be multiples of 8");
    COMA_REQUIRE(((uintptr_t)x | (uintptr_t)y) % 16 == 0 && !scale == !shift, "bad arguments");
    launch_pdl(upsample2x_affine_act_kernel, dim3(blocks_for(B * 4 * H * W * (C / 8))), dim3(256), 0, (cudaStream_t)stream, 
        (const __half *)x, (int)B, (int)H, (int)W, (int)C, ldx, scale, shift, act, (__half *)y, ldy);
    return check_launch("upsample2x_affine_act_kernel");
}

extern "C" int coma_im2col3x3_f16(const void *x, int64_t B, int64_t H, int64_t W, int64_t C, int64_t ldx, int stride, int pad,
                                  int upsample, const float *scale, const float *shift, int act, void *out, int64_t ldo,
                                  coma_stream_t stream) {
    COMA_REQUIRE(x && out, "null pointer");
    COMA_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && ldx >= C, "bad sizes");
    COMA_REQUIRE((stride == 1 || stride == 2) && (pad == 0 || pad == 1) && (upsample == 0 || upsample == 1), "bad conv geometry");
    COMA_REQUIRE(ldo >= 9 * C && ldo % 8 == 0, "ldo must be >= 9*C and a multiple of 8");
    COMA_REQUIRE(!scale == !shift, "scale and shift come together");
    const int64_t Hin = upsample ? 2 * H : H, Win = upsample ? 2 * W : W;
    // stride 1: same size; stride 2 with pad 1 (UNet Downsample2D): floor((H+2-3)/2)+1; stride 2, pad 0 after the VAE
    // encoder's (0,1,0,1) padding: floor((H+1-3)/2)+1
    const int64_t Ho = stride == 1 ? Hin : (pad ? (Hin + 2 - 3) / 2 + 1 : (Hin + 1 - 3) / 2 + 1);
    const int64_t Wo = stride == 1 ? Win : (pad ? (Win + 2 - 3) / 2 + 1 : (Win + 1 - 3) / 2 + 1);
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = (C % 8 == 0) && (ldx % 8 == 0) && (((uintptr_t)x | (uintptr_t)out) % 16 == 0);
    const long long rows = B * Ho * Wo;
    if (vec)
        launch_pdl(im2col3x3_kernel<true>, dim3(blocks_for(rows * 9 * (C / 8))), dim3(256), 0, st, (const __half *)x, (int)B, (int)H, (int)W, (int)C, ldx,
                                                                            (int)Ho, (int)Wo, stride, pad, upsample, scale, shift, act,
                                                                            (__half *)out, ldo);
    else
        launch_pdl(im2col3x3_kernel<false>, dim3(blocks_for(rows * 9 * C)), dim3(256), 0, st, (const __half *)x, (int)B, (int)H, (int)W, (int)C, ldx, (int)Ho,
                                                                       (int)Wo, stride, pad, upsample, scale, shift, act,
                                                                       (__half *)out, ldo);
    if (int e = check_launch("im2col3x3_kernel")) return e;
    if (ldo > 9 * C) {
        launch_pdl(zero_cols_kernel, dim3(blocks_for(rows * (ldo - 9 * C))), dim3(256), 0, st, (__half *)out, rows, (int)(9 * C), ldo);
        return check_launch("zero_cols_kernel");
    }
    return 0;
}

extern "C" int coma_layernorm_f16(const void *x, int64_t M, int64_t C, int64_t ldx, const float *gamma, const float *beta, float eps,
                                  void *y, int64_t ldy, coma_stream_t stream) {
    COMA_REQUIRE(x && y && gamma && beta, "null pointer");
    COMA_REQUIRE(M > 0 && C > 0 && ldx >= C && ldy >= C, "bad sizes");
    const bool vec = C % 8 == 0 && C <= 2048 && ldx % 8 == 0 && ldy % 8 == 0 &&
                     ((uintptr_t)x | (uintptr_t)y | (uintptr_t)gamma | (uintptr_t)beta) % 16 == 0;
    const dim3 grid((unsigned)((M + 7) / 8));
    cudaStream_t st = (cudaStream_t)stream;
    if (vec && C <= 512)
        launch_pdl(layernorm_vec_kernel<2>, grid, dim3(256), 0, st, (const __half *)x, M, (int)C, ldx, gamma, beta, eps, (__half *)y, ldy);
    else if (vec && C <= 1024)
        launch_pdl(layernorm_vec_kernel<4>, grid, dim3(256), 0, st, (const __half *)x, M, (int)C, ldx, gamma, beta, eps, (__half *)y, ldy);
    else if (vec)
        launch_pdl(layernorm_vec_kernel<8>, grid, dim3(256), 0, st, (const __half *)x, M, (int)C, ldx, gamma, beta, eps, (__half *)y, ldy);
    else
        launch_pdl(layernorm_kernel, grid, dim3(256), 0, st, (const __half *)x, M, (int)C, ldx, gamma, beta, eps, (__half *)y, ldy);
    return check_launch("layernorm_kernel");
}

extern "C" int coma_layernorm_stats_f16(const void *x, int64_t M, int64_t C, int64_t ldx, float eps, float *out, coma_stream_t stream) {
    COMA_REQUIRE(x && out, "null pointer");
    COMA_REQUIRE(M > 0 && C > 0 && ldx >= C && C % 8 == 0 && C <= 2048 && ldx % 8 == 0, "needs C % 8 == 0, C <= 2048, ldx % 8 == 0");
    COMA_REQUIRE((uintptr_t)x % 16 == 0 && (uintptr_t)out % 8 == 0, "x must be 16-byte aligned, out 8-byte aligned");
    const dim3 grid((unsigned)((M + 7) / 8));
    cudaStream_t st = (cudaStream_t)stream;
    if (C <= 512) launch_pdl(layernorm_stats_kernel<2>, grid, dim3(256), 0, st, (const __half *)x, M, (int)C, ldx, eps, (float2 *)out);
    else if (C <= 1024) launch_pdl(layernorm_stats_kernel<4>, grid, dim3(256), 0, st, (const __half *)x, M, (int)C, ldx, eps, (float2 *)out);
    else launch_pdl(layernorm_stats_kernel<8>, grid, dim3(256), 0, st, (const __half *)x, M, (int)C, ldx, eps, (float2 *)out);
    return check_launch("layernorm_stats_kernel");
}

extern "C" int coma_softmax_rows_f16(void *s, int64_t R, int64_t L, int64_t ld, coma_stream_t stream) {
    COMA_REQUIRE(s, "null pointer");
    COMA_REQUIRE(R > 0 && L > 0 && ld >= L && R < (1LL << 31), "bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    if (ld <= 1024) {
        launch_pdl(softmax_rows_warp_kernel, dim3((unsigned)((R + 7) / 8)), dim3(256), 0, st, (__half *)s, R, (int)L, ld, 0);
        return check_launch("softmax_rows_warp_kernel");
    }
    if (ld <= 4096 && ld % 8 == 0 && (uintptr_t)s % 16 == 0) {
        launch_pdl(softmax_rows_block_kernel, dim3((unsigned)R), dim3(256), 0, st, (__half *)s, (int)L, ld);
        return check_launch("softmax_rows_block_kernel");
    }
    launch_pdl(softmax_rows_kernel, dim3((unsigned)R), dim3(256), 0, st, (__half *)s, (int)L, ld);
    return check_launch("softmax_rows_kernel");
}

extern "C" int coma_softmax_rows_causal_f16(void *s, int64_t R, int64_t S, int64_t L, int64_t ld, coma_stream_t stream) {
    COMA_REQUIRE(s, "null pointer");
    COMA_REQUIRE(R > 0 && S > 0 && L > 0 && ld >= L && ld <= 1024 && R % S == 0 && R < (1LL << 31), "bad sizes (causal rows: ld <= 1024)");
    launch_pdl(softmax_rows_warp_kernel, dim3((unsigned)((R + 7) / 8)), dim3(256), 0, (cudaStream_t)stream, (__half *)s, R, (int)L, ld, (int)S);
    return check_launch("softmax_rows_warp_kernel");
}

extern "C" int coma_geglu_f16(const void *h, int64_t M, int64_t C, int64_t ldh, void *y, int64_t ldy, coma_stream_t stream) {
    COMA_REQUIRE(h && y, "null pointer");
    COMA_REQUIRE(M > 0 && C > 0 && C % 2 == 0 && ldh >= 2 * C && ldy >= C && ldh % 2 == 0 && ldy % 2 == 0, "bad sizes");
    launch_pdl(geglu_kernel, dim3(blocks_for(M * (C / 2))), dim3(256), 0, (cudaStream_t)stream, (const __half *)h, M, (int)C, ldh, (__half *)y, ldy);
    return check_launch("geglu_kernel");
}

extern "C" int coma_transpose_heads_f16(const void *v, int64_t B, int64_t L, int64_t heads, int64_t d, int64_t ldv, void *vt,
                                        int64_t Lpad, coma_stream_t stream) {
    COMA_REQUIRE(v && vt, "null pointer");
    COMA_REQUIRE(B > 0 && L > 0 && heads > 0 && d > 0 && Lpad >= L && ldv >= heads * d && B * heads <= 65535, "bad sizes");
    dim3 grid((unsigned)((Lpad + 31) / 32), (unsigned)((d + 31) / 32), (unsigned)(B * heads));
    launch_pdl(transpose_heads_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const __half *)v, (int)L, (int)heads, (int)d, ldv, (__half *)vt,
                                                                 (int)Lpad);
    return check_launch("transpose_heads_kernel");
}

extern "C" int coma_timestep_embedding_f16(const float *t, int64_t B, int64_t dim, void *out, coma_stream_t stream) {
    COMA_REQUIRE(t && out, "null pointer");
    COMA_REQUIRE(B > 0 && dim > 0 && dim % 2 == 0, "bad sizes");
    launch_pdl(timestep_embedding_kernel, dim3((unsigned)((B * dim / 2 + 127) / 128)), dim3(128), 0, (cudaStream_t)stream, t, (int)B, (int)dim, (__half *)out);
    return check_launch("timestep_embedding_kernel");
}

extern "C" int coma_silu_f16(const void *x, int64_t n, void *y, coma_stream_t stream) {
    COMA_REQUIRE(x && y && n > 0, "bad arguments");
    launch_pdl(silu_kernel, dim3(blocks_for(n)), dim3(256), 0, (cudaStream_t)stream, (const __half *)x, n, (__half *)y);
    return check_launch("silu_kernel");
}
