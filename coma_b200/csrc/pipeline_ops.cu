// HP-A loop kernels (everything in AdaptiveMaskInpaintPipeline.__call__'s denoising loop that is not the UNet / VAE):
//   cfg_ddim_step        classifier-free guidance + DDIMScheduler.step (eta = 0)   utils/adaptive_mask_inpainting.py:1009-1017
//   assemble_unet_input  cat([latents]*2) ++ mask ++ masked_image_latents -> 9-ch NHWC fp16           :990-996
//   adaptive mask        area test, k x (3x3) dilation == (2k+1)^2 box max, AND default mask, binarise,
//                        masked image, nearest /8 mask                              :1123-1141, :166-206, :239, :690
//   image_to_u8          (x/2+0.5).clamp(0,1)*255 truncated to uint8, HWC            :1111-1115
//   sample_latents       mean + exp(0.5*logvar)*noise, scaled                        :675-684 (DiagonalGaussianDistribution.sample)
#include <cuda_fp16.h>
#include <math.h>

#include "common.cuh"

namespace coma {

// eps: fp32 [2B*P, ld] rows (first B*P rows unconditional, next B*P text-conditioned), C = 4 latent channels.
__global__ void cfg_ddim_kernel(const float *__restrict__ eps, long long half_rows, int ld, int C, float guidance,
                                const float *__restrict__ x, float sqrt_a_t, float sqrt_1m_a_t, float sqrt_a_prev,
                                float sqrt_1m_a_prev, float *__restrict__ x_prev, float *__restrict__ x0) {
    pdl_trigger();
    pdl_wait();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= half_rows * C) return;
    const long long r = i / C;
    const int c = (int)(i % C);
    const float eu = eps[r * ld + c], et = eps[(r + half_rows) * ld + c];
    const float e = eu + guidance * (et - eu);                       // :1011-1012
    const float xv = x[i];
    const float p0 = (xv - sqrt_1m_a_t * e) / sqrt_a_t;               // pred_original_sample
    x0[i] = p0;
    x_prev[i] = sqrt_a_prev * p0 + sqrt_1m_a_prev * e;                // eta = 0: no noise term
}

// in9[(b', p), 0:4] = latents[b, p, :], [4] = mask64[b, p], [5:9] = masked_latents[b, p, :], b' in {b, b + B} (CFG duplicate)
__global__ void assemble_input_kernel(const float *__restrict__ lat, const float *__restrict__ mask64,
                                      const float *__restrict__ masked_lat, long long rows, __half *__restrict__ out, int ldo) {
    pdl_trigger();
    pdl_wait();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * 9) return;
    const long long r = i / 9;
    const int c = (int)(i % 9);
    const float v = c < 4 ? lat[r * 4 + c] : (c == 4 ? mask64[r] : masked_lat[r * 4 + (c - 5)]);
    const __half h = __float2half_rn(v);
    out[r * ldo + c] = h;
    out[(r + rows) * ldo + c] = h;
}

// ---- adaptive mask --------------------------------------------------------------------------------------------------
__global__ void mask_area_kernel(const uint8_t *__restrict__ seg, int n, unsigned long long *__restrict__ area) {
    pdl_trigger();
    pdl_wait();
    unsigned long long s = 0;
    const uint8_t *m = seg + (size_t)blockIdx.y * n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s += m[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(area + blockIdx.y, s);
}

// horizontal (dir = 0) or vertical (dir = 1) running max of radius k over a u8 image; pixels outside the image are
// ignored — identical to k iterations of cv2.dilate with a 3x3 ones kernel (default border handling).
__global__ void box_max_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, int H, int W, int k, int dir) {
    pdl_trigger();
    pdl_wait();
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
    if (x >= W) return;
    const uint8_t *src = in + (size_t)b * H * W;
    uint8_t m = 0;
    if (dir == 0) {
        const int lo = max(0, x - k), hi = min(W - 1, x + k);
        for (int j = lo; j <= hi; ++j) m = max(m, src[(size_t)y * W + j]);
    } else {
        const int lo = max(0, y - k), hi = min(H - 1, y + k);
        for (int j = lo; j <= hi; ++j) m = max(m, src[(size_t)j * W + x]);
    }
    out[((size_t)b * H + y) * W + x] = m;
}

// final select + derived tensors. dil: dilated segmentation (or the raw one for k = 0); def8: default mask 0..255.
__global__ void mask_finalize_kernel(const uint8_t *__restrict__ dil, const uint8_t *__restrict__ def8,
                                     const unsigned long long *__restrict__ area, float area_thres, int force_default, int H,
                                     int W, const float *__restrict__ image /* [B,H,W,3] in [-1,1] */, uint8_t *__restrict__ mask_out,
                                     __half *__restrict__ masked_image /* [B,H,W,ldm] */, int ldm, float *__restrict__ mask_small,
                                     int *__restrict__ used_default) {
    pdl_trigger();
    pdl_wait();
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
    if (x >= W) return;
    const bool use_def = force_default || ((float)area[b] < area_thres);            // :1130
    const size_t p = ((size_t)b * H + y) * W + x;
    const uint8_t d = def8[(size_t)y * W + x];
    // default branch: mask = default/255 binarised at 0.5 (:1132, :202-203); adapted: logical_and(dilated, default) (:1136-1137)
    const bool m = use_def ? (d >= 128) : (dil[p] != 0 && d != 0);
    mask_out[p] = m ? 1 : 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) masked_image[p * ldm + c] = __float2half_rn(m ? 0.0f : image[p * 3 + c]);  // image*(mask<0.5)
    if ((x & 7) == 0 && (y & 7) == 0)                                                 // nearest /8 reads pixel (8i, 8j) (:690)
        mask_small[((size_t)b * (H / 8) + (y >> 3)) * (W / 8) + (x >> 3)] = m ? 1.0f : 0.0f;
    if (x == 0 && y == 0 && used_default) used_default[b] = use_def ? 1 : 0;
}

__global__ void image_to_u8_kernel(const float *__restrict__ img, long long rows, int ld, uint8_t *__restrict__ out) {
    pdl_trigger();
    pdl_wait();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * 3) return;
    const float v = fminf(fmaxf(img[(i / 3) * ld + (i % 3)] * 0.5f + 0.5f, 0.0f), 1.0f);
    out[i] = (uint8_t)(v * 255.0f);  // numpy astype(uint8) truncates
}

__global__ void sample_latents_kernel(const float *__restrict__ mean, const float *__restrict__ logvar,
                                      const float *__restrict__ noise, long long n, float scaling, float *__restrict__ out) {
    pdl_trigger();
    pdl_wait();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (mean[i] + expf(0.5f * logvar[i]) * noise[i]) * scaling;
}

}  // namespace coma

using namespace coma;

extern "C" int coma_cfg_ddim_step_f32(const float *eps, int64_t half_rows, int64_t ld, int64_t C, float guidance, const float *x,
                                      double alpha_t, double alpha_prev, float *x_prev, float *x0, coma_stream_t stream) {
    COMA_REQUIRE(eps && x && x_prev && x0, "null pointer");
    COMA_REQUIRE(half_rows > 0 && C > 0 && ld >= C, "bad sizes");
    COMA_REQUIRE(alpha_t > 0.0 && alpha_t <= 1.0 && alpha_prev > 0.0 && alpha_prev <= 1.0, "alphas_cumprod out of range");
    const long long n = half_rows * C;
    launch_pdl(cfg_ddim_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, 
        eps, half_rows, (int)ld, (int)C, guidance, x, (float)sqrt(alpha_t), (float)sqrt(1.0 - alpha_t), (float)sqrt(alpha_prev),
        (float)sqrt(1.0 - alpha_prev), x_prev, x0);
    return check_launch("cfg_ddim_kernel");
}

extern "C" int coma_assemble_unet_input_f16(const float *latents, const float *mask64, const float *masked_latents, int64_t rows,
                                            void *out, int64_t ldo, coma_stream_t stream) {
    COMA_REQUIRE(latents && mask64 && masked_latents && out, "null pointer");
    COMA_REQUIRE(rows > 0 && ldo >= 9, "bad sizes");
    launch_pdl(assemble_input_kernel, dim3((unsigned)((rows * 9 + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, latents, mask64, masked_latents, rows,
                                                                                             (__half *)out, (int)ldo);
    return check_launch("assemble_input_kernel");
}

extern "C" int coma_adaptive_mask_u8(const uint8_t *seg, const uint8_t *default_mask, int64_t B, int64_t H, int64_t W, int dilate_iters,
                                     float area_thres, int force_default, const float *image, uint8_t *scratch, uint8_t *mask_out,
                                     void *masked_image, int64_t ldm, float *mask_small, int *used_default,
                                     unsigned long long *area_ws, coma_stream_t stream) {
    COMA_REQUIRE(seg && default_mask && image && scratch && mask_out && masked_image && mask_small && area_ws, "null pointer");
    COMA_REQUIRE(B > 0 && H > 0 && W > 0 && H % 8 == 0 && W % 8 == 0 && dilate_iters >= 0 && ldm >= 3 && B <= 65535 && H <= 65535,
                 "bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(area_ws, 0, sizeof(unsigned long long) * B, st);
    launch_pdl(mask_area_kernel, dim3(dim3(32, (unsigned)B)), dim3(256), 0, st, seg, (int)(H * W), area_ws);
    if (int e = check_launch("mask_area_kernel")) return e;
    const uint8_t *dil = seg;
    if (dilate_iters > 0) {  // separable (2k+1)^2 box max: rows into scratch[0], columns into scratch[1]
        uint8_t *t0 = scratch, *t1 = scratch + (size_t)B * H * W;
        dim3 grid((unsigned)((W + 127) / 128), (unsigned)H, (unsigned)B);
        launch_pdl(box_max_kernel, dim3(grid), dim3(128), 0, st, seg, t0, (int)H, (int)W, dilate_iters, 0);
        if (int e = check_launch("box_max_kernel")) return e;
        launch_pdl(box_max_kernel, dim3(grid), dim3(128), 0, st, t0, t1, (int)H, (int)W, dilate_iters, 1);
        if (int e = check_launch("box_max_kernel")) return e;
        dil = t1;
    }
    dim3 grid((unsigned)((W + 127) / 128), (unsigned)H, (unsigned)B);
    launch_pdl(mask_finalize_kernel, dim3(grid), dim3(128), 0, st, dil, default_mask, area_ws, area_thres, force_default, (int)H, (int)W, image, mask_out,
                                              (__half *)masked_image, (int)ldm, mask_small, used_default);
    return check_launch("mask_finalize_kernel");
}

extern "C" int coma_image_to_u8(const float *img, int64_t rows, int64_t ld, uint8_t *out, coma_stream_t stream) {
    COMA_REQUIRE(img && out && rows > 0 && ld >= 3, "bad arguments");
    launch_pdl(image_to_u8_kernel, dim3((unsigned)((rows * 3 + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, img, rows, (int)ld, out);
    return check_launch("image_to_u8_kernel");
}

extern "C" int coma_sample_latents_f32(const float *mean, const float *logvar, const float *noise, int64_t n, float scaling,
                                       float *out, coma_stream_t stream) {
    COMA_REQUIRE(mean && logvar && noise && out && n > 0, "bad arguments");
    launch_pdl(sample_latents_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, mean, logvar, noise, n, scaling, out);
    return check_launch("sample_latents_kernel");
}
