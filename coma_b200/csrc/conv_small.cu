// C1 — 3x3 convolution with a handful of output channels (the VAE decoder's conv_out 128 -> 3, the UNet's conv_out 320 -> 4),
// fused with the GroupNorm affine + SiLU of its input. diffusers layers reached from utils/adaptive_mask_inpainting.py:1086,:1112
// (AutoencoderKL.decode -> decoder.conv_norm_out / conv_act / conv_out) and :1001 (UNet conv_norm_out / conv_act / conv_out).
//
// On the tensor-core implicit GEMM this layer pads N = 3 to a 64-wide tile (13 TFLOP/s, 0.54 ms per B = 4 decode at 512^2) and
// needs a separate pass that writes the normalised + activated 268 MB tensor first (0.17 ms). Here a CTA owns a 32 x 16 pixel tile:
// it stages the 34 x 18 halo in shared memory 32 channels at a time, applying x*scale[b,c] + shift[b,c] -> SiLU on the way in
// (fp16 in smem, zero outside the image = the convolution's zero padding of the ACTIVATED tensor), and every thread accumulates two
// pixels x Cout outputs in fp32 on the CUDA cores; weights sit in shared memory as fp32 and are read as broadcast LDS.128.
// HBM traffic: the input once (x 1.2 for the halo, mostly L2 hits) + a 3-channel output; no intermediate tensor.
#include <cuda_fp16.h>

#include "common.cuh"

namespace coma {

constexpr int CS_TW = 32, CS_TH = 16, CS_CC = 32;            // tile width / height (pixels), channels per smem chunk (54 KB: 4 CTAs/SM)
constexpr int CS_HW = CS_TW + 2, CS_HH = CS_TH + 2;          // halo tile
constexpr int CS_PIX_STRIDE = CS_CC + 8;                     // halves per staged pixel (+16 B: conflict-free LDS.128 across x)
constexpr int CS_NOUT = 4;                                   // accumulators per pixel (Cout <= 4)

__global__ void __launch_bounds__(256)
    conv3x3_small_n_kernel(const __half *__restrict__ x, int B, int H, int W, int C, long long ldx, const float *__restrict__ scale,
                           const float *__restrict__ shift, int act, const __half *__restrict__ wt, long long ldw, int Cout,
                           const float *__restrict__ bias, float *__restrict__ out32, __half *__restrict__ out16, long long ldo) {
    extern __shared__ __align__(16) unsigned char cs_smem[];
    __half *sx = reinterpret_cast<__half *>(cs_smem);                                        // [CS_HH][CS_HW][CS_PIX_STRIDE]
    float *sw = reinterpret_cast<float *>(cs_smem + (size_t)CS_HH * CS_HW * CS_PIX_STRIDE * 2);   // [9][CS_CC][CS_NOUT]
    float *ss = sw + 9 * CS_CC * CS_NOUT;                                                     // scale[CS_CC], shift[CS_CC]
    pdl_trigger();
    pdl_wait();
    const int tiles_x = (W + CS_TW - 1) / CS_TW;
    const int b = blockIdx.y, ty0 = (blockIdx.x / tiles_x) * CS_TH, tx0 = (blockIdx.x % tiles_x) * CS_TW;
    const int tid = threadIdx.x, px = tid & 15, py = tid >> 4;   // thread -> pixels (px, py) and (px + 16, py) of the tile
    float2 acc[2][2];   // [pixel][output pair (n0,n1) / (n2,n3)]: packed FP32x2 accumulation, the input value broadcast
#pragma unroll
    for (int q = 0; q < 2; ++q) acc[q][0] = acc[q][1] = make_float2(0.f, 0.f);

    for (int c0 = 0; c0 < C; c0 += CS_CC) {
        __syncthreads();   // previous chunk fully consumed
        // weights of this channel chunk: sw[tap][c][n] = W[n][tap*C + c0 + c]  (K order (ky, kx, cin), zero for n >= Cout)
        for (int i = tid; i < 9 * CS_CC * CS_NOUT; i += 256) {
            const int n = i % CS_NOUT, c = (i / CS_NOUT) % CS_CC, tap = i / (CS_NOUT * CS_CC);
            sw[i] = (n < Cout && c0 + c < C) ? __half2float(wt[(long long)n * ldw + (long long)tap * C + c0 + c]) : 0.f;
        }
        if (tid < CS_CC) {
            const int c = c0 + tid;
            ss[tid] = (scale && c < C) ? scale[(long long)b * C + c] : 1.f;
            ss[CS_CC + tid] = (shift && c < C) ? shift[(long long)b * C + c] : 0.f;
        }
        __syncthreads();
        // halo tile, 8 channels (16 bytes) per item, affine + activation applied on the way in
        for (int i = tid; i < CS_HH * CS_HW * (CS_CC / 8); i += 256) {
            const int c8 = i % (CS_CC / 8), hx = (i / (CS_CC / 8)) % CS_HW, hy = i / ((CS_CC / 8) * CS_HW);
            const int gy = ty0 + hy - 1, gx = tx0 + hx - 1, c = c0 + c8 * 8;
            uint4 o = make_uint4(0u, 0u, 0u, 0u);
            if (gy >= 0 && gy < H && gx >= 0 && gx < W && c < C) {
                o = *reinterpret_cast<const uint4 *>(x + ((long long)(b * H + gy) * W + gx) * ldx + c);
                if (scale) {
                    __half2 *h2 = reinterpret_cast<__half2 *>(&o);
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        float2 v = __half22float2(h2[t]);
                        v.x = fmaf(v.x, ss[c8 * 8 + 2 * t], ss[CS_CC + c8 * 8 + 2 * t]);
                        v.y = fmaf(v.y, ss[c8 * 8 + 2 * t + 1], ss[CS_CC + c8 * 8 + 2 * t + 1]);
                        if (act == 1) {
                            v.x = __fdividef(v.x, 1.0f + __expf(-v.x));
                            v.y = __fdividef(v.y, 1.0f + __expf(-v.y));
                        }
                        h2[t] = __floats2half2_rn(v.x, v.y);
                    }
                }
            }
            *reinterpret_cast<uint4 *>(sx + ((size_t)hy * CS_HW + hx) * CS_PIX_STRIDE + c8 * 8) = o;
        }
        __syncthreads();
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
            const int dy = tap / 3, dx = tap % 3;
            const __half *p0 = sx + ((size_t)(py + dy) * CS_HW + px + dx) * CS_PIX_STRIDE;
            const __half *p1 = p0 + 16 * CS_PIX_STRIDE;
            const float4 *wp = reinterpret_cast<const float4 *>(sw + (size_t)tap * CS_CC * CS_NOUT);
#pragma unroll 2
            for (int c8 = 0; c8 < CS_CC / 8; ++c8) {
                const uint4 a = *reinterpret_cast<const uint4 *>(p0 + c8 * 8), bq = *reinterpret_cast<const uint4 *>(p1 + c8 * 8);
                const __half2 *ah = reinterpret_cast<const __half2 *>(&a), *bh = reinterpret_cast<const __half2 *>(&bq);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const float2 av = __half22float2(ah[t]), bv = __half22float2(bh[t]);
                    const float4 w0 = wp[c8 * 8 + 2 * t], w1 = wp[c8 * 8 + 2 * t + 1];   // weights of channels 2t, 2t+1: (n0..n3)
                    const float2 w0a = make_float2(w0.x, w0.y), w0b = make_float2(w0.z, w0.w), w1a = make_float2(w1.x, w1.y), w1b = make_float2(w1.z, w1.w);
                    acc[0][0] = __ffma2_rn(make_float2(av.x, av.x), w0a, acc[0][0]);
                    acc[0][1] = __ffma2_rn(make_float2(av.x, av.x), w0b, acc[0][1]);
                    acc[0][0] = __ffma2_rn(make_float2(av.y, av.y), w1a, acc[0][0]);
                    acc[0][1] = __ffma2_rn(make_float2(av.y, av.y), w1b, acc[0][1]);
                    acc[1][0] = __ffma2_rn(make_float2(bv.x, bv.x), w0a, acc[1][0]);
                    acc[1][1] = __ffma2_rn(make_float2(bv.x, bv.x), w0b, acc[1][1]);
                    acc[1][0] = __ffma2_rn(make_float2(bv.y, bv.y), w1a, acc[1][0]);
                    acc[1][1] = __ffma2_rn(make_float2(bv.y, bv.y), w1b, acc[1][1]);
                }
            }
        }
    }
    const int gy = ty0 + py;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int gx = tx0 + px + 16 * q;
        if (gy < H && gx < W) {
            const long long row = ((long long)(b * H + gy) * W + gx) * ldo;
            const float a4[4] = {acc[q][0].x, acc[q][0].y, acc[q][1].x, acc[q][1].y};
            for (int n = 0; n < Cout; ++n) {
                const float v = a4[n] + (bias ? bias[n] : 0.f);
                if (out32) out32[row + n] = v;
                if (out16) out16[row + n] = __float2half_rn(v);
            }
        }
    }
}

}  // namespace coma

extern "C" int coma_conv3x3_small_n_f16(const void *x, int64_t B, int64_t H, int64_t W, int64_t C, int64_t ldx, const float *scale,
                                        const float *shift, int act, const void *Wt, int64_t ldw, int64_t Cout, const float *bias,
                                        float *out_f32, void *out_f16, int64_t ldo, coma_stream_t stream) {
    using namespace coma;
    COMA_REQUIRE(x && Wt && (out_f32 || out_f16), "null pointer");
    COMA_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0 && ldx % 8 == 0 && ldx >= C, "bad input shape (C, ldx multiples of 8)");
    COMA_REQUIRE(Cout >= 1 && Cout <= CS_NOUT && ldw >= 9 * C && ldo >= Cout, "Cout must be 1..4");
    COMA_REQUIRE((scale == nullptr) == (shift == nullptr), "scale and shift come together");
    COMA_REQUIRE(act == 0 || act == 1, "act must be 0 or 1 (SiLU on the normalised input)");
    COMA_REQUIRE((uintptr_t)x % 16 == 0 && B <= 65535, "x must be 16-byte aligned");
    const size_t smem = (size_t)CS_HH * CS_HW * CS_PIX_STRIDE * 2 + (size_t)9 * CS_CC * CS_NOUT * 4 + 2 * CS_CC * 4;
    static bool attr[16] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 16 && !attr[dev]) {
        cudaError_t e = cudaFuncSetAttribute(conv3x3_small_n_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(conv3x3_small_n_kernel): %s", cudaGetErrorString(e));
            return (int)e;
        }
        attr[dev] = true;
    }
    const int64_t tiles = ((W + CS_TW - 1) / CS_TW) * ((H + CS_TH - 1) / CS_TH);
    launch_pdl(conv3x3_small_n_kernel, dim3((unsigned)tiles, (unsigned)B), dim3(256), smem, (cudaStream_t)stream, (const __half *)x, (int)B, (int)H,
               (int)W, (int)C, (long long)ldx, scale, shift, act, (const __half *)Wt, (long long)ldw, (int)Cout, bias, out_f32, (__half *)out_f16,
               (long long)ldo);
    return check_launch("conv3x3_small_n_kernel");
}
