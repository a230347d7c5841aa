// Pipe-throughput micro-benchmark for B200 (sm_100a): cycles per warp-instruction per SM sub-partition for the
// instruction classes the K3 kernel is made of. Build & run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench tools/ubench_pipes.cu && /tmp/ubench
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;
constexpr int CH = 8;  // independent chains per thread

#define BODY(NAME, DECL, STEP, SINK)                                                         \
    __global__ void NAME(float *out, float seed) {                                           \
        DECL;                                                                                \
        for (int i = 0; i < ITERS; ++i) {                                                    \
            _Pragma("unroll") for (int c = 0; c < CH; ++c) { STEP; }                         \
        }                                                                                    \
        float acc = 0.f;                                                                     \
        _Pragma("unroll") for (int c = 0; c < CH; ++c) acc += SINK;                          \
        if (acc == 123.456f) out[0] = acc;                                                   \
    }

BODY(k_ffma, float x[CH]; for (int c = 0; c < CH; ++c) x[c] = seed + c, x[c] = fmaf(x[c], 0.999f, 0.001f), x[c])
BODY(k_ffma2, float2 x[CH]; for (int c = 0; c < CH; ++c) x[c] = make_float2(seed + c, seed - c),
     x[c] = __ffma2_rn(x[c], make_float2(0.999f, 0.999f), make_float2(0.001f, 0.001f)), (x[c].x + x[c].y))
BODY(k_ex2, float x[CH]; for (int c = 0; c < CH; ++c) x[c] = seed * 0.01f + c * 0.001f,
     asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[c])), x[c])
BODY(k_sqrt, float x[CH]; for (int c = 0; c < CH; ++c) x[c] = seed + c,
     asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(x[c])), x[c])
BODY(k_rsqrt, float x[CH]; for (int c = 0; c < CH; ++c) x[c] = seed + c,
     asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(x[c])), x[c])
BODY(k_fmnmx, float x[CH]; for (int c = 0; c < CH; ++c) x[c] = seed + c, x[c] = fminf(fabsf(x[c]) , seed + i), x[c])
BODY(k_lop3, float x[CH]; for (int c = 0; c < CH; ++c) x[c] = seed + c,
     x[c] = __int_as_float((__float_as_int(x[c]) & 0x7fffffff) ^ (i << 3)), x[c])
// mix: 4 FFMA2 + 1 EX2 per chain step (K3-like ratio 7.5 : 2 is approximated by 8 FFMA2 + 2 MUFU below)
BODY(k_mix_ffma2_ex2, float2 x[CH]; for (int c = 0; c < CH; ++c) x[c] = make_float2(seed * 0.01f + c, seed * 0.01f - c),
     { x[c] = __ffma2_rn(x[c], make_float2(0.999f, 0.999f), make_float2(0.001f, 0.001f));
       x[c] = __ffma2_rn(x[c], make_float2(0.998f, 0.998f), make_float2(0.002f, 0.002f));
       x[c] = __ffma2_rn(x[c], make_float2(0.997f, 0.997f), make_float2(0.003f, 0.003f));
       x[c] = __ffma2_rn(x[c], make_float2(0.996f, 0.996f), make_float2(0.004f, 0.004f));
       asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[c].x)); }, (x[c].x + x[c].y))
BODY(k_mix_ffma_ex2, float x[CH]; for (int c = 0; c < CH; ++c) x[c] = seed * 0.01f + c,
     { x[c] = fmaf(x[c], 0.999f, 0.001f); x[c] = fmaf(x[c], 0.998f, 0.002f); x[c] = fmaf(x[c], 0.997f, 0.003f);
       x[c] = fmaf(x[c], 0.996f, 0.004f); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[c])); }, x[c])

template <typename K>
void run(const char *name, K kern, int instr_per_step, int warps_per_sm) {
    float *out;
    cudaMalloc(&out, 4);
    int blocks = 148, threads = warps_per_sm * 32;
    kern<<<blocks, threads>>>(out, 1.0f);
    cudaDeviceSynchronize();
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a);
    kern<<<blocks, threads>>>(out, 1.0f);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    int clk_khz;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double cycles = ms * 1e-3 * clk_khz * 1e3;
    double warp_instr_per_smsp = (double)ITERS * CH * instr_per_step * warps_per_sm / 4.0;
    printf("%-18s warps/SM=%2d  %.3f ms  cycles/warp-instr/SMSP = %.2f  (assuming %d kHz)\n", name, warps_per_sm, ms,
           cycles / warp_instr_per_smsp, clk_khz);
    cudaFree(out);
}

int main() {
    for (int w : {8, 16, 32}) {
        run("FFMA", k_ffma, 1, w);
        run("FFMA2", k_ffma2, 1, w);
        run("MUFU.EX2", k_ex2, 1, w);
        run("MUFU.SQRT", k_sqrt, 1, w);
        run("MUFU.RSQ", k_rsqrt, 1, w);
        run("FMNMX", k_fmnmx, 1, w);
        run("LOP3x2", k_lop3, 2, w);
        run("4FFMA2+EX2 (per 5)", k_mix_ffma2_ex2, 5, w);
        run("4FFMA+EX2 (per 5)", k_mix_ffma_ex2, 5, w);
    }
    return 0;
}
