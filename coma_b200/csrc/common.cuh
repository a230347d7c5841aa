// Shared helpers for the coma_b200 CUDA kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <utility>

#include "../../include/coma_b200.h"

namespace coma {

constexpr int kNumSM = 148;  // B200

void set_error(const char *fmt, ...);
void count_launch(int n = 1);
void note_kernel(const char *name);

inline int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    count_launch();
    note_kernel(what);
    return 0;
}

#define COMA_REQUIRE(cond, msg)                          \
    do {                                                 \
        if (!(cond)) {                                   \
            coma::set_error("%s: %s", __func__, msg);    \
            return COMA_E_BADARG;                        \
        }                                                \
    } while (0)

// ---- programmatic dependent launch (PDL): the UNet / VAE forward is ~500 short dependent kernels replayed from a CUDA
// graph; with the launch attribute below a kernel's CTAs are scheduled while its predecessor drains, run their prologue
// (barrier init, TMEM allocation, descriptor prefetch) and block in pdl_wait() until the predecessor's memory is visible.
// Every kernel launched through launch_pdl() calls pdl_trigger() and then pdl_wait() before it touches global memory.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    static const bool no_pdl = getenv("COMA_NO_PDL") != nullptr;  // A/B switch: plain stream-ordered launches
    cfg.attrs = attr;
    cfg.numAttrs = no_pdl ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float warp_max_nan(float v) {
    // max that propagates NaN like torch.max
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float u = __shfl_xor_sync(0xffffffffu, v, o);
        v = (u != u || v != v) ? __int_as_float(0x7fc00000) : fmaxf(u, v);
    }
    return v;
}

// Streaming (read-once / write-once) global accesses: keep them out of L1.
__device__ __forceinline__ float ld_stream(const float *p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

}  // namespace coma
