// common.cuh -- shared helpers for libsln_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/sln_b200.h"

namespace sln {

// thread-local error text behind sln_last_error_string()
void set_error(const char *fmt, ...);

#define SLN_REQUIRE(cond, code, ...)                 \
    do {                                             \
        if (!(cond)) {                               \
            ::sln::set_error(__VA_ARGS__);           \
            return (code);                           \
        }                                            \
    } while (0)

#define SLN_CUDA_OK(expr)                                                                  \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            ::sln::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),       \
                             __FILE__, __LINE__);                                          \
            return SLN_ERR_CUDA;                                                           \
        }                                                                                  \
    } while (0)

#define SLN_LAUNCH_OK(name)                                                                \
    do {                                                                                   \
        cudaError_t _e = cudaGetLastError();                                               \
        if (_e != cudaSuccess) {                                                           \
            ::sln::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));     \
            return SLN_ERR_CUDA;                                                           \
        }                                                                                  \
    } while (0)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

int sm_count();   // of the current device (not cached)

// ---------------------------------------------------------------------------
// Programmatic dependent launch (sm_90+): a kernel launched with launch_chain(pdl = true) may be scheduled while its
// predecessor in the stream is still running; pdl_prologue() at the very top of BOTH kernels makes the predecessor
// allow that as soon as all of its CTAs are resident and makes the successor block until the predecessor has
// completed and flushed -- only the launch latency and CTA placement overlap, never the bodies.  Without the launch
// attribute both instructions are no-ops.  SLN_PDL=0 in the environment turns the attribute off (A/B).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void pdl_prologue()
{
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

bool pdl_enabled();   // lib.cu: getenv("SLN_PDL") != "0", cached

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                       bool pdl, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl && pdl_enabled()) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------------------
// One axis of the crop sampling grid, bit-compatible with crop_and_resize.c:44-56
// (scale and sample position, un-fused fp32; the single-sample case is evaluated in
// double like the reference's `0.5 * (y1 + y2) * (image_height - 1)`), :58/:80 (range
// test) and :71-73/:89-91 (floorf / ceilf taps, lerp = pos - floor).
// lo == INVALID_TAP marks a sample outside [0, extent-1] (or a NaN position, which is
// undefined behaviour in the reference and is treated as "outside" here).
// hi is not stored: ceilf(pos) == lo + (lerp != 0) exactly, because pos - floorf(pos)
// is exact in fp32.
// ---------------------------------------------------------------------------
constexpr int INVALID_TAP = -(1 << 30);

struct Tap {
    int lo;
    float lerp;
};

__device__ __forceinline__ float axis_scale(float a1, float a2, int extent, int crop)
{
    return crop > 1 ? __fdiv_rn(__fmul_rn(__fsub_rn(a2, a1), (float)(extent - 1)), (float)(crop - 1))
                    : 0.f;
}

__device__ __forceinline__ Tap axis_tap(float a1, float a2, float scale, int extent, int crop, int k)
{
    const float em1 = (float)(extent - 1);
    float pos;
    if (crop > 1) {
        pos = __fadd_rn(__fmul_rn(a1, em1), __fmul_rn((float)k, scale));
    } else {
        pos = (float)(0.5 * (double)__fadd_rn(a1, a2) * (double)(extent - 1));
    }
    Tap t;
    if (!(pos >= 0.f && pos <= em1)) {   // also catches NaN
        t.lo = INVALID_TAP;
        t.lerp = 0.f;
    } else {
        const float fl = floorf(pos);
        t.lo = (int)fl;
        t.lerp = __fsub_rn(pos, fl);
    }
    return t;
}

// bilinear blend with the reference's rounding sequence (crop_and_resize.c:102-106)
__device__ __forceinline__ float lerp2(float tl, float tr, float bl, float br, float xl, float yl)
{
    const float top = __fadd_rn(tl, __fmul_rn(__fsub_rn(tr, tl), xl));
    const float bot = __fadd_rn(bl, __fmul_rn(__fsub_rn(br, bl), xl));
    return __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), yl));
}

template <int VEC> struct VecT;
template <> struct VecT<1> { using type = float; };
template <> struct VecT<4> { using type = float4; };

__device__ __forceinline__ float ldg_vec(const float *p) { return __ldg(p); }
__device__ __forceinline__ float4 ldg_vec(const float4 *p) { return __ldg(p); }

__device__ __forceinline__ float4 make_splat(float v, float4 *) { return make_float4(v, v, v, v); }
__device__ __forceinline__ float make_splat(float v, float *) { return v; }

__device__ __forceinline__ float lerp2v(float tl, float tr, float bl, float br, float xl, float yl)
{
    return lerp2(tl, tr, bl, br, xl, yl);
}
__device__ __forceinline__ float4 lerp2v(float4 tl, float4 tr, float4 bl, float4 br, float xl, float yl)
{
    return make_float4(lerp2(tl.x, tr.x, bl.x, br.x, xl, yl), lerp2(tl.y, tr.y, bl.y, br.y, xl, yl),
                       lerp2(tl.z, tr.z, bl.z, br.z, xl, yl), lerp2(tl.w, tr.w, bl.w, br.w, xl, yl));
}

}  // namespace sln
