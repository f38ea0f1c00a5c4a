// rpn.cu -- RPN output re-layout in front of proposal_layer.
//
// Reference: RPN.forward (modal/modals.py:388-412) turns each level's conv outputs [B, 2a, H, W] / [B, 4a, H, W] into
// [B, H*W*a, 2] / [B, H*W*a, 4] with permute(0,2,3,1).contiguous().view and applies Softmax(dim=2); MaskRCNN.predict
// (model.py:553-563) then concatenates the five levels along dim 1.  That is 10 permute copies, 5 softmax launches and
// 3 concatenations of 262 k x 6 floats per image.  Anchor (y, x, k) of a level sits at row (y*W + x)*a + k, and its
// value j comes from channel 2k + j (4k + j for the deltas), so per level the whole thing is one NCHW -> NHWC
// transpose written at the level's row offset.  One launch does all levels and all three outputs; the inverse launch
// scatters the gradients of the logits and deltas back to the conv outputs for the training step.
//
// Softmax: p_j = exp(l_j - max(l_0, l_1)) / (exp(l_0 - max) + exp(l_1 - max)), the same expression and operation
// order as torch's softmax kernels (float accumulation), with full-precision expf.
#include "common.cuh"

namespace sln {

constexpr int RPN_MAX_LEVELS = 8;
constexpr int RPN_MAX_A = 8;

struct RpnLevels {
    const float *logits[RPN_MAX_LEVELS];
    const float *bbox[RPN_MAX_LEVELS];
    float *d_logits[RPN_MAX_LEVELS];      // gradient launch: destinations
    float *d_bbox[RPN_MAX_LEVELS];
    int hw[RPN_MAX_LEVELS];
    int pos0[RPN_MAX_LEVELS + 1];         // first position of each level in the concatenation; [n] = total
    int n;
};

__device__ __forceinline__ int rpn_find_level(const RpnLevels &lv, int p)
{
    int l = 0;
#pragma unroll
    for (int k = 1; k < RPN_MAX_LEVELS; ++k)
        if (k < lv.n && p >= lv.pos0[k]) l = k;
    return l;
}

// thread = one position (y, x) of one level of one image: C = 2a logits + 4a deltas
template <int A_T>
__global__ void __launch_bounds__(256)
rpn_pack_kernel(RpnLevels lv, int B, int a_rt, int nhwc, float *__restrict__ out_logits, float *__restrict__ out_probs,
                float *__restrict__ out_bbox)
{
    const int a = A_T > 0 ? A_T : a_rt;
    const int P = lv.pos0[lv.n];
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)B * P) return;
    const int b = (int)(t / P), p = (int)(t - (long long)b * P);
    const int l = rpn_find_level(lv, p);
    const int q = p - lv.pos0[l], hw = lv.hw[l];
    const size_t row = ((size_t)b * P + p) * a;                       // first anchor row of this position
    if (lv.logits[l]) {
        const float *src = lv.logits[l] + (size_t)b * 2 * a * hw;
        const size_t sc = nhwc ? 1 : (size_t)hw, sp = nhwc ? (size_t)(2 * a) : 1;
        float2 v[A_T > 0 ? A_T : RPN_MAX_A];
#pragma unroll
        for (int k = 0; k < (A_T > 0 ? A_T : RPN_MAX_A); ++k)
            if (k < a) v[k] = make_float2(__ldg(src + (2 * k) * sc + q * sp), __ldg(src + (2 * k + 1) * sc + q * sp));
#pragma unroll
        for (int k = 0; k < (A_T > 0 ? A_T : RPN_MAX_A); ++k) {
            if (k >= a) break;
            if (out_logits) reinterpret_cast<float2 *>(out_logits)[row + k] = v[k];
            if (out_probs) {
                const float m = fmaxf(v[k].x, v[k].y);
                const float e0 = expf(__fsub_rn(v[k].x, m)), e1 = expf(__fsub_rn(v[k].y, m));
                const float s = __fadd_rn(e0, e1);
                reinterpret_cast<float2 *>(out_probs)[row + k] = make_float2(__fdiv_rn(e0, s), __fdiv_rn(e1, s));
            }
        }
    }
    if (lv.bbox[l] && out_bbox) {
        const float *src = lv.bbox[l] + (size_t)b * 4 * a * hw;
        const size_t sc = nhwc ? 1 : (size_t)hw, sp = nhwc ? (size_t)(4 * a) : 1;
        float4 v[A_T > 0 ? A_T : RPN_MAX_A];
#pragma unroll
        for (int k = 0; k < (A_T > 0 ? A_T : RPN_MAX_A); ++k)
            if (k < a)
                v[k] = make_float4(__ldg(src + (4 * k) * sc + q * sp), __ldg(src + (4 * k + 1) * sc + q * sp),
                                   __ldg(src + (4 * k + 2) * sc + q * sp), __ldg(src + (4 * k + 3) * sc + q * sp));
#pragma unroll
        for (int k = 0; k < (A_T > 0 ? A_T : RPN_MAX_A); ++k) {
            if (k >= a) break;
            reinterpret_cast<float4 *>(out_bbox)[row + k] = v[k];
        }
    }
}

// inverse: gradients w.r.t. the packed logits / deltas -> gradients of the conv outputs (every element written once)
__global__ void __launch_bounds__(256)
rpn_unpack_kernel(RpnLevels lv, int B, int a, int nhwc, const float *__restrict__ g_logits, const float *__restrict__ g_bbox)
{
    const int P = lv.pos0[lv.n];
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)B * P) return;
    const int b = (int)(t / P), p = (int)(t - (long long)b * P);
    const int l = rpn_find_level(lv, p);
    const int q = p - lv.pos0[l], hw = lv.hw[l];
    const size_t row = ((size_t)b * P + p) * a;
    if (lv.d_logits[l]) {
        float *dst = lv.d_logits[l] + (size_t)b * 2 * a * hw;
        const size_t sc = nhwc ? 1 : (size_t)hw, sp = nhwc ? (size_t)(2 * a) : 1;
        for (int k = 0; k < a; ++k) {
            const float2 g = g_logits ? __ldg(reinterpret_cast<const float2 *>(g_logits) + row + k) : make_float2(0.f, 0.f);
            dst[(2 * k) * sc + q * sp] = g.x;
            dst[(2 * k + 1) * sc + q * sp] = g.y;
        }
    }
    if (lv.d_bbox[l]) {
        float *dst = lv.d_bbox[l] + (size_t)b * 4 * a * hw;
        const size_t sc = nhwc ? 1 : (size_t)hw, sp = nhwc ? (size_t)(4 * a) : 1;
        for (int k = 0; k < a; ++k) {
            const float4 g = g_bbox ? __ldg(reinterpret_cast<const float4 *>(g_bbox) + row + k) : make_float4(0.f, 0.f, 0.f, 0.f);
            dst[(4 * k) * sc + q * sp] = g.x;
            dst[(4 * k + 1) * sc + q * sp] = g.y;
            dst[(4 * k + 2) * sc + q * sp] = g.z;
            dst[(4 * k + 3) * sc + q * sp] = g.w;
        }
    }
}

static int rpn_levels(RpnLevels &lv, const int *heights, const int *widths, int n_levels, int B, int a, long long *threads)
{
    SLN_REQUIRE(n_levels >= 1 && n_levels <= RPN_MAX_LEVELS, SLN_ERR_ARG, "rpn: 1..%d levels (got %d)", RPN_MAX_LEVELS, n_levels);
    SLN_REQUIRE(B >= 0 && a >= 1 && a <= RPN_MAX_A, SLN_ERR_ARG, "rpn: 1..%d anchors per location (got %d)", RPN_MAX_A, a);
    SLN_REQUIRE(heights && widths, SLN_ERR_ARG, "null pointer");
    lv.n = n_levels;
    long long pos = 0;
    for (int l = 0; l < n_levels; ++l) {
        SLN_REQUIRE(heights[l] >= 0 && widths[l] >= 0, SLN_ERR_ARG, "negative level size");
        lv.pos0[l] = (int)pos;
        lv.hw[l] = heights[l] * widths[l];
        pos += (long long)heights[l] * widths[l];
        SLN_REQUIRE(pos * a < (1ll << 31) / 4, SLN_ERR_ARG, "rpn: too many anchors");
    }
    for (int l = n_levels; l <= RPN_MAX_LEVELS; ++l) lv.pos0[l] = (int)pos;
    lv.pos0[n_levels] = (int)pos;
    *threads = pos * B;
    return SLN_OK;
}

}  // namespace sln

using namespace sln;

extern "C" int sln_rpn_pack(const float *const *logits, const float *const *bbox, const int *heights, const int *widths,
                            int n_levels, int B, int a, int layout, float *out_logits, float *out_probs, float *out_bbox,
                            void *stream)
{
    RpnLevels lv = {};
    long long threads = 0;
    int rc = rpn_levels(lv, heights, widths, n_levels, B, a, &threads);
    if (rc != SLN_OK) return rc;
    SLN_REQUIRE(layout == SLN_LAYOUT_NCHW || layout == SLN_LAYOUT_NHWC, SLN_ERR_LAYOUT, "rpn: bad layout %d", layout);
    if (threads == 0) return SLN_OK;
    SLN_REQUIRE(logits || bbox, SLN_ERR_ARG, "null pointer");
    for (int l = 0; l < n_levels; ++l) {
        lv.logits[l] = logits ? logits[l] : nullptr;
        lv.bbox[l] = bbox ? bbox[l] : nullptr;
    }
    SLN_REQUIRE((reinterpret_cast<uintptr_t>(out_logits) & 7u) == 0 && (reinterpret_cast<uintptr_t>(out_probs) & 7u) == 0 &&
                    (reinterpret_cast<uintptr_t>(out_bbox) & 15u) == 0,
                SLN_ERR_LAYOUT, "rpn outputs must be 8 / 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const unsigned blocks = (unsigned)((threads + 255) / 256);
    if (a == 3)
        rpn_pack_kernel<3><<<blocks, 256, 0, st>>>(lv, B, a, layout == SLN_LAYOUT_NHWC, out_logits, out_probs, out_bbox);
    else
        rpn_pack_kernel<0><<<blocks, 256, 0, st>>>(lv, B, a, layout == SLN_LAYOUT_NHWC, out_logits, out_probs, out_bbox);
    SLN_LAUNCH_OK("rpn_pack_kernel");
    return SLN_OK;
}

extern "C" int sln_rpn_unpack_grads(const float *g_logits, const float *g_bbox, const int *heights, const int *widths,
                                    int n_levels, int B, int a, int layout, float *const *d_logits, float *const *d_bbox,
                                    void *stream)
{
    RpnLevels lv = {};
    long long threads = 0;
    int rc = rpn_levels(lv, heights, widths, n_levels, B, a, &threads);
    if (rc != SLN_OK) return rc;
    SLN_REQUIRE(layout == SLN_LAYOUT_NCHW || layout == SLN_LAYOUT_NHWC, SLN_ERR_LAYOUT, "rpn: bad layout %d", layout);
    if (threads == 0) return SLN_OK;
    for (int l = 0; l < n_levels; ++l) {
        lv.d_logits[l] = d_logits ? d_logits[l] : nullptr;
        lv.d_bbox[l] = d_bbox ? d_bbox[l] : nullptr;
    }
    SLN_REQUIRE((reinterpret_cast<uintptr_t>(g_logits) & 7u) == 0 && (reinterpret_cast<uintptr_t>(g_bbox) & 15u) == 0, SLN_ERR_LAYOUT,
                "rpn gradients must be 8 / 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    rpn_unpack_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(lv, B, a, layout == SLN_LAYOUT_NHWC, g_logits, g_bbox);
    SLN_LAUNCH_OK("rpn_unpack_kernel");
    return SLN_OK;
}
