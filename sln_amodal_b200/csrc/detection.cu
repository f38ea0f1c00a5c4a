// detection.cu -- the elementwise front of refine_detections (reference modal/Functions.py:453-557) in one launch.
//
// Per ROI: class = argmax of the class probabilities (:468, first maximum), class-specific deltas scaled by
// RPN_BBOX_STD_DEV (:436-450), apply_box_deltas in normalised coordinates (:77-98), scale to pixels, clip to the
// window (:423-433), round half-to-even (:485), and the keep filter `class_id > 0 [and score >= min_confidence]`
// (:488-493).  The reference spends ~25 small torch kernels on this; every operation below is rounded separately in the
// same order, `exp` goes through double (within 1 ulp of torch's CPU exp, like proposal.cu).
//
// Output is laid out for sln_nms: dets [N][5] = (y1, x1, y2, x2, score) and a class id per box.  Boxes that fail the
// filter get score -inf and a unique negative class, so the class-aware NMS keeps them (nothing shares their class) at
// the very end of its score-ordered output; *n_excluded says how many to drop from the tail.
#include "common.cuh"

#include <math.h>

namespace sln {

__global__ void __launch_bounds__(128)
refine_decode_kernel(const float *__restrict__ rois, const float *__restrict__ probs, const float *__restrict__ deltas,
                     int N, int K, float s0, float s1, float s2, float s3, float img_h, float img_w, float wy1,
                     float wx1, float wy2, float wx2, float min_conf, int use_min_conf, float *__restrict__ dets,
                     int *__restrict__ cls_nms, int *__restrict__ class_ids, int *__restrict__ n_excluded)
{
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);      // warp per ROI: the K probabilities are one row
    if (i >= N) return;
    // argmax, first maximum (torch.max on the CPU reference); NaN never wins
    float best = -INFINITY;
    int arg = 0x7fffffff;
    for (int k = lane; k < K; k += 32) {
        const float p = __ldg(probs + (size_t)i * K + k);
        if (p > best) { best = p; arg = k; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
        if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
    }
    if (arg == 0x7fffffff) arg = 0;                                        // all NaN / -inf: class 0 (filtered out)
    if (lane != 0) return;
    const float score = __ldg(probs + (size_t)i * K + arg);
    const float4 r = *reinterpret_cast<const float4 *>(rois + 4 * (size_t)i);
    const float4 dl = *reinterpret_cast<const float4 *>(deltas + ((size_t)i * K + arg) * 4);
    const float dy = __fmul_rn(dl.x, s0), dx = __fmul_rn(dl.y, s1), dh = __fmul_rn(dl.z, s2), dw = __fmul_rn(dl.w, s3);
    float height = __fsub_rn(r.z, r.x), width = __fsub_rn(r.w, r.y);
    float cy = __fadd_rn(r.x, __fmul_rn(0.5f, height)), cx = __fadd_rn(r.y, __fmul_rn(0.5f, width));
    cy = __fadd_rn(cy, __fmul_rn(dy, height));
    cx = __fadd_rn(cx, __fmul_rn(dx, width));
    height = __fmul_rn(height, (float)exp((double)dh));
    width = __fmul_rn(width, (float)exp((double)dw));
    float y1 = __fsub_rn(cy, __fmul_rn(0.5f, height)), x1 = __fsub_rn(cx, __fmul_rn(0.5f, width));
    float y2 = __fadd_rn(y1, height), x2 = __fadd_rn(x1, width);
    y1 = __fmul_rn(y1, img_h); x1 = __fmul_rn(x1, img_w); y2 = __fmul_rn(y2, img_h); x2 = __fmul_rn(x2, img_w);
    // torch.clamp(min, max): NaN passes through
    auto clampf = [](float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); };
    y1 = rintf(clampf(y1, wy1, wy2)); x1 = rintf(clampf(x1, wx1, wx2));
    y2 = rintf(clampf(y2, wy1, wy2)); x2 = rintf(clampf(x2, wx1, wx2));
    const bool keep = arg > 0 && (!use_min_conf || score >= min_conf);
    float *o = dets + 5 * (size_t)i;
    o[0] = y1; o[1] = x1; o[2] = y2; o[3] = x2;
    o[4] = keep ? score : -INFINITY;
    cls_nms[i] = keep ? arg : -(i + 1);
    class_ids[i] = arg;
    if (!keep) atomicAdd(n_excluded, 1);                                    // integer count: order-independent
}

}  // namespace sln

extern "C" int sln_refine_decode(const float *rois, const float *probs, const float *deltas, int N, int K,
                                 const float *std_dev_host, float img_h, float img_w, const float *window_host,
                                 float min_confidence, float *dets, int *cls_nms, int *class_ids, int *n_excluded,
                                 void *stream)
{
    SLN_REQUIRE(N >= 0 && K >= 1, SLN_ERR_ARG, "need N >= 0 and K >= 1 (got %d, %d)", N, K);
    SLN_REQUIRE(n_excluded && std_dev_host && window_host, SLN_ERR_ARG, "null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SLN_CUDA_OK(cudaMemsetAsync(n_excluded, 0, sizeof(int), st));
    if (N == 0) return SLN_OK;
    SLN_REQUIRE(rois && probs && deltas && dets && cls_nms && class_ids, SLN_ERR_ARG, "null pointer");
    SLN_REQUIRE((reinterpret_cast<uintptr_t>(rois) & 15u) == 0 && (reinterpret_cast<uintptr_t>(deltas) & 15u) == 0,
                SLN_ERR_LAYOUT, "rois / deltas must be 16-byte aligned");
    const int use_min = min_confidence != 0.f;                              // `if config.DETECTION_MIN_CONFIDENCE:` (:490)
    sln::refine_decode_kernel<<<sln::cdiv(N, 4), 128, 0, st>>>(rois, probs, deltas, N, K, std_dev_host[0], std_dev_host[1],
                                                               std_dev_host[2], std_dev_host[3], img_h, img_w,
                                                               window_host[0], window_host[1], window_host[2],
                                                               window_host[3], min_confidence, use_min, dets, cls_nms,
                                                               class_ids, n_excluded);
    SLN_LAUNCH_OK("refine_decode_kernel");
    return SLN_OK;
}
