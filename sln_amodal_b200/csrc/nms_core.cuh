// nms_core.cuh -- internal interface shared by nms.cu (the sln_nms entry point) and
// proposal.cu (which feeds already-sorted boxes straight into the mask + scan stages).
#pragma once

#include "common.cuh"

namespace sln {

// Workspace carved up for one NMS problem of n boxes.
struct NmsBuffers {
    float4 *boxes;                 // [n] (c0,c1,c2,c3) in visiting (score) order
    float *areas;                  // [n] (c3-c1+1)*(c2-c0+1), pth_nms.py:16
    int *cls;                      // [n] class id in visiting order (only if class-aware)
    int *order;                    // [n] original index of the box at each visiting position
    int *rank;                     // [n] visiting position of each original box
    int *tickets;                  // [ceil(n/256)] row-tile tickets of the rank kernel's fused gather
    unsigned long long *mask;      // [n][W] IoU>=thresh bitmask, W = ceil(n/64)
    void *stage;                   // scratch of the two-stage pipeline (see nms.cu)
    void *fix;                     // scratch of the parallel fixed-point resolve
    void *sparse;                  // scratch of the sparse (binned) path
};

size_t nms_buffers_bytes(int n);
void nms_carve(void *ws, int n, NmsBuffers &b);

// Mask + greedy scan over boxes that are ALREADY in visiting order.
//   order == nullptr : kept entries are visiting positions themselves
//   keep64 / keep32  : either may be nullptr
int nms_sorted_launch(const float4 *boxes, const float *areas, const int *cls, const int *order, int n,
                      float thresh, int max_keep, unsigned long long *mask, int64_t *keep64, int *keep32,
                      int *num_keep, cudaStream_t st, void *stage, void *fix, void *sparse, bool sparse_only = false);

// Total order used everywhere a "descending score, stable" sort is needed:
// larger key first; NaN sorts as the largest value (torch.sort's convention);
// -0.0 == +0.0.  Ties are broken by ascending index.
__device__ __forceinline__ unsigned score_key(float s)
{
    unsigned u = __float_as_uint(s);
    if (s != s) return 0xffffffffu;
    if (u == 0x80000000u) u = 0u;
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// rank[i] = number of elements that precede element i in (score desc, tie id asc) order,
// score of element i = scores[i*stride], tie id = tie_ids[i] (or i when tie_ids is null;
// tie ids must be distinct).  rank must be zero-filled before the launch.
// O(n^2) compares spread over the whole grid: a stable, deterministic sort for the
// n <= ~16k this path sees, with no multi-pass radix machinery.
int rank_sort_launch(const float *scores, int stride, const int *tie_ids, int n, int *rank, cudaStream_t st);

}  // namespace sln
