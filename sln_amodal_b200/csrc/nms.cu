// nms.cu -- greedy NMS on the device, bit-compatible with the reference CPU extension.
//
// Reference semantics: nms/src/nms.c:4-69 (cpu_nms) behind nms/pth_nms.py:5-24.
// GPU baseline being replaced: nms/src/cuda/nms_kernel.cu:26-70 (full n x n mask incl.
// the useless lower triangle, ">" instead of ">=", FMA-contracted) followed by a
// synchronous D2H copy of the mask and a serial CPU scan (nms/src/nms_cuda.c:33-58).
//
// Pipeline (all on the caller's stream, no host round trip):
//   1. rank_kernel      stable rank of every box by (score desc, index asc)   O(n^2) compares, whole grid
//   2. gather_kernel    boxes/areas/classes permuted into visiting order
//   3. mask_kernel      upper-triangle 64x64 tiles of the "IoU >= thresh" bit matrix.  The
//                       IoU test reproduces nms.c's un-fused fp32 sequence; the divide is
//                       replaced by an exact sign test with a guarded fall-back to the
//                       IEEE divide (see iou_ge below)
//   4. scan_kernel      one CTA walks the 64-box chunks in order: a warp resolves each
//                       chunk against its diagonal tile with shuffles, the other warps
//                       OR the kept rows into the `removed` words in shared memory
#include "nms_core.cuh"

namespace sln {

// ---------------------------------------------------------------------------
// 1. rank sort
// ---------------------------------------------------------------------------
constexpr int RANK_ROWS = 256;     // elements ranked per CTA (one per thread)
constexpr int RANK_COLS = 1024;    // elements compared against per CTA (staged in smem)

// MODE 0: every column index is below every row index of this CTA  -> count key_j >= key_i
// MODE 1: every column index is above every row index              -> count key_j >  key_i
// MODE 2: index ranges overlap, or explicit tie ids                 -> full (key, tie) compare
template <int MODE>
__device__ __forceinline__ int rank_count(const uint4 *__restrict__ k4, const int4 *__restrict__ t4, int n4,
                                          unsigned ki, int ti)
{
    int cnt = 0;
#pragma unroll 4
    for (int q = 0; q < n4; ++q) {
        const uint4 k = k4[q];      // broadcast 16-byte shared load: four keys per instruction
        if (MODE == 0) {
            cnt += (k.x >= ki) + (k.y >= ki) + (k.z >= ki) + (k.w >= ki);
        } else if (MODE == 1) {
            cnt += (k.x > ki) + (k.y > ki) + (k.z > ki) + (k.w > ki);
        } else {
            const int4 t = t4[q];
            cnt += (k.x > ki || (k.x == ki && t.x < ti)) + (k.y > ki || (k.y == ki && t.y < ti)) +
                   (k.z > ki || (k.z == ki && t.z < ti)) + (k.w > ki || (k.w == ki && t.w < ti));
        }
    }
    return cnt;
}

__global__ void __launch_bounds__(RANK_ROWS)
rank_kernel(const float *__restrict__ scores, int stride, const int *__restrict__ tie_ids, int n,
            int *__restrict__ rank)
{
    __shared__ __align__(16) unsigned s_key[RANK_COLS];
    __shared__ __align__(16) int s_tie[RANK_COLS];
    const int i0 = blockIdx.x * RANK_ROWS, i = i0 + threadIdx.x;
    const int j0 = blockIdx.y * RANK_COLS;
    const int jn = min(RANK_COLS, n - j0);
    // pad the tile to a multiple of 4 with entries that never count (key 0 with the largest tie id
    // loses every compare except against key 0 rows in MODE 0, handled by clamping below)
    for (int t = threadIdx.x; t < RANK_COLS; t += RANK_ROWS) {
        unsigned k = 0u;
        int tie = 0x7fffffff;
        if (t < jn) {
            k = score_key(__ldg(scores + (size_t)(j0 + t) * stride));
            tie = tie_ids ? __ldg(tie_ids + j0 + t) : (j0 + t);
        }
        s_key[t] = k;
        s_tie[t] = tie;
    }
    __syncthreads();
    if (i >= n) return;
    const unsigned ki = score_key(__ldg(scores + (size_t)i * stride));
    const int ti = tie_ids ? __ldg(tie_ids + i) : i;
    const int n4 = (jn + 3) >> 2, pad = n4 * 4 - jn;
    const uint4 *k4 = reinterpret_cast<const uint4 *>(s_key);
    const int4 *t4 = reinterpret_cast<const int4 *>(s_tie);
    int cnt;
    if (tie_ids == nullptr && j0 + jn <= i0) {
        cnt = rank_count<0>(k4, t4, n4, ki, ti);
        if (ki == 0u) cnt -= pad;                      // padded keys (0) satisfy 0 >= 0
    } else if (tie_ids == nullptr && j0 >= i0 + RANK_ROWS) {
        cnt = rank_count<1>(k4, t4, n4, ki, ti);
    } else {
        cnt = rank_count<2>(k4, t4, n4, ki, ti);
    }
    if (cnt) atomicAdd(rank + i, cnt);
}

int rank_sort_launch(const float *scores, int stride, const int *tie_ids, int n, int *rank, cudaStream_t st)
{
    if (n <= 0) return SLN_OK;
    dim3 grid(cdiv(n, RANK_ROWS), cdiv(n, RANK_COLS));
    SLN_REQUIRE(grid.y <= 65535, SLN_ERR_ARG, "rank sort: n=%d too large", n);
    rank_kernel<<<grid, RANK_ROWS, 0, st>>>(scores, stride, tie_ids, n, rank);
    SLN_LAUNCH_OK("rank_kernel");
    return SLN_OK;
}

// ---------------------------------------------------------------------------
// 2. gather into visiting order
// ---------------------------------------------------------------------------
__global__ void nms_gather_kernel(const float *__restrict__ dets, const int *__restrict__ class_ids,
                                  const int *__restrict__ rank, int n, float4 *__restrict__ boxes,
                                  float *__restrict__ areas, int *__restrict__ cls, int *__restrict__ order)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float c0 = dets[5 * (size_t)i + 0], c1 = dets[5 * (size_t)i + 1];
    const float c2 = dets[5 * (size_t)i + 2], c3 = dets[5 * (size_t)i + 3];
    const int r = rank[i];
    boxes[r] = make_float4(c0, c1, c2, c3);
    // pth_nms.py:16  areas = (x2 - x1 + 1) * (y2 - y1 + 1), one rounding per op
    areas[r] = __fmul_rn(__fadd_rn(__fsub_rn(c3, c1), 1.f), __fadd_rn(__fsub_rn(c2, c0), 1.f));
    order[r] = i;
    if (class_ids) cls[r] = class_ids[i];
}

// ---------------------------------------------------------------------------
// 3. IoU >= thresh bit matrix
// ---------------------------------------------------------------------------
// nms.c:51-61 restated:  w = max(0, min(x2)-max(x1)+1), h likewise, inter = w*h,
// ovr = inter / (area_i + area_j - inter), suppressed iff ovr >= thresh -- every
// operation rounded to fp32 separately (the CPU build has no FMA).
// The divide is avoided without changing the decision: d = fma(-thresh, den, inter)
// is inter - thresh*den rounded once; when den > 0 and |d| > 2^-20 * inter the true
// ratio is more than 2^-20 (relative) away from thresh, far outside the 2^-24 band in
// which the rounded quotient could land on the other side, so sign(d) decides.
// Anything else (near ties, den <= 0, inf/NaN) takes the IEEE divide.
__device__ __forceinline__ bool iou_ge(const float4 a, const float aa, const float4 b, const float ab,
                                       const float thresh)
{
    const float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
    const float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
    const float w = fmaxf(0.f, __fadd_rn(__fsub_rn(xx2, xx1), 1.f));
    const float h = fmaxf(0.f, __fadd_rn(__fsub_rn(yy2, yy1), 1.f));
    const float inter = __fmul_rn(w, h);
    const float den = __fsub_rn(__fadd_rn(aa, ab), inter);
    const float d = __fmaf_rn(-thresh, den, inter);
    if (den > 0.f && fabsf(d) > __fmul_rn(inter, 9.5367431640625e-07f)) return d > 0.f;
    return __fdiv_rn(inter, den) >= thresh;
}

// One 64-thread group per tile (row block rb, column block cb >= rb); 4 groups per CTA.
// Thread t owns row rb*64+t and produces one 64-bit word.  Diagonal tiles carry the
// full symmetric relation (bit t itself cleared).
constexpr int MASK_GROUPS = 4;

template <bool CLS>
__global__ void __launch_bounds__(64 * MASK_GROUPS)
nms_mask_kernel(const float4 *__restrict__ boxes, const float *__restrict__ areas, const int *__restrict__ cls,
                int n, int W, long long n_tiles, float thresh, unsigned long long *__restrict__ mask)
{
    __shared__ float4 s_box[MASK_GROUPS][64];
    __shared__ float s_area[MASK_GROUPS][64];
    __shared__ int s_cls[MASK_GROUPS][64];
    const int grp = threadIdx.x >> 6, t = threadIdx.x & 63;
    const long long tile = (long long)blockIdx.x * MASK_GROUPS + grp;
    const bool active = tile < n_tiles;
    int rb = 0, cb = 0;
    if (active) {
        // tiles are numbered row block by row block: row rb holds W-rb tiles (cb = rb..W-1)
        // start(rb) = rb*W - rb*(rb-1)/2
        const double Wd = (double)W + 0.5;
        rb = (int)(Wd - sqrt(Wd * Wd - 2.0 * (double)tile));
        if (rb < 0) rb = 0;
        if (rb > W - 1) rb = W - 1;
        while (rb > 0 && (long long)rb * W - (long long)rb * (rb - 1) / 2 > tile) --rb;
        while ((long long)(rb + 1) * W - (long long)(rb + 1) * rb / 2 <= tile) ++rb;
        cb = rb + (int)(tile - ((long long)rb * W - (long long)rb * (rb - 1) / 2));
        const int j = cb * 64 + t;
        if (j < n) {
            s_box[grp][t] = boxes[j];
            s_area[grp][t] = areas[j];
            if (CLS) s_cls[grp][t] = cls[j];
        }
    }
    __syncthreads();
    if (!active) return;
    const int i = rb * 64 + t;
    if (i >= n) return;
    const float4 bi = boxes[i];
    const float ai = areas[i];
    const int ci = CLS ? cls[i] : 0;
    const int jn = min(64, n - cb * 64);
    unsigned long long bits = 0ull;
#pragma unroll 4
    for (int k = 0; k < jn; ++k) {
        bool hit = iou_ge(bi, ai, s_box[grp][k], s_area[grp][k], thresh);
        if (CLS) hit = hit && (s_cls[grp][k] == ci);
        bits |= (unsigned long long)hit << k;
    }
    if (rb == cb) bits &= ~(1ull << t);
    mask[(size_t)i * W + cb] = bits;
}

// ---------------------------------------------------------------------------
// 4. greedy scan (single CTA)
// ---------------------------------------------------------------------------
constexpr int SCAN_THREADS = 1024;

__global__ void __launch_bounds__(SCAN_THREADS)
nms_scan_kernel(const unsigned long long *__restrict__ mask, const int *__restrict__ order, int n, int W,
                int max_keep, int64_t *__restrict__ keep64, int *__restrict__ keep32,
                int *__restrict__ num_keep)
{
    extern __shared__ unsigned long long s_removed[];      // [W]
    __shared__ unsigned long long s_kept;
    __shared__ int s_total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int w = tid; w < W; w += SCAN_THREADS) s_removed[w] = 0ull;
    if (tid == 0) { s_total = 0; s_kept = 0ull; }
    __syncthreads();

    int total = 0;
    // diagonal tile rows of the current chunk, prefetched one chunk ahead by warp 0
    unsigned long long d0 = 0ull, d1 = 0ull;
    if (warp == 0) {
        if (lane < n) d0 = mask[(size_t)lane * W];
        if (lane + 32 < n) d1 = mask[(size_t)(lane + 32) * W];
    }
    for (int c = 0; c < W; ++c) {
        if (warp == 0) {
            // issue the next chunk's diagonal loads now; they complete during this step
            unsigned long long n0 = 0ull, n1 = 0ull;
            if (c + 1 < W) {
                const int r0 = (c + 1) * 64 + lane, r1 = r0 + 32;
                if (r0 < n) n0 = __ldg(mask + (size_t)r0 * W + c + 1);
                if (r1 < n) n1 = __ldg(mask + (size_t)r1 * W + c + 1);
            }
            const int nb = min(64, n - c * 64);
            const unsigned long long valid = nb == 64 ? ~0ull : ((1ull << nb) - 1ull);
            unsigned long long cand = ~s_removed[c] & valid;
            unsigned long long kept = 0ull;
            while (cand) {
                const int b = __ffsll((long long)cand) - 1;
                kept |= 1ull << b;
                const unsigned long long ra = __shfl_sync(0xffffffffu, d0, b & 31);
                const unsigned long long rb_ = __shfl_sync(0xffffffffu, d1, b & 31);
                const unsigned long long row = b < 32 ? ra : rb_;
                cand &= ~row;
                cand &= ~(1ull << b);
            }
            int cnt = __popcll(kept);
            if (total + cnt > max_keep) {      // keep only the first (max_keep-total) survivors
                int need = max_keep - total;
                unsigned long long k2 = 0ull, rest = kept;
                while (need-- > 0) {
                    const int b = __ffsll((long long)rest) - 1;
                    k2 |= 1ull << b;
                    rest &= ~(1ull << b);
                }
                kept = k2;
                cnt = __popcll(kept);
            }
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int b = lane + 32 * half;
                if ((kept >> b) & 1ull) {
                    const int slot = total + __popcll(kept & ((1ull << b) - 1ull));
                    const int pos = c * 64 + b;
                    const int idx = order ? order[pos] : pos;
                    if (keep64) keep64[slot] = idx;
                    if (keep32) keep32[slot] = idx;
                }
            }
            if (lane == 0) { s_kept = kept; s_total = total + cnt; }
            d0 = n0;
            d1 = n1;
        }
        __syncthreads();
        const unsigned long long kept = s_kept;
        total = s_total;
        if (total >= max_keep) break;
        if (kept) {
            for (int w = c + 1 + tid; w < W; w += SCAN_THREADS) {
                unsigned long long acc = 0ull, k = kept;
                const unsigned long long *col = mask + (size_t)c * 64 * W + w;
                while (k) {                    // eight independent loads in flight per batch
                    unsigned long long v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        v[u] = 0ull;
                        if (k) {
                            const int b = __ffsll((long long)k) - 1;
                            k &= k - 1ull;
                            v[u] = __ldg(col + (size_t)b * W);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) acc |= v[u];
                }
                s_removed[w] |= acc;
            }
        }
        __syncthreads();
    }
    if (tid == 0) *num_keep = total;
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
size_t nms_buffers_bytes(int n)
{
    const size_t W = (size_t)cdiv(n, 64);
    size_t b = 0;
    b += align_up(sizeof(float4) * (size_t)n, 256);
    b += align_up(sizeof(float) * (size_t)n, 256);
    b += 3 * align_up(sizeof(int) * (size_t)n, 256);
    b += align_up(sizeof(unsigned long long) * (size_t)n * W, 256);
    return b;
}

void nms_carve(void *ws, int n, NmsBuffers &b)
{
    unsigned char *p = static_cast<unsigned char *>(ws);
    b.boxes = reinterpret_cast<float4 *>(p); p += align_up(sizeof(float4) * (size_t)n, 256);
    b.areas = reinterpret_cast<float *>(p);  p += align_up(sizeof(float) * (size_t)n, 256);
    b.cls = reinterpret_cast<int *>(p);      p += align_up(sizeof(int) * (size_t)n, 256);
    b.order = reinterpret_cast<int *>(p);    p += align_up(sizeof(int) * (size_t)n, 256);
    b.rank = reinterpret_cast<int *>(p);     p += align_up(sizeof(int) * (size_t)n, 256);
    b.mask = reinterpret_cast<unsigned long long *>(p);
}

int nms_sorted_launch(const float4 *boxes, const float *areas, const int *cls, const int *order, int n,
                      float thresh, int max_keep, unsigned long long *mask, int64_t *keep64, int *keep32,
                      int *num_keep, cudaStream_t st)
{
    if (max_keep <= 0 || max_keep > n) max_keep = n;
    if (n == 0) {
        SLN_CUDA_OK(cudaMemsetAsync(num_keep, 0, sizeof(int), st));
        return SLN_OK;
    }
    const int W = cdiv(n, 64);
    const long long n_tiles = (long long)W * (W + 1) / 2;
    const long long n_blocks = (n_tiles + MASK_GROUPS - 1) / MASK_GROUPS;
    SLN_REQUIRE(n_blocks < (1ll << 31), SLN_ERR_ARG, "nms: n=%d too large", n);
    if (cls)
        nms_mask_kernel<true><<<(unsigned)n_blocks, 64 * MASK_GROUPS, 0, st>>>(boxes, areas, cls, n, W, n_tiles, thresh, mask);
    else
        nms_mask_kernel<false><<<(unsigned)n_blocks, 64 * MASK_GROUPS, 0, st>>>(boxes, areas, cls, n, W, n_tiles, thresh, mask);
    SLN_LAUNCH_OK("nms_mask_kernel");
    const size_t smem = sizeof(unsigned long long) * (size_t)W;
    SLN_REQUIRE(smem <= 200 * 1024, SLN_ERR_ARG, "nms: n=%d too large for the scan kernel", n);
    if (smem > 48 * 1024)
        SLN_CUDA_OK(cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    nms_scan_kernel<<<1, SCAN_THREADS, smem, st>>>(mask, order, n, W, max_keep, keep64, keep32, num_keep);
    SLN_LAUNCH_OK("nms_scan_kernel");
    return SLN_OK;
}

}  // namespace sln

using namespace sln;

extern "C" size_t sln_nms_workspace_bytes(int n)
{
    if (n < 0) return 0;
    return nms_buffers_bytes(n);
}

extern "C" int sln_nms(const float *dets, const int *class_ids, int n, float thresh, int max_keep,
                       int64_t *keep, int *num_keep, void *workspace, size_t workspace_bytes, void *stream)
{
    SLN_REQUIRE(n >= 0, SLN_ERR_ARG, "negative n");
    SLN_REQUIRE(num_keep != nullptr, SLN_ERR_ARG, "null num_keep");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n == 0) {
        SLN_CUDA_OK(cudaMemsetAsync(num_keep, 0, sizeof(int), st));
        return SLN_OK;
    }
    SLN_REQUIRE(dets && keep, SLN_ERR_ARG, "null pointer");
    SLN_REQUIRE(workspace && workspace_bytes >= nms_buffers_bytes(n), SLN_ERR_WORKSPACE,
                "nms workspace: need %zu bytes, got %zu", nms_buffers_bytes(n), workspace_bytes);
    NmsBuffers b;
    nms_carve(workspace, n, b);
    SLN_CUDA_OK(cudaMemsetAsync(b.rank, 0, sizeof(int) * (size_t)n, st));
    int rc = rank_sort_launch(dets + 4, 5, nullptr, n, b.rank, st);
    if (rc != SLN_OK) return rc;
    nms_gather_kernel<<<cdiv(n, 256), 256, 0, st>>>(dets, class_ids, b.rank, n, b.boxes, b.areas, b.cls, b.order);
    SLN_LAUNCH_OK("nms_gather_kernel");
    return nms_sorted_launch(b.boxes, b.areas, class_ids ? b.cls : nullptr, b.order, n, thresh, max_keep, b.mask,
                             keep, nullptr, num_keep, st);
}
