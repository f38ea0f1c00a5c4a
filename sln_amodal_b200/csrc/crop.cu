// crop.cu -- RoIAlign (TF-style crop_and_resize) forward / backward for sm_100a.
//
// Reference semantics: roialign/roi_align/src/crop_and_resize.c:6-112 (forward),
// :157-252 (backward); GPU baseline being replaced:
// roialign/roi_align/src/cuda/crop_and_resize_kernel.cu:10-82 / :84-165.
//
// Design (DESIGN.md section "RoIAlign"):
//   forward, NHWC   one CTA per ROI; threadIdx.x walks the channel vectors of one
//                   sample point (float4, 512 B per warp-load, 4 loads in flight per
//                   sample, two samples per iteration), threadIdx.y walks sample
//                   points; per-ROI tap tables live in shared memory so the box is
//                   decoded once per CTA, not once per output element.
//   forward, NCHW   one CTA per (ROI, channel chunk); a warp owns one channel plane
//                   at a time and its lanes walk the ph*pw samples, so the writes are
//                   contiguous and the 4 taps come through L1.
//   backward, NHWC  gather form, no atomics, no memset: one CTA per 8x8 pixel tile of
//                   one image; it bins that image's ROIs against the tile (ordered
//                   ballot compaction, so the list keeps the original box order), and
//                   one warp per destination pixel then sums every contribution in
//                   the reference's serial order (box, y, x, tap) in registers and
//                   writes the pixel exactly once.  Result is bit-identical to the
//                   reference's CPU backward and independent of scheduling.
#include "common.cuh"

namespace sln {

// ===========================================================================
// forward, NHWC
// ===========================================================================
struct PyramidMaps {
    const float *map[8];
    int H[8];
    int W[8];
};

// img/out addressed in units of VEC floats.  LEVELS: take the source map from
// `pm` by level[r]; otherwise use pm.map[0].
template <int VEC, bool LEVELS>
__global__ void __launch_bounds__(256)
crop_fwd_nhwc_kernel(PyramidMaps pm, int B, int C, const float *__restrict__ boxes,
                     const int *__restrict__ box_ind, const int *__restrict__ level, int n_levels,
                     int ph, int pw, float ext, float *__restrict__ out)
{
    using V = typename VecT<VEC>::type;
    extern __shared__ Tap s_tab[];          // [ph] y taps, then [pw] x taps
    Tap *ytab = s_tab;
    Tap *xtab = s_tab + ph;

    const int r = blockIdx.x;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int nthr = blockDim.x * blockDim.y;
    const int b = box_ind[r];
    int lv = 0;
    if (LEVELS) lv = level[r];
    const bool ok = (b >= 0 && b < B) && (!LEVELS || (lv >= 0 && lv < n_levels));
    // select the level with static indices (dynamic indexing would spill the
    // by-value parameter struct to local memory)
    int H = pm.H[0], W = pm.W[0];
    const float *mp = pm.map[0];
    if (LEVELS && ok) {
#pragma unroll
        for (int l = 1; l < 8; ++l)
            if (l == lv) { H = pm.H[l]; W = pm.W[l]; mp = pm.map[l]; }
    }
    const V *__restrict__ img = reinterpret_cast<const V *>(mp);
    V *__restrict__ o = reinterpret_cast<V *>(out);
    const int CV = C / VEC;
    const int S = ph * pw;
    const size_t out_base = (size_t)r * S * CV;

    if (!ok) {   // reference GPU kernel skips such boxes (kernel.cu:34-38): rows stay zero
        V z = make_splat(0.f, (V *)nullptr);
        for (int i = tid; i < S * CV; i += nthr) o[out_base + i] = z;
        return;
    }

    const float y1 = boxes[4 * r + 0], x1 = boxes[4 * r + 1];
    const float y2 = boxes[4 * r + 2], x2 = boxes[4 * r + 3];
    if (tid < ph) {
        ytab[tid] = axis_tap(y1, y2, axis_scale(y1, y2, H, ph), H, ph, tid);
    } else if (tid < ph + pw) {
        const int k = tid - ph;
        xtab[k] = axis_tap(x1, x2, axis_scale(x1, x2, W, pw), W, pw, k);
    }
    for (int k = tid + nthr; k < ph + pw; k += nthr) {   // crops larger than the block
        if (k < ph) ytab[k] = axis_tap(y1, y2, axis_scale(y1, y2, H, ph), H, ph, k);
        else xtab[k - ph] = axis_tap(x1, x2, axis_scale(x1, x2, W, pw), W, pw, k - ph);
    }
    __syncthreads();

    const V vext = make_splat(ext, (V *)nullptr);
    const size_t img_base = (size_t)b * H * W * CV;
    const int sstep = blockDim.y;

    for (int cv = threadIdx.x; cv < CV; cv += blockDim.x) {
        int s = threadIdx.y;
        // two sample points per iteration: 8 independent 16-byte loads in flight
        for (; s + sstep < S; s += 2 * sstep) {
            const int sA = s, sB = s + sstep;
            const int yA = sA / pw, xA = sA - yA * pw;
            const int yB = sB / pw, xB = sB - yB * pw;
            const Tap tyA = ytab[yA], txA = xtab[xA];
            const Tap tyB = ytab[yB], txB = xtab[xB];
            const bool vA = (tyA.lo != INVALID_TAP) && (txA.lo != INVALID_TAP);
            const bool vB = (tyB.lo != INVALID_TAP) && (txB.lo != INVALID_TAP);
            V a0, a1, a2, a3, b0, b1, b2, b3;
            if (vA) {
                const int yh = tyA.lo + (tyA.lerp != 0.f), xh = txA.lo + (txA.lerp != 0.f);
                const V *p0 = img + img_base + ((size_t)tyA.lo * W) * CV + cv;
                const V *p1 = img + img_base + ((size_t)yh * W) * CV + cv;
                a0 = ldg_vec(p0 + (size_t)txA.lo * CV);
                a1 = ldg_vec(p0 + (size_t)xh * CV);
                a2 = ldg_vec(p1 + (size_t)txA.lo * CV);
                a3 = ldg_vec(p1 + (size_t)xh * CV);
            }
            if (vB) {
                const int yh = tyB.lo + (tyB.lerp != 0.f), xh = txB.lo + (txB.lerp != 0.f);
                const V *p0 = img + img_base + ((size_t)tyB.lo * W) * CV + cv;
                const V *p1 = img + img_base + ((size_t)yh * W) * CV + cv;
                b0 = ldg_vec(p0 + (size_t)txB.lo * CV);
                b1 = ldg_vec(p0 + (size_t)xh * CV);
                b2 = ldg_vec(p1 + (size_t)txB.lo * CV);
                b3 = ldg_vec(p1 + (size_t)xh * CV);
            }
            const V ra = vA ? lerp2v(a0, a1, a2, a3, txA.lerp, tyA.lerp) : vext;
            const V rb = vB ? lerp2v(b0, b1, b2, b3, txB.lerp, tyB.lerp) : vext;
            __stcs(o + out_base + (size_t)sA * CV + cv, ra);
            __stcs(o + out_base + (size_t)sB * CV + cv, rb);
        }
        if (s < S) {
            const int y = s / pw, x = s - y * pw;
            const Tap ty = ytab[y], tx = xtab[x];
            V res = vext;
            if (ty.lo != INVALID_TAP && tx.lo != INVALID_TAP) {
                const int yh = ty.lo + (ty.lerp != 0.f), xh = tx.lo + (tx.lerp != 0.f);
                const V *p0 = img + img_base + ((size_t)ty.lo * W) * CV + cv;
                const V *p1 = img + img_base + ((size_t)yh * W) * CV + cv;
                const V a0 = ldg_vec(p0 + (size_t)tx.lo * CV);
                const V a1 = ldg_vec(p0 + (size_t)xh * CV);
                const V a2 = ldg_vec(p1 + (size_t)tx.lo * CV);
                const V a3 = ldg_vec(p1 + (size_t)xh * CV);
                res = lerp2v(a0, a1, a2, a3, tx.lerp, ty.lerp);
            }
            __stcs(o + out_base + (size_t)s * CV + cv, res);
        }
    }
}

// ===========================================================================
// forward, NCHW (the reference's native layout; API default for NCHW tensors)
// ===========================================================================
__global__ void __launch_bounds__(256)
crop_fwd_nchw_kernel(const float *__restrict__ img, int B, int C, int H, int W,
                     const float *__restrict__ boxes, const int *__restrict__ box_ind,
                     int ph, int pw, float ext, int c_chunk, float *__restrict__ out)
{
    extern __shared__ Tap s_tab[];
    Tap *ytab = s_tab;
    Tap *xtab = s_tab + ph;
    const int r = blockIdx.x;
    const int c0 = blockIdx.y * c_chunk;
    const int c1 = min(C, c0 + c_chunk);
    const int tid = threadIdx.x;
    const int S = ph * pw;
    const int b = box_ind[r];
    float *__restrict__ o = out + ((size_t)r * C) * S;
    if (b < 0 || b >= B) {
        for (int i = c0 * S + tid; i < c1 * S; i += blockDim.x) o[i] = 0.f;
        return;
    }
    const float y1 = boxes[4 * r + 0], x1 = boxes[4 * r + 1];
    const float y2 = boxes[4 * r + 2], x2 = boxes[4 * r + 3];
    for (int k = tid; k < ph + pw; k += blockDim.x) {
        if (k < ph) ytab[k] = axis_tap(y1, y2, axis_scale(y1, y2, H, ph), H, ph, k);
        else xtab[k - ph] = axis_tap(x1, x2, axis_scale(x1, x2, W, pw), W, pw, k - ph);
    }
    __syncthreads();
    const int lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    const size_t plane = (size_t)H * W;
    for (int s = lane; s < S; s += 32) {
        const int y = s / pw, x = s - y * pw;
        const Tap ty = ytab[y], tx = xtab[x];
        const bool valid = ty.lo != INVALID_TAP && tx.lo != INVALID_TAP;
        const int yh = ty.lo + (ty.lerp != 0.f), xh = tx.lo + (tx.lerp != 0.f);
        const int o00 = ty.lo * W + tx.lo, o01 = ty.lo * W + xh;
        const int o10 = yh * W + tx.lo, o11 = yh * W + xh;
        for (int c = c0 + warp; c < c1; c += nwarp) {
            float v = ext;
            if (valid) {
                const float *p = img + ((size_t)b * C + c) * plane;
                v = lerp2(__ldg(p + o00), __ldg(p + o01), __ldg(p + o10), __ldg(p + o11), tx.lerp, ty.lerp);
            }
            __stcs(o + (size_t)c * S + s, v);
        }
    }
}

// ===========================================================================
// backward, NHWC
// ===========================================================================
// Pixel window [y0,y1]x[x0,x1] (inclusive, clamped) that a ROI's valid samples can
// touch, or an empty window.  Sample positions are monotone in k, so the first and
// last sample bound the window; a conservative superset is fine (exact per-sample
// tests happen in the accumulate phase).
struct RoiWin {
    short y0, y1, x0, x1;
};

__device__ __forceinline__ void axis_window(float a1, float a2, int extent, int crop, int &w0, int &w1)
{
    const float em1 = (float)(extent - 1);
    const float scale = axis_scale(a1, a2, extent, crop);
    float p0, p1;
    if (crop > 1) {
        p0 = __fmul_rn(a1, em1);
        p1 = __fadd_rn(p0, __fmul_rn((float)(crop - 1), scale));
    } else {
        p0 = p1 = (float)(0.5 * (double)__fadd_rn(a1, a2) * (double)(extent - 1));
    }
    if (!(p0 == p0) || !(p1 == p1)) { w0 = 1; w1 = 0; return; }   // NaN: nothing
    const float lo = fminf(p0, p1), hi = fmaxf(p0, p1);
    if (hi < 0.f || lo > em1) { w0 = 1; w1 = 0; return; }
    w0 = (int)floorf(fmaxf(lo, 0.f));
    w1 = (int)ceilf(fminf(hi, em1));
}

// prep 1: per-ROI windows
__global__ void crop_bwd_windows_kernel(const float *__restrict__ boxes, const int *__restrict__ box_ind,
                                        const int *__restrict__ level, int which_level, int N, int B,
                                        int H, int W, int ph, int pw, RoiWin *__restrict__ win,
                                        int *__restrict__ counts)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= N) return;
    RoiWin w;
    w.y0 = 1; w.y1 = 0; w.x0 = 1; w.x1 = 0;
    const int b = box_ind[r];
    const bool use = (b >= 0 && b < B) && (level == nullptr || level[r] == which_level);
    if (use) {
        int a0, a1, c0, c1;
        axis_window(boxes[4 * r + 0], boxes[4 * r + 2], H, ph, a0, a1);
        axis_window(boxes[4 * r + 1], boxes[4 * r + 3], W, pw, c0, c1);
        if (a0 <= a1 && c0 <= c1) {
            w.y0 = (short)a0; w.y1 = (short)a1; w.x0 = (short)c0; w.x1 = (short)c1;
            atomicAdd(counts + b, 1);     // integer count: order-independent
        }
    }
    win[r] = w;
}

// prep 2: per-image ROI lists in original box order (stable partition by box_ind).
// One CTA per image; ordered ballot compaction into one compact array: image b's
// list starts at sum(counts[0..b-1]) (counts come from the windows kernel).
__global__ void __launch_bounds__(256)
crop_bwd_lists_kernel(const int *__restrict__ box_ind, const RoiWin *__restrict__ win, int N,
                      const int *__restrict__ counts, int *__restrict__ lists)
{
    __shared__ int s_warp[8];
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int base = 0;
    for (int i = 0; i < b; ++i) base += counts[i];
    for (int start = 0; start < N; start += 256) {
        const int r = start + tid;
        bool take = false;
        if (r < N) {
            const RoiWin w = win[r];
            take = (box_ind[r] == b) && (w.y0 <= w.y1);
        }
        const unsigned m = __ballot_sync(0xffffffffu, take);
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        int off = 0, total = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const int c = s_warp[w];
            if (w < warp) off += c;
            total += c;
        }
        if (take) lists[base + off + __popc(m & ((1u << lane) - 1u))] = r;
        base += total;
        __syncthreads();
    }
}

template <bool EXACT>
__device__ __forceinline__ float accum1(float s, float g, float wy, float wx)
{
    if (EXACT) return __fadd_rn(s, __fmul_rn(wx, __fmul_rn(wy, g)));   // crop_and_resize.c:241-247
    return __fmaf_rn(__fmul_rn(wy, wx), g, s);
}
template <bool EXACT>
__device__ __forceinline__ float accum(float s, float g, float wy, float wx) { return accum1<EXACT>(s, g, wy, wx); }
template <bool EXACT>
__device__ __forceinline__ float4 accum(float4 s, float4 g, float wy, float wx)
{
    return make_float4(accum1<EXACT>(s.x, g.x, wy, wx), accum1<EXACT>(s.y, g.y, wy, wx),
                       accum1<EXACT>(s.z, g.z, wy, wx), accum1<EXACT>(s.w, g.w, wy, wx));
}

constexpr int BWD_TILE = 8;            // 8x8 destination pixels per CTA
constexpr int BWD_PIX = BWD_TILE * BWD_TILE;
constexpr int BWD_THREADS = 256;       // 8 warps: warp w owns tile row w
constexpr int BWD_MAX_ROUND = 64;      // list entries examined per ROI round (<= BWD_THREADS)
constexpr int BWD_CAP = 32;            // visit entries per destination pixel per visit round

// Shared memory per ROI round of `CH` examined list entries (dynamic):
//   Tap ytab[CH][ph], xtab[CH][pw]     tap tables of the accepted ROIs
//   int roi[CH]                        accepted ROI ids, original order
//   u16 yr[CH][8], xr[CH][8]           per tile row / column: sample range lo | hi<<8
//   visit lists, entry-major so the 64 pixel threads write conflict-free:
//   u32 v_sid[CAP][64]; float v_wy[CAP][64], v_wx[CAP][64]; int v_cnt[64]
static size_t bwd_smem_bytes(int CH, int ph, int pw)
{
    return (size_t)CH * ((size_t)(ph + pw) * sizeof(Tap) + sizeof(int) + 2 * BWD_TILE * sizeof(unsigned short)) +
           (size_t)BWD_CAP * BWD_PIX * 12 + BWD_PIX * sizeof(int);
}

// Backward kernel.  Three kinds of work, all deterministic:
//   A  (CTA)          examine the next CH entries of this image's ROI list, keep (in order) those
//                     whose pixel window meets the tile
//   B  (CTA)          tap tables of the kept ROIs; per tile row/column the contiguous range of
//                     samples that touch it
//   B' (64 threads)   one thread per destination pixel walks the kept ROIs in order and emits
//                     its "visits" (sample id, wy, wx) in the reference's order (ROI, y, x, tap)
//   C  (8 warps)      warp = tile row; for each of its 8 pixels it streams the pixel's visit list:
//                     loads issued four visits ahead, lanes = channel vectors, accumulation in
//                     registers; every pixel is written exactly once at the end.
// EXACT: sum += wx*(wy*g) with every operation rounded like crop_and_resize.c:241-247 (bit-
// identical to the reference CPU backward); otherwise sum = fma(wy*wx, g, sum) (<= 1 ulp per
// term away, same order, still deterministic).
template <int VEC, bool EXACT>
__global__ void __launch_bounds__(BWD_THREADS)
crop_bwd_nhwc_kernel(const float *__restrict__ grads, const float *__restrict__ boxes,
                     const RoiWin *__restrict__ win, const int *__restrict__ lists,
                     const int *__restrict__ counts, int C, int ph, int pw,
                     float *__restrict__ grad_image, int B, int H, int W, int tiles_x, int tiles_y, int CH)
{
    using V = typename VecT<VEC>::type;
    extern __shared__ __align__(16) unsigned char s_raw[];
    Tap *ytab = reinterpret_cast<Tap *>(s_raw);                          // [CH][ph]
    Tap *xtab = ytab + (size_t)CH * ph;                                  // [CH][pw]
    unsigned *v_sid = reinterpret_cast<unsigned *>(xtab + (size_t)CH * pw);   // [CAP][64]
    float *v_wy = reinterpret_cast<float *>(v_sid + BWD_CAP * BWD_PIX);  // [CAP][64]
    float *v_wx = v_wy + BWD_CAP * BWD_PIX;                              // [CAP][64]
    int *v_cnt = reinterpret_cast<int *>(v_wx + BWD_CAP * BWD_PIX);      // [64]
    int *s_roi = v_cnt + BWD_PIX;                                        // [CH]
    unsigned short *s_yr = reinterpret_cast<unsigned short *>(s_roi + CH);   // [CH][8]
    unsigned short *s_xr = s_yr + (size_t)CH * BWD_TILE;                 // [CH][8]
    __shared__ int s_warp[8];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int t = blockIdx.x;
    const int tx_i = t % tiles_x; t /= tiles_x;
    const int ty_i = t % tiles_y; t /= tiles_y;
    const int b = t;
    const int ty0 = ty_i * BWD_TILE, tx0 = tx_i * BWD_TILE;
    const int ty1 = min(ty0 + BWD_TILE, H) - 1, tx1 = min(tx0 + BWD_TILE, W) - 1;
    const int CV = C / VEC;
    const int cv = blockIdx.y * 32 + lane;       // this lane's channel vector
    const bool cv_ok = cv < CV;

    V acc[BWD_TILE];
#pragma unroll
    for (int p = 0; p < BWD_TILE; ++p) acc[p] = make_splat(0.f, (V *)nullptr);

    // this image's ROI list (original box order) inside the compact list array
    int list_off = 0;
    for (int i = 0; i < b; ++i) list_off += counts[i];
    const int n_list = counts[b];
    const int *__restrict__ list = lists + list_off;
    const V *__restrict__ g = reinterpret_cast<const V *>(grads) + cv;
    const int S = ph * pw;
    // pixel-thread role (threads 0..63): destination pixel (prow, pcol) of the tile
    const int prow = tid >> 3, pcol = tid & 7;
    const int ppy = ty0 + prow, ppx = tx0 + pcol;

    for (int scan = 0; scan < n_list; scan += CH) {
        // ---- A: examine list[scan, scan+CH)
        const int li = scan + tid;
        bool take = false;
        int r = -1;
        if (tid < CH && li < n_list) {
            r = list[li];
            const RoiWin w = win[r];
            take = !(w.y1 < ty0 || w.y0 > ty1 || w.x1 < tx0 || w.x0 > tx1);
        }
        const unsigned m = __ballot_sync(0xffffffffu, take);
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        int off = 0, n_chunk = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const int c = s_warp[w];
            if (w < warp) off += c;
            n_chunk += c;
        }
        if (take) s_roi[off + __popc(m & ((1u << lane) - 1u))] = r;
        __syncthreads();
        if (n_chunk == 0) continue;     // uniform: every thread read the same s_warp values

        // ---- B: tap tables, then per-row / per-column sample ranges of the kept ROIs
        for (int i = tid; i < n_chunk * (ph + pw); i += BWD_THREADS) {
            const int q = i / (ph + pw), k = i - q * (ph + pw);
            const int rr = s_roi[q];
            if (k < ph) {
                const float a1 = boxes[4 * rr + 0], a2 = boxes[4 * rr + 2];
                ytab[q * ph + k] = axis_tap(a1, a2, axis_scale(a1, a2, H, ph), H, ph, k);
            } else {
                const float a1 = boxes[4 * rr + 1], a2 = boxes[4 * rr + 3];
                xtab[q * pw + (k - ph)] = axis_tap(a1, a2, axis_scale(a1, a2, W, pw), W, pw, k - ph);
            }
        }
        __syncthreads();
        for (int i = tid; i < n_chunk * 2 * BWD_TILE; i += BWD_THREADS) {
            const int q = i / (2 * BWD_TILE), j = i - q * (2 * BWD_TILE);
            const bool isx = j >= BWD_TILE;
            const int jj = isx ? j - BWD_TILE : j;
            const int pix = (isx ? tx0 : ty0) + jj;
            const Tap *tab = isx ? (xtab + q * pw) : (ytab + q * ph);
            const int cnt = isx ? pw : ph;
            int lo = 255, hi = 0;
            bool any = false;
            for (int k = 0; k < cnt; ++k) {
                const Tap tp = tab[k];
                if (tp.lo == INVALID_TAP) continue;
                const int h = tp.lo + (tp.lerp != 0.f);
                if (tp.lo == pix || h == pix) {
                    if (!any) lo = k;
                    hi = k;
                    any = true;
                }
            }
            const unsigned short packed = any ? (unsigned short)(lo | (hi << 8)) : (unsigned short)0x00ff;
            (isx ? s_xr : s_yr)[q * BWD_TILE + jj] = packed;
        }
        __syncthreads();

        // ---- B' / C: visit rounds.  Cursor of the pixel thread: (q, y, x, tap)
        int cq = 0, cy = -1, cx = -1, ctap = 0;
        for (;;) {
            int more = 0;
            if (tid < BWD_PIX) {
                int cnt = 0;
                while (cq < n_chunk) {
                    const unsigned yrng = s_yr[cq * BWD_TILE + prow], xrng = s_xr[cq * BWD_TILE + pcol];
                    const int ylo = yrng & 0xff, yhi = yrng >> 8, xlo = xrng & 0xff, xhi = xrng >> 8;
                    if (ylo > yhi || xlo > xhi) { ++cq; cy = -1; continue; }
                    if (cy < 0) { cy = ylo; cx = xlo; ctap = 0; }
                    const unsigned sid0 = (unsigned)s_roi[cq] * (unsigned)S;
                    bool full = false;
                    for (; cy <= yhi && !full; ++cy) {
                        const Tap tyy = ytab[cq * ph + cy];
                        const bool top = (tyy.lo == ppy), bot = (tyy.lo + (tyy.lerp != 0.f) == ppy);
                        for (; cx <= xhi && !full; ++cx) {
                            const Tap txx = xtab[cq * pw + cx];
                            const bool lft = (txx.lo == ppx), rgt = (txx.lo + (txx.lerp != 0.f) == ppx);
                            for (; ctap < 4; ++ctap) {         // reference tap order TL, TR, BL, BR
                                const bool useb = ctap >> 1, user = ctap & 1;
                                if (!((useb ? bot : top) && (user ? rgt : lft))) continue;
                                if (cnt == BWD_CAP) { full = true; break; }
                                v_sid[cnt * BWD_PIX + tid] = sid0 + (unsigned)(cy * pw + cx);
                                v_wy[cnt * BWD_PIX + tid] = useb ? tyy.lerp : __fsub_rn(1.f, tyy.lerp);
                                v_wx[cnt * BWD_PIX + tid] = user ? txx.lerp : __fsub_rn(1.f, txx.lerp);
                                ++cnt;
                            }
                            if (full) break;
                            ctap = 0;
                        }
                        if (full) break;
                        cx = xlo;
                    }
                    if (full) { more = 1; break; }
                    ++cq; cy = -1;
                }
                v_cnt[tid] = cnt;
            }
            more = __syncthreads_or(more);

            // ---- C: stream the visit lists.  warp = tile row, lanes = channel vectors
            if (cv_ok) {
#pragma unroll
                for (int p = 0; p < BWD_TILE; ++p) {
                    const int pix = warp * BWD_TILE + p;
                    const int cnt = v_cnt[pix];
                    V a = acc[p];
                    int i = 0;
                    for (; i + 4 <= cnt; i += 4) {
                        V gv[4];
                        float wy[4], wx[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const unsigned sid = v_sid[(i + u) * BWD_PIX + pix];
                            wy[u] = v_wy[(i + u) * BWD_PIX + pix];
                            wx[u] = v_wx[(i + u) * BWD_PIX + pix];
                            gv[u] = ldg_vec(g + (size_t)sid * CV);
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) a = accum<EXACT>(a, gv[u], wy[u], wx[u]);
                    }
                    for (; i < cnt; ++i) {
                        const unsigned sid = v_sid[i * BWD_PIX + pix];
                        const V gv = ldg_vec(g + (size_t)sid * CV);
                        a = accum<EXACT>(a, gv, v_wy[i * BWD_PIX + pix], v_wx[i * BWD_PIX + pix]);
                    }
                    acc[p] = a;
                }
            }
            if (!more) break;
            __syncthreads();        // lists are rewritten by the next visit round
        }
        __syncthreads();            // tables / lists are rewritten by the next ROI round
    }

    // ---- write every pixel of the tile exactly once (zeros included)
    const int py = ty0 + warp;
    if (py <= ty1 && cv_ok) {
        V *__restrict__ o = reinterpret_cast<V *>(grad_image);
#pragma unroll
        for (int p = 0; p < BWD_TILE; ++p) {
            const int px = tx0 + p;
            if (px <= tx1) __stcs(o + (((size_t)b * H + py) * W + px) * CV + cv, acc[p]);
        }
    }
}

// ===========================================================================
// layout converters: per image, [C][HW] <-> [HW][C]
// ===========================================================================
__global__ void __launch_bounds__(256)
transpose_kernel(const float *__restrict__ src, float *__restrict__ dst, int rows, int cols)
{
    // src [batch][rows][cols] -> dst [batch][cols][rows]
    __shared__ float tile[32][33];
    const size_t boff = (size_t)blockIdx.z * rows * cols;
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int j = ty; j < 32; j += 8) {
        const int r = r0 + j, c = c0 + tx;
        if (r < rows && c < cols) tile[j][tx] = __ldg(src + boff + (size_t)r * cols + c);
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int c = c0 + j, r = r0 + tx;
        if (r < rows && c < cols) dst[boff + (size_t)c * rows + r] = tile[tx][j];
    }
}

static int launch_transpose(const float *src, float *dst, int batch, int rows, int cols, cudaStream_t st)
{
    if (batch == 0 || rows == 0 || cols == 0) return SLN_OK;
    SLN_REQUIRE(cdiv(rows, 32) <= 65535 && batch <= 65535, SLN_ERR_ARG,
                "transpose: rows/batch too large (%d, %d)", rows, batch);
    dim3 grid(cdiv(cols, 32), cdiv(rows, 32), batch);
    transpose_kernel<<<grid, 256, 0, st>>>(src, dst, rows, cols);
    SLN_LAUNCH_OK("transpose_kernel");
    return SLN_OK;
}

// ===========================================================================
// host-side launch logic
// ===========================================================================
static bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// max_crop: 255 for the backward (sample ranges are packed in bytes), 4096 for the forward
static int check_crop_args(const void *image, const void *boxes, const void *box_ind, const void *out,
                           int B, int C, int H, int W, int N, int ph, int pw, int max_crop = 255)
{
    SLN_REQUIRE(B >= 0 && C >= 0 && H >= 0 && W >= 0 && N >= 0, SLN_ERR_ARG, "negative size");
    SLN_REQUIRE(ph >= 1 && pw >= 1 && ph <= max_crop && pw <= max_crop, SLN_ERR_ARG,
                "crop size %dx%d outside [1,%d]", ph, pw, max_crop);
    SLN_REQUIRE(H <= 32767 && W <= 32767, SLN_ERR_ARG, "map side > 32767");
    if (N > 0 && C > 0) {
        SLN_REQUIRE(boxes && box_ind && out, SLN_ERR_ARG, "null pointer");
        SLN_REQUIRE(B == 0 || H == 0 || W == 0 || image, SLN_ERR_ARG, "null image");
    }
    return SLN_OK;
}

static void fwd_block_shape(int CV, dim3 &block)
{
    int bx = CV < 256 ? CV : 256;
    if (bx >= 32) bx = (bx / 32) * 32;          // whole warps along the channel axis
    int by = 256 / bx;
    if (by < 1) by = 1;
    block = dim3(bx, by, 1);
}

static int crop_fwd_nhwc(const PyramidMaps &pm, int n_levels, bool levels, int B, int C,
                         const float *boxes, const int *box_ind, const int *level, int N, int ph, int pw,
                         float ext, float *out, cudaStream_t st)
{
    bool vec4 = (C % 4 == 0) && aligned16(out);
    for (int l = 0; l < n_levels; ++l) vec4 = vec4 && aligned16(pm.map[l]);
    const int CV = vec4 ? C / 4 : C;
    dim3 block;
    fwd_block_shape(CV, block);
    const size_t smem = sizeof(Tap) * (size_t)(ph + pw);
    dim3 grid(N);
    if (vec4) {
        if (levels) crop_fwd_nhwc_kernel<4, true><<<grid, block, smem, st>>>(pm, B, C, boxes, box_ind, level, n_levels, ph, pw, ext, out);
        else crop_fwd_nhwc_kernel<4, false><<<grid, block, smem, st>>>(pm, B, C, boxes, box_ind, level, n_levels, ph, pw, ext, out);
    } else {
        if (levels) crop_fwd_nhwc_kernel<1, true><<<grid, block, smem, st>>>(pm, B, C, boxes, box_ind, level, n_levels, ph, pw, ext, out);
        else crop_fwd_nhwc_kernel<1, false><<<grid, block, smem, st>>>(pm, B, C, boxes, box_ind, level, n_levels, ph, pw, ext, out);
    }
    SLN_LAUNCH_OK("crop_fwd_nhwc_kernel");
    return SLN_OK;
}

template <int VEC, bool EXACT>
static int launch_bwd(const float *grads, const float *boxes, const RoiWin *win, const int *lists,
                      const int *counts, int C, int ph, int pw, float *grad_image, int B, int H,
                      int W, cudaStream_t st)
{
    const int tiles_x = cdiv(W, BWD_TILE), tiles_y = cdiv(H, BWD_TILE);
    const int chunks = cdiv(C / VEC, 32);
    // list entries examined per ROI round: keep the CTA under ~56 KB of shared memory
    int CH = BWD_MAX_ROUND;
    while (CH > 1 && bwd_smem_bytes(CH, ph, pw) > 56 * 1024) CH /= 2;
    const size_t smem = bwd_smem_bytes(CH, ph, pw);
    SLN_REQUIRE(smem <= 200 * 1024, SLN_ERR_ARG, "crop %dx%d too large for the backward kernel", ph, pw);
    auto kern = crop_bwd_nhwc_kernel<VEC, EXACT>;
    if (smem > 48 * 1024)
        SLN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SLN_REQUIRE(chunks <= 65535, SLN_ERR_ARG, "too many channel chunks");
    dim3 grid((unsigned)((size_t)tiles_x * tiles_y * B), chunks);
    kern<<<grid, BWD_THREADS, smem, st>>>(grads, boxes, win, lists, counts, C, ph, pw, grad_image, B, H, W,
                                           tiles_x, tiles_y, CH);
    SLN_LAUNCH_OK("crop_bwd_nhwc_kernel");
    return SLN_OK;
}

static size_t bwd_ws_bytes(int N, int B)
{
    return align_up(sizeof(RoiWin) * (size_t)N, 256) + align_up(sizeof(int) * (size_t)N, 256) +
           align_up(sizeof(int) * (size_t)(B + 1), 256);
}

static int crop_bwd_nhwc(const float *grads, const float *boxes, const int *box_ind, const int *level,
                         int which_level, int N, int C, int ph, int pw, float *grad_image, int B, int H,
                         int W, bool exact, void *ws, size_t ws_bytes, cudaStream_t st)
{
    if (B == 0 || C == 0 || H == 0 || W == 0) return SLN_OK;
    SLN_REQUIRE((size_t)cdiv(W, BWD_TILE) * cdiv(H, BWD_TILE) * B < (1ull << 31), SLN_ERR_ARG, "too many tiles");
    SLN_REQUIRE(ws_bytes >= bwd_ws_bytes(N, B), SLN_ERR_WORKSPACE, "crop bwd workspace: need %zu bytes, got %zu",
                bwd_ws_bytes(N, B), ws_bytes);
    SLN_REQUIRE(ws != nullptr, SLN_ERR_WORKSPACE, "null workspace");
    unsigned char *p = static_cast<unsigned char *>(ws);
    RoiWin *win = reinterpret_cast<RoiWin *>(p);
    p += align_up(sizeof(RoiWin) * (size_t)N, 256);
    int *lists = reinterpret_cast<int *>(p);
    p += align_up(sizeof(int) * (size_t)N, 256);
    int *counts = reinterpret_cast<int *>(p);

    SLN_CUDA_OK(cudaMemsetAsync(counts, 0, sizeof(int) * (size_t)(B + 1), st));
    if (N > 0) {
        crop_bwd_windows_kernel<<<cdiv(N, 256), 256, 0, st>>>(boxes, box_ind, level, which_level, N, B, H, W,
                                                             ph, pw, win, counts);
        SLN_LAUNCH_OK("crop_bwd_windows_kernel");
        crop_bwd_lists_kernel<<<B, 256, 0, st>>>(box_ind, win, N, counts, lists);
        SLN_LAUNCH_OK("crop_bwd_lists_kernel");
    }
    const bool vec4 = (C % 4 == 0) && aligned16(grads) && aligned16(grad_image);
    if (vec4) {
        if (exact) return launch_bwd<4, true>(grads, boxes, win, lists, counts, C, ph, pw, grad_image, B, H, W, st);
        return launch_bwd<4, false>(grads, boxes, win, lists, counts, C, ph, pw, grad_image, B, H, W, st);
    }
    if (exact) return launch_bwd<1, true>(grads, boxes, win, lists, counts, C, ph, pw, grad_image, B, H, W, st);
    return launch_bwd<1, false>(grads, boxes, win, lists, counts, C, ph, pw, grad_image, B, H, W, st);
}

}  // namespace sln

// ===========================================================================
// C ABI
// ===========================================================================
using namespace sln;

extern "C" int sln_crop_and_resize_fwd(const float *image, int B, int C, int H, int W, int layout,
                                       const float *boxes, const int *box_ind, int N, int ph, int pw,
                                       float ext, float *crops, void *stream)
{
    int rc = check_crop_args(image, boxes, box_ind, crops, B, C, H, W, N, ph, pw, 4096);
    if (rc != SLN_OK) return rc;
    if (N == 0 || C == 0) return SLN_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (layout == SLN_LAYOUT_NHWC) {
        PyramidMaps pm{};
        pm.map[0] = image; pm.H[0] = H; pm.W[0] = W;
        return crop_fwd_nhwc(pm, 1, false, B, C, boxes, box_ind, nullptr, N, ph, pw, ext, crops, st);
    }
    SLN_REQUIRE(layout == SLN_LAYOUT_NCHW, SLN_ERR_LAYOUT, "unknown layout %d", layout);
    // channel chunks so that small-N calls still fill the machine
    int c_chunk = C;
    const int want_ctas = 4 * sm_count();
    if (N < want_ctas) {
        const int split = cdiv(want_ctas, N);
        c_chunk = cdiv(C, split);
        if (c_chunk < 8) c_chunk = C < 8 ? C : 8;
    }
    dim3 grid(N, cdiv(C, c_chunk));
    SLN_REQUIRE(grid.y <= 65535, SLN_ERR_ARG, "too many channel chunks");
    const size_t smem = sizeof(Tap) * (size_t)(ph + pw);
    crop_fwd_nchw_kernel<<<grid, 256, smem, st>>>(image, B, C, H, W, boxes, box_ind, ph, pw, ext, c_chunk, crops);
    SLN_LAUNCH_OK("crop_fwd_nchw_kernel");
    return SLN_OK;
}

extern "C" size_t sln_crop_and_resize_bwd_workspace_bytes(int N, int B)
{
    if (N < 0 || B < 0) return 0;
    return bwd_ws_bytes(N, B);
}

extern "C" int sln_crop_and_resize_bwd(const float *grads, const float *boxes, const int *box_ind, int N,
                                       int C, int ph, int pw, float *grad_image, int B, int H, int W,
                                       int layout, int flags, void *workspace, size_t workspace_bytes, void *stream)
{
    SLN_REQUIRE(layout == SLN_LAYOUT_NHWC, SLN_ERR_LAYOUT,
                "crop backward is NHWC-only; convert with sln_nchw_to_nhwc / sln_nhwc_to_nchw");
    int rc = check_crop_args(grad_image, boxes, box_ind, grad_image, B, C, H, W, N, ph, pw);
    if (rc != SLN_OK) return rc;
    SLN_REQUIRE(N == 0 || C == 0 || grads, SLN_ERR_ARG, "null grads");
    return crop_bwd_nhwc(grads, boxes, box_ind, nullptr, 0, N, C, ph, pw, grad_image, B, H, W,
                         (flags & SLN_BWD_EXACT) != 0, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

extern "C" int sln_pyramid_crop_fwd(const float *const *maps_host, const int *H_host, const int *W_host,
                                    int n_levels, int B, int C, const float *boxes, const int *box_ind,
                                    const int *level, int N, int ph, int pw, float ext, float *crops,
                                    void *stream)
{
    SLN_REQUIRE(n_levels >= 1 && n_levels <= 8, SLN_ERR_ARG, "n_levels %d outside [1,8]", n_levels);
    SLN_REQUIRE(maps_host && H_host && W_host && level, SLN_ERR_ARG, "null pointer");
    PyramidMaps pm{};
    for (int l = 0; l < n_levels; ++l) {
        int rc = check_crop_args(maps_host[l], boxes, box_ind, crops, B, C, H_host[l], W_host[l], N, ph, pw, 4096);
        if (rc != SLN_OK) return rc;
        pm.map[l] = maps_host[l]; pm.H[l] = H_host[l]; pm.W[l] = W_host[l];
    }
    if (N == 0 || C == 0) return SLN_OK;
    return crop_fwd_nhwc(pm, n_levels, true, B, C, boxes, box_ind, level, N, ph, pw, ext, crops,
                         static_cast<cudaStream_t>(stream));
}

extern "C" int sln_pyramid_crop_bwd_level(const float *grads, const float *boxes, const int *box_ind,
                                          const int *level, int which_level, int N, int C, int ph, int pw,
                                          float *grad_image, int B, int H, int W, int flags, void *workspace,
                                          size_t workspace_bytes, void *stream)
{
    int rc = check_crop_args(grad_image, boxes, box_ind, grad_image, B, C, H, W, N, ph, pw);
    if (rc != SLN_OK) return rc;
    SLN_REQUIRE(N == 0 || C == 0 || grads, SLN_ERR_ARG, "null grads");
    return crop_bwd_nhwc(grads, boxes, box_ind, level, which_level, N, C, ph, pw, grad_image, B, H, W,
                         (flags & SLN_BWD_EXACT) != 0, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

extern "C" int sln_nchw_to_nhwc(const float *src, float *dst, int B, int C, int H, int W, void *stream)
{
    SLN_REQUIRE(B >= 0 && C >= 0 && H >= 0 && W >= 0, SLN_ERR_ARG, "negative size");
    SLN_REQUIRE((size_t)H * W < (1ull << 31), SLN_ERR_ARG, "plane too large");
    if ((size_t)B * C * H * W == 0) return SLN_OK;
    SLN_REQUIRE(src && dst && src != dst, SLN_ERR_ARG, "null or aliased pointers");
    return launch_transpose(src, dst, B, C, H * W, static_cast<cudaStream_t>(stream));
}

extern "C" int sln_nhwc_to_nchw(const float *src, float *dst, int B, int C, int H, int W, void *stream)
{
    SLN_REQUIRE(B >= 0 && C >= 0 && H >= 0 && W >= 0, SLN_ERR_ARG, "negative size");
    SLN_REQUIRE((size_t)H * W < (1ull << 31), SLN_ERR_ARG, "plane too large");
    if ((size_t)B * C * H * W == 0) return SLN_OK;
    SLN_REQUIRE(src && dst && src != dst, SLN_ERR_ARG, "null or aliased pointers");
    return launch_transpose(src, dst, B, H * W, C, static_cast<cudaStream_t>(stream));
}
