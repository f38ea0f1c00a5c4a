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

// Supertiles: the destination map is cut into at most 8x8 supertiles per image (side a
// multiple of the 8-pixel tile); the prep kernels build, per supertile, the ordered list
// of ROIs whose window meets it, so a tile CTA scans tens of entries instead of the
// whole image's ROI list.
struct SuperGrid {
    int side;      // pixels per supertile side
    int nx, ny;    // supertiles per image
};

static SuperGrid super_grid(int H, int W)
{
    SuperGrid g;
    int side = 32;
    while (cdiv(H, side) > 8 || cdiv(W, side) > 8) side *= 2;
    g.side = side;
    g.nx = cdiv(W, side);
    g.ny = cdiv(H, side);
    return g;
}

// prep 1: per-ROI windows + per-supertile counts (integer atomics: order-independent)
__global__ void crop_bwd_windows_kernel(const float *__restrict__ boxes, const int *__restrict__ box_ind,
                                        const int *__restrict__ level, int which_level, int N, int B,
                                        int H, int W, int ph, int pw, SuperGrid sg, RoiWin *__restrict__ win,
                                        int *__restrict__ st_count)
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
            for (int sy = a0 / sg.side; sy <= a1 / sg.side; ++sy)
                for (int sx = c0 / sg.side; sx <= c1 / sg.side; ++sx)
                    atomicAdd(st_count + ((size_t)b * sg.ny + sy) * sg.nx + sx, 1);
        }
    }
    win[r] = w;
}

// prep 2: exclusive scan of the supertile counts (single CTA; n_st = B*ny*nx)
__global__ void __launch_bounds__(1024) crop_bwd_scan_kernel(const int *__restrict__ st_count, int n_st,
                                                             int *__restrict__ st_off)
{
    __shared__ int s[1024];
    __shared__ int s_carry;
    const int t = threadIdx.x;
    if (t == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n_st; base += 1024) {
        const int v = base + t < n_st ? st_count[base + t] : 0;
        s[t] = v;
        __syncthreads();
        for (int off = 1; off < 1024; off <<= 1) {
            const int add = t >= off ? s[t - off] : 0;
            __syncthreads();
            s[t] += add;
            __syncthreads();
        }
        const int carry = s_carry;
        if (base + t < n_st) st_off[base + t] = carry + s[t] - v;
        __syncthreads();
        if (t == 1023) s_carry = carry + s[1023];
        __syncthreads();
    }
    if (t == 0) st_off[n_st] = s_carry;
}

// prep 3: ordered fill.  One CTA per supertile scans all ROIs in index order (ballot
// compaction keeps the original box order) and stores each hit's window.
struct ListEntry {
    RoiWin win;
    int roi;
};

__global__ void __launch_bounds__(256)
crop_bwd_fill_kernel(const int *__restrict__ box_ind, const RoiWin *__restrict__ win, int N, SuperGrid sg,
                     const int *__restrict__ st_off, ListEntry *__restrict__ entries)
{
    __shared__ int s_warp[8];
    const int st = blockIdx.x;
    const int sx = st % sg.nx, sy = (st / sg.nx) % sg.ny, b = st / (sg.nx * sg.ny);
    const int y0 = sy * sg.side, y1 = y0 + sg.side - 1, x0 = sx * sg.side, x1 = x0 + sg.side - 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int base = st_off[st];
    if (st_off[st + 1] == base) return;
    for (int start = 0; start < N; start += 256) {
        const int r = start + tid;
        bool take = false;
        RoiWin w;
        if (r < N && box_ind[r] == b) {
            w = win[r];
            take = (w.y0 <= w.y1) && !(w.y1 < y0 || w.y0 > y1 || w.x1 < x0 || w.x0 > x1);
        }
        const unsigned m = __ballot_sync(0xffffffffu, take);
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        int off = 0, total = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int c = s_warp[k];
            if (k < warp) off += c;
            total += c;
        }
        if (take) {
            ListEntry e;
            e.win = w;
            e.roi = r;
            entries[base + off + __popc(m & ((1u << lane) - 1u))] = e;
        }
        base += total;
        __syncthreads();
    }
}

template <bool EXACT>
__device__ __forceinline__ float accum1(float s, float g, float wy, float wx, float w)
{
    if (EXACT) return __fadd_rn(s, __fmul_rn(wx, __fmul_rn(wy, g)));   // crop_and_resize.c:241-247
    return __fmaf_rn(w, g, s);
}
template <bool EXACT>
__device__ __forceinline__ float accum(float s, float g, float wy, float wx, float w) { return accum1<EXACT>(s, g, wy, wx, w); }
template <bool EXACT>
__device__ __forceinline__ float4 accum(float4 s, float4 g, float wy, float wx, float w)
{
    return make_float4(accum1<EXACT>(s.x, g.x, wy, wx, w), accum1<EXACT>(s.y, g.y, wy, wx, w),
                       accum1<EXACT>(s.z, g.z, wy, wx, w), accum1<EXACT>(s.w, g.w, wy, wx, w));
}

constexpr int BWD_ROWS = 8;            // tile rows = warps per CTA
constexpr int BWD_THREADS = 32 * BWD_ROWS;
constexpr int BWD_MAX_CH = 64;         // accepted ROIs staged in shared memory per round

// Shared memory per round of at most CH accepted ROIs (dynamic):
//   Tap ytab[CH][ph], xtab[CH][pw]     tap tables
//   int roi[CH]                        accepted ROI ids, original order
//   u16 yr[CH][8], xr[CH][8]           per tile row / column: sample range lo | hi<<8
static size_t bwd_smem_bytes(int CH, int ph, int pw)
{
    return (size_t)CH * ((size_t)(ph + pw) * sizeof(Tap) + sizeof(int) + 2 * 8 * sizeof(unsigned short));
}

// Backward kernel: gather form, no atomics, every destination pixel written exactly once.
//   CTA  = tile of 8 rows x TW columns of one image (x a channel chunk of 32*NV vectors)
//   A    scan the supertile's ordered ROI list, keep (in order) up to CH ROIs whose window meets the tile
//   B    tap tables of the kept ROIs; per tile row/column the contiguous range of samples touching it
//   C    warp = tile row, lanes = channel vectors: for each kept ROI (original order), each of its
//        samples touching the row, each pixel of the row: acc += w * g, in the reference's serial
//        order (ROI, y, x, tap TL/TR/BL/BR; crop_and_resize.c:190-250), accumulators in registers.
// EXACT: each term is wx*(wy*g) with every operation rounded like crop_and_resize.c:241-247 (bit-
// identical to the reference CPU backward); otherwise fma(wy*wx, g, acc) (<= 1 ulp per term).
template <int VEC, int NV, int TW, bool EXACT>
__global__ void __launch_bounds__(BWD_THREADS)
crop_bwd_nhwc_kernel(const float *__restrict__ grads, const float *__restrict__ boxes,
                     const ListEntry *__restrict__ entries, const int *__restrict__ st_off, SuperGrid sg,
                     int C, int ph, int pw, float *__restrict__ grad_image, int B, int H, int W,
                     int tiles_x, int tiles_y, int CH)
{
    using V = typename VecT<VEC>::type;
    extern __shared__ __align__(16) unsigned char s_raw[];
    Tap *ytab = reinterpret_cast<Tap *>(s_raw);                          // [CH][ph]
    Tap *xtab = ytab + (size_t)CH * ph;                                  // [CH][pw]
    int *s_roi = reinterpret_cast<int *>(xtab + (size_t)CH * pw);        // [CH]
    unsigned short *s_yr = reinterpret_cast<unsigned short *>(s_roi + CH);   // [CH][8]
    unsigned short *s_xr = s_yr + (size_t)CH * 8;                        // [CH][8]
    __shared__ int s_warp[BWD_ROWS];
    __shared__ int s_last;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int t = blockIdx.x;
    const int tx_i = t % tiles_x; t /= tiles_x;
    const int ty_i = t % tiles_y; t /= tiles_y;
    const int b = t;
    const int ty0 = ty_i * BWD_ROWS, tx0 = tx_i * TW;
    const int ty1 = min(ty0 + BWD_ROWS, H) - 1, tx1 = min(tx0 + TW, W) - 1;
    const int CV = C / VEC;
    const int cvbase = blockIdx.y * (32 * NV) + lane;

    V acc[TW][NV];
#pragma unroll
    for (int p = 0; p < TW; ++p)
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[p][v] = make_splat(0.f, (V *)nullptr);

    const int st = (b * sg.ny + ty0 / sg.side) * sg.nx + tx0 / sg.side;
    const int l0 = st_off[st], n_list = st_off[st + 1] - l0;
    const ListEntry *__restrict__ list = entries + l0;
    const V *__restrict__ g = reinterpret_cast<const V *>(grads) + cvbase;
    const int S = ph * pw;
    const int py = ty0 + warp;

    int scan = 0;
    while (scan < n_list) {
        // ---- A: accumulate accepted ROIs (in order) until the round is full or the list ends
        int n_acc = 0;
        while (scan < n_list && n_acc < CH) {
            const int li = scan + tid;
            bool take = false;
            int r = -1;
            if (li < n_list) {
                const ListEntry e = list[li];
                r = e.roi;
                take = !(e.win.y1 < ty0 || e.win.y0 > ty1 || e.win.x1 < tx0 || e.win.x0 > tx1);
            }
            const unsigned m = __ballot_sync(0xffffffffu, take);
            if (lane == 0) s_warp[warp] = __popc(m);
            if (tid == 0) s_last = -1;
            __syncthreads();
            int off = 0, total = 0;
#pragma unroll
            for (int k = 0; k < BWD_ROWS; ++k) {
                const int c = s_warp[k];
                if (k < warp) off += c;
                total += c;
            }
            const int slot = n_acc + off + __popc(m & ((1u << lane) - 1u));
            if (take && slot < CH) s_roi[slot] = r;
            const bool overflow = n_acc + total > CH;
            if (overflow && take && slot == CH - 1) s_last = li;     // last entry that still fits
            __syncthreads();
            if (overflow) {
                scan = s_last + 1;
                n_acc = CH;
            } else {
                scan += BWD_THREADS;
                n_acc += total;
            }
            // s_warp / s_last are rewritten next iteration only after the barrier above
        }
        if (n_acc == 0) break;
        const int n_chunk = n_acc;

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
        for (int i = tid; i < n_chunk * (BWD_ROWS + TW); i += BWD_THREADS) {
            const int q = i / (BWD_ROWS + TW), j = i - q * (BWD_ROWS + TW);
            const bool isx = j >= BWD_ROWS;
            const int jj = isx ? j - BWD_ROWS : j;
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
            (isx ? s_xr : s_yr)[q * 8 + jj] = packed;
        }
        __syncthreads();

        // ---- C: accumulate
        if (py <= ty1) {
            for (int q = 0; q < n_chunk; ++q) {
                const unsigned yrng = s_yr[q * 8 + warp];
                const int ylo = yrng & 0xff, yhi = yrng >> 8;
                if (ylo > yhi) continue;
                const Tap *yt = ytab + q * ph;
                const Tap *xt = xtab + q * pw;
                const V *gr = g + (size_t)s_roi[q] * S * CV;
#pragma unroll
                for (int p = 0; p < TW; ++p) {
                    const int px = tx0 + p;
                    const unsigned xrng = s_xr[q * 8 + p];
                    const int xlo = xrng & 0xff, xhi = xrng >> 8;
                    if (xlo > xhi) continue;
                    for (int y = ylo; y <= yhi; ++y) {
                        const Tap tyy = yt[y];
                        const bool top = (tyy.lo == py);
                        const bool bot = (tyy.lo + (tyy.lerp != 0.f) == py);
                        const float wy_t = __fsub_rn(1.f, tyy.lerp), wy_b = tyy.lerp;
                        for (int x = xlo; x <= xhi; ++x) {
                            const Tap txx = xt[x];
                            const bool lft = (txx.lo == px);
                            const bool rgt = (txx.lo + (txx.lerp != 0.f) == px);
                            const float wx_l = __fsub_rn(1.f, txx.lerp), wx_r = txx.lerp;
                            V gv[NV];
#pragma unroll
                            for (int v = 0; v < NV; ++v)
                                if (cvbase + 32 * v < CV) gv[v] = ldg_vec(gr + (size_t)(y * pw + x) * CV + 32 * v);
                            if (top != bot && lft != rgt) {          // the generic case: exactly one tap lands here
                                const float wy = top ? wy_t : wy_b, wx = lft ? wx_l : wx_r;
                                const float w = __fmul_rn(wy, wx);
#pragma unroll
                                for (int v = 0; v < NV; ++v) acc[p][v] = accum<EXACT>(acc[p][v], gv[v], wy, wx, w);
                            } else {                                  // integral sample position: several taps coincide
#pragma unroll
                                for (int tap = 0; tap < 4; ++tap) {  // reference order TL, TR, BL, BR
                                    const bool hit = ((tap >> 1) ? bot : top) && ((tap & 1) ? rgt : lft);
                                    if (hit) {
                                        const float wy = (tap >> 1) ? wy_b : wy_t, wx = (tap & 1) ? wx_r : wx_l;
                                        const float w = __fmul_rn(wy, wx);
#pragma unroll
                                        for (int v = 0; v < NV; ++v) acc[p][v] = accum<EXACT>(acc[p][v], gv[v], wy, wx, w);
                                    }
                                }
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();            // tables are rewritten by the next round
    }

    // ---- write every pixel of the tile exactly once (zeros included)
    if (py <= ty1) {
        V *__restrict__ o = reinterpret_cast<V *>(grad_image);
#pragma unroll
        for (int p = 0; p < TW; ++p) {
            const int px = tx0 + p;
            if (px <= tx1) {
#pragma unroll
                for (int v = 0; v < NV; ++v)
                    if (cvbase + 32 * v < CV) __stcs(o + (((size_t)b * H + py) * W + px) * CV + cvbase + 32 * v, acc[p][v]);
            }
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

struct BwdWs {
    RoiWin *win;
    int *st_count;
    int *st_off;
    ListEntry *entries;
};

static size_t bwd_ws_bytes(int N, int B)
{
    // every ROI can meet at most 64 supertiles of its image
    return align_up(sizeof(RoiWin) * (size_t)N, 256) + 2 * align_up(sizeof(int) * ((size_t)B * 64 + 1), 256) +
           align_up(sizeof(ListEntry) * (size_t)N * 64, 256);
}

template <int VEC, int NV, int TW, bool EXACT>
static int launch_bwd(const float *grads, const float *boxes, const BwdWs &ws, SuperGrid sg, int C, int ph, int pw,
                      float *grad_image, int B, int H, int W, cudaStream_t st)
{
    const int tiles_x = cdiv(W, TW), tiles_y = cdiv(H, BWD_ROWS);
    const int chunks = cdiv(C / VEC, 32 * NV);
    int CH = BWD_MAX_CH;
    while (CH > 1 && bwd_smem_bytes(CH, ph, pw) > 40 * 1024) CH /= 2;
    const size_t smem = bwd_smem_bytes(CH, ph, pw);
    SLN_REQUIRE(smem <= 200 * 1024, SLN_ERR_ARG, "crop %dx%d too large for the backward kernel", ph, pw);
    auto kern = crop_bwd_nhwc_kernel<VEC, NV, TW, EXACT>;
    if (smem > 48 * 1024)
        SLN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SLN_REQUIRE(chunks <= 65535, SLN_ERR_ARG, "too many channel chunks");
    SLN_REQUIRE((size_t)tiles_x * tiles_y * B < (1ull << 31), SLN_ERR_ARG, "too many tiles");
    dim3 grid((unsigned)((size_t)tiles_x * tiles_y * B), chunks);
    kern<<<grid, BWD_THREADS, smem, st>>>(grads, boxes, ws.entries, ws.st_off, sg, C, ph, pw, grad_image, B, H, W,
                                           tiles_x, tiles_y, CH);
    SLN_LAUNCH_OK("crop_bwd_nhwc_kernel");
    return SLN_OK;
}

template <int VEC, bool EXACT>
static int dispatch_bwd(const float *grads, const float *boxes, const BwdWs &ws, SuperGrid sg, int C, int ph, int pw,
                        float *grad_image, int B, int H, int W, cudaStream_t st)
{
    const int CV = C / VEC;
    // two channel vectors per lane halve the control work per byte, but also the CTA count:
    // only use them when the map alone provides enough tiles to fill the machine
    const long long tiles4 = (long long)cdiv(W, 4) * cdiv(H, BWD_ROWS) * B;
    if (CV > 32 && tiles4 >= 6LL * sm_count())
        return launch_bwd<VEC, 2, 4, EXACT>(grads, boxes, ws, sg, C, ph, pw, grad_image, B, H, W, st);
    return launch_bwd<VEC, 1, 4, EXACT>(grads, boxes, ws, sg, C, ph, pw, grad_image, B, H, W, st);
}

static int crop_bwd_nhwc(const float *grads, const float *boxes, const int *box_ind, const int *level,
                         int which_level, int N, int C, int ph, int pw, float *grad_image, int B, int H,
                         int W, bool exact, void *wsp, size_t ws_bytes, cudaStream_t st)
{
    if (B == 0 || C == 0 || H == 0 || W == 0) return SLN_OK;
    SLN_REQUIRE(ws_bytes >= bwd_ws_bytes(N, B), SLN_ERR_WORKSPACE, "crop bwd workspace: need %zu bytes, got %zu",
                bwd_ws_bytes(N, B), ws_bytes);
    SLN_REQUIRE(wsp != nullptr, SLN_ERR_WORKSPACE, "null workspace");
    unsigned char *p = static_cast<unsigned char *>(wsp);
    BwdWs ws;
    ws.win = reinterpret_cast<RoiWin *>(p);        p += align_up(sizeof(RoiWin) * (size_t)N, 256);
    ws.st_count = reinterpret_cast<int *>(p);      p += align_up(sizeof(int) * ((size_t)B * 64 + 1), 256);
    ws.st_off = reinterpret_cast<int *>(p);        p += align_up(sizeof(int) * ((size_t)B * 64 + 1), 256);
    ws.entries = reinterpret_cast<ListEntry *>(p);
    const SuperGrid sg = super_grid(H, W);
    const int n_st = B * sg.nx * sg.ny;

    SLN_CUDA_OK(cudaMemsetAsync(ws.st_count, 0, sizeof(int) * (size_t)(n_st + 1), st));
    if (N > 0) {
        crop_bwd_windows_kernel<<<cdiv(N, 256), 256, 0, st>>>(boxes, box_ind, level, which_level, N, B, H, W, ph, pw,
                                                             sg, ws.win, ws.st_count);
        SLN_LAUNCH_OK("crop_bwd_windows_kernel");
    }
    crop_bwd_scan_kernel<<<1, 1024, 0, st>>>(ws.st_count, n_st, ws.st_off);
    SLN_LAUNCH_OK("crop_bwd_scan_kernel");
    if (N > 0) {
        crop_bwd_fill_kernel<<<n_st, 256, 0, st>>>(box_ind, ws.win, N, sg, ws.st_off, ws.entries);
        SLN_LAUNCH_OK("crop_bwd_fill_kernel");
    }
    const bool vec4 = (C % 4 == 0) && aligned16(grads) && aligned16(grad_image);
    if (vec4) {
        if (exact) return dispatch_bwd<4, true>(grads, boxes, ws, sg, C, ph, pw, grad_image, B, H, W, st);
        return dispatch_bwd<4, false>(grads, boxes, ws, sg, C, ph, pw, grad_image, B, H, W, st);
    }
    if (exact) return dispatch_bwd<1, true>(grads, boxes, ws, sg, C, ph, pw, grad_image, B, H, W, st);
    return dispatch_bwd<1, false>(grads, boxes, ws, sg, C, ph, pw, grad_image, B, H, W, st);
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
