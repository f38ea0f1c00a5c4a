// crop_bwd_tma.cuh -- RoIAlign backward, bulk-async ("TMA") form.  Included by crop.cu inside namespace sln.
//
// Reference semantics: roialign/roi_align/src/crop_and_resize.c:157-252 (serial scatter, order (box, y, x, tap));
// GPU baseline being replaced: cuda/crop_and_resize_kernel.cu:84-165 (atomicAdd scatter).
//
// Why another form: the strip / tile-owner kernels (crop.cu) spend ~7 warp instructions on search and addressing
// per useful FMA and stall on every gradient load (ncu, round 1: 389 M warp instructions for 50 M warp-FMAs at 14x14,
// 48 % of the stall samples on the first FMA after the loads).  Here
//   * one PRODUCER warp per CTA does all the search -- it walks the supertile's ordered ROI list, intersects each
//     ROI's tap tables with the CTA's 8x8 pixel tile, and turns the hit into a plan: a pair descriptor (per tile row
//     the range of sample rows that reach it, per staged sample the tile column of its left tap and its lerp) and one
//     slot per sample row.  The gradient vectors of a slot (the contiguous run of samples sx0..sx1 of crop row sy:
//     n * C * 4 bytes) are fetched by ONE cp.async.bulk issued by one lane into a ring of shared-memory slots; an
//     mbarrier per slot counts the bytes in (complete_tx) and the consumers out.
//   * eight CONSUMER warps own one pixel row of the tile each, with the row's accumulators (8 pixels x C channels)
//     in registers.  A consumer visits only the slots of sample rows that touch its pixel row (it knows them from
//     the pair descriptor), reads each staged vector from shared memory once, and applies both x taps with a
//     warp-uniform switch on the tap's tile column.  No ballots, no shuffles, no global loads, no search.
//   * every output pixel is written exactly once, zeros included (no memset, no atomics).
// Per destination pixel the terms are added in the reference's order (ROI, y, x, tap TL/TR/BL/BR): slots are issued
// in (ROI, sy) order, samples inside a slot ascend in sx, and a pixel row belongs to one warp.  EXACT = the
// reference's rounding sequence wx*(wy*g) per term, incl. the zero-weight taps of integral sample positions
// (bit-identical to crop_and_resize.c); otherwise one fma per term with the fused weight wy*wx.
#pragma once

namespace bwdtma {

constexpr int TH = 8;                  // tile rows = consumer warps
constexpr int TW = 8;                  // tile columns = accumulator pixels per consumer lane
constexpr int NCONS = TH;
constexpr int THREADS = 32 * (NCONS + 1);
#ifndef SLN_BWD_SLOT_S
#define SLN_BWD_SLOT_S 8
#endif
#ifndef SLN_BWD_NSTG
#define SLN_BWD_NSTG 8
#endif
constexpr int SLOT_S = SLN_BWD_SLOT_S; // samples per slot (a longer run of one sample row takes several slots)
constexpr int NSTG = SLN_BWD_NSTG;     // slots in the ring
constexpr int NPAIR = 4;               // pair descriptors in flight
constexpr int MAX_POOL = 32;           // tap tables live one tap per lane
constexpr int CH_MAX = 256;            // channels per CTA (two float4 per consumer lane)

struct SlotDesc {
    float yl;        // lerp of the sample row
    int ylo_rel;     // tile row of its top tap (-1 .. 7)
    int s_begin;     // first staged sample of the slot, relative to the pair's sx0
    int n_s;         // samples in the slot
};
struct XEnt {
    int pl;          // tile column of the left tap (-1 .. 7)
    float xl;        // lerp
};
struct PairDesc {
    int c0;          // ring counter of the pair's first slot
    int nchunk;      // slots per sample row
    int end;         // 1: no more pairs for this tile
    int pad;
    int2 rows[TH];   // per tile row: (first sample row, count), in sample rows relative to the pair's sy0
    XEnt x[MAX_POOL];
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "SLN_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra SLN_DONE;\n"
        "bra SLN_WAIT;\n"
        "SLN_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy (SASS UBLKCP); completion is counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

template <bool EXACT>
__device__ __forceinline__ float4 term(float4 a, float4 g, float w, float wy, float wx)
{
    if (EXACT) {   // crop_and_resize.c:241-247: image += wx * (wy * g), every operation rounded
        return make_float4(__fadd_rn(a.x, __fmul_rn(wx, __fmul_rn(wy, g.x))), __fadd_rn(a.y, __fmul_rn(wx, __fmul_rn(wy, g.y))),
                           __fadd_rn(a.z, __fmul_rn(wx, __fmul_rn(wy, g.z))), __fadd_rn(a.w, __fmul_rn(wx, __fmul_rn(wy, g.w))));
    }
    return make_float4(__fmaf_rn(w, g.x, a.x), __fmaf_rn(w, g.y, a.y), __fmaf_rn(w, g.z, a.z), __fmaf_rn(w, g.w, a.w));
}

// acc[K] += weight * g for a warp-uniform, run-time K: a switch over statically indexed registers
#define SLN_TMA_ADD(K, W, WY, WX)                                                  \
    {                                                                              \
        acc[K][0] = term<EXACT>(acc[K][0], g0, W, WY, WX);                         \
        if (NV == 2) acc[K][1] = term<EXACT>(acc[K][1], g1, W, WY, WX);            \
    }

// FULL: the CTA's channel chunk fills every lane's vectors (CH == 128 * NV), so no lane predicates
template <int NV, bool EXACT, bool FULL>
__global__ void __launch_bounds__(THREADS, 2)
crop_bwd_tma_kernel(const float *__restrict__ grads, const Tap *__restrict__ taps,
                    const ListEntry *__restrict__ entries, const int *__restrict__ st_off,
                    const int *__restrict__ st_count, const BwdLevel *__restrict__ lv_table,
                    BwdTileBases TB, int C, int ph, int pw)
{
    extern __shared__ __align__(128) unsigned char s_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // ---- which tile
    int l = TB.lvl[0];
#pragma unroll
    for (int k = 1; k < BWD_MAX_LEVELS; ++k)
        if (k < TB.n_levels && (int)blockIdx.x >= TB.base[k]) l = TB.lvl[k];
    const BwdLevel L = lv_table[l];
    const int t = blockIdx.x - L.tile_base;
    const int q1 = fast_div(t, L.tiles_x, L.rcp_tiles_x);
    const int tx_i = t - q1 * L.tiles_x;
    const int b = fast_div(q1, L.tiles_y, L.rcp_tiles_y);
    const int ty_i = q1 - b * L.tiles_y;
    const int H = L.H, W = L.W;
    const int y0 = ty_i * TH, x0 = tx_i * TW;
    const int y1 = min(y0 + TH, H) - 1, x1 = min(x0 + TW, W) - 1;
    const int c0 = blockIdx.y * CH_MAX;                    // first channel of this CTA
    const int CH = min(C - c0, CH_MAX);
    const int chb = CH * 4;                                // bytes per staged sample
    const int slot_bytes = SLOT_S * chb;

    // ---- shared memory carve-up
    unsigned char *slots = s_raw;                                                   // [NSTG][SLOT_S * chb]
    SlotDesc *sdesc = reinterpret_cast<SlotDesc *>(s_raw + (size_t)NSTG * SLOT_S * CH_MAX * 4);
    PairDesc *pdesc = reinterpret_cast<PairDesc *>(sdesc + NSTG);
    uint64_t *bars = reinterpret_cast<uint64_t *>(pdesc + NPAIR);
    // slots issued so far.  A consumer learns about a slot from the pair descriptor, which is published BEFORE the
    // pair's slots are issued (a pair may need more slots than the ring has), so it could reach the parity wait of
    // use c of a slot while use c - NSTG is still unfilled -- and a parity wait cannot tell "two phases behind"
    // from "done".  It therefore first waits until the producer has issued slot c (which the producer only does
    // after use c - NSTG was released), and only then on the slot's mbarrier.
    volatile unsigned *issued = reinterpret_cast<volatile unsigned *>(bars + 2 * NSTG + 2 * NPAIR);
    const uint32_t bar0 = smem_u32(bars);
    // slot_full[i] = bar0 + 8 i, slot_empty[i] = bar0 + 8 (NSTG + i), pair_full[q] = bar0 + 8 (2 NSTG + q),
    // pair_empty[q] = bar0 + 8 (2 NSTG + NPAIR + q)
    const uint32_t slot_full = bar0, slot_empty = bar0 + 8 * NSTG, pair_full = bar0 + 16 * NSTG,
                   pair_empty = bar0 + 16 * NSTG + 8 * NPAIR;
    if (threadIdx.x == 0) {
        for (int i = 0; i < NSTG; ++i) {
            mbar_init(slot_full + 8 * i, 1);               // the producer's arrive.expect_tx (+ the bytes)
            mbar_init(slot_empty + 8 * i, 2);              // the two pixel rows a sample row reaches
        }
        for (int q = 0; q < NPAIR; ++q) {
            mbar_init(pair_full + 8 * q, 1);
            mbar_init(pair_empty + 8 * q, NCONS);
        }
        *issued = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    if (warp == NCONS) {
        // =========================================================== producer
        const int st = L.st_base + (b * L.sg.ny + (y0 >> L.sg_shift)) * L.sg.nx + (x0 >> L.sg_shift);
        const int n_list = st_count[st];
        const ListEntry *__restrict__ list = entries + (n_list ? st_off[st] : 0);
        const int S = ph * pw;
        unsigned counter = 0, pcount = 0;
        for (int base = 0; base < n_list; base += 32) {
            const int li = base + lane;
            int my_roi = -1;
            bool take = false;
            if (li < n_list) {
                const ListEntry e = list[li];
                my_roi = e.roi;
                take = !(e.win.y1 < y0 || e.win.y0 > y1 || e.win.x1 < x0 || e.win.x0 > x1);
            }
            unsigned todo = __ballot_sync(0xffffffffu, take);
            if (!todo) continue;
            // tap tables of the next ROI are in flight while the current one is planned
            Tap ty_n, tx_n;
            auto fetch_taps = [&](int r) {
                const Tap *__restrict__ tp = taps + (size_t)r * (ph + pw);
                ty_n.lo = INVALID_TAP; ty_n.lerp = 0.f; tx_n.lo = INVALID_TAP; tx_n.lerp = 0.f;
                if (lane < ph) ty_n = ld_tap(tp + lane);
                if (lane < pw) tx_n = ld_tap(tp + ph + lane);
            };
            int r_n = __shfl_sync(0xffffffffu, my_roi, __ffs(todo) - 1);
            todo &= todo - 1;
            fetch_taps(r_n);
            while (r_n >= 0) {
                const int r = r_n;
                const Tap ty = ty_n, tx = tx_n;
                r_n = -1;
                if (todo) {
                    r_n = __shfl_sync(0xffffffffu, my_roi, __ffs(todo) - 1);
                    todo &= todo - 1;
                    fetch_taps(r_n);
                }
                const bool yv = ty.lo != INVALID_TAP, xv = tx.lo != INVALID_TAP;
                const int yhi = ty.lo + (ty.lerp != 0.f), xhi = tx.lo + (tx.lerp != 0.f);
                const unsigned ym = __ballot_sync(0xffffffffu, yv && ty.lo <= y1 && yhi >= y0);
                const unsigned xm = __ballot_sync(0xffffffffu, xv && tx.lo <= x1 && xhi >= x0);
                if (!ym || !xm) continue;
                // sample positions are monotone in the sample index, so both masks are contiguous runs
                const int sy0 = __ffs(ym) - 1, nsy = __popc(ym), sx0 = __ffs(xm) - 1, nsx = __popc(xm);
                const int nchunk = (nsx + SLOT_S - 1) / SLOT_S;
#ifdef SLN_BWD_DEBUG
                if (lane == 0 && ((ym >> sy0) != ((nsy == 32) ? 0xffffffffu : ((1u << nsy) - 1u)) ||
                                  (xm >> sx0) != ((nsx == 32) ? 0xffffffffu : ((1u << nsx) - 1u)))) {
                    printf("producer: blk %d roi %d non-contiguous masks ym %08x xm %08x\n", blockIdx.x, r, ym, xm);
                    __trap();
                }
#endif
                const unsigned q = pcount % NPAIR;
                mbar_wait(pair_empty + 8 * q, ((pcount / NPAIR) & 1) ^ 1);
                PairDesc *pd = pdesc + q;
                if ((xm >> lane) & 1u) {
                    XEnt e;
                    e.pl = tx.lo - x0;
                    e.xl = tx.lerp;
                    pd->x[lane - sx0] = e;
                }
                unsigned mine = 0;
#pragma unroll
                for (int j = 0; j < TH; ++j) {
                    const unsigned mj = __ballot_sync(0xffffffffu, yv && (ty.lo == y0 + j || (yhi == y0 + j && ty.lo != yhi)));
                    if (lane == j) mine = mj;
                }
                if (lane < TH) pd->rows[lane] = make_int2(mine ? __ffs(mine) - 1 - sy0 : 0, __popc(mine));
                if (lane == 0) {
                    pd->c0 = (int)counter;
                    pd->nchunk = nchunk;
                    pd->end = 0;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(pair_full + 8 * q);
                ++pcount;
                const float *gr = grads + ((size_t)r * S + sx0) * C + c0;
                for (int i = 0; i < nsy; ++i) {
                    const int sy = sy0 + i;
                    const int ylo = __shfl_sync(0xffffffffu, ty.lo, sy);
                    const float yl = __shfl_sync(0xffffffffu, ty.lerp, sy);
                    const int rel = ylo - y0;
                    const int n_cons = (rel >= 0 ? 1 : 0) + ((yl != 0.f && rel + 1 < TH) ? 1 : 0);   // rel <= 7, rel + 1 >= 0 here
#ifdef SLN_BWD_DEBUG
                    if (lane == 0 && (n_cons < 1 || rel < -1 || rel > 7)) {
                        printf("producer: blk %d roi %d row %d rel %d yl %f n_cons %d ym %08x\n", blockIdx.x, r, sy, rel, yl, n_cons, ym);
                        __trap();
                    }
#endif
                    for (int c = 0; c < nchunk; ++c) {
                        const unsigned slot = counter % NSTG;
                        mbar_wait(slot_empty + 8 * slot, ((counter / NSTG) & 1) ^ 1);
                        const int sb = c * SLOT_S, n_s = min(SLOT_S, nsx - sb);
                        const uint32_t dst = smem_u32(slots + (size_t)slot * slot_bytes);
                        const float *src = gr + ((size_t)sy * pw + sb) * C;
                        if (lane == 0) {
                            SlotDesc d;
                            d.yl = yl; d.ylo_rel = rel; d.s_begin = sb; d.n_s = n_s;
                            sdesc[slot] = d;
                            if (n_cons == 1) mbar_arrive(slot_empty + 8 * slot);     // the other row is outside the tile
                            mbar_arrive_expect_tx(slot_full + 8 * slot, (uint32_t)(n_s * chb));
                            if (CH == C) bulk_g2s(dst, src, (uint32_t)(n_s * chb), slot_full + 8 * slot);
                        }
                        if (CH != C) {                       // channel chunk of a wider map: one copy per sample
                            __syncwarp();
                            if (lane < n_s) bulk_g2s(dst + lane * chb, src + (size_t)lane * C, (uint32_t)chb, slot_full + 8 * slot);
                        }
                        ++counter;
                        __syncwarp();
                        if (lane == 0) *issued = counter;
                    }
                }
            }
        }
        const unsigned q = pcount % NPAIR;
        mbar_wait(pair_empty + 8 * q, ((pcount / NPAIR) & 1) ^ 1);
        if (lane == 0) {
            pdesc[q].end = 1;
            mbar_arrive(pair_full + 8 * q);
        }
        return;
    }

    // =============================================================== consumers
    const int j = warp;                                     // tile row
    const bool ok0 = FULL || lane * 4 < CH, ok1 = NV == 2 && (FULL || lane * 4 + 128 < CH);
    float4 acc[TW][NV];
#pragma unroll
    for (int k = 0; k < TW; ++k)
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[k][v] = make_float4(0.f, 0.f, 0.f, 0.f);

    for (unsigned pc = 0;; ++pc) {
        const unsigned q = pc % NPAIR;
        mbar_wait(pair_full + 8 * q, (pc / NPAIR) & 1);
        const PairDesc *pd = pdesc + q;
        const int4 hdr = *reinterpret_cast<const int4 *>(pd);      // c0, nchunk, end
        if (hdr.z) break;
        const int2 rr = pd->rows[j];
        const int t_end = (rr.x + rr.y) * hdr.y;
        for (int tt = rr.x * hdr.y; tt < t_end; ++tt) {
            const unsigned c = (unsigned)hdr.x + (unsigned)tt;
            const unsigned slot = c % NSTG;
            while ((int)(*issued - c) <= 0) {}
            mbar_wait(slot_full + 8 * slot, (c / NSTG) & 1);
            const SlotDesc sd = sdesc[slot];
#ifdef SLN_BWD_DEBUG
            if (lane == 0 && (sd.n_s < 1 || sd.n_s > SLOT_S || sd.s_begin < 0 || sd.s_begin >= MAX_POOL ||
                              (sd.ylo_rel != j && sd.ylo_rel != j - 1))) {
                printf("consumer: blk %d warp %d pc %u c %u (c0 %d nchunk %d rows %d+%d tt %d) slot %u sd(yl %f rel %d sb %d n %d) issued %u\n",
                       blockIdx.x, j, pc, c, hdr.x, hdr.y, rr.x, rr.y, tt, slot, sd.yl, sd.ylo_rel, sd.s_begin, sd.n_s, *issued);
                __trap();
            }
#endif
            const bool top = sd.ylo_rel == j;
            const float wy = top ? __fsub_rn(1.f, sd.yl) : sd.yl;
            // integral sample row: the reference's bottom taps land on the same pixel row with weight 0
            const int npass = (EXACT && top && sd.yl == 0.f) ? 2 : 1;
            const unsigned char *sp = slots + (size_t)slot * slot_bytes + lane * 16;
            const XEnt *xe = pd->x + sd.s_begin;
            for (int s = sd.n_s; s > 0; --s, sp += chb, ++xe) {
                const XEnt e = *xe;
                float4 g0, g1;
                if (FULL) {
                    g0 = *reinterpret_cast<const float4 *>(sp);
                    if (NV == 2) g1 = *reinterpret_cast<const float4 *>(sp + 512);
                } else {
                    g0 = make_float4(0.f, 0.f, 0.f, 0.f);
                    g1 = g0;
                    if (ok0) g0 = *reinterpret_cast<const float4 *>(sp);
                    if (ok1) g1 = *reinterpret_cast<const float4 *>(sp + 512);
                }
                const float wl = __fsub_rn(1.f, e.xl);
                if (!EXACT) {
                    const float a = __fmul_rn(wy, wl), bb = __fmul_rn(wy, e.xl);
                    switch (e.pl) {
                    case -1: SLN_TMA_ADD(0, bb, 0.f, 0.f) break;
                    case 0: SLN_TMA_ADD(0, a, 0.f, 0.f) SLN_TMA_ADD(1, bb, 0.f, 0.f) break;
                    case 1: SLN_TMA_ADD(1, a, 0.f, 0.f) SLN_TMA_ADD(2, bb, 0.f, 0.f) break;
                    case 2: SLN_TMA_ADD(2, a, 0.f, 0.f) SLN_TMA_ADD(3, bb, 0.f, 0.f) break;
                    case 3: SLN_TMA_ADD(3, a, 0.f, 0.f) SLN_TMA_ADD(4, bb, 0.f, 0.f) break;
                    case 4: SLN_TMA_ADD(4, a, 0.f, 0.f) SLN_TMA_ADD(5, bb, 0.f, 0.f) break;
                    case 5: SLN_TMA_ADD(5, a, 0.f, 0.f) SLN_TMA_ADD(6, bb, 0.f, 0.f) break;
                    case 6: SLN_TMA_ADD(6, a, 0.f, 0.f) SLN_TMA_ADD(7, bb, 0.f, 0.f) break;
                    default: SLN_TMA_ADD(7, a, 0.f, 0.f) break;
                    }
                } else {
                    // reference order per sample: TL, TR (this row as the top row), then BL, BR when the sample row is
                    // integral; the right tap coincides with the left one when the sample column is integral
                    const int pr = e.pl + (e.xl != 0.f);
                    for (int pass = 0; pass < npass; ++pass) {
                        const float w_y = pass == 0 ? wy : sd.yl;
#pragma unroll
                        for (int side = 0; side < 2; ++side) {
                            const int p = side == 0 ? e.pl : pr;
                            const float w_x = side == 0 ? wl : e.xl;
                            switch (p) {
                            case 0: SLN_TMA_ADD(0, 0.f, w_y, w_x) break;
                            case 1: SLN_TMA_ADD(1, 0.f, w_y, w_x) break;
                            case 2: SLN_TMA_ADD(2, 0.f, w_y, w_x) break;
                            case 3: SLN_TMA_ADD(3, 0.f, w_y, w_x) break;
                            case 4: SLN_TMA_ADD(4, 0.f, w_y, w_x) break;
                            case 5: SLN_TMA_ADD(5, 0.f, w_y, w_x) break;
                            case 6: SLN_TMA_ADD(6, 0.f, w_y, w_x) break;
                            case 7: SLN_TMA_ADD(7, 0.f, w_y, w_x) break;
                            default: break;                  // -1 / 8: the neighbouring tile's pixel
                            }
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(slot_empty + 8 * slot);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(pair_empty + 8 * q);
    }

    // ---- write the row exactly once (zeros included)
    const int py = y0 + j;
    if (py > y1) return;
    float4 *__restrict__ o = reinterpret_cast<float4 *>(L.out + (((size_t)b * H + py) * W + x0) * C + c0) + lane;
    const size_t pix = (size_t)C / 4;
#pragma unroll
    for (int k = 0; k < TW; ++k) {
        if (x0 + k > x1) break;
        if (ok0) __stcs(o + k * pix, acc[k][0]);
        if (ok1) __stcs(o + k * pix + 32, acc[k][1]);
    }
}
#undef SLN_TMA_ADD

static size_t smem_bytes()
{
    return (size_t)NSTG * SLOT_S * CH_MAX * 4 + sizeof(SlotDesc) * NSTG + sizeof(PairDesc) * NPAIR +
           sizeof(uint64_t) * (2 * NSTG + 2 * NPAIR) + 16 + 128;
}

}  // namespace bwdtma
