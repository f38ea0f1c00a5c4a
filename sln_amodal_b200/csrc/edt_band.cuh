// edt_band.cuh -- per-column logic of the banded exact EDT (A14).
//
// The column pass D(y,x) = min_s (y-s)^2 + g(s,x)^2 is cut into 32-row BANDS so that no warp ever walks a
// chain longer than 32 rows:
//   build : one warp per (32-column segment, band) builds the lower envelope of the band's own sites
//           (foreground pixels with a finite row distance, plus the zero pixels that bound a vertical run)
//           as a stack of packed entries (s:11 | t:11 | g:10), t = first row from which the entry wins.
//   eval  : one warp per (segment, band) evaluates its 32 rows as the minimum over its own envelope and the
//           envelopes of the bands its vertical runs continue into, nearest band first, until the row gap
//           alone exceeds what is already known.
// What makes the band envelopes usable from OUTSIDE their band:
//   * a run that continues ABOVE the band ("top-open") is built with its first breakpoint clamped to row 0
//     instead of the band's first row, so its entries are the true lower envelope for every row >= 0 --
//     the bands above read the leading entries [0, n_first);
//   * a run that continues BELOW the band ("bottom-open") keeps every entry that wins anywhere up to row
//     H-1 -- the bands below read the trailing entries [last_base, total);
//   * sites beyond a zero pixel of the column can never beat that zero pixel, so a run only ever needs the
//     bands it spans, and any SUPERSET of the useful bands is exact: the pruning (row gap squared >= the
//     largest value still standing) only saves work.
// Everything is integer and exact.  The functions below are written per LANE (= one column) and compile
// for the device and for the host: tests/host_sim/edt_band_sim.cpp runs them on the CPU against the oracle,
// because the GPU is minutes away from the build container.  Warp-level shortcuts go through SLN_WARP_ANY,
// which is a vote on the device and the lane's own predicate on the host (an inactive lane stays inactive,
// so both readings execute the same per-lane work).
//
// Test-infrastructure note: the host instantiation is a SIMULATION OF THE DEVICE CODE for CPU tests, not a
// CPU fallback -- nothing in the product calls it.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define SLN_HD __host__ __device__ __forceinline__
#else
#define SLN_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define SLN_WARP_ANY(p) __any_sync(0xffffffffu, (p))
#else
#define SLN_WARP_ANY(p) (p)
#endif

namespace sln {
namespace edtband {

constexpr int BAND = 32;             // rows per band
constexpr int SLOTS = 34;            // stack slots per (band, column): 32 rows + the bounding zero above and below
constexpr int GQ_INF = 0xffff;       // row distance of a pixel whose row holds no zero pixel
constexpr int NONE_D = 1 << 20;      // "no zero pixel on that side of the segment"
constexpr int T_NEVER = 1 << 20;     // breakpoint of a non-existent next entry
constexpr unsigned FULLW = 0xffffffffu;

SLN_HD int hd_min(int a, int b) { return a < b ? a : b; }
SLN_HD int hd_max(int a, int b) { return a > b ? a : b; }

SLN_HD int hd_clz(unsigned v)        // v != 0
{
#if defined(__CUDA_ARCH__)
    return __clz((int)v);
#else
    return __builtin_clz(v);
#endif
}
SLN_HD int hd_ffs(unsigned v)        // 1-based, 0 for v == 0
{
#if defined(__CUDA_ARCH__)
    return __ffs((int)v);
#else
    return __builtin_ffs((int)v);
#endif
}

// floor(a / b) for |a| < 2^23, 0 < b < 2^13: one reciprocal and an exact fix-up
SLN_HD int floor_div_small(int a, int b)
{
#if defined(__CUDA_ARCH__)
    int q = __float2int_rd(__fdividef((float)a, (float)b));      // within 1 of the true floor
#else
    int q = (int)__builtin_floorf((float)a / (float)b);
#endif
    const int r = a - q * b;
    if (r < 0) --q; else if (r >= b) ++q;
    return q;
}

// ---- packed envelope entries
SLN_HD int pk_s(unsigned e) { return (int)(e & 0x7ffu); }
SLN_HD int pk_t(unsigned e) { return (int)((e >> 11) & 0x7ffu); }
SLN_HD int pk_g2(unsigned e) { const int gq = (int)(e >> 22); return gq * gq; }
SLN_HD unsigned pk_make(int s, int t, int gq) { return (unsigned)s | ((unsigned)t << 11) | ((unsigned)gq << 22); }
SLN_HD int env_f(int y, int s, int g2) { return (y - s) * (y - s) + g2; }
SLN_HD int env_at(int y, unsigned e) { return env_f(y, pk_s(e), pk_g2(e)); }

// ---- per (band, column) word written by the build pass next to the column's foreground bits
//   n_first   : entries [0, n_first) belong to a first run that continues above the band (0 otherwise)
//   last_base : entries [last_base, total) belong to a last run that continues below the band (= total otherwise)
//   gmin      : smallest row distance among the column's sites in this band (1023: no site) -- no candidate from this
//               band can be below (row gap)^2 + gmin^2, which lets the evaluation pass skip the band without a walk
SLN_HD unsigned meta_make(int n_first, int last_base, int total, bool top_open, bool bot_open, int gmin)
{
    return (unsigned)n_first | ((unsigned)last_base << 6) | ((unsigned)total << 12) | ((unsigned)top_open << 18) |
           ((unsigned)bot_open << 19) | ((unsigned)gmin << 20);
}
SLN_HD int meta_gmin(unsigned m) { return (int)((m >> 20) & 1023u); }
SLN_HD int meta_n_first(unsigned m) { return (int)(m & 63u); }
SLN_HD int meta_last_base(unsigned m) { return (int)((m >> 6) & 63u); }
SLN_HD int meta_total(unsigned m) { return (int)((m >> 12) & 63u); }
SLN_HD bool meta_top_open(unsigned m) { return (m >> 18) & 1u; }
SLN_HD bool meta_bot_open(unsigned m) { return (m >> 19) & 1u; }

// Row distance of pixel `lane` of a 32-pixel segment.  z: zero mask of the segment (bit k <=> pixel k is zero);
// ldist: distance from pixel 0 to the nearest zero left of the segment (>= 1, NONE_D if none); rdist: from pixel 31
// to the nearest zero right of it.  Finite distances are <= 1023 on the shapes this path takes (W <= 1024).
SLN_HD int row_dist(unsigned z, int lane, int ldist, int rdist)
{
    if ((z >> lane) & 1u) return 0;
    const unsigned below = z & ((1u << lane) - 1u);
    const unsigned above = z >> lane;                              // bit 0 (the pixel itself) is clear
    const int dl = below ? lane - (31 - hd_clz(below)) : lane + ldist;
    const int dr = above ? hd_ffs(above) - 1 : (31 - lane) + rdist;
    const int d = hd_min(dl, dr);
    return d > 1023 ? GQ_INF : d;
}

// ---- the column's envelope stack while it is being built: entries live in sc[k * stride]; the top three are
// mirrored in registers so that a pop never waits for the entry it uncovers (refill issued two pops ahead)
constexpr int PK_D = 3;
struct Stack {
    int q, base, ystart;         // top index (-1: empty), first entry and first breakpoint of the open run
    unsigned *top;               // address of entry q (walked, never recomputed: 64-bit multiplies cost more than the pops)
    unsigned e[PK_D];            // entries q, q-1, q-2 (garbage below index 0)
    bool open;
};

SLN_HD void pk_pop(Stack &c, int stride)
{
    --c.q;
    c.top -= stride;
#pragma unroll
    for (int i = 0; i + 1 < PK_D; ++i) c.e[i] = c.e[i + 1];
    if (c.q >= PK_D - 1) c.e[PK_D - 1] = *(c.top - (PK_D - 1) * stride);
}

SLN_HD void pk_push(Stack &c, int stride, unsigned e)
{
    ++c.q;
    c.top += stride;
#pragma unroll
    for (int i = PK_D - 1; i > 0; --i) c.e[i] = c.e[i - 1];
    c.e[0] = e;
    *c.top = e;
}

// add the parabola of site (u, gq); entries that would only win after row `limit` are dropped
SLN_HD void pk_insert(Stack &c, int stride, int u, int gq, int limit)
{
    const int gu2 = gq * gq;
    while (c.q >= c.base) {
        const int t = pk_t(c.e[0]);
        if (env_at(t, c.e[0]) > env_f(t, u, gu2)) pk_pop(c, stride);
        else break;
    }
    if (c.q < c.base) {
        pk_push(c, stride, pk_make(u, c.ystart, gq));
    } else {
        const int st = pk_s(c.e[0]);
        const int w = 1 + floor_div_small(u * u - st * st + gu2 - pk_g2(c.e[0]), 2 * (u - st));
        if (w <= limit) pk_push(c, stride, pk_make(u, w, gq));
    }
}

struct BuildResult {
    unsigned fgw;                // bit r <=> pixel (yb + r, x) is foreground
    unsigned meta;
};

// Build pass of one column of one band.
//   f        : the segment's flag word (bit r <=> row yb + r holds a foreground pixel in this 32-column segment)
//   rows     : rows of the band inside the map (32 except for a cut last band)
//   above_zero / above_fg : the pixel right above the band is a zero pixel / a foreground pixel (both false at y = 0)
//   below_zero / below_fg : same for the pixel right below the band
//   row(r, z, ld, rd)     : zero mask and outside distances of row yb + r (only called for rows with the flag bit set)
template <class RowFn>
SLN_HD BuildResult band_build_lane(unsigned f, int rows, int yb, int H, int lane, bool above_zero, bool above_fg,
                                   bool below_zero, bool below_fg, unsigned *sc, int stride, RowFn row)
{
    Stack c;
    c.q = -1; c.base = 0; c.ystart = 0; c.open = false;
    c.top = sc - stride;
#pragma unroll
    for (int i = 0; i < PK_D; ++i) c.e[i] = 0u;
    unsigned fgw = 0u;
    int n_first = 0, gmin = 1023;
    bool first_open = false, top_open = false;
    for (int r = 0; r < rows; ++r) {
        const bool rowbit = (f >> r) & 1u;
        if (!SLN_WARP_ANY(rowbit || c.open)) continue;              // 32 background pixels, nothing open
        int gv = 0;
        if (rowbit) {
            unsigned z;
            int ld, rd;
            row(r, z, ld, rd);
            gv = row_dist(z, lane, ld, rd);
        }
        const int y = yb + r;
        if (gv == 0) {
            if (c.open) {                                           // the zero pixel at y closes the run
                pk_insert(c, stride, y, 0, y - 1);
                gmin = 0;
                c.open = false;
                if (first_open) { n_first = c.q + 1; first_open = false; }
            }
        } else {
            fgw |= 1u << r;
            if (!c.open) {
                c.open = true;
                c.base = c.q + 1;
                if (r == 0 && !above_zero) {
                    c.ystart = 0;                                   // continues above the band (or starts at the map's top)
                    if (above_fg) { top_open = true; first_open = true; }
                } else {
                    c.ystart = y;
                    pk_insert(c, stride, y - 1, 0, H - 1);      // the zero pixel just above the run
                    gmin = 0;
                }
            }
            if (gv != GQ_INF) {
                pk_insert(c, stride, y, gv, H - 1);
                gmin = hd_min(gmin, gv);
            }
        }
    }
    bool bot_open = false;
    int last_base = 0;
    if (c.open) {
        if (below_zero) {
            pk_insert(c, stride, yb + rows, 0, yb + rows - 1);
            gmin = 0;
            c.open = false;
            if (first_open) { n_first = c.q + 1; first_open = false; }
        } else if (below_fg) {
            bot_open = true;
            last_base = c.base;
        }
    }
    const int total = c.q + 1;
    if (first_open) n_first = total;
    if (!bot_open) last_base = total;
    BuildResult res;
    res.fgw = fgw;
    res.meta = meta_make(n_first, last_base, total, top_open, bot_open, gmin);
    return res;
}

// two words per (band, column): foreground bits and the meta word
struct alignas(8) Words2 {
    unsigned x, y;
};

SLN_HD unsigned ld_ro(const unsigned *p)
{
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
SLN_HD Words2 ld_ro(const Words2 *p)
{
#if defined(__CUDA_ARCH__)
    const uint2 t = __ldg(reinterpret_cast<const uint2 *>(p));
    Words2 w;
    w.x = t.x;
    w.y = t.y;
    return w;
#else
    return *p;
#endif
}

// Evaluation pass of one column of one band.  best[r * best_stride] receives the squared distance of pixel (yb + r, x)
// for the column's foreground rows (`cap` where the column's runs see no zero pixel at all; other rows: unspecified):
// the own-band walk stores every row once, the bands above / below then lower the rows of the runs that continue into them.
//   own[k * own_stride]                                  : entry k of this band's stack (staged in shared memory)
//   stk_col + bb * band_stride                           : band bb's stack, same column (entry k at + k * slot_stride)
//   meta_col[bb * meta_stride]                           : the two words of band bb, same column
//   stage(bp, lo, hi, need, ptr, stride)                 : warp-wide; makes entries [lo, hi) of the stack at bp readable
//                                                          as ptr[k * stride] for the lanes with `need`
// Why the staging: a row's winner in a neighbouring band is usually a DIFFERENT entry for every row (inside a disc the
// winner of a row sits tens of rows further out), so a walk touches most of that band's entries one after the other --
// as dependent L2 / DRAM round trips of a lone warp that was 30 us per walk; staged, the loads are all in flight at once.
// Everything is addressed by walking pointers (index lambdas cost 50-60 instructions per row in address arithmetic).
template <class StageFn>
SLN_HD void band_eval_lane(int b, int nb, int yb, int rows, int cap, unsigned fgw, unsigned mw, const unsigned *own,
                           int own_stride, const unsigned *stk_col, size_t band_stride, int slot_stride,
                           const Words2 *meta_col, int meta_stride, int *best, int best_stride, StageFn stage)
{
    const int total = meta_total(mw);
    const int n1 = fgw == FULLW ? BAND : hd_ffs(~fgw) - 1;              // rows of the run that touches the band's top
    const int nl = fgw == FULLW ? BAND : hd_clz(~fgw);                  // rows of the run that touches its bottom
    int mx_up = 0, mx_dn = 0;
    {                                                                   // the band's own envelope: one monotone walk
        const unsigned *op = own;
        const unsigned *const last = own + (total - 1) * own_stride;
        unsigned e = total > 0 ? *op : 0u;
        int s = pk_s(e), g2 = total > 0 ? pk_g2(e) : cap;
        int tn = total > 1 ? pk_t(op[own_stride]) : T_NEVER;
        int *o = best;
        const int d0 = total > 0 ? 1 : 0;                               // no entries: every row reads `cap`
        for (int r = 0; r < rows; ++r, o += best_stride) {
            const int y = yb + r;
            while (tn <= y) {
                op += own_stride;
                e = *op;
                s = pk_s(e);
                g2 = pk_g2(e);
                tn = op != last ? pk_t(op[own_stride]) : T_NEVER;
            }
            const int dy = (y - s) * d0;
            const int v = dy * dy + g2;
            *o = v;
            if (r < n1) mx_up = hd_max(mx_up, v);
            if (r >= BAND - nl) mx_dn = hd_max(mx_dn, v);
        }
    }
    // ---- bands above: rows of the first run, trailing entries of those bands
    {                                   // (every lane runs the loop: the votes and the staging inside are warp-wide)
        bool cont = meta_top_open(mw);
        int mx = mx_up;
        const Words2 *mp = meta_col + (size_t)b * meta_stride;
        const unsigned *bp = stk_col + (size_t)b * band_stride;
        Words2 wn;
        wn.x = wn.y = 0u;
        if (b >= 1) wn = ld_ro(mp - meta_stride);                       // the next band's words, one band ahead
        for (int d = 1; b - d >= 0; ++d) {
            mp -= meta_stride;
            bp -= band_stride;
            const int gap = BAND * (d - 1) + 1;
            const bool act = cont && gap * gap < mx;
            if (!SLN_WARP_ANY(act)) break;
            const Words2 w2 = wn;
            if (b - d >= 1) wn = ld_ro(mp - meta_stride);
            const int tot2 = meta_total(w2.y), lb2 = meta_last_base(w2.y), gm = meta_gmin(w2.y);
            const bool need = act && tot2 > lb2 && gap * gap + gm * gm < mx;
            if (SLN_WARP_ANY(need)) {
                const unsigned *ep;
                int es;
                stage(bp, lb2, tot2, need, ep, es);
                if (need) {
                    const unsigned *sp = ep + (tot2 - 1) * es;
                    const unsigned *const first = ep + lb2 * es;
                    unsigned e = *sp;
                    int s = pk_s(e), g2 = pk_g2(e), t = pk_t(e);
                    int *o = best + (n1 - 1) * best_stride;
                    mx = 0;
                    for (int y = yb + n1 - 1; y >= yb; --y, o -= best_stride) {
                        while (t > y && sp != first) {
                            sp -= es;
                            e = *sp;
                            s = pk_s(e);
                            g2 = pk_g2(e);
                            t = pk_t(e);
                        }
                        const int c = (y - s) * (y - s) + g2;
                        int v = *o;
                        if (c < v) {
                            v = c;
                            *o = c;
                        }
                        mx = hd_max(mx, v);
                    }
                }
            }
            if (act) cont = w2.x == FULLW && meta_top_open(w2.y);
        }
    }
    // ---- bands below: rows of the last run, leading entries of those bands
    {
        bool cont = meta_bot_open(mw);
        int mx = mx_dn;
        const Words2 *mp = meta_col + (size_t)b * meta_stride;
        const unsigned *bp = stk_col + (size_t)b * band_stride;
        Words2 wn;
        wn.x = wn.y = 0u;
        if (b + 1 < nb) wn = ld_ro(mp + meta_stride);
        for (int d = 1; b + d < nb; ++d) {
            mp += meta_stride;
            bp += band_stride;
            const int gap = BAND * (d - 1) + 1;
            const bool act = cont && gap * gap < mx;
            if (!SLN_WARP_ANY(act)) break;
            const Words2 w2 = wn;
            if (b + d + 1 < nb) wn = ld_ro(mp + meta_stride);
            const int nf2 = meta_n_first(w2.y), gm = meta_gmin(w2.y);
            const bool need = act && nf2 > 0 && gap * gap + gm * gm < mx;
            if (SLN_WARP_ANY(need)) {
                const unsigned *ep;
                int es;
                stage(bp, 0, nf2, need, ep, es);
                if (need) {
                    const unsigned *sp = ep;
                    const unsigned *const last = ep + (nf2 - 1) * es;
                    unsigned e = *sp;
                    int s = pk_s(e), g2 = pk_g2(e);
                    int tn = nf2 > 1 ? pk_t(sp[es]) : T_NEVER;
                    int *o = best + (BAND - nl) * best_stride;
                    mx = 0;
                    for (int y = yb + BAND - nl; y < yb + BAND; ++y, o += best_stride) {
                        while (tn <= y) {
                            sp += es;
                            e = *sp;
                            s = pk_s(e);
                            g2 = pk_g2(e);
                            tn = sp != last ? pk_t(sp[es]) : T_NEVER;
                        }
                        const int c = (y - s) * (y - s) + g2;
                        int v = *o;
                        if (c < v) {
                            v = c;
                            *o = c;
                        }
                        mx = hd_max(mx, v);
                    }
                }
            }
            if (act) cont = w2.x == FULLW && meta_bot_open(w2.y);
        }
    }
}

}  // namespace edtband
}  // namespace sln
