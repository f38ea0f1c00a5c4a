/*
 * sln_oracle.c -- CPU restatement of the SLN-Amodal detection-head hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the CUDA
 * library in sln_amodal_b200/csrc.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * path never links, imports or calls anything under oracle/.
 *
 * Each function restates the arithmetic of one reference routine (file:line
 * under /root/reference given at each function) with plain pointers instead of
 * TH tensors.  The restatement is pinned by tests/test_oracle_pin.py, which
 * compares it bit-for-bit with the reference's own unmodified C sources
 * compiled into oracle/_ref/ (see oracle/Makefile), and with scipy for the EDT
 * (the reference has no EDT: "parity unpinned" for that one op).
 *
 * Build: gcc -O2 -std=c99 -ffp-contract=off (no -march, no fast-math): fp32
 * operations stay un-fused, which is what defines the reference bit pattern.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* One axis of the crop sampling grid.
 * Follows crop_and_resize.c:44-56 (scale, sample position), :58 / :80 (range
 * test), :71-73 / :89-91 (floorf/ceilf taps and lerp weight). */
typedef struct {
    int valid; /* sample inside [0, extent-1] */
    int lo;    /* floorf(pos) */
    int hi;    /* ceilf(pos)  */
    float lerp;
} orc_tap;

static void orc_axis(float a1, float a2, int extent, int crop, orc_tap *taps)
{
    const float scale = (crop > 1) ? (a2 - a1) * (extent - 1) / (crop - 1) : 0;
    for (int k = 0; k < crop; ++k) {
        const float pos = (crop > 1) ? a1 * (extent - 1) + k * scale
                                     : 0.5 * (a1 + a2) * (extent - 1);
        orc_tap t;
        t.valid = !(pos < 0 || pos > extent - 1);
        t.lo = t.hi = 0;
        t.lerp = 0.f;
        if (t.valid) {
            t.lo = (int)floorf(pos);
            t.hi = (int)ceilf(pos);
            t.lerp = pos - t.lo;
        }
        taps[k] = t;
    }
}

/* crop_and_resize forward, NCHW in / NCHW out.
 * Reference: CropAndResizePerBox + crop_and_resize_forward,
 * roialign/roi_align/src/crop_and_resize.c:6-112, 115-154.
 * Returns 0, or -1 if a box_ind is out of range (the reference exit(-1)s at
 * crop_and_resize.c:39-42). */
int orc_crop_and_resize_fwd(const float *image, int B, int C, int H, int W,
                            const float *boxes, const int *box_ind, int N,
                            int ph, int pw, float ext, float *crops)
{
    orc_tap *ty = (orc_tap *)malloc(sizeof(orc_tap) * (size_t)(ph + pw));
    orc_tap *tx = ty + ph;
    const size_t plane = (size_t)H * W;
    for (int r = 0; r < N; ++r) {
        const int b = box_ind[r];
        if (b < 0 || b >= B) { free(ty); return -1; }
        orc_axis(boxes[4 * r + 0], boxes[4 * r + 2], H, ph, ty);
        orc_axis(boxes[4 * r + 1], boxes[4 * r + 3], W, pw, tx);
        for (int c = 0; c < C; ++c) {
            const float *src = image + ((size_t)b * C + c) * plane;
            float *dst = crops + ((size_t)r * C + c) * ph * pw;
            for (int y = 0; y < ph; ++y) {
                for (int x = 0; x < pw; ++x) {
                    float v = ext;
                    if (ty[y].valid && tx[x].valid) {
                        const float tl = src[(size_t)ty[y].lo * W + tx[x].lo];
                        const float tr = src[(size_t)ty[y].lo * W + tx[x].hi];
                        const float bl = src[(size_t)ty[y].hi * W + tx[x].lo];
                        const float br = src[(size_t)ty[y].hi * W + tx[x].hi];
                        const float top = tl + (tr - tl) * tx[x].lerp;
                        const float bot = bl + (br - bl) * tx[x].lerp;
                        v = top + (bot - top) * ty[y].lerp;
                    }
                    dst[y * pw + x] = v;
                }
            }
        }
    }
    free(ty);
    return 0;
}

/* crop_and_resize backward w.r.t. the image, NCHW.
 * Reference: crop_and_resize_backward, crop_and_resize.c:157-252.  The
 * accumulation order per destination pixel is (box, y, x, tap TL/TR/BL/BR),
 * exactly the reference's serial loop order (:190-250), so the fp32 sums are
 * reproducible bit for bit. */
int orc_crop_and_resize_bwd(const float *grads, const float *boxes,
                            const int *box_ind, int N, int C, int ph, int pw,
                            float *grad_image, int B, int H, int W)
{
    orc_tap *ty = (orc_tap *)malloc(sizeof(orc_tap) * (size_t)(ph + pw));
    orc_tap *tx = ty + ph;
    const size_t plane = (size_t)H * W;
    memset(grad_image, 0, sizeof(float) * (size_t)B * C * plane);
    for (int r = 0; r < N; ++r) {
        const int b = box_ind[r];
        if (b < 0 || b >= B) { free(ty); return -1; }
        orc_axis(boxes[4 * r + 0], boxes[4 * r + 2], H, ph, ty);
        orc_axis(boxes[4 * r + 1], boxes[4 * r + 3], W, pw, tx);
        for (int y = 0; y < ph; ++y) {
            if (!ty[y].valid) continue;
            const float yl = ty[y].lerp;
            for (int x = 0; x < pw; ++x) {
                if (!tx[x].valid) continue;
                const float xl = tx[x].lerp;
                for (int c = 0; c < C; ++c) {
                    float *dst = grad_image + ((size_t)b * C + c) * plane;
                    const float g = grads[(((size_t)r * C + c) * ph + y) * pw + x];
                    const float dtop = (1 - yl) * g;
                    dst[(size_t)ty[y].lo * W + tx[x].lo] += (1 - xl) * dtop;
                    dst[(size_t)ty[y].lo * W + tx[x].hi] += xl * dtop;
                    const float dbot = yl * g;
                    dst[(size_t)ty[y].hi * W + tx[x].lo] += (1 - xl) * dbot;
                    dst[(size_t)ty[y].hi * W + tx[x].hi] += xl * dbot;
                }
            }
        }
    }
    free(ty);
    return 0;
}

/* Greedy NMS over a caller-supplied visiting order and precomputed areas.
 * Reference: cpu_nms, nms/src/nms.c:4-69.  "+1" pixel convention (:55-56),
 * un-fused fp32, IEEE divide, suppress when ovr >= thresh (:58-61).  Kept
 * original indices are written in visiting order; *num_out gets the count. */
int orc_nms(const float *boxes, int stride, const int64_t *order,
            const float *areas, int64_t n, float thresh, int64_t *keep,
            int64_t *num_out)
{
    unsigned char *dead = (unsigned char *)calloc((size_t)(n > 0 ? n : 1), 1);
    int64_t kept = 0;
    for (int64_t pi = 0; pi < n; ++pi) {
        const int64_t i = order[pi];
        if (dead[i]) continue;
        keep[kept++] = i;
        const float *bi = boxes + i * stride;
        const float ia = areas[i];
        for (int64_t pj = pi + 1; pj < n; ++pj) {
            const int64_t j = order[pj];
            if (dead[j]) continue;
            const float *bj = boxes + j * stride;
            const float xx1 = fmaxf(bi[0], bj[0]);
            const float yy1 = fmaxf(bi[1], bj[1]);
            const float xx2 = fminf(bi[2], bj[2]);
            const float yy2 = fminf(bi[3], bj[3]);
            const float w = fmaxf(0.0, xx2 - xx1 + 1);
            const float h = fmaxf(0.0, yy2 - yy1 + 1);
            const float inter = w * h;
            const float ovr = inter / (ia + areas[j] - inter);
            if (ovr >= thresh) dead[j] = 1;
        }
    }
    *num_out = kept;
    free(dead);
    return 0;
}

/* Exact squared Euclidean distance transform of one binary map: for every
 * pixel, the squared distance to the nearest ZERO pixel (0 on zero pixels).
 * Not in the reference; semantics pinned to
 * scipy.ndimage.distance_transform_edt(map)**2 by tests/test_oracle_pin.py.
 * A map with no zero pixel yields `big` = (H+W)^2 everywhere (documented
 * saturation; scipy is undefined there).
 * Method: column pass (distance to nearest zero along y), then for each pixel
 * an outward search along x that stops once d*d >= best.  All integer. */
void orc_edt_sq(const unsigned char *map, int H, int W, int32_t *out)
{
    const int32_t inf = H + W;
    int32_t *g = (int32_t *)malloc(sizeof(int32_t) * (size_t)H * W);
    for (int x = 0; x < W; ++x) {
        int32_t d = inf;
        for (int y = 0; y < H; ++y) {
            d = map[(size_t)y * W + x] ? (d < inf ? d + 1 : inf) : 0;
            g[(size_t)y * W + x] = d;
        }
        d = inf;
        for (int y = H - 1; y >= 0; --y) {
            d = map[(size_t)y * W + x] ? (d < inf ? d + 1 : inf) : 0;
            if (d < g[(size_t)y * W + x]) g[(size_t)y * W + x] = d;
        }
    }
    for (int y = 0; y < H; ++y) {
        const int32_t *gr = g + (size_t)y * W;
        for (int x = 0; x < W; ++x) {
            int64_t best = (int64_t)gr[x] * gr[x];
            for (int d = 1; (int64_t)d * d < best; ++d) {
                if (x - d < 0 && x + d >= W) break;
                if (x - d >= 0) {
                    const int64_t v = (int64_t)d * d + (int64_t)gr[x - d] * gr[x - d];
                    if (v < best) best = v;
                }
                if (x + d < W) {
                    const int64_t v = (int64_t)d * d + (int64_t)gr[x + d] * gr[x + d];
                    if (v < best) best = v;
                }
            }
            const int64_t cap = (int64_t)inf * inf;
            out[(size_t)y * W + x] = (int32_t)(best < cap ? best : cap);
        }
    }
    free(g);
}

/* Second, independent EDT restatement -- and the executable form of the banded column pass planned for the CUDA
 * kernel: row pass first (g = distance to the nearest zero along x), then per column one lower
 * envelope PER BAND of `band` rows over the band's rows plus the row just above and just below it (site s has height
 * g(s)^2; rows without a zero are no sites), and every pixel takes the minimum over the envelopes of the bands in
 * reach: its own, then outwards while the row gap squared is below the best value so far.  Any superset of bands is
 * exact (every candidate is the true squared distance to some zero pixel), so the pruning only saves work.
 * tests/test_oracle_pin.py holds it to orc_edt_sq and scipy for band heights 1 .. H. */
static int64_t ceil_div64(int64_t a, int64_t b)      /* b > 0 */
{
    int64_t q = a / b;
    if ((a % b != 0) && (a > 0)) ++q;
    return q;
}

void orc_edt_sq_banded(const unsigned char *map, int H, int W, int band, int32_t *out)
{
    const int32_t inf = H + W;
    const int64_t cap = (int64_t)inf * inf;
    if (band < 1) band = 1;
    const int nb = (H + band - 1) / band;
    int32_t *g = (int32_t *)malloc(sizeof(int32_t) * (size_t)H * W);
    for (int y = 0; y < H; ++y) {                     /* row pass */
        const unsigned char *m = map + (size_t)y * W;
        int32_t *gr = g + (size_t)y * W;
        int32_t d = inf;
        for (int x = 0; x < W; ++x) { d = m[x] ? (d < inf ? d + 1 : inf) : 0; gr[x] = d; }
        d = inf;
        for (int x = W - 1; x >= 0; --x) { d = m[x] ? (d < inf ? d + 1 : inf) : 0; if (d < gr[x]) gr[x] = d; }
    }
    /* per band: site rows v[], first winning row z[] (z[0] = -inf), count */
    int *v = (int *)malloc(sizeof(int) * (size_t)nb * (band + 2));
    int64_t *z = (int64_t *)malloc(sizeof(int64_t) * (size_t)nb * (band + 2));
    int *cnt = (int *)malloc(sizeof(int) * (size_t)nb);
    for (int x = 0; x < W; ++x) {
        for (int b = 0; b < nb; ++b) {                /* phase A: band-local envelopes */
            int *vb = v + (size_t)b * (band + 2);
            int64_t *zb = z + (size_t)b * (band + 2);
            int k = -1;
            const int s0 = b * band - 1 < 0 ? 0 : b * band - 1;
            const int s1 = (b + 1) * band > H - 1 ? H - 1 : (b + 1) * band;      /* inclusive */
            for (int s = s0; s <= s1; ++s) {
                const int64_t gs = g[(size_t)s * W + x];
                if (gs >= inf) continue;
                const int64_t fs = gs * gs + (int64_t)s * s;
                int64_t start = INT64_MIN;
                while (k >= 0) {
                    const int64_t gv = g[(size_t)vb[k] * W + x];
                    const int64_t fv = gv * gv + (int64_t)vb[k] * vb[k];
                    start = ceil_div64(fs - fv, 2 * (int64_t)(s - vb[k]));       /* first row where s is at least as good */
                    if (start <= zb[k]) --k; else break;
                }
                if (k < 0) start = INT64_MIN;
                ++k;
                vb[k] = s;
                zb[k] = start;
            }
            cnt[b] = k + 1;
        }
        for (int y = 0; y < H; ++y) {                 /* phase B: minimum over the bands in reach */
            const int b = y / band;
            int64_t best = INT64_MAX;
            for (int d = 0; d < nb; ++d) {
                int any = 0;
                for (int side = 0; side < (d ? 2 : 1); ++side) {
                    const int bb = side ? b + d : b - d;
                    if (bb < 0 || bb >= nb) continue;
                    int64_t gap = 0;                  /* rows between y and the nearest site row of band bb */
                    if (bb < b) { const int last = (bb + 1) * band > H - 1 ? H - 1 : (bb + 1) * band; gap = y > last ? y - last : 0; }
                    if (bb > b) { const int first = bb * band - 1; gap = first > y ? first - y : 0; }
                    if (best != INT64_MAX && gap * gap >= best) continue;
                    any = 1;
                    const int n = cnt[bb];
                    if (n == 0) continue;
                    const int *vb = v + (size_t)bb * (band + 2);
                    const int64_t *zb = z + (size_t)bb * (band + 2);
                    int lo = 0, hi = n - 1;           /* largest k with z[k] <= y */
                    while (lo < hi) { const int mid = (lo + hi + 1) / 2; if (zb[mid] <= y) lo = mid; else hi = mid - 1; }
                    const int64_t gv = g[(size_t)vb[lo] * W + x];
                    const int64_t val = gv * gv + (int64_t)(y - vb[lo]) * (y - vb[lo]);
                    if (val < best) best = val;
                }
                if (!any && d > 0) break;
            }
            out[(size_t)y * W + x] = (int32_t)(best < cap ? best : cap);
        }
    }
    free(cnt); free(z); free(v); free(g);
}

/* Layer ("sem-dist") target decode in closed form.
 * Reference: amodal_train.py:236-271 (load_layer2) driving
 * modal/Functions.py:1012-1095 (get_image_labals, objectID_to_masks,
 * max_objectID, maskID_to_objectIDs, objIDs_to_sindistanceLayer).
 *   label bit i (i<32)   : object i visible at this pixel
 *   label bit 32+i       : object i present but occluded at this pixel
 *   n_obj = first shift s such that no non-zero label has its low word's top
 *           set bit at s (max_objectID, Functions.py:1074-1079)
 *   channel 0 of object i      |= bit i
 *   channel min(d, L-1)        |= bit 32+i,  d = 1 + #set bits of the high
 *                                 word below position i (Functions.py:1050-1064,
 *                                 amodal_train.py:253-259)
 * out is u8 [n_max][L][H][W]; planes for i >= n_obj are zero.  Returns n_obj
 * (not clamped to n_max).  The python restatement that loops exactly like the
 * reference lives in oracle/oracle.py; tests pin this closed form to it. */
int orc_layer_decode(const uint64_t *label, int H, int W, int L, int n_max,
                     unsigned char *out)
{
    const size_t px = (size_t)H * W;
    uint32_t top_seen = 0;
    for (size_t p = 0; p < px; ++p) {
        const uint32_t lo = (uint32_t)(label[p] & 0xffffffffu);
        if (lo) top_seen |= 1u << (31 - __builtin_clz(lo));
    }
    int n_obj = 0;
    while (n_obj < 32 && ((top_seen >> n_obj) & 1u)) ++n_obj;
    memset(out, 0, (size_t)n_max * L * px);
    const int n_eff = n_obj < n_max ? n_obj : n_max;
    for (size_t p = 0; p < px; ++p) {
        const uint64_t v = label[p];
        if (!v) continue;
        const uint32_t lo = (uint32_t)(v & 0xffffffffu);
        const uint32_t hi = (uint32_t)(v >> 32);
        for (int i = 0; i < n_eff; ++i) {
            if ((lo >> i) & 1u) out[((size_t)i * L + 0) * px + p] = 1;
            if ((hi >> i) & 1u) {
                int d = 1 + __builtin_popcount(hi & ((1u << i) - 1u));
                if (d > L - 1) d = L - 1;
                out[((size_t)i * L + d) * px + p] = 1;
            }
        }
    }
    return n_obj;
}

/* ---------------------------------------------------------------------------
 * Host scan over the 64x64-tile IoU bit matrix of the reference's GPU NMS: restatement of nms_cuda.c:33-58 (which
 * needs THC and cannot be compiled here).  mask [n][ceil(n/64)] u64 as written by nms_kernel.cu:26-70; keep receives
 * the positions (in the sorted order the kernel was given) of the survivors; returns their count.  Used by the
 * speed-only comparator of the reference's CUDA path (bench.py extra.ref_cuda).
 * ------------------------------------------------------------------------- */
long orc_nms_mask_scan(const unsigned long long *mask, int n, long long *keep)
{
    const int col_blocks = (n + 63) / 64;
    unsigned long long *remv = (unsigned long long *)calloc((size_t)col_blocks, sizeof(unsigned long long));
    long num = 0;
    for (int i = 0; i < n; i++) {
        const int nblock = i / 64, inblock = i % 64;
        if (!(remv[nblock] & (1ULL << inblock))) {
            keep[num++] = i;
            const unsigned long long *p = mask + (size_t)i * col_blocks;
            for (int j = nblock; j < col_blocks; j++) remv[j] |= p[j];
        }
    }
    free(remv);
    return num;
}
