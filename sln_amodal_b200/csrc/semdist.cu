// semdist.cu -- semantics-aware distance-map target encoding: the bit-packed layer
// codec and an exact squared Euclidean distance transform.
//
// Layer codec reference: AmodalDataset.load_layer2 (amodal_train.py:236-271) driving
// get_image_labals / objectID_to_masks / max_objectID / maskID_to_objectIDs /
// objIDs_to_sindistanceLayer (modal/Functions.py:1012-1095).  The numpy reference does
// one full-image compare per (object x label piece) after an np.unique over 1 M u64;
// here every pixel is decoded independently with bit operations:
//     visible plane of object i    <- bit i
//     plane min(d, L-1) of object i <- bit 32+i,  d = 1 + popcount(hi & ((1<<i)-1))
//     n_obj = first s such that no label's low word has its top set bit at s
//
// EDT: NOT in the reference.  Separable and exact on integers:
//     pass 1 (rows)    g(y,x)  = |x - nearest zero in row y|         warp-parallel scans
//     pass 2 (columns) D(y,x)  = min_y' (y-y')^2 + g(y',x)^2          bounded outward search:
//                      start from g(y,x)^2 and walk d = 1,2,.. while d^2 < best.
// The search cost of a pixel equals its true distance, so background pixels cost
// nothing and the whole pass is coalesced along x.  The intermediate g is u16 and is
// produced/consumed map-chunk by map-chunk so that it lives in L2, not HBM.
#include <stdlib.h>

#include "common.cuh"
#include "edt_band.cuh"

namespace sln {

// ---------------------------------------------------------------------------
// layer codec
// ---------------------------------------------------------------------------
// One thread decodes 8 consecutive pixels for every object plane.  Per object it gathers
// bit i of the 8 labels into an 8-bit mask and expands it to 8 bytes with a 256-entry
// shared-memory table (one 8-byte store per plane, 256 B per warp).  The occluded bits
// (high word) are only looked at when one of the 8 pixels has that bit set.
// The same pass ORs (a) the top visible bit of every label (max_objectID rule) and
// (b) every label bit below n_max into per-image words; layer_fixup_kernel then derives
// n_obj and, only if an object >= n_obj had pixels, clears those planes (the reference
// drops objects after the first never-visible one, Functions.py:1074-1079).
// LD_PX pixels per thread: 16 when the plane size allows (one 16-byte store per plane and thread: the kernel is bound by
// its instruction count, 20 stores + loop per thread, so twice the pixels per thread is ~40 % fewer warp instructions), else 8

__device__ __forceinline__ unsigned long long expand_bits(const unsigned long long *__restrict__ lut, unsigned m)
{
    return lut[m & 0xffu];
}

template <int LD_PX>
__global__ void __launch_bounds__(256)
layer_decode_kernel(const unsigned long long *__restrict__ label, size_t px_per_image, int L, int n_max,
                    unsigned char *__restrict__ out, unsigned *__restrict__ seen /* [B][2]: top bits, any bits */)
{
    static_assert(LD_PX == 8 || LD_PX == 16, "8 or 16 pixels per thread");
    pdl_prologue();
    __shared__ unsigned long long s_lut[256];
    {
        unsigned long long v = 0ull;
        for (int k = 0; k < 8; ++k) v |= (unsigned long long)((threadIdx.x >> k) & 1u) << (8 * k);
        s_lut[threadIdx.x] = v;
    }
    __syncthreads();
    const int b = blockIdx.y;
    const size_t p0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * LD_PX;
    unsigned top = 0u, any = 0u;
    if (p0 < px_per_image) {
        const unsigned long long *lab = label + (size_t)b * px_per_image + p0;
        unsigned char *o = out + (size_t)b * n_max * L * px_per_image + p0;
        const bool full = (p0 + LD_PX <= px_per_image) && ((px_per_image % LD_PX) == 0);
        unsigned lo[LD_PX], hi[LD_PX];
        if (full) {
            const ulonglong2 *l2 = reinterpret_cast<const ulonglong2 *>(lab);
#pragma unroll
            for (int k = 0; k < LD_PX / 2; ++k) {
                const ulonglong2 v = __ldg(l2 + k);
                lo[2 * k] = (unsigned)v.x; hi[2 * k] = (unsigned)(v.x >> 32);
                lo[2 * k + 1] = (unsigned)v.y; hi[2 * k + 1] = (unsigned)(v.y >> 32);
            }
        } else {
#pragma unroll
            for (int k = 0; k < LD_PX; ++k) {
                unsigned long long v = 0ull;
                if (p0 + k < px_per_image) v = __ldg(lab + k);
                lo[k] = (unsigned)v; hi[k] = (unsigned)(v >> 32);
            }
        }
        unsigned lo_or = 0u, hi_or = 0u;
#pragma unroll
        for (int k = 0; k < LD_PX; ++k) {
            lo_or |= lo[k];
            hi_or |= hi[k];
            if (lo[k]) top |= 1u << (31 - __clz(lo[k]));
        }
        any = lo_or | hi_or;
        unsigned char *dst = o;                     // walks the planes (i major, l minor)
        if (L == 1) {
            // one plane per object: visible and occluded pixels share channel 0 (amodal_train.py:256-257)
            unsigned lh[LD_PX];
#pragma unroll
            for (int k = 0; k < LD_PX; ++k) lh[k] = lo[k] | hi[k];
            for (int i = 0; i < n_max; ++i, dst += px_per_image) {
                unsigned m = 0u;
                if ((any >> i) & 1u) {
#pragma unroll
                    for (int k = 0; k < LD_PX; ++k) m |= ((lh[k] >> i) & 1u) << k;
                }
                const unsigned long long bytes = expand_bits(s_lut, m);
                if (full) {
                    if (LD_PX == 16) {
                        const unsigned long long b2 = expand_bits(s_lut, m >> 8);
                        __stcs(reinterpret_cast<uint4 *>(dst), make_uint4((unsigned)bytes, (unsigned)(bytes >> 32), (unsigned)b2, (unsigned)(b2 >> 32)));
                    } else {
                        __stcs(reinterpret_cast<unsigned long long *>(dst), bytes);
                    }
                } else {
                    for (int k = 0; k < LD_PX && p0 + k < px_per_image; ++k)
                        dst[k] = (unsigned char)((k < 8 ? bytes : expand_bits(s_lut, m >> 8)) >> (8 * (k & 7)));
                }
            }
        } else {
            for (int i = 0; i < n_max; ++i) {
                unsigned mv = 0u;                   // visible mask of object i over the 8 pixels
                if ((lo_or >> i) & 1u) {
#pragma unroll
                    for (int k = 0; k < LD_PX; ++k) mv |= ((lo[k] >> i) & 1u) << k;
                }
                const bool occ = (hi_or >> i) & 1u;
                for (int l = 0; l < L; ++l, dst += px_per_image) {
                    unsigned m = l == 0 ? mv : 0u;
                    if (occ) {
#pragma unroll
                        for (int k = 0; k < LD_PX; ++k) {
                            if ((hi[k] >> i) & 1u) {
                                const int d = min(1 + __popc(hi[k] & ((1u << i) - 1u)), L - 1);
                                m |= (unsigned)(d == l) << k;
                            }
                        }
                    }
                    const unsigned long long bytes = expand_bits(s_lut, m);
                    if (full) {
                        if (LD_PX == 16) {
                            const unsigned long long b2 = expand_bits(s_lut, m >> 8);
                            __stcs(reinterpret_cast<uint4 *>(dst), make_uint4((unsigned)bytes, (unsigned)(bytes >> 32), (unsigned)b2, (unsigned)(b2 >> 32)));
                        } else {
                            __stcs(reinterpret_cast<unsigned long long *>(dst), bytes);
                        }
                    } else {
                        for (int k = 0; k < LD_PX && p0 + k < px_per_image; ++k)
                            dst[k] = (unsigned char)((k < 8 ? bytes : expand_bits(s_lut, m >> 8)) >> (8 * (k & 7)));
                    }
                }
            }
        }
    }
    top = __reduce_or_sync(0xffffffffu, top);
    any = __reduce_or_sync(0xffffffffu, any);
    if ((threadIdx.x & 31) == 0) {                  // OR: order-independent
        if (top) atomicOr(seen + 2 * b, top);
        if (any) atomicOr(seen + 2 * b + 1, any);
    }
}

// n_obj per image; planes of objects in [n_obj, n_max) are cleared when any of them got pixels
__global__ void __launch_bounds__(256)
layer_fixup_kernel(const unsigned *__restrict__ seen, size_t px_per_image, int L, int n_max,
                   unsigned char *__restrict__ out, int *__restrict__ n_obj_out)
{
    pdl_prologue();
    const int b = blockIdx.y;
    const unsigned top = seen[2 * b], any = seen[2 * b + 1];
    const int n_obj = __ffs(~top) == 0 ? 32 : __ffs(~top) - 1;     // max_objectID, Functions.py:1074-1079
    if (blockIdx.x == 0 && threadIdx.x == 0) n_obj_out[b] = n_obj;
    if (n_obj >= n_max) return;
    const unsigned upper = n_max >= 32 ? 0xffffffffu : ((1u << n_max) - 1u);
    if (((any & upper) >> n_obj) == 0u) return;                     // nothing was written there
    unsigned char *o = out + ((size_t)b * n_max + n_obj) * L * px_per_image;
    const size_t bytes = (size_t)(n_max - n_obj) * L * px_per_image;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
    if ((reinterpret_cast<uintptr_t>(o) & 15u) == 0) {
        const size_t n16 = bytes >> 4;
        const uint4 zz = make_uint4(0u, 0u, 0u, 0u);
        for (size_t i = tid; i < n16; i += nthr) reinterpret_cast<uint4 *>(o)[i] = zz;
        for (size_t i = (n16 << 4) + tid; i < bytes; i += nthr) o[i] = 0;
    } else {
        for (size_t i = tid; i < bytes; i += nthr) o[i] = 0;
    }
}

// ---------------------------------------------------------------------------
// EDT pass 1: distance to the nearest zero along each row (u16, 0xffff = none)
// ---------------------------------------------------------------------------
constexpr unsigned short G_INF = 0xffffu;
constexpr int ROW_SEG = 32;                 // pixels per lane per chunk
constexpr int ROW_CHUNK = 32 * ROW_SEG;     // pixels per warp per chunk

// zero-mask of the lane's 32-pixel segment starting at x0 (bit k <=> pixel x0+k == 0)
__device__ __forceinline__ unsigned seg_zero_mask(const unsigned char *__restrict__ src, int x0, int W, bool vec_ok)
{
    unsigned z = 0u;
    if (x0 >= W) return 0u;
    if (vec_ok && x0 + ROW_SEG <= W) {
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(src + x0));
        const uint4 b = __ldg(reinterpret_cast<const uint4 *>(src + x0 + 16));
        if ((a.x | a.y | a.z | a.w | b.x | b.y | b.z | b.w) == 0u) return 0xffffffffu;     // background
        const unsigned w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            // 0xff per zero byte -> one bit per byte -> the four bits gathered at 24..27 by one multiply
            const unsigned eq = __vcmpeq4(w[q], 0u) & 0x01010101u;
            z |= ((eq * 0x01020408u) >> 24) << (4 * q);
        }
    } else {
        for (int k = 0; k < ROW_SEG && x0 + k < W; ++k) z |= (unsigned)(__ldg(src + x0 + k) == 0) << k;
    }
    return z;
}

// One warp per row.  Lane l owns 32 consecutive pixels of each 1024-pixel chunk; the
// nearest zero outside the segment comes from warp scans of the per-lane last / first
// zero positions, plus a carry between chunks (second, backward sweep only for W > 1024).
template <typename GT>      // u16 for the search / generic envelope kernels, u32 for the packed envelope kernel
__global__ void __launch_bounds__(256)
edt_rows_kernel(const unsigned char *__restrict__ maps, int H, int W, long long n_rows, GT *__restrict__ g,
                unsigned *__restrict__ flags, int tiles_y, int tiles_x, bool skip_zero_segments)
{
    constexpr bool WIDE = sizeof(GT) == 4;
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const unsigned char *src = maps + row * W;
    GT *dst = g + row * W;
    const bool vec_ok = (W % 16) == 0 && (((reinterpret_cast<uintptr_t>(maps) | reinterpret_cast<uintptr_t>(g)) & 15u) == 0);
    const int n_chunks = (W + ROW_CHUNK - 1) / ROW_CHUNK;
    constexpr int NONE_R = 1 << 29;
    int carry_left = -NONE_R;                    // last zero in earlier chunks
    for (int c = 0; c < n_chunks; ++c) {
        const int x0 = c * ROW_CHUNK + lane * ROW_SEG;
        const unsigned z = seg_zero_mask(src, x0, W, vec_ok);
        if (skip_zero_segments && (c + 1) * ROW_CHUNK <= W && __all_sync(0xffffffffu, z == 0xffffffffu)) {
            carry_left = (c + 1) * ROW_CHUNK - 1;    // 1024 background pixels: no flag bit, nothing to write
            continue;
        }
        const int my_last = z ? x0 + 31 - __clz(z) : -NONE_R;
        const int my_first = z ? x0 + __ffs(z) - 1 : NONE_R;
        int left = my_last;                      // inclusive max-scan, then shift
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, left, o);
            if (lane >= o) left = max(left, v);
        }
        const int chunk_last = __shfl_sync(0xffffffffu, left, 31);
        left = __shfl_up_sync(0xffffffffu, left, 1);
        if (lane == 0) left = -NONE_R;
        left = max(left, carry_left);
        int right = my_first;
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_down_sync(0xffffffffu, right, o);
            if (lane + o < 32) right = min(right, v);
        }
        right = __shfl_down_sync(0xffffffffu, right, 1);
        if (lane == 31) right = NONE_R;
        carry_left = max(carry_left, chunk_last);
        // tile flags for the column pass: bit (y & 31) of flags[map][y / 32][x / 32] <=> that 32-pixel segment of
        // row y holds a non-zero pixel (a segment cut by the right border counts as non-empty: harmless)
        if (x0 < W && z != 0xffffffffu) {
            const long long m = row / H;
            const int y = (int)(row - m * H);
            atomicOr(flags + ((size_t)m * tiles_y + (y >> 5)) * tiles_x + (x0 >> 5), 1u << (y & 31));
        }
        if (z == 0xffffffffu && x0 + ROW_SEG <= W && skip_zero_segments) {
            // every pixel of the segment is a zero pixel (the common case: instance masks are mostly
            // background).  Its flag bit stays clear and the envelope kernels never read g there, so
            // nothing is written at all.
        } else if (z == 0xffffffffu && vec_ok && x0 + ROW_SEG <= W) {
            uint4 *o = reinterpret_cast<uint4 *>(dst + x0);
            const uint4 zz = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
            for (int q = 0; q < (WIDE ? 8 : 4); ++q) o[q] = zz;
        } else if (x0 < W) {
            // distance of pixel k = min(k - nearest zero bit at or below k, nearest zero bit at or above k - k),
            // falling back to the zeros left / right of the segment; 8 pixels per step, 16-byte stores
            const bool vst = vec_ok && x0 + ROW_SEG <= W;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                unsigned dv[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int k = 8 * q + e;
                    const unsigned below = z & (0xffffffffu >> (31 - k));      // zero pixels at positions <= k
                    const unsigned above = z >> k;                             // zero pixels at positions >= k
                    const int lz = below ? x0 + 31 - __clz(below) : left;
                    const int rz = above ? x0 + k + __ffs(above) - 1 : right;
                    const int dl = min(x0 + k - lz, (int)G_INF), dr = min(rz - (x0 + k), (int)G_INF);
                    dv[e] = (unsigned)min(dl, dr);
                    if (!vst && x0 + k < W) dst[x0 + k] = (GT)dv[e];
                }
                if (vst) {
                    if (WIDE) {
                        reinterpret_cast<uint4 *>(dst + x0)[2 * q] = make_uint4(dv[0], dv[1], dv[2], dv[3]);
                        reinterpret_cast<uint4 *>(dst + x0)[2 * q + 1] = make_uint4(dv[4], dv[5], dv[6], dv[7]);
                    } else {
                        reinterpret_cast<uint4 *>(dst + x0)[q] = make_uint4(dv[0] | (dv[1] << 16), dv[2] | (dv[3] << 16),
                                                                            dv[4] | (dv[5] << 16), dv[6] | (dv[7] << 16));
                    }
                }
            }
        }
    }
    if (n_chunks > 1) {                          // zeros in later chunks
        __syncwarp();
        int carry_right = NONE_R;
        for (int c = n_chunks - 1; c >= 0; --c) {
            const int x0 = c * ROW_CHUNK + lane * ROW_SEG;
            const unsigned z = seg_zero_mask(src, x0, W, vec_ok);
            if (carry_right != NONE_R && !(skip_zero_segments && z == 0xffffffffu && x0 + ROW_SEG <= W)) {
                for (int k = 0; k < ROW_SEG && x0 + k < W; ++k) {
                    const int dr = min(carry_right - (x0 + k), (int)G_INF);
                    if (dr < (int)dst[x0 + k]) dst[x0 + k] = (GT)dr;
                }
            }
            int first = z ? x0 + __ffs(z) - 1 : NONE_R;
            for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
            carry_right = min(carry_right, first);
        }
    }
}

// ---------------------------------------------------------------------------
// EDT pass 2: column search.  One thread per 4 horizontally adjacent pixels
// (8-byte g loads, 16-byte stores); a warp covers 128 pixels of one row.
// ---------------------------------------------------------------------------
__device__ __forceinline__ int sq_or_cap(unsigned gv, int cap)
{
    return gv == G_INF ? cap : (int)min(gv * gv, (unsigned)cap);
}

template <bool VEC>
__global__ void __launch_bounds__(256)
edt_cols_kernel(const unsigned short *__restrict__ g, int H, int W, int cap, int *__restrict__ out)
{
    const int x = (blockIdx.x * 32 + (threadIdx.x & 31)) * 4;
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int m = blockIdx.z;
    if (x >= W || y >= H) return;
    const unsigned short *gm = g + (size_t)m * H * W + x;
    int *om = out + (size_t)m * H * W + (size_t)y * W + x;
    int best[4];
    int gmax = 0;                      // search radius bound: max over the 4 pixels of best
    auto load4 = [&](int yy, unsigned v[4]) {
        const unsigned short *p = gm + (size_t)yy * W;
        if (VEC) {
            const uint2 t = __ldg(reinterpret_cast<const uint2 *>(p));
            v[0] = t.x & 0xffffu; v[1] = t.x >> 16; v[2] = t.y & 0xffffu; v[3] = t.y >> 16;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = (x + k < W) ? (unsigned)__ldg(p + k) : 0u;
        }
    };
    unsigned v0[4];
    load4(y, v0);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        best[k] = sq_or_cap(v0[k], cap);
        gmax = max(gmax, best[k]);
    }
    for (int d = 1; d * d < gmax; ++d) {
        const bool up = y - d >= 0, dn = y + d < H;
        if (!up && !dn) break;
        const int dd = d * d;
        if (up) {
            unsigned v[4];
            load4(y - d, v);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (v[k] != G_INF) best[k] = min(best[k], (int)min((long long)dd + (long long)v[k] * v[k], (long long)cap));
        }
        if (dn) {
            unsigned v[4];
            load4(y + d, v);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (v[k] != G_INF) best[k] = min(best[k], (int)min((long long)dd + (long long)v[k] * v[k], (long long)cap));
        }
        gmax = max(max(best[0], best[1]), max(best[2], best[3]));
    }
    if (VEC) {
        __stcs(reinterpret_cast<int4 *>(om), make_int4(best[0], best[1], best[2], best[3]));
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (x + k < W) om[k] = best[k];
    }
}

// Pass 2, exact lower-envelope scan (Meijster et al., "A general algorithm for computing
// distance transforms in linear time", phase 2) -- cost per pixel independent of the blob
// size, unlike the outward search of edt_cols_kernel (kept as the fallback for W < 32).
// thread = one column; the warp walks the rows in lock step, so g reads and output writes
// are coalesced and rows whose 32 pixels are all background cost a single vote.
// Per column, every maximal run of foreground pixels is one envelope problem over the
// parabolas f_s(y) = (y-s)^2 + g(s)^2 of the run's pixels plus the zero pixels that bound
// it.  The envelope stack (s, first row t where s wins) is kept IN the output column
// itself: entry k lives in out[k][x]; t is strictly increasing so k <= t[k], and the
// backward sweep at row y only needs entries with t <= y, i.e. k <= y -- it never
// overwrites an entry it still has to read.
__device__ __forceinline__ int floor_div(int a, int b)      // b > 0
{
    int q = a / b;
    if ((a % b != 0) && (a < 0)) --q;
    return q;
}

// zero a 32-row x 32-column tile of the output with 16-byte stores (lane = 4 columns of one of 4 rows).
// All eight addresses are formed before the first store: a store holds its address registers until the
// LSU has taken it, and reusing one pair for the next address serialises the stores on that release.
__device__ __forceinline__ void zero_tile(int *tile_base, int W, int rows, int lane)
{
    const int c4 = (lane & 7) * 4, r0 = lane >> 3;
    int4 *p[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) p[k] = reinterpret_cast<int4 *>(tile_base + (size_t)(r0 + 4 * k) * W + c4);
    if (rows == 32) {
#pragma unroll
        for (int k = 0; k < 8; ++k) asm volatile("st.global.cs.v4.s32 [%0], {0, 0, 0, 0};" ::"l"(p[k]) : "memory");
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (r0 + 4 * k < rows) asm volatile("st.global.cs.v4.s32 [%0], {0, 0, 0, 0};" ::"l"(p[k]) : "memory");
    }
}

// state of one column's envelope stack
struct ColState {
    int q;                      // index of the top entry (-1: empty)
    int s_top, t_top, g2_top;   // top entry (parabola row, first row it wins) and g(s_top)^2
    int base, ystart;           // first entry / first row of the open run
    bool open;
};

struct ColCtx {
    const unsigned short *gc;   // g column
    int *oc;                    // output column (doubles as the entry stack)
    int W, H;
};

__device__ __forceinline__ int env_f(int y, int sidx, int g2) { return (y - sidx) * (y - sidx) + g2; }

__device__ __forceinline__ void env_load_top(ColState &c, const ColCtx &k)
{
    const int e = k.oc[(size_t)c.q * k.W];
    c.s_top = e & 0xffff;
    c.t_top = (e >> 16) & 0xffff;
    const int gv = k.gc[(size_t)c.s_top * k.W];
    c.g2_top = gv * gv;         // entries only ever hold finite g
}

__device__ __forceinline__ void env_store_top(const ColState &c, const ColCtx &k)
{
    k.oc[(size_t)c.q * k.W] = c.s_top | (c.t_top << 16);
}

// add parabola (u, gu2); entries that would only win after row `limit` are dropped
__device__ __forceinline__ void env_insert(ColState &c, const ColCtx &k, int u, int gu2, int limit)
{
    while (c.q >= c.base && env_f(c.t_top, c.s_top, c.g2_top) > env_f(c.t_top, u, gu2)) {
        --c.q;
        if (c.q >= c.base) env_load_top(c, k);
    }
    if (c.q < c.base) {
        c.q = c.base;
        c.s_top = u; c.t_top = c.ystart; c.g2_top = gu2;
        env_store_top(c, k);
    } else {
        const int w = 1 + floor_div(u * u - c.s_top * c.s_top + gu2 - c.g2_top, 2 * (u - c.s_top));
        if (w <= limit) {
            ++c.q;
            c.s_top = u; c.t_top = w; c.g2_top = gu2;
            env_store_top(c, k);
        }
    }
}

__device__ __forceinline__ void env_forward_row(ColState &c, const ColCtx &k, int y, int gv)
{
    if (gv == 0) {
        if (c.open) {                           // the zero pixel at y closes the run [ystart, y-1]
            env_insert(c, k, y, 0, y - 1);
            c.open = false;
        }
    } else {
        if (!c.open) {
            c.open = true;
            c.base = c.q + 1;
            c.ystart = y;
            if (y > 0) env_insert(c, k, y - 1, 0, k.H - 1);     // the zero pixel just above the run
        }
        if (gv != G_INF) env_insert(c, k, y, gv * gv, k.H - 1);
    }
}

__device__ __forceinline__ int env_backward_row(ColState &c, const ColCtx &k, int y, int gv, int cap)
{
    if (gv == 0) return 0;
    if (c.q < 0) return cap;                    // no zero pixel anywhere on this column's runs
    const int val = min(env_f(y, c.s_top, c.g2_top), cap);
    if (y == c.t_top) {
        --c.q;
        if (c.q >= 0) env_load_top(c, k);
    }
    return val;
}

#ifndef SLN_ENV_AHEAD
#define SLN_ENV_AHEAD 4
#endif
constexpr int ENV_AHEAD = SLN_ENV_AHEAD;     // rows of g fetched per batch (independent loads in flight)

// thread = one column; the warp (32 adjacent columns = one flag segment) walks the rows in lock step
// (coalesced g reads and output writes), 32-row tile by tile: tiles whose flag word is zero hold only
// background in these 32 columns -- the forward sweep skips them outright and the backward sweep just
// stores zeros; inside a non-empty tile rows whose 32 pixels are background cost one vote.
__global__ void __launch_bounds__(128)
edt_cols_envelope_kernel(const unsigned short *__restrict__ g, const unsigned *__restrict__ flags, int tiles_y,
                         int tiles_x, int H, int W, int cap, int *__restrict__ out)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int seg = (blockIdx.x * blockDim.x + (threadIdx.x & ~31)) >> 5;     // warp-uniform
    if (x >= W) return;
    const unsigned act = __activemask();
    const int m = blockIdx.y;
    const unsigned short *__restrict__ gbase = g + (size_t)m * H * W + x;
    int *__restrict__ obase = out + (size_t)m * H * W + x;
    const unsigned *__restrict__ fl = flags + (size_t)m * tiles_y * tiles_x + seg;
    ColCtx k0{gbase, obase, W, H};
    ColState c0{-1, 0, 0, 0, 0, 0, false};

    // ---- forward: build the envelopes
    for (int ty = 0; ty < tiles_y; ++ty) {
        const unsigned f = __ldg(fl + (size_t)ty * tiles_x);
        if (f == 0u && !__any_sync(act, c0.open)) continue;
        const int yb = ty << 5, ye = min(H, yb + 32);
        for (int y0 = yb; y0 < ye; y0 += ENV_AHEAD) {
            unsigned r[ENV_AHEAD];
#pragma unroll
            for (int j = 0; j < ENV_AHEAD; ++j) {
                r[j] = 0u;
                if (y0 + j < ye && ((f >> ((y0 + j) & 31)) & 1u)) r[j] = __ldg(gbase + (size_t)(y0 + j) * W);
            }
#pragma unroll
            for (int j = 0; j < ENV_AHEAD; ++j) {
                const int y = y0 + j;
                if (y >= ye) break;
                if (__all_sync(act, r[j] == 0u && !c0.open)) continue;
                env_forward_row(c0, k0, y, (int)r[j]);
            }
        }
    }

    // ---- backward: evaluate
    if (c0.q >= 0) env_load_top(c0, k0);
    for (int ty = tiles_y - 1; ty >= 0; --ty) {
        const unsigned f = __ldg(fl + (size_t)ty * tiles_x);
        const int yb = ty << 5, ye = min(H, yb + 32);
        if (f == 0u) {                              // 32 rows x 32 columns of background
            if (act == 0xffffffffu) zero_tile(obase - (threadIdx.x & 31) + (size_t)yb * W, W, ye - yb, threadIdx.x & 31);
            else for (int y = ye - 1; y >= yb; --y) __stcs(obase + (size_t)y * W, 0);
            continue;
        }
        for (int y0 = ye - 1; y0 >= yb; y0 -= ENV_AHEAD) {
            unsigned r[ENV_AHEAD];
#pragma unroll
            for (int j = 0; j < ENV_AHEAD; ++j) {
                r[j] = 0u;
                if (y0 - j >= yb && ((f >> ((y0 - j) & 31)) & 1u)) r[j] = __ldg(gbase + (size_t)(y0 - j) * W);
            }
#pragma unroll
            for (int j = 0; j < ENV_AHEAD; ++j) {
                const int y = y0 - j;
                if (y < yb) break;
                int v = 0;
                if (!__all_sync(act, r[j] == 0u)) v = env_backward_row(c0, k0, y, (int)r[j], cap);
                __stcs(obase + (size_t)y * W, v);
            }
        }
    }
}

// Packed variant for H <= 2048, W <= 1024 (the 1024^2 sem-dist maps).  An envelope entry fits one word
// (s:11 | t:11 | g:10), so a pop needs no dependent second load, and the three entries below the top are
// mirrored in registers (e0 = top, e1, e2) with the refill load issued at every pop and consumed two pops
// later -- the L2 round trips leave the serial chain of the busy columns.
//
// The stack lives in the column's own g (u32 per pixel here): a push at row u has index k <= u (closed runs
// keep at most one entry per row, the open run one per parabola seen so far), and rows <= u of g have
// already been read by the forward sweep.  The backward sweep does not need g, only which pixels are
// foreground: the forward sweep leaves one word per (32-row tile, column) with those bits (`fgcol`).
// With the stack out of the way the output is written exactly once: non-empty tiles by the envelope warps,
// background tiles by the fill blocks (blockIdx.z == 1) that run beside them -- the bandwidth-bound zero
// fill overlaps the latency-bound envelope chains.
#ifndef SLN_EDT_MIRROR
#define SLN_EDT_MIRROR 3
#endif
constexpr int PK_D = SLN_EDT_MIRROR;   // stack entries mirrored in registers
struct PCol {
    int q, base, ystart;
    unsigned e[PK_D];           // entries q, q-1, .. q-PK_D+1 (garbage where the index is < 0); e[0] is the top
    bool open;
};

__device__ __forceinline__ int pk_s(unsigned e) { return (int)(e & 0x7ffu); }
__device__ __forceinline__ int pk_t(unsigned e) { return (int)((e >> 11) & 0x7ffu); }
__device__ __forceinline__ int pk_g2(unsigned e) { const int gq = (int)(e >> 22); return gq * gq; }
__device__ __forceinline__ unsigned pk_make(int sidx, int t, int gq) { return (unsigned)sidx | ((unsigned)t << 11) | ((unsigned)gq << 22); }

__device__ __forceinline__ void pk_pop(PCol &c, const unsigned *sc, int W)
{
    --c.q;
#pragma unroll
    for (int i = 0; i + 1 < PK_D; ++i) c.e[i] = c.e[i + 1];
    // plain (L1-cached) accesses on purpose: a sector holds the entries of 8 neighbouring columns, which pop
    // at nearly the same time -- with L2-only loads the kernel is 3.8x slower (2.66 ms vs 0.70 ms, 320 maps).
    // g is read with ld.cs (coherent), never ld.nc, so a thread always sees the entry it wrote over g.
    // The refill is consumed PK_D-1 pops later and reads an entry pushed at least PK_D-1 pushes ago: with a
    // shallow mirror the load chased the thread's own store through L2 (store -> load round trip per pop).
    if (c.q >= PK_D - 1) c.e[PK_D - 1] = sc[(size_t)(c.q - (PK_D - 1)) * W];
}

__device__ __forceinline__ void pk_push(PCol &c, unsigned *sc, int W, unsigned e)
{
    ++c.q;
#pragma unroll
    for (int i = PK_D - 1; i > 0; --i) c.e[i] = c.e[i - 1];
    c.e[0] = e;
    sc[(size_t)c.q * W] = e;
}

// floor(a / b) for |a| < 2^23, 0 < b < 2^13 (the packed kernel's ranges): one reciprocal and an exact fix-up
__device__ __forceinline__ int floor_div_small(int a, int b)
{
    int q = __float2int_rd(__fdividef((float)a, (float)b));       // within 1 of the true floor
    const int r = a - q * b;
    if (r < 0) --q; else if (r >= b) ++q;
    return q;
}

#ifdef SLN_EDT_NOINLINE
__device__ __noinline__ void pk_insert(PCol &c, unsigned *sc, int W, int u, int gq, int limit)
#else
__device__ __forceinline__ void pk_insert(PCol &c, unsigned *sc, int W, int u, int gq, int limit)
#endif
{
    const int gu2 = gq * gq;
    while (c.q >= c.base) {
        const int t = pk_t(c.e[0]);
        if (env_f(t, pk_s(c.e[0]), pk_g2(c.e[0])) > env_f(t, u, gu2)) pk_pop(c, sc, W);
        else break;
    }
    if (c.q < c.base) {
        pk_push(c, sc, W, pk_make(u, c.ystart, gq));
    } else {
        const int st = pk_s(c.e[0]);
        const int w = 1 + floor_div_small(u * u - st * st + gu2 - pk_g2(c.e[0]), 2 * (u - st));
        if (w <= limit) pk_push(c, sc, W, pk_make(u, w, gq));
    }
}

#ifndef SLN_EDT_COLS_BLOCK
#define SLN_EDT_COLS_BLOCK 128
#endif

// Zero fill of the background tiles of one map: warp `wid` of `nw` takes tile rows wid, wid + nw, ..
// warp = tile rows (32 output rows each), walked along x: one store instruction covers four adjacent
// segments = 512 contiguous bytes of one row, so runs of background are written as long bursts
// (A/B on 320 maps: 704 -> 676 us against column-wise 128-byte pieces at a 4 KB stride).
// (a segment cut by the right border is flagged on every row, so only full-width tiles get here)
__device__ __forceinline__ void edt_fill_rows(const unsigned *__restrict__ flm, int *__restrict__ om, int tiles_y, int tiles_x,
                                              int H, int W, int wid, int nw, int lane)
{
    constexpr unsigned FULL = 0xffffffffu;
    for (int ty = wid; ty < tiles_y; ty += nw) {
        const int yb = ty << 5, rows = min(H, yb + 32) - yb;
        for (int s0 = 0; s0 < tiles_x; s0 += 32) {
            const unsigned f = s0 + lane < tiles_x ? __ldg(flm + (size_t)ty * tiles_x + s0 + lane) : 1u;
            const unsigned emp = __ballot_sync(FULL, f == 0u);          // empty segments s0 .. s0+31
            if (!emp) continue;
#pragma unroll
            for (int grp = 0; grp < 8; ++grp) {                         // 4 segments = 128 columns
                if (!((emp >> (4 * grp)) & 0xfu)) continue;
                const bool mine = (emp >> (4 * grp + (lane >> 3))) & 1u;
                int *p = om + (size_t)yb * W + (size_t)(s0 + 4 * grp) * 32 + 4 * lane;
                if (mine) {
#pragma unroll 8
                    for (int y = 0; y < rows; ++y)
                        asm volatile("st.global.cs.v4.s32 [%0], {0, 0, 0, 0};" ::"l"(p + (size_t)y * W) : "memory");
                }
            }
        }
    }
}

// Split form of the packed column pass: the zero fill as its own launch of a few fat CTAs.  Each takes 227 KB of
// dynamic shared memory it never touches, so it owns its SM: the envelope kernel (launched right behind it as a
// programmatic dependent with a 4-KB request per CTA) cannot be placed beside it, and its latency-bound chains run on
// the other SMs without queueing behind the fill's stores in the LSU (lg-throttle was 3.1 stall cycles per issue when
// both roles shared every SM, and the two roles' times simply added up).
constexpr int EDT_FILL_THREADS = 1024;
constexpr int EDT_FILL_SMEM = 227 * 1024;
constexpr int EDT_ENV_SPLIT_SMEM = 4 * 1024;

__global__ void __launch_bounds__(EDT_FILL_THREADS, 1)
edt_fill_kernel(const unsigned *__restrict__ flags, int tiles_y, int tiles_x, int H, int W, int mc, int *__restrict__ out)
{
    asm volatile("griddepcontrol.launch_dependents;");
    const int lane = threadIdx.x & 31;
    const int wpc = EDT_FILL_THREADS >> 5;
    // item = (map, tile row); consecutive warps take consecutive tile rows of one map
    const long long items = (long long)mc * tiles_y;
    for (long long it = (long long)blockIdx.x * wpc + (threadIdx.x >> 5); it < items; it += (long long)gridDim.x * wpc) {
        const int m = (int)(it / tiles_y), ty = (int)(it - (long long)m * tiles_y);
        edt_fill_rows(flags + (size_t)m * tiles_y * tiles_x, out + (size_t)m * H * W, ty + 1, tiles_x, H, W, ty, 1 << 30, lane);
    }
}


// The warp's 32 columns are one flag segment.  All tile flags of the segment (<= 64 words) are fetched once
// into two registers per lane and broadcast by shuffle, so the sweeps visit only the non-empty 32-row tiles
// (bit mask `ne`), and the loads of the next 8-row batch -- possibly in a far-away tile -- are issued before
// the current batch is processed.  Lanes right of the border stay alive (no loads, no stores) so the shuffles
// are always full-warp.
__global__ void __launch_bounds__(SLN_EDT_COLS_BLOCK)
edt_cols_envelope_packed_kernel(unsigned *__restrict__ g, const unsigned *__restrict__ flags, unsigned *__restrict__ fgcol,
                                int tiles_y, int tiles_x, int H, int W, int cap, int *__restrict__ out)
{
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int x0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31);             // warp-uniform
    if (x0 >= W) {
        // split form (sln_edt_sq): one CTA past the right border keeps this grid from completing before the fill
        // kernel it was launched behind -- the next operation in the stream is ordered after this grid only
        if (blockIdx.x * blockDim.x >= W && blockIdx.y == 0) asm volatile("griddepcontrol.wait;" ::: "memory");
        return;
    }
    const int x = x0 + lane;
    const bool valid = x < W;
    const int m = blockIdx.y;
    const unsigned *__restrict__ fl = flags + (size_t)m * tiles_y * tiles_x + (x0 >> 5);
    const unsigned fw0 = lane < tiles_y ? __ldg(fl + (size_t)lane * tiles_x) : 0u;
    const unsigned fw1 = lane + 32 < tiles_y ? __ldg(fl + (size_t)(lane + 32) * tiles_x) : 0u;
    const unsigned long long ne = (unsigned long long)__ballot_sync(FULL, fw0 != 0u) |
                                  ((unsigned long long)__ballot_sync(FULL, fw1 != 0u) << 32);
    int *__restrict__ oc = out + (size_t)m * H * W + x;

    if (blockIdx.z == 1) {                          // ---- fill role: zero the background tiles
        edt_fill_rows(flags + (size_t)m * tiles_y * tiles_x, out + (size_t)m * H * W, tiles_y, tiles_x, H, W,
                      blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), gridDim.x * (blockDim.x >> 5), lane);
        return;
    }
#ifdef SLN_EDT_PROBE_NOENV
    return;                                         // timing probe (wrong output): fill role alone
#endif
    if (ne == 0ull) return;

    unsigned *__restrict__ sc = g + (size_t)m * H * W + x;                    // g column, then the entry stack
    unsigned *__restrict__ fgc = fgcol + (size_t)m * tiles_y * W + x;
    PCol c;
    c.q = -1; c.base = 0; c.ystart = 0; c.open = false;
#pragma unroll
    for (int i = 0; i < PK_D; ++i) c.e[i] = 0u;

    // ---- forward: build the envelopes
    {
        unsigned long long rem = ne;
        int ty = __ffsll((long long)rem) - 1, b = 0;
        unsigned f = __shfl_sync(FULL, ty < 32 ? fw0 : fw1, ty & 31);
        unsigned r[ENV_AHEAD], rn[ENV_AHEAD];
#pragma unroll
        for (int j = 0; j < ENV_AHEAD; ++j) {
            const int y = (ty << 5) + j;
            r[j] = (valid && y < H && ((f >> (y & 31)) & 1u)) ? __ldcs(sc + (size_t)y * W) : 0u;
        }
        unsigned fgw = 0u;
        while (ty >= 0) {
            const int y0 = (ty << 5) + b * ENV_AHEAD;
            int nty = ty, nb = b + 1;
            if (nb == 32 / ENV_AHEAD || y0 + ENV_AHEAD >= H) {
                rem &= rem - 1ull;
                nty = rem ? __ffsll((long long)rem) - 1 : -1;
                nb = 0;
            }
            unsigned nf = f;
            if (nty >= 0) {
                if (nty != ty) nf = __shfl_sync(FULL, nty < 32 ? fw0 : fw1, nty & 31);
#pragma unroll
                for (int j = 0; j < ENV_AHEAD; ++j) {
                    const int y = (nty << 5) + nb * ENV_AHEAD + j;
                    rn[j] = (valid && y < H && ((nf >> (y & 31)) & 1u)) ? __ldcs(sc + (size_t)y * W) : 0u;
                }
            }
#pragma unroll
            for (int j = 0; j < ENV_AHEAD; ++j) {
                const int y = y0 + j;
                if (y >= H) break;
                const int gv = (int)r[j];
                if (__all_sync(FULL, gv == 0 && !c.open)) continue;
                if (gv == 0) {
                    if (c.open) {                       // the zero pixel at y closes the run [ystart, y-1]
                        pk_insert(c, sc, W, y, 0, y - 1);
                        c.open = false;
                    }
                } else {
                    fgw |= 1u << (y & 31);
                    if (!c.open) {
                        c.open = true;
                        c.base = c.q + 1;
                        c.ystart = y;
                        if (y > 0) pk_insert(c, sc, W, y - 1, 0, H - 1);      // the zero pixel just above the run
                    }
                    if (gv != (int)G_INF) pk_insert(c, sc, W, y, gv, H - 1);
                }
            }
            if (nb == 0) {
                if (valid) __stcg(fgc + (size_t)ty * W, fgw);
                fgw = 0u;
                if (nty != ty + 1) {                    // the next tile (if any) is background: close what is open
                    const int ye = (ty << 5) + 32;
                    if (ye < H && c.open) {
                        pk_insert(c, sc, W, ye, 0, ye - 1);
                        c.open = false;
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < ENV_AHEAD; ++j) r[j] = rn[j];
            ty = nty; b = nb; f = nf;
        }
    }

    // ---- backward: evaluate the non-empty tiles, bottom one first
    {
        unsigned long long rem = ne;
        int ty = 63 - __clzll((long long)rem);
        unsigned fgw = valid ? __ldcg(fgc + (size_t)ty * W) : 0u;
        while (ty >= 0) {
            rem &= ~(1ull << ty);
            const int nty = rem ? 63 - __clzll((long long)rem) : -1;
            unsigned nfgw = 0u;
            if (nty >= 0 && valid) nfgw = __ldcg(fgc + (size_t)nty * W);
            const int yb = ty << 5;
#pragma unroll 8
            for (int y = min(H, yb + 32) - 1; y >= yb; --y) {
                int v = 0;
                if ((fgw >> (y & 31)) & 1u) {
                    if (c.q < 0) {
                        v = cap;                    // no zero pixel anywhere on this column's runs
                    } else {
                        v = min(env_f(y, pk_s(c.e[0]), pk_g2(c.e[0])), cap);
                        if (y == pk_t(c.e[0])) pk_pop(c, sc, W);
                    }
                }
                if (valid) __stcs(oc + (size_t)y * W, v);
            }
            ty = nty; fgw = nfgw;
        }
    }
}

// (one column per thread was measured fastest: 1476 / 1926 / 3205 us for 1 / 2 / 4 columns per thread on 320
// 1024^2 maps -- the scan is latency bound, warps in flight matter more than instruction count)

#ifndef SLN_EDT_FILL_CTAS_DEFAULT
#define SLN_EDT_FILL_CTAS_DEFAULT 0
#endif
constexpr size_t EDT_CHUNK_BYTES = 2048ull << 20;   // the column pass needs thousands of columns in flight: big chunks

// shapes the packed envelope kernel takes (entry fields s:11 | t:11 | g:10, two flag words per segment)
static bool edt_packed_shape(int H, int W) { return W >= 128 && W % 4 == 0 && H <= 2048 && W <= 1024; }

// number of fat fill CTAs of the split form (0: fill role inside the envelope launch); SLN_EDT_FILL_CTAS overrides (A/B)
static int edt_fill_ctas()
{
    const char *e = getenv("SLN_EDT_FILL_CTAS");
    int f = e ? atoi(e) : SLN_EDT_FILL_CTAS_DEFAULT;
    const int cap = sm_count() / 2;
    return f < 0 ? 0 : (f > cap ? cap : f);
}

static int edt_chunk_maps(int M, int H, int W)
{
    const size_t per = (edt_packed_shape(H, W) ? sizeof(unsigned) : sizeof(unsigned short)) * (size_t)H * W;
    size_t c = per ? EDT_CHUNK_BYTES / per : (size_t)M;
    if (c < 1) c = 1;
    if (c > (size_t)M) c = (size_t)M;
    return (int)c;
}


// ---------------------------------------------------------------------------
// Banded exact EDT (round 2): two launches, no row-distance buffer.
// The per-column logic is edt_band.cuh (shared with the CPU simulation under tests/host_sim); the kernels below
// add the row pass, the work distribution and the zero fill.
//   edt_band_build_kernel : CTA = (band of 32 rows, map).  Loads the zero masks of its 34 rows (one extra above and
//       below) -- 1 bit per pixel in shared memory -- and, per row that holds foreground, the distance from every
//       32-pixel segment to the nearest zero pixel outside it (two warp scans).  A band without foreground is
//       zero-filled on the spot (128 KB of streaming stores); otherwise the empty 32x32 tiles are zero-filled and the
//       non-empty ones get their envelope stacks built, one warp per tile, the row distance of a pixel computed on the
//       fly from (mask, left, right) -- g never exists in memory.
//   edt_band_eval_kernel  : persistent warps that draw the non-empty tiles from a work list, launched as a programmatic
//       dependent; only those tiles are evaluated and written (own stack staged in shared memory, the bands a run
//       continues into read through L2).
// ---------------------------------------------------------------------------
namespace eb = edtband;
constexpr int EB_WARPS = 8;

__device__ __forceinline__ void st_zero16(int *p)
{
    asm volatile("st.global.cs.v4.s32 [%0], {0, 0, 0, 0};" ::"l"(p) : "memory");
}

#ifndef SLN_EB_CTAS
#define SLN_EB_CTAS 5
#endif
__global__ void __launch_bounds__(EB_WARPS * 32, SLN_EB_CTAS)
edt_band_build_kernel(const unsigned char *__restrict__ maps, int H, int W, int nb, unsigned *__restrict__ list,
                      unsigned *__restrict__ counters, uint2 *__restrict__ meta, unsigned *__restrict__ stk, int *__restrict__ out)
{
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int NROW = eb::BAND + 2;
    constexpr int PER_WARP = (NROW + EB_WARPS - 1) / EB_WARPS;
    __shared__ unsigned s_mask[NROW][32];
    __shared__ int s_ld[eb::BAND][32];
    __shared__ int s_rd[eb::BAND][32];
    __shared__ unsigned s_rowbits[eb::BAND];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x, m = blockIdx.y;
    const int tx = W >> 5;
    const int yb = b * eb::BAND, rows = min(eb::BAND, H - yb);
    const unsigned char *__restrict__ src = maps + (size_t)m * H * W;

    // ---- zero masks of rows yb-1 .. yb+32 (rows outside the map: no zero pixels), loads first
    unsigned zr[PER_WARP];
#pragma unroll
    for (int j = 0; j < PER_WARP; ++j) {
        const int i = warp + j * EB_WARPS, yy = yb - 1 + i;
        zr[j] = 0u;
        if (i < NROW && yy >= 0 && yy < H) zr[j] = seg_zero_mask(src + (size_t)yy * W, lane * 32, W, true);
    }
#pragma unroll
    for (int j = 0; j < PER_WARP; ++j) {
        const int i = warp + j * EB_WARPS;
        if (i >= NROW) break;
        const unsigned z = zr[j];
        s_mask[i][lane] = z;
        if (i >= 1 && i <= eb::BAND) {
            const unsigned rb = __ballot_sync(FULL, lane < tx && z != FULL) & (yb - 1 + i < H ? FULL : 0u);
            if (lane == 0) s_rowbits[i - 1] = rb;
            if (rb) {                                   // the row holds foreground: nearest zero outside every segment
                const int x0 = lane * 32;
                int left = z ? x0 + 31 - __clz(z) : -eb::NONE_D;
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(FULL, left, o);
                    if (lane >= o) left = max(left, v);
                }
                left = __shfl_up_sync(FULL, left, 1);
                if (lane == 0) left = -eb::NONE_D;
                int right = z ? x0 + __ffs(z) - 1 : eb::NONE_D;
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_down_sync(FULL, right, o);
                    if (lane + o < 32) right = min(right, v);
                }
                right = __shfl_down_sync(FULL, right, 1);
                if (lane == 31) right = eb::NONE_D;
                s_ld[i - 1][lane] = left == -eb::NONE_D ? eb::NONE_D : x0 - left;
                s_rd[i - 1][lane] = right == eb::NONE_D ? eb::NONE_D : right - (x0 + 31);
            }
        }
    }
    __syncthreads();

    const unsigned anym = __reduce_or_sync(FULL, s_rowbits[lane]);          // segments that hold foreground
    int *__restrict__ om = out + (size_t)m * H * W + (size_t)yb * W;
    if (anym == 0u) {                                                        // background band: zeros, done
        const int c4 = threadIdx.x * 4;
        if (c4 < W) {
#pragma unroll 8
            for (int r = 0; r < rows; ++r) st_zero16(om + (size_t)r * W + c4);
        }
        return;
    }
    // the band's non-empty tiles join the evaluation pass's work list (order: whatever the atomics give; every tile is
    // evaluated independently, so the result does not depend on it)
    if (warp == 0) {
        const int n = __popc(anym);
        unsigned base = 0u;
        if (lane == 0) base = atomicAdd(counters, (unsigned)n);
        base = __shfl_sync(FULL, base, 0);
        if (lane < n) list[base + lane] = (unsigned)((((size_t)m * nb + b) << 5) | __fns(anym, 0, lane + 1));
    }
    // flag word of segment `lane`: bit r <=> row yb + r holds foreground there
    unsigned fmine = 0u;
#pragma unroll
    for (int r = 0; r < eb::BAND; ++r) fmine |= ((s_rowbits[r] >> lane) & 1u) << r;

    // ---- zero fill of the band's empty tiles: rows dealt over the warps, one store instruction = 512 contiguous bytes
    const unsigned segmask = tx >= 32 ? FULL : ((1u << tx) - 1u);
    const unsigned emp = ~anym & segmask;
    if (emp) {
        for (int r = warp; r < rows; r += EB_WARPS) {
#pragma unroll
            for (int grp = 0; grp < 8; ++grp) {
                if (!((emp >> (4 * grp)) & 0xfu)) continue;
                if ((emp >> (4 * grp + (lane >> 3))) & 1u) st_zero16(om + (size_t)r * W + grp * 128 + lane * 4);
            }
        }
    }
    // ---- envelope stacks of the non-empty tiles, dealt over the warps by rank
    unsigned ne = anym;
    for (int k = 0; ne; ++k) {
        const int s = __ffs(ne) - 1;
        ne &= ne - 1u;
        if ((k % EB_WARPS) != warp) continue;
        const unsigned f = __shfl_sync(FULL, fmine, s);
        const int x = s * 32 + lane;
        const bool above = yb > 0, below = yb + eb::BAND < H;
        const bool az = above && ((s_mask[0][s] >> lane) & 1u), bz = below && ((s_mask[NROW - 1][s] >> lane) & 1u);
        unsigned *sc = stk + ((size_t)m * nb + b) * eb::SLOTS * W + x;
        const eb::BuildResult res = eb::band_build_lane(f, rows, yb, H, lane, az, above && !az, bz, below && !bz, sc, W,
                                                        [&](int r, unsigned &z, int &ld, int &rd) {
                                                            z = s_mask[r + 1][s];
                                                            ld = s_ld[r][s];
                                                            rd = s_rd[r][s];
                                                        });
        meta[((size_t)m * nb + b) * W + x] = make_uint2(res.fgw, res.meta);
    }
}

constexpr int EV_WARPS = 4;      // 8.3 KB of shared memory per warp: one stack buffer (own band, then each neighbour) + the minima
constexpr int EV_CTAS_PER_SM = 6;

// Persistent: every warp draws tiles from the list the build pass left (ticket counter), so the few bands that hold a
// blob's tiles are spread over the whole GPU instead of queueing behind one another in one CTA.  The next tile's ticket,
// list entry and words are fetched while the current tile is processed.  A tile's running minima stay in shared memory
// until every band in reach has been seen; the tile is then written once (zeros where the pixel is background).
__global__ void __launch_bounds__(EV_WARPS * 32, EV_CTAS_PER_SM)
edt_band_eval_kernel(const unsigned *__restrict__ list, unsigned *__restrict__ counters, const uint2 *__restrict__ meta,
                     const unsigned *__restrict__ stk, int H, int W, int nb, int cap, int *__restrict__ out)
{
    constexpr unsigned FULL = 0xffffffffu;
    pdl_prologue();
    __shared__ unsigned s_buf[EV_WARPS][eb::SLOTS][32];
    __shared__ int s_best[EV_WARPS][eb::BAND][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned count = __ldcg(counters);
    auto draw = [&]() {                                    // lane 0 holds the ticket; broadcast where it is needed
        unsigned t = 0u;
        if (lane == 0) t = atomicAdd(counters + 1, 1u);
        return t;
    };
    // every lane copies ITS column's entries [kmin, kmax) of the stack at bp (kmin / kmax: over the lanes that need them);
    // a lane only ever reads back what it wrote itself, so no barrier is involved
    auto stage = [&](const unsigned *bp, int lo, int hi, bool need, const unsigned *&ptr, int &stride) {
        const int kmin = __reduce_min_sync(FULL, need ? lo : eb::SLOTS), kmax = __reduce_max_sync(FULL, need ? hi : 0);
        // global -> shared without a register in between (LDGSTS): every entry of the range is in flight at once, where a
        // load / store loop unrolled by four waited out one L2 round trip per four entries (26 % of this kernel's stall
        // samples).  A lane waits for its own copies only -- it reads back nothing else.
        const unsigned *src = bp + (size_t)kmin * W;
        unsigned dst = (unsigned)__cvta_generic_to_shared(&s_buf[warp][kmin][lane]);
        for (int q = kmin; q < kmax; ++q, src += W, dst += 32 * 4)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
        asm volatile("cp.async.wait_all;" ::: "memory");
        ptr = &s_buf[warp][0][lane];
        stride = 32;
    };
    // tickets are drawn two tiles ahead: the atomic's round trip (7 % of the stall samples when its result was needed at
    // once) overlaps a whole tile; the list entry and the words of the next tile are fetched one tile ahead
    unsigned idx = __shfl_sync(FULL, draw(), 0);
    if (idx >= count) return;
    unsigned idx_next = draw();
    unsigned tile = __ldg(list + idx);
    uint2 mm = __ldg(meta + (size_t)(tile >> 5) * W + (tile & 31u) * 32 + lane);
    for (;;) {
        const unsigned idx_n = __shfl_sync(FULL, idx_next, 0);
        if (idx_n < count) idx_next = draw();              // (a warp that saw the end of the list stops drawing)
        unsigned tile_n = 0u;
        uint2 mm_n = make_uint2(0u, 0u);
        if (idx_n < count) {
            tile_n = __ldg(list + idx_n);
            mm_n = __ldg(meta + (size_t)(tile_n >> 5) * W + (tile_n & 31u) * 32 + lane);
        }
        const int s = (int)(tile & 31u);
        const size_t mb = tile >> 5;                                   // m * nb + b
        const int m = (int)(mb / (unsigned)nb), b = (int)(mb - (size_t)m * nb);
        const size_t band0 = (size_t)m * nb;
        const int yb = b * eb::BAND, rows = min(eb::BAND, H - yb);
        const int x = s * 32 + lane;
        const unsigned *own;
        int own_stride;
        stage(stk + mb * eb::SLOTS * W + x, 0, eb::meta_total(mm.y), true, own, own_stride);
        eb::band_eval_lane(b, nb, yb, rows, cap, mm.x, mm.y, own, own_stride, stk + band0 * eb::SLOTS * W + x,
                           (size_t)eb::SLOTS * W, W, reinterpret_cast<const eb::Words2 *>(meta) + band0 * W + x, W,
                           &s_best[warp][0][lane], 32, stage);
        int *__restrict__ oc = out + (size_t)m * H * W + (size_t)yb * W + x;
#pragma unroll 8
        for (int r = 0; r < rows; ++r) __stcs(oc + (size_t)r * W, ((mm.x >> r) & 1u) ? s_best[warp][r][lane] : 0);
        if (idx_n >= count) break;
        tile = tile_n;
        mm = mm_n;
    }
}

// shapes the banded kernels take: whole 32-column segments, entry fields s:11 | t:11 | g:10
static bool edt_band_shape(int H, int W) { return W >= 32 && W % 32 == 0 && W <= 1024 && H <= 2048; }
static bool edt_band_enabled()
{
    const char *e = getenv("SLN_EDT_IMPL");       // "legacy": the whole-column envelope kernels of round 1 (A/B)
    return !(e && e[0] == 'l');
}
static size_t edt_band_map_bytes(int H, int W)
{
    const size_t nb = (size_t)cdiv(H, eb::BAND);
    return nb * ((size_t)eb::SLOTS * W * sizeof(unsigned) + (size_t)W * sizeof(uint2) + 32 * sizeof(unsigned)) + 1;
}
static int edt_band_chunk_maps(int M, int H, int W)
{
    size_t c = (3072ull << 20) / edt_band_map_bytes(H, W);
    if (c < 1) c = 1;
    if (c > (size_t)M) c = (size_t)M;
    return (int)c;
}

}  // namespace sln

using namespace sln;

extern "C" int sln_layer_decode(const uint64_t *label, int B, int H, int W, int L, int n_max, uint8_t *out,
                                int *n_obj, uint32_t *scratch, void *stream)
{
    SLN_REQUIRE(B >= 0 && H >= 0 && W >= 0, SLN_ERR_ARG, "negative size");
    SLN_REQUIRE(L >= 1 && n_max >= 0 && n_max <= 32, SLN_ERR_ARG, "need L >= 1 and 0 <= n_max <= 32 (got %d, %d)", L, n_max);
    if (B == 0) return SLN_OK;
    SLN_REQUIRE(n_obj && scratch, SLN_ERR_ARG, "null pointer");
    SLN_REQUIRE(B <= 65535, SLN_ERR_ARG, "B > 65535");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t px = (size_t)H * W;
    SLN_CUDA_OK(cudaMemsetAsync(scratch, 0, sizeof(uint32_t) * 2 * (size_t)B, st));
    if (px == 0) {
        SLN_CUDA_OK(cudaMemsetAsync(n_obj, 0, sizeof(int) * (size_t)B, st));
        return SLN_OK;
    }
    SLN_REQUIRE(label && (out || n_max == 0), SLN_ERR_ARG, "null pointer");
    SLN_REQUIRE((reinterpret_cast<uintptr_t>(out) & 7u) == 0 && (reinterpret_cast<uintptr_t>(label) & 15u) == 0,
                SLN_ERR_LAYOUT, "label must be 16-byte and out 8-byte aligned");
    const bool wide = px % 16 == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0;
    const int ld_px = wide ? 16 : 8;
    const size_t threads = (px + ld_px - 1) / ld_px;
    SLN_REQUIRE((threads + 255) / 256 < (1ull << 31), SLN_ERR_ARG, "image too large");
    const dim3 grid((unsigned)((threads + 255) / 256), B);
    if (wide)
        layer_decode_kernel<16><<<grid, 256, 0, st>>>(reinterpret_cast<const unsigned long long *>(label), px, L, n_max, out, scratch);
    else
        layer_decode_kernel<8><<<grid, 256, 0, st>>>(reinterpret_cast<const unsigned long long *>(label), px, L, n_max, out, scratch);
    SLN_LAUNCH_OK("layer_decode_kernel");
    // grid-strided clearing loop; one CTA per SM and image is enough to stream the zeros when an image needs them.
    // Programmatic dependent of the decode kernel: its CTAs are placed while the decode's last wave drains.
    SLN_CUDA_OK(launch_chain(layer_fixup_kernel, dim3(sm_count(), B), dim3(256), 0, st, true, (const unsigned *)scratch, px, L, n_max,
                             (unsigned char *)out, n_obj));
    SLN_LAUNCH_OK("layer_fixup_kernel");
    return SLN_OK;
}

static size_t edt_flags_bytes(int mc, int H, int W)
{
    return align_up(sizeof(unsigned) * (size_t)mc * cdiv(H, 32) * cdiv(W, 32), 256);
}
static size_t edt_g_bytes(int mc, int H, int W)
{
    return align_up((edt_packed_shape(H, W) ? sizeof(unsigned) : sizeof(unsigned short)) * (size_t)mc * H * W, 256);
}
static size_t edt_fgcol_bytes(int mc, int H, int W)
{
    return edt_packed_shape(H, W) ? align_up(sizeof(unsigned) * (size_t)mc * cdiv(H, 32) * W, 256) : 0;
}
extern "C" size_t sln_edt_workspace_bytes(int M, int H, int W)
{
    if (M <= 0 || H <= 0 || W <= 0) return 0;
    const int mc = edt_chunk_maps(M, H, W);
    size_t need = edt_g_bytes(mc, H, W) + edt_flags_bytes(mc, H, W) + edt_fgcol_bytes(mc, H, W);
    if (edt_band_shape(H, W)) {      // the banded kernels' stacks + words + flags (either path may be taken at run time)
        const size_t nb = (size_t)cdiv(H, eb::BAND), bc = (size_t)edt_band_chunk_maps(M, H, W);
        const size_t band = align_up(bc * nb * eb::SLOTS * W * sizeof(unsigned), 256) + align_up(bc * nb * W * sizeof(uint2), 256) +
                            align_up(bc * nb * 32 * sizeof(unsigned), 256) + 256;
        if (band > need) need = band;
    }
    return need;
}
extern "C" int sln_edt_sq(const uint8_t *maps, int M, int H, int W, int32_t *out, void *workspace,
                          size_t workspace_bytes, void *stream)
{
    SLN_REQUIRE(M >= 0 && H >= 0 && W >= 0, SLN_ERR_ARG, "negative size");
    if (M == 0 || H == 0 || W == 0) return SLN_OK;
    SLN_REQUIRE(W <= 65534 && H <= 65534 && (long long)(H + W) * (H + W) < (1ll << 31), SLN_ERR_ARG,
                "map %dx%d too large for i32 squared distances", H, W);
    SLN_REQUIRE(maps && out, SLN_ERR_ARG, "null pointer");
    SLN_REQUIRE(workspace && workspace_bytes >= sln_edt_workspace_bytes(M, H, W), SLN_ERR_WORKSPACE,
                "edt workspace: need %zu bytes, got %zu", sln_edt_workspace_bytes(M, H, W), workspace_bytes);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (edt_band_shape(H, W) && edt_band_enabled() &&
        ((reinterpret_cast<uintptr_t>(maps) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(workspace)) & 15u) == 0) {
        const int nb = cdiv(H, eb::BAND), bc = edt_band_chunk_maps(M, H, W);
        unsigned char *w0 = static_cast<unsigned char *>(workspace);
        unsigned *stk = reinterpret_cast<unsigned *>(w0);
        uint2 *meta = reinterpret_cast<uint2 *>(w0 + align_up((size_t)bc * nb * eb::SLOTS * W * sizeof(unsigned), 256));
        unsigned *tlist = reinterpret_cast<unsigned *>(reinterpret_cast<unsigned char *>(meta) + align_up((size_t)bc * nb * W * sizeof(uint2), 256));
        unsigned *counters = reinterpret_cast<unsigned *>(reinterpret_cast<unsigned char *>(tlist) + align_up((size_t)bc * nb * 32 * sizeof(unsigned), 256));
        const int bcap = (H + W) * (H + W);
        SLN_REQUIRE((size_t)bc * nb < (1ull << 27), SLN_ERR_ARG, "too many bands per chunk");
        for (int m0 = 0; m0 < M; m0 += bc) {
            const int mc = (M - m0) < bc ? (M - m0) : bc;
            SLN_REQUIRE(mc <= 65535, SLN_ERR_ARG, "too many maps per chunk");
            SLN_CUDA_OK(cudaMemsetAsync(counters, 0, 2 * sizeof(unsigned), st));      // tiles listed, tickets drawn
            edt_band_build_kernel<<<dim3(nb, mc), EB_WARPS * 32, 0, st>>>(maps + (size_t)m0 * H * W, H, W, nb, tlist, counters, meta, stk,
                                                                          out + (size_t)m0 * H * W);
            SLN_LAUNCH_OK("edt_band_build_kernel");
            SLN_CUDA_OK(launch_chain(edt_band_eval_kernel, dim3(sm_count() * EV_CTAS_PER_SM), dim3(EV_WARPS * 32), 0, st, true,
                                     (const unsigned *)tlist, counters, (const uint2 *)meta, (const unsigned *)stk, H, W, nb, bcap,
                                     out + (size_t)m0 * H * W));
            SLN_LAUNCH_OK("edt_band_eval_kernel");
        }
        return SLN_OK;
    }
    const int chunk = edt_chunk_maps(M, H, W);
    unsigned char *wsb = static_cast<unsigned char *>(workspace);
    unsigned short *g = reinterpret_cast<unsigned short *>(wsb);              // u32 elements on the packed path
    unsigned *flags = reinterpret_cast<unsigned *>(wsb + edt_g_bytes(chunk, H, W));
    unsigned *fgcol = reinterpret_cast<unsigned *>(wsb + edt_g_bytes(chunk, H, W) + edt_flags_bytes(chunk, H, W));
    const int tiles_y = cdiv(H, 32), tiles_x = cdiv(W, 32);
    const int cap = (H + W) * (H + W);
    const bool vec = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15u) == 0) &&
                     ((reinterpret_cast<uintptr_t>(workspace) & 15u) == 0);
    SLN_REQUIRE(cdiv(H, 8) <= 65535, SLN_ERR_ARG, "H too large");
    // the packed envelope kernel carries g inside its entries and consults the per-row flag bits, so the row
    // pass may leave all-background segments of g unwritten
    const bool packed = vec && edt_packed_shape(H, W);
    for (int m0 = 0; m0 < M; m0 += chunk) {
        const int mc = (M - m0) < chunk ? (M - m0) : chunk;
        const long long n_rows = (long long)mc * H;
        const long long blocks = (n_rows + 7) / 8;
        SLN_REQUIRE(blocks < (1ll << 31) && mc <= 65535, SLN_ERR_ARG, "too many rows");
        SLN_CUDA_OK(cudaMemsetAsync(flags, 0, edt_flags_bytes(mc, H, W), st));
        if (packed)
            edt_rows_kernel<unsigned><<<(unsigned)blocks, 256, 0, st>>>(maps + (size_t)m0 * H * W, H, W, n_rows,
                                                                          reinterpret_cast<unsigned *>(g), flags, tiles_y, tiles_x, true);
        else
            edt_rows_kernel<unsigned short><<<(unsigned)blocks, 256, 0, st>>>(maps + (size_t)m0 * H * W, H, W, n_rows, g, flags,
                                                                                tiles_y, tiles_x, false);
        SLN_LAUNCH_OK("edt_rows_kernel");
        const dim3 cgrid(cdiv(W, 128), cdiv(H, 8), mc);
        if (packed && edt_fill_ctas() > 0) {
            // split form: fat fill CTAs that own their SMs, the envelope kernel as a programmatic dependent beside them
            SLN_CUDA_OK(cudaFuncSetAttribute(edt_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, EDT_FILL_SMEM));
            edt_fill_kernel<<<edt_fill_ctas(), EDT_FILL_THREADS, EDT_FILL_SMEM, st>>>(flags, tiles_y, tiles_x, H, W, mc, out + (size_t)m0 * H * W);
            SLN_LAUNCH_OK("edt_fill_kernel");
            SLN_CUDA_OK(launch_chain(edt_cols_envelope_packed_kernel, dim3(cdiv(W, SLN_EDT_COLS_BLOCK) + 1, mc, 1), dim3(SLN_EDT_COLS_BLOCK),
                                     (size_t)EDT_ENV_SPLIT_SMEM, st, true, reinterpret_cast<unsigned *>(g), (const unsigned *)flags, fgcol,
                                     tiles_y, tiles_x, H, W, cap, out + (size_t)m0 * H * W));
        } else if (packed) {
            edt_cols_envelope_packed_kernel<<<dim3(cdiv(W, SLN_EDT_COLS_BLOCK), mc, 2), SLN_EDT_COLS_BLOCK, 0, st>>>(
                reinterpret_cast<unsigned *>(g), flags, fgcol, tiles_y, tiles_x, H, W, cap, out + (size_t)m0 * H * W);
        } else if (vec && W >= 128) {
            edt_cols_envelope_kernel<<<dim3(cdiv(W, 128), mc), 128, 0, st>>>(g, flags, tiles_y, tiles_x, H, W, cap, out + (size_t)m0 * H * W);
        } else if (vec) {
            edt_cols_kernel<true><<<cgrid, 256, 0, st>>>(g, H, W, cap, out + (size_t)m0 * H * W);
        } else {
            edt_cols_kernel<false><<<cgrid, 256, 0, st>>>(g, H, W, cap, out + (size_t)m0 * H * W);
        }
        SLN_LAUNCH_OK("edt_cols_kernel");
    }
    return SLN_OK;
}
