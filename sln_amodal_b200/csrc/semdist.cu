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
#include "common.cuh"

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
constexpr int LD_PX = 8;     // pixels per thread

__device__ __forceinline__ unsigned long long expand_bits(const unsigned long long *__restrict__ lut, unsigned m)
{
    return lut[m & 0xffu];
}

__global__ void __launch_bounds__(256)
layer_decode_kernel(const unsigned long long *__restrict__ label, size_t px_per_image, int L, int n_max,
                    unsigned char *__restrict__ out, unsigned *__restrict__ seen /* [B][2]: top bits, any bits */)
{
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
                    __stcs(reinterpret_cast<unsigned long long *>(dst), bytes);
                } else {
                    for (int k = 0; k < LD_PX && p0 + k < px_per_image; ++k) dst[k] = (unsigned char)(bytes >> (8 * k));
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
                        __stcs(reinterpret_cast<unsigned long long *>(dst), bytes);
                    } else {
                        for (int k = 0; k < LD_PX && p0 + k < px_per_image; ++k) dst[k] = (unsigned char)(bytes >> (8 * k));
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
        const unsigned w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int q = 0; q < 8; ++q) {
#pragma unroll
            for (int e = 0; e < 4; ++e)
                z |= (unsigned)(((w[q] >> (8 * e)) & 0xffu) == 0u) << (4 * q + e);
        }
    } else {
        for (int k = 0; k < ROW_SEG && x0 + k < W; ++k) z |= (unsigned)(__ldg(src + x0 + k) == 0) << k;
    }
    return z;
}

// One warp per row.  Lane l owns 32 consecutive pixels of each 1024-pixel chunk; the
// nearest zero outside the segment comes from warp scans of the per-lane last / first
// zero positions, plus a carry between chunks (second, backward sweep only for W > 1024).
__global__ void __launch_bounds__(256)
edt_rows_kernel(const unsigned char *__restrict__ maps, int W, long long n_rows, unsigned short *__restrict__ g)
{
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const unsigned char *src = maps + row * W;
    unsigned short *dst = g + row * W;
    const bool vec_ok = (W % 16) == 0 && ((reinterpret_cast<uintptr_t>(maps) & 15u) == 0);
    const int n_chunks = (W + ROW_CHUNK - 1) / ROW_CHUNK;
    constexpr int NONE_R = 1 << 29;
    int carry_left = -NONE_R;                    // last zero in earlier chunks
    for (int c = 0; c < n_chunks; ++c) {
        const int x0 = c * ROW_CHUNK + lane * ROW_SEG;
        const unsigned z = seg_zero_mask(src, x0, W, vec_ok);
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
        if (z == 0xffffffffu && vec_ok && x0 + ROW_SEG <= W) {
            // every pixel of the segment is a zero pixel (the common case: instance masks are
            // mostly background): all distances are 0
            uint4 *o = reinterpret_cast<uint4 *>(dst + x0);
            const uint4 zz = make_uint4(0u, 0u, 0u, 0u);
            o[0] = zz; o[1] = zz; o[2] = zz; o[3] = zz;
        } else if (x0 < W) {
            // distance of pixel k = min(k - nearest zero bit at or below k, nearest zero bit at or above k - k),
            // falling back to the zeros left / right of the segment; 8 pixels per 16-byte store
            const bool vst = vec_ok && x0 + ROW_SEG <= W;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                unsigned pk[4];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int k = 8 * q + e;
                    const unsigned below = z & (0xffffffffu >> (31 - k));      // zero pixels at positions <= k
                    const unsigned above = z >> k;                             // zero pixels at positions >= k
                    const int lz = below ? x0 + 31 - __clz(below) : left;
                    const int rz = above ? x0 + k + __ffs(above) - 1 : right;
                    const int dl = min(x0 + k - lz, (int)G_INF), dr = min(rz - (x0 + k), (int)G_INF);
                    const unsigned dv = (unsigned)min(dl, dr);
                    if (e & 1) pk[e >> 1] |= dv << 16; else pk[e >> 1] = dv;
                    if (!vst && x0 + k < W) dst[x0 + k] = (unsigned short)dv;
                }
                if (vst) reinterpret_cast<uint4 *>(dst + x0)[q] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
        }
    }
    if (n_chunks > 1) {                          // zeros in later chunks
        __syncwarp();
        int carry_right = NONE_R;
        for (int c = n_chunks - 1; c >= 0; --c) {
            const int x0 = c * ROW_CHUNK + lane * ROW_SEG;
            const unsigned z = seg_zero_mask(src, x0, W, vec_ok);
            if (carry_right != NONE_R) {
                for (int k = 0; k < ROW_SEG && x0 + k < W; ++k) {
                    const int dr = min(carry_right - (x0 + k), (int)G_INF);
                    if (dr < (int)dst[x0 + k]) dst[x0 + k] = (unsigned short)dr;
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

constexpr int ENV_AHEAD = 8;     // rows of g fetched per batch (independent loads in flight)

template <int NC> struct GRow;
template <> struct GRow<1> { unsigned short v; };
template <> struct GRow<2> { unsigned v; };
template <> struct GRow<4> { uint2 v; };

template <int NC>
__device__ __forceinline__ void grow_load(const unsigned short *p, unsigned (&gv)[4])
{
    gv[0] = gv[1] = gv[2] = gv[3] = 0u;
    if (NC == 1) {
        gv[0] = __ldg(p);
    } else if (NC == 2) {
        const unsigned t = __ldg(reinterpret_cast<const unsigned *>(p));
        gv[0] = t & 0xffffu; gv[1] = t >> 16;
    } else {
        const uint2 t = __ldg(reinterpret_cast<const uint2 *>(p));
        gv[0] = t.x & 0xffffu; gv[1] = t.x >> 16; gv[2] = t.y & 0xffffu; gv[3] = t.y >> 16;
    }
}

// thread = NC adjacent columns; the warp walks the rows in lock step (coalesced g reads and
// output writes; rows whose pixels are all background cost a vote and a store).
template <int NC>
__global__ void __launch_bounds__(128)
edt_cols_envelope_kernel(const unsigned short *__restrict__ g, int H, int W, int cap, int *__restrict__ out)
{
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * NC;
    if (x >= W) return;
    const unsigned act = __activemask();
    const int m = blockIdx.y;
    const unsigned short *__restrict__ gbase = g + (size_t)m * H * W + x;
    int *__restrict__ obase = out + (size_t)m * H * W + x;
    ColCtx k0{gbase + 0, obase + 0, W, H}, k1{gbase + 1, obase + 1, W, H};
    ColCtx k2{gbase + 2, obase + 2, W, H}, k3{gbase + 3, obase + 3, W, H};
    ColState c0{-1, 0, 0, 0, 0, 0, false}, c1 = c0, c2 = c0, c3 = c0;

    // ---- forward: build the envelopes
    for (int y0 = 0; y0 < H; y0 += ENV_AHEAD) {
        unsigned r[ENV_AHEAD][4];
#pragma unroll
        for (int j = 0; j < ENV_AHEAD; ++j) {
            r[j][0] = r[j][1] = r[j][2] = r[j][3] = 0u;
            if (y0 + j < H) grow_load<NC>(gbase + (size_t)(y0 + j) * W, r[j]);
        }
#pragma unroll
        for (int j = 0; j < ENV_AHEAD; ++j) {
            const int y = y0 + j;
            if (y >= H) break;
            const bool idle = (r[j][0] | r[j][1] | r[j][2] | r[j][3]) == 0u && !(c0.open | c1.open | c2.open | c3.open);
            if (__all_sync(act, idle)) continue;
            env_forward_row(c0, k0, y, (int)r[j][0]);
            if (NC >= 2) env_forward_row(c1, k1, y, (int)r[j][1]);
            if (NC >= 4) {
                env_forward_row(c2, k2, y, (int)r[j][2]);
                env_forward_row(c3, k3, y, (int)r[j][3]);
            }
        }
    }

    // ---- backward: evaluate
    if (c0.q >= 0) env_load_top(c0, k0);
    if (NC >= 2 && c1.q >= 0) env_load_top(c1, k1);
    if (NC >= 4 && c2.q >= 0) env_load_top(c2, k2);
    if (NC >= 4 && c3.q >= 0) env_load_top(c3, k3);
    for (int y0 = H - 1; y0 >= 0; y0 -= ENV_AHEAD) {
        unsigned r[ENV_AHEAD][4];
#pragma unroll
        for (int j = 0; j < ENV_AHEAD; ++j) {
            r[j][0] = r[j][1] = r[j][2] = r[j][3] = 0u;
            if (y0 - j >= 0) grow_load<NC>(gbase + (size_t)(y0 - j) * W, r[j]);
        }
#pragma unroll
        for (int j = 0; j < ENV_AHEAD; ++j) {
            const int y = y0 - j;
            if (y < 0) break;
            int v[4] = {0, 0, 0, 0};
            if (!__all_sync(act, (r[j][0] | r[j][1] | r[j][2] | r[j][3]) == 0u)) {
                v[0] = env_backward_row(c0, k0, y, (int)r[j][0], cap);
                if (NC >= 2) v[1] = env_backward_row(c1, k1, y, (int)r[j][1], cap);
                if (NC >= 4) {
                    v[2] = env_backward_row(c2, k2, y, (int)r[j][2], cap);
                    v[3] = env_backward_row(c3, k3, y, (int)r[j][3], cap);
                }
            }
            int *dst = obase + (size_t)y * W;
            if (NC == 1) __stcs(dst, v[0]);
            else if (NC == 2) __stcs(reinterpret_cast<int2 *>(dst), make_int2(v[0], v[1]));
            else __stcs(reinterpret_cast<int4 *>(dst), make_int4(v[0], v[1], v[2], v[3]));
        }
    }
}

// one column per thread: measured fastest (1476 / 1926 / 3205 us for 1 / 2 / 4 columns on 320
// 1024^2 maps) -- the scan is latency bound, so warps in flight matter more than instruction count
constexpr int ENV_COLS = 1;

constexpr size_t EDT_CHUNK_BYTES = 2048ull << 20;   // the column pass needs thousands of columns in flight: big chunks

static int edt_chunk_maps(int M, int H, int W)
{
    const size_t per = sizeof(unsigned short) * (size_t)H * W;
    size_t c = per ? EDT_CHUNK_BYTES / per : (size_t)M;
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
    const size_t threads = (px + LD_PX - 1) / LD_PX;
    SLN_REQUIRE((threads + 255) / 256 < (1ull << 31), SLN_ERR_ARG, "image too large");
    layer_decode_kernel<<<dim3((unsigned)((threads + 255) / 256), B), 256, 0, st>>>(
        reinterpret_cast<const unsigned long long *>(label), px, L, n_max, out, scratch);
    SLN_LAUNCH_OK("layer_decode_kernel");
    layer_fixup_kernel<<<dim3(4 * sm_count(), B), 256, 0, st>>>(scratch, px, L, n_max, out, n_obj);
    SLN_LAUNCH_OK("layer_fixup_kernel");
    return SLN_OK;
}

extern "C" size_t sln_edt_workspace_bytes(int M, int H, int W)
{
    if (M <= 0 || H <= 0 || W <= 0) return 0;
    return sizeof(unsigned short) * (size_t)edt_chunk_maps(M, H, W) * H * W;
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
    unsigned short *g = static_cast<unsigned short *>(workspace);
    const int chunk = edt_chunk_maps(M, H, W);
    const int cap = (H + W) * (H + W);
    const bool vec = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15u) == 0) &&
                     ((reinterpret_cast<uintptr_t>(workspace) & 15u) == 0);
    SLN_REQUIRE(cdiv(H, 8) <= 65535, SLN_ERR_ARG, "H too large");
    for (int m0 = 0; m0 < M; m0 += chunk) {
        const int mc = (M - m0) < chunk ? (M - m0) : chunk;
        const long long n_rows = (long long)mc * H;
        const long long blocks = (n_rows + 7) / 8;
        SLN_REQUIRE(blocks < (1ll << 31) && mc <= 65535, SLN_ERR_ARG, "too many rows");
        edt_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(maps + (size_t)m0 * H * W, W, n_rows, g);
        SLN_LAUNCH_OK("edt_rows_kernel");
        const dim3 cgrid(cdiv(W, 128), cdiv(H, 8), mc);
        if (vec && W >= 128) {
            edt_cols_envelope_kernel<ENV_COLS><<<dim3(cdiv(W, 128 * ENV_COLS), mc), 128, 0, st>>>(g, H, W, cap, out + (size_t)m0 * H * W);
        } else if (vec) {
            edt_cols_kernel<true><<<cgrid, 256, 0, st>>>(g, H, W, cap, out + (size_t)m0 * H * W);
        } else {
            edt_cols_kernel<false><<<cgrid, 256, 0, st>>>(g, H, W, cap, out + (size_t)m0 * H * W);
        }
        SLN_LAUNCH_OK("edt_cols_kernel");
    }
    return SLN_OK;
}
