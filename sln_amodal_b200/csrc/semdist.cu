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
constexpr int LD_PX = 8;     // pixels per thread

__global__ void __launch_bounds__(256)
layer_presence_kernel(const unsigned long long *__restrict__ label, size_t px_per_image,
                      unsigned *__restrict__ top_seen)
{
    const int b = blockIdx.y;
    const unsigned long long *lab = label + (size_t)b * px_per_image;
    unsigned seen = 0u;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < px_per_image;
         p += (size_t)gridDim.x * blockDim.x) {
        const unsigned lo = (unsigned)(__ldg(lab + p) & 0xffffffffull);
        if (lo) seen |= 1u << (31 - __clz(lo));
    }
    seen = __reduce_or_sync(0xffffffffu, seen);
    if ((threadIdx.x & 31) == 0 && seen) atomicOr(top_seen + b, seen);   // OR: order-independent
}

__global__ void __launch_bounds__(256)
layer_decode_kernel(const unsigned long long *__restrict__ label, size_t px_per_image, int L, int n_max,
                    const unsigned *__restrict__ top_seen, unsigned char *__restrict__ out,
                    int *__restrict__ n_obj_out)
{
    const int b = blockIdx.y;
    const unsigned seen = top_seen[b];
    const int n_obj = __ffs(~seen) == 0 ? 32 : __ffs(~seen) - 1;     // max_objectID, Functions.py:1074-1079
    if (blockIdx.x == 0 && threadIdx.x == 0) n_obj_out[b] = n_obj;
    const int n_eff = min(n_obj, n_max);
    const size_t p0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * LD_PX;
    if (p0 >= px_per_image) return;
    const unsigned long long *lab = label + (size_t)b * px_per_image + p0;
    unsigned char *o = out + (size_t)b * n_max * L * px_per_image + p0;
    const bool full = (p0 + LD_PX <= px_per_image) && ((px_per_image % LD_PX) == 0);

    unsigned lo[LD_PX], hi[LD_PX];
#pragma unroll
    for (int k = 0; k < LD_PX; ++k) {
        unsigned long long v = 0ull;
        if (full || p0 + k < px_per_image) v = __ldg(lab + k);
        lo[k] = (unsigned)(v & 0xffffffffull);
        hi[k] = (unsigned)(v >> 32);
    }
    for (int i = 0; i < n_max; ++i) {
        // per pixel: channel hit by the occluded bit of object i (or -1)
        int ch[LD_PX];
        unsigned vis = 0u;
#pragma unroll
        for (int k = 0; k < LD_PX; ++k) {
            ch[k] = -1;
            if (i < n_eff) {
                vis |= ((lo[k] >> i) & 1u) << k;
                if ((hi[k] >> i) & 1u) ch[k] = min(1 + __popc(hi[k] & ((1u << i) - 1u)), L - 1);
            }
        }
        for (int l = 0; l < L; ++l) {
            unsigned long long bytes = 0ull;
#pragma unroll
            for (int k = 0; k < LD_PX; ++k) {
                const unsigned on = ((l == 0) && ((vis >> k) & 1u)) || (ch[k] == l);
                bytes |= (unsigned long long)on << (8 * k);
            }
            unsigned char *dst = o + ((size_t)i * L + l) * px_per_image;
            if (full) {
                __stcs(reinterpret_cast<unsigned long long *>(dst), bytes);
            } else {
                for (int k = 0; k < LD_PX && p0 + k < px_per_image; ++k) dst[k] = (unsigned char)(bytes >> (8 * k));
            }
        }
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
        if (x0 < W) {
            unsigned short d[ROW_SEG];
            int lz = left;
#pragma unroll
            for (int k = 0; k < ROW_SEG; ++k) {
                if ((z >> k) & 1u) lz = x0 + k;
                d[k] = (unsigned short)min(x0 + k - lz, (int)G_INF);
            }
            int rz = right;
#pragma unroll
            for (int k = ROW_SEG - 1; k >= 0; --k) {
                if ((z >> k) & 1u) rz = x0 + k;
                d[k] = (unsigned short)min((int)d[k], min(rz - (x0 + k), (int)G_INF));
            }
            if (vec_ok && x0 + ROW_SEG <= W) {
                uint4 *o = reinterpret_cast<uint4 *>(dst + x0);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint4 v;
                    v.x = d[8 * q + 0] | ((unsigned)d[8 * q + 1] << 16);
                    v.y = d[8 * q + 2] | ((unsigned)d[8 * q + 3] << 16);
                    v.z = d[8 * q + 4] | ((unsigned)d[8 * q + 5] << 16);
                    v.w = d[8 * q + 6] | ((unsigned)d[8 * q + 7] << 16);
                    o[q] = v;
                }
            } else {
#pragma unroll
                for (int k = 0; k < ROW_SEG; ++k)
                    if (x0 + k < W) dst[x0 + k] = d[k];
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

constexpr size_t EDT_CHUNK_BYTES = 48ull << 20;     // intermediate kept under ~half of L2

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
    SLN_CUDA_OK(cudaMemsetAsync(scratch, 0, sizeof(uint32_t) * (size_t)B, st));
    if (px == 0) {
        SLN_CUDA_OK(cudaMemsetAsync(n_obj, 0, sizeof(int) * (size_t)B, st));
        return SLN_OK;
    }
    SLN_REQUIRE(label && (out || n_max == 0), SLN_ERR_ARG, "null pointer");
    SLN_REQUIRE((reinterpret_cast<uintptr_t>(out) & 7u) == 0, SLN_ERR_LAYOUT, "out must be 8-byte aligned");
    int gx = (int)((px + 256 * 8 - 1) / (256 * 8));
    if (gx > 8 * sm_count()) gx = 8 * sm_count();
    layer_presence_kernel<<<dim3(gx, B), 256, 0, st>>>(reinterpret_cast<const unsigned long long *>(label), px, scratch);
    SLN_LAUNCH_OK("layer_presence_kernel");
    const size_t threads = (px + LD_PX - 1) / LD_PX;
    SLN_REQUIRE((threads + 255) / 256 < (1ull << 31), SLN_ERR_ARG, "image too large");
    layer_decode_kernel<<<dim3((unsigned)((threads + 255) / 256), B), 256, 0, st>>>(
        reinterpret_cast<const unsigned long long *>(label), px, L, n_max, scratch, out, n_obj);
    SLN_LAUNCH_OK("layer_decode_kernel");
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
        if (vec)
            edt_cols_kernel<true><<<cgrid, 256, 0, st>>>(g, H, W, cap, out + (size_t)m0 * H * W);
        else
            edt_cols_kernel<false><<<cgrid, 256, 0, st>>>(g, H, W, cap, out + (size_t)m0 * H * W);
        SLN_LAUNCH_OK("edt_cols_kernel");
    }
    return SLN_OK;
}
