// gray.cu -- fixed-point RGB -> grey, bit-exact with cv2 (15-bit) and PIL (16-bit).
//
// Reference call sites: cv2.cvtColor(COLOR_RGB2GRAY) pdf_image_segmentation.py:1348;
// PIL convert('L') pdf_image_segmentation.py:1323,1549,1599,1699,1758,1804,2988,3072.
//
// Roofline: HBM-bound, 4 algorithmic bytes per pixel (3 read + 1 written).
// Layout: a thread converts 16 consecutive pixels of one row: 48 source bytes fetched as the
// 16-byte-aligned superset (3-4 LDG.128; the row base of a tightly packed RGB page is only
// 2-byte aligned: 3*2550 = 7650), realigned in registers with funnel shifts, reduced with two
// dp4a per pixel (coefficients split into high and low bytes), and stored as one STG.128.
#include "internal.cuh"
#include "pixel.cuh"

namespace {

// 12 bytes (w0,w1,w2) = 4 RGB pixels -> 4 grey bytes packed little-endian.
template <int MODE>
__device__ __forceinline__ uint32_t gray4(uint32_t w0, uint32_t w1, uint32_t w2)
{
    uint32_t y0 = gray1<MODE>(w0);
    uint32_t y1 = gray1<MODE>(__funnelshift_r(w0, w1, 24));
    uint32_t y2 = gray1<MODE>(__funnelshift_r(w1, w2, 16));
    uint32_t y3 = gray1<MODE>(w2 >> 8);
    return y0 | (y1 << 8) | (y2 << 16) | (y3 << 24);
}

template <int MODE>
__device__ __forceinline__ uint8_t gray_scalar(const uint8_t *p)
{
    return (uint8_t)gray1<MODE>((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16));
}

constexpr int GRAY_THREADS = 128;
constexpr int PX_PER_THREAD = 16;

template <int MODE, bool DST_ALIGNED>
__global__ void __launch_bounds__(GRAY_THREADS) rgb2gray_kernel(Plane src, Plane dst, int width, int height, int chunk_blocks)
{
    // blockIdx.x -> (global row, block of chunks inside the row)
    const int64_t grow = blockIdx.x / chunk_blocks;
    const int cb = blockIdx.x - (int)(grow * chunk_blocks);
    const int img = (int)(grow / height);
    const int y = (int)(grow - (int64_t)img * height);
    const int x0 = (cb * GRAY_THREADS + threadIdx.x) * PX_PER_THREAD;
    if (x0 >= width) return;
    const uint8_t *srow = src.p + img * src.bs + y * src.rs;
    uint8_t *drow = dst.p + img * dst.bs + y * dst.rs;
    const uint8_t *sp = srow + 3 * (int64_t)x0;

    if (x0 + PX_PER_THREAD <= width) {
        const uintptr_t a = (uintptr_t)sp;
        const uint4 *ap = (const uint4 *)(a & ~(uintptr_t)15);
        const int s = (int)(a & 15);
        uint32_t w[17];
        uint4 q0 = __ldg(ap), q1 = __ldg(ap + 1), q2 = __ldg(ap + 2);
        uint4 q3 = make_uint4(0, 0, 0, 0);
        if (s) q3 = __ldg(ap + 3);   // the 4th chunk holds valid bytes only when misaligned
        w[0] = q0.x; w[1] = q0.y; w[2] = q0.z; w[3] = q0.w;
        w[4] = q1.x; w[5] = q1.y; w[6] = q1.z; w[7] = q1.w;
        w[8] = q2.x; w[9] = q2.y; w[10] = q2.z; w[11] = q2.w;
        w[12] = q3.x; w[13] = q3.y; w[14] = q3.z; w[15] = q3.w; w[16] = 0;
        const int ws = s >> 2;          // word shift (uniform over the row)
        const int bs = (s & 3) * 8;     // bit shift
        uint32_t v[12];
        // ws is uniform across the block's row: resolve with a uniform switch so v[] stays in registers.
        switch (ws) {
        case 0:
#pragma unroll
            for (int i = 0; i < 12; ++i) v[i] = __funnelshift_r(w[i], w[i + 1], bs);
            break;
        case 1:
#pragma unroll
            for (int i = 0; i < 12; ++i) v[i] = __funnelshift_r(w[i + 1], w[i + 2], bs);
            break;
        case 2:
#pragma unroll
            for (int i = 0; i < 12; ++i) v[i] = __funnelshift_r(w[i + 2], w[i + 3], bs);
            break;
        default:
#pragma unroll
            for (int i = 0; i < 12; ++i) v[i] = __funnelshift_r(w[i + 3], w[i + 4], bs);
            break;
        }
        uint4 o;
        o.x = gray4<MODE>(v[0], v[1], v[2]);
        o.y = gray4<MODE>(v[3], v[4], v[5]);
        o.z = gray4<MODE>(v[6], v[7], v[8]);
        o.w = gray4<MODE>(v[9], v[10], v[11]);
        if (DST_ALIGNED) {
            *(uint4 *)(drow + x0) = o;
        } else {
            uint32_t ow[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
            for (int i = 0; i < 16; ++i) drow[x0 + i] = (uint8_t)(ow[i >> 2] >> ((i & 3) * 8));
        }
    } else {
        for (int x = x0; x < width; ++x) drow[x] = gray_scalar<MODE>(srow + 3 * (int64_t)x);
    }
}

// ---------------------------------------------------------------------------------------------
// Ragged crop front end: ONE pass over a packed batch of crops of different sizes produces
//   * the PIL grey of every crop in a canvas batch (what Canny reads; pdf_image_segmentation.py:1549),
//   * the exact grey moments sum / sum of squares / non-zero count (np.var at :1805, :2989, :3073),
//   * the HSV mask count S>30 & V>40 & V<240 of RGB crops (:1571-1577); RGB crops are 3 bytes per pixel or 4 (RGBX,
//     the layout PIL keeps in memory, so a host can hand its images over without repacking them),
// so the source (3 B/px) is read once instead of three times.  blockIdx.y = crop, a warp owns a row (four rows per
// warp and CTA), a lane 16
// consecutive pixels: 48 source bytes as the aligned superset (like rgb2gray), the moments of the 16 grey bytes with
// dp4a (sum: dot with 0x01010101, sum of squares: dot with itself), one 128-bit store into the canvas.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void load48(const uint8_t *sp, uint32_t (&v)[12])
{
    const uintptr_t a = (uintptr_t)sp;
    const uint4 *ap = (const uint4 *)(a & ~(uintptr_t)15);
    const int s = (int)(a & 15);
    const uint4 q0 = __ldg(ap), q1 = __ldg(ap + 1), q2 = __ldg(ap + 2);
    if (s == 0) {
        v[0] = q0.x; v[1] = q0.y; v[2] = q0.z; v[3] = q0.w; v[4] = q1.x; v[5] = q1.y; v[6] = q1.z; v[7] = q1.w;
        v[8] = q2.x; v[9] = q2.y; v[10] = q2.z; v[11] = q2.w;
        return;
    }
    const uint4 q3 = __ldg(ap + 3);        // holds valid bytes whenever the start is misaligned
    const uint32_t w[17] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w, 0u};
    const int ws = s >> 2, bs = (s & 3) * 8;
    switch (ws) {
    case 0:
#pragma unroll
        for (int i = 0; i < 12; ++i) v[i] = __funnelshift_r(w[i], w[i + 1], bs);
        break;
    case 1:
#pragma unroll
        for (int i = 0; i < 12; ++i) v[i] = __funnelshift_r(w[i + 1], w[i + 2], bs);
        break;
    case 2:
#pragma unroll
        for (int i = 0; i < 12; ++i) v[i] = __funnelshift_r(w[i + 2], w[i + 3], bs);
        break;
    default:
#pragma unroll
        for (int i = 0; i < 12; ++i) v[i] = __funnelshift_r(w[i + 3], w[i + 4], bs);
        break;
    }
}

// grey byte and HSV mask bit of the four pixels held by three words
__device__ __forceinline__ uint32_t gray_mask4(uint32_t w0, uint32_t w1, uint32_t w2, const uint32_t *sdiv, uint32_t &mask_cnt)
{
    const uint32_t px[4] = {w0, __funnelshift_r(w0, w1, 24), __funnelshift_r(w1, w2, 16), w2 >> 8};
    uint32_t out = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t r = px[j] & 0xFFu, g = (px[j] >> 8) & 0xFFu, b = (px[j] >> 16) & 0xFFu;
        mask_cnt += hsv_mask_px(r, g, b, sdiv) ? 1u : 0u;
        out |= gray1<SYNSEG_GRAY_PIL>(px[j]) << (8 * j);
    }
    return out;
}

// the same for four RGBX pixels (one word each; the fourth byte is ignored: its grey coefficient is zero)
__device__ __forceinline__ uint32_t gray_mask4x(uint32_t p0, uint32_t p1, uint32_t p2, uint32_t p3, const uint32_t *sdiv, uint32_t &mask_cnt)
{
    const uint32_t px[4] = {p0, p1, p2, p3};
    uint32_t out = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t r = px[j] & 0xFFu, g = (px[j] >> 8) & 0xFFu, b = (px[j] >> 16) & 0xFFu;
        mask_cnt += hsv_mask_px(r, g, b, sdiv) ? 1u : 0u;
        out |= gray1<SYNSEG_GRAY_PIL>(px[j]) << (8 * j);
    }
    return out;
}

__global__ void __launch_bounds__(256) crop_front_kernel(const uint8_t *base, const CropTask *tasks, Plane dst, unsigned long long *res)
{
    __shared__ uint32_t sdiv[256];
    __shared__ unsigned long long sh[8][4];
    const CropTask t = tasks[blockIdx.y];
    if ((int)blockIdx.x * 8 >= t.height) return;                     // the grid covers the tallest crop of the batch
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    sdiv[threadIdx.x] = hsv_sdiv(threadIdx.x);
    __syncthreads();
    const uint8_t *src = base + t.offset;
    uint8_t *dimg = dst.p + (int64_t)blockIdx.y * dst.bs;
    const int W = t.width;
    // The rows of this warp (y_first, y_first + ystep, ...) are cut into groups of 16 pixels and the (row, group) pairs
    // are dealt out to the lanes as ONE sequence: a 699-pixel row has 44 groups, which would fill 32 + 12 lanes row by
    // row, but four such rows fill 5.5 rounds of 32 lanes.  The partial last group of a row runs in the same round as
    // the full groups of its neighbours: only its loads differ (pixel by pixel, nothing beyond column W-1 is touched;
    // pixels outside the row enter as black: grey 0, not masked in, nothing added to the moments).
    const int G = (W + 15) >> 4;
    const int ystep = gridDim.x * 8, y_first = blockIdx.x * 8 + warp;
    const int nrows = y_first < t.height ? (t.height - 1 - y_first) / ystep + 1 : 0;
    const int total = nrows * G;
    int r = 0, gi = lane;
    while (gi >= G) { gi -= G; ++r; }
    uint32_t rs = 0, rss = 0, rnz = 0, rmk = 0;     // per lane: <= (4 rows * 2048 groups / 32 + 1) * 16 px * 65025 fits 32 bits
    for (int q = lane; q < total; q += 32) {
        const int y = y_first + r * ystep, x0 = gi << 4;
        const uint8_t *srow = src + (int64_t)y * t.row_stride;
        uint8_t *drow = dimg + (int64_t)y * dst.rs;
        const bool full = x0 + 16 <= W;
        uint4 o;
        if (t.channels >= 3) {                                       // warp-uniform (one crop per CTA)
            uint32_t v[16];                                          // one word per pixel: R | G << 8 | B << 16 | (ignored) << 24
            if (t.channels == 4) {                                   // RGBX words (4-byte aligned by contract)
                const uint32_t *wp = (const uint32_t *)(srow + 4 * (int64_t)x0);
                if (full && ((uintptr_t)wp & 15) == 0) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) { const uint4 u = __ldg((const uint4 *)wp + k); v[4 * k] = u.x; v[4 * k + 1] = u.y; v[4 * k + 2] = u.z; v[4 * k + 3] = u.w; }
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = (x0 + i < W) ? __ldg(wp + i) : 0u;
                }
            } else if (full) {
                uint32_t w[12];
                load48(srow + 3 * (int64_t)x0, w);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    v[4 * k] = w[3 * k]; v[4 * k + 1] = __funnelshift_r(w[3 * k], w[3 * k + 1], 24);
                    v[4 * k + 2] = __funnelshift_r(w[3 * k + 1], w[3 * k + 2], 16); v[4 * k + 3] = w[3 * k + 2] >> 8;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    v[i] = 0u;
                    if (x0 + i < W) {
                        const uint8_t *p = srow + 3 * (int64_t)(x0 + i);
                        v[i] = (uint32_t)__ldg(p) | ((uint32_t)__ldg(p + 1) << 8) | ((uint32_t)__ldg(p + 2) << 16);
                    }
                }
            }
            o.x = gray_mask4x(v[0], v[1], v[2], v[3], sdiv, rmk); o.y = gray_mask4x(v[4], v[5], v[6], v[7], sdiv, rmk);
            o.z = gray_mask4x(v[8], v[9], v[10], v[11], sdiv, rmk); o.w = gray_mask4x(v[12], v[13], v[14], v[15], sdiv, rmk);
        } else {
            const uint8_t *sp = srow + x0;
            if (full && ((uintptr_t)sp & 15) == 0) o = __ldg((const uint4 *)sp);
            else {
                uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
                for (int i = 0; i < 16; ++i) w[i >> 2] |= ((x0 + i < W) ? (uint32_t)__ldg(sp + i) : 0u) << (8 * (i & 3));
                o = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
        *(uint4 *)(drow + x0) = o;                                   // canvas rows are 16-byte aligned and padded; columns >= W are never read
        rs = __dp4a(o.x, 0x01010101u, rs); rs = __dp4a(o.y, 0x01010101u, rs); rs = __dp4a(o.z, 0x01010101u, rs); rs = __dp4a(o.w, 0x01010101u, rs);
        rss = __dp4a(o.x, o.x, rss); rss = __dp4a(o.y, o.y, rss); rss = __dp4a(o.z, o.z, rss); rss = __dp4a(o.w, o.w, rss);
        rnz += __popc(__vcmpne4(o.x, 0u) & 0x01010101u) + __popc(__vcmpne4(o.y, 0u) & 0x01010101u) +
               __popc(__vcmpne4(o.z, 0u) & 0x01010101u) + __popc(__vcmpne4(o.w, 0u) & 0x01010101u);
        gi += 32;
        while (gi >= G) { gi -= G; ++r; }
    }
    unsigned long long s = rs, ss = rss, nz = rnz, mk = rmk;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        s += __shfl_down_sync(0xffffffffu, s, d); ss += __shfl_down_sync(0xffffffffu, ss, d);
        nz += __shfl_down_sync(0xffffffffu, nz, d); mk += __shfl_down_sync(0xffffffffu, mk, d);
    }
    if (lane == 0) { sh[warp][0] = s; sh[warp][1] = ss; sh[warp][2] = nz; sh[warp][3] = mk; }
    __syncthreads();
    if (threadIdx.x < 4) {
        unsigned long long tot = 0;
        for (int w = 0; w < 8; ++w) tot += sh[w][threadIdx.x];
        if (tot) atomicAdd(res + 8 * (int64_t)blockIdx.y + 3 + threadIdx.x, tot);
    }
}

}  // namespace

int launch_crop_front(synseg_ctx *ctx, const void *base, const CropTask *tasks, int n, const synseg_img *gray_canvas, uint64_t *res,
                      cudaStream_t st)
{
    if (!plane_aligned(gray_canvas, 16)) { synseg_set_error("crop_front: the grey canvas must be 16-byte aligned"); return SYNSEG_E_INVALID; }
    crop_front_kernel<<<dim3(cdiv(gray_canvas->height, 32), n), 256, 0, st>>>((const uint8_t *)base, tasks, plane_of(gray_canvas),
                                                                            (unsigned long long *)res);
    SS_LAUNCH_CHECK(ctx, "crop_front", st);
    return SYNSEG_OK;
}

int launch_rgb2gray(synseg_ctx *ctx, const synseg_img *rgb, const synseg_img *gray, int mode, cudaStream_t st)
{
    const int width = rgb->width, height = rgb->height, batch = rgb->batch;
    const int chunks = cdiv(width, PX_PER_THREAD);
    const int chunk_blocks = cdiv(chunks, GRAY_THREADS);
    const int64_t nblocks = (int64_t)batch * height * chunk_blocks;
    if (nblocks > 0x7fffffffLL) { synseg_set_error("rgb2gray: batch too large"); return SYNSEG_E_INVALID; }
    Plane s = plane_of(rgb), d = plane_of(gray);
    const bool al = plane_aligned(gray, 16);
    dim3 grid((unsigned)nblocks), block(GRAY_THREADS);
    if (mode == SYNSEG_GRAY_CV) {
        if (al) rgb2gray_kernel<SYNSEG_GRAY_CV, true><<<grid, block, 0, st>>>(s, d, width, height, chunk_blocks);
        else rgb2gray_kernel<SYNSEG_GRAY_CV, false><<<grid, block, 0, st>>>(s, d, width, height, chunk_blocks);
    } else {
        if (al) rgb2gray_kernel<SYNSEG_GRAY_PIL, true><<<grid, block, 0, st>>>(s, d, width, height, chunk_blocks);
        else rgb2gray_kernel<SYNSEG_GRAY_PIL, false><<<grid, block, 0, st>>>(s, d, width, height, chunk_blocks);
    }
    SS_LAUNCH_CHECK(ctx, "rgb2gray", st);
    return SYNSEG_OK;
}

extern "C" SYNSEG_EXPORT int synseg_rgb2gray(synseg_ctx *ctx, const synseg_img *rgb, const synseg_img *gray, int mode, void *stream)
{
    if (!ctx) { synseg_set_error("synseg_rgb2gray: ctx is NULL"); return SYNSEG_E_INVALID; }
    SS_ENTER(ctx, stream);
    SS_TRY(validate_img(rgb, "rgb", 3));
    SS_TRY(validate_img(gray, "gray", 1));
    if (!same_shape(rgb, gray)) { synseg_set_error("synseg_rgb2gray: shape mismatch"); return SYNSEG_E_INVALID; }
    if (mode != SYNSEG_GRAY_CV && mode != SYNSEG_GRAY_PIL) { synseg_set_error("synseg_rgb2gray: bad mode %d", mode); return SYNSEG_E_INVALID; }
    return launch_rgb2gray(ctx, rgb, gray, mode, (cudaStream_t)stream);
}
