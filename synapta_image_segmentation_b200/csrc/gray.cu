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

}  // namespace

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
    SS_TRY(validate_img(rgb, "rgb", 3));
    SS_TRY(validate_img(gray, "gray", 1));
    if (!same_shape(rgb, gray)) { synseg_set_error("synseg_rgb2gray: shape mismatch"); return SYNSEG_E_INVALID; }
    if (mode != SYNSEG_GRAY_CV && mode != SYNSEG_GRAY_PIL) { synseg_set_error("synseg_rgb2gray: bad mode %d", mode); return SYNSEG_E_INVALID; }
    return launch_rgb2gray(ctx, rgb, gray, mode, (cudaStream_t)stream);
}
