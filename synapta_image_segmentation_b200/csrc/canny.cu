// canny.cu -- cv2.Canny(g, lo, hi) (aperture 3, L1 gradient), bit-exact.
//
// Reference call sites: pdf_image_segmentation.py:1324,1366,1550,1600,1700,1759 (all Canny(.,50,150)).
// Algorithm (OpenCV 4.13 semantics):
//   Sobel 3x3 on the replicate-bordered image -> mag = |dx|+|dy| with a ring of zeros outside the
//   image -> for mag > lo, integer non-maximum suppression along the quantised gradient direction
//   (TG22 = 13573 = tan(22.5 deg) << 15) -> kept pixels with mag > hi are strong -> the output is
//   every kept pixel 8-connected (through kept pixels) to a strong one.
// Stage 1 (this file): one tile kernel, grey tile + 2-pixel halo staged in shared memory, Sobel and
//   magnitude for tile + 1 halo kept in shared memory, NMS from shared memory; the result is two bit
//   planes (warp ballots): kept = survived NMS, strong = kept and mag > hi.
// Stage 2 (ccl.cu): hysteresis = run-based union-find over the kept pixels + "component holds a strong
//   pixel" flag; no host round trip, no iteration count that depends on the image.
//
// Roofline: HBM-bound, 1.25 algorithmic bytes per pixel (1 read + two bit planes written) for stage 1.
#include "internal.cuh"

namespace {

constexpr int TW = 128, TH = 16;           // output tile
constexpr int GP = TW + 8;                 // grey tile pitch (TW + 4 used)
constexpr int MP = TW + 4;                 // magnitude / gradient tile pitch (TW + 2 used)

__global__ void __launch_bounds__(256) canny_classes_kernel(Plane src, BitPlane kept, BitPlane strong, int width, int height, int lo, int hi)
{
    __shared__ uint8_t g[(TH + 4) * GP];
    __shared__ uint16_t mag[(TH + 2) * MP];
    __shared__ uint32_t dxy[(TH + 2) * MP];   // dx (low 16, signed) | dy (high 16, signed)

    const int img = blockIdx.z;
    const int tx0 = blockIdx.x * TW, ty0 = blockIdx.y * TH;
    const uint8_t *base = src.p + img * src.bs;
    const int tid = threadIdx.x;

    // grey tile rows ty0-2 .. ty0+TH+1, cols tx0-2 .. tx0+TW+1 (replicate)
    for (int i = tid; i < (TH + 4) * (TW + 4); i += 256) {
        const int ly = i / (TW + 4), lx = i - ly * (TW + 4);
        const int gy = min(max(ty0 + ly - 2, 0), height - 1);
        const int gx = min(max(tx0 + lx - 2, 0), width - 1);
        g[ly * GP + lx] = __ldg(base + gy * src.rs + gx);
    }
    __syncthreads();

    // gradient for rows ty0-1 .. ty0+TH, cols tx0-1 .. tx0+TW
    for (int i = tid; i < (TH + 2) * (TW + 2); i += 256) {
        const int ly = i / (TW + 2), lx = i - ly * (TW + 2);
        const int gy = ty0 + ly - 1, gx = tx0 + lx - 1;
        const uint8_t *c = g + (ly + 1) * GP + (lx + 1);   // centre in the grey tile
        const int a00 = c[-GP - 1], a01 = c[-GP], a02 = c[-GP + 1];
        const int a10 = c[-1], a12 = c[1];
        const int a20 = c[GP - 1], a21 = c[GP], a22 = c[GP + 1];
        const int dx = (a02 + 2 * a12 + a22) - (a00 + 2 * a10 + a20);
        const int dy = (a20 + 2 * a21 + a22) - (a00 + 2 * a01 + a02);
        const bool in = gy >= 0 && gy < height && gx >= 0 && gx < width;
        mag[ly * MP + lx] = in ? (uint16_t)(abs(dx) + abs(dy)) : (uint16_t)0;
        dxy[ly * MP + lx] = ((uint32_t)dx & 0xffffu) | ((uint32_t)dy << 16);
    }
    __syncthreads();

    const int lx = tid & (TW - 1);
    const int gx = tx0 + lx;
    if (tx0 + (lx & ~31) >= width) return;          // whole warp outside the image (warp-uniform)
    for (int ly = tid >> 7; ly < TH; ly += 2) {
        const int gy = ty0 + ly;
        if (gy >= height) break;
        const uint16_t *mc = mag + (ly + 1) * MP + (lx + 1);
        const int m = mc[0];
        int c = 0;
        if (m > lo && gx < width) {
            const uint32_t d = dxy[(ly + 1) * MP + (lx + 1)];
            const int xs = (int)(short)(d & 0xffffu), ys = (int)(short)(d >> 16);
            const int ax = abs(xs), ay = abs(ys) << 15;
            const int tg22x = ax * 13573;
            bool keep;
            if (ay < tg22x) keep = m > mc[-1] && m >= mc[1];
            else {
                const int tg67x = tg22x + (ax << 16);
                if (ay > tg67x) keep = m > mc[-MP] && m >= mc[MP];
                else {
                    const int s = ((xs ^ ys) < 0) ? -1 : 1;
                    keep = m > mc[-MP - s] && m > mc[MP + s];
                }
            }
            if (keep) c = (m > hi) ? 2 : 1;
        }
        const uint32_t kw = __ballot_sync(0xffffffffu, c != 0), sw = __ballot_sync(0xffffffffu, c == 2);
        if ((tid & 31) == 0) {
            const int64_t o = (int64_t)gy * kept.wpr + (gx >> 5);
            kept.p[img * kept.bs + o] = kw;
            strong.p[img * strong.bs + o] = sw;
        }
    }
}

}  // namespace

int launch_canny_classes(synseg_ctx *ctx, const synseg_img *gray, BitPlane kept, BitPlane strong, int lo, int hi, cudaStream_t st)
{
    dim3 grid(cdiv(gray->width, TW), cdiv(gray->height, TH), gray->batch);
    canny_classes_kernel<<<grid, 256, 0, st>>>(plane_of(gray), kept, strong, gray->width, gray->height, lo, hi);
    SS_LAUNCH_CHECK(ctx, "canny_classes", st);
    return SYNSEG_OK;
}

size_t canny_scratch_bytes(int width, int height, int batch)
{
    return 2 * ((size_t)bit_wpr(width) * height * batch * 4 + 256) + hysteresis_scratch_bytes(width, height, batch) + 256;
}

int run_canny(synseg_ctx *ctx, const synseg_img *gray, const synseg_img *edges_u8, BitPlane edges_bits, bool or_bits,
              int lo, int hi, cudaStream_t st)
{
    const int W = gray->width, H = gray->height, B = gray->batch;
    const int wpr = bit_wpr(W);
    const size_t plane_bytes = (size_t)wpr * H * B * 4;
    void *p;
    SS_TRY(arena_alloc(ctx, plane_bytes, &p, st));
    BitPlane kept{(uint32_t *)p, wpr, (int64_t)wpr * H};
    SS_TRY(arena_alloc(ctx, plane_bytes, &p, st));
    BitPlane strong{(uint32_t *)p, wpr, (int64_t)wpr * H};
    SS_TRY(launch_canny_classes(ctx, gray, kept, strong, lo, hi, st));
    return run_hysteresis(ctx, kept, strong, W, H, B, edges_u8, edges_bits, or_bits, st);
}

extern "C" SYNSEG_EXPORT int synseg_canny(synseg_ctx *ctx, const synseg_img *gray, const synseg_img *edges, int lo, int hi, void *stream)
{
    if (!ctx) { synseg_set_error("synseg_canny: ctx is NULL"); return SYNSEG_E_INVALID; }
    SS_TRY(validate_img(gray, "gray", 1));
    SS_TRY(validate_img(edges, "edges", 1));
    if (!same_shape(gray, edges)) { synseg_set_error("synseg_canny: shape mismatch"); return SYNSEG_E_INVALID; }
    if (lo < 0 || hi < lo) { synseg_set_error("synseg_canny: need 0 <= lo <= hi"); return SYNSEG_E_INVALID; }
    if (gray->batch > 65535) { synseg_set_error("synseg_canny: batch > 65535"); return SYNSEG_E_INVALID; }
    SS_TRY(arena_ensure(ctx, canny_scratch_bytes(gray->width, gray->height, gray->batch)));
    arena_begin(ctx);
    return run_canny(ctx, gray, edges, BitPlane{nullptr, 0, 0}, false, lo, hi, (cudaStream_t)stream);
}
