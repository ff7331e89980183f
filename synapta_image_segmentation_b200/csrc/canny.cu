// canny.cu -- cv2.Canny(g, lo, hi) (aperture 3, L1 gradient), bit-exact.
//
// Reference call sites: pdf_image_segmentation.py:1324,1366,1550,1600,1700,1759 (all Canny(.,50,150)).
// Algorithm (OpenCV 4.13 semantics):
//   Sobel 3x3 on the replicate-bordered image -> mag = |dx|+|dy| with a ring of zeros outside the
//   image -> for mag > lo, integer non-maximum suppression along the quantised gradient direction
//   (TG22 = 13573 = tan(22.5 deg) << 15) -> kept pixels with mag > hi are strong -> the output is
//   every kept pixel 8-connected (through kept pixels) to a strong one.
// Stage 1 (this file): warp-autonomous strips, no block-level barrier.  A warp owns a strip of 512 columns
//   (16 per lane, one 128-bit load per lane and row; lanes 0 and 31 are halo lanes, 480 output columns) and
//   marches down a band of rows with a three-row grey ring in registers.  Each step it computes one row of
//   Sobel magnitudes with dp4a on byte windows, stores mag | direction << 12 (direction only where
//   mag > lo) into a three-row, transposed, conflict-free shared-memory ring, and runs the integer
//   non-maximum suppression of the previous row for the candidate pixels only.  The result is two bit
//   planes: kept = survived NMS, strong = kept and mag > hi (two lanes -> one 32-bit word).
// Stage 2 (ccl.cu): hysteresis = run-based union-find over the kept pixels + "component holds a strong
//   pixel" flag; no host round trip, no iteration count that depends on the image.
//
// Roofline: HBM-bound, 1.25 algorithmic bytes per pixel (1 read + two bit planes written) for stage 1.
#include "internal.cuh"
#include "pixel.cuh"

namespace {

constexpr int CN_OUT_W = 480;           // output columns per warp strip (lanes 1..30 x 16)
constexpr int CN_WARPS = 4;             // independent warps per CTA
constexpr unsigned FULL = 0xffffffffu;

struct CnParams {
    Plane src;
    BitPlane kept, strong;
    int width, height, lo, hi;
    int strips, bands, band_h;
    int64_t tasks;
};

struct GRow { uint32_t w[6]; };         // [left neighbour's last word, own 4 words, right neighbour's first word]

__device__ __forceinline__ GRow load_grow(const uint8_t *base, int64_t rs, int y, int H, int x, int W, bool aligned, bool live)
{
    GRow g;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (live) v = load16_rep(base + (int64_t)min(max(y, 0), H - 1) * rs, x, W, aligned);
    g.w[1] = v.x; g.w[2] = v.y; g.w[3] = v.z; g.w[4] = v.w;
    g.w[0] = __shfl_up_sync(FULL, v.w, 1);
    g.w[5] = __shfl_down_sync(FULL, v.x, 1);
    return g;
}

__device__ __forceinline__ int dp4a_us(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// bytes (x-1, x, x+1, x+2) of column j of a row
template <int J>
__device__ __forceinline__ uint32_t window(const GRow &g)
{
    constexpr int B = 3 + J;            // byte offset in the 24-byte array
    return __funnelshift_r(g.w[B >> 2], g.w[(B >> 2) + 1], 8 * (B & 3));
}

// Sobel magnitude (+ direction where mag > lo) of column J of the middle row b; returns mag | dir << 12
template <int J>
__device__ __forceinline__ uint32_t mag_dir(const GRow &a, const GRow &b, const GRow &c, int lo)
{
    const uint32_t wa = window<J>(a), wb = window<J>(b), wc = window<J>(c);
    int dx = dp4a_us(wa, 0x000100FFu, 0);            // (-1, 0, 1, 0)
    dx = dp4a_us(wb, 0x000200FEu, dx);               // (-2, 0, 2, 0)
    dx = dp4a_us(wc, 0x000100FFu, dx);
    int dy = dp4a_us(wc, 0x00010201u, 0);            // ( 1, 2, 1, 0)
    dy = dp4a_us(wa, 0x00FFFEFFu, dy);               // (-1,-2,-1, 0)
    const int ax = abs(dx), ayv = abs(dy);
    const uint32_t m = (uint32_t)(ax + ayv);
    uint32_t dir = 0;
    if ((int)m > lo) {
        const int ay = ayv << 15, tg22 = ax * 13573;
        if (ay >= tg22) {
            const int tg67 = tg22 + (ax << 16);
            dir = (ay > tg67) ? 1u : (((dx ^ dy) < 0) ? 3u : 2u);
        }
    }
    return m | (dir << 12);
}

struct MagRing { uint16_t v[3][16][32]; };   // [ring row][column within lane][lane]

// magnitude of column j (-1..16) of `lane` in ring row `slot`
__device__ __forceinline__ int ring_mag(const MagRing &R, int slot, int lane, int j)
{
    if (j < 0) { j = 15; --lane; } else if (j > 15) { j = 0; ++lane; }
    return R.v[slot][j][lane] & 0xFFF;
}

template <int J>
__device__ __forceinline__ void mag_step(MagRing &R, int slot, int lane, const GRow &a, const GRow &b, const GRow &c, int lo,
                                         int x, int W, bool row_in, uint32_t &cand)
{
    uint32_t v = 0;
    if (row_in && x + J < W && x + J >= 0) v = mag_dir<J>(a, b, c, lo);
    R.v[slot][J][lane] = (uint16_t)v;
    if ((int)(v & 0xFFFu) > lo) cand |= 1u << J;
    if constexpr (J < 15) mag_step<J + 1>(R, slot, lane, a, b, c, lo, x, W, row_in, cand);
}

__global__ void __launch_bounds__(32 * CN_WARPS) canny_classes_kernel(CnParams p, bool aligned)
{
    __shared__ MagRing rings[CN_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t task = (int64_t)blockIdx.x * CN_WARPS + warp;
    if (task >= p.tasks) return;                      // warp-uniform
    const int strip = (int)(task % p.strips); task /= p.strips;
    const int band = (int)(task % p.bands);
    const int img = (int)(task / p.bands);
    MagRing &R = rings[warp];

    const int W = p.width, H = p.height, lo = p.lo, hi = p.hi;
    const int x = strip * CN_OUT_W - 16 + 16 * lane;  // first of this lane's 16 columns
    const int y0 = band * p.band_h, y1 = min(y0 + p.band_h, H);
    const uint8_t *base = p.src.p + img * p.src.bs;
    const int64_t rs = p.src.rs;
    const bool live = x <= W;                          // lanes further right feed no output
    const bool out_lane = lane >= 1 && lane <= 30 && x < W;

    // prime: magnitude rows y0-1 and y0
    GRow g0 = load_grow(base, rs, y0 - 2, H, x, W, aligned, live);
    GRow g1 = load_grow(base, rs, y0 - 1, H, x, W, aligned, live);
    GRow g2 = load_grow(base, rs, y0, H, x, W, aligned, live);
    uint32_t cand_prev = 0, cand_cur = 0, cand_next = 0;
    mag_step<0>(R, (y0 + 2) % 3, lane, g0, g1, g2, lo, x, W, y0 - 1 >= 0, cand_prev);   // row y0-1 -> slot (y0-1) mod 3
    g0 = g1; g1 = g2; g2 = load_grow(base, rs, y0 + 1, H, x, W, aligned, live);
    mag_step<0>(R, y0 % 3, lane, g0, g1, g2, lo, x, W, true, cand_cur);
    (void)cand_prev;

    for (int y = y0; y < y1; ++y) {
        // magnitude row y+1 from grey rows y, y+1, y+2
        g0 = g1; g1 = g2; g2 = load_grow(base, rs, y + 2, H, x, W, aligned, live);
        cand_next = 0;
        mag_step<0>(R, (y + 1) % 3, lane, g0, g1, g2, lo, x, W, y + 1 < H, cand_next);
        __syncwarp();

        // non-maximum suppression of row y, candidates only
        uint32_t kept16 = 0, strong16 = 0;
        if (out_lane) {
            const int sc = y % 3, su = (y + 2) % 3, sd = (y + 1) % 3;
            uint32_t c = cand_cur;
            while (c) {
                const int j = __ffs((int)c) - 1;
                c &= c - 1;
                const uint32_t v = R.v[sc][j][lane];
                const int m = (int)(v & 0xFFFu), dir = (int)(v >> 12);
                bool keep;
                if (dir == 0) keep = m > ring_mag(R, sc, lane, j - 1) && m >= ring_mag(R, sc, lane, j + 1);
                else if (dir == 1) keep = m > ring_mag(R, su, lane, j) && m >= ring_mag(R, sd, lane, j);
                else if (dir == 2) keep = m > ring_mag(R, su, lane, j - 1) && m > ring_mag(R, sd, lane, j + 1);
                else keep = m > ring_mag(R, su, lane, j + 1) && m > ring_mag(R, sd, lane, j - 1);
                if (keep) { kept16 |= 1u << j; if (m > hi) strong16 |= 1u << j; }
            }
        }
        const uint32_t k_up = __shfl_down_sync(FULL, kept16, 1), s_up = __shfl_down_sync(FULL, strong16, 1);
        if (out_lane && (lane & 1) && lane <= 29) {
            const int64_t o = (int64_t)y * p.kept.wpr + (x >> 5);
            p.kept.p[img * p.kept.bs + o] = kept16 | (k_up << 16);
            p.strong.p[img * p.strong.bs + o] = strong16 | (s_up << 16);
        }
        cand_cur = cand_next;
        __syncwarp();          // ring row (y+2) mod 3 == (y-1) mod 3 is overwritten next iteration
    }
}

}  // namespace

int launch_canny_classes(synseg_ctx *ctx, const synseg_img *gray, BitPlane kept, BitPlane strong, int lo, int hi, cudaStream_t st)
{
    CnParams p;
    p.src = plane_of(gray);
    p.kept = kept; p.strong = strong;
    p.width = gray->width; p.height = gray->height; p.lo = lo; p.hi = hi;
    p.strips = cdiv(gray->width, CN_OUT_W);
    // bands: about 24 resident warps per SM; a band recomputes 2 magnitude rows of its neighbours
    const int64_t rows_total = (int64_t)gray->height * gray->batch * p.strips;
    int band_h = (int)(rows_total / (24 * (int64_t)ctx->sm_count));
    band_h = band_h < 32 ? 32 : (band_h > 128 ? 128 : band_h);
    if (ctx->tune_canny_band > 0) band_h = ctx->tune_canny_band;
    if (band_h > gray->height) band_h = gray->height;
    p.band_h = band_h;
    p.bands = cdiv(gray->height, band_h);
    p.tasks = (int64_t)gray->batch * p.bands * p.strips;
    canny_classes_kernel<<<(unsigned)cdiv(p.tasks, CN_WARPS), 32 * CN_WARPS, 0, st>>>(p, plane_aligned(gray, 16));
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
