// threshold.cu -- cv2.adaptiveThreshold(ADAPTIVE_THRESH_MEAN_C), bit-exact.
//
// No reference call site (north-star primitive); oracle = cv2 4.13.0:
//   mean = round(boxsum / bs^2) over BORDER_REPLICATE  (bs odd -> never a tie, so
//          mean = floor((2s + n) / 2n) exactly),
//   THRESH_BINARY: 255 iff g - mean > -C ;  THRESH_BINARY_INV: 255 iff g - mean <= -C.
//
// Roofline: HBM-bound, 2 algorithmic bytes per pixel (1 read + 1 written).
// Design: a CTA owns a strip of columns and a band of rows and marches down the band keeping the
// vertical running column sums (new row in, old row out) in registers, 4 columns per thread.
// Each row the horizontal window sums come from one block-wide inclusive scan of the column sums
// (local 4 + warp shuffle scan + cross-warp redux) parked in shared memory: S = P[x+r] - P[x-r-1].
// Every source byte is fetched from HBM once (the second, "row out", read hits L2); no integral
// image is ever written.  Output is either a u8 {0,255} plane or a bit plane (8 lanes -> 1 word).
#include "internal.cuh"

namespace {

struct AdParams {
    Plane src;
    Plane dst;       // u8 output (OUT_BITS = false)
    BitPlane bits;   // bit output (OUT_BITS = true)
    int width, height;
    int r;           // bs / 2
    int lead;        // round_up(r, 32): columns of left halo the strip carries
    int out_w;       // output columns per strip (multiple of 32)
    int strips;
    int bands;
    int band_h;
    int C;
    int invert;
    uint32_t n;      // bs * bs
    uint64_t magic;  // ceil(2^48 / 2n)
};

__device__ __forceinline__ uint32_t load4_clamped(const uint8_t *row, int cx, int width, bool aligned)
{
    if (aligned && cx >= 0 && cx + 3 < width) return __ldg((const uint32_t *)(row + cx));
    uint32_t v = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int c = min(max(cx + j, 0), width - 1);
        v |= (uint32_t)__ldg(row + c) << (8 * j);
    }
    return v;
}

template <bool OUT_BITS>
__global__ void __launch_bounds__(1024) adaptive_mean_kernel(AdParams p, bool src_aligned, bool dst_aligned)
{
    extern __shared__ uint32_t smem[];
    const int T = blockDim.x;
    const int ncols = 4 * T;
    uint32_t *P[2] = {smem, smem + (ncols + 1)};             // inclusive prefix with P[0] = 0
    uint32_t *wt[2] = {smem + 2 * (ncols + 1), smem + 2 * (ncols + 1) + 32};

    int bid = blockIdx.x;
    const int band = bid % p.bands; bid /= p.bands;
    const int strip = bid % p.strips;
    const int img = bid / p.strips;

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int x_strip = strip * p.out_w;
    const int cx = x_strip - p.lead + 4 * t;                 // first of this thread's 4 columns
    const int y0 = band * p.band_h;
    const int y1 = min(y0 + p.band_h, p.height);
    const uint8_t *base = p.src.p + img * p.src.bs;
    const int r = p.r, H = p.height, W = p.width;

    if (t == 0) { P[0][0] = 0; P[1][0] = 0; }

    // running column sums over rows [y - r, y + r] (replicate)
    uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    for (int dy = -r; dy <= r; ++dy) {
        int yy = min(max(y0 + dy, 0), H - 1);
        uint32_t v = load4_clamped(base + yy * p.src.rs, cx, W, src_aligned);
        c0 += v & 255; c1 += (v >> 8) & 255; c2 += (v >> 16) & 255; c3 += v >> 24;
    }

    const bool out_thread = (4 * t >= p.lead) && (4 * t < p.lead + p.out_w) && (cx < W);
    const int li = 4 * t;   // local column index of c0

    for (int y = y0; y < y1; ++y) {
        const int buf = y & 1;
        // issue next iteration's loads early
        const int yn = min(y + r + 1, H - 1), yo = max(y - r, 0);
        uint32_t vnew = load4_clamped(base + yn * p.src.rs, cx, W, src_aligned);
        uint32_t vold = load4_clamped(base + yo * p.src.rs, cx, W, src_aligned);
        uint32_t vcen = 0;
        if (out_thread) vcen = load4_clamped(base + (int64_t)y * p.src.rs, cx, W, src_aligned);

        uint32_t a0 = c0, a1 = a0 + c1, a2 = a1 + c2, a3 = a2 + c3;
        uint32_t v = a3;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t nb = __shfl_up_sync(0xffffffffu, v, d);
            if (lane >= d) v += nb;
        }
        if (lane == 31) wt[buf][warp] = v;
        __syncthreads();
        uint32_t w = (lane < warp) ? wt[buf][lane] : 0u;
        w = __reduce_add_sync(0xffffffffu, w);
        const uint32_t off = w + v - a3;
        uint32_t *Pp = P[buf] + 1 + li;
        Pp[0] = off + a0; Pp[1] = off + a1; Pp[2] = off + a2; Pp[3] = off + a3;
        __syncthreads();

        uint32_t nib = 0;
        if (out_thread) {
            const uint32_t *Pq = P[buf];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t s = Pq[li + j + r + 1] - Pq[li + j - r];
                uint32_t num = 2 * s + p.n;
                int mean = (int)__umul64hi((uint64_t)num << 16, p.magic);
                int g = (vcen >> (8 * j)) & 255;
                int diff = g - mean;
                bool on = p.invert ? (diff <= -p.C) : (diff > -p.C);
                if (cx + j < W && on) nib |= 1u << j;
            }
        }
        if (OUT_BITS) {
            uint32_t wv = nib << (4 * (lane & 7));
            wv |= __shfl_xor_sync(0xffffffffu, wv, 1);
            wv |= __shfl_xor_sync(0xffffffffu, wv, 2);
            wv |= __shfl_xor_sync(0xffffffffu, wv, 4);
            if (out_thread && (lane & 7) == 0) p.bits.p[img * p.bits.bs + (int64_t)y * p.bits.wpr + (cx >> 5)] = wv;
        } else if (out_thread) {
            uint8_t *drow = p.dst.p + img * p.dst.bs + y * p.dst.rs;
            uint32_t bytes = ((nib * 0x00204081u) & 0x01010101u) * 0xFFu;
            if (dst_aligned && cx + 3 < W) *(uint32_t *)(drow + cx) = bytes;
            else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (cx + j < W) drow[cx + j] = (uint8_t)(bytes >> (8 * j));
            }
        }

        c0 += (vnew & 255) - (vold & 255);
        c1 += ((vnew >> 8) & 255) - ((vold >> 8) & 255);
        c2 += ((vnew >> 16) & 255) - ((vold >> 16) & 255);
        c3 += (vnew >> 24) - (vold >> 24);
    }
}

}  // namespace

int launch_adaptive_mean(synseg_ctx *ctx, const synseg_img *gray, const synseg_img *out_u8, BitPlane out_bits,
                         int block_size, int C, int invert, cudaStream_t st)
{
    AdParams p;
    p.src = plane_of(gray);
    p.width = gray->width; p.height = gray->height;
    p.r = block_size / 2;
    p.lead = (int)align_up((size_t)p.r, 32);
    const bool to_bits = (out_u8 == nullptr);
    if (!to_bits) p.dst = plane_of(out_u8); else p.dst = Plane{nullptr, 0, 0};
    p.bits = out_bits;
    // one strip if the row + halos fit 1024 threads x 4 columns, otherwise strips of 4096 - 2*lead
    int need = (int)align_up((size_t)p.width, 32) + 2 * p.lead;
    int T;
    if (need <= 4096) { T = (int)align_up((size_t)cdiv(need, 4), 32); p.out_w = 4 * T - 2 * p.lead; p.strips = 1; }
    else { T = 1024; p.out_w = 4096 - 2 * p.lead; p.strips = cdiv(p.width, p.out_w); }
    if (T < 64) T = 64;
    // bands: enough CTAs for ~4 per SM, band height 32..256
    int64_t rows_total = (int64_t)gray->height * gray->batch * p.strips;
    int band_h = (int)(rows_total / (4 * (int64_t)ctx->sm_count));
    band_h = band_h < 32 ? 32 : (band_h > 256 ? 256 : band_h);
    if (band_h > gray->height) band_h = gray->height;
    p.band_h = band_h;
    p.bands = cdiv(gray->height, band_h);
    p.C = C; p.invert = invert;
    p.n = (uint32_t)block_size * block_size;
    const uint64_t d = 2ull * p.n;
    p.magic = ((1ull << 48) + d - 1) / d;
    const int64_t nblocks = (int64_t)gray->batch * p.strips * p.bands;
    const size_t smem = (size_t)(2 * (4 * T + 1) + 64) * sizeof(uint32_t);
    const bool sal = plane_aligned(gray, 4);
    const bool dal = to_bits ? true : plane_aligned(out_u8, 4);
    if (to_bits) adaptive_mean_kernel<true><<<(unsigned)nblocks, T, smem, st>>>(p, sal, dal);
    else adaptive_mean_kernel<false><<<(unsigned)nblocks, T, smem, st>>>(p, sal, dal);
    SS_LAUNCH_CHECK(ctx, "adaptive_mean", st);
    return SYNSEG_OK;
}

extern "C" SYNSEG_EXPORT int synseg_adaptive_mean(synseg_ctx *ctx, const synseg_img *gray, const synseg_img *out, int block_size, int C,
                                    int invert, void *stream)
{
    if (!ctx) { synseg_set_error("synseg_adaptive_mean: ctx is NULL"); return SYNSEG_E_INVALID; }
    SS_TRY(validate_img(gray, "gray", 1));
    SS_TRY(validate_img(out, "out", 1));
    if (!same_shape(gray, out)) { synseg_set_error("synseg_adaptive_mean: shape mismatch"); return SYNSEG_E_INVALID; }
    if (block_size < 3 || block_size > 255 || !(block_size & 1)) {
        synseg_set_error("synseg_adaptive_mean: block_size must be odd in 3..255 (got %d)", block_size);
        return SYNSEG_E_INVALID;
    }
    return launch_adaptive_mean(ctx, gray, out, BitPlane{nullptr, 0, 0}, block_size, C, invert ? 1 : 0, (cudaStream_t)stream);
}
