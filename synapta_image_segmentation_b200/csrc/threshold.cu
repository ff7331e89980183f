// threshold.cu -- cv2.adaptiveThreshold(ADAPTIVE_THRESH_MEAN_C), bit-exact.
//
// No reference call site (north-star primitive); oracle = cv2 4.13.0:
//   mean = round(boxsum / bs^2) over BORDER_REPLICATE  (bs odd -> never a tie, so
//          mean = floor((2s + n) / 2n) exactly),
//   THRESH_BINARY: 255 iff g - mean > -C ;  THRESH_BINARY_INV: 255 iff g - mean <= -C.
//
// Roofline: HBM-bound, 2 algorithmic bytes per pixel with a u8 output (1 read + 1 written), 1.125 with
// the bit-plane output the page pipeline uses.
// Design: warp-autonomous strips, no block-level barrier.  A warp owns a strip of 512 columns (16 per
// lane, one 128-bit load per lane and row) and marches down a band of rows keeping the vertical running
// column sums (row y+r+1 in, row y-r out) in registers, two 16-bit sums per register (255 rows x 255 fit).
// Each row the horizontal window sums come from prefixes of the column sums parked in shared memory, S = P[x+r] - P[x-r-1]:
// the generic kernel (any block size) scans them across the warp and keeps a transposed, conflict-free tile; the kernels with a
// compile-time radius (the page pipeline's block sizes 51 and 25) keep lane-local prefixes in a lane-major tile read with 128-bit
// loads at immediate offsets and fold the few neighbouring lane totals a window spans into the constant of the test.
// The threshold test needs no division: mean >= g + C  <=>  2s + n >= 2n (g + C)  <=>  s >= n g + n C - (n-1)/2,
// and its sign bit is funnel-shifted straight into the output word.  Rows on which no output pixel of the strip has
// g <= 255 - C (blank paper: the mean cannot exceed 255) skip the prefix and the test altogether (warp-uniform).
// Every source byte comes from HBM once; the "row out" and centre-row reads of a band hit L2.
// Output is either a u8 {0,255} plane or a bit plane (two lanes -> one 32-bit word).
#include <stdlib.h>

#include "internal.cuh"
#include "pixel.cuh"

namespace {

constexpr int CPL = 16;                 // columns per lane
constexpr int SW = 32 * CPL;            // columns per warp strip
#ifndef SYNSEG_AD_WARPS
#define SYNSEG_AD_WARPS 1
#endif
constexpr int AD_WARPS = SYNSEG_AD_WARPS;   // independent warps per CTA (1: a finished warp frees its registers at once)
#ifndef SYNSEG_AD_MINBLOCKS
#define SYNSEG_AD_MINBLOCKS (16 / SYNSEG_AD_WARPS)
#endif
constexpr unsigned FULL = 0xffffffffu;

struct AdParams {
    Plane src;
    Plane dst;       // u8 output (OUT_BITS = false)
    BitPlane bits;   // bit output (OUT_BITS = true)
    int width, height;
    int r;           // bs / 2
    int lead;        // round_up(r + 1, 32): columns of left halo the strip carries
    int out_w;       // output columns per strip (multiple of 32)
    int strips, bands, band_h;
    int n;           // bs * bs
    int k2m1;        // n C - (n - 1) / 2 - 1
    int C;           // clamped to [-256, 256] (g - mean lies in [-255, 255])
    int skip_ok;     // 1 <= C <= 128: the blank-row shortcut applies
    uint32_t skip_add;
    int invert;
    int64_t tasks;   // batch * bands * strips
};

// cs[2q] holds columns 4q (low half) and 4q+2 (high half); cs[2q+1] columns 4q+1 and 4q+3
__device__ __forceinline__ void add16(uint32_t cs[8], const uint4 v)
{
    cs[0] += __byte_perm(v.x, 0, 0x4240); cs[1] += __byte_perm(v.x, 0, 0x4341);
    cs[2] += __byte_perm(v.y, 0, 0x4240); cs[3] += __byte_perm(v.y, 0, 0x4341);
    cs[4] += __byte_perm(v.z, 0, 0x4240); cs[5] += __byte_perm(v.z, 0, 0x4341);
    cs[6] += __byte_perm(v.w, 0, 0x4240); cs[7] += __byte_perm(v.w, 0, 0x4341);
}
__device__ __forceinline__ void sub16(uint32_t cs[8], const uint4 v)
{
    cs[0] -= __byte_perm(v.x, 0, 0x4240); cs[1] -= __byte_perm(v.x, 0, 0x4341);
    cs[2] -= __byte_perm(v.y, 0, 0x4240); cs[3] -= __byte_perm(v.y, 0, 0x4341);
    cs[4] -= __byte_perm(v.z, 0, 0x4240); cs[5] -= __byte_perm(v.z, 0, 0x4341);
    cs[6] -= __byte_perm(v.w, 0, 0x4240); cs[7] -= __byte_perm(v.w, 0, 0x4341);
}

#ifndef SYNSEG_AD_DEPTH
#define SYNSEG_AD_DEPTH 4
#endif
constexpr int AD_DEPTH = SYNSEG_AD_DEPTH;   // rows in flight per warp (cp.async ring; a power of two)
#ifndef SYNSEG_AD_PRO
#define SYNSEG_AD_PRO 0
#endif
constexpr int AD_PRO = SYNSEG_AD_PRO;   // prologue of a band (the 2r + 1 rows of the first window): 0 = cp.async through shared memory, n = n loads in registers

__device__ __forceinline__ uint32_t bytes_of_nib(uint32_t nib) { return ((nib * 0x00204081u) & 0x01010101u) * 0xFFu; }

// ALIGNED (source rows 16-byte aligned) is a template parameter so that the byte-load fallback does not sit between
// the hot instructions of the common case (the loop body has to stay inside the instruction cache).
//
// R > 0 fixes the window radius at compile time (the page pipeline's block sizes, 51 at 300 DPI and 25 at 150 DPI).  The prefix
// tile is then lane-major with a pitch of 20 words (five 16-byte units: eight consecutive lanes hit eight different bank groups, so
// 128-bit accesses are conflict-free): a lane parks its 16 prefixes with 4 STS.128 and fetches the 2 x 16 prefixes its window
// sums need (columns x + R and x - R - 1, statically known words of the neighbouring lanes' chunks) with 2 x 5 LDS.128 instead of
// 32 scalar loads at run-time computed addresses (per tested pixel: PRMT, IMAD, IADD3, SHF).  R = 0 is the generic kernel.
constexpr int PL_PITCH = 20;
template <bool OUT_BITS, bool ALIGNED, int R>
__global__ void __launch_bounds__(32 * AD_WARPS, SYNSEG_AD_MINBLOCKS) adaptive_mean_kernel(AdParams p, bool dst_aligned)
{
    constexpr bool src_aligned = ALIGNED;
    struct WarpSmem {
        uint4 ring[AD_DEPTH][3][32];                 // cp.async row ring: [step][new | old | centre][lane]
        uint32_t P[2][32 * PL_PITCH];                // prefix tile, double-buffered (R = 0: [column within lane][lane], pitch 33: conflict-free both ways)
    };
    __shared__ __align__(16) WarpSmem Sm[AD_WARPS];
    constexpr int PRO_SLOTS = (int)(sizeof(WarpSmem) / 512);          // the whole per-warp area as row slots for the prologue (22)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t task = (int64_t)blockIdx.x * AD_WARPS + warp;
    if (task >= p.tasks) return;                      // warp-uniform
    const int strip = (int)(task % p.strips); task /= p.strips;
    const int band = (int)(task % p.bands);
    const int img = (int)(task / p.bands);

    const int W = p.width, H = p.height, r = R ? R : p.r;
    const int n = R ? (2 * R + 1) * (2 * R + 1) : p.n;
    const int x = strip * p.out_w - p.lead + CPL * lane;     // first of this lane's 16 columns
    const int y0 = band * p.band_h, y1 = min(y0 + p.band_h, H);
    const uint8_t *base = p.src.p + img * p.src.bs;
    const bool out_lane = (CPL * lane >= p.lead) && (CPL * lane < p.lead + p.out_w) && (x < W);
    // Aligned rows: every lane reads one 16-byte unit per row at a clamped address (clamp16_x); the edge pixel is
    // replicated in registers when the row is consumed (fix16_rep; a no-op compare for interior lanes).
    const int xl = clamp16_x(x, W);
    const EdgeFix efix = make_edge_fix(x, W);
    auto ldraw = [&](int yy) -> uint4 { return __ldg((const uint4 *)(base + (int64_t)yy * p.src.rs + xl)); };

    // running column sums over rows [y - r, y + r] (replicate)
    uint32_t cs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (src_aligned) {
        if (AD_PRO == 0) {
            // the 2r + 1 rows of the first window through cp.async: the whole per-warp shared-memory area serves as 22 row slots, so
            // the prologue of a band waits for memory three times (r = 25) instead of seven times with eight loads in registers
            uint4 *scr = (uint4 *)&Sm[warp] + lane;
            for (int dy = -r; dy <= r; dy += PRO_SLOTS) {
                const int cnt = min(PRO_SLOTS, r - dy + 1);
                for (int q = 0; q < cnt; ++q) cp_async16(scr + q * 32, base + (int64_t)min(max(y0 + dy + q, 0), H - 1) * p.src.rs + xl);
                cp_async_commit();
                cp_async_wait<0>();
                for (int q = 0; q < cnt; ++q) add16(cs, apply_edge_fix(scr[q * 32], efix));
            }
        } else {
            for (int dy = -r; dy <= r; dy += (AD_PRO ? AD_PRO : 1)) {       // AD_PRO loads in flight, then as many accumulations
                uint4 t[AD_PRO ? AD_PRO : 1];
#pragma unroll
                for (int q = 0; q < AD_PRO; ++q) if (dy + q <= r) t[q] = ldraw(min(max(y0 + dy + q, 0), H - 1));
#pragma unroll
                for (int q = 0; q < AD_PRO; ++q) if (dy + q <= r) add16(cs, apply_edge_fix(t[q], efix));
            }
        }
    } else {
        for (int dy = -r; dy <= r; ++dy) add16(cs, load16_rep(base + (int64_t)min(max(y0 + dy, 0), H - 1) * p.src.rs, x, W, false));
    }
    const int kx = CPL * lane + r, ky = CPL * lane - r - 1;  // prefix indices of column j: kx + j, ky + j
    uint32_t colmask = 0;                                    // output columns of this lane inside the image
    if (out_lane) colmask = (W - x >= 16) ? 0xFFFFu : ((1u << (W - x)) - 1u);

    // Software pipeline (aligned rows): the three row loads of step y + AD_DEPTH are issued with cp.async into a
    // per-warp shared-memory ring while step y is computed -- four rows of latency cover without a single extra
    // register (every lane reads back only what it copied itself, so no barrier is needed).
    uint4 *ring = &Sm[warp].ring[0][0][lane];
    // running pointers (no 64-bit multiplies in the row loop): rows y+r+1 (clamped to H-1), y-r (clamped to 0) and y
    int y_pf = y0;                                                    // next step to prefetch
    const uint8_t *p_new = base + (int64_t)min(y0 + r + 1, H - 1) * p.src.rs + xl;
    const uint8_t *p_old = base + (int64_t)max(y0 - r, 0) * p.src.rs + xl;
    const uint8_t *p_cen = base + (int64_t)y0 * p.src.rs + xl;
    const int64_t rs = p.src.rs;
    auto issue_async = [&](int slot) {
        if (y_pf < y1) {
            cp_async16(ring + (slot * 3 + 0) * 32, p_new);
            cp_async16(ring + (slot * 3 + 1) * 32, p_old);
            cp_async16(ring + (slot * 3 + 2) * 32, p_cen);
        }
        cp_async_commit();
        if (y_pf + r + 1 < H - 1) p_new += rs;                        // row of the next step: min(y_pf + 1 + r + 1, H - 1)
        if (y_pf - r >= 0) p_old += rs;                               //                       max(y_pf + 1 - r, 0)
        p_cen += rs;
        ++y_pf;
    };
    if (src_aligned) {
#pragma unroll
        for (int d = 0; d < AD_DEPTH; ++d) issue_async(d);
    }
    int slot = 0;
    uint32_t *obits = OUT_BITS ? p.bits.p + img * p.bits.bs + (int64_t)y0 * p.bits.wpr + (max(x, 0) >> 5) : nullptr;

    for (int y = y0; y < y1; ++y) {
        uint32_t *Pb = Sm[warp].P[y & 1];
        uint4 vnew, vold, vcen;
        if (src_aligned) {
            cp_async_wait<AD_DEPTH - 1>();
            vnew = apply_edge_fix(ring[(slot * 3 + 0) * 32], efix);
            vold = apply_edge_fix(ring[(slot * 3 + 1) * 32], efix);
            vcen = apply_edge_fix(ring[(slot * 3 + 2) * 32], efix);
            issue_async(slot);
            slot = (slot + 1) & (AD_DEPTH - 1);
        } else {
            vnew = load16_rep(base + (int64_t)min(y + r + 1, H - 1) * p.src.rs, x, W, false);
            vold = load16_rep(base + (int64_t)max(y - r, 0) * p.src.rs, x, W, false);
            vcen = load16_rep(base + (int64_t)y * p.src.rs, x, W, false);
        }

        // Row shortcut: mean <= 255, so a pixel with g + C > 255 can never satisfy mean >= g + C.  When no output
        // pixel of the whole strip row has g <= 255 - C (blank paper) the result is known without the window sums.
        if (p.skip_ok) {
            bool need = false;
            if (out_lane) {
                const uint32_t add = p.skip_add;               // 0x01010101 * (127 - (C - 1))
                const uint32_t a0 = ~vcen.x, a1 = ~vcen.y, a2 = ~vcen.z, a3 = ~vcen.w;   // 255 - g > C - 1 ?
                need = ((((a0 + add) | a0) | ((a1 + add) | a1) | ((a2 + add) | a2) | ((a3 + add) | a3)) & 0x80808080u) != 0u;
            }
            if (!__any_sync(FULL, need)) {
                const uint32_t cbits = p.invert ? 0u : colmask;
                if (OUT_BITS) {
                    const uint32_t other = __shfl_xor_sync(FULL, cbits, 1);
                    if (out_lane && !(lane & 1)) obits[(int64_t)(y - y0) * p.bits.wpr] = cbits | (other << 16);
                } else if (out_lane) {
                    uint8_t *drow = p.dst.p + img * p.dst.bs + (int64_t)y * p.dst.rs + x;
                    const uint32_t fill = p.invert ? 0u : 0xFFFFFFFFu;
                    if (dst_aligned && x + 15 < W) *(uint4 *)drow = make_uint4(fill, fill, fill, fill);
                    else {
                        const int nvalid = min(16, W - x);
                        for (int j = 0; j < nvalid; ++j) drow[j] = (uint8_t)fill;
                    }
                }
                sub16(cs, vold);
                add16(cs, vnew);
                __syncwarp();
                continue;
            }
        }

        // inclusive prefix of the 16 column sums of this lane
        uint32_t a[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t e = cs[2 * q], o = cs[2 * q + 1];
            a[4 * q + 0] = (q ? a[4 * q - 1] : 0u) + (e & 0xFFFFu);
            a[4 * q + 1] = a[4 * q + 0] + (o & 0xFFFFu);
            a[4 * q + 2] = a[4 * q + 1] + (e >> 16);
            a[4 * q + 3] = a[4 * q + 2] + (o >> 16);
        }
        uint32_t bits16 = 0;
        // mean >= g + C  <=>  2s + n >= 2n (g + C)  <=>  s >= n g + k2  (n odd)  <=>  sign(n g + k2 - 1 - s) set
        if (R > 0) {
            // No warp-wide scan: a window spans at most LB + LXMAX + 1 lanes, so the tile holds the lane-LOCAL prefixes and the
            // difference of the two lanes' offsets -- a sum of a few neighbouring lane totals, fetched with independent shuffles
            // instead of five dependent ones -- is folded into the constant of the test.
            constexpr int XF = R >> 2;                     // first 4-word chunk holding a window-end word (counted from the own chunk 0)
            constexpr int LB = (R + 16) / 16;              // lanes back to the chunk row of the first window-start word
            constexpr int Y0 = 16 * LB - R - 1;            // its word in that row (>= 0)
            constexpr int YF = Y0 >> 2;
            constexpr int LXMIN = R >> 4, LXMAX = (R + 15) >> 4;      // lane offsets of the window-end words
            uint4 *own = (uint4 *)(Pb + lane * PL_PITCH);
#pragma unroll
            for (int c = 0; c < 4; ++c) own[c] = make_uint4(a[4 * c], a[4 * c + 1], a[4 * c + 2], a[4 * c + 3]);
            uint32_t Cs[LB + LXMAX + 1];                   // Cs[i] = totals of the lanes l - LB .. l - LB + i - 1
            Cs[0] = 0;
#pragma unroll
            for (int m = -LB; m < LXMAX; ++m) {
                const uint32_t t = m == 0 ? a[15] : (m < 0 ? __shfl_up_sync(FULL, a[15], m < 0 ? -m : 1) : __shfl_down_sync(FULL, a[15], m > 0 ? m : 1));
                Cs[m + LB + 1] = Cs[m + LB] + t;
            }
            __syncwarp();
            if (out_lane) {
                // local prefix word 16 lane + R + j (window end) and 16 lane - R - 1 + j (one before the window start), j = 0 .. 15
                uint32_t xw[20], yw[20];
#pragma unroll
                for (int c = 0; c < 5; ++c) {
                    const int fx = XF + c, fy = YF + c;
                    const uint4 tx = *(const uint4 *)(Pb + (lane + (fx >> 2)) * PL_PITCH + 4 * (fx & 3));
                    const uint4 ty = *(const uint4 *)(Pb + (lane - LB + (fy >> 2)) * PL_PITCH + 4 * (fy & 3));
                    xw[4 * c] = tx.x; xw[4 * c + 1] = tx.y; xw[4 * c + 2] = tx.z; xw[4 * c + 3] = tx.w;
                    yw[4 * c] = ty.x; yw[4 * c + 1] = ty.y; yw[4 * c + 2] = ty.z; yw[4 * c + 3] = ty.w;
                }
                int K[2][2];                               // [window end in lane l + LXMIN + ix][window start in lane l - LB + iy]
#pragma unroll
                for (int ix = 0; ix < 2; ++ix)
#pragma unroll
                    for (int iy = 0; iy < 2; ++iy) K[ix][iy] = p.k2m1 - (int)(Cs[(LXMIN + ix <= LXMAX ? LXMIN + ix : LXMAX) + LB] - Cs[iy]);
                const uint32_t cw[4] = {vcen.x, vcen.y, vcen.z, vcen.w};
#pragma unroll
                for (int j = 15; j >= 0; --j) {
                    const int g = (int)__byte_perm(cw[j >> 2], 0, 0x4440 + (j & 3));
                    const int e = g * n + K[((R + j) >> 4) - LXMIN][(Y0 + j) >> 4] - (int)xw[(R & 3) + j] + (int)yw[(Y0 & 3) + j];
                    bits16 = __funnelshift_l((uint32_t)e, bits16, 1);
                }
                if (!p.invert) bits16 = ~bits16;
                bits16 &= colmask;
            }
        } else {
            uint32_t v = a[15];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t nb = __shfl_up_sync(FULL, v, d);
                if (lane >= d) v += nb;
            }
            const uint32_t off = v - a[15];
#pragma unroll
            for (int j = 0; j < 16; ++j) Pb[j * 33 + lane] = off + a[j];
            __syncwarp();
            if (out_lane) {
                const uint32_t cw[4] = {vcen.x, vcen.y, vcen.z, vcen.w};
#pragma unroll
                for (int j = 15; j >= 0; --j) {
                    const int X = kx + j, Y = ky + j;              // Y >= 0 because lead >= r + 1
                    const int g = (int)((cw[j >> 2] >> (8 * (j & 3))) & 255u);
                    const int e = g * n + p.k2m1 - (int)Pb[(X & 15) * 33 + (X >> 4)] + (int)Pb[(Y & 15) * 33 + (Y >> 4)];
                    bits16 = __funnelshift_l((uint32_t)e, bits16, 1);
                }
                if (!p.invert) bits16 = ~bits16;
                bits16 &= colmask;
            }
        }
        if (OUT_BITS) {
            const uint32_t other = __shfl_xor_sync(FULL, bits16, 1);
            SS_DEVICE_ASSERT(!(out_lane && !(lane & 1)) || (y < H && (max(x, 0) >> 5) < p.bits.wpr));
            if (out_lane && !(lane & 1)) obits[(int64_t)(y - y0) * p.bits.wpr] = bits16 | (other << 16);
        } else if (out_lane) {
            uint8_t *drow = p.dst.p + img * p.dst.bs + (int64_t)y * p.dst.rs + x;
            if (dst_aligned && x + 15 < W) {
                *(uint4 *)drow = make_uint4(bytes_of_nib(bits16 & 15u), bytes_of_nib((bits16 >> 4) & 15u),
                                            bytes_of_nib((bits16 >> 8) & 15u), bytes_of_nib(bits16 >> 12));
            } else {
                const int nvalid = min(16, W - x);
                for (int j = 0; j < nvalid; ++j) drow[j] = ((bits16 >> j) & 1u) ? 255 : 0;
            }
        }
        sub16(cs, vold);
        add16(cs, vnew);
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
    p.lead = (int)align_up((size_t)p.r + 1, 32);
    p.out_w = (SW - p.lead - p.r) & ~31;
    const bool to_bits = (out_u8 == nullptr);
    if (!to_bits) p.dst = plane_of(out_u8); else p.dst = Plane{nullptr, 0, 0};
    p.bits = out_bits;
    p.strips = cdiv(p.width, p.out_w);
    // Bands: a band re-reads block_size - 1 rows of its neighbours (L2 hits, cheap: load + accumulate only).  Short
    // bands win because the cost of a band depends on its content (blank rows skip the prefix and the test), so
    // many small tasks balance the SMs better (measured on B200: 320 rows 1.16 ms, 64 rows 0.81 ms per 50 pages with the first strip kernel; 64-96 rows stay best).
    const int64_t rows_total = (int64_t)gray->height * gray->batch * p.strips;
    // (the prologue of a band, 2r + 1 latency-bound row loads, is 13 % of the stall samples at 52-row bands: 25-page chains now get 96 rows too)
    int band_h = (int)(rows_total / (32 * (int64_t)ctx->sm_count));
    band_h = band_h < 32 ? 32 : (band_h > 96 ? 96 : band_h);
    if (ctx->tune_ad_band > 0) band_h = ctx->tune_ad_band;
    if (band_h > gray->height) band_h = gray->height;
    p.band_h = band_h;
    p.bands = cdiv(gray->height, band_h);
    p.n = block_size * block_size;
    p.C = C < -256 ? -256 : (C > 256 ? 256 : C);
    p.k2m1 = p.n * p.C - (p.n - 1) / 2 - 1;
    p.skip_ok = (p.C >= 1 && p.C <= 128) ? 1 : 0;
    p.skip_add = 0x01010101u * (uint32_t)(127 - (p.C - 1));
    p.invert = invert;
    p.tasks = (int64_t)gray->batch * p.bands * p.strips;
    const bool sal = plane_aligned(gray, 16);
    const bool dal = to_bits ? true : plane_aligned(out_u8, 16);
    const unsigned nblocks = (unsigned)cdiv(p.tasks, AD_WARPS);
    static const bool generic_only = getenv("SYNSEG_AD_GENERIC") != nullptr;      // parity / tuning: always the run-time radius kernel
    if (to_bits) {
        if (sal && p.r == 25 && !generic_only) adaptive_mean_kernel<true, true, 25><<<nblocks, 32 * AD_WARPS, 0, st>>>(p, dal);
        else if (sal && p.r == 12 && !generic_only) adaptive_mean_kernel<true, true, 12><<<nblocks, 32 * AD_WARPS, 0, st>>>(p, dal);
        else if (sal) adaptive_mean_kernel<true, true, 0><<<nblocks, 32 * AD_WARPS, 0, st>>>(p, dal);
        else adaptive_mean_kernel<true, false, 0><<<nblocks, 32 * AD_WARPS, 0, st>>>(p, dal);
    } else {
        if (sal && p.r == 25 && !generic_only) adaptive_mean_kernel<false, true, 25><<<nblocks, 32 * AD_WARPS, 0, st>>>(p, dal);
        else if (sal && p.r == 12 && !generic_only) adaptive_mean_kernel<false, true, 12><<<nblocks, 32 * AD_WARPS, 0, st>>>(p, dal);
        else if (sal) adaptive_mean_kernel<false, true, 0><<<nblocks, 32 * AD_WARPS, 0, st>>>(p, dal);
        else adaptive_mean_kernel<false, false, 0><<<nblocks, 32 * AD_WARPS, 0, st>>>(p, dal);
    }
    SS_LAUNCH_CHECK(ctx, "adaptive_mean", st);
    return SYNSEG_OK;
}

extern "C" SYNSEG_EXPORT int synseg_adaptive_mean(synseg_ctx *ctx, const synseg_img *gray, const synseg_img *out, int block_size, int C,
                                    int invert, void *stream)
{
    if (!ctx) { synseg_set_error("synseg_adaptive_mean: ctx is NULL"); return SYNSEG_E_INVALID; }
    SS_ENTER(ctx, stream);
    SS_TRY(validate_img(gray, "gray", 1));
    SS_TRY(validate_img(out, "out", 1));
    if (!same_shape(gray, out)) { synseg_set_error("synseg_adaptive_mean: shape mismatch"); return SYNSEG_E_INVALID; }
    if (block_size < 3 || block_size > 255 || !(block_size & 1)) {
        synseg_set_error("synseg_adaptive_mean: block_size must be odd in 3..255 (got %d)", block_size);
        return SYNSEG_E_INVALID;
    }
    return launch_adaptive_mean(ctx, gray, out, BitPlane{nullptr, 0, 0}, block_size, C, invert ? 1 : 0, (cudaStream_t)stream);
}
