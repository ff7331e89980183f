// colors.cu -- batched dominant colours of a ragged batch of crops (BASELINE.json configs[3]).
//
// Reference: OCRProcessor._extract_dominant_colors, pdf_image_segmentation.py:1566-1594:
//   RGB -> HSV, mask = S > 30 & V > 40 & V < 240 (:1571-1574); fewer than 100 masked pixels -> [] (:1577);
//   otherwise KMeans(min(5, n)) over (an UNSEEDED random sample of 5000 of) the masked pixels (:1581-1590),
//   centres truncated to int and printed as '#rrggbb' (:1591-1592).
// The mask, the masked-pixel count and the `[]` decision are reproduced exactly.  The clustering is the deterministic
// histogram form (an APPROXIMATION of the reference's sampled KMeans, labelled as such everywhere): exact 4096-bin
// (R>>4, G>>4, B>>4) histogram with per-bin channel sums, then weighted Lloyd iterations over the bin centroids,
// started from the heaviest bins.  No sampling, no randomness; bit-identical with oracle/colors_port.py:
//   points    p_b = chan_sum_b / count_b                      (IEEE f64 divisions)
//   start     centres = points of the k heaviest bins (ties: lower bin index first), k = min(n_colors, non-empty bins)
//   iteration assign every bin to the nearest centre, distance (dr*dr + dg*dg) + db*db in f64 without fused
//             multiply-add, ties to the lower centre index; centre_j = (integer sum of chan_sum over its bins) /
//             (integer sum of count) -- empty clusters keep their centre; stop after `iters` rounds or when no
//             centre moved.
// One CTA per crop: the histogram lives in shared memory (64 KB: counts + three channel sums) and is filled with
// warp-aggregated atomics (lanes holding the same bin elect a leader that adds the group's count and channel sums
// once), so nothing but the crop itself is read from HBM and nothing but the result is written.
// Roofline: HBM-bound, 3 (RGB) or 4 (RGBX) algorithmic bytes per pixel.
#include "internal.cuh"
#include "pixel.cuh"

#include <algorithm>

namespace {

constexpr int CBINS = 4096;
constexpr int CTHREADS = 512;
constexpr int CMAXK = 8;
typedef unsigned long long u64;

struct ColorSmem {
    uint32_t h[CBINS];            // masked pixels per bin
    uint32_t cs[3 * CBINS];       // channel sums per bin (<= 255 * 2^24 pixels per crop)
    uint32_t sdiv[256];
    uint16_t list[CBINS];         // non-empty bins in ascending order
    uint32_t warp_cnt[CTHREADS / 32];
    u64 red[CTHREADS / 32];
    double centre[CMAXK][3];
    u64 acc[CMAXK][4];            // per cluster: count, sum r, sum g, sum b
    u64 picked;
    int n_list, changed;
    u64 total;
};

__device__ __forceinline__ void add_pixel(ColorSmem &S, bool on, uint32_t rr, uint32_t gg, uint32_t bb, int lane, uint32_t &cnt)
{
    const unsigned m = __ballot_sync(0xffffffffu, on);
    if (m == 0u) return;                                   // white / grey / black stretch: nothing to add
    cnt += __popc(m);
    if (on) {
        const int bin = ((rr >> 4) << 8) | ((gg >> 4) << 4) | (bb >> 4);
        const unsigned peers = __match_any_sync(m, bin);
        const unsigned sr = __reduce_add_sync(peers, rr), sg = __reduce_add_sync(peers, gg), sb = __reduce_add_sync(peers, bb);
        if (lane == __ffs(peers) - 1) {
            atomicAdd(&S.h[bin], (unsigned)__popc(peers));
            atomicAdd(&S.cs[3 * bin], sr); atomicAdd(&S.cs[3 * bin + 1], sg); atomicAdd(&S.cs[3 * bin + 2], sb);
        }
    }
}

__device__ __forceinline__ u64 block_max_u64(ColorSmem &S, u64 v)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { const u64 o = __shfl_down_sync(0xffffffffu, v, d); v = o > v ? o : v; }
    __syncthreads();
    if (lane == 0) S.red[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        u64 b = 0;
        for (int w = 0; w < CTHREADS / 32; ++w) b = S.red[w] > b ? S.red[w] : b;
        S.picked = b;
    }
    __syncthreads();
    return S.picked;
}

// out: per crop `stride` = 2 + n_colors words: { mask_px, k, (cluster pixels << 24 | R << 16 | G << 8 | B) x k, 0 ... }
__global__ void __launch_bounds__(CTHREADS) crop_colors_kernel(const uint8_t *base, const CropTask *tasks, int n_colors, int iters, int min_px,
                                                               u64 *out, uint32_t *hist_out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ColorSmem &S = *(ColorSmem *)smem_raw;
    const CropTask t = tasks[blockIdx.x];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int stride = 2 + n_colors;
    u64 *o = out + (int64_t)t.out_index * stride;

    for (int i = threadIdx.x; i < CBINS; i += CTHREADS) S.h[i] = 0;
    for (int i = threadIdx.x; i < 3 * CBINS; i += CTHREADS) S.cs[i] = 0;
    if (threadIdx.x < 256) S.sdiv[threadIdx.x] = hsv_sdiv(threadIdx.x);
    if (threadIdx.x < 4 * CMAXK) ((u64 *)S.acc)[threadIdx.x] = 0ull;
    if (threadIdx.x == 0) S.total = 0;
    __syncthreads();

    // ---- histogram of the masked pixels: one warp per row, a lane per pixel -----------------------------------
    uint32_t cnt = 0;
    if (t.channels >= 3) {
        const uint8_t *src = base + t.offset;
        const int W = t.width;
        for (int y = warp; y < t.height; y += CTHREADS / 32) {
            const uint8_t *row = src + (int64_t)y * t.row_stride;
            if (t.channels == 4) {
                const uint32_t *wp = (const uint32_t *)row;                 // RGBX words (4-byte aligned by contract)
                for (int x0 = 0; x0 < W; x0 += 32) {
                    const int x = x0 + lane;
                    bool on = false;
                    uint32_t rr = 0, gg = 0, bb = 0;
                    if (x < W) {
                        const uint32_t v = __ldg(wp + x);
                        rr = v & 255u; gg = (v >> 8) & 255u; bb = (v >> 16) & 255u;
                        on = hsv_mask_px(rr, gg, bb, S.sdiv);
                    }
                    add_pixel(S, on, rr, gg, bb, lane, cnt);
                }
            } else {
                for (int x0 = 0; x0 < W; x0 += 32) {
                    const int x = x0 + lane;
                    bool on = false;
                    uint32_t rr = 0, gg = 0, bb = 0;
                    if (x < W) {
                        const uint8_t *p = row + 3 * x;
                        rr = __ldg(p); gg = __ldg(p + 1); bb = __ldg(p + 2);
                        on = hsv_mask_px(rr, gg, bb, S.sdiv);
                    }
                    add_pixel(S, on, rr, gg, bb, lane, cnt);
                }
            }
        }
    }
    if (lane == 0 && cnt) atomicAdd(&S.total, (u64)cnt);
    __syncthreads();
    const u64 total = S.total;
    if (hist_out) {
        uint32_t *ho = hist_out + (int64_t)t.out_index * CBINS;
        for (int i = threadIdx.x; i < CBINS; i += CTHREADS) ho[i] = S.h[i];
    }
    if (threadIdx.x < stride) o[threadIdx.x] = threadIdx.x == 0 ? total : 0ull;
    if (total < (u64)min_px) return;                       // the reference's `len(pixels) < 100 -> []`

    // ---- non-empty bins, ascending (ballot compaction, CTHREADS bins per round) ---------------------------------
    if (threadIdx.x == 0) S.n_list = 0;
    __syncthreads();
    for (int b0 = 0; b0 < CBINS; b0 += CTHREADS) {
        const int bin = b0 + threadIdx.x;
        const bool ne = S.h[bin] != 0u;
        const unsigned m = __ballot_sync(0xffffffffu, ne);
        if (lane == 0) S.warp_cnt[warp] = __popc(m);
        __syncthreads();
        int off = S.n_list;
        for (int w = 0; w < warp; ++w) off += S.warp_cnt[w];
        if (ne) S.list[off + __popc(m & ((1u << lane) - 1u))] = (uint16_t)bin;
        __syncthreads();
        if (threadIdx.x == 0) { int s = 0; for (int w = 0; w < CTHREADS / 32; ++w) s += S.warp_cnt[w]; S.n_list += s; }
        __syncthreads();
    }
    const int nb = S.n_list;
    const int k = n_colors < nb ? n_colors : nb;

    // ---- start: the k heaviest bins, ties to the lower bin index (keys are unique) ----------------------------
    u64 below = ~0ull;
    for (int j = 0; j < k; ++j) {
        u64 best = 0;
        for (int i = threadIdx.x; i < nb; i += CTHREADS) {
            const int bin = S.list[i];
            const u64 key = ((u64)S.h[bin] << 12) | (u64)(CBINS - 1 - bin);
            if (key < below && key > best) best = key;
        }
        below = block_max_u64(S, best);
        if (threadIdx.x < 3) {
            const int bin = CBINS - 1 - (int)(below & (CBINS - 1));
            S.centre[j][threadIdx.x] = (double)S.cs[3 * bin + threadIdx.x] / (double)S.h[bin];
        }
    }
    __syncthreads();

    // ---- weighted Lloyd iterations over the bin centroids -------------------------------------------------------
    for (int it = 0; it < iters; ++it) {
        if (threadIdx.x < 4 * CMAXK) ((u64 *)S.acc)[threadIdx.x] = 0ull;
        if (threadIdx.x == 0) S.changed = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < nb; i += CTHREADS) {
            const int bin = S.list[i];
            const uint32_t w = S.h[bin], sr = S.cs[3 * bin], sg = S.cs[3 * bin + 1], sb = S.cs[3 * bin + 2];
            const double pr = (double)sr / (double)w, pg = (double)sg / (double)w, pb = (double)sb / (double)w;
            int a = 0;
            double dbest = 0.0;
            for (int j = 0; j < k; ++j) {
                const double dr = __dsub_rn(pr, S.centre[j][0]), dg = __dsub_rn(pg, S.centre[j][1]), db = __dsub_rn(pb, S.centre[j][2]);
                const double d = __dadd_rn(__dadd_rn(__dmul_rn(dr, dr), __dmul_rn(dg, dg)), __dmul_rn(db, db));
                if (j == 0 || d < dbest) { dbest = d; a = j; }
            }
            atomicAdd(&S.acc[a][0], (u64)w); atomicAdd(&S.acc[a][1], (u64)sr); atomicAdd(&S.acc[a][2], (u64)sg); atomicAdd(&S.acc[a][3], (u64)sb);
        }
        __syncthreads();
        if (threadIdx.x < 3 * k) {
            const int j = threadIdx.x / 3, c = threadIdx.x % 3;
            if (S.acc[j][0]) {
                const double nc = (double)S.acc[j][1 + c] / (double)S.acc[j][0];
                if (nc != S.centre[j][c]) { S.centre[j][c] = nc; S.changed = 1; }
            }
        }
        __syncthreads();
        const int moved = S.changed;
        __syncthreads();
        if (!moved) break;                                 // a fixed point: further rounds would repeat this one
    }
    if (threadIdx.x < k) {
        const int j = threadIdx.x;
        const u64 r = (u64)(int)S.centre[j][0], g = (u64)(int)S.centre[j][1], b = (u64)(int)S.centre[j][2];
        o[2 + j] = (S.acc[j][0] << 24) | (r << 16) | (g << 8) | b;
    }
    if (threadIdx.x == 0) o[1] = (u64)k;
}

}  // namespace

extern "C" SYNSEG_EXPORT int synseg_colors_crops(synseg_ctx *ctx, const void *base, const synseg_crop *crops_host, int32_t n, int32_t n_colors,
                                                 int32_t iters, int32_t min_pixels, uint64_t *out, uint32_t *hist_out, void *stream)
{
    if (!ctx) { synseg_set_error("synseg_colors_crops: ctx is NULL"); return SYNSEG_E_INVALID; }
    if (n <= 0) return SYNSEG_OK;
    if (!base || !crops_host || !out) { synseg_set_error("synseg_colors_crops: NULL argument"); return SYNSEG_E_INVALID; }
    if (n_colors < 1 || n_colors > CMAXK || iters < 0) { synseg_set_error("synseg_colors_crops: need 1 <= n_colors <= %d, iters >= 0", CMAXK); return SYNSEG_E_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<CropTask> tasks(n);
    for (int i = 0; i < n; ++i) {
        const synseg_crop &c = crops_host[i];
        if (c.width <= 0 || c.height <= 0 || (c.channels != 1 && c.channels != 3 && c.channels != 4) || c.row_stride < (int64_t)c.width * c.channels) {
            synseg_set_error("synseg_colors_crops: bad crop %d", i); return SYNSEG_E_INVALID;
        }
        if ((int64_t)c.width * c.height > ((int64_t)1 << 24)) {      // 32-bit channel sums per bin
            synseg_set_error("synseg_colors_crops: crop %d has more than 2^24 pixels", i); return SYNSEG_E_INVALID;
        }
        if (c.channels == 4 && ((((uintptr_t)base + c.offset) | (uint64_t)c.row_stride) & 3)) {
            synseg_set_error("synseg_colors_crops: RGBX crop %d is not 4-byte aligned", i); return SYNSEG_E_INVALID;
        }
        tasks[i] = CropTask{(int64_t)c.offset, c.row_stride, c.width, c.height, c.channels, i};
    }
    // largest crops first: one CTA per crop, the hardware hands CTAs out in order (longest-processing-time-first balance)
    std::stable_sort(tasks.begin(), tasks.end(), [](const CropTask &a, const CropTask &b) {
        return (int64_t)a.width * a.height * (a.channels >= 3) > (int64_t)b.width * b.height * (b.channels >= 3);
    });
    SS_TRY(arena_ensure(ctx, sizeof(CropTask) * (size_t)n + 4096));
    arena_begin(ctx);
    void *p;
    SS_TRY(arena_alloc(ctx, sizeof(CropTask) * (size_t)n, &p, st));
    // pageable source: the driver stages the bytes before cudaMemcpyAsync returns, so the vector may die with this call
    SS_CUDA(cudaMemcpyAsync(p, tasks.data(), sizeof(CropTask) * (size_t)n, cudaMemcpyHostToDevice, st));
    static bool attr_set = false;
    if (!attr_set) {
        SS_CUDA(cudaFuncSetAttribute(crop_colors_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ColorSmem)));
        attr_set = true;
    }
    crop_colors_kernel<<<n, CTHREADS, sizeof(ColorSmem), st>>>((const uint8_t *)base, (const CropTask *)p, n_colors, iters, min_pixels,
                                                             (u64 *)out, hist_out);
    SS_LAUNCH_CHECK(ctx, "crop_colors", st);
    return SYNSEG_OK;
}
