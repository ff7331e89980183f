// colors.cu -- batched dominant colours of a ragged batch of crops (BASELINE.json configs[3]).
//
// Reference: OCRProcessor._extract_dominant_colors, pdf_image_segmentation.py:1566-1594:
//   RGB -> HSV, mask = S > 30 & V > 40 & V < 240 (:1571-1574); fewer than 100 masked pixels -> [] (:1577);
//   otherwise KMeans(min(5, n)) over (an UNSEEDED random sample of 5000 of) the masked pixels (:1581-1590),
//   centres truncated to int and printed as '#rrggbb' (:1591-1592).
// The mask, the masked-pixel count and the `[]` decision are reproduced exactly.  The clustering is the deterministic
// histogram form (an APPROXIMATION of the reference's sampled KMeans, labelled as such everywhere): exact 4096-bin
// (R>>4, G>>4, B>>4) histogram with per-bin channel sums, then weighted Lloyd iterations over the bin centroids,
// started from the heaviest bins.  No sampling, no randomness, no floating point (fixed point, Q = 64 = 1/64 of a grey
// level, so the result does not depend on contraction or summation order); bit-identical with oracle/colors_port.py:
//   points    p_b = (Q * chan_sum_b + count_b / 2) / count_b          (integer division: bin centroid rounded to 1/64)
//   start     centres = points of the k heaviest bins (ties: lower bin index first), k = min(n_colors, non-empty bins)
//   iteration assign every bin to the nearest centre (squared distance in int32, < 2^31; ties to the lower centre index);
//             centre_j = (sum of count_b * p_b + N_j / 2) / N_j with N_j = sum of count_b over its bins -- empty
//             clusters keep their centre; stop after `iters` rounds or when no centre moved.
//   output    centre / Q (truncation, like the reference's `.astype(int)`), cluster pixel counts of the last assignment.
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
    uint16_t list[CBINS];         // non-empty bins in ascending order (bits 0-11; bits 12-15: cluster of the bin)
    uint32_t warp_cnt[CTHREADS / 32];
    u64 red[CTHREADS / 32];
    int32_t centre[CMAXK][3];     // Q6 fixed point
    u64 acc[CMAXK][4];            // per cluster: [0] = pixels of the last assignment
    u64 part[CTHREADS / 32][4];   // per warp: count, sum of count * point (r, g, b) of its cluster in this round
    u64 picked;
    int n_list, changed;
    u64 total;
};

// Flat colour areas (bars, fills) send the same bin from every lane for thousands of pixels in a row: such warp-uniform
// groups are accumulated in registers (one open run per warp) and reach shared memory once, when the bin changes or the
// warp is done -- the shared-memory atomics of 16 warps hammering one address were the bottleneck before.
struct RunAcc { int bin; uint32_t n, r, g, b; };

__device__ __forceinline__ void flush_run(ColorSmem &S, RunAcc &a, int lane)
{
    if (a.bin < 0) return;                                  // warp-uniform
    const unsigned n = __reduce_add_sync(0xffffffffu, a.n), r = __reduce_add_sync(0xffffffffu, a.r),
                   g = __reduce_add_sync(0xffffffffu, a.g), b = __reduce_add_sync(0xffffffffu, a.b);
    if (lane == 0) {
        atomicAdd(&S.h[a.bin], n);
        atomicAdd(&S.cs[3 * a.bin], r); atomicAdd(&S.cs[3 * a.bin + 1], g); atomicAdd(&S.cs[3 * a.bin + 2], b);
    }
    a.bin = -1; a.n = a.r = a.g = a.b = 0u;
}

__device__ __forceinline__ void add_pixel(ColorSmem &S, RunAcc &acc, bool on, uint32_t rr, uint32_t gg, uint32_t bb, int lane, uint32_t &cnt)
{
    const unsigned m = __ballot_sync(0xffffffffu, on);
    if (m == 0u) return;                                   // white / grey / black stretch: nothing to add
    cnt += __popc(m);
    const int bin = ((rr >> 4) << 8) | ((gg >> 4) << 4) | (bb >> 4);
    const int b0 = __shfl_sync(0xffffffffu, bin, __ffs(m) - 1);
    if (__ballot_sync(0xffffffffu, on && bin != b0) == 0u) {   // one bin for the whole group: extend (or open) the run
        if (b0 != acc.bin) { flush_run(S, acc, lane); acc.bin = b0; }
        if (on) { acc.n += 1u; acc.r += rr; acc.g += gg; acc.b += bb; }
    } else if (on) {                                       // mixed group: lanes holding the same bin elect a leader
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
        RunAcc acc; acc.bin = -1; acc.n = acc.r = acc.g = acc.b = 0u;
        for (int y = warp; y < t.height; y += CTHREADS / 32) {
            const uint8_t *row = src + (int64_t)y * t.row_stride;
            // four groups of 32 pixels per step: the loads of a step are issued before any of them is consumed
            for (int x0 = 0; x0 < W; x0 += 128) {
                uint32_t v[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int x = x0 + 32 * q + lane;
                    v[q] = 0u;                                       // outside the row: black, never masked in
                    if (x < W) {
                        if (t.channels == 4) v[q] = __ldg((const uint32_t *)row + x);      // RGBX words (4-byte aligned by contract)
                        else {
                            const uint8_t *p = row + 3 * x;
                            v[q] = (uint32_t)__ldg(p) | ((uint32_t)__ldg(p + 1) << 8) | ((uint32_t)__ldg(p + 2) << 16);
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (x0 + 32 * q >= W) break;                     // warp-uniform
                    const uint32_t rr = v[q] & 255u, gg = (v[q] >> 8) & 255u, bb = (v[q] >> 16) & 255u;
                    add_pixel(S, acc, hsv_mask_px(rr, gg, bb, S.sdiv), rr, gg, bb, lane, cnt);
                }
            }
        }
        flush_run(S, acc, lane);
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

    // ---- bin centroids in fixed point, in place of the channel sums (the sums are not needed any more) -----------
    for (int i = threadIdx.x; i < nb; i += CTHREADS) {
        const int bin = S.list[i];
        const u64 w = S.h[bin];
#pragma unroll
        for (int c = 0; c < 3; ++c) S.cs[3 * bin + c] = (uint32_t)((64ull * S.cs[3 * bin + c] + w / 2) / w);
    }
    __syncthreads();

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
            S.centre[j][threadIdx.x] = (int32_t)S.cs[3 * bin + threadIdx.x];
        }
    }
    __syncthreads();

    // ---- weighted Lloyd iterations over the bin centroids -------------------------------------------------------
    // Two passes per round.  Assign: a thread owns a bin and writes the index of its nearest centre into the four spare
    // bits of the bin's 16-bit list entry.  Accumulate, cluster-major: the 16 warps split into k groups, the warps of a
    // group share the list; a lane adds the members of its cluster it meets to 64-bit registers, the warp folds them
    // with shuffles once per round and leaves one partial per warp -- no atomics, no `redux` (the per-cluster REDUX
    // form spent 58 % of the kernel's stall samples on the uniform-register round trips).
    const int nparts = (CTHREADS / 32) / k;                // warps per cluster (>= 2 for k <= 8)
    for (int it = 0; it < iters; ++it) {
        if (threadIdx.x == 0) S.changed = 0;
        for (int i = threadIdx.x; i < nb; i += CTHREADS) {
            const int bin = S.list[i] & (CBINS - 1);
            const int p0 = (int)S.cs[3 * bin], p1 = (int)S.cs[3 * bin + 1], p2 = (int)S.cs[3 * bin + 2];
            int a = 0, dbest = 0x7fffffff;
            for (int j = 0; j < k; ++j) {
                const int d0 = p0 - S.centre[j][0], d1 = p1 - S.centre[j][1], d2 = p2 - S.centre[j][2];
                const int d = d0 * d0 + d1 * d1 + d2 * d2;
                if (d < dbest) { dbest = d; a = j; }
            }
            S.list[i] = (uint16_t)(bin | (a << 12));
        }
        __syncthreads();
        if (warp < k * nparts) {
            const int j = warp % k, part = warp / k;
            u64 N = 0, T0 = 0, T1 = 0, T2 = 0;
            for (int i = part * 32 + lane; i < nb; i += 32 * nparts) {
                const int e = S.list[i];
                if ((e >> 12) == j) {
                    const int bin = e & (CBINS - 1);
                    const uint32_t w = S.h[bin];
                    N += w; T0 += (u64)w * S.cs[3 * bin]; T1 += (u64)w * S.cs[3 * bin + 1]; T2 += (u64)w * S.cs[3 * bin + 2];
                }
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                N += __shfl_down_sync(0xffffffffu, N, d); T0 += __shfl_down_sync(0xffffffffu, T0, d);
                T1 += __shfl_down_sync(0xffffffffu, T1, d); T2 += __shfl_down_sync(0xffffffffu, T2, d);
            }
            if (lane == 0) { S.part[warp][0] = N; S.part[warp][1] = T0; S.part[warp][2] = T1; S.part[warp][3] = T2; }
        }
        __syncthreads();
        if (threadIdx.x < 3 * k) {
            const int j = threadIdx.x / 3, c = threadIdx.x % 3;
            u64 N = 0, T = 0;
            for (int q = 0; q < nparts; ++q) { N += S.part[j + k * q][0]; T += S.part[j + k * q][1 + c]; }
            if (c == 0) S.acc[j][0] = N;
            if (N) {
                const int32_t nc = (int32_t)((T + N / 2) / N);
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
        const u64 r = (u64)(S.centre[j][0] >> 6), g = (u64)(S.centre[j][1] >> 6), b = (u64)(S.centre[j][2] >> 6);
        o[2 + j] = (S.acc[j][0] << 24) | (r << 16) | (g << 8) | b;
    }
    if (threadIdx.x == 0) o[1] = (u64)k;
}

}  // namespace

extern "C" SYNSEG_EXPORT int synseg_colors_crops(synseg_ctx *ctx, const void *base, const synseg_crop *crops_host, int32_t n, int32_t n_colors,
                                                 int32_t iters, int32_t min_pixels, uint64_t *out, uint32_t *hist_out, void *stream)
{
    if (!ctx) { synseg_set_error("synseg_colors_crops: ctx is NULL"); return SYNSEG_E_INVALID; }
    SS_ENTER(ctx, stream);
    if (n <= 0) return SYNSEG_OK;
    if (!base || !crops_host || !out) { synseg_set_error("synseg_colors_crops: NULL argument"); return SYNSEG_E_INVALID; }
    if (n_colors < 1 || n_colors > CMAXK || iters < 0) { synseg_set_error("synseg_colors_crops: need 1 <= n_colors <= %d, iters >= 0", CMAXK); return SYNSEG_E_INVALID; }
    if (min_pixels < 1) min_pixels = 1;                 // a crop without a masked pixel has no colour
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
    SS_CUDA(cudaFuncSetAttribute(crop_colors_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ColorSmem)));   // per device; cheap
    crop_colors_kernel<<<n, CTHREADS, sizeof(ColorSmem), st>>>((const uint8_t *)base, (const CropTask *)p, n_colors, iters, min_pixels,
                                                             (u64 *)out, hist_out);
    SS_LAUNCH_CHECK(ctx, "crop_colors", st);
    return SYNSEG_OK;
}
