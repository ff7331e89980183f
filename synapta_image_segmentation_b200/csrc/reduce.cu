// reduce.cu -- exact integer reductions over regions: grey moments, HSV mask + colour histogram,
// ordered gather of masked pixels.
//
// Reference call sites:
//   np.sum(x > 0)                      pdf_image_segmentation.py:1371,1376,1439,1560,1561,1616
//   np.var(L)                          :1805 (>1500 photo), :2989 (<10 / >100), :3073 (>1000); old_algo:1007
//   RGB2HSV + S>30 & V>40 & V<240      :1571-1575 ; `len(pixels) < 100 -> []` :1577
//   img_array[mask] (raster order)     :1575 (the list np.random.choice samples from, :1581-1583)
//
// Variance is returned as exact integer moments (n, sum, sum of squares); the host forms
// (n*ss - s*s) / n^2 with big-integer arithmetic, so threshold decisions match np.var except within
// ~1e-9 of a threshold.
// Roofline: HBM-bound, 1 (grey) or 3 (RGB) algorithmic bytes per pixel.
#include "internal.cuh"
#include "pixel.cuh"

namespace {

__device__ __forceinline__ synseg_roi roi_of(const synseg_roi *rois, int i, int width, int height)
{
    if (rois) return rois[i];
    synseg_roi r; r.image = i; r.x = 0; r.y = 0; r.width = width; r.height = height;
    return r;
}

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
    return v;
}

// src_kind: 0 grey, 1 RGB via PIL grey, 2 RGB via cv2 grey
__global__ void __launch_bounds__(256) moments_kernel(Plane src, int width, int height, int src_kind, const synseg_roi *rois,
                                                      unsigned long long *out)
{
    const synseg_roi r = roi_of(rois, blockIdx.y, width, height);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint8_t *base = src.p + r.image * src.bs;
    unsigned long long s = 0, ss = 0, nz = 0;
    for (int y = blockIdx.x * 8 + warp; y < r.height; y += gridDim.x * 8) {
        const uint8_t *row = base + (int64_t)(r.y + y) * src.rs;
        unsigned int rs = 0, rss = 0, rnz = 0;     // per row: <= 65535 * 65025 / 32 lanes fits u32 per lane
        for (int x = lane; x < r.width; x += 32) {
            uint32_t v;
            if (src_kind == 0) v = __ldg(row + r.x + x);
            else {
                const uint8_t *p = row + 3 * (int64_t)(r.x + x);
                const uint32_t rgbx = (uint32_t)__ldg(p) | ((uint32_t)__ldg(p + 1) << 8) | ((uint32_t)__ldg(p + 2) << 16);
                v = gray_dyn(rgbx, src_kind == 1 ? SYNSEG_GRAY_PIL : SYNSEG_GRAY_CV);
            }
            rs += v; rss += v * v; rnz += (v != 0);
        }
        s += rs; ss += rss; nz += rnz;
    }
    s = warp_sum_u64(s); ss = warp_sum_u64(ss); nz = warp_sum_u64(nz);
    __shared__ unsigned long long sh[8][3];
    if (lane == 0) { sh[warp][0] = s; sh[warp][1] = ss; sh[warp][2] = nz; }
    __syncthreads();
    if (threadIdx.x < 3) {
        unsigned long long t = 0;
        for (int w = 0; w < 8; ++w) t += sh[w][threadIdx.x];
        if (t) atomicAdd(out + 3 * (int64_t)blockIdx.y + threadIdx.x, t);
    }
}

constexpr int HBINS = 4096;

template <bool WITH_SUMS>
__global__ void __launch_bounds__(256) hsv_hist_kernel(Plane src, int width, int height, const synseg_roi *rois,
                                                       unsigned long long *count, uint32_t *hist, unsigned long long *chan_sum,
                                                       uint32_t *row_count, int max_rows)
{
    extern __shared__ uint32_t smh[];
    uint32_t *sdiv = smh;                  // 256
    uint32_t *h = smh + 256;               // HBINS
    uint32_t *cs = h + HBINS;              // HBINS * 3 (WITH_SUMS)
    const synseg_roi r = roi_of(rois, blockIdx.y, width, height);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 256; i += 256) sdiv[i] = hsv_sdiv(i);
    for (int i = threadIdx.x; i < HBINS; i += 256) h[i] = 0;
    if (WITH_SUMS) for (int i = threadIdx.x; i < 3 * HBINS; i += 256) cs[i] = 0;
    __syncthreads();
    const uint8_t *base = src.p + r.image * src.bs;
    unsigned int total = 0;
    for (int y = blockIdx.x * 8 + warp; y < r.height; y += gridDim.x * 8) {
        const uint8_t *row = base + (int64_t)(r.y + y) * src.rs + 3 * (int64_t)r.x;
        unsigned int rc = 0;
        for (int x0 = 0; x0 < r.width; x0 += 32) {
            const int x = x0 + lane;
            bool on = false;
            uint32_t rr = 0, gg = 0, bb = 0;
            if (x < r.width) {
                const uint8_t *p = row + 3 * x;
                rr = __ldg(p); gg = __ldg(p + 1); bb = __ldg(p + 2);
                on = hsv_mask_px(rr, gg, bb, sdiv);
            }
            const unsigned m = __ballot_sync(0xffffffffu, on);
            rc += __popc(m);
            if (on && hist) {
                const int bin = ((rr >> 4) << 8) | ((gg >> 4) << 4) | (bb >> 4);
                const unsigned peers = __match_any_sync(m, bin);
                const bool leader = lane == __ffs(peers) - 1;
                if (WITH_SUMS) {
                    const unsigned sr = __reduce_add_sync(peers, rr), sg = __reduce_add_sync(peers, gg), sb = __reduce_add_sync(peers, bb);
                    if (leader) { atomicAdd(&cs[3 * bin], sr); atomicAdd(&cs[3 * bin + 1], sg); atomicAdd(&cs[3 * bin + 2], sb); }
                }
                if (leader) atomicAdd(&h[bin], (unsigned)__popc(peers));
            }
        }
        if (lane == 0 && row_count && y < max_rows) row_count[(int64_t)blockIdx.y * max_rows + y] = rc;
        total += rc;
    }
    if (lane == 0 && total) atomicAdd(count + blockIdx.y, (unsigned long long)total);
    __syncthreads();
    if (hist) {
        for (int i = threadIdx.x; i < HBINS; i += 256) {
            const uint32_t v = h[i];
            if (v) {
                atomicAdd(&hist[(int64_t)blockIdx.y * HBINS + i], v);
                if (WITH_SUMS) {
                    unsigned long long *c = chan_sum + ((int64_t)blockIdx.y * HBINS + i) * 3;
                    atomicAdd(c, (unsigned long long)cs[3 * i]); atomicAdd(c + 1, (unsigned long long)cs[3 * i + 1]);
                    atomicAdd(c + 2, (unsigned long long)cs[3 * i + 2]);
                }
            }
        }
    }
}

// one warp per requested rank
__global__ void __launch_bounds__(256) hsv_gather_kernel(Plane src, synseg_roi r, const unsigned long long *row_prefix, const long long *ranks,
                                                         int n, uint8_t *out)
{
    __shared__ uint32_t sdiv[256];
    for (int i = threadIdx.x; i < 256; i += 256) sdiv[i] = hsv_sdiv(i);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + warp;
    if (i >= n) return;
    const unsigned long long rank = (unsigned long long)ranks[i];
    // largest y with row_prefix[y] <= rank  (row_prefix has height + 1 entries, exclusive prefix)
    int lo = 0, hi = r.height;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (row_prefix[mid] <= rank) lo = mid; else hi = mid;
    }
    const int y = lo;
    unsigned int need = (unsigned int)(rank - row_prefix[y]);   // in-row rank
    const uint8_t *row = src.p + r.image * src.bs + (int64_t)(r.y + y) * src.rs + 3 * (int64_t)r.x;
    for (int x0 = 0; x0 < r.width; x0 += 32) {
        const int x = x0 + lane;
        bool on = false;
        uint32_t rr = 0, gg = 0, bb = 0;
        if (x < r.width) {
            const uint8_t *p = row + 3 * x;
            rr = __ldg(p); gg = __ldg(p + 1); bb = __ldg(p + 2);
            on = hsv_mask_px(rr, gg, bb, sdiv);
        }
        const unsigned m = __ballot_sync(0xffffffffu, on);
        const unsigned c = __popc(m);
        if (need < c) {
            const int sel = __fns(m, 0, need + 1);     // lane of the (need+1)-th set bit
            if (lane == sel) { out[3 * i] = (uint8_t)rr; out[3 * i + 1] = (uint8_t)gg; out[3 * i + 2] = (uint8_t)bb; }
            return;
        }
        need -= c;
    }
}

}  // namespace

int launch_moments(synseg_ctx *ctx, const synseg_img *src, int src_kind, const synseg_roi *rois, int32_t n_rois, uint64_t *out,
                   cudaStream_t st)
{
    SS_CUDA(cudaMemsetAsync(out, 0, sizeof(uint64_t) * 3 * (size_t)n_rois, st));
    int gx = rois ? 16 : cdiv(src->height, 8 * 4);
    if (gx < 1) gx = 1;
    if (gx > 1024) gx = 1024;
    moments_kernel<<<dim3(gx, n_rois), 256, 0, st>>>(plane_of(src), src->width, src->height, src_kind, rois, (unsigned long long *)out);
    SS_LAUNCH_CHECK(ctx, "moments", st);
    return SYNSEG_OK;
}

extern "C" SYNSEG_EXPORT int synseg_moments(synseg_ctx *ctx, const synseg_img *src, int src_kind, const synseg_roi *rois, int32_t n_rois,
                              uint64_t *out, void *stream)
{
    if (!ctx) { synseg_set_error("synseg_moments: ctx is NULL"); return SYNSEG_E_INVALID; }
    SS_ENTER(ctx, stream);
    if (src_kind < 0 || src_kind > 2) { synseg_set_error("synseg_moments: bad src_kind %d", src_kind); return SYNSEG_E_INVALID; }
    SS_TRY(validate_img(src, "src", src_kind ? 3 : 1));
    if (!out) { synseg_set_error("synseg_moments: out is NULL"); return SYNSEG_E_INVALID; }
    if (!rois) n_rois = src->batch;
    if (n_rois <= 0) return SYNSEG_OK;
    if (n_rois > 65535) { synseg_set_error("synseg_moments: more than 65535 regions per call"); return SYNSEG_E_INVALID; }
    return launch_moments(ctx, src, src_kind, rois, n_rois, out, (cudaStream_t)stream);
}

extern "C" SYNSEG_EXPORT int synseg_hsv_mask_hist(synseg_ctx *ctx, const synseg_img *rgb, const synseg_roi *rois, int32_t n_rois, uint64_t *count,
                                    uint32_t *hist, uint64_t *chan_sum, uint32_t *row_count, int32_t max_rows, void *stream)
{
    if (!ctx) { synseg_set_error("synseg_hsv_mask_hist: ctx is NULL"); return SYNSEG_E_INVALID; }
    SS_ENTER(ctx, stream);
    SS_TRY(validate_img(rgb, "rgb", 3));
    if (!count) { synseg_set_error("synseg_hsv_mask_hist: count is NULL"); return SYNSEG_E_INVALID; }
    if (chan_sum && !hist) { synseg_set_error("synseg_hsv_mask_hist: chan_sum needs hist"); return SYNSEG_E_INVALID; }
    if (!rois) n_rois = rgb->batch;
    if (n_rois <= 0) return SYNSEG_OK;
    if (n_rois > 65535) { synseg_set_error("synseg_hsv_mask_hist: more than 65535 regions per call"); return SYNSEG_E_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    SS_CUDA(cudaMemsetAsync(count, 0, sizeof(uint64_t) * (size_t)n_rois, st));
    if (hist) SS_CUDA(cudaMemsetAsync(hist, 0, sizeof(uint32_t) * HBINS * (size_t)n_rois, st));
    if (chan_sum) SS_CUDA(cudaMemsetAsync(chan_sum, 0, sizeof(uint64_t) * 3 * HBINS * (size_t)n_rois, st));
    if (row_count) SS_CUDA(cudaMemsetAsync(row_count, 0, sizeof(uint32_t) * (size_t)max_rows * n_rois, st));
    const int gx = rois ? 8 : (rgb->height >= 1024 ? 32 : 8);
    const size_t smem = (256 + HBINS + (chan_sum ? 3 * HBINS : 0)) * sizeof(uint32_t);
    if (!(ctx->attr_done & ATTR_HSV_HIST)) {
        SS_CUDA(cudaFuncSetAttribute(hsv_hist_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
        ctx->attr_done |= ATTR_HSV_HIST;
    }
    if (chan_sum)
        hsv_hist_kernel<true><<<dim3(gx, n_rois), 256, smem, st>>>(plane_of(rgb), rgb->width, rgb->height, rois, (unsigned long long *)count,
                                                                  hist, (unsigned long long *)chan_sum, row_count, max_rows);
    else
        hsv_hist_kernel<false><<<dim3(gx, n_rois), 256, smem, st>>>(plane_of(rgb), rgb->width, rgb->height, rois, (unsigned long long *)count,
                                                                   hist, nullptr, row_count, max_rows);
    SS_LAUNCH_CHECK(ctx, "hsv_hist", st);
    return SYNSEG_OK;
}

extern "C" SYNSEG_EXPORT int synseg_hsv_mask_gather(synseg_ctx *ctx, const synseg_img *rgb, const synseg_roi *roi_host, const uint64_t *row_prefix,
                                      const int64_t *ranks, int32_t n, uint8_t *out_rgb, void *stream)
{
    if (!ctx) { synseg_set_error("synseg_hsv_mask_gather: ctx is NULL"); return SYNSEG_E_INVALID; }
    SS_ENTER(ctx, stream);
    SS_TRY(validate_img(rgb, "rgb", 3));
    if (!row_prefix || !ranks || !out_rgb) { synseg_set_error("synseg_hsv_mask_gather: NULL buffer"); return SYNSEG_E_INVALID; }
    if (n <= 0) return SYNSEG_OK;
    synseg_roi r;
    if (roi_host) r = *roi_host; else { r.image = 0; r.x = 0; r.y = 0; r.width = rgb->width; r.height = rgb->height; }
    if (r.image < 0 || r.image >= rgb->batch || r.x < 0 || r.y < 0 || r.x + r.width > rgb->width || r.y + r.height > rgb->height) {
        synseg_set_error("synseg_hsv_mask_gather: region outside the image"); return SYNSEG_E_INVALID;
    }
    hsv_gather_kernel<<<cdiv(n, 8), 256, 0, (cudaStream_t)stream>>>(plane_of(rgb), r, (const unsigned long long *)row_prefix,
                                                                    (const long long *)ranks, n, out_rgb);
    SS_LAUNCH_CHECK(ctx, "hsv_gather", (cudaStream_t)stream);
    return SYNSEG_OK;
}
