// pipeline.cu -- the fused page pipeline and the per-crop grid-line pipeline.
//
// synseg_detect_pages: RGB pages -> component boxes, bit-exact with the cv2 chain of SURVEY.md 8(d)
//   grey(cv2) -> adaptiveThreshold(MEAN_C, BINARY_INV, bs, C) | Canny(lo,hi) -> dilate(k x k)
//   -> morphologyEx(CLOSE, k x k) -> connectedComponentsWithStats(8).
// It is the raster replacement for the inside of _detect_visual_regions / _detect_by_drawings
// (pdf_image_segmentation.py:3105-3146, 3511-3557): the boxes it returns play the role of the
// drawing-command rectangles and are filtered / merged on the host with the reference's own rules.
//
// HBM plan per page: the RGB page is read once (3 B/px) and the grey page written once (1 B/px);
// every mask after that is a bit plane (1/8 B/px): the adaptive threshold writes bits directly, the
// Canny hysteresis ORs its result into the same plane, the dilate (folded with the close's dilate into
// one (2k-1) pass) and the erode run on bit planes, and the labelling reads bits.  The bit planes of a
// 50-page batch (53 MB each) stay in the 126 MB L2; the grey plane (420 MB) is re-read from HBM by the two stencils.
// Stream plan: a batch is cut into page chunks that run as independent chains on the caller's stream and on side streams
// owned by the context (fork / join through events; SYNSEG_OVERLAP chunks on SYNSEG_STREAMS streams, default 3 on 3), so
// the latency-bound union-find of one chunk overlaps the issue-bound stencils of another.
//
// synseg_grid_counts: per crop, grey -> Canny(50,150) -> OPEN(kw x 1, it=2) / OPEN(1 x kh, it=2)
//   -> non-zero counts, i.e. _detect_grid (pdf_image_segmentation.py:1546-1564) and the visual part of
//   _detect_chart_subtype (:1365-1376); the Canny map can be handed back for the host-side consumers
//   (HoughLinesP :1327,1387,1701 and findContours :1762).
// synseg_hints_crops: the same per crop plus grey moments and the HSV mask count for a ragged batch of crops.
// synseg_detect_pages_host: pages in host memory -> tables in host memory (staging ring + copy stream inside).
#include "internal.cuh"

#include <stdlib.h>

#include <algorithm>

static size_t detect_scratch_bytes(int W, int H, int B, int max_labels, bool need_gray)
{
    const size_t gray = need_gray ? (size_t)align_up((size_t)W, 16) * H * B + 256 : 0;
    const size_t bits = 2 * ((size_t)bit_wpr(W) * H * B * 4 + 256);
    const size_t canny = canny_scratch_bytes(W, H, B) + 512;
    const size_t ccl = ccl_stats_scratch_bytes(W, H, B, max_labels) + 4096;
    return gray + bits + (canny > ccl ? canny : ccl) + 4096;
}

static int overlap_streams(synseg_ctx *ctx)
{
    if (ctx->ev_split_fork) return SYNSEG_OK;
    for (int i = 0; i < 3; ++i) {
        SS_CUDA(cudaStreamCreateWithFlags(&ctx->aux[i], cudaStreamNonBlocking));
        SS_CUDA(cudaEventCreateWithFlags(&ctx->ev_split_join[i], cudaEventDisableTiming));
    }
    SS_CUDA(cudaEventCreateWithFlags(&ctx->ev_split_fork, cudaEventDisableTiming));
    return SYNSEG_OK;
}

// One chain over the pages of `rgb` on stream `st`, scratch from the arena at `arena_base`.
// channels 1: `rgb` holds grey pages and the chain starts at the threshold (nothing is converted or copied).
static int detect_chain(synseg_ctx *ctx, const synseg_img *rgb, const synseg_detect_params *prm, const synseg_img *gray_out,
                        int32_t *n_labels, int32_t *stats, double *centroids, cudaStream_t st, size_t arena_base)
{
    const int W = rgb->width, H = rgb->height, B = rgb->batch;
    const bool grey_in = prm->channels == 1;
    arena_rebase(ctx, arena_base);
    void *p;
    synseg_img gray;
    if (grey_in) gray = *rgb;
    else if (gray_out) gray = *gray_out;
    else {
        gray = *rgb;
        gray.row_stride = (int64_t)align_up((size_t)W, 16);
        gray.batch_stride = gray.row_stride * H;
        SS_TRY(arena_alloc(ctx, (size_t)gray.batch_stride * B, &p, st));
        gray.data = p;
    }
    const int wpr = bit_wpr(W);
    const size_t plane_bytes = (size_t)wpr * H * B * 4;
    SS_TRY(arena_alloc(ctx, plane_bytes, &p, st));
    BitPlane cur{(uint32_t *)p, wpr, (int64_t)wpr * H};
    SS_TRY(arena_alloc(ctx, plane_bytes, &p, st));
    BitPlane other{(uint32_t *)p, wpr, (int64_t)wpr * H};

    const size_t mark = arena_mark(ctx);
    if (!grey_in && canny_rgb_supported(rgb, &gray)) {
        // one TMA-fed pass over the RGB bytes writes the grey plane and the Canny classes; the threshold reads the grey plane once
        SS_TRY(run_front_rgb(ctx, rgb, &gray, cur, prm->block_size, prm->C, prm->canny_lo, prm->canny_hi, st));
    } else {
        if (!grey_in) SS_TRY(launch_rgb2gray(ctx, rgb, &gray, SYNSEG_GRAY_CV, st));
        SS_TRY(launch_adaptive_mean(ctx, &gray, nullptr, cur, prm->block_size, prm->C, 1, st));
        SS_TRY(run_canny(ctx, &gray, nullptr, cur, /*or_bits=*/true, prm->canny_lo, prm->canny_hi, st));
    }
    arena_release(ctx, mark);
    // dilate(k) then close(k) = dilate(k), dilate(k), erode(k) = dilate(2k-1, anchor 2*(k/2)), erode(k)
    const int k = prm->k;
    bool fused = false;                 // all four 1-D passes in one shared-memory kernel when the geometry fits (morph_fused.cu)
    SS_TRY(launch_bit_dilate_erode(ctx, cur, other, W, H, B, 2 * (k - 1) + 1, 2 * (k / 2), k, k / 2, &fused, st));
    if (fused) { const BitPlane t = cur; cur = other; other = t; }
    else {
        SS_TRY(run_bitmorph(ctx, cur, other, W, H, B, SYNSEG_MORPH_DILATE, k, k, k / 2, k / 2, 2, st));
        SS_TRY(run_bitmorph(ctx, cur, other, W, H, B, SYNSEG_MORPH_ERODE, k, k, k / 2, k / 2, 1, st));
    }
    CclMask m; m.u8 = nullptr; m.bits = cur; m.width = W; m.height = H; m.batch = B;
    return run_ccl_stats(ctx, m, nullptr, n_labels, stats, centroids, prm->max_labels, st);
}

static int check_detect_args(const char *who, const synseg_img *pages, const synseg_detect_params *prm, const synseg_img *gray_out)
{
    if (prm->channels != 0 && prm->channels != 1 && prm->channels != 3) { synseg_set_error("%s: channels must be 3 (RGB) or 1 (grey)", who); return SYNSEG_E_INVALID; }
    SS_TRY(validate_img(pages, "pages", prm->channels == 1 ? 1 : 3));
    if (gray_out) {
        if (prm->channels == 1) { synseg_set_error("%s: gray_out is meaningless for grey pages", who); return SYNSEG_E_INVALID; }
        SS_TRY(validate_img(gray_out, "gray_out", 1));
        if (!same_shape(pages, gray_out)) { synseg_set_error("%s: gray_out shape mismatch", who); return SYNSEG_E_INVALID; }
    }
    if (prm->block_size < 3 || prm->block_size > 255 || !(prm->block_size & 1) || prm->k < 1 || prm->max_labels < 1 || prm->canny_lo < 0 ||
        prm->canny_hi < prm->canny_lo) {
        synseg_set_error("%s: bad parameters", who); return SYNSEG_E_INVALID;
    }
    if (pages->width > 32766 || pages->height > 32766 || pages->batch > 65535) { synseg_set_error("%s: image or batch too large", who); return SYNSEG_E_INVALID; }
    return SYNSEG_OK;
}

// The chains of one call; on return every side stream has been joined into `st` -- also when a launch failed in between.
static int detect_pages_impl(synseg_ctx *ctx, const synseg_img *rgb, const synseg_detect_params *prm, const synseg_img *gray_out,
                             int32_t *n_labels, int32_t *stats, double *centroids, cudaStream_t st, size_t extra_scratch)
{
    const int W = rgb->width, H = rgb->height, B = rgb->batch;
    const int chunks = ctx->prof_on ? 1 : ctx->overlap;      // per-kernel profiling needs every kernel alone on one stream
    const int ml = prm->max_labels;
    const bool need_gray = gray_out == nullptr && prm->channels != 1;
    if (chunks >= 2 && B >= 2 * chunks) {
        // Page chunks as independent chains, dealt round-robin to the caller's stream and the context's side streams, each
        // stream with its own part of the arena: the latency-bound labelling of one chunk runs beside the issue- / HBM-bound
        // front end of another (measured: 2.34 -> 2.18 ms per 50 pages with two chunks on two streams in round 1; 1.47 -> 1.39 with two, 1.34 with three chunks on three streams now).  The side streams
        // fork from and join into the caller's stream through events, so the call stays asynchronous on that stream.
        SS_TRY(overlap_streams(ctx));
        const int ns = chunks < ctx->overlap_streams ? chunks : ctx->overlap_streams;
        const int per = cdiv(B, chunks);
        const size_t region = align_up(detect_scratch_bytes(W, H, per, ml, need_gray) + SS_GUARD_SLACK, 256);
        const size_t total = ns * region;
        SS_TRY(arena_ensure(ctx, total > extra_scratch ? total : extra_scratch));
        SS_CUDA(cudaEventRecord(ctx->ev_split_fork, st));
        for (int i = 1; i < ns; ++i) SS_CUDA(cudaStreamWaitEvent(ctx->aux[i - 1], ctx->ev_split_fork, 0));
        int rc = SYNSEG_OK;
        for (int c = 0, p0 = 0; p0 < B && rc == SYNSEG_OK; ++c, p0 += per) {
            const int np = B - p0 < per ? B - p0 : per;
            const int lane = c % ns;
            synseg_img v = *rgb, g;
            v.data = (uint8_t *)rgb->data + (int64_t)p0 * rgb->batch_stride; v.batch = np;
            if (gray_out) { g = *gray_out; g.data = (uint8_t *)gray_out->data + (int64_t)p0 * gray_out->batch_stride; g.batch = np; }
            rc = detect_chain(ctx, &v, prm, gray_out ? &g : nullptr, n_labels + p0, stats + (size_t)p0 * ml * 5,
                              centroids ? centroids + (size_t)p0 * ml * 2 : nullptr, lane ? ctx->aux[lane - 1] : st, lane * region);
        }
        for (int i = 1; i < ns; ++i) {       // always join, so an error return never leaves a side stream running unordered
            const int e1 = synseg_check_cuda(cudaEventRecord(ctx->ev_split_join[i - 1], ctx->aux[i - 1]), "join record");
            const int e2 = e1 ? e1 : synseg_check_cuda(cudaStreamWaitEvent(st, ctx->ev_split_join[i - 1], 0), "join wait");
            if (rc == SYNSEG_OK && e2) rc = e2;
        }
        return rc;
    }
    const size_t need = detect_scratch_bytes(W, H, B, ml, need_gray);
    SS_TRY(arena_ensure(ctx, need > extra_scratch ? need : extra_scratch));
    return detect_chain(ctx, rgb, prm, gray_out, n_labels, stats, centroids, st, 0);
}

extern "C" SYNSEG_EXPORT int synseg_detect_pages(synseg_ctx *ctx, const synseg_img *rgb, const synseg_detect_params *prm, const synseg_img *gray_out,
                                   int32_t *n_labels, int32_t *stats, double *centroids, void *stream)
{
    if (!ctx || !prm) { synseg_set_error("synseg_detect_pages: NULL ctx/params"); return SYNSEG_E_INVALID; }
    SS_ENTER(ctx, stream);
    SS_TRY(check_detect_args("synseg_detect_pages", rgb, prm, gray_out));
    if (!n_labels || !stats || !centroids) { synseg_set_error("synseg_detect_pages: NULL result buffer"); return SYNSEG_E_INVALID; }
    return detect_pages_impl(ctx, rgb, prm, gray_out, n_labels, stats, centroids, (cudaStream_t)stream, 0);
}

// pages -> component tables -> candidate regions with crop moments, one stream, no host round trip
static int detect_regions_impl(synseg_ctx *ctx, const synseg_img *pages, const synseg_detect_params *prm, const synseg_region_params *rp,
                               int32_t *n_labels, int32_t *stats, double *centroids, synseg_region *regions, int32_t *n_regions, int32_t *flags,
                               cudaStream_t st)
{
    const int ch = prm->channels == 1 ? 1 : 3;
    SS_TRY(detect_pages_impl(ctx, pages, prm, nullptr, n_labels, stats, centroids, st, regions_scratch_bytes(pages->batch, prm->max_labels)));
    arena_begin(ctx);          // every chain has been joined into `st`: the region kernels follow them in stream order
    return run_regions(ctx, n_labels, stats, prm->max_labels, pages, ch, rp, regions, n_regions, flags, st);
}

extern "C" SYNSEG_EXPORT int synseg_detect_regions(synseg_ctx *ctx, const synseg_img *pages, const synseg_detect_params *prm,
                                                   const synseg_region_params *rp, int32_t *n_labels, int32_t *stats, double *centroids,
                                                   synseg_region *regions, int32_t *n_regions, int32_t *flags, void *stream)
{
    if (!ctx || !prm) { synseg_set_error("synseg_detect_regions: NULL ctx/params"); return SYNSEG_E_INVALID; }
    SS_ENTER(ctx, stream);
    SS_TRY(check_detect_args("synseg_detect_regions", pages, prm, nullptr));
    SS_TRY(validate_region_params(rp, "synseg_detect_regions"));
    if (!n_labels || !stats || !regions || !n_regions || !flags) { synseg_set_error("synseg_detect_regions: NULL result buffer"); return SYNSEG_E_INVALID; }
    return detect_regions_impl(ctx, pages, prm, rp, n_labels, stats, centroids, regions, n_regions, flags, (cudaStream_t)stream);
}

// grey -> Canny(50,150) -> OPEN(kw x 1, it=2) / OPEN(1 x kh, it=2) -> counts of ONE image (view->batch == 1).
// out3 = {h_count, v_count, edge_px} must be zero on entry.  The scratch arena must already be large enough.
static int grid_counts_one(synseg_ctx *ctx, const synseg_img *view, int channels, int gray_mode, int kw, int kh, uint64_t *out3,
                           const synseg_img *edges_out, cudaStream_t st)
{
    arena_begin(ctx);
    const int W = view->width, H = view->height;
    synseg_img gray = *view;
    void *p;
    if (channels == 3) {
        gray.row_stride = (int64_t)align_up((size_t)W, 16);
        gray.batch_stride = gray.row_stride * H;
        SS_TRY(arena_alloc(ctx, (size_t)gray.batch_stride, &p, st));
        gray.data = p;
        SS_TRY(launch_rgb2gray(ctx, view, &gray, gray_mode, st));
    }
    const int wpr = bit_wpr(W);
    const size_t pb = (size_t)wpr * H * 4;
    SS_TRY(arena_alloc(ctx, pb, &p, st)); BitPlane edges{(uint32_t *)p, wpr, (int64_t)wpr * H};
    SS_TRY(arena_alloc(ctx, pb, &p, st)); BitPlane a{(uint32_t *)p, wpr, (int64_t)wpr * H};
    SS_TRY(arena_alloc(ctx, pb, &p, st)); BitPlane b{(uint32_t *)p, wpr, (int64_t)wpr * H};
    SS_TRY(run_canny(ctx, &gray, nullptr, edges, false, 50, 150, st));
    SS_TRY(launch_count_bits(ctx, edges, W, H, 1, out3 + 2, 0, st));
    if (edges_out) SS_TRY(launch_unpack_bits(ctx, edges, edges_out, st));
    const int ekw = kw > 0 ? kw : (W / 20 > 20 ? W / 20 : 20);
    const int ekh = kh > 0 ? kh : (H / 20 > 20 ? H / 20 : 20);
    // horizontal lines: OPEN with (ekw x 1), iterations 2
    SS_CUDA(cudaMemcpyAsync(a.p, edges.p, pb, cudaMemcpyDeviceToDevice, st));
    {
        BitPlane c = a, o = b;
        SS_TRY(run_bitmorph(ctx, c, o, W, H, 1, SYNSEG_MORPH_OPEN, ekw, 1, ekw / 2, 0, 2, st));
        SS_TRY(launch_count_bits(ctx, c, W, H, 1, out3 + 0, 0, st));
    }
    // vertical lines: OPEN with (1 x ekh), iterations 2  (column passes ping-pong edges -> a -> b)
    {
        BitPlane c = edges, o = a;
        SS_TRY(run_bitmorph(ctx, c, o, W, H, 1, SYNSEG_MORPH_OPEN, 1, ekh, 0, ekh / 2, 2, st));
        SS_TRY(launch_count_bits(ctx, c, W, H, 1, out3 + 1, 0, st));
    }
    return SYNSEG_OK;
}

static size_t grid_counts_scratch(int mw, int mh)
{
    const size_t gray_bytes = (size_t)align_up((size_t)mw, 16) * mh + 256;
    const size_t plane_bytes = (size_t)bit_wpr(mw) * mh * 4 + 256;
    return gray_bytes + 3 * plane_bytes + canny_scratch_bytes(mw, mh, 1) + 8192;
}

extern "C" SYNSEG_EXPORT int synseg_grid_counts(synseg_ctx *ctx, const synseg_img *src, int channels, int gray_mode, const synseg_roi *rois_host,
                                  int32_t n_rois, int kw, int kh, uint64_t *out, const synseg_img *edges_out, void *stream)
{
    if (!ctx) { synseg_set_error("synseg_grid_counts: ctx is NULL"); return SYNSEG_E_INVALID; }
    SS_ENTER(ctx, stream);
    if (channels != 1 && channels != 3) { synseg_set_error("synseg_grid_counts: channels must be 1 or 3"); return SYNSEG_E_INVALID; }
    SS_TRY(validate_img(src, "src", channels));
    if (!out) { synseg_set_error("synseg_grid_counts: out is NULL"); return SYNSEG_E_INVALID; }
    if (edges_out) SS_TRY(validate_img(edges_out, "edges_out", 1));
    cudaStream_t st = (cudaStream_t)stream;
    synseg_roi whole;
    if (!rois_host) {
        if (src->batch != 1) { synseg_set_error("synseg_grid_counts: rois required for batched sources"); return SYNSEG_E_INVALID; }
        whole.image = 0; whole.x = 0; whole.y = 0; whole.width = src->width; whole.height = src->height;
        rois_host = &whole; n_rois = 1;
    }
    if (n_rois <= 0) return SYNSEG_OK;
    if (edges_out && edges_out->batch < n_rois) { synseg_set_error("synseg_grid_counts: edges_out batch < n_rois"); return SYNSEG_E_INVALID; }
    // scratch for the largest region
    int mw = 0, mh = 0;
    for (int i = 0; i < n_rois; ++i) {
        const synseg_roi &r = rois_host[i];
        if (r.image < 0 || r.image >= src->batch || r.x < 0 || r.y < 0 || r.width <= 0 || r.height <= 0 || r.x + r.width > src->width ||
            r.y + r.height > src->height) {
            synseg_set_error("synseg_grid_counts: region %d outside the image", i); return SYNSEG_E_INVALID;
        }
        if (edges_out && (r.width > edges_out->width || r.height > edges_out->height)) {
            synseg_set_error("synseg_grid_counts: region %d larger than edges_out", i); return SYNSEG_E_INVALID;
        }
        if (r.width > mw) mw = r.width;
        if (r.height > mh) mh = r.height;
    }
    SS_TRY(arena_ensure(ctx, grid_counts_scratch(mw, mh)));
    SS_CUDA(cudaMemsetAsync(out, 0, sizeof(uint64_t) * 3 * (size_t)n_rois, st));
    for (int i = 0; i < n_rois; ++i) {
        const synseg_roi &r = rois_host[i];
        synseg_img view = *src;
        view.data = (uint8_t *)src->data + r.image * src->batch_stride + (int64_t)r.y * src->row_stride + (int64_t)r.x * channels;
        view.width = r.width; view.height = r.height; view.batch = 1;
        synseg_img eo;
        if (edges_out) {
            eo = *edges_out;
            eo.data = (uint8_t *)edges_out->data + (int64_t)i * edges_out->batch_stride;
            eo.width = r.width; eo.height = r.height; eo.batch = 1;
        }
        SS_TRY(grid_counts_one(ctx, &view, channels, gray_mode, kw, kh, out + 3 * (size_t)i, edges_out ? &eo : nullptr, st));
    }
    return SYNSEG_OK;
}

// ---- ragged batches of crops ------------------------------------------------------------------------------------
// The crops are sorted by size and cut into chunks; a chunk is processed by ONE launch per stage on a canvas batch
// (planes sized for the tallest / widest crop of the chunk, per-crop width / height in a device table that the
// kernels consult: BitPlane::dims).  16 launches per chunk instead of ~25 per crop.
__global__ void __launch_bounds__(256) scatter_results_kernel(const uint64_t *sorted, const CropTask *tasks, int n, uint64_t *out)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= 8 * n) return;
    out[8 * (int64_t)tasks[i >> 3].out_index + (i & 7)] = sorted[i];
}

static size_t ragged_chunk_bytes(int wc, int hc, int n)
{
    const size_t gray = (size_t)align_up((size_t)wc, 16) * hc * n + 256;
    const size_t plane = (size_t)bit_wpr(wc) * hc * n * 4 + 256;
    return gray + 3 * plane + canny_scratch_bytes(wc, hc, n) + 4096;
}

constexpr size_t RAGGED_CHUNK_BUDGET = (size_t)1 << 30;    // scratch per chunk
constexpr int RAGGED_CHUNK_MAX = 1024;                      // crops per chunk

static int hints_crops_ragged(synseg_ctx *ctx, const void *base, const synseg_crop *crops, int n, int kw, int kh, uint64_t *out, cudaStream_t st)
{
    // order: by height, then width (chunks of similar crops waste few canvas rows)
    std::vector<int> order(n);
    for (int i = 0; i < n; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int a, int b) {
        if (crops[a].height != crops[b].height) return crops[a].height < crops[b].height;
        if (crops[a].width != crops[b].width) return crops[a].width < crops[b].width;
        return a < b;
    });
    std::vector<CropTask> tasks(n);
    std::vector<int2> dims(n);
    for (int j = 0; j < n; ++j) {
        const synseg_crop &c = crops[order[j]];
        tasks[j] = CropTask{(int64_t)c.offset, c.row_stride, c.width, c.height, c.channels, order[j]};
        dims[j] = make_int2(c.width, c.height);
    }
    // chunks: consecutive crops while the canvas fits the scratch budget
    struct Chunk { int first, count, wc, hc; };
    std::vector<Chunk> chunks;
    size_t worst = 0;
    int chunk_max = RAGGED_CHUNK_MAX;
    if (const char *e = getenv("SYNSEG_RAGGED_CHUNK")) { const int v = atoi(e); if (v >= 1 && v < chunk_max) chunk_max = v; }
    for (int j = 0; j < n;) {
        Chunk ch{j, 0, 0, 0};
        while (j < n && ch.count < chunk_max) {
            const int wc = tasks[j].width > ch.wc ? tasks[j].width : ch.wc, hc = tasks[j].height > ch.hc ? tasks[j].height : ch.hc;
            if (ch.count > 0 && ragged_chunk_bytes(wc, hc, ch.count + 1) > RAGGED_CHUNK_BUDGET) break;
            ch.wc = wc; ch.hc = hc; ++ch.count; ++j;
        }
        const size_t need = ragged_chunk_bytes(ch.wc, ch.hc, ch.count);
        if (need > worst) worst = need;
        chunks.push_back(ch);
    }
    const size_t table_bytes = align_up(sizeof(CropTask) * (size_t)n, 256) + align_up(sizeof(int2) * (size_t)n, 256) + align_up(64 * (size_t)n, 256);
    SS_TRY(arena_ensure(ctx, table_bytes + worst + 4096));
    arena_begin(ctx);
    void *p;
    SS_TRY(arena_alloc(ctx, sizeof(CropTask) * (size_t)n, &p, st)); CropTask *d_tasks = (CropTask *)p;
    SS_TRY(arena_alloc(ctx, sizeof(int2) * (size_t)n, &p, st)); int2 *d_dims = (int2 *)p;
    SS_TRY(arena_alloc(ctx, 64 * (size_t)n, &p, st)); uint64_t *res = (uint64_t *)p;
    // pageable sources: the driver stages the bytes before cudaMemcpyAsync returns, so the vectors may die with this call
    SS_CUDA(cudaMemcpyAsync(d_tasks, tasks.data(), sizeof(CropTask) * (size_t)n, cudaMemcpyHostToDevice, st));
    SS_CUDA(cudaMemcpyAsync(d_dims, dims.data(), sizeof(int2) * (size_t)n, cudaMemcpyHostToDevice, st));
    SS_CUDA(cudaMemsetAsync(res, 0, 64 * (size_t)n, st));
    const int fkw = 2 * (kw - 1) + 1, fkh = 2 * (kh - 1) + 1, fax = 2 * (kw / 2), fay = 2 * (kh / 2);      // OPEN, iterations = 2, folded
    const size_t mark = arena_mark(ctx);
    for (const Chunk &ch : chunks) {
        arena_release(ctx, mark);
        const int W = ch.wc, H = ch.hc, B = ch.count;
        synseg_img gray;
        gray.width = W; gray.height = H; gray.batch = B; gray._pad = 0;
        gray.row_stride = (int64_t)align_up((size_t)W, 16); gray.batch_stride = gray.row_stride * H;
        SS_TRY(arena_alloc(ctx, (size_t)gray.batch_stride * B, &p, st)); gray.data = p;
        const int wpr = bit_wpr(W);
        const size_t pb = (size_t)wpr * H * B * 4;
        const int2 *dm = d_dims + ch.first;
        uint64_t *r = res + 8 * (size_t)ch.first;
        SS_TRY(arena_alloc(ctx, pb, &p, st)); BitPlane edges{(uint32_t *)p, wpr, (int64_t)wpr * H, dm};
        SS_TRY(arena_alloc(ctx, pb, &p, st)); BitPlane a{(uint32_t *)p, wpr, (int64_t)wpr * H, dm};
        SS_TRY(arena_alloc(ctx, pb, &p, st)); BitPlane b{(uint32_t *)p, wpr, (int64_t)wpr * H, dm};
        SS_TRY(launch_crop_front(ctx, base, d_tasks + ch.first, B, &gray, r, st));
        SS_TRY(run_canny(ctx, &gray, nullptr, edges, false, 50, 150, st));
        SS_TRY(launch_count_bits(ctx, edges, W, H, B, r + 2, 8, st));
        // horizontal lines: erode then dilate with (fkw x 1); vertical lines with (1 x fkh); `edges` is never written
        if (fkw > 1) {
            SS_TRY(launch_bitmorph_h(ctx, edges, a, W, H, B, SYNSEG_MORPH_ERODE, fkw, fax, st));
            SS_TRY(launch_bitmorph_h(ctx, a, b, W, H, B, SYNSEG_MORPH_DILATE, fkw, fax, st));
        }
        SS_TRY(launch_count_bits(ctx, fkw > 1 ? b : edges, W, H, B, r + 0, 8, st));
        if (fkh > 1) {
            SS_TRY(launch_bitmorph_v(ctx, edges, a, W, H, B, SYNSEG_MORPH_ERODE, fkh, fay, st));
            SS_TRY(launch_bitmorph_v(ctx, a, b, W, H, B, SYNSEG_MORPH_DILATE, fkh, fay, st));
        }
        SS_TRY(launch_count_bits(ctx, fkh > 1 ? b : edges, W, H, B, r + 1, 8, st));
    }
    scatter_results_kernel<<<cdiv(8 * (int64_t)n, 256), 256, 0, st>>>(res, d_tasks, n, out);
    SS_LAUNCH_CHECK(ctx, "scatter_results", st);
    return SYNSEG_OK;
}

// Batched per-crop hint quantities for n crops of different sizes packed in one device buffer (config 4:
// 10k cropped figure regions).  No host synchronisation: everything is queued on `stream`.
extern "C" SYNSEG_EXPORT int synseg_hints_crops(synseg_ctx *ctx, const void *base, const synseg_crop *crops_host, int32_t n, int kw, int kh,
                                  uint64_t *out, void *stream)
{
    if (!ctx) { synseg_set_error("synseg_hints_crops: ctx is NULL"); return SYNSEG_E_INVALID; }
    SS_ENTER(ctx, stream);
    if (n <= 0) return SYNSEG_OK;
    if (!base || !crops_host || !out) { synseg_set_error("synseg_hints_crops: NULL argument"); return SYNSEG_E_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    int mw = 0, mh = 0;
    bool has_rgbx = false;
    for (int i = 0; i < n; ++i) {
        const synseg_crop &c = crops_host[i];
        if (c.width <= 0 || c.height <= 0 || (c.channels != 1 && c.channels != 3 && c.channels != 4) || c.row_stride < (int64_t)c.width * c.channels) {
            synseg_set_error("synseg_hints_crops: bad crop %d", i); return SYNSEG_E_INVALID;
        }
        if (c.channels == 4) {
            has_rgbx = true;
            if ((((uintptr_t)base + c.offset) | (uint64_t)c.row_stride) & 3) { synseg_set_error("synseg_hints_crops: RGBX crop %d is not 4-byte aligned", i); return SYNSEG_E_INVALID; }
        }
        if (c.width > mw) mw = c.width;
        if (c.height > mh) mh = c.height;
    }
    if (mw > 32766 || mh > 32766) { synseg_set_error("synseg_hints_crops: crop larger than 32766"); return SYNSEG_E_INVALID; }
    // fixed structuring elements within the register / van Herk kernels' range: ragged batches, one launch per stage and chunk
    if (kw >= 1 && kh >= 1 && 2 * (kw - 1) + 1 <= 226 && 2 * (kh - 1) + 1 <= 384 && !getenv("SYNSEG_HINTS_PER_CROP"))
        return hints_crops_ragged(ctx, base, crops_host, n, kw, kh, out, st);
    // per-image structuring elements (kw / kh <= 0: the chart rule max(20, W / 20)) or very long ones: crop by crop
    if (has_rgbx) { synseg_set_error("synseg_hints_crops: RGBX crops need fixed structuring elements (1 <= kw <= 113, 1 <= kh <= 192)"); return SYNSEG_E_INVALID; }
    SS_TRY(arena_ensure(ctx, grid_counts_scratch(mw, mh)));
    SS_CUDA(cudaMemsetAsync(out, 0, sizeof(uint64_t) * 8 * (size_t)n, st));
    for (int i = 0; i < n; ++i) {
        const synseg_crop &c = crops_host[i];
        synseg_img view;
        view.data = (uint8_t *)base + c.offset; view.width = c.width; view.height = c.height; view.row_stride = c.row_stride;
        view.batch = 1; view._pad = 0; view.batch_stride = c.row_stride * c.height;
        uint64_t *o = out + 8 * (size_t)i;
        SS_TRY(grid_counts_one(ctx, &view, c.channels, SYNSEG_GRAY_PIL, kw, kh, o, nullptr, st));
        SS_TRY(launch_moments(ctx, &view, c.channels == 3 ? 1 : 0, nullptr, 1, o + 3, st));
        if (c.channels == 3) SS_TRY(synseg_hsv_mask_hist(ctx, &view, nullptr, 1, o + 6, nullptr, nullptr, nullptr, 0, stream));
    }
    return SYNSEG_OK;
}

// ---- hints on regions of pages that are already in HBM ----------------------------------------------------------
// The detector knows every region's crop rectangle (synseg_region px/py/pw/ph) and the pages are still resident: the hint
// kernels read the regions in place (a crop descriptor = byte offset + the page's row stride) instead of taking re-uploaded crops.
static int rois_to_crops(const char *who, const synseg_img *pages, int channels, const synseg_roi *rois, int32_t n, std::vector<synseg_crop> &crops)
{
    crops.resize(n);
    for (int i = 0; i < n; ++i) {
        const synseg_roi &r = rois[i];
        if (r.image < 0 || r.image >= pages->batch || r.x < 0 || r.y < 0 || r.width <= 0 || r.height <= 0 || r.x + r.width > pages->width ||
            r.y + r.height > pages->height) {
            synseg_set_error("%s: region %d outside the pages", who, i); return SYNSEG_E_INVALID;
        }
        synseg_crop &c = crops[i];
        c.offset = (uint64_t)((int64_t)r.image * pages->batch_stride + (int64_t)r.y * pages->row_stride + (int64_t)r.x * channels);
        c.width = r.width; c.height = r.height; c.row_stride = pages->row_stride; c.channels = channels; c._pad = 0;
    }
    return SYNSEG_OK;
}

extern "C" SYNSEG_EXPORT int synseg_hints_rois(synseg_ctx *ctx, const synseg_img *pages, int channels, const synseg_roi *rois_host, int32_t n, int kw, int kh,
                                               uint64_t *out, void *stream)
{
    if (!ctx) { synseg_set_error("synseg_hints_rois: ctx is NULL"); return SYNSEG_E_INVALID; }
    SS_ENTER(ctx, stream);
    if (channels != 1 && channels != 3) { synseg_set_error("synseg_hints_rois: channels must be 1 or 3"); return SYNSEG_E_INVALID; }
    SS_TRY(validate_img(pages, "pages", channels));
    if (n <= 0) return SYNSEG_OK;
    if (!rois_host || !out) { synseg_set_error("synseg_hints_rois: NULL argument"); return SYNSEG_E_INVALID; }
    std::vector<synseg_crop> crops;
    SS_TRY(rois_to_crops("synseg_hints_rois", pages, channels, rois_host, n, crops));
    return synseg_hints_crops(ctx, pages->data, crops.data(), n, kw, kh, out, stream);
}

extern "C" SYNSEG_EXPORT int synseg_colors_rois(synseg_ctx *ctx, const synseg_img *pages, const synseg_roi *rois_host, int32_t n, int32_t n_colors,
                                                int32_t iters, int32_t min_pixels, uint64_t *out, uint32_t *hist_out, void *stream)
{
    if (!ctx) { synseg_set_error("synseg_colors_rois: ctx is NULL"); return SYNSEG_E_INVALID; }
    SS_ENTER(ctx, stream);
    SS_TRY(validate_img(pages, "pages", 3));
    if (n <= 0) return SYNSEG_OK;
    if (!rois_host || !out) { synseg_set_error("synseg_colors_rois: NULL argument"); return SYNSEG_E_INVALID; }
    std::vector<synseg_crop> crops;
    SS_TRY(rois_to_crops("synseg_colors_rois", pages, 3, rois_host, n, crops));
    return synseg_colors_crops(ctx, pages->data, crops.data(), n, n_colors, iters, min_pixels, out, hist_out, stream);
}

// ---- host-buffer entry points ---------------------------------------------------------------------------------
#include <sys/syscall.h>
#include <unistd.h>

static void page_slots_free(synseg_ctx *ctx)
{
    synseg_ctx::HostStream &h = ctx->hs;
    for (int i = 0; i < synseg_ctx::HostStream::MAXS; ++i) {
        synseg_ctx::HostStream::PageSlot &q = h.ps[i];
        if (q.pages) cudaFreeHost(q.pages);
        if (q.ints) cudaFreeHost(q.ints);
        if (q.stats) cudaFreeHost(q.stats);
        if (q.regions) cudaFreeHost(q.regions);
        if (q.finished) cudaEventDestroy(q.finished);
        memset(&q, 0, sizeof(q));
    }
    h.ps_n = 0;
}

void host_stream_release(synseg_ctx *ctx)
{
    synseg_ctx::HostStream &h = ctx->hs;
    page_slots_free(ctx);
    if (!h.ready) return;
    for (int i = 0; i < synseg_ctx::HostStream::MAXS; ++i) {
        if (h.pages[i]) cudaFree(h.pages[i]);
        if (h.raw[i]) cudaFree(h.raw[i]);
        if (h.n_labels[i]) cudaFree(h.n_labels[i]);
        if (h.stats[i]) cudaFree(h.stats[i]);
        if (h.n_regions[i]) cudaFree(h.n_regions[i]);
        if (h.centroids[i]) cudaFree(h.centroids[i]);
        if (h.regions[i]) cudaFree(h.regions[i]);
        if (h.copied[i]) cudaEventDestroy(h.copied[i]);
        if (h.done[i]) cudaEventDestroy(h.done[i]);
    }
    if (h.copy) cudaStreamDestroy(h.copy);
    memset(&h, 0, sizeof(h));
}

constexpr int RING = 3;

static int host_stream_prepare(synseg_ctx *ctx, size_t slot_bytes, int slot_pages, int max_labels, int max_regions, size_t raw_bytes)
{
    synseg_ctx::HostStream &h = ctx->hs;
    if (h.ready && h.slot_bytes >= slot_bytes && h.slot_pages >= slot_pages && h.max_labels == max_labels && h.max_regions >= max_regions &&
        h.raw_bytes >= raw_bytes) return SYNSEG_OK;
    if (h.ps_n) { synseg_set_error("host staging: the page slots are initialised with another geometry (release them first)"); return SYNSEG_E_INVALID; }
    SS_CUDA(cudaDeviceSynchronize());
    if (h.ready) {
        if (h.slot_bytes > slot_bytes) slot_bytes = h.slot_bytes;
        if (h.slot_pages > slot_pages) slot_pages = h.slot_pages;
        if (h.max_regions > max_regions) max_regions = h.max_regions;
        if (h.raw_bytes > raw_bytes) raw_bytes = h.raw_bytes;
    }
    host_stream_release(ctx);
    if (max_regions < 1) max_regions = 1;
    SS_CUDA(cudaStreamCreateWithFlags(&h.copy, cudaStreamNonBlocking));
    for (int i = 0; i < RING; ++i) {
        SS_CUDA(cudaMalloc(&h.pages[i], slot_bytes));
        if (raw_bytes) SS_CUDA(cudaMalloc(&h.raw[i], raw_bytes));
        SS_CUDA(cudaMalloc(&h.n_labels[i], sizeof(int32_t) * slot_pages));
        SS_CUDA(cudaMalloc(&h.stats[i], sizeof(int32_t) * 5 * (size_t)slot_pages * max_labels));
        SS_CUDA(cudaMalloc(&h.centroids[i], sizeof(double) * 2 * (size_t)slot_pages * max_labels));
        SS_CUDA(cudaMalloc(&h.n_regions[i], sizeof(int32_t) * 2 * slot_pages));
        SS_CUDA(cudaMalloc(&h.regions[i], sizeof(synseg_region) * (size_t)slot_pages * max_regions));
        SS_CUDA(cudaEventCreateWithFlags(&h.copied[i], cudaEventDisableTiming));
        SS_CUDA(cudaEventCreateWithFlags(&h.done[i], cudaEventDisableTiming));
    }
    h.slot_bytes = slot_bytes; h.slot_pages = slot_pages; h.max_labels = max_labels; h.max_regions = max_regions; h.raw_bytes = raw_bytes; h.ready = true;
    return SYNSEG_OK;
}

// Row stride of staged pages on the device: RGB rows keep the host stride (rgb2gray takes any); grey rows are padded to
// 16 bytes so that the stencils read them with 128-bit loads.
static int64_t device_row_stride(int width, int channels, int64_t host_row_stride)
{
    return channels == 1 ? (int64_t)align_up((size_t)width, 16) : host_row_stride;
}

// Grey pages whose host layout is contiguous but not 16-byte pitched take the raw route: ONE 1-D copy into raw[s] on the
// copy stream (a 2-D H2D copy of 2550-byte rows reaches a third of the PCIe rate: 165,000 row descriptors per 50 pages),
// then a device-to-device 2-D copy into the pitched slot on the compute stream (HBM speed, ~0.15 ms per 50 pages).
static bool raw_route(int width, int channels, int64_t row_stride, int64_t page_stride, int height)
{
    return channels == 1 && row_stride != device_row_stride(width, channels, row_stride) && page_stride == row_stride * height;
}

// np pages host -> device staging slot s on the copy stream
static int stage_pages(synseg_ctx *ctx, int s, const uint8_t *src, int width, int height, int channels, int64_t row_stride,
                       int64_t page_stride, int np)
{
    synseg_ctx::HostStream &h = ctx->hs;
    cudaStream_t cs = h.copy;
    uint8_t *dst = h.pages[s];
    const int64_t dr = device_row_stride(width, channels, row_stride);
    const size_t dpage = (size_t)dr * height;
    if (row_stride == dr && page_stride == (int64_t)dpage) SS_CUDA(cudaMemcpyAsync(dst, src, dpage * np, cudaMemcpyHostToDevice, cs));
    else if (row_stride == dr) SS_CUDA(cudaMemcpy2DAsync(dst, dpage, src, (size_t)page_stride, dpage, np, cudaMemcpyHostToDevice, cs));
    else if (raw_route(width, channels, row_stride, page_stride, height))
        SS_CUDA(cudaMemcpyAsync(h.raw[s], src, (size_t)page_stride * np, cudaMemcpyHostToDevice, cs));       // re-pitched in run_ring_slot
    else
        for (int i = 0; i < np; ++i)
            SS_CUDA(cudaMemcpy2DAsync(dst + i * dpage, (size_t)dr, src + (int64_t)i * page_stride, (size_t)row_stride, (size_t)width * channels, height,
                                      cudaMemcpyHostToDevice, cs));
    return SYNSEG_OK;
}

struct HostOut {           // HOST result arrays of a job (any of them may be NULL), indexed from the job's first page
    int32_t *n_labels, *stats;
    double *centroids;
    synseg_region *regions;
    int32_t *n_regions, *flags;
};

// One chunk of pages already resident in ring slot s (its `copied` event recorded on the copy stream): pipeline + D2H on `st`.
static int run_ring_slot(synseg_ctx *ctx, int s, int width, int height, int channels, int64_t host_row_stride, int64_t host_page_stride, int np,
                         const synseg_detect_params *prm,
                         const synseg_region_params *rp, const HostOut &o, int p0, cudaStream_t st)
{
    synseg_ctx::HostStream &h = ctx->hs;
    const int ml = prm->max_labels;
    SS_CUDA(cudaStreamWaitEvent(st, h.copied[s], 0));
    if (raw_route(width, channels, host_row_stride, host_page_stride, height))
        SS_CUDA(cudaMemcpy2DAsync(h.pages[s], (size_t)device_row_stride(width, channels, host_row_stride), h.raw[s], (size_t)host_row_stride, (size_t)width,
                                  (size_t)height * np, cudaMemcpyDeviceToDevice, st));
    synseg_img img;
    img.data = h.pages[s]; img.width = width; img.height = height; img.row_stride = device_row_stride(width, channels, host_row_stride);
    img.batch = np; img._pad = 0; img.batch_stride = img.row_stride * height;
    double *cen = o.centroids ? h.centroids[s] : nullptr;
    if (rp) SS_TRY(detect_regions_impl(ctx, &img, prm, rp, h.n_labels[s], h.stats[s], cen, h.regions[s], h.n_regions[s], h.n_regions[s] + np, st));
    else SS_TRY(detect_pages_impl(ctx, &img, prm, nullptr, h.n_labels[s], h.stats[s], cen, st, 0));
    if (o.n_labels) SS_CUDA(cudaMemcpyAsync(o.n_labels + p0, h.n_labels[s], sizeof(int32_t) * np, cudaMemcpyDeviceToHost, st));
    if (o.stats) SS_CUDA(cudaMemcpyAsync(o.stats + (size_t)p0 * ml * 5, h.stats[s], sizeof(int32_t) * 5 * (size_t)np * ml, cudaMemcpyDeviceToHost, st));
    if (o.centroids) SS_CUDA(cudaMemcpyAsync(o.centroids + (size_t)p0 * ml * 2, h.centroids[s], sizeof(double) * 2 * (size_t)np * ml, cudaMemcpyDeviceToHost, st));
    if (rp) {
        SS_CUDA(cudaMemcpyAsync(o.n_regions + p0, h.n_regions[s], sizeof(int32_t) * np, cudaMemcpyDeviceToHost, st));
        SS_CUDA(cudaMemcpyAsync(o.flags + p0, h.n_regions[s] + np, sizeof(int32_t) * np, cudaMemcpyDeviceToHost, st));
        SS_CUDA(cudaMemcpyAsync(o.regions + (size_t)p0 * rp->max_regions, h.regions[s], sizeof(synseg_region) * (size_t)np * rp->max_regions,
                                cudaMemcpyDeviceToHost, st));
    }
    SS_CUDA(cudaEventRecord(h.done[s], st));
    h.used[s] = true;
    return SYNSEG_OK;
}

// Pages in HOST memory (pinned for full PCIe speed and true overlap; pageable works, slower) -> results in HOST memory.
// The library stages the pages through three device slots of `chunk_pages` pages: a copy stream moves chunk i+1, i+2
// while the pipeline runs on chunk i and the results of chunk i-1 travel back.  Returns after everything is queued;
// `stream` waits for the last chunk, so synchronising `stream` completes the call.
static int host_pipeline(synseg_ctx *ctx, const char *who, const void *host_pages, int32_t width, int32_t height, int64_t row_stride, int64_t page_stride,
                         int32_t n_pages, const synseg_detect_params *prm, const synseg_region_params *rp, int32_t chunk_pages, const HostOut &o,
                         cudaStream_t st)
{
    const int ch = prm->channels == 1 ? 1 : 3;
    if (prm->channels != 0 && prm->channels != 1 && prm->channels != 3) { synseg_set_error("%s: channels must be 3 (RGB) or 1 (grey)", who); return SYNSEG_E_INVALID; }
    if (width <= 0 || height <= 0 || n_pages < 0 || chunk_pages <= 0 || row_stride < ch * (int64_t)width || page_stride < row_stride * height || prm->max_labels < 1) {
        synseg_set_error("%s: bad shape / strides", who); return SYNSEG_E_INVALID;
    }
    if (n_pages == 0) return SYNSEG_OK;
    if (chunk_pages > n_pages) chunk_pages = n_pages;
    const size_t dpage = (size_t)device_row_stride(width, ch, row_stride) * height;
    SS_TRY(host_stream_prepare(ctx, dpage * chunk_pages, chunk_pages, prm->max_labels, rp ? rp->max_regions : 1,
                               raw_route(width, ch, row_stride, page_stride, height) ? (size_t)page_stride * chunk_pages : 0));
    synseg_ctx::HostStream &h = ctx->hs;
    // ring slots are private to the host entry points and every use ends with the slot's `done` event on the stream that
    // computed on it; the copy stream waits for exactly that event before it overwrites the slot -- also across calls
    const int n_chunks = cdiv(n_pages, chunk_pages);
    for (int c = 0; c < n_chunks; ++c) {
        const int s = h.next; h.next = (h.next + 1) % RING;
        const int p0 = c * chunk_pages, np = (n_pages - p0 < chunk_pages) ? n_pages - p0 : chunk_pages;
        if (h.used[s]) SS_CUDA(cudaStreamWaitEvent(h.copy, h.done[s], 0));       // slot s free again (the chunk that used it last has finished)
        SS_TRY(stage_pages(ctx, s, (const uint8_t *)host_pages + (int64_t)p0 * page_stride, width, height, ch, row_stride, page_stride, np));
        SS_CUDA(cudaEventRecord(h.copied[s], h.copy));
        SS_TRY(run_ring_slot(ctx, s, width, height, ch, row_stride, page_stride, np, prm, rp, o, p0, st));
    }
    return SYNSEG_OK;
}

extern "C" SYNSEG_EXPORT int synseg_detect_pages_host(synseg_ctx *ctx, const void *host_rgb, int32_t width, int32_t height,
                                        int64_t row_stride, int64_t page_stride, int32_t n_pages, const synseg_detect_params *prm,
                                        int32_t chunk_pages, int32_t *n_labels_host, int32_t *stats_host, double *centroids_host,
                                        void *stream)
{
    if (!ctx || !prm || !host_rgb || !n_labels_host || !stats_host) { synseg_set_error("synseg_detect_pages_host: NULL argument"); return SYNSEG_E_INVALID; }
    SS_ENTER(ctx, stream);
    const HostOut o{n_labels_host, stats_host, centroids_host, nullptr, nullptr, nullptr};
    return host_pipeline(ctx, "synseg_detect_pages_host", host_rgb, width, height, row_stride, page_stride, n_pages, prm, nullptr, chunk_pages, o,
                         (cudaStream_t)stream);
}

extern "C" SYNSEG_EXPORT int synseg_detect_regions_host(synseg_ctx *ctx, const void *host_pages, int32_t width, int32_t height, int64_t row_stride,
                                                        int64_t page_stride, int32_t n_pages, const synseg_detect_params *prm,
                                                        const synseg_region_params *rp, int32_t chunk_pages, int32_t *n_labels_host,
                                                        int32_t *stats_host, synseg_region *regions_host, int32_t *n_regions_host,
                                                        int32_t *flags_host, void *stream)
{
    if (!ctx || !prm || !host_pages || !regions_host || !n_regions_host || !flags_host) {
        synseg_set_error("synseg_detect_regions_host: NULL argument"); return SYNSEG_E_INVALID;
    }
    SS_ENTER(ctx, stream);
    SS_TRY(validate_region_params(rp, "synseg_detect_regions_host"));
    const HostOut o{n_labels_host, stats_host, nullptr, regions_host, n_regions_host, flags_host};
    return host_pipeline(ctx, "synseg_detect_regions_host", host_pages, width, height, row_stride, page_stride, n_pages, prm, rp, chunk_pages, o,
                         (cudaStream_t)stream);
}

// ---- renderer-facing page slots ----------------------------------------------------------------------------------
// NUMA node of the GPU (sysfs), -1 when unknown or the machine has a single node.
static int gpu_numa_node(int device)
{
    char bus[32] = "";
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) { cudaGetLastError(); return -1; }
    for (char *c = bus; *c; ++c) if (*c >= 'A' && *c <= 'Z') *c += 'a' - 'A';
    char path[128];
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    if (node < 0 || node >= 1024) return -1;
    if (access("/sys/devices/system/node/node1", F_OK) != 0) return -1;     // single node: nothing to choose
    return node;
}

// Pinned allocation whose pages are placed on `node` (set_mempolicy(MPOL_PREFERRED) around the allocation and the first
// touch; cudaHostAlloc pins the pages where they are).  node < 0: plain cudaHostAlloc.
static int pinned_alloc_on_node(void **out, size_t bytes, int node)
{
    unsigned long mask[16] = {0};
    bool bound = false;
    if (node >= 0) {
        mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
        bound = syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, mask, sizeof(mask) * 8) == 0;
    }
    const cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocDefault);
    if (bound) syscall(SYS_set_mempolicy, 0 /* MPOL_DEFAULT */, nullptr, 0);
    return synseg_check_cuda(e, "cudaHostAlloc(page slot)");
}

extern "C" SYNSEG_EXPORT int synseg_page_slots_release(synseg_ctx *ctx)
{
    if (!ctx) { synseg_set_error("synseg_page_slots_release: ctx is NULL"); return SYNSEG_E_INVALID; }
    DeviceScope scope(ctx->device);
    if (ctx->hs.ps_n) { cudaDeviceSynchronize(); page_slots_free(ctx); }
    return SYNSEG_OK;
}

extern "C" SYNSEG_EXPORT int synseg_page_slots_numa_node(const synseg_ctx *ctx) { return ctx && ctx->hs.ps_n ? ctx->hs.ps_numa : -1; }

extern "C" SYNSEG_EXPORT int synseg_page_slots_init(synseg_ctx *ctx, int32_t width, int32_t height, int32_t channels, int32_t pages_per_slot,
                                                    int32_t n_slots, int32_t max_labels, int32_t max_regions)
{
    if (!ctx) { synseg_set_error("synseg_page_slots_init: ctx is NULL"); return SYNSEG_E_INVALID; }
    if (width <= 0 || height <= 0 || (channels != 1 && channels != 3) || pages_per_slot < 1 || n_slots < 2 || n_slots > RING || max_labels < 1 ||
        max_regions < 1 || max_regions > 1024) {
        synseg_set_error("synseg_page_slots_init: bad arguments (channels 1 or 3, 2 <= n_slots <= %d)", RING); return SYNSEG_E_INVALID;
    }
    DeviceScope scope(ctx->device);
    synseg_ctx::HostStream &h = ctx->hs;
    if (h.ps_n) { cudaDeviceSynchronize(); page_slots_free(ctx); }
    const int64_t rs = (int64_t)align_up((size_t)width * channels, 16);
    const int64_t page = rs * height;
    SS_TRY(host_stream_prepare(ctx, (size_t)device_row_stride(width, channels, rs) * height * pages_per_slot, pages_per_slot, max_labels, max_regions, 0));
    const int node = gpu_numa_node(ctx->device);
    for (int i = 0; i < n_slots; ++i) {
        synseg_ctx::HostStream::PageSlot &q = h.ps[i];
        void *p;
        SS_TRY(pinned_alloc_on_node(&p, (size_t)page * pages_per_slot, node)); q.pages = (uint8_t *)p;
        SS_TRY(pinned_alloc_on_node(&p, sizeof(int32_t) * 3 * (size_t)pages_per_slot, node)); q.ints = (int32_t *)p;
        SS_TRY(pinned_alloc_on_node(&p, sizeof(int32_t) * 5 * (size_t)pages_per_slot * max_labels, node)); q.stats = (int32_t *)p;
        SS_TRY(pinned_alloc_on_node(&p, sizeof(synseg_region) * (size_t)pages_per_slot * max_regions, node)); q.regions = (synseg_region *)p;
        SS_CUDA(cudaEventCreateWithFlags(&q.finished, cudaEventDisableTiming));
        q.state = 0; q.n_pages = 0;
    }
    h.ps_n = n_slots; h.ps_next = 0; h.ps_width = width; h.ps_height = height; h.ps_channels = channels; h.ps_pages = pages_per_slot;
    h.ps_max_labels = max_labels; h.ps_max_regions = max_regions; h.ps_numa = node; h.ps_row_stride = rs; h.ps_page_stride = page;
    return SYNSEG_OK;
}

extern "C" SYNSEG_EXPORT int synseg_page_slot_acquire(synseg_ctx *ctx, int32_t *slot, void **host_pages, int64_t *row_stride, int64_t *page_stride)
{
    if (!ctx || !slot || !host_pages) { synseg_set_error("synseg_page_slot_acquire: NULL argument"); return SYNSEG_E_INVALID; }
    synseg_ctx::HostStream &h = ctx->hs;
    if (!h.ps_n) { synseg_set_error("synseg_page_slot_acquire: call synseg_page_slots_init first"); return SYNSEG_E_INVALID; }
    synseg_ctx::HostStream::PageSlot &q = h.ps[h.ps_next];
    if (q.state == 2) { synseg_set_error("synseg_page_slot_acquire: slot %d holds results nobody waited for (synseg_page_slot_wait)", h.ps_next); return SYNSEG_E_INVALID; }
    q.state = 1;
    *slot = h.ps_next; *host_pages = q.pages;
    if (row_stride) *row_stride = h.ps_row_stride;
    if (page_stride) *page_stride = h.ps_page_stride;
    h.ps_next = (h.ps_next + 1) % h.ps_n;
    return SYNSEG_OK;
}

extern "C" SYNSEG_EXPORT int synseg_page_slot_submit(synseg_ctx *ctx, int32_t slot, int32_t n_pages, const synseg_detect_params *prm,
                                                     const synseg_region_params *rp, void *stream)
{
    if (!ctx || !prm) { synseg_set_error("synseg_page_slot_submit: NULL argument"); return SYNSEG_E_INVALID; }
    SS_ENTER(ctx, stream);
    synseg_ctx::HostStream &h = ctx->hs;
    if (slot < 0 || slot >= h.ps_n || h.ps[slot].state != 1) { synseg_set_error("synseg_page_slot_submit: slot %d was not acquired", slot); return SYNSEG_E_INVALID; }
    if (n_pages < 1 || n_pages > h.ps_pages) { synseg_set_error("synseg_page_slot_submit: n_pages outside 1..%d", h.ps_pages); return SYNSEG_E_INVALID; }
    SS_TRY(validate_region_params(rp, "synseg_page_slot_submit"));
    const int ch = prm->channels == 1 ? 1 : 3;
    if (ch != h.ps_channels || prm->max_labels != h.ps_max_labels || rp->max_regions != h.ps_max_regions) {
        synseg_set_error("synseg_page_slot_submit: channels / max_labels / max_regions differ from synseg_page_slots_init"); return SYNSEG_E_INVALID;
    }
    if (prm->block_size < 3 || prm->block_size > 255 || !(prm->block_size & 1) || prm->k < 1 || prm->canny_lo < 0 || prm->canny_hi < prm->canny_lo) {
        synseg_set_error("synseg_page_slot_submit: bad parameters"); return SYNSEG_E_INVALID;
    }
    synseg_ctx::HostStream::PageSlot &q = h.ps[slot];
    const int s = slot;                  // host slot i is staged through ring slot i (n_slots <= RING)
    if (h.used[s]) SS_CUDA(cudaStreamWaitEvent(h.copy, h.done[s], 0));
    SS_TRY(stage_pages(ctx, s, q.pages, h.ps_width, h.ps_height, ch, h.ps_row_stride, h.ps_page_stride, n_pages));
    SS_CUDA(cudaEventRecord(h.copied[s], h.copy));
    const HostOut o{q.ints, q.stats, nullptr, q.regions, q.ints + h.ps_pages, q.ints + 2 * h.ps_pages};
    SS_TRY(run_ring_slot(ctx, s, h.ps_width, h.ps_height, ch, h.ps_row_stride, h.ps_page_stride, n_pages, prm, rp, o, 0, (cudaStream_t)stream));
    SS_CUDA(cudaEventRecord(q.finished, (cudaStream_t)stream));
    q.state = 2; q.n_pages = n_pages;
    return SYNSEG_OK;
}

extern "C" SYNSEG_EXPORT int synseg_page_slot_wait(synseg_ctx *ctx, int32_t slot, const int32_t **n_regions, const int32_t **flags,
                                                   const synseg_region **regions, const int32_t **n_labels, const int32_t **stats)
{
    if (!ctx) { synseg_set_error("synseg_page_slot_wait: ctx is NULL"); return SYNSEG_E_INVALID; }
    synseg_ctx::HostStream &h = ctx->hs;
    if (slot < 0 || slot >= h.ps_n || h.ps[slot].state != 2) { synseg_set_error("synseg_page_slot_wait: slot %d was not submitted", slot); return SYNSEG_E_INVALID; }
    synseg_ctx::HostStream::PageSlot &q = h.ps[slot];
    SS_CUDA(cudaEventSynchronize(q.finished));
    q.state = 0;
    if (n_labels) *n_labels = q.ints;
    if (n_regions) *n_regions = q.ints + h.ps_pages;
    if (flags) *flags = q.ints + 2 * h.ps_pages;
    if (regions) *regions = q.regions;
    if (stats) *stats = q.stats;
    return SYNSEG_OK;
}
