// regions.cu -- component tables -> candidate regions on the device, with the exact grey moments of every region crop.
//
// The host stage of the detector (synapta_image_segmentation_b200/detector.py: candidate_regions, geometry.py) restates
// the reference's box rules in Python; this file is the same arithmetic on the device so that a page batch leaves the GPU
// as ready-to-score regions in ONE stream, without a host round trip between the component tables and the crop moments:
//
//   per component (x, y, w, h, area in px; row 0 = background)          rect = px * (72 / dpi) in PDF points
//     big  : 5000 < area_pt < 0.8 page  and  w_pt, h_pt > min_extent   -> region "raster_cc"          (_detect_by_drawings' area
//     small: not big, area_pt < 0.8 page                               -> clustering input             filter, pdf_image_segmentation.py:3549)
//   greedy clustering of the small rects in component order             (_cluster_drawings :3559-3594, _drawing_distance :3596-3618):
//     an unused seed absorbs every unused rect whose gap to the SEED is < 100 pt; >= 3 members make a cluster
//   cluster -> bbox of the members, padded 10 pt, clamped to the page, 5000 < area < 0.8 page   (_detect_by_drawings :3531-3555)
//   merge: a cluster region is dropped when more than half of it lies inside a region kept before it  (_overlaps_with_existing :3620-3636,
//          _detect_visual_regions :3122-3144)
//   crop  : BoundingBox.to_pixels (round half to even of pt * dpi / 72, :3649) clamped to the page
//   moments: sum and sum of squares of the crop's grey pixels (PIL grey for RGB pages) -> np.var (:2988-2989)
//
// Exactness: every f64 operation is an explicit IEEE multiply / add / subtract in the host code's order (no contraction into
// FMA), so boxes and areas are bit-identical with the Python floats.  The one operation that cannot be replayed bit for bit
// is `(dx**2 + dy**2)**0.5 < 100` (libm pow on the host): the kernel decides on dx*dx + dy*dy < 10000 and raises
// SYNSEG_REGION_FLAG_AMBIGUOUS for a page holding a pair within 1e-9 (relative) of the threshold; the host then
// recomputes that page from the component table (which is returned as well).  Pages with more components than
// max_labels or more regions than max_regions are flagged the same way.
#include "internal.cuh"
#include "pixel.cuh"

namespace {

constexpr int RG_THREADS = 256;

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }

struct RgParams {
    double s;            // 72 / dpi          (points per pixel)
    double inv_s;        // dpi / 72          (pixels per point)
    double pw, ph;       // page size in points
    double min_extent;   // points
    int width, height;   // page size in pixels
    int max_labels, max_regions;
};

struct Box { double x0, y0, x1, y1; };

__device__ __forceinline__ double box_area(const Box &b) { return dmul(dsub(b.x1, b.x0), dsub(b.y1, b.y0)); }

// _overlaps_with_existing: intersection > 0.5 * area(candidate)
__device__ __forceinline__ bool overlaps_half(const Box &c, const Box &e)
{
    const double xo = fmax(0.0, dsub(fmin(c.x1, e.x1), fmax(c.x0, e.x0)));
    const double yo = fmax(0.0, dsub(fmin(c.y1, e.y1), fmax(c.y0, e.y0)));
    return dmul(xo, yo) > dmul(box_area(c), 0.5);
}

__device__ __forceinline__ void emit_region(synseg_region *out, const RgParams &p, const Box &b, int kind, int count)
{
    synseg_region r;
    r.x0 = b.x0; r.y0 = b.y0; r.x1 = b.x1; r.y1 = b.y1;
    // BoundingBox.to_pixels: int(round(v * dpi / 72)) (Python rounds half to even = rint), then the detector's clamp
    const int X0 = (int)rint(dmul(b.x0, p.inv_s)), Y0 = (int)rint(dmul(b.y0, p.inv_s));
    const int X1 = (int)rint(dmul(b.x1, p.inv_s)), Y1 = (int)rint(dmul(b.y1, p.inv_s));
    int w = max(1, X1 - X0), h = max(1, Y1 - Y0);
    const int x = min(max(X0, 0), p.width - 1), y = min(max(Y0, 0), p.height - 1);
    w = max(1, min(w, p.width - x)); h = max(1, min(h, p.height - y));
    r.px = x; r.py = y; r.pw = w; r.ph = h;
    r.kind = kind; r.count = count;
    r.sum = 0; r.sum_sq = 0;
    *out = r;
}

// One CTA per page.  scratch: int4 small[max_labels] per page (pixel coordinates x0, y0, x1, y1 of the small rects, in
// component order); dynamic shared memory: the `used` bit set of the clustering.
__global__ void __launch_bounds__(RG_THREADS) regions_kernel(const int32_t *n_labels, const int32_t *stats, RgParams p, int4 *small_all,
                                                             int4 *clusters_all, synseg_region *regions, int32_t *n_regions, int32_t *flags)
{
    extern __shared__ uint32_t used[];                 // cdiv(max_labels, 32) words
    __shared__ int warp_cnt[2][RG_THREADS / 32];
    __shared__ int base_cnt[2];                        // primaries / smalls emitted so far
    __shared__ int cl_cnt, cl_minx, cl_miny, cl_maxx, cl_maxy, n_clusters, sh_flags;
    const int img = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int32_t *st = stats + (int64_t)img * p.max_labels * 5;
    int4 *small = small_all + (int64_t)img * p.max_labels;
    int4 *clusters = clusters_all + (int64_t)img * p.max_labels;     // (minx, miny) / (maxx, maxy) / count packed in two int4 halves
    synseg_region *out = regions + (int64_t)img * p.max_regions;
    int n = n_labels[img];
    if (tid == 0) { base_cnt[0] = 0; base_cnt[1] = 0; n_clusters = 0; sh_flags = 0; }
    for (int i = tid; i < (p.max_labels + 31) / 32; i += RG_THREADS) used[i] = 0u;
    __syncthreads();
    if (n < 0) {                                       // more components than max_labels: the host retries with a larger table
        if (tid == 0) { n_regions[img] = 0; flags[img] = SYNSEG_REGION_FLAG_LABELS; }
        return;
    }
    const double page_cap = dmul(dmul(p.pw, p.ph), 0.8);

    // ---- 1. classify the components, ordered compaction of the big ones (-> regions) and the small ones ----------------
    for (int k0 = 1; k0 < n; k0 += RG_THREADS) {
        const int k = k0 + tid;
        bool big = false, sm = false;
        int x = 0, y = 0, w = 0, h = 0, area = 0;
        Box b{0, 0, 0, 0};
        if (k < n) {
            x = st[5 * k]; y = st[5 * k + 1]; w = st[5 * k + 2]; h = st[5 * k + 3]; area = st[5 * k + 4];
            b.x0 = dmul((double)x, p.s); b.y0 = dmul((double)y, p.s); b.x1 = dmul((double)(x + w), p.s); b.y1 = dmul((double)(y + h), p.s);
            const double a = box_area(b);
            big = a > 5000.0 && a < page_cap && dsub(b.x1, b.x0) > p.min_extent && dsub(b.y1, b.y0) > p.min_extent;
            sm = !big && a < page_cap;
        }
        const unsigned mb = __ballot_sync(0xffffffffu, big), ms = __ballot_sync(0xffffffffu, sm);
        if (lane == 0) { warp_cnt[0][warp] = __popc(mb); warp_cnt[1][warp] = __popc(ms); }
        __syncthreads();
        int off_b = base_cnt[0], off_s = base_cnt[1];
        for (int q = 0; q < warp; ++q) { off_b += warp_cnt[0][q]; off_s += warp_cnt[1][q]; }
        off_b += __popc(mb & ((1u << lane) - 1u)); off_s += __popc(ms & ((1u << lane) - 1u));
        if (big) {
            if (off_b < p.max_regions) emit_region(out + off_b, p, b, SYNSEG_REGION_CC, area);
            else atomicOr(&sh_flags, SYNSEG_REGION_FLAG_CAPACITY);
        }
        SS_DEVICE_ASSERT(!sm || off_s < p.max_labels);
        if (sm) small[off_s] = make_int4(x, y, x + w, y + h);
        __syncthreads();
        if (tid == 0) {
            int tb = 0, ts = 0;
            for (int q = 0; q < RG_THREADS / 32; ++q) { tb += warp_cnt[0][q]; ts += warp_cnt[1][q]; }
            base_cnt[0] += tb; base_cnt[1] += ts;
        }
        __syncthreads();
    }
    const int n_primary = min(base_cnt[0], p.max_regions), ns = base_cnt[1];

    // ---- 2. greedy clustering of the small rects (seed order = component order) -----------------------------------------
    for (int i = 0; i < ns; ++i) {
        if ((used[i >> 5] >> (i & 31)) & 1u) continue;            // block-uniform: `used` is stable between the barriers
        const int4 si = small[i];
        if (tid == 0) { cl_cnt = 1; cl_minx = si.x; cl_miny = si.y; cl_maxx = si.z; cl_maxy = si.w; used[i >> 5] |= 1u << (i & 31); }
        __syncthreads();
        const double a0 = dmul((double)si.x, p.s), b0 = dmul((double)si.y, p.s), a1 = dmul((double)si.z, p.s), b1 = dmul((double)si.w, p.s);
        for (int j = tid; j < ns; j += RG_THREADS) {
            if (j == i || ((used[j >> 5] >> (j & 31)) & 1u)) continue;
            const int4 sj = small[j];
            const double c0 = dmul((double)sj.x, p.s), d0 = dmul((double)sj.y, p.s), c1 = dmul((double)sj.z, p.s), d1 = dmul((double)sj.w, p.s);
            bool near;
            if (a0 <= c1 && a1 >= c0 && b0 <= d1 && b1 >= d0) near = true;                       // touching / overlapping: distance 0
            else {
                const double dx = fmax(0.0, fmax(dsub(a0, c1), dsub(c0, a1))), dy = fmax(0.0, fmax(dsub(b0, d1), dsub(d0, b1)));
                const double v = dadd(dmul(dx, dx), dmul(dy, dy));
                near = v < 10000.0;
                if (fabs(v - 10000.0) <= 1e-5) atomicOr(&sh_flags, SYNSEG_REGION_FLAG_AMBIGUOUS);
            }
            if (near) {
                atomicOr(&used[j >> 5], 1u << (j & 31));
                atomicAdd(&cl_cnt, 1);
                atomicMin(&cl_minx, sj.x); atomicMin(&cl_miny, sj.y); atomicMax(&cl_maxx, sj.z); atomicMax(&cl_maxy, sj.w);
            }
        }
        __syncthreads();
        if (tid == 0 && cl_cnt >= 3) {
            SS_DEVICE_ASSERT(2 * n_clusters + 1 < p.max_labels);
            clusters[2 * n_clusters] = make_int4(cl_minx, cl_miny, cl_maxx, cl_maxy);
            clusters[2 * n_clusters + 1] = make_int4(cl_cnt, 0, 0, 0);
            ++n_clusters;
        }
        __syncthreads();
    }

    // ---- 3. clusters -> padded regions, merged behind the primaries (sequential: every test sees the regions kept before) ----
    if (tid == 0) {
        int n_out = n_primary;
        for (int c = 0; c < n_clusters; ++c) {
            const int4 m = clusters[2 * c];
            const int cnt = clusters[2 * c + 1].x;
            Box b;
            b.x0 = fmax(0.0, dsub(dmul((double)m.x, p.s), 10.0)); b.y0 = fmax(0.0, dsub(dmul((double)m.y, p.s), 10.0));
            b.x1 = fmin(p.pw, dadd(dmul((double)m.z, p.s), 10.0)); b.y1 = fmin(p.ph, dadd(dmul((double)m.w, p.s), 10.0));
            const double a = box_area(b);
            if (!(a > 5000.0 && a < page_cap)) continue;
            bool dup = false;
            for (int e = 0; e < n_out && !dup; ++e) {
                const Box eb{out[e].x0, out[e].y0, out[e].x1, out[e].y1};
                dup = overlaps_half(b, eb);
            }
            if (dup) continue;
            if (n_out < p.max_regions) emit_region(out + n_out++, p, b, SYNSEG_REGION_CLUSTER, cnt);
            else sh_flags |= SYNSEG_REGION_FLAG_CAPACITY;
        }
        n_regions[img] = n_out;
        flags[img] = sh_flags;
    }
}

__device__ __forceinline__ unsigned long long warp_sum64(unsigned long long v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
    return v;
}

// Grey moments of every region crop: grid (row slabs, max_regions, pages); a CTA takes every gridDim.x-th group of 8 rows.
// channels 3: PIL grey of the RGB pixels (the crop the reference scores is `image.convert('L')`, :2988); channels 1: the grey page itself.
template <int CH>
__global__ void __launch_bounds__(256) region_moments_kernel(Plane src, const int32_t *n_regions, synseg_region *regions, int max_regions)
{
    const int img = blockIdx.z, j = blockIdx.y;
    if (j >= n_regions[img]) return;
    synseg_region *r = regions + (int64_t)img * max_regions + j;
    const int X = r->px, Y = r->py, Wc = r->pw, Hc = r->ph;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint8_t *base = src.p + img * src.bs;
    unsigned long long s = 0, ss = 0;
    for (int y = blockIdx.x * 8 + warp; y < Hc; y += gridDim.x * 8) {
        const uint8_t *row = base + (int64_t)(Y + y) * src.rs + (int64_t)CH * X;
        unsigned int rs = 0, rss = 0;                  // per lane and row: <= 2048 px * 65025 fits 32 bits for rows up to 65535 px
        if (CH == 1) {
            // head to a 4-byte boundary, aligned words, tail
            const int head = min(Wc, (int)((4 - ((uintptr_t)row & 3)) & 3));
            if (lane < head) { const uint32_t v = __ldg(row + lane); rs += v; rss += v * v; }
            const uint32_t *w4 = (const uint32_t *)(row + head);
            const int nw = (Wc - head) >> 2;
            for (int q = lane; q < nw; q += 32) {
                const uint32_t v = __ldg(w4 + q);
                rs += __dp4a(v, 0x01010101u, 0u); rss += __dp4a(v, v, 0u);
            }
            const int t0 = head + 4 * nw;
            if (t0 + lane < Wc) { const uint32_t v = __ldg(row + t0 + lane); rs += v; rss += v * v; }
        } else {
            for (int x = lane; x < Wc; x += 32) {
                const uint8_t *q = row + 3 * x;
                const uint32_t rgbx = (uint32_t)__ldg(q) | ((uint32_t)__ldg(q + 1) << 8) | ((uint32_t)__ldg(q + 2) << 16);
                const uint32_t v = gray1<SYNSEG_GRAY_PIL>(rgbx);
                rs += v; rss += v * v;
            }
        }
        s += rs; ss += rss;
    }
    s = warp_sum64(s); ss = warp_sum64(ss);
    __shared__ unsigned long long sh[8][2];
    if (lane == 0) { sh[warp][0] = s; sh[warp][1] = ss; }
    __syncthreads();
    if (threadIdx.x < 2) {
        unsigned long long t = 0;
        for (int w = 0; w < 8; ++w) t += sh[w][threadIdx.x];
        if (t) atomicAdd((unsigned long long *)(threadIdx.x ? &r->sum_sq : &r->sum), t);
    }
}

}  // namespace

size_t regions_scratch_bytes(int batch, int max_labels) { return 2 * (sizeof(int4) * (size_t)batch * max_labels + 256) + 256; }

int validate_region_params(const synseg_region_params *rp, const char *who)
{
    if (!rp || !(rp->dpi > 0.0) || !(rp->page_width_pt > 0.0) || !(rp->page_height_pt > 0.0) || rp->max_regions < 1 || rp->max_regions > 1024) {
        synseg_set_error("%s: bad region parameters (dpi, page size in points > 0; 1 <= max_regions <= 1024)", who);
        return SYNSEG_E_INVALID;
    }
    return SYNSEG_OK;
}

// The arena must hold regions_scratch_bytes(batch, max_labels) from its current top.
int run_regions(synseg_ctx *ctx, const int32_t *n_labels, const int32_t *stats, int32_t max_labels, const synseg_img *pages, int channels,
                const synseg_region_params *rp, synseg_region *regions, int32_t *n_regions, int32_t *flags, cudaStream_t st)
{
    const int B = pages->batch;
    RgParams p;
    p.s = 72.0 / rp->dpi; p.inv_s = rp->dpi / 72.0;
    p.pw = rp->page_width_pt; p.ph = rp->page_height_pt; p.min_extent = rp->min_extent_pt;
    p.width = pages->width; p.height = pages->height; p.max_labels = max_labels; p.max_regions = rp->max_regions;
    void *q;
    SS_TRY(arena_alloc(ctx, sizeof(int4) * (size_t)B * max_labels, &q, st)); int4 *small = (int4 *)q;
    SS_TRY(arena_alloc(ctx, sizeof(int4) * (size_t)B * max_labels, &q, st)); int4 *clusters = (int4 *)q;
    const size_t smem = sizeof(uint32_t) * (size_t)((max_labels + 31) / 32);
    regions_kernel<<<B, RG_THREADS, smem, st>>>(n_labels, stats, p, small, clusters, regions, n_regions, flags);
    SS_LAUNCH_CHECK(ctx, "regions", st);
    const dim3 grid(8, rp->max_regions, B);
    if (channels == 1) region_moments_kernel<1><<<grid, 256, 0, st>>>(plane_of(pages), n_regions, regions, rp->max_regions);
    else region_moments_kernel<3><<<grid, 256, 0, st>>>(plane_of(pages), n_regions, regions, rp->max_regions);
    SS_LAUNCH_CHECK(ctx, "region_moments", st);
    return SYNSEG_OK;
}

extern "C" SYNSEG_EXPORT int synseg_regions_from_stats(synseg_ctx *ctx, const int32_t *n_labels, const int32_t *stats, int32_t max_labels,
                                                       const synseg_img *pages, int channels, const synseg_region_params *rp,
                                                       synseg_region *regions, int32_t *n_regions, int32_t *flags, void *stream)
{
    if (!ctx) { synseg_set_error("synseg_regions_from_stats: ctx is NULL"); return SYNSEG_E_INVALID; }
    SS_ENTER(ctx, stream);
    if (channels != 1 && channels != 3) { synseg_set_error("synseg_regions_from_stats: channels must be 1 or 3"); return SYNSEG_E_INVALID; }
    SS_TRY(validate_img(pages, "pages", channels));
    SS_TRY(validate_region_params(rp, "synseg_regions_from_stats"));
    if (!n_labels || !stats || !regions || !n_regions || !flags || max_labels < 1 || max_labels > 262144) {
        synseg_set_error("synseg_regions_from_stats: bad arguments"); return SYNSEG_E_INVALID;
    }
    if (pages->batch > 65535) { synseg_set_error("synseg_regions_from_stats: batch > 65535"); return SYNSEG_E_INVALID; }
        SS_TRY(arena_ensure(ctx, regions_scratch_bytes(pages->batch, max_labels)));
    arena_begin(ctx);
    return run_regions(ctx, n_labels, stats, max_labels, pages, channels, rp, regions, n_regions, flags, (cudaStream_t)stream);
}
