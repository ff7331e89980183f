// phash.cu -- 64-bit DCT perceptual hash per region + replicated Hamming dedup.
//
// Not in the reference: its only hash is md5(png bytes)[:8] used as a segment id
// (pdf_image_segmentation.py:3782) and id-dedup (:3886-3887).  This is the north-star's cross-page
// duplicate-figure key.  Defined in exact integer arithmetic so that the GPU,
// the C oracle (oracle/synseg_oracle.c:orc_phash) and every rank agree bit for bit:
//   grey -> 32x32 cell means q = (256*sum + cnt/2)/cnt -> F = C q C^T with C[u][x] =
//   lround(16384 cos(pi (2x+1) u / 64)), u < 8 -> bit k = 2 F[k] > (sorted[31] + sorted[32]).
// Two kernels: cell means with one CTA per (row group, region) -- every pixel is read once, thread per column --
// then one CTA per region for the 32-point integer DCTs, the median and the bits.
#include "internal.cuh"
#include "pixel.cuh"

namespace {

constexpr int PH_MAXW = 8192;

// Stage 1: cell means.  grid = (32 row groups, regions): CTA (i, r) reduces row group i of region r to its
// 32 cell values q[i][0..31] (thread per column, rows of the group unrolled by 4 for memory-level parallelism).
template <int SRC_KIND>
__device__ __forceinline__ void phash_cells_one(uint32_t *colsum, const Plane &src, int width, int height, const synseg_roi *rois, int32_t *qbuf, int ri)
{
    synseg_roi r;
    if (rois) r = rois[ri];
    else { r.image = ri; r.x = 0; r.y = 0; r.width = width; r.height = height; }
    const int tid = threadIdx.x, i = blockIdx.x;
    const uint8_t *base = src.p + r.image * src.bs;
    const int w = r.width, h = r.height;
    const int y0 = (int)((long long)i * h / 32);
    int y1 = (int)((long long)(i + 1) * h / 32);
    if (y1 <= y0) y1 = y0 + 1;
    constexpr int bpp = SRC_KIND == 0 ? 1 : 3;
    constexpr int MODE = SRC_KIND == 1 ? SYNSEG_GRAY_PIL : SYNSEG_GRAY_CV;
    const int64_t row_bytes = (int64_t)width * bpp;                  // bytes of an image row that may be read
    // Four pixels per thread and row.  Their 4*bpp bytes are fetched as aligned 32-bit words and funnel-shifted by the
    // byte phase of the row (the same for every thread of the row): a third of the load instructions of a byte loop.
    // The row loop of the word path is branch-free, so its loads are issued four rows ahead of their use.  Threads
    // whose words could cross the ends of the image row (any phase) take the byte path.
    for (int x = 4 * tid; x < w; x += 4 * 256) {
        uint32_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;
        const int npx = min(4, w - x);
        const int64_t off = (int64_t)(r.x + x) * bpp;                // byte offset of the first pixel inside the row
        const uint8_t *p0 = base + (int64_t)(r.y + y0) * src.rs + off;
        if (npx == 4 && off >= 3 && off + 4 * (bpp + 1) <= row_bytes) {
#pragma unroll 4
            for (int y = y0; y < y1; ++y, p0 += src.rs) {
                const int ph = (int)((uintptr_t)p0 & 3);
                const uint32_t *wp = (const uint32_t *)(p0 - ph);
                const int sh = 8 * ph;
                if (bpp == 3) {
                    const uint32_t w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2), w3 = __ldg(wp + 3);
                    const uint32_t b0 = __funnelshift_r(w0, w1, sh), b1 = __funnelshift_r(w1, w2, sh), b2 = __funnelshift_r(w2, w3, sh);
                    s0 += gray1<MODE>(b0); s1 += gray1<MODE>(__funnelshift_r(b0, b1, 24));
                    s2 += gray1<MODE>(__funnelshift_r(b1, b2, 16)); s3 += gray1<MODE>(b2 >> 8);
                } else {
                    const uint32_t g4 = __funnelshift_r(__ldg(wp), __ldg(wp + 1), sh);
                    s0 += g4 & 255u; s1 += (g4 >> 8) & 255u; s2 += (g4 >> 16) & 255u; s3 += g4 >> 24;
                }
            }
        } else {
            for (int y = y0; y < y1; ++y, p0 += src.rs) {
                uint32_t px[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    px[k] = 0;
                    if (k < npx) {
                        const uint8_t *q = p0 + k * bpp;
                        px[k] = bpp == 3 ? ((uint32_t)__ldg(q) | ((uint32_t)__ldg(q + 1) << 8) | ((uint32_t)__ldg(q + 2) << 16)) : (uint32_t)__ldg(q);
                    }
                }
                if (bpp == 3) { s0 += gray1<MODE>(px[0]); s1 += gray1<MODE>(px[1]); s2 += gray1<MODE>(px[2]); s3 += gray1<MODE>(px[3]); }
                else { s0 += px[0]; s1 += px[1]; s2 += px[2]; s3 += px[3]; }
            }
        }
        colsum[x] = s0;
        if (npx > 1) colsum[x + 1] = s1;
        if (npx > 2) colsum[x + 2] = s2;
        if (npx > 3) colsum[x + 3] = s3;
    }
    __syncthreads();
    if (tid < 32) {
        const int j = tid;
        const int x0 = (int)((long long)j * w / 32);
        int x1 = (int)((long long)(j + 1) * w / 32);
        if (x1 <= x0) x1 = x0 + 1;
        unsigned long long s = 0;
        for (int x = x0; x < x1; ++x) s += colsum[x];
        const unsigned long long c = (unsigned long long)(y1 - y0) * (unsigned long long)(x1 - x0);
        qbuf[((int64_t)ri * 32 + i) * 32 + j] = (int32_t)((256ull * s + c / 2) / c);
    }
}

// grid = (32 row groups, G): CTA (i, g) handles regions g, g + G, ... below n = min(*count, n_rois) (count: device-resident
// list length of the indirect form, or NULL).  A list whose length is only known on the device is launched with G <= 1024, not
// with one grid row per capacity slot: 8,000 slots x 32 empty CTAs cost 0.37 ms per call (profiles/r2_launches_summary.txt).
template <int SRC_KIND>
__global__ void __launch_bounds__(256) phash_cells_kernel(Plane src, int width, int height, const synseg_roi *rois,
                                                          int32_t *qbuf, const int32_t *count, int n_rois)
{
    __shared__ uint32_t colsum[PH_MAXW];
    const int n = count ? min(max(*count, 0), n_rois) : n_rois;
    for (int ri = blockIdx.y; ri < n; ri += gridDim.y) {
        phash_cells_one<SRC_KIND>(colsum, src, width, height, rois, qbuf, ri);
        __syncthreads();
    }
}

// Stage 2: one CTA per region: integer DCT of the 32x32 cell values, low 8x8 block, median bits.
__global__ void __launch_bounds__(256) phash_dct_kernel(const int32_t *qbuf, const int32_t *basis, unsigned long long *out,
                                                        const int32_t *count, int n_rois)
{
    __shared__ int32_t q[32][33];
    __shared__ long long T[8][33];
    __shared__ long long F[64];
    __shared__ int32_t cb[8 * 32];
    __shared__ long long med2;
    __shared__ unsigned int hbits[2];
    const int tid = threadIdx.x;
    cb[tid] = basis[tid];
    const int n = count ? min(max(*count, 0), n_rois) : n_rois;
    for (int ri = blockIdx.x; ri < n; ri += gridDim.x) {
    if (tid == 0) med2 = 0;
    for (int k = tid; k < 1024; k += 256) q[k >> 5][k & 31] = qbuf[(int64_t)ri * 1024 + k];
    __syncthreads();
    // T[v][y] = sum_x C[v][x] q[y][x]
    {
        const int v = tid >> 5, y = tid & 31;
        long long acc = 0;
#pragma unroll
        for (int x = 0; x < 32; ++x) acc += (long long)cb[v * 32 + x] * q[y][x];
        T[v][y] = acc;
    }
    __syncthreads();
    if (tid < 64) {
        const int u = tid >> 3, v = tid & 7;
        long long acc = 0;
#pragma unroll
        for (int y = 0; y < 32; ++y) acc += (long long)cb[u * 32 + y] * T[v][y];
        F[tid] = acc;
    }
    __syncthreads();
    if (tid < 64) {
        const long long f = F[tid];
        int rank = 0;
        for (int m = 0; m < 64; ++m) { const long long g = F[m]; rank += (g < f) || (g == f && m < tid); }
        if (rank == 31 || rank == 32) atomicAdd((unsigned long long *)&med2, (unsigned long long)f);
    }
    __syncthreads();
    if (tid < 64) {
        const bool bit = 2 * F[tid] > med2;
        const unsigned m = __ballot_sync(0xffffffffu, bit);
        if ((tid & 31) == 0) hbits[tid >> 5] = m;
    }
    __syncthreads();
    if (tid == 0) {
        // bit k of the hash (MSB first) is coefficient k; ballot bit l of word wv is coefficient 32*wv + l
        const unsigned long long hi = __brev(hbits[0]), lo = __brev(hbits[1]);
        out[ri] = (hi << 32) | lo;
    }
    __syncthreads();
    }
}

// keep[i] = 0 iff some j with key[j] < key[i] has popcount(hash[i] ^ hash[j]) <= max_hamming.  All pairs, split over
// blockIdx.y slices of j so that a few thousand regions still fill the GPU; keep[] is preset to 1 by the launcher.
// A slice of the (hash, key) pairs is staged in shared memory and scanned by every thread of the CTA.
constexpr int DEDUP_TILE = 1024;
__global__ void __launch_bounds__(256) phash_dedup_kernel(const unsigned long long *hashes, const unsigned long long *keys, int n,
                                                          int max_hamming, uint8_t *keep)
{
    __shared__ unsigned long long sh[DEDUP_TILE], sk[DEDUP_TILE];
    const int i = blockIdx.x * 256 + threadIdx.x;
    const int j0 = blockIdx.y * DEDUP_TILE, j1 = min(j0 + DEDUP_TILE, n);
    for (int j = j0 + threadIdx.x; j < j1; j += 256) { sh[j - j0] = hashes[j]; sk[j - j0] = keys[j]; }
    __syncthreads();
    if (i >= n) return;
    const unsigned long long h = hashes[i], k = keys[i];
    bool dup = false;
    for (int j = 0; j < j1 - j0; ++j) dup |= (sk[j] < k) && (__popcll(sh[j] ^ h) <= max_hamming);
    if (dup) keep[i] = 0;
}

__global__ void __launch_bounds__(256) select_rois_kernel(const int32_t *n_labels, const int32_t *stats, int batch, int max_labels,
                                                          long long page_base, int min_area, int max_area, int min_w, int min_h,
                                                          synseg_roi *rois, unsigned long long *keys, int32_t *count, int capacity)
{
    const int img = blockIdx.y;
    const int k = blockIdx.x * 256 + threadIdx.x;
    int n = n_labels[img];
    if (n < 0) n = max_labels;           // over capacity: only the first max_labels rows exist
    if (k < 1 || k >= n) return;
    const int32_t *s = stats + ((int64_t)img * max_labels + k) * 5;
    const int w = s[2], h = s[3];
    const long long box = (long long)w * h;
    if (box < min_area || box > max_area || w < min_w || h < min_h) return;
    const int slot = atomicAdd(count, 1);
    if (slot >= capacity) return;        // *count keeps counting: count > capacity tells the caller that entries were dropped
    synseg_roi r; r.image = img; r.x = s[0]; r.y = s[1]; r.width = w; r.height = h;
    rois[slot] = r;
    keys[slot] = ((unsigned long long)(page_base + img) << 16) | (unsigned long long)k;
}

}  // namespace

static int run_phash(synseg_ctx *ctx, const synseg_img *src, int src_kind, const synseg_roi *rois, int n, const int32_t *count,
                     uint64_t *out, cudaStream_t st)
{
    SS_TRY(arena_ensure(ctx, (size_t)n * 1024 * sizeof(int32_t) + 256));
    arena_begin(ctx);
    void *p;
    SS_TRY(arena_alloc(ctx, (size_t)n * 1024 * sizeof(int32_t), &p, st));
    int32_t *qbuf = (int32_t *)p;
    const int gy = count ? (n < 1024 ? n : 1024) : n;          // device-resident length: a fixed grid that strides over the list
    if (src_kind == 0) phash_cells_kernel<0><<<dim3(32, gy), 256, 0, st>>>(plane_of(src), src->width, src->height, rois, qbuf, count, n);
    else if (src_kind == 1) phash_cells_kernel<1><<<dim3(32, gy), 256, 0, st>>>(plane_of(src), src->width, src->height, rois, qbuf, count, n);
    else phash_cells_kernel<2><<<dim3(32, gy), 256, 0, st>>>(plane_of(src), src->width, src->height, rois, qbuf, count, n);
    SS_LAUNCH_CHECK(ctx, "phash_cells", st);
    phash_dct_kernel<<<count ? (n < 4096 ? n : 4096) : n, 256, 0, st>>>(qbuf, ctx->phash_basis, (unsigned long long *)out, count, n);
    SS_LAUNCH_CHECK(ctx, "phash_dct", st);
    return SYNSEG_OK;
}

extern "C" SYNSEG_EXPORT int synseg_select_rois(synseg_ctx *ctx, const int32_t *n_labels, const int32_t *stats, int32_t batch,
                                                int32_t max_labels, int64_t page_base, int32_t min_area, int32_t max_area, int32_t min_w,
                                                int32_t min_h, synseg_roi *rois, uint64_t *keys, int32_t *count, int32_t capacity,
                                                void *stream)
{
    if (!ctx) { synseg_set_error("synseg_select_rois: ctx is NULL"); return SYNSEG_E_INVALID; }
    SS_ENTER(ctx, stream);
    if (!n_labels || !stats || !rois || !keys || !count || batch <= 0 || max_labels < 1 || capacity < 1 || max_labels > 65535) {
        synseg_set_error("synseg_select_rois: bad arguments"); return SYNSEG_E_INVALID;
    }
    select_rois_kernel<<<dim3(cdiv(max_labels, 256), batch), 256, 0, (cudaStream_t)stream>>>(
        n_labels, stats, batch, max_labels, (long long)page_base, min_area, max_area, min_w, min_h, rois, (unsigned long long *)keys, count,
        capacity);
    SS_LAUNCH_CHECK(ctx, "select_rois", (cudaStream_t)stream);
    return SYNSEG_OK;
}

extern "C" SYNSEG_EXPORT int synseg_phash_indirect(synseg_ctx *ctx, const synseg_img *src, int src_kind, const synseg_roi *rois,
                                                   const int32_t *count, int32_t capacity, uint64_t *out, void *stream)
{
    if (!ctx) { synseg_set_error("synseg_phash_indirect: ctx is NULL"); return SYNSEG_E_INVALID; }
    SS_ENTER(ctx, stream);
    if (src_kind < 0 || src_kind > 2) { synseg_set_error("synseg_phash_indirect: bad src_kind %d", src_kind); return SYNSEG_E_INVALID; }
    SS_TRY(validate_img(src, "src", src_kind ? 3 : 1));
    if (!rois || !count || !out || capacity < 1) { synseg_set_error("synseg_phash_indirect: bad arguments"); return SYNSEG_E_INVALID; }
    if (src->width > PH_MAXW) { synseg_set_error("synseg_phash_indirect: width > %d", PH_MAXW); return SYNSEG_E_INVALID; }
    if (capacity > 65535) { synseg_set_error("synseg_phash_indirect: capacity > 65535"); return SYNSEG_E_INVALID; }
    return run_phash(ctx, src, src_kind, rois, capacity, count, out, (cudaStream_t)stream);
}

extern "C" SYNSEG_EXPORT int synseg_phash(synseg_ctx *ctx, const synseg_img *src, int src_kind, const synseg_roi *rois, int32_t n_rois, uint64_t *out,
                            void *stream)
{
    if (!ctx) { synseg_set_error("synseg_phash: ctx is NULL"); return SYNSEG_E_INVALID; }
    SS_ENTER(ctx, stream);
    if (src_kind < 0 || src_kind > 2) { synseg_set_error("synseg_phash: bad src_kind %d", src_kind); return SYNSEG_E_INVALID; }
    SS_TRY(validate_img(src, "src", src_kind ? 3 : 1));
    if (!out) { synseg_set_error("synseg_phash: out is NULL"); return SYNSEG_E_INVALID; }
    if (!rois) n_rois = src->batch;
    if (n_rois <= 0) return SYNSEG_OK;
    if (src->width > PH_MAXW) { synseg_set_error("synseg_phash: width > %d", PH_MAXW); return SYNSEG_E_INVALID; }
    for (int32_t o = 0; o < n_rois; o += 32768) {       // grid.y limit
        const int32_t n = n_rois - o < 32768 ? n_rois - o : 32768;
        synseg_img view = *src;
        if (!rois) { view.data = (uint8_t *)src->data + (int64_t)o * src->batch_stride; view.batch = n; }
        SS_TRY(run_phash(ctx, &view, src_kind, rois ? rois + o : nullptr, n, nullptr, out + o, (cudaStream_t)stream));
    }
    return SYNSEG_OK;
}

extern "C" SYNSEG_EXPORT int synseg_phash_dedup(synseg_ctx *ctx, const uint64_t *hashes, const uint64_t *keys, int32_t n, int32_t max_hamming,
                                  uint8_t *keep, void *stream)
{
    if (!ctx) { synseg_set_error("synseg_phash_dedup: ctx is NULL"); return SYNSEG_E_INVALID; }
    SS_ENTER(ctx, stream);
    if (n <= 0) return SYNSEG_OK;
    if (!hashes || !keys || !keep) { synseg_set_error("synseg_phash_dedup: NULL buffer"); return SYNSEG_E_INVALID; }
    SS_CUDA(cudaMemsetAsync(keep, 1, (size_t)n, (cudaStream_t)stream));
    phash_dedup_kernel<<<dim3(cdiv(n, 256), cdiv(n, DEDUP_TILE)), 256, 0, (cudaStream_t)stream>>>((const unsigned long long *)hashes,
                                                                                                  (const unsigned long long *)keys, n, max_hamming, keep);
    SS_LAUNCH_CHECK(ctx, "phash_dedup", (cudaStream_t)stream);
    return SYNSEG_OK;
}
