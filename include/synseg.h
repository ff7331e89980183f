/*
 * synseg.h -- C ABI of libsynseg.so, the B200 (sm_100a) raster region-detection hot path.
 *
 * The reference (ashr2k/synapta-image-segmentation) has no FFI of its own: it is a single Python
 * file whose pixel arithmetic is delegated to the cv2 / PIL / numpy wheels.  Every entry point here
 * therefore replaces one *call site* of those wheels in /root/reference/pdf_image_segmentation.py
 * ("S:" below) or its old-algorithm twin ("O:"), or one of the north-star raster primitives that
 * have no reference call site (oracle = the cv2 4.13.0 primitive).
 * INTEGRATION.md shows the ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *  - plain C types only; no CUDA/torch types.  `stream` is a cudaStream_t passed as void*.
 *  - every function returns 0 on success, <0 on error (SYNSEG_E_*); synseg_last_error() gives text.
 *    "Nothing found" (no components, empty mask) is a valid result, never an error.
 *  - all image/result pointers are DEVICE pointers owned by the caller; calls are asynchronous on
 *    `stream`.  The library owns only the opaque context (scratch arena, grown on demand or
 *    pre-sized with synseg_reserve); one context per (process, GPU), not thread-safe.  The scratch arena
 *    is shared by all calls on a context: a call issued to a different stream than the previous one is made to wait
 *    for it (an event per call), so calls on one context never overlap each other; use one context per concurrent
 *    pipeline.  No call changes the caller's current CUDA device.
 *  - images are row-major, 8-bit unless stated, described by synseg_img: `batch` images of
 *    height x width, `row_stride` / `batch_stride` in BYTES.  RGB images are interleaved HWC
 *    (3 bytes per pixel; width counts pixels).  Any stride/alignment is accepted; rows whose base
 *    and stride are 16-byte aligned take the vectorised path.
 */
#ifndef SYNSEG_H
#define SYNSEG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SYNSEG_VERSION 200

#define SYNSEG_OK 0
#define SYNSEG_E_INVALID (-1)   /* bad argument */
#define SYNSEG_E_CUDA (-2)      /* CUDA runtime error (text in synseg_last_error) */
#define SYNSEG_E_NOMEM (-3)     /* scratch arena could not be grown */
#define SYNSEG_E_CAPACITY (-4)  /* caller-provided result capacity too small */

typedef struct synseg_ctx synseg_ctx;

typedef struct synseg_img {
    void *data;           /* device pointer to pixel (0,0) of image 0 */
    int32_t width;        /* pixels */
    int32_t height;       /* rows */
    int64_t row_stride;   /* bytes between rows */
    int32_t batch;        /* number of images (>= 1) */
    int32_t _pad;
    int64_t batch_stride; /* bytes between images */
} synseg_img;

/* Rectangular region of one image of a batch (crop given by reference, no copy). */
typedef struct synseg_roi {
    int32_t image;        /* index into the batch */
    int32_t x, y;         /* top-left pixel */
    int32_t width, height;
} synseg_roi;

/* ---- context -------------------------------------------------------------------------------- */
int synseg_version(void);
const char *synseg_last_error(void);
int synseg_create(int device, synseg_ctx **out);
int synseg_destroy(synseg_ctx *ctx);
/* Pre-size the scratch arena so that no call below allocates (bytes as returned by
 * synseg_scratch_bytes for the largest batch you will submit). */
int synseg_reserve(synseg_ctx *ctx, size_t bytes);
size_t synseg_scratch_bytes(int32_t width, int32_t height, int32_t batch);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
int64_t synseg_launch_count(const synseg_ctx *ctx);

/* Memory-safety build (compile with -DSYNSEG_GUARD, tools/guard_run.sh): every scratch allocation sits between canary zones that are
 * compared when the public call returns, and the kernels' index assertions are real device asserts.  Returns the number of
 * damaged zones seen so far; *zones_checked = zones compared; *guard_build = 1 in such a build (a normal build returns 0 / 0 / 0). */
int64_t synseg_guard_violations(const synseg_ctx *ctx, int64_t *zones_checked, int32_t *guard_build);

/* Per-kernel device timing for bench.py: begin records a start event on `stream`; every kernel the
 * library launches afterwards is followed by an event; end returns (kernel name, milliseconds) pairs in
 * launch order (value = number of launches, <0 on error). */
int synseg_profile_begin(synseg_ctx *ctx, void *stream);
int synseg_profile_end(synseg_ctx *ctx, const char **names, float *ms, int cap);

/* ---- colour -------------------------------------------------------------------------------- */
/* RGB -> grey.  mode 0: cv2.cvtColor(COLOR_RGB2GRAY) 15-bit fixed point (S:1348);
 *               mode 1: PIL Image.convert('L') 16-bit fixed point (S:1323,1549,1599,1699,1758,1804,2988,3072). */
#define SYNSEG_GRAY_CV 0
#define SYNSEG_GRAY_PIL 1
int synseg_rgb2gray(synseg_ctx *ctx, const synseg_img *rgb, const synseg_img *gray, int mode, void *stream);

/* ---- threshold / edges --------------------------------------------------------------------- */
/* cv2.adaptiveThreshold(g, 255, ADAPTIVE_THRESH_MEAN_C, invert ? THRESH_BINARY_INV : THRESH_BINARY,
 * block_size, C): box mean over BORDER_REPLICATE.  block_size odd, 3..255.   */
int synseg_adaptive_mean(synseg_ctx *ctx, const synseg_img *gray, const synseg_img *out, int block_size, int C,
                         int invert, void *stream);
/* cv2.Canny(g, lo, hi) (aperture 3, L1 gradient): S:1324,1366,1550,1600,1700,1759.  Output {0,255}. */
int synseg_canny(synseg_ctx *ctx, const synseg_img *gray, const synseg_img *edges, int lo, int hi, void *stream);

/* ---- morphology ---------------------------------------------------------------------------- */
#define SYNSEG_MORPH_ERODE 0
#define SYNSEG_MORPH_DILATE 1
#define SYNSEG_MORPH_OPEN 2   /* same values as cv2.MORPH_* */
#define SYNSEG_MORPH_CLOSE 3
#define SYNSEG_MORPH_BINARY 1 /* flag: caller guarantees pixels are 0 or 255 -> bit-packed fast path */
/* cv2.erode / dilate / morphologyEx(op, getStructuringElement(MORPH_RECT,(kw,kh)), anchor=(ax,ay),
 * iterations) with the default constant border.  ax/ay = -1 means the centre (kw/2, kh/2).
 * Call sites: S:1370,1375 (1 x max(20,H/20), max(20,W/20) x 1, OPEN, it=2); S:1556,1557 (25x1, 1x25). */
int synseg_morph(synseg_ctx *ctx, const synseg_img *src, const synseg_img *dst, int op, int kw, int kh, int ax, int ay,
                 int iterations, int flags, void *stream);

/* ---- connected components ------------------------------------------------------------------ */
/* cv2.connectedComponentsWithStats(mask, connectivity=8, CV_32S) (also replaces the
 * findContours(EXTERNAL)+boundingRect bar test at S:1403-1404).
 *   labels     : optional int32 image (same width/height/batch; row_stride in bytes), may be NULL
 *   n_labels   : int32[batch]   number of labels including background
 *   stats      : int32[batch][max_labels][5]  = left, top, width, height, area  (row 0 = background)
 *   centroids  : double[batch][max_labels][2] = mean x, mean y
 * Component numbering is cv2's: raster order of the first 2x2 block touching each component.
 * If an image has more than max_labels labels, n_labels is set to -(required) for that image and
 * only the first max_labels rows are written. */
int synseg_ccl_stats(synseg_ctx *ctx, const synseg_img *mask, const synseg_img *labels, int32_t *n_labels,
                     int32_t *stats, double *centroids, int32_t max_labels, void *stream);

/* ---- reductions ---------------------------------------------------------------------------- */
/* Exact integer moments of 8-bit regions: out[i] = {sum, sum of squares, count of non-zero} (uint64 x 3)
 * for each of n_rois regions.  Replaces np.sum(x > 0) (S:1371,1376,1439,1560,1561,1616) and np.var
 * (S:1805,2989,3073; O:1007: var = (n*ss - s*s)/n^2).  rois is a DEVICE array; NULL = whole images.
 * src_kind 0: 8-bit grey image; 1: RGB image reduced through PIL grey; 2: RGB through cv2 grey. */
int synseg_moments(synseg_ctx *ctx, const synseg_img *src, int src_kind, const synseg_roi *rois, int32_t n_rois,
                   uint64_t *out, void *stream);

/* Dominant-colour front end (S:1571-1578): cv2 RGB2HSV S/V + mask S>30 & V>40 & V<240.
 *   count    : uint64[n_rois]  masked-pixel count (the `< 100 -> []` decision, S:1577)
 *   hist     : uint32[n_rois][4096] histogram of (R>>4, G>>4, B>>4) over masked pixels, may be NULL
 *   chan_sum : uint64[n_rois][4096][3] exact per-bin channel sums, may be NULL
 *   row_count: uint32[n_rois][max_rows] masked pixels per region row (for ordered sampling), may be NULL */
int synseg_hsv_mask_hist(synseg_ctx *ctx, const synseg_img *rgb, const synseg_roi *rois, int32_t n_rois,
                         uint64_t *count, uint32_t *hist, uint64_t *chan_sum, uint32_t *row_count, int32_t max_rows,
                         void *stream);
/* Gather the RGB of the k-th masked pixel (raster order, as numpy's img[mask]) for each rank in
 * `ranks` (device int64[n]); row_prefix is the exclusive prefix sum of row_count for that region. */
int synseg_hsv_mask_gather(synseg_ctx *ctx, const synseg_img *rgb, const synseg_roi *roi_host, const uint64_t *row_prefix,
                           const int64_t *ranks, int32_t n, uint8_t *out_rgb, void *stream);

/* ---- perceptual hash ----------------------------------------------------------------------- */
/* 64-bit DCT perceptual hash of each region (PIL grey -> 32x32 cell means -> integer DCT-II ->
 * 8x8 low frequencies -> median bits), exact integer arithmetic (oracle/synseg_oracle.c:orc_phash).
 * Not in the reference (it has only md5(png)[:8], S:3782); SURVEY.md 8a B8.  src_kind as synseg_moments. */
int synseg_phash(synseg_ctx *ctx, const synseg_img *src, int src_kind, const synseg_roi *rois, int32_t n_rois,
                 uint64_t *out, void *stream);
/* Replicated duplicate removal after the all-gather: keep[i] = 0 iff some j with key[j] < key[i] has
 * popcount(hash[i]^hash[j]) <= max_hamming. */
int synseg_phash_dedup(synseg_ctx *ctx, const uint64_t *hashes, const uint64_t *keys, int32_t n, int32_t max_hamming,
                       uint8_t *keep, void *stream);

/* ---- the exchange step over all GPUs ------------------------------------------------------------- */
/* One process per GPU.  Bootstrap: rank 0 calls synseg_comm_unique_id and hands the 128 bytes to the other ranks by any
 * means (the host code uses torch.distributed's store); every rank then calls synseg_comm_init(ctx, id, rank, world).
 * NCCL is bound at run time (the libnccl.so.2 already in the process, else the system one); world == 1 needs none. */
int synseg_comm_unique_id(uint8_t *id128);
int synseg_comm_init(synseg_ctx *ctx, const uint8_t *id128, int32_t rank, int32_t world);
int synseg_comm_destroy(synseg_ctx *ctx);
int synseg_comm_info(const synseg_ctx *ctx, int32_t *rank, int32_t *world, int32_t *nccl_version);
/* Cross-page duplicate removal in one call, everything queued on `stream`, no host synchronisation: this rank's
 * (hashes[i], keys[i]) for i < *count (count: DEVICE int32, clamped to capacity) travel in ONE ncclAllGather of
 * fixed-capacity blocks; then every rank computes, identically,
 *   all_keys  uint64[world * capacity]  the keys of all ranks in ascending order (all_hashes alongside, may be NULL),
 *   keep      uint8 [world * capacity]  0 iff an entry with a smaller key lies within max_hamming bits,
 *   n_total   DEVICE int32              number of valid entries.
 * Keys must be unique over all ranks (page index << 16 | region index). */
int synseg_dedup_exchange(synseg_ctx *ctx, const uint64_t *hashes, const uint64_t *keys, const int32_t *count, int32_t capacity,
                          int32_t max_hamming, uint64_t *all_keys, uint64_t *all_hashes, uint8_t *keep, int32_t *n_total, void *stream);

/* ---- fused page pipeline ------------------------------------------------------------------- */
typedef struct synseg_detect_params {
    int32_t block_size;   /* adaptive threshold window (odd) : (dpi/6)|1      -> 51 @300 DPI */
    int32_t C;            /* adaptive threshold offset       : 10                              */
    int32_t canny_lo;     /* 50  (S:1324) */
    int32_t canny_hi;     /* 150 (S:1324) */
    int32_t k;            /* dilate / close square size     : int(10*dpi/72)|1 -> 41 @300 DPI */
    int32_t max_labels;   /* capacity of stats / centroids per page */
    int32_t channels;     /* 3 (or 0): interleaved RGB pages; 1: grey ('L') pages -- the handoff of S:3638-3657 may be RGB or L, and
                             detection reads grey only (a renderer that produces 'L' saves two thirds of the PCIe bytes) */
    int32_t _pad;
} synseg_detect_params;

/* pages -> component boxes:  grey(cv2) -> adaptive(INV) | Canny -> dilate(k) -> close(k) -> CCL(8)+stats.
 * Same outputs as synseg_ccl_stats (labels omitted).  Bit-exact with the cv2 chain of SURVEY.md 8(d).
 * `rgb` holds RGB pages (params->channels 3 or 0) or grey pages (params->channels 1: the chain starts at the threshold).
 * gray_out may be NULL; if given (RGB pages only) it receives the cv2 grey pages (kept for downstream features). */
int synseg_detect_pages(synseg_ctx *ctx, const synseg_img *rgb, const synseg_detect_params *params,
                        const synseg_img *gray_out, int32_t *n_labels, int32_t *stats, double *centroids, void *stream);

/* The same pipeline for pages in HOST memory (the rasterisation handoff of S:3638-3657: RGB or L u8 (params->channels), `row_stride` bytes per
 * row, `page_stride` bytes per page; pinned memory gives full PCIe speed and true copy/compute overlap, pageable
 * memory works but is staged by the driver).  The library copies `chunk_pages` pages at a time through three device
 * staging slots on its own copy stream, runs synseg_detect_pages on each chunk and copies the tables back to
 *   n_labels_host int32[n_pages], stats_host int32[n_pages][max_labels][5], centroids_host double[n_pages][max_labels][2] (may be NULL).
 * It returns when everything is queued; `stream` waits for the last chunk, so synchronising `stream` (or the device)
 * completes the call and the host buffers must stay valid until then. */
int synseg_detect_pages_host(synseg_ctx *ctx, const void *host_rgb, int32_t width, int32_t height, int64_t row_stride,
                             int64_t page_stride, int32_t n_pages, const synseg_detect_params *params, int32_t chunk_pages,
                             int32_t *n_labels_host, int32_t *stats_host, double *centroids_host, void *stream);

/* Device-side selection of candidate component boxes for hashing (no host round trip): for every image,
 * every component k >= 1 of `stats` (as written by synseg_detect_pages / synseg_ccl_stats) with
 * min_area <= w*h <= max_area, w >= min_w, h >= min_h is appended to rois[] (image, x, y, w, h) and
 * keys[] ((page_base + image) << 16 | k).  count: device int32 (must be zeroed by the caller before the
 * first call; appends are cumulative).  Only the first `capacity` entries are stored; *count keeps counting, so
 * *count > capacity afterwards signals the overflow (readers clamp it to capacity).  Order inside rois[] is
 * unspecified; keys identify the entries. */
int synseg_select_rois(synseg_ctx *ctx, const int32_t *n_labels, const int32_t *stats, int32_t batch, int32_t max_labels,
                       int64_t page_base, int32_t min_area, int32_t max_area, int32_t min_w, int32_t min_h,
                       synseg_roi *rois, uint64_t *keys, int32_t *count, int32_t capacity, void *stream);
/* synseg_phash over a device-resident roi list whose length lives in device memory (`count`, capped at
 * capacity): hashes[i] for i < *count. */
int synseg_phash_indirect(synseg_ctx *ctx, const synseg_img *src, int src_kind, const synseg_roi *rois, const int32_t *count,
                          int32_t capacity, uint64_t *out, void *stream);

/* ---- candidate regions (the box rules of the reference, on the device) ------------------------------------------------ */
#define SYNSEG_REGION_CC 1          /* one large component: the analogue of an embedded-image rect (S:2876-2881)        */
#define SYNSEG_REGION_CLUSTER 2     /* >= 3 small components clustered by the drawing rule (S:3559-3594), padded 10 pt  */
#define SYNSEG_REGION_FLAG_LABELS 1     /* page has more components than max_labels: table truncated, no regions         */
#define SYNSEG_REGION_FLAG_AMBIGUOUS 2  /* a rect pair lies within 1e-9 of the 100 pt cluster threshold: recompute on host */
#define SYNSEG_REGION_FLAG_CAPACITY 4   /* more than max_regions regions: list truncated                                 */

typedef struct synseg_region {
    double x0, y0, x1, y1;        /* box in PDF points, top-left origin (BoundingBox, S:101-122)                        */
    int32_t px, py, pw, ph;       /* the crop in page pixels: round(pt * dpi / 72) (S:3649), clamped to the page        */
    int32_t kind;                 /* SYNSEG_REGION_CC | SYNSEG_REGION_CLUSTER                                           */
    int32_t count;                /* CC: component area in pixels; CLUSTER: number of clustered components              */
    uint64_t sum, sum_sq;         /* exact grey moments of the crop (PIL grey for RGB pages): np.var(L) of S:2988-2989  */
} synseg_region;                  /* 72 bytes */

typedef struct synseg_region_params {
    double dpi;                   /* raster resolution: px = pt * dpi / 72                                              */
    double page_width_pt, page_height_pt;
    double min_extent_pt;         /* a CC region must exceed this in both directions (50 pt, S:3450)                    */
    int32_t max_regions;          /* capacity of the region list per page (<= 1024)                                     */
    int32_t _pad;
} synseg_region_params;

/* Component tables (as written by synseg_detect_pages / synseg_ccl_stats) -> candidate regions of every page + the grey
 * moments of every region crop of `pages` (channels 3: RGB through PIL grey; 1: grey pages).  The device form of
 * detector.py:candidate_regions: area / extent classes (S:3549, S:3450), greedy clustering of the small boxes
 * (_cluster_drawings S:3559-3594, _drawing_distance S:3596-3618), padded cluster boxes (_detect_by_drawings S:3531-3555),
 * duplicate rule (_overlaps_with_existing S:3620-3636).  Bit-identical with the host arithmetic; pages the device cannot
 * decide exactly are flagged (SYNSEG_REGION_FLAG_*) and must be recomputed on the host from the table.
 *   regions : synseg_region[batch][max_regions], n_regions / flags : int32[batch]  (all DEVICE). */
int synseg_regions_from_stats(synseg_ctx *ctx, const int32_t *n_labels, const int32_t *stats, int32_t max_labels,
                              const synseg_img *pages, int channels, const synseg_region_params *rparams,
                              synseg_region *regions, int32_t *n_regions, int32_t *flags, void *stream);

/* synseg_detect_pages + synseg_regions_from_stats in one call on one stream (device pages). centroids may be NULL. */
int synseg_detect_regions(synseg_ctx *ctx, const synseg_img *pages, const synseg_detect_params *params,
                          const synseg_region_params *rparams, int32_t *n_labels, int32_t *stats, double *centroids,
                          synseg_region *regions, int32_t *n_regions, int32_t *flags, void *stream);

/* The same for pages in HOST memory (see synseg_detect_pages_host): per page the library returns, in HOST memory,
 *   n_regions_host int32[n_pages], flags_host int32[n_pages], regions_host synseg_region[n_pages][max_regions],
 *   n_labels_host int32[n_pages] and stats_host int32[n_pages][max_labels][5] (both may be NULL: the table is only
 *   needed to recompute flagged pages on the host). */
int synseg_detect_regions_host(synseg_ctx *ctx, const void *host_pages, int32_t width, int32_t height, int64_t row_stride,
                               int64_t page_stride, int32_t n_pages, const synseg_detect_params *params,
                               const synseg_region_params *rparams, int32_t chunk_pages, int32_t *n_labels_host,
                               int32_t *stats_host, synseg_region *regions_host, int32_t *n_regions_host, int32_t *flags_host,
                               void *stream);

/* ---- renderer-facing page slots (the rasterisation step before the path, S:3638-3657) ----------- */
/* The reference renders a region with MuPDF, PNG-encodes it and decodes it again with PIL (S:3651-3655).  Here a
 * rasteriser writes its pages (RGB or L, params->channels) STRAIGHT into pinned host memory owned by the library:
 *   synseg_page_slots_init    allocates n_slots (2..4) pinned page buffers of pages_per_slot pages (rows padded to 16
 *                             bytes) on the GPU's NUMA node, their device staging buffers and pinned result buffers;
 *   synseg_page_slot_acquire  returns the next slot whose previous submission has been waited for: the pointer to write
 *                             pages into and its strides;
 *   synseg_page_slot_submit   queues H2D (own copy stream), detection, regions + moments and the D2H of the results for
 *                             the first n_pages pages of the slot; returns at once;
 *   synseg_page_slot_wait     blocks until that submission has finished and hands out the slot's result arrays (valid
 *                             until the slot is acquired again): n_regions / flags / n_labels int32[n_pages],
 *                             regions synseg_region[n_pages][max_regions], stats int32[n_pages][max_labels][5];
 *   synseg_page_slots_release frees everything (also done by synseg_destroy). */
int synseg_page_slots_init(synseg_ctx *ctx, int32_t width, int32_t height, int32_t channels, int32_t pages_per_slot,
                           int32_t n_slots, int32_t max_labels, int32_t max_regions);
int synseg_page_slot_acquire(synseg_ctx *ctx, int32_t *slot, void **host_pages, int64_t *row_stride, int64_t *page_stride);
int synseg_page_slot_submit(synseg_ctx *ctx, int32_t slot, int32_t n_pages, const synseg_detect_params *params,
                            const synseg_region_params *rparams, void *stream);
int synseg_page_slot_wait(synseg_ctx *ctx, int32_t slot, const int32_t **n_regions, const int32_t **flags,
                          const synseg_region **regions, const int32_t **n_labels, const int32_t **stats);
int synseg_page_slots_release(synseg_ctx *ctx);
/* NUMA node of the pinned slot memory (-1: unknown / single node) -- bench.py reports it. */
int synseg_page_slots_numa_node(const synseg_ctx *ctx);

/* Per-crop grid-line counts of _detect_grid (S:1546-1564) / _detect_chart_subtype (S:1365-1376):
 * grey (gray_mode) -> Canny(50,150) -> OPEN(kw x 1, it=2) and OPEN(1 x kh, it=2) -> non-zero counts.
 *   out : uint64[n_rois][3] = {h_count, v_count, edge pixel count}
 * kw/kh <= 0 select the chart rule max(20, W/20) / max(20, H/20) per region.
 * edges_out (optional) receives each region's Canny map in a caller buffer of batch=n_rois images
 * of edges_out->width x height (regions larger than that are an error). */
int synseg_grid_counts(synseg_ctx *ctx, const synseg_img *rgb_or_gray, int channels, int gray_mode,
                       const synseg_roi *rois_host, int32_t n_rois, int kw, int kh, uint64_t *out,
                       const synseg_img *edges_out, void *stream);

/* One crop of a ragged batch: `channels` (1 grey / 3 RGB / 4 RGBX, fourth byte ignored -- the layout PIL keeps RGB
 * images in, so a host can hand them over without repacking; offset and row_stride then multiples of 4) x width x
 * height pixels at byte `offset` of a packed device buffer, rows `row_stride` bytes apart. */
typedef struct synseg_crop {
    uint64_t offset;
    int32_t width, height;
    int64_t row_stride;
    int32_t channels;
    int32_t _pad;
} synseg_crop;

/* Batched form of the deterministic per-crop hint quantities (the O:887-1010 drivers over many crops; BASELINE.json
 * configs[3]).  For crop i, out[8*i ..] = { h_count, v_count, edge_px      (as synseg_grid_counts with PIL grey: S:1546-1564,
 *                                           sum, sum_sq, non_zero          (PIL grey moments -> np.var: S:1805, 2989, 3073, O:1007),
 *                                           mask_px                        (HSV mask count, S:1574-1577; 0 for grey crops),
 *                                           0 }.
 * crops_host is a HOST array; everything is queued on `stream` without synchronising.
 * With fixed structuring elements (1 <= kw <= 113, 1 <= kh <= 192) the crops are sorted by size, cut into chunks and
 * every stage runs as ONE launch per chunk over a ragged batch (per-crop width / height in a device table);
 * kw <= 0 / kh <= 0 (the chart rule max(20, W/20) of S:1368-1373) or longer elements run crop by crop (no RGBX). */
int synseg_hints_crops(synseg_ctx *ctx, const void *base, const synseg_crop *crops_host, int32_t n, int kw, int kh,
                       uint64_t *out, void *stream);

/* Batched dominant colours of a ragged batch of crops (the batched form of OCRProcessor._extract_dominant_colors,
 * S:1566-1594; BASELINE.json configs[3]).  Exact: the HSV mask S > 30 & V > 40 & V < 240 (S:1571-1574), the masked
 * pixel count and the `fewer than min_pixels -> no colours` decision (S:1577, min_pixels = 100).  The clustering is a
 * deterministic APPROXIMATION of the reference's KMeans over an unseeded random sample (S:1581-1590): exact 4096-bin
 * (R>>4, G>>4, B>>4) histogram with per-bin channel sums, weighted Lloyd iterations over the bin centroids started
 * from the heaviest bins (arithmetic spelled out in csrc/colors.cu; bit-identical CPU restatement: oracle/colors_port.py).
 * out[(2 + n_colors) * i ..] = { mask_px, k, (cluster pixels << 24 | R << 16 | G << 8 | B) for the k <= n_colors
 * centres (truncated like `.astype(int)`, S:1591), 0 ... }.  hist_out (optional) receives uint32[4096] per crop.
 * Grey crops (channels 1) have no saturated pixel: mask_px = 0, k = 0.  1 <= n_colors <= 8; a crop holds <= 2^24 pixels.
 * crops_host is a HOST array; everything is queued on `stream` without synchronising. */
int synseg_colors_crops(synseg_ctx *ctx, const void *base, const synseg_crop *crops_host, int32_t n, int32_t n_colors,
                        int32_t iters, int32_t min_pixels, uint64_t *out, uint32_t *hist_out, void *stream);

/* The two batched hint calls on regions of pages that are ALREADY in device memory (rois_host: HOST array of (page, x, y, w, h),
 * e.g. the px/py/pw/ph of synseg_region): the kernels read the regions in place, nothing is packed or uploaded again.
 * out as synseg_hints_crops / synseg_colors_crops.  channels 3: RGB pages; 1: grey pages (hints only).  The page allocation must be
 * readable 16 bytes past its last pixel (the 128-bit loads of a region touching the last row may look that far; any
 * cudaMalloc / torch allocation that is not an exact multiple of 512 bytes is). */
int synseg_hints_rois(synseg_ctx *ctx, const synseg_img *pages, int channels, const synseg_roi *rois_host, int32_t n, int kw, int kh,
                      uint64_t *out, void *stream);
int synseg_colors_rois(synseg_ctx *ctx, const synseg_img *pages, const synseg_roi *rois_host, int32_t n, int32_t n_colors,
                       int32_t iters, int32_t min_pixels, uint64_t *out, uint32_t *hist_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SYNSEG_H */
