// internal.cuh -- shared declarations for libsynseg.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "../../include/synseg.h"

#define SYNSEG_EXPORT __attribute__((visibility("default")))

#ifndef SYNSEG_OVERLAP_DEFAULT
#define SYNSEG_OVERLAP_DEFAULT 3
#endif

#ifndef __CUDA_ARCH__
#define SYNSEG_HOST 1
#endif

// ---- memory-safety build (SYNSEG_NVCC_EXTRA="-DSYNSEG_GUARD", see tools/guard_run.sh) -------------------------------
// compute-sanitizer is closed on this pool; this build is the substitute:
//  * every scratch-arena allocation sits between two 256-byte canary zones (0xA5) that are filled when the allocation is
//    made and compared on the host when the public call that made it returns (the call synchronises its stream first);
//    a changed canary is counted (synseg_guard_violations) and described in synseg_last_error;
//  * SS_DEVICE_ASSERT(index in range) inside the kernels is a real device assert (the kernel stops, every later call fails).
#ifdef SYNSEG_GUARD
#include <assert.h>
#define SS_DEVICE_ASSERT(cond) assert(cond)
constexpr size_t SS_GUARD_ZONE = 256;
constexpr size_t SS_GUARD_SLACK = (size_t)1 << 20;     // added to every scratch estimate: room for the canary zones
#else
#define SS_DEVICE_ASSERT(cond) ((void)0)
constexpr size_t SS_GUARD_ZONE = 0;
constexpr size_t SS_GUARD_SLACK = 0;
#endif

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
struct synseg_ctx {
    int device;
    int sm_count;
    uint8_t *arena;       // scratch arena (device)
    size_t arena_bytes;
    size_t arena_top;     // bump pointer, reset at the start of every public call
    int64_t launches;     // kernels launched through this context
    int32_t *phash_basis; // device int32[8*32]
    struct GuardRec { size_t off, bytes; };             // guard build: payload offsets of the allocations of the running call
    std::vector<GuardRec> guard_recs;
    bool in_capture = false;   // the current public call runs inside a stream capture (set by CallGuard): the guard build must not synchronise
    int64_t guard_violations, guard_checked;
    void *comm;           // ncclComm_t of the dedup exchange (exchange.cu), NULL on a single GPU
    int comm_world, comm_rank;
    uint32_t attr_done;   // ATTR_* bits: cudaFuncSetAttribute calls already made on THIS device (the attribute is per device)
    // stream ordering of the shared scratch arena: every public call records `ev_last` on its stream when it returns; a
    // call arriving on a different stream first waits for it (SS_ENTER)
    cudaStream_t last_stream;
    cudaEvent_t ev_last;
    bool last_valid;
    int tune_ad_band;     // experiment knobs (env SYNSEG_TUNE_AD_BAND / SYNSEG_TUNE_CANNY_BAND), 0 = automatic
    int tune_canny_band;
    // Stage overlap inside synseg_detect_pages (pipeline.cu): side streams owned by the context, forked from and joined
    // back into the caller's stream with events, so the call stays asynchronous on the caller's stream.
    int overlap;                            // env SYNSEG_OVERLAP: page chunks per call, dealt round-robin to `overlap_streams` streams (1 = one chain)
    int overlap_streams;                    // env SYNSEG_STREAMS: the caller's stream + up to 3 side streams (default 2)
    cudaStream_t aux[3];                    // side streams (created on first use)
    cudaEvent_t ev_split_fork, ev_split_join[3];
    // host-buffer streaming (synseg_detect_pages_host / synseg_detect_regions_host / page slots): device staging ring +
    // copy stream, created on first use.  Ring slots rotate ACROSS calls and carry their own `done` event, so the first
    // copies of a call only wait for the work that last used the same slot (not for everything queued before).
    struct HostStream {
        static constexpr int MAXS = 4;
        cudaStream_t copy;                 // H2D stream
        uint8_t *pages[MAXS];              // staging slots for `slot_pages` pages each
        uint8_t *raw[MAXS];                // grey pages whose host rows are not 16-byte multiples land here first (1-D copy at full
        size_t raw_bytes;                  // PCIe speed) and are re-pitched on the device; a strided 2-D H2D copy runs at a third of it
        int32_t *n_labels[MAXS], *stats[MAXS];
        int32_t *n_regions[MAXS];          // n_regions[slot_pages] followed by flags[slot_pages]
        double *centroids[MAXS];
        synseg_region *regions[MAXS];
        cudaEvent_t copied[MAXS], done[MAXS];
        bool used[MAXS];                   // `done` has been recorded at least once
        int next;                          // next ring slot
        size_t slot_bytes;                 // bytes per page slot buffer
        int slot_pages, max_labels, max_regions;
        bool ready;
        // renderer-facing pinned page slots (synseg_page_slot_*)
        struct PageSlot {
            uint8_t *pages;                // pinned host pages
            int32_t *ints;                 // pinned: n_labels | n_regions | flags, ps_pages each
            int32_t *stats;                // pinned
            synseg_region *regions;        // pinned
            cudaEvent_t finished;
            int state;                     // 0 free, 1 acquired, 2 submitted
            int n_pages;
        } ps[MAXS];
        int ps_n, ps_next, ps_width, ps_height, ps_channels, ps_pages, ps_max_labels, ps_max_regions, ps_numa;
        int64_t ps_row_stride, ps_page_stride;
    } hs;
    // optional per-kernel timing (synseg_profile_*): one event after every launch on the profiled stream
    bool prof_on;
    cudaEvent_t prof_start;
    std::vector<cudaEvent_t> prof_events;
    std::vector<const char *> prof_names;
    size_t prof_used;
};

enum : uint32_t { ATTR_BITMORPH_H = 1u, ATTR_BITMORPH_VH = 2u, ATTR_BITMORPH_V = 4u, ATTR_MORPH_U8_H = 8u, ATTR_MORPH_U8_V = 16u, ATTR_HSV_HIST = 32u,
                  ATTR_FRONT = 64u, ATTR_MORPH2D = 128u, ATTR_HYST_SWEEP = 256u };

// Entry guard of every public call that queues work (SS_ENTER below):
//  * makes the context's device current for the duration of the call and restores the caller's device afterwards
//    (a context for GPU 1 used under current device 0 must neither change the caller's device nor mix devices);
//  * orders the call behind the previous call on this context when that one was issued to a DIFFERENT stream -- all calls
//    share one scratch arena, so two in-flight calls are only safe when the second waits for the first.
void guard_flush(synseg_ctx *ctx);

struct CallGuard {
    synseg_ctx *c;
    cudaStream_t st;
    int prev;
    bool switched, capturing;
    CallGuard(synseg_ctx *ctx, cudaStream_t stream) : c(ctx), st(stream), prev(-1), switched(false), capturing(false)
    {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != c->device) { cudaSetDevice(c->device); switched = true; }
        // A call that is being captured into a CUDA graph is ordered by the graph: an event recorded during capture must not be waited
        // for by streams outside it (and vice versa), so the cross-stream ordering of the scratch arena is left to whoever replays.
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(st, &cs) == cudaSuccess) capturing = cs != cudaStreamCaptureStatusNone; else cudaGetLastError();
        c->in_capture = capturing;
        if (!capturing && c->last_valid && c->last_stream != st && c->ev_last) cudaStreamWaitEvent(st, c->ev_last, 0);
    }
    ~CallGuard()
    {
#ifdef SYNSEG_GUARD
        if (!capturing) guard_flush(c);
#endif
        c->in_capture = false;
        if (!capturing) {
            if (c->ev_last && cudaEventRecord(c->ev_last, st) == cudaSuccess) { c->last_stream = st; c->last_valid = true; }
            else cudaGetLastError();
        }
        if (switched) cudaSetDevice(prev);
    }
};
#define SS_ENTER(ctx, stream) CallGuard _ss_guard((ctx), (cudaStream_t)(stream))

// Restores the caller's current device (for the few internal paths that must call cudaSetDevice outside SS_ENTER).
struct DeviceScope {
    int prev; bool switched;
    explicit DeviceScope(int dev) : prev(-1), switched(false) { if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) { cudaSetDevice(dev); switched = true; } }
    ~DeviceScope() { if (switched) cudaSetDevice(prev); }
};

void prof_mark(synseg_ctx *ctx, const char *name, cudaStream_t st);
void host_stream_release(synseg_ctx *ctx);
void comm_release(synseg_ctx *ctx);

void synseg_set_error(const char *fmt, ...);
int synseg_check_cuda(cudaError_t e, const char *what);

#define SS_CUDA(call)                                                    \
    do {                                                                 \
        int _rc = synseg_check_cuda((call), #call);                      \
        if (_rc) return _rc;                                             \
    } while (0)

#define SS_TRY(call)                 \
    do {                             \
        int _rc = (call);            \
        if (_rc) return _rc;         \
    } while (0)

#define SS_LAUNCH_CHECK(ctx, name, stream)                               \
    do {                                                                 \
        (ctx)->launches++;                                               \
        int _rc = synseg_check_cuda(cudaGetLastError(), name);           \
        if (_rc) return _rc;                                             \
        if ((ctx)->prof_on) prof_mark((ctx), name, (stream));            \
    } while (0)

// Scratch arena: bump allocation, 256-byte aligned.  arena_begin() at the start of a public call.
void arena_begin(synseg_ctx *ctx);
int arena_alloc(synseg_ctx *ctx, size_t bytes, void **out, cudaStream_t stream);
static inline size_t arena_mark(const synseg_ctx *ctx) { return ctx->arena_top; }
// Every point where scratch is about to be reused first compares the canaries of what was allocated so far (guard build; no-op otherwise).
void guard_flush(synseg_ctx *ctx);
static inline void arena_release(synseg_ctx *ctx, size_t mark) { guard_flush(ctx); ctx->arena_top = mark; }
static inline void arena_rebase(synseg_ctx *ctx, size_t base) { guard_flush(ctx); ctx->arena_top = base; }
// Grows the arena to hold `bytes` in total (synchronises the device when it has to reallocate).
int arena_ensure(synseg_ctx *ctx, size_t bytes);

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

int validate_img(const synseg_img *im, const char *name, int channels);
static inline bool same_shape(const synseg_img *a, const synseg_img *b)
{
    return a->width == b->width && a->height == b->height && a->batch == b->batch;
}

// A plane view used by kernels (by value).
struct Plane {
    uint8_t *p;
    int64_t rs;  // row stride (bytes)
    int64_t bs;  // batch stride (bytes)
};
static inline Plane plane_of(const synseg_img *im) { return Plane{(uint8_t *)im->data, im->row_stride, im->batch_stride}; }
static inline bool plane_aligned(const synseg_img *im, int a)
{
    return (((uintptr_t)im->data | (uintptr_t)im->row_stride | (uintptr_t)(im->batch > 1 ? im->batch_stride : 0)) % a) == 0;
}

// Bit-packed masks: one bit per pixel, LSB = leftmost pixel, `wpr` 32-bit words per row
// (multiple of 4 so rows are 16-byte aligned), bits at x >= width are always 0.
// Ragged batches (crops of different sizes processed by ONE launch per stage): the planes of a batch share one canvas
// geometry (wpr / bs sized for the widest / tallest image) and `dims` gives the real width and height of every image;
// a kernel that finds dims != NULL takes its width / height from there and skips tasks that lie outside the image.
// Words at or beyond cdiv(width, 32) and rows at or beyond height are then undefined and never read.
struct BitPlane {
    uint32_t *p;
    int wpr;        // words per row
    int64_t bs;     // words per image
    const int2 *dims;   // per-image (width, height) of a ragged batch, or NULL: every image has the launch's width x height
};
static inline int bit_wpr(int width) { return (int)align_up((size_t)cdiv(width, 32), 4); }

// ------------------------------------------------------------------------------------------------
// internal launchers (one per .cu file)
// ------------------------------------------------------------------------------------------------
int launch_rgb2gray(synseg_ctx *ctx, const synseg_img *rgb, const synseg_img *gray, int mode, cudaStream_t st);

int launch_adaptive_mean(synseg_ctx *ctx, const synseg_img *gray, const synseg_img *out_u8, BitPlane out_bits,
                         int block_size, int C, int invert, cudaStream_t st);

// Canny stage 1: grey -> two bit planes: kept (survived non-maximum suppression, mag > lo) and strong (kept, mag > hi).
int launch_canny_classes(synseg_ctx *ctx, const synseg_img *gray, BitPlane kept, BitPlane strong, int lo, int hi, cudaStream_t st);
// Full Canny.  Exactly one of edges_u8 / edges_bits receives the result (the other NULL / {nullptr}).
// or_bits: when edges_bits is given, OR into it instead of overwriting.
int run_canny(synseg_ctx *ctx, const synseg_img *gray, const synseg_img *edges_u8, BitPlane edges_bits, bool or_bits,
              int lo, int hi, cudaStream_t st);

// Fused RGB front end (canny.cu): TMA-fed RGB -> grey + Canny classes, then threshold + hysteresis into `mask`.
bool canny_rgb_supported(const synseg_img *rgb, const synseg_img *gray);
int launch_canny_rgb(synseg_ctx *ctx, const synseg_img *rgb, const synseg_img *gray, BitPlane kept, BitPlane strong, int lo, int hi, cudaStream_t st);
int run_front_rgb(synseg_ctx *ctx, const synseg_img *rgb, const synseg_img *gray, BitPlane mask, int block_size, int C, int lo, int hi, cudaStream_t st);

// bit packing
int launch_pack_bits(synseg_ctx *ctx, const synseg_img *src, BitPlane dst, cudaStream_t st);
int launch_unpack_bits(synseg_ctx *ctx, BitPlane src, const synseg_img *dst, cudaStream_t st);
int launch_count_bits(synseg_ctx *ctx, BitPlane src, int width, int height, int batch, uint64_t *out, int out_stride,
                      cudaStream_t st);

// 1-D rect morphology passes on bit planes. op 0 erode, 1 dilate. Window for output i is
// [i - anchor, i - anchor + k - 1] clipped to the image.
int launch_bitmorph_h(synseg_ctx *ctx, BitPlane src, BitPlane dst, int width, int height, int batch, int op, int k,
                      int anchor, cudaStream_t st);
int launch_bitmorph_v(synseg_ctx *ctx, BitPlane src, BitPlane dst, int width, int height, int batch, int op, int k,
                      int anchor, cudaStream_t st);
// dilate(K1 x K1, a1) then erode(k2 x k2, a2) in one shared-memory kernel (morph_fused.cu); *done = false: not applicable, nothing launched
int launch_bit_dilate_erode(synseg_ctx *ctx, BitPlane src, BitPlane dst, int width, int height, int batch, int K1, int a1, int k2, int a2,
                            bool *done, cudaStream_t st);
// Generic 8-bit 1-D passes.
int launch_morph_u8_h(synseg_ctx *ctx, const synseg_img *src, const synseg_img *dst, int op, int k, int anchor,
                      cudaStream_t st);
int launch_morph_u8_v(synseg_ctx *ctx, const synseg_img *src, const synseg_img *dst, int op, int k, int anchor,
                      cudaStream_t st);
// Whole rect op (erode/dilate/open/close with iterations folded) on bit planes.
// The result is left in `cur` (the two planes are swapped as passes ping-pong).
int run_bitmorph(synseg_ctx *ctx, BitPlane &cur, BitPlane &other, int width, int height, int batch, int op, int kw, int kh,
                 int ax, int ay, int iterations, cudaStream_t st);

// Connected components.  Mask given either as u8 plane (non-zero = foreground) or bit plane.
struct CclMask {
    const synseg_img *u8;  // or NULL
    BitPlane bits;         // used when u8 == NULL
    int width, height, batch;
};
int run_ccl_stats(synseg_ctx *ctx, const CclMask &m, const synseg_img *labels, int32_t *n_labels, int32_t *stats,
                  double *centroids, int32_t max_labels, cudaStream_t st);
// Hysteresis: keep the pixels of `kept` whose 8-connected component (through kept pixels) holds a `strong` pixel.
int run_hysteresis(synseg_ctx *ctx, BitPlane kept, BitPlane strong, int width, int height, int batch, const synseg_img *edges_u8,
                   BitPlane edges_bits, bool or_bits, cudaStream_t st);

// Hysteresis by propagation sweeps (hyst_sweep.cu); sw: device int32[n_sweeps]; sw[n_sweeps - 1] != 0 afterwards = not converged
int launch_hyst_sweeps(synseg_ctx *ctx, BitPlane kept, BitPlane strong, BitPlane out, bool or_bits, int width, int height, int batch,
                       int32_t *sw, int n_sweeps, bool *used, cudaStream_t st);

size_t ccl_label_scratch_bytes(int width, int height, int batch);
size_t hysteresis_scratch_bytes(int width, int height, int batch);
size_t canny_scratch_bytes(int width, int height, int batch);
size_t ccl_stats_scratch_bytes(int width, int height, int batch, int max_labels);

// Ragged crop front end (gray.cu): packed crops -> PIL-grey canvas + grey moments + HSV mask count in one pass over the
// source.  tasks / res are DEVICE arrays; res[8*j + 3..6] += {sum, sum_sq, non_zero, mask_px} of crop j.
struct CropTask {
    int64_t offset, row_stride;     // bytes from `base`
    int32_t width, height, channels, out_index;
};
int launch_crop_front(synseg_ctx *ctx, const void *base, const CropTask *tasks, int n, const synseg_img *gray_canvas, uint64_t *res,
                      cudaStream_t st);

// regions.cu: component tables -> candidate regions + crop moments (the arena must hold regions_scratch_bytes from its top)
size_t regions_scratch_bytes(int batch, int max_labels);
int validate_region_params(const synseg_region_params *rp, const char *who);
int run_regions(synseg_ctx *ctx, const int32_t *n_labels, const int32_t *stats, int32_t max_labels, const synseg_img *pages, int channels,
                const synseg_region_params *rp, synseg_region *regions, int32_t *n_regions, int32_t *flags, cudaStream_t st);

int launch_moments(synseg_ctx *ctx, const synseg_img *src, int src_kind, const synseg_roi *rois, int32_t n_rois,
                   uint64_t *out, cudaStream_t st);
