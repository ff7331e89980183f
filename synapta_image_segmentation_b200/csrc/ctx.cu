// ctx.cu -- context, scratch arena, error reporting.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>

#include "internal.cuh"

static thread_local char g_err[512] = "";

void synseg_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int synseg_check_cuda(cudaError_t e, const char *what)
{
    if (e == cudaSuccess) return 0;
    synseg_set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return SYNSEG_E_CUDA;
}

extern "C" SYNSEG_EXPORT int synseg_version(void) { return SYNSEG_VERSION; }
extern "C" SYNSEG_EXPORT const char *synseg_last_error(void) { return g_err; }

extern "C" SYNSEG_EXPORT int synseg_create(int device, synseg_ctx **out)
{
    if (!out) { synseg_set_error("synseg_create: out is NULL"); return SYNSEG_E_INVALID; }
    *out = nullptr;
    int count = 0;
    SS_CUDA(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) { synseg_set_error("synseg_create: no CUDA device %d", device); return SYNSEG_E_INVALID; }
    DeviceScope scope(device);          // the caller's current device is restored on return
    cudaDeviceProp prop;
    SS_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        synseg_set_error("synseg_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                         prop.minor);
        return SYNSEG_E_INVALID;
    }
    synseg_ctx *c = new synseg_ctx();
    c->arena = nullptr; c->arena_bytes = 0; c->arena_top = 0; c->launches = 0; c->phash_basis = nullptr;
    c->prof_on = false; c->prof_start = nullptr; c->prof_used = 0;
    c->device = device;
    c->comm = nullptr; c->comm_world = 1; c->comm_rank = 0;
    c->guard_violations = 0; c->guard_checked = 0;
    c->attr_done = 0; c->last_stream = nullptr; c->ev_last = nullptr; c->last_valid = false;
    memset(&c->hs, 0, sizeof(c->hs));
    const char *e1 = getenv("SYNSEG_TUNE_AD_BAND"), *e2 = getenv("SYNSEG_TUNE_CANNY_BAND");
    c->tune_ad_band = e1 ? atoi(e1) : 0;
    c->tune_canny_band = e2 ? atoi(e2) : 0;
    c->sm_count = prop.multiProcessorCount;
    const char *e3 = getenv("SYNSEG_OVERLAP");
    const char *e4 = getenv("SYNSEG_STREAMS");
    c->overlap = e3 ? atoi(e3) : SYNSEG_OVERLAP_DEFAULT;
    c->overlap_streams = e4 ? atoi(e4) : 3;      // 3 chunks on 3 streams: 1.34 ms per 50-page step against 1.39-1.41 with 2 on 2 (current kernels)
    if (c->overlap_streams < 2) c->overlap_streams = 2;
    if (c->overlap_streams > 4) c->overlap_streams = 4;
    c->ev_split_fork = nullptr;
    for (int i = 0; i < 3; ++i) { c->aux[i] = nullptr; c->ev_split_join[i] = nullptr; }
    // integer DCT basis of the perceptual hash (same formula as oracle/synseg_oracle.c:orc_phash_basis)
    int32_t basis[8 * 32];
    for (int u = 0; u < 8; ++u)
        for (int x = 0; x < 32; ++x) basis[u * 32 + x] = (int32_t)lround(16384.0 * cos(M_PI * (2 * x + 1) * u / 64.0));
    int rc = synseg_check_cuda(cudaMalloc(&c->phash_basis, sizeof(basis)), "cudaMalloc(phash basis)");
    if (!rc) rc = synseg_check_cuda(cudaMemcpy(c->phash_basis, basis, sizeof(basis), cudaMemcpyHostToDevice), "cudaMemcpy(basis)");
    if (!rc) rc = synseg_check_cuda(cudaEventCreateWithFlags(&c->ev_last, cudaEventDisableTiming), "cudaEventCreate(ev_last)");
    if (rc) { if (c->phash_basis) cudaFree(c->phash_basis); delete c; return rc; }
    *out = c;
    return SYNSEG_OK;
}

extern "C" SYNSEG_EXPORT int synseg_destroy(synseg_ctx *ctx)
{
    if (!ctx) return SYNSEG_OK;
    DeviceScope scope(ctx->device);
    cudaDeviceSynchronize();
    if (ctx->ev_last) cudaEventDestroy(ctx->ev_last);
    if (ctx->arena) cudaFree(ctx->arena);
    if (ctx->phash_basis) cudaFree(ctx->phash_basis);
    host_stream_release(ctx);
    comm_release(ctx);
    if (ctx->ev_split_fork) cudaEventDestroy(ctx->ev_split_fork);
    for (int i = 0; i < 3; ++i) {
        if (ctx->aux[i]) cudaStreamDestroy(ctx->aux[i]);
        if (ctx->ev_split_join[i]) cudaEventDestroy(ctx->ev_split_join[i]);
    }
    if (ctx->prof_start) cudaEventDestroy(ctx->prof_start);
    for (cudaEvent_t e : ctx->prof_events) cudaEventDestroy(e);
    delete ctx;
    return SYNSEG_OK;
}

extern "C" SYNSEG_EXPORT int64_t synseg_launch_count(const synseg_ctx *ctx) { return ctx ? ctx->launches : 0; }

void arena_begin(synseg_ctx *ctx) { guard_flush(ctx); ctx->arena_top = 0; }

int arena_ensure(synseg_ctx *ctx, size_t bytes)
{
    if (bytes + SS_GUARD_SLACK <= ctx->arena_bytes) return SYNSEG_OK;
    DeviceScope scope(ctx->device);
    SS_CUDA(cudaDeviceSynchronize());
    if (ctx->arena) { cudaFree(ctx->arena); ctx->arena = nullptr; ctx->arena_bytes = 0; }
    size_t want = align_up(bytes + SS_GUARD_SLACK, (size_t)1 << 20);
    cudaError_t e = cudaMalloc(&ctx->arena, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        synseg_set_error("scratch arena: cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e));
        return SYNSEG_E_NOMEM;
    }
    ctx->arena_bytes = want;
    return SYNSEG_OK;
}

int arena_alloc(synseg_ctx *ctx, size_t bytes, void **out, cudaStream_t st)
{
    size_t off = align_up(ctx->arena_top, 256) + SS_GUARD_ZONE;
    const size_t end = align_up(off + bytes, 256) + SS_GUARD_ZONE;
    if (end > ctx->arena_bytes) {
        synseg_set_error("scratch arena overflow: need %zu, have %zu (internal sizing error)", end, ctx->arena_bytes);
        return SYNSEG_E_NOMEM;
    }
    *out = ctx->arena + off;
    ctx->arena_top = end;
#ifdef SYNSEG_GUARD
    // canary zones right below the payload and right above its 256-byte rounded end, filled in stream order before any kernel of this call
    SS_CUDA(cudaMemsetAsync(ctx->arena + off - SS_GUARD_ZONE, 0xA5, SS_GUARD_ZONE, st));
    SS_CUDA(cudaMemsetAsync(ctx->arena + end - SS_GUARD_ZONE, 0xA5, SS_GUARD_ZONE, st));
    ctx->guard_recs.push_back(synseg_ctx::GuardRec{off, end - SS_GUARD_ZONE - off});
#else
    (void)st;
#endif
    return SYNSEG_OK;
}

// Guard build: called when a public call returns and wherever scratch is about to be reused (arena_begin / arena_release /
// arena_rebase).  Waits for the device, compares every canary zone of the allocations made since the last flush, counts and
// reports the damaged ones.
void guard_flush(synseg_ctx *ctx)
{
#ifdef SYNSEG_GUARD
    if (ctx->guard_recs.empty()) return;
    if (ctx->in_capture) { ctx->guard_recs.clear(); return; }     // a captured call: no synchronisation allowed, its canaries are not compared
    std::vector<synseg_ctx::GuardRec> recs;
    recs.swap(ctx->guard_recs);
    if (cudaDeviceSynchronize() != cudaSuccess) { cudaGetLastError(); return; }
    unsigned char host[2 * 256];
    for (const synseg_ctx::GuardRec &r : recs) {
        if (r.off + r.bytes + SS_GUARD_ZONE > ctx->arena_bytes) continue;            // the arena was re-allocated meanwhile
        if (cudaMemcpy(host, ctx->arena + r.off - SS_GUARD_ZONE, SS_GUARD_ZONE, cudaMemcpyDeviceToHost) != cudaSuccess ||
            cudaMemcpy(host + 256, ctx->arena + r.off + r.bytes, SS_GUARD_ZONE, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); continue; }
        ctx->guard_checked += 2;
        int bad_lo = -1, bad_hi = -1;
        for (int i = 0; i < 256; ++i) { if (host[i] != 0xA5 && bad_lo < 0) bad_lo = i; if (host[256 + i] != 0xA5 && bad_hi < 0) bad_hi = i; }
        if (bad_lo >= 0 || bad_hi >= 0) {
            ctx->guard_violations++;
            synseg_set_error("SYNSEG_GUARD: canary of the scratch allocation at arena offset %zu (%zu bytes) was overwritten (%s zone, byte %d)",
                             r.off, r.bytes, bad_lo >= 0 ? "lower" : "upper", bad_lo >= 0 ? bad_lo : bad_hi);
            fprintf(stderr, "%s\n", synseg_last_error());
        }
    }
#else
    (void)ctx;
#endif
}

// Number of damaged canary zones seen so far / zones compared (both 0 in a normal build); *guard_build = 1 in a SYNSEG_GUARD build.
extern "C" SYNSEG_EXPORT int64_t synseg_guard_violations(const synseg_ctx *ctx, int64_t *zones_checked, int32_t *guard_build)
{
#ifdef SYNSEG_GUARD
    if (guard_build) *guard_build = 1;
#else
    if (guard_build) *guard_build = 0;
#endif
    if (zones_checked) *zones_checked = ctx ? ctx->guard_checked : 0;
    return ctx ? ctx->guard_violations : 0;
}

extern "C" SYNSEG_EXPORT int synseg_reserve(synseg_ctx *ctx, size_t bytes)
{
    if (!ctx) { synseg_set_error("synseg_reserve: ctx is NULL"); return SYNSEG_E_INVALID; }
    return arena_ensure(ctx, bytes);
}

// Upper bound of the scratch any single public call needs for a batch of width x height images:
// class map (1 B/px) + grey (1 B/px) + 2x2-block labels (1 B/px) + a few bit planes + accumulators.
extern "C" SYNSEG_EXPORT size_t synseg_scratch_bytes(int32_t width, int32_t height, int32_t batch)
{
    size_t px = (size_t)align_up((size_t)width, 128) * (size_t)(height + 2);
    size_t per = 4 * px + 8 * ((size_t)bit_wpr(width) * 4 * height) + ((size_t)1 << 16);
    return per * (size_t)batch + ((size_t)8 << 20);
}

int validate_img(const synseg_img *im, const char *name, int channels)
{
    if (!im || !im->data) { synseg_set_error("%s: image is NULL", name); return SYNSEG_E_INVALID; }
    if (im->width <= 0 || im->height <= 0 || im->batch <= 0) {
        synseg_set_error("%s: bad shape %d x %d x %d", name, im->batch, im->height, im->width);
        return SYNSEG_E_INVALID;
    }
    if (im->row_stride < (int64_t)im->width * channels) {
        synseg_set_error("%s: row_stride %lld < %d bytes per row", name, (long long)im->row_stride, im->width * channels);
        return SYNSEG_E_INVALID;
    }
    if (im->batch > 1 && im->batch_stride < im->row_stride * (int64_t)im->height) {
        synseg_set_error("%s: batch_stride %lld too small", name, (long long)im->batch_stride);
        return SYNSEG_E_INVALID;
    }
    return SYNSEG_OK;
}

// ---- per-kernel timing ---------------------------------------------------------------------------
// All kernels of a call run back to back on one stream, so the time between the events recorded after
// consecutive launches is the duration of the later kernel (plus any memset/memcpy issued in between).
void prof_mark(synseg_ctx *ctx, const char *name, cudaStream_t st)
{
    if (ctx->prof_used == ctx->prof_events.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        ctx->prof_events.push_back(e);
        ctx->prof_names.push_back(name);
    }
    ctx->prof_names[ctx->prof_used] = name;
    cudaEventRecord(ctx->prof_events[ctx->prof_used], st);
    ctx->prof_used++;
}

extern "C" SYNSEG_EXPORT int synseg_profile_begin(synseg_ctx *ctx, void *stream)
{
    if (!ctx) { synseg_set_error("synseg_profile_begin: ctx is NULL"); return SYNSEG_E_INVALID; }
    SS_ENTER(ctx, stream);
    if (!ctx->prof_start) SS_CUDA(cudaEventCreate(&ctx->prof_start));
    ctx->prof_used = 0;
    ctx->prof_on = true;
    SS_CUDA(cudaEventRecord(ctx->prof_start, (cudaStream_t)stream));
    return SYNSEG_OK;
}

// Stops profiling, waits for the last recorded event and returns up to `cap` (name, milliseconds) pairs in
// launch order.  names[i] points to a static string.  Returns the number of launches recorded (may exceed cap).
extern "C" SYNSEG_EXPORT int synseg_profile_end(synseg_ctx *ctx, const char **names, float *ms, int cap)
{
    if (!ctx) { synseg_set_error("synseg_profile_end: ctx is NULL"); return SYNSEG_E_INVALID; }
    ctx->prof_on = false;
    const int n = (int)ctx->prof_used;
    if (n == 0) return 0;
    SS_CUDA(cudaEventSynchronize(ctx->prof_events[n - 1]));
    cudaEvent_t prev = ctx->prof_start;
    for (int i = 0; i < n; ++i) {
        float t = 0.f;
        SS_CUDA(cudaEventElapsedTime(&t, prev, ctx->prof_events[i]));
        if (i < cap) { names[i] = ctx->prof_names[i]; ms[i] = t; }
        prev = ctx->prof_events[i];
    }
    return n;
}
