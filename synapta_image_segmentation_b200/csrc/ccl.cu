// ccl.cu -- 8-connected component labelling with fused statistics, bit-exact with
// cv2.connectedComponentsWithStats(mask, 8, CV_32S) including cv2's label numbering, and the
// hysteresis stage of Canny (canny.cu) on the same union-find core.
//
// No reference call site for the labelling itself (north-star primitive); it also
// replaces findContours(RETR_EXTERNAL)+boundingRect at pdf_image_segmentation.py:1403-1404.
//
// Data layout: the mask is a bit plane (1 bit / pixel).  The unit of the union-find is a BLOCK RUN:
//   * a block is a 2x2 pixel cell (all foreground pixels of a cell are mutually 8-connected);
//   * blocks bx, bx+1 of one block row are linked iff column 2bx+1 and column 2bx+2 both hold a
//     foreground pixel in the two pixel rows of the block row (any such pair is 8-adjacent);
//   * a block run is a maximal chain of linked blocks; it is identified by its first block, and only
//     that block owns an entry in the parent array L (int32 per block, touched sparsely).
// A group of 32 or 16 lanes owns one block row (half-warp rows where a row wastes fewer lanes; the union kernel
// always uses a full warp); a lane owns a 64-pixel chunk (one 64-bit word per pixel row); runs are
// found with bit arithmetic and a ballot-based carry across lanes.  Vertical contacts between block
// rows are unions between runs, issued once per touching (run, run) pair:
//   t = top pixel row of the block row, u = pixel row above it
//   event A at x: t[x] & !t[x-1] & (u[x] | u[x-1])  -> union(run of t[x], run of u[x] or else u[x-1])
//   event B at x: u[x] & !u[x-1] & t[x-1]           -> union(run of t[x-1], run of u[x])
// so a solid blob costs one atomic per block row instead of four per block.  Unions are lock-free
// (atomicMin towards the smaller block index, find with path halving); the root of a component is its
// first block in 2x2-block raster order, which is exactly the order in which cv2's block-based
// algorithm numbers components, so labels are 1 + rank of the root (root bit plane + per-row counts
// + scan).  Statistics are reduced per run piece (popcounts of the 64-bit masks) -> warp redux over
// lanes sharing a label -> per-CTA shared-memory cache -> one global atomic set per CTA and label.
// Centroids are sum/area in f64 like cv2.
//
// Roofline: HBM/L2-bound on the bit plane; algorithmic bytes 1/8 per pixel read (+4 per pixel when a
// label image is requested).  The parent array is touched only at run starts.
#include <stdlib.h>

#include "internal.cuh"

namespace {

typedef unsigned long long u64;
constexpr u64 EVEN = 0x5555555555555555ULL;
constexpr unsigned FULL = 0xffffffffu;
constexpr int ROWS_PER_CTA = 8;     // one warp per block row

struct RowGeom {
    const uint32_t *bits; int wpr; int64_t wbs;   // bit plane: words per row / per image
    int width, height;
    int bw, bh;        // blocks per block row, block rows
    int64_t bper;      // parent-array entries per image
    int nseg;          // segments per row: G lanes x 64 pixels each (G = lanes per block row, 32 or 16)
    int cpr;           // 64-pixel chunks per row = cdiv(width, 64)
    const int2 *dims;  // ragged batch (internal.cuh): per-image width / height; everything above then describes the canvas
    const int32_t *need;   // hysteresis behind the propagation sweeps (hyst_sweep.cu): the kernels return at once for image i when need[i * need_stride] == 0
    int need_stride;
};

// Lanes per block row.  G = 32: one warp per block row.  G = 16: a warp owns two consecutive block rows, one per
// half-warp (shuffles with width 16, ballots split per half) -- a 2550-pixel row (40 chunks) then needs 3 segments of
// 16 lanes instead of 2 of 32, i.e. 25 % fewer warp instructions per row.  Every *_sync intrinsic is still executed
// by the full warp.

// 64 pixels of one row; bits at x >= width (and rows outside the image) read as 0
__device__ __forceinline__ u64 load_chunk(const uint32_t *row, int chunk, int width)
{
    const int rem = width - 64 * chunk;
    if (row == nullptr || rem <= 0) return 0;
    const uint2 w = __ldg((const uint2 *)row + chunk);
    u64 v = ((u64)w.y << 32) | w.x;
    if (rem < 64) v &= (1ULL << rem) - 1ULL;
    return v;
}

// ---- run structure of one block row ---------------------------------------------------------
struct RunCarry { uint32_t topbit; int start; };   // warp-uniform state carried from one segment to the next
struct Runs {
    u64 occE;      // even bit 2i: block i of this chunk holds a foreground pixel
    u64 startE;    // even bit 2i: block i starts a run
    int carryIn;   // first block (column index) of the run that holds the last block of the previous chunk
};

// v = OR of the two pixel rows of the block row (this lane's chunk).  All 32 lanes must call; `gl` = lane within
// its group of G lanes (one group per block row).
template <int G>
__device__ __forceinline__ Runs analyze_runs(u64 v, int chunk, int gl, RunCarry &c)
{
    const uint32_t hi = (uint32_t)(v >> 63);
    uint32_t pb = __shfl_up_sync(FULL, hi, 1, G);
    if (gl == 0) pb = c.topbit;
    const u64 vprev = (v << 1) | pb;                   // vprev[x] = v[x-1]
    Runs r;
    r.occE = (v | (v >> 1)) & EVEN;
    r.startE = r.occE & ~(v & vprev);                  // not linked to the previous block
    unsigned has = __ballot_sync(FULL, r.startE != 0);
    if (G == 16) has = (has >> (threadIdx.x & 16)) & 0xFFFFu;      // this half-warp's lanes
    const int my_last = r.startE ? chunk * 32 + ((63 - __clzll((long long)r.startE)) >> 1) : -1;
    const unsigned below = has & ((1u << gl) - 1u);
    const int from_lane = __shfl_sync(FULL, my_last, below ? 31 - __clz((int)below) : 0, G);
    r.carryIn = below ? from_lane : c.start;
    const int last_all = __shfl_sync(FULL, my_last, has ? 31 - __clz((int)has) : 0, G);
    if (has) c.start = last_all;
    c.topbit = __shfl_sync(FULL, hi, G - 1, G);
    return r;
}

// first block (column index) of the run holding the block at even bit position `pos` of this chunk
__device__ __forceinline__ int run_start(const Runs &r, int chunk, int pos)
{
    const u64 m = r.startE & ((2ULL << pos) - 1ULL);
    return m ? chunk * 32 + ((63 - __clzll((long long)m)) >> 1) : r.carryIn;
}

// Next run piece of a chunk (maximal chain of linked blocks inside the chunk).  `rem` = blocks not yet
// visited (even bits).  Returns the pixel-column range mask of the piece and its run's first block.
__device__ __forceinline__ u64 next_piece(const Runs &r, int chunk, u64 &rem, int &first_block)
{
    const int p = __ffsll((long long)rem) - 1;
    const u64 brk = (~r.occE | r.startE) & EVEN;
    const u64 above = (p >= 62) ? 0ULL : (brk & (~0ULL << (p + 2)));
    const int e = above ? __ffsll((long long)above) - 1 : 64;
    const u64 range = ((e >= 64) ? ~0ULL : ((1ULL << e) - 1ULL)) & (~0ULL << p);
    rem &= ~range;
    first_block = ((r.startE >> p) & 1ULL) ? chunk * 32 + (p >> 1) : r.carryIn;   // a piece that is not a start begins at p == 0
    return range;
}

// pixel-column range of the run piece (inside this chunk) that holds the block at even bit position `pos`
__device__ __forceinline__ u64 piece_range(const Runs &r, int pos)
{
    const u64 m = r.startE & ((2ULL << pos) - 1ULL);
    const int sp = m ? 63 - __clzll((long long)m) : 0;          // a piece without a start in the chunk begins at bit 0
    const u64 brk = (~r.occE | r.startE) & EVEN;
    const u64 above = (pos >= 62) ? 0ULL : (brk & (~0ULL << (pos + 2)));
    const int e = above ? __ffsll((long long)above) - 1 : 64;
    return ((e >= 64) ? ~0ULL : ((1ULL << e) - 1ULL)) & (~0ULL << sp);
}

// ---- union-find -----------------------------------------------------------------------------
// Parents only ever decrease and every value written is an ancestor, so stale reads are harmless;
// L2-coherent loads because other SMs write concurrently.
__device__ __forceinline__ int32_t uf_find(int32_t *L, int32_t i)
{
    for (;;) {
        const int32_t p = __ldcg(L + i);
        if (p == i) return i;
        const int32_t gp = __ldcg(L + p);
        if (gp == p) return p;
        atomicMin(L + i, gp);          // path halving; only non-root entries are touched
        i = gp;
    }
}

__device__ __forceinline__ void uf_union(int32_t *L, int32_t a, int32_t b)
{
    bool done;
    do {
        a = uf_find(L, a);
        b = uf_find(L, b);
        if (a < b) { const int32_t old = atomicMin(&L[b], a); done = (old == b); b = old; }
        else if (b < a) { const int32_t old = atomicMin(&L[a], b); done = (old == a); a = old; }
        else done = true;
    } while (!done);
}

// lane = lane within the group (0..G-1); rows outside the image read as empty (null row pointers) so that a half-warp
// whose block row does not exist still takes part in the warp-wide intrinsics
#define ROW_PROLOGUE(first_row)                                                             \
    if (g_.need && __ldg(g_.need + blockIdx.y * g_.need_stride) == 0) return;      /* the sweeps converged for this image */ \
    RowGeom g = g_;                                                                          \
    if (g_.dims) { const int2 d_ = g_.dims[blockIdx.y]; g.width = d_.x; g.height = d_.y; g.bh = (d_.y + 1) >> 1; } \
    const int lane = threadIdx.x & (G - 1);                                                  \
    const int by = (first_row) + (blockIdx.x * ROWS_PER_CTA + (threadIdx.x >> 5)) * (32 / G) + ((threadIdx.x & 31) / G); \
    const int img = blockIdx.y;                                                              \
    const bool row_ok = by < g.bh;                                                           \
    const bool warp_ok = (by - ((threadIdx.x & 31) / G)) < g.bh;     /* first row of this warp exists */ \
    const uint32_t *pbase = g.bits + img * g.wbs;                                            \
    const uint32_t *r0 = row_ok ? pbase + (int64_t)(2 * by) * g.wpr : nullptr;               \
    const uint32_t *r1 = (row_ok && 2 * by + 1 < g.height) ? r0 + g.wpr : nullptr;

// every run start becomes its own parent
template <int G>
__global__ void __launch_bounds__(256) rccl_init_kernel(RowGeom g_, int32_t *Lall)
{
    ROW_PROLOGUE(0)
    if (!warp_ok) return;
    int32_t *L = Lall + img * g.bper + (int64_t)by * g.bw;
    RunCarry c{0u, -1};
    for (int s = 0; s < g.nseg; ++s) {
        const int chunk = s * G + lane;
        const u64 v = load_chunk(r0, chunk, g.width) | load_chunk(r1, chunk, g.width);
        const Runs r = analyze_runs<G>(v, chunk, lane, c);
        u64 st = r.startE;
        while (st) {
            const int bx = chunk * 32 + ((__ffsll((long long)st) - 1) >> 1);
            st &= st - 1;
            SS_DEVICE_ASSERT(bx >= 0 && bx < g.bw && by < g.bh);
            L[bx] = by * g.bw + bx;
        }
    }
}

// unions between the runs of block row `by` and those of block row by - 1.
// HYST (hysteresis): a union between two runs that BOTH hold a strong pixel is skipped -- both are kept anyway, and
// every weak run still reaches a strong run of its component through unions that involve at least one weak run
// (take a path to the nearest strong run: all its edges but none beyond have a weak end).  The test is made on the
// run pieces inside this lane's chunk (a subset of the runs), so it only ever skips safely.
// resident CTAs per SM the labelling's union kernel is compiled for: the kernel waits on L2 (row loads, parent look-ups), more warps hide
// more of it -- 5 (48 registers, 56 bytes of spills) measured 0.096 ms per 50 pages against 0.109 at 3 (68 registers), 0.103 at 4, 0.100 at 6
#ifndef SYNSEG_MERGE_MINB
#define SYNSEG_MERGE_MINB 5
#endif
template <bool HYST, int G>
__global__ void __launch_bounds__(256, HYST ? 1 : SYNSEG_MERGE_MINB) rccl_merge_kernel(RowGeom g_, int32_t *Lall, BitPlane strong)
{
    ROW_PROLOGUE(1)
    if (!warp_ok) return;
    const uint32_t *u1 = row_ok ? r0 - g.wpr : nullptr, *u0 = row_ok ? u1 - g.wpr : nullptr;    // pixel rows 2by-1, 2by-2
    int32_t *L = Lall + img * g.bper;
    const int cur_base = by * g.bw, up_base = (by - 1) * g.bw;
    RunCarry cc{0u, -1}, cu{0u, -1};
    uint32_t t_top = 0, u_top = 0;
    for (int s = 0; s < g.nseg; ++s) {
        const int chunk = s * G + lane;
        const u64 t = load_chunk(r0, chunk, g.width), a1 = load_chunk(r1, chunk, g.width);
        const u64 u = load_chunk(u1, chunk, g.width), b0 = load_chunk(u0, chunk, g.width);
        // (an early exit for segments without any pixel -- one vote after the four loads -- was measured slower: 0.117 -> 0.155 ms per 50 pages)
        const Runs rc = analyze_runs<G>(t | a1, chunk, lane, cc);
        const Runs ru = analyze_runs<G>(u | b0, chunk, lane, cu);
        const uint32_t th = (uint32_t)(t >> 63), uh = (uint32_t)(u >> 63);
        uint32_t tp = __shfl_up_sync(FULL, th, 1, G), up = __shfl_up_sync(FULL, uh, 1, G);
        if (lane == 0) { tp = t_top; up = u_top; }
        t_top = __shfl_sync(FULL, th, G - 1, G); u_top = __shfl_sync(FULL, uh, G - 1, G);
        const u64 tl = (t << 1) | tp, ul = (u << 1) | up;      // tl[x] = t[x-1]
        u64 ev_a = t & ~tl & (u | ul);
        u64 ev_b = u & ~ul & tl;
        u64 sc = 0, su = 0;                                     // strong pixels of the current / upper block row
        if (HYST && (ev_a | ev_b)) {
            const uint32_t *s0 = strong.p + img * strong.bs + (int64_t)(2 * by) * strong.wpr;
            sc = load_chunk(s0, chunk, g.width) | load_chunk(r1 ? s0 + strong.wpr : nullptr, chunk, g.width);
            su = load_chunk(s0 - strong.wpr, chunk, g.width) | load_chunk(s0 - 2 * strong.wpr, chunk, g.width);
        }
        while (ev_a) {
            const int x = __ffsll((long long)ev_a) - 1;
            ev_a &= ev_a - 1;
            const int ux = ((u >> x) & 1ULL) ? x : x - 1;       // upper pixel of the contact (-1: last pixel of the previous chunk)
            if (HYST && ux >= 0 && (sc & piece_range(rc, x & ~1)) && (su & piece_range(ru, ux & ~1))) continue;
            const int cs = run_start(rc, chunk, x & ~1);
            const int us = (ux >= 0) ? run_start(ru, chunk, ux & ~1) : ru.carryIn;
            SS_DEVICE_ASSERT(cs >= 0 && cs < g.bw && us >= 0 && us < g.bw);
            uf_union(L, cur_base + cs, up_base + us);
        }
        while (ev_b) {
            const int x = __ffsll((long long)ev_b) - 1;
            ev_b &= ev_b - 1;
            if (HYST && x > 0 && (sc & piece_range(rc, (x - 1) & ~1)) && (su & piece_range(ru, x & ~1))) continue;
            const int us = run_start(ru, chunk, x & ~1);
            const int cs = (x > 0) ? run_start(rc, chunk, (x - 1) & ~1) : rc.carryIn;
            uf_union(L, cur_base + cs, up_base + us);
        }
    }
}

// Two consecutive block rows per warp (G = 32): the pixel rows and the run analysis of block row `by` serve twice, as the lower
// partner of by - 1 and as the upper partner of by + 1 -- six row loads and three analyses for two rows of unions instead of eight and
// four, and all six loads of a segment are in flight together.  Same events, same unions as rccl_merge_kernel.
template <bool HYST>
__device__ __forceinline__ void merge_events(int32_t *L, int chunk, int width, u64 t, u64 u, uint32_t tp, uint32_t up, const Runs &rc, const Runs &ru,
                                             int cur_base, int up_base, const uint32_t *s_cur0, const uint32_t *s_cur1, const uint32_t *s_up1,
                                             const uint32_t *s_up0)
{
    const u64 tl = (t << 1) | tp, ul = (u << 1) | up;      // tl[x] = t[x-1]
    u64 ev_a = t & ~tl & (u | ul);
    u64 ev_b = u & ~ul & tl;
    u64 sc = 0, su = 0;                                     // strong pixels of the current / upper block row
    if (HYST && (ev_a | ev_b)) {
        sc = load_chunk(s_cur0, chunk, width) | load_chunk(s_cur1, chunk, width);
        su = load_chunk(s_up1, chunk, width) | load_chunk(s_up0, chunk, width);
    }
    while (ev_a) {
        const int x = __ffsll((long long)ev_a) - 1;
        ev_a &= ev_a - 1;
        const int ux = ((u >> x) & 1ULL) ? x : x - 1;       // upper pixel of the contact (-1: last pixel of the previous chunk)
        if (HYST && ux >= 0 && (sc & piece_range(rc, x & ~1)) && (su & piece_range(ru, ux & ~1))) continue;
        const int cs = run_start(rc, chunk, x & ~1);
        const int us = (ux >= 0) ? run_start(ru, chunk, ux & ~1) : ru.carryIn;
        uf_union(L, cur_base + cs, up_base + us);
    }
    while (ev_b) {
        const int x = __ffsll((long long)ev_b) - 1;
        ev_b &= ev_b - 1;
        if (HYST && x > 0 && (sc & piece_range(rc, (x - 1) & ~1)) && (su & piece_range(ru, x & ~1))) continue;
        const int us = run_start(ru, chunk, x & ~1);
        const int cs = (x > 0) ? run_start(rc, chunk, (x - 1) & ~1) : rc.carryIn;
        uf_union(L, cur_base + cs, up_base + us);
    }
}

template <bool HYST>
__global__ void __launch_bounds__(256, HYST ? 1 : SYNSEG_MERGE_MINB) rccl_merge2_kernel(RowGeom g_, int32_t *Lall, BitPlane strong)
{
    constexpr int G = 32;
    if (g_.need && __ldg(g_.need + blockIdx.y * g_.need_stride) == 0) return;      // the sweeps converged for this image
    int width = g_.width, height = g_.height, bh = g_.bh;
    if (g_.dims) { const int2 d_ = g_.dims[blockIdx.y]; width = d_.x; height = d_.y; bh = (d_.y + 1) >> 1; }
    const int bw = g_.bw, wpr = g_.wpr, nseg = g_.nseg;
    const int lane = threadIdx.x & 31;
    const int by = 1 + 2 * (blockIdx.x * ROWS_PER_CTA + (threadIdx.x >> 5));       // this warp: block rows by and by + 1
    const int img = blockIdx.y;
    if (by >= bh) return;                                                        // warp-uniform
    const bool two = by + 1 < bh;
    const uint32_t *pbase = g_.bits + img * g_.wbs;
    // pixel rows p[i] = 2 by - 2 + i: block row by - 1 (0, 1), by (2, 3), by + 1 (4, 5); rows outside the image read as empty
    const uint32_t *p[6], *sp[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const int y = 2 * by - 2 + i;
        const bool ok = y < height && (i < 4 || two);
        p[i] = ok ? pbase + (int64_t)y * wpr : nullptr;
        sp[i] = (HYST && ok) ? strong.p + img * strong.bs + (int64_t)y * strong.wpr : nullptr;
    }
    int32_t *L = Lall + img * g_.bper;
    RunCarry c0{0u, -1}, c1{0u, -1}, c2{0u, -1};
    uint32_t top1 = 0, top2 = 0, top3 = 0, top4 = 0;       // last bit of the previous segment of the pixel rows p[1] .. p[4]
    for (int s = 0; s < nseg; ++s) {
        const int chunk = s * G + lane;
        u64 v[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) v[i] = load_chunk(p[i], chunk, width);
        const Runs r0 = analyze_runs<G>(v[0] | v[1], chunk, lane, c0);
        const Runs r1 = analyze_runs<G>(v[2] | v[3], chunk, lane, c1);
        const uint32_t h1 = (uint32_t)(v[1] >> 63), h2 = (uint32_t)(v[2] >> 63);
        uint32_t up1 = __shfl_up_sync(FULL, h1, 1), tp2 = __shfl_up_sync(FULL, h2, 1);
        if (lane == 0) { up1 = top1; tp2 = top2; }
        top1 = __shfl_sync(FULL, h1, 31); top2 = __shfl_sync(FULL, h2, 31);
        merge_events<HYST>(L, chunk, width, v[2], v[1], tp2, up1, r1, r0, by * bw, (by - 1) * bw, sp[2], sp[3], sp[1], sp[0]);
        if (two) {                                          // warp-uniform
            const Runs r2 = analyze_runs<G>(v[4] | v[5], chunk, lane, c2);
            const uint32_t h3 = (uint32_t)(v[3] >> 63), h4 = (uint32_t)(v[4] >> 63);
            uint32_t up3 = __shfl_up_sync(FULL, h3, 1), tp4 = __shfl_up_sync(FULL, h4, 1);
            if (lane == 0) { up3 = top3; tp4 = top4; }
            top3 = __shfl_sync(FULL, h3, 31); top4 = __shfl_sync(FULL, h4, 31);
            merge_events<HYST>(L, chunk, width, v[4], v[3], tp4, up3, r2, r1, (by + 1) * bw, by * bw, sp[4], sp[5], sp[3], sp[2]);
        }
    }
}

// Full compression of every run start.
//  MODE 0 (labelling): the roots of the block row go to a bit plane (even-bit layout, one u64 per chunk)
//                      and their number to row_count.
//  MODE 1 (hysteresis): the root of every run piece that holds a strong pixel is flagged.
template <int MODE, int G>
__global__ void __launch_bounds__(256) rccl_compress_kernel(RowGeom g_, int32_t *Lall, u64 *rootbits, int32_t *row_count,
                                                            BitPlane strong, uint32_t *flags, int64_t fper)
{
    ROW_PROLOGUE(0)
    if (!warp_ok) return;
    int32_t *L = Lall + img * g.bper;
    const int cur_base = by * g.bw;
    RunCarry c{0u, -1};
    int nroots = 0;
    const uint32_t *s0 = nullptr, *s1 = nullptr;
    if (MODE == 1 && row_ok) {
        s0 = strong.p + img * strong.bs + (int64_t)(2 * by) * strong.wpr;
        s1 = (2 * by + 1 < g.height) ? s0 + strong.wpr : nullptr;
    }
    for (int s = 0; s < g.nseg; ++s) {
        const int chunk = s * G + lane;
        const u64 v = load_chunk(r0, chunk, g.width) | load_chunk(r1, chunk, g.width);
        const Runs r = analyze_runs<G>(v, chunk, lane, c);
        u64 st = r.startE, rootE = 0;
        while (st) {
            const int p = __ffsll((long long)st) - 1;
            st &= st - 1;
            const int32_t b = cur_base + chunk * 32 + (p >> 1);
            const int32_t root = uf_find(L, b);
            if (root != b) L[b] = root; else rootE |= 1ULL << p;
        }
        if (MODE == 0) {
            if (row_ok && chunk < g.cpr) rootbits[((int64_t)img * g.bh + by) * g.cpr + chunk] = rootE;
            nroots += __popcll(rootE);
        } else {
            const u64 sv = load_chunk(s0, chunk, g.width) | load_chunk(s1, chunk, g.width);
            u64 rem = sv ? r.occE : 0ULL;
            while (rem) {
                int fb;
                const u64 range = next_piece(r, chunk, rem, fb);
                if (sv & range) {
                    const int32_t root = uf_find(L, cur_base + fb);
                    uint32_t *fw = flags + img * fper + (root >> 5);
                    const uint32_t bit = 1u << (root & 31);
                    if (!(__ldcg(fw) & bit)) atomicOr(fw, bit);
                }
            }
        }
    }
    if (MODE == 0) {
        nroots = __reduce_add_sync(G == 32 ? FULL : (0xFFFFu << (threadIdx.x & 16)), nroots);      // over the lanes of this block row
        if (lane == 0 && row_ok) row_count[(int64_t)img * g.bh + by] = nroots;
    }
}

// one CTA per image: exclusive scan of the per-row root counts (in place), total -> n_roots[img]
__global__ void __launch_bounds__(256) ccl_scan_kernel(int n, int32_t *cnt, int32_t *n_roots)
{
    const int img = blockIdx.x;
    int32_t *c = cnt + (int64_t)img * n;
    __shared__ int ws[8];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 256) {
        const int i = base + threadIdx.x;
        const int v = (i < n) ? c[i] : 0;
        int s = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { int nb = __shfl_up_sync(FULL, s, d); if ((threadIdx.x & 31) >= d) s += nb; }
        if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = s;
        __syncthreads();
        int woff = 0;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) woff += ws[w];
        const int excl = carry + woff + s - v;
        if (i < n) c[i] = excl;
        __syncthreads();
        if (threadIdx.x == 255) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) n_roots[img] = carry;
}

// roots get L[root] = -(label) - 1 with label = 1 + rank in block raster order
template <int G>
__global__ void __launch_bounds__(256) rccl_assign_kernel(RowGeom g_, int32_t *Lall, const u64 *rootbits, const int32_t *row_off)
{
    ROW_PROLOGUE(0)
    (void)r0; (void)r1; (void)pbase;
    if (!warp_ok) return;
    int32_t *L = Lall + img * g.bper + (int64_t)by * g.bw;
    int base = row_ok ? row_off[(int64_t)img * g.bh + by] : 0;
    for (int s = 0; s < g.nseg; ++s) {
        const int chunk = s * G + lane;
        u64 rb = (row_ok && chunk < g.cpr) ? rootbits[((int64_t)img * g.bh + by) * g.cpr + chunk] : 0ULL;
        const int n = __popcll(rb);
        int incl = n;
#pragma unroll
        for (int d = 1; d < G; d <<= 1) { int nb = __shfl_up_sync(FULL, incl, d, G); if (lane >= d) incl += nb; }
        int rank = base + incl - n;
        while (rb) {
            const int bx = chunk * 32 + ((__ffsll((long long)rb) - 1) >> 1);
            rb &= rb - 1;
            L[bx] = -(rank + 1) - 1;
            ++rank;
        }
        base += __shfl_sync(FULL, incl, G - 1, G);
    }
}

// ---- statistics ---------------------------------------------------------------------------
struct StatAcc {          // SoA accumulators, [batch][cap]
    int32_t *minx, *miny, *maxx, *maxy, *area;
    u64 *sumx, *sumy;
    int cap;
};

__global__ void stats_init_kernel(StatAcc a, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    a.minx[i] = 0x7fffffff; a.miny[i] = 0x7fffffff; a.maxx[i] = -1; a.maxy[i] = -1; a.area[i] = 0;
    a.sumx[i] = 0; a.sumy[i] = 0;
}

constexpr int NSLOT = 32;
struct SlotCache {
    int key[NSLOT];
    int minx[NSLOT], miny[NSLOT], maxx[NSLOT], maxy[NSLOT], area[NSLOT];
    u64 sumx[NSLOT], sumy[NSLOT];
    // Background (label 0): only its bounding box is measured.  Its area and coordinate sums are the image totals minus the
    // foreground's: every foreground contribution is also added here and leaves the CTA as ONE negative contribution to the
    // accumulators of label 0 (two's complement, exact); stats_finalize adds W H, H W (W - 1) / 2 and W H (H - 1) / 2.
    int bg_minx, bg_miny, bg_maxx, bg_maxy, fg_area;
    u64 fg_sumx, fg_sumy;
};

struct Contrib { int minx, miny, maxx, maxy, area; unsigned int sumx, sumy; };

// sum of the set bit positions of a 64-bit mask
__device__ __forceinline__ unsigned pos_sum(u64 m)
{
    return (unsigned)__popcll(m & 0xAAAAAAAAAAAAAAAAULL) + 2u * __popcll(m & 0xCCCCCCCCCCCCCCCCULL) +
           4u * __popcll(m & 0xF0F0F0F0F0F0F0F0ULL) + 8u * __popcll(m & 0xFF00FF00FF00FF00ULL) +
           16u * __popcll(m & 0xFFFF0000FFFF0000ULL) + 32u * __popcll(m & 0xFFFFFFFF00000000ULL);
}

// pixels m0 (row y) and m1 (row y + 1) of a 64-pixel chunk starting at column x0; m0 | m1 != 0
__device__ __forceinline__ Contrib contrib_of(u64 m0, u64 m1, int x0, int y)
{
    Contrib v;
    const int n0 = __popcll(m0), n1 = __popcll(m1);
    const u64 both = m0 | m1;
    v.area = n0 + n1;
    v.minx = x0 + __ffsll((long long)both) - 1;
    v.maxx = x0 + 63 - __clzll((long long)both);
    v.miny = m0 ? y : y + 1;
    v.maxy = m1 ? y + 1 : y;
    v.sumx = (unsigned)v.area * (unsigned)x0 + pos_sum(m0) + pos_sum(m1);
    v.sumy = (unsigned)v.area * (unsigned)y + (unsigned)n1;
    return v;
}

// Lanes in `peers` share `key`: reduce their contributions; the leader adds into the CTA cache or global.
__device__ __forceinline__ void accumulate_group(unsigned peers, int key, Contrib v, SlotCache &sc, const StatAcc &a, int64_t img_off)
{
    v.minx = __reduce_min_sync(peers, v.minx); v.miny = __reduce_min_sync(peers, v.miny);
    v.maxx = __reduce_max_sync(peers, v.maxx); v.maxy = __reduce_max_sync(peers, v.maxy);
    v.area = __reduce_add_sync(peers, v.area);
    v.sumx = __reduce_add_sync(peers, v.sumx); v.sumy = __reduce_add_sync(peers, v.sumy);
    if ((int)(threadIdx.x & 31) != __ffs((int)peers) - 1) return;
    if (v.area == 0) return;
    atomicAdd(&sc.fg_area, v.area); atomicAdd(&sc.fg_sumx, (u64)v.sumx); atomicAdd(&sc.fg_sumy, (u64)v.sumy);
    if (key >= a.cap) return;                           // over capacity: reported through n_labels
    const int slot = key & (NSLOT - 1);
    const int old = atomicCAS(&sc.key[slot], -1, key);
    if (old == -1 || old == key) {
        atomicMin(&sc.minx[slot], v.minx); atomicMin(&sc.miny[slot], v.miny);
        atomicMax(&sc.maxx[slot], v.maxx); atomicMax(&sc.maxy[slot], v.maxy);
        atomicAdd(&sc.area[slot], v.area);
        atomicAdd(&sc.sumx[slot], (u64)v.sumx); atomicAdd(&sc.sumy[slot], (u64)v.sumy);
    } else {
        const int64_t i = img_off + key;
        atomicMin(&a.minx[i], v.minx); atomicMin(&a.miny[i], v.miny);
        atomicMax(&a.maxx[i], v.maxx); atomicMax(&a.maxy[i], v.maxy);
        atomicAdd(&a.area[i], v.area);
        atomicAdd(&a.sumx[i], (u64)v.sumx); atomicAdd(&a.sumy[i], (u64)v.sumy);
    }
}

// Per-warp staging for the optional label image: block labels and pixel words of one 2048-pixel segment.
struct LabelStage { int32_t lab[1024]; u64 px[2][32]; };

#ifndef SYNSEG_FINAL_MINB
#define SYNSEG_FINAL_MINB 5
#endif
template <bool WRITE_LABELS, int G>
__global__ void __launch_bounds__(256, WRITE_LABELS ? 1 : SYNSEG_FINAL_MINB) rccl_final_kernel(RowGeom g_, const int32_t *Lall, Plane labels, bool labels_al16, StatAcc a)
{
    static_assert(!WRITE_LABELS || G == 32, "the label image path stages one block row per warp");
    __shared__ SlotCache sc;
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    if (threadIdx.x < NSLOT) {
        const int t = threadIdx.x;
        sc.key[t] = -1; sc.minx[t] = 0x7fffffff; sc.miny[t] = 0x7fffffff; sc.maxx[t] = -1; sc.maxy[t] = -1;
        sc.area[t] = 0; sc.sumx[t] = 0; sc.sumy[t] = 0;
        if (t == 0) { sc.bg_minx = 0x7fffffff; sc.bg_miny = 0x7fffffff; sc.bg_maxx = -1; sc.bg_maxy = -1; sc.fg_area = 0; sc.fg_sumx = 0; sc.fg_sumy = 0; }
    }
    __syncthreads();
    ROW_PROLOGUE(0)
    const int64_t img_off = (int64_t)img * a.cap;
    if (warp_ok) {
        const int32_t *L = Lall + img * g.bper;
        const int cur_base = by * g.bw;
        const int y = 2 * by;
        const bool has_row1 = y + 1 < g.height;
        LabelStage *stage = WRITE_LABELS ? (LabelStage *)dyn_smem + (threadIdx.x >> 5) : nullptr;
        RunCarry c{0u, -1};
        for (int s = 0; s < g.nseg; ++s) {
            const int chunk = s * G + lane;
            const int x0 = 64 * chunk;
            const u64 a0 = load_chunk(r0, chunk, g.width), a1 = load_chunk(r1, chunk, g.width);
            const Runs r = analyze_runs<G>(a0 | a1, chunk, lane, c);
            // background (label 0): all lanes reduce together (both block rows of a warp belong to the same image)
            {
                const int remw = row_ok ? g.width - x0 : 0;
                const u64 vm = remw >= 64 ? ~0ULL : (remw <= 0 ? 0ULL : ((1ULL << remw) - 1ULL));
                const u64 b0 = ~a0 & vm, b1 = has_row1 ? (~a1 & vm) : 0ULL, bb = b0 | b1;
                int mnx = bb ? x0 + __ffsll((long long)bb) - 1 : 0x7fffffff, mxx = bb ? x0 + 63 - __clzll((long long)bb) : -1;
                int mny = b0 ? y : (b1 ? y + 1 : 0x7fffffff), mxy = b1 ? y + 1 : (b0 ? y : -1);
                mnx = __reduce_min_sync(FULL, mnx); mny = __reduce_min_sync(FULL, mny);
                mxx = __reduce_max_sync(FULL, mxx); mxy = __reduce_max_sync(FULL, mxy);
                if ((threadIdx.x & 31) == 0 && mxx >= 0) {
                    atomicMin(&sc.bg_minx, mnx); atomicMin(&sc.bg_miny, mny); atomicMax(&sc.bg_maxx, mxx); atomicMax(&sc.bg_maxy, mxy);
                }
            }
            if (WRITE_LABELS) {
                int4 *z = (int4 *)(stage->lab + lane * 32);
#pragma unroll
                for (int i = 0; i < 8; ++i) z[i] = make_int4(0, 0, 0, 0);
                stage->px[0][lane] = a0; stage->px[1][lane] = a1;
            }
            u64 rem = r.occE;
            for (;;) {
                const bool has = rem != 0;
                const unsigned active = __ballot_sync(FULL, has);
                if (!active) break;
                if (has) {
                    int fb;
                    const int p = __ffsll((long long)rem) - 1;
                    (void)p;
                    const u64 range = next_piece(r, chunk, rem, fb);
                    int32_t v = L[cur_base + fb];
                    if (v >= 0) v = L[v];
                    const int label = -v - 1;
                    if (WRITE_LABELS) {
                        const int e = 64 - __clzll((long long)range);     // one past the last column of the piece
                        for (int q = p; q < e; q += 2) stage->lab[lane * 32 + (q >> 1)] = label;
                    }
                    const unsigned peers = __match_any_sync(active, label);
                    accumulate_group(peers, label, contrib_of(a0 & range, a1 & range, x0, y), sc, a, img_off);
                }
            }
            if (WRITE_LABELS) {
                __syncwarp();
                const int seg_x = s * 2048;
                for (int rr = 0; rr < (has_row1 ? 2 : 1); ++rr) {
                    int32_t *lrow = (int32_t *)(labels.p + img * labels.bs + (int64_t)(y + rr) * labels.rs);
#pragma unroll 4
                    for (int i = 0; i < 16; ++i) {
                        const int xx = 128 * i + 4 * lane;
                        const int x = seg_x + xx;
                        if (x >= g.width) continue;
                        const uint32_t bits = (uint32_t)(stage->px[rr][xx >> 6] >> (xx & 63)) & 0xFu;
                        const int l01 = stage->lab[xx >> 1], l23 = stage->lab[(xx >> 1) + 1];
                        const int4 o = make_int4((bits & 1u) ? l01 : 0, (bits & 2u) ? l01 : 0, (bits & 4u) ? l23 : 0, (bits & 8u) ? l23 : 0);
                        if (labels_al16 && x + 3 < g.width) *(int4 *)(lrow + x) = o;
                        else {
                            lrow[x] = o.x;
                            if (x + 1 < g.width) lrow[x + 1] = o.y;
                            if (x + 2 < g.width) lrow[x + 2] = o.z;
                            if (x + 3 < g.width) lrow[x + 3] = o.w;
                        }
                    }
                }
                __syncwarp();
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < NSLOT && sc.key[threadIdx.x] >= 0 && sc.area[threadIdx.x] > 0) {
        const int t = threadIdx.x;
        const int64_t i = img_off + sc.key[t];
        atomicMin(&a.minx[i], sc.minx[t]); atomicMin(&a.miny[i], sc.miny[t]);
        atomicMax(&a.maxx[i], sc.maxx[t]); atomicMax(&a.maxy[i], sc.maxy[t]);
        atomicAdd(&a.area[i], sc.area[t]);
        atomicAdd(&a.sumx[i], sc.sumx[t]); atomicAdd(&a.sumy[i], sc.sumy[t]);
    }
    if (threadIdx.x == NSLOT) {                       // label 0: bounding box of the background, minus the foreground's sums
        if (sc.bg_maxx >= 0) {
            atomicMin(&a.minx[img_off], sc.bg_minx); atomicMin(&a.miny[img_off], sc.bg_miny);
            atomicMax(&a.maxx[img_off], sc.bg_maxx); atomicMax(&a.maxy[img_off], sc.bg_maxy);
        }
        if (sc.fg_area) {
            atomicAdd(&a.area[img_off], -sc.fg_area);
            atomicAdd(&a.sumx[img_off], 0ULL - sc.fg_sumx); atomicAdd(&a.sumy[img_off], 0ULL - sc.fg_sumy);
        }
    }
}

// width / height: of the image (per image from `dims` for a ragged batch): label 0 holds minus the foreground totals (see SlotCache)
__global__ void stats_finalize_kernel(StatAcc a, const int32_t *n_roots, int batch, int32_t *n_labels, int32_t *stats, double *centroids,
                                      int width, int height, const int2 *dims)
{
    const int img = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = n_roots[img] + 1;
    if (k == 0) n_labels[img] = (n <= a.cap) ? n : -n;
    if (k >= a.cap || k >= n) return;
    const int64_t i = (int64_t)img * a.cap + k;
    int32_t *s = stats + i * 5;
    if (k == 0) {
        if (dims) { width = dims[img].x; height = dims[img].y; }
        const u64 W = (u64)width, H = (u64)height;
        a.area[i] += (int)(W * H);
        a.sumx[i] += H * (W * (W - 1) / 2);
        a.sumy[i] += W * (H * (H - 1) / 2);
    }
    const int area = a.area[i];
    if (area > 0) {
        s[0] = a.minx[i]; s[1] = a.miny[i]; s[2] = a.maxx[i] - a.minx[i] + 1; s[3] = a.maxy[i] - a.miny[i] + 1; s[4] = area;
        if (centroids) {
            centroids[2 * i] = (double)a.sumx[i] / (double)area;
            centroids[2 * i + 1] = (double)a.sumy[i] / (double)area;
        }
    } else {   // only possible for the background of an all-foreground image; cv2 4.13 reports exactly this row
        s[0] = -1; s[1] = 0x7fffffff; s[2] = 0; s[3] = 0; s[4] = 0;
        if (centroids) {
            centroids[2 * i] = __longlong_as_double(0x7ff8000000000000LL);
            centroids[2 * i + 1] = __longlong_as_double(0x7ff8000000000000LL);
        }
    }
}

// ---- hysteresis output -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t bytes_of_nibble(uint32_t nib) { return ((nib * 0x00204081u) & 0x01010101u) * 0xFFu; }

// keep the pixels of every run piece whose component was flagged (holds a strong pixel)
template <bool OUT_BITS, int G>
__global__ void __launch_bounds__(256) rccl_hyst_final_kernel(RowGeom g_, const int32_t *Lall, const uint32_t *flags, int64_t fper,
                                                              Plane out, bool out_al16, BitPlane obits, bool or_bits)
{
    ROW_PROLOGUE(0)
    if (!warp_ok) return;
    const int32_t *L = Lall + img * g.bper;
    const uint32_t *F = flags + img * fper;
    const int cur_base = by * g.bw;
    const int y = 2 * by;
    const bool has_row1 = y + 1 < g.height;
    RunCarry c{0u, -1};
    for (int s = 0; s < g.nseg; ++s) {
        const int chunk = s * G + lane;
        const u64 a0 = load_chunk(r0, chunk, g.width), a1 = load_chunk(r1, chunk, g.width);
        const Runs r = analyze_runs<G>(a0 | a1, chunk, lane, c);
        u64 keep = 0, rem = r.occE;
        while (rem) {
            int fb;
            const u64 range = next_piece(r, chunk, rem, fb);
            const int32_t root = L[cur_base + fb];          // fully compressed by rccl_compress_kernel<1>
            if ((F[root >> 5] >> (root & 31)) & 1u) keep |= range;
        }
        if (!row_ok || chunk >= g.cpr) continue;
        const u64 o0 = a0 & keep, o1 = a1 & keep;
        if (OUT_BITS) {
            uint2 *p0 = (uint2 *)(obits.p + img * obits.bs + (int64_t)y * obits.wpr) + chunk;
            uint2 w0 = make_uint2((uint32_t)o0, (uint32_t)(o0 >> 32));
            if (or_bits) { const uint2 e = *p0; w0.x |= e.x; w0.y |= e.y; }
            *p0 = w0;
            if (has_row1) {
                uint2 *p1 = (uint2 *)((uint32_t *)p0 + obits.wpr);
                uint2 w1 = make_uint2((uint32_t)o1, (uint32_t)(o1 >> 32));
                if (or_bits) { const uint2 e = *p1; w1.x |= e.x; w1.y |= e.y; }
                *p1 = w1;
            }
        } else {
            const int x0 = 64 * chunk;
            for (int rr = 0; rr < (has_row1 ? 2 : 1); ++rr) {
                const u64 o = rr ? o1 : o0;
                uint8_t *orow = out.p + img * out.bs + (int64_t)(y + rr) * out.rs + x0;
                if (out_al16 && x0 + 64 <= g.width) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t h = (uint32_t)(o >> (16 * q)) & 0xFFFFu;
                        ((uint4 *)orow)[q] = make_uint4(bytes_of_nibble(h & 15u), bytes_of_nibble((h >> 4) & 15u),
                                                        bytes_of_nibble((h >> 8) & 15u), bytes_of_nibble(h >> 12));
                    }
                } else {
                    const int n = min(64, g.width - x0);
                    for (int j = 0; j < n; ++j) orow[j] = ((o >> j) & 1ULL) ? 255 : 0;
                }
            }
        }
    }
}

// lanes per block row that waste the fewest lanes on this width (ties -> 32)
int pick_group(int width)
{
    const int chunks = cdiv(width, 64);
    return (cdiv(chunks, 16) * 16 < cdiv(chunks, 32) * 32) ? 16 : 32;
}

RowGeom geom_of(BitPlane bits, int width, int height, int G)
{
    RowGeom g;
    g.bits = bits.p; g.wpr = bits.wpr; g.wbs = bits.bs;
    g.width = width; g.height = height;
    g.bw = (width + 1) / 2; g.bh = (height + 1) / 2;
    g.bper = (int64_t)align_up((size_t)g.bw * g.bh, 4);
    g.nseg = cdiv(width, 64 * G);
    g.cpr = cdiv(width, 64);
    g.dims = bits.dims;
    g.need = nullptr; g.need_stride = 0;
    return g;
}

inline dim3 row_grid(const RowGeom &g, int batch, int first_row, int G) { return dim3(cdiv(g.bh - first_row, ROWS_PER_CTA * (32 / G)), batch); }

// launch `kernel<..., G>` for G = 16 or 32
#define LAUNCH_G(G_, kernel16, kernel32, grid, smem, st, ...)                    \
    do {                                                                         \
        if ((G_) == 16) kernel16<<<grid, 256, smem, st>>>(__VA_ARGS__);          \
        else kernel32<<<grid, 256, smem, st>>>(__VA_ARGS__);                     \
    } while (0)

// init + merge: after this every run start's parent chain ends at the root of its component
int run_union_find(synseg_ctx *ctx, const RowGeom &g, int G, int batch, int32_t *L, const BitPlane *strong, cudaStream_t st)
{
    LAUNCH_G(G, rccl_init_kernel<16>, rccl_init_kernel<32>, row_grid(g, batch, 0, G), 0, st, g, L);
    SS_LAUNCH_CHECK(ctx, "ccl_init", st);
    if (g.bh > 1) {
        // The union kernel works on whole-warp block rows (G = 32), two consecutive rows per warp one after the other (rccl_merge2_kernel).
        // Half-warp rows -- two rows of a warp side by side, as in the purely analytic kernels -- were slower here: divergent event loops,
        // colliding atomics of adjacent rows (0.23 -> 0.25 ms in round 1, 0.119 against 0.109 ms now).  -DSYNSEG_MERGE1 selects the older
        // kernel with one block row per warp (with -DSYNSEG_MERGE_G16 also its half-warp form for the labelling).
        RowGeom g32 = g;
        g32.nseg = cdiv(g.width, 64 * 32);
#ifndef SYNSEG_MERGE1
        const dim3 grid2(cdiv(cdiv(g32.bh - 1, 2), ROWS_PER_CTA), batch);
        if (strong) rccl_merge2_kernel<true><<<grid2, 256, 0, st>>>(g32, L, *strong);
        else rccl_merge2_kernel<false><<<grid2, 256, 0, st>>>(g32, L, BitPlane{nullptr, 0, 0});
#else
        const dim3 grid = row_grid(g32, batch, 1, 32);
        if (strong) rccl_merge_kernel<true, 32><<<grid, 256, 0, st>>>(g32, L, *strong);
#ifdef SYNSEG_MERGE_G16
        else if (G == 16) rccl_merge_kernel<false, 16><<<row_grid(g, batch, 1, 16), 256, 0, st>>>(g, L, BitPlane{nullptr, 0, 0});
#endif
        else rccl_merge_kernel<false, 32><<<grid, 256, 0, st>>>(g32, L, BitPlane{nullptr, 0, 0});
#endif
        SS_LAUNCH_CHECK(ctx, "ccl_merge", st);
    }
    return SYNSEG_OK;
}

}  // namespace

size_t ccl_label_scratch_bytes(int width, int height, int batch)
{
    const int64_t bw = (width + 1) / 2, bh = (height + 1) / 2;
    return align_up((size_t)(bw * bh), 4) * batch * sizeof(int32_t) + 256;
}

size_t hysteresis_scratch_bytes(int width, int height, int batch)
{
    const int64_t bw = (width + 1) / 2, bh = (height + 1) / 2;
    return ccl_label_scratch_bytes(width, height, batch) + (size_t)cdiv(bw * bh, 32) * 4 * batch + 32 * (size_t)batch + 1024;
}

size_t ccl_stats_scratch_bytes(int width, int height, int batch, int max_labels)
{
    const int64_t bh = (height + 1) / 2;
    const size_t plane = (size_t)bit_wpr(width) * height * batch * 4 + 256;      // packed copy of a u8 mask
    return ccl_label_scratch_bytes(width, height, batch) + plane + (size_t)bh * cdiv(width, 64) * 8 * batch + (size_t)bh * batch * 4 +
           (size_t)batch * max_labels * 36 + (size_t)batch * 4 + 16 * 256;
}

int run_ccl_stats(synseg_ctx *ctx, const CclMask &m, const synseg_img *labels, int32_t *n_labels, int32_t *stats,
                  double *centroids, int32_t max_labels, cudaStream_t st)
{
    const int batch = m.batch;
    void *p;
    BitPlane bits = m.bits;
    if (m.u8) {
        const int wpr = bit_wpr(m.width);
        SS_TRY(arena_alloc(ctx, (size_t)wpr * m.height * batch * 4, &p, st));
        bits = BitPlane{(uint32_t *)p, wpr, (int64_t)wpr * m.height};
        SS_TRY(launch_pack_bits(ctx, m.u8, bits, st));
    }
    const int G = labels ? 32 : pick_group(m.width);          // the label-image path stages one block row per warp
    const RowGeom g = geom_of(bits, m.width, m.height, G);
    SS_TRY(arena_alloc(ctx, (size_t)g.bper * batch * 4, &p, st));
    int32_t *L = (int32_t *)p;
    SS_TRY(arena_alloc(ctx, (size_t)g.bh * g.cpr * 8 * batch, &p, st));
    u64 *rootbits = (u64 *)p;
    SS_TRY(arena_alloc(ctx, (size_t)g.bh * batch * 4, &p, st));
    int32_t *row_count = (int32_t *)p;
    SS_TRY(arena_alloc(ctx, (size_t)batch * 4, &p, st));
    int32_t *n_roots = (int32_t *)p;
    StatAcc a;
    a.cap = max_labels;
    const size_t nacc = (size_t)batch * max_labels;
    SS_TRY(arena_alloc(ctx, nacc * 4, &p, st)); a.minx = (int32_t *)p;
    SS_TRY(arena_alloc(ctx, nacc * 4, &p, st)); a.miny = (int32_t *)p;
    SS_TRY(arena_alloc(ctx, nacc * 4, &p, st)); a.maxx = (int32_t *)p;
    SS_TRY(arena_alloc(ctx, nacc * 4, &p, st)); a.maxy = (int32_t *)p;
    SS_TRY(arena_alloc(ctx, nacc * 4, &p, st)); a.area = (int32_t *)p;
    SS_TRY(arena_alloc(ctx, nacc * 8, &p, st)); a.sumx = (u64 *)p;
    SS_TRY(arena_alloc(ctx, nacc * 8, &p, st)); a.sumy = (u64 *)p;

    SS_TRY(run_union_find(ctx, g, G, batch, L, nullptr, st));
    const dim3 grid = row_grid(g, batch, 0, G);
    LAUNCH_G(G, (rccl_compress_kernel<0, 16>), (rccl_compress_kernel<0, 32>), grid, 0, st, g, L, rootbits, row_count, BitPlane{nullptr, 0, 0},
             (uint32_t *)nullptr, (int64_t)0);
    SS_LAUNCH_CHECK(ctx, "ccl_compress", st);
    ccl_scan_kernel<<<batch, 256, 0, st>>>(g.bh, row_count, n_roots);
    SS_LAUNCH_CHECK(ctx, "ccl_scan", st);
    LAUNCH_G(G, rccl_assign_kernel<16>, rccl_assign_kernel<32>, grid, 0, st, g, L, rootbits, row_count);
    SS_LAUNCH_CHECK(ctx, "ccl_assign", st);
    stats_init_kernel<<<(unsigned)cdiv(nacc, 256), 256, 0, st>>>(a, (int64_t)nacc);
    SS_LAUNCH_CHECK(ctx, "stats_init", st);
    if (labels) {
        const bool al16 = plane_aligned(labels, 16);
        rccl_final_kernel<true, 32><<<grid, 256, ROWS_PER_CTA * sizeof(LabelStage), st>>>(g, L, plane_of(labels), al16, a);
    } else {
        LAUNCH_G(G, (rccl_final_kernel<false, 16>), (rccl_final_kernel<false, 32>), grid, 0, st, g, L, Plane{nullptr, 0, 0}, false, a);
    }
    SS_LAUNCH_CHECK(ctx, "ccl_final", st);
    stats_finalize_kernel<<<dim3(cdiv(max_labels, 128), batch), 128, 0, st>>>(a, n_roots, batch, n_labels, stats, centroids, m.width, m.height, g.dims);
    SS_LAUNCH_CHECK(ctx, "stats_finalize", st);
    return SYNSEG_OK;
}

int run_hysteresis(synseg_ctx *ctx, BitPlane kept, BitPlane strong, int width, int height, int batch, const synseg_img *edges_u8,
                   BitPlane edges_bits, bool or_bits, cudaStream_t st)
{
    const int G = pick_group(width);
    RowGeom g = geom_of(kept, width, height, G);
    void *p;
    SS_TRY(arena_alloc(ctx, (size_t)g.bper * batch * 4, &p, st));
    int32_t *L = (int32_t *)p;
    const int64_t fper = cdiv((int64_t)g.bw * g.bh, 32);
    SS_TRY(arena_alloc(ctx, (size_t)fper * 4 * batch, &p, st));
    uint32_t *flags = (uint32_t *)p;
    // Propagation sweeps first (hyst_sweep.cu): on print nearly every kept pixel is strong or touches a strong one and a few
    // bit-parallel sweeps reach the fixed point; the union-find below then finds *need == 0 and its kernels return at once.
#ifndef SYNSEG_HS_SWEEPS
#define SYNSEG_HS_SWEEPS 4
#endif
    constexpr int N_SWEEPS = SYNSEG_HS_SWEEPS;
    if (!edges_u8 && !getenv("SYNSEG_NO_HYST_SWEEPS")) {
        SS_TRY(arena_alloc(ctx, sizeof(int32_t) * (N_SWEEPS + 1) * (size_t)batch, &p, st));
        int32_t *sw = (int32_t *)p;
        bool used = false;
        SS_TRY(launch_hyst_sweeps(ctx, kept, strong, edges_bits, or_bits, width, height, batch, sw, N_SWEEPS, &used, st));
        if (used) { g.need = sw + (N_SWEEPS - 1); g.need_stride = N_SWEEPS + 1; }
    }
    SS_CUDA(cudaMemsetAsync(flags, 0, (size_t)fper * 4 * batch, st));
    SS_TRY(run_union_find(ctx, g, G, batch, L, &strong, st));
    const dim3 grid = row_grid(g, batch, 0, G);
    LAUNCH_G(G, (rccl_compress_kernel<1, 16>), (rccl_compress_kernel<1, 32>), grid, 0, st, g, L, (u64 *)nullptr, (int32_t *)nullptr, strong, flags, fper);
    SS_LAUNCH_CHECK(ctx, "hyst_flag", st);
    if (edges_u8)
        LAUNCH_G(G, (rccl_hyst_final_kernel<false, 16>), (rccl_hyst_final_kernel<false, 32>), grid, 0, st, g, L, flags, fper, plane_of(edges_u8),
                 plane_aligned(edges_u8, 16), BitPlane{nullptr, 0, 0}, false);
    else
        LAUNCH_G(G, (rccl_hyst_final_kernel<true, 16>), (rccl_hyst_final_kernel<true, 32>), grid, 0, st, g, L, flags, fper, Plane{nullptr, 0, 0}, false,
                 edges_bits, or_bits);
    SS_LAUNCH_CHECK(ctx, "hyst_final", st);
    return SYNSEG_OK;
}

extern "C" SYNSEG_EXPORT int synseg_ccl_stats(synseg_ctx *ctx, const synseg_img *mask, const synseg_img *labels, int32_t *n_labels,
                                int32_t *stats, double *centroids, int32_t max_labels, void *stream)
{
    if (!ctx) { synseg_set_error("synseg_ccl_stats: ctx is NULL"); return SYNSEG_E_INVALID; }
    SS_ENTER(ctx, stream);
    SS_TRY(validate_img(mask, "mask", 1));
    if (labels) {
        SS_TRY(validate_img(labels, "labels", 4));
        if (!same_shape(mask, labels)) { synseg_set_error("synseg_ccl_stats: labels shape mismatch"); return SYNSEG_E_INVALID; }
        if (((uintptr_t)labels->data | (uintptr_t)labels->row_stride | (uintptr_t)labels->batch_stride) & 3) {
            synseg_set_error("synseg_ccl_stats: labels must be 4-byte aligned"); return SYNSEG_E_INVALID;
        }
    }
    if (!n_labels || !stats || !centroids || max_labels < 1) { synseg_set_error("synseg_ccl_stats: bad result buffers"); return SYNSEG_E_INVALID; }
    if (mask->width > 32766 || mask->height > 32766) { synseg_set_error("synseg_ccl_stats: image larger than 32766"); return SYNSEG_E_INVALID; }
    if (mask->batch > 65535) { synseg_set_error("synseg_ccl_stats: batch > 65535"); return SYNSEG_E_INVALID; }
    SS_TRY(arena_ensure(ctx, ccl_stats_scratch_bytes(mask->width, mask->height, mask->batch, max_labels)));
    arena_begin(ctx);
    CclMask m; m.u8 = mask; m.bits = BitPlane{nullptr, 0, 0}; m.width = mask->width; m.height = mask->height; m.batch = mask->batch;
    return run_ccl_stats(ctx, m, labels, n_labels, stats, centroids, max_labels, (cudaStream_t)stream);
}
