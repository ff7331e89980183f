// canny.cu -- cv2.Canny(g, lo, hi) (aperture 3, L1 gradient), bit-exact.
//
// Reference call sites: pdf_image_segmentation.py:1324,1366,1550,1600,1700,1759 (all Canny(.,50,150)).
// Algorithm (OpenCV 4.13 semantics):
//   Sobel 3x3 on the replicate-bordered image -> mag = |dx|+|dy| with a ring of zeros outside the
//   image -> for mag > lo, integer non-maximum suppression along the quantised gradient direction
//   (TG22 = 13573 = tan(22.5 deg) << 15) -> kept pixels with mag > hi are strong -> the output is
//   every kept pixel 8-connected (through kept pixels) to a strong one.
// Stage 1 (this file): warp-autonomous strips, no block-level barrier.  A warp owns a strip of 512 columns
//   (16 per lane, one 128-bit load per lane and row; lanes 0 and 31 are halo lanes, 480 output columns) and
//   marches down a band of rows.  All arithmetic is packed 16-bit SIMD (two pixels per register; VIADD.16x2 /
//   VIADDMNMX.S16x2 are native on sm_100a): per grey row the horizontal Sobel partials s = [1 2 1] and
//   d = [-1 0 1] are formed once from the even / odd byte lanes and kept for three rows in registers;
//   dx = dA + 2 dB + dC, dy = sC - sA, mag = |dx| + |dy|.  mag / dx / dy go to a three-row, lane-transposed
//   (conflict-free) shared-memory ring; the integer non-maximum suppression of the previous row runs for the
//   candidate pixels (mag > lo) only, reading its neighbours from the ring: a strip row full of candidates first has its steep
//   vertical-gradient ones decided for whole lanes at once in packed arithmetic (nms_vertical_packed), the rest are listed and
//   dealt out evenly to the lanes.  Warp-uniform shortcuts: a row whose 3x18 neighbourhood is constant in every lane (blank
//   paper) skips the gradient and the ring stores; a row equal to the five before it repeats the previous output row.
//   The result is two bit planes: kept = survived NMS, strong = kept and mag > hi (two lanes -> one word).
// Stage 2: hysteresis = bit-parallel propagation sweeps (hyst_sweep.cu), then, for the images they do not settle, the run-based
//   union-find over the kept pixels + "component holds a strong pixel" flag (ccl.cu); no host round trip.
//
// Roofline: HBM-bound, 1.25 algorithmic bytes per pixel (1 read + two bit planes written) for stage 1.
#include <cuda.h>
#include <stdlib.h>

#include "internal.cuh"
#include "pixel.cuh"

namespace {

constexpr int CN_OUT_W = 480;           // output columns per warp strip (lanes 1..30 x 16)
#ifndef SYNSEG_CN_WARPS
#define SYNSEG_CN_WARPS 1
#endif
constexpr int CN_WARPS = SYNSEG_CN_WARPS;   // independent warps per CTA (1: a finished warp frees its registers at once)
#ifndef SYNSEG_CN_MINBLOCKS
#define SYNSEG_CN_MINBLOCKS (19 / SYNSEG_CN_WARPS)   // resident warps per SM the register allocation aims at (17..20 measured equal, 21 slower)
#endif
constexpr unsigned FULL = 0xffffffffu;
constexpr int CN_DEPTH = 4;             // grey rows in flight per warp (cp.async ring)

struct CnParams {
    Plane src;
    BitPlane kept, strong;
    int width, height, lo, hi;
    int strips, bands, band_h;
    int64_t tasks;
};

// Horizontal partials of one grey row for this lane's 16 columns, two columns per register:
// register 2k holds columns 4k (low half) and 4k+2 (high half), register 2k+1 columns 4k+1 and 4k+3.
struct HRow {
    uint32_t s[8];   // g[x-1] + 2 g[x] + g[x+1]
    uint32_t d[8];   // g[x+1] - g[x-1] + 256
    uint32_t rep;    // first pixel replicated into four bytes
    bool uni;        // the 16 columns and the two neighbouring words all hold that pixel value
};

__device__ __forceinline__ void make_hrow(HRow &h, const uint4 v, uint32_t wl, uint32_t wr)
{
    const uint32_t rep = __byte_perm(v.x, 0, 0x0000);
    h.rep = rep;
    h.uni = (((v.x ^ rep) | (v.y ^ rep) | (v.z ^ rep)) | ((v.w ^ rep) | (wl ^ rep) | (wr ^ rep))) == 0u;
    if (h.uni) {
        const uint32_t s4 = (v.x & 0xFFu) * 0x00040004u;
#pragma unroll
        for (int i = 0; i < 8; ++i) { h.s[i] = s4; h.d[i] = 0x01000100u; }
    } else {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t E[5], O[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { E[k] = w[k] & 0x00FF00FFu; O[k] = __byte_perm(w[k], 0, 0x4341); }
        E[4] = wr & 0x00FF00FFu;
        const uint32_t Om1 = __byte_perm(wl, 0, 0x4341);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t Op = __funnelshift_r(k ? O[k - 1] : Om1, O[k], 16);    // columns 4k-1, 4k+1
            const uint32_t En = __funnelshift_r(E[k], E[k + 1], 16);              // columns 4k+2, 4k+4
            h.s[2 * k] = Op + O[k] + 2u * E[k];
            h.d[2 * k] = O[k] + 0x01000100u - Op;
            h.s[2 * k + 1] = E[k] + En + 2u * O[k];
            h.d[2 * k + 1] = En + 0x01000100u - E[k];
        }
    }
}

// magnitudes of three rows (NMS looks one row up and down); dx / dy only of the row being suppressed and the row being
// produced (two slots, row parity)
struct MagRing { uint32_t mag[3][8][32], dx[2][8][32], dy[2][8][32]; };   // [ring row][pair register][lane]

__device__ __forceinline__ int col_pr(int c) { return 2 * (c >> 2) + (c & 1); }
__device__ __forceinline__ int col_half(int c) { return (c >> 1) & 1; }

// magnitude of column c (-1..16) of `lane` in ring row `slot` (rows flagged in zmask are all zero)
__device__ __forceinline__ int ring_mag(const MagRing &R, uint32_t zmask, int slot, int lane, int c)
{
    if ((zmask >> slot) & 1u) return 0;
    if (c < 0) { c = 15; --lane; } else if (c > 15) { c = 0; ++lane; }
    return (int)((R.mag[slot][col_pr(c)][lane] >> (16 * col_half(c))) & 0xFFFFu);
}

struct CnState {
    uint32_t zmask;      // ring rows known to be all zero (warp-uniform)
    uint32_t kpair;      // (0x7FFF - lo) in both halves: mag + k has its sign bit set iff mag > lo
    uint32_t cv16;       // bit c: column c of this lane lies inside the image
};

// Magnitude row (middle row B) -> ring slot.  Returns the candidate mask in gather order:
// bit 4t + i <-> pair register 2i + (t >> 1), half t & 1  <->  column 4i + (t >> 1) + 2 (t & 1).
__device__ __forceinline__ uint32_t produce_row(MagRing &R, CnState &st, int slot, int par, int lane, const HRow &A, const HRow &B, const HRow &C,
                                                bool row_in)
{
    const bool blank = A.uni & B.uni & C.uni & (A.rep == B.rep) & (B.rep == C.rep);
    if (!row_in || __all_sync(FULL, blank)) { st.zmask |= 1u << slot; return 0u; }
    st.zmask &= ~(1u << slot);
    uint32_t cm[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint32_t dxb = A.d[i] + C.d[i] + 2u * B.d[i];                    // dx + 1024 per half
        const uint32_t dx = __vadd2(dxb, 0xFC00FC00u);
        const uint32_t dy = __vadd2(C.s[i] + 0x04000400u - A.s[i], 0xFC00FC00u);
        uint32_t mag = __vadd2(__vmaxs2(dx, __vneg2(dx)), __vmaxs2(dy, __vneg2(dy)));
        if (st.cv16 != 0xFFFFu) {                                              // lanes touching the image border
            const int c0 = 4 * (i >> 1) + (i & 1);
            mag &= (((st.cv16 >> c0) & 1u) ? 0xFFFFu : 0u) | (((st.cv16 >> (c0 + 2)) & 1u) ? 0xFFFF0000u : 0u);
        }
        R.mag[slot][i][lane] = mag; R.dx[par][i][lane] = dx; R.dy[par][i][lane] = dy;
        cm[i] = __vadd2(mag, st.kpair);
    }
    uint32_t z = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) z |= ((__byte_perm(cm[2 * i], cm[2 * i + 1], 0x7531) & 0x80808080u) >> 7) << i;
    z = (z | (z >> 4)) & 0x00FF00FFu;
    return (z | (z >> 8)) & 0xFFFFu;
}

// Per-warp staging of the balanced NMS: the candidates of a strip row are listed and dealt out evenly to the lanes.
struct NmsStage { uint16_t list[512]; uint32_t kept[32], strong[32]; };

// non-maximum suppression of candidate b (gather order) of lane `owner` in ring row y; results are OR-ed into the stage
__device__ __forceinline__ void nms_one(const MagRing &R, const CnState &st, NmsStage &S, int y, int owner, int b, int hi)
{
    const int sc = y % 3, su = (y + 2) % 3, sd = (y + 1) % 3;
    const int i = b & 3, t = b >> 2;
    const int pr = 2 * i + (t >> 1), sh = 16 * (t & 1);
    const int col = 4 * i + (t >> 1) + 2 * (t & 1);
    const int m = (int)((R.mag[sc][pr][owner] >> sh) & 0xFFFFu);
    const int dx = (int)(short)(R.dx[y & 1][pr][owner] >> sh), dy = (int)(short)(R.dy[y & 1][pr][owner] >> sh);
    const int ax = abs(dx), ay = abs(dy) << 15, tg22 = ax * 13573;
    bool keep;
    if (ay < tg22) keep = m > ring_mag(R, st.zmask, sc, owner, col - 1) && m >= ring_mag(R, st.zmask, sc, owner, col + 1);
    else {
        const int tg67 = tg22 + (ax << 16);
        if (ay > tg67) keep = m > ring_mag(R, st.zmask, su, owner, col) && m >= ring_mag(R, st.zmask, sd, owner, col);
        else {
            const int s = ((dx ^ dy) < 0) ? -1 : 1;
            keep = m > ring_mag(R, st.zmask, su, owner, col - s) && m > ring_mag(R, st.zmask, sd, owner, col + s);
        }
    }
    if (keep) {
        atomicOr(&S.kept[owner], 1u << col);
        if (m > hi) atomicOr(&S.strong[owner], 1u << col);
    }
}

// sign bits of eight packed 16-bit pair registers -> 16-bit mask in gather order (see produce_row)
__device__ __forceinline__ uint32_t sign_bits_gather(const uint32_t w[8])
{
    uint32_t z = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) z |= ((__byte_perm(w[2 * i], w[2 * i + 1], 0x7531) & 0x80808080u) >> 7) << i;
    z = (z | (z >> 4)) & 0x00FF00FFu;
    return (z | (z >> 8)) & 0xFFFFu;
}

// gather order -> column order, both 16-bit halves of x at once: bit 4t + i -> bit 4i + (t >> 1) + 2 (t & 1), i.e. the index bits
// (b3 b2 b1 b0) become (b1 b0 b2 b3): three delta swaps
__device__ __forceinline__ uint32_t gather_to_columns(uint32_t x)
{
    uint32_t t;
    t = ((x >> 6) ^ x) & 0x00CC00CCu; x ^= t ^ (t << 6);
    t = ((x >> 3) ^ x) & 0x0A0A0A0Au; x ^= t ^ (t << 3);
    t = ((x >> 1) ^ x) & 0x22222222u; x ^= t ^ (t << 1);
    return x;
}

#ifndef SYNSEG_CN_VGATE
#define SYNSEG_CN_VGATE 160
#endif
constexpr int CN_VGATE = SYNSEG_CN_VGATE;   // candidates in a strip row from which the packed pass below runs first (0 = never)

// Candidates whose gradient is steep enough to be in cv2's vertical class without the 32-bit products -- |dy| >= 3 |dx| implies
// |dy| << 15 > |dx| (13573 + 65536) -- are suppressed for all 16 pixels of a lane at once in packed 16-bit arithmetic: their two
// neighbours are the pixels above and below, which sit in the same register position of the other ring rows (no list, no dealing
// out, no per-candidate address arithmetic).  The tops and bottoms of words, rules, bars and frames put hundreds of such candidates
// into one strip row (9 rounds of the list path, ~80 instructions each, against ~230 here).  Returns kept | strong << 16 in column
// order and removes the candidates it decided from `cand` (gather order).  Every magnitude is < 2^15, so the sign of a packed
// difference is the comparison.
__device__ __forceinline__ uint32_t nms_vertical_packed(const MagRing &R, const CnState &st, int y, int lane, int hi, uint32_t &cand)
{
    const int sc = y % 3, su = (y + 2) % 3, sd = (y + 1) % 3, par = y & 1;
    const bool zu = (st.zmask >> su) & 1u, zd = (st.zmask >> sd) & 1u;
    const uint32_t hpair = (uint32_t)min(hi, 0x7FFF) * 0x00010001u;
    uint32_t vs[8], ks[8], ss[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint32_t m = R.mag[sc][i][lane], up = zu ? 0u : R.mag[su][i][lane], dn = zd ? 0u : R.mag[sd][i][lane];
        const uint32_t dx = R.dx[par][i][lane], dy = R.dy[par][i][lane];
        const uint32_t ax = __vmaxs2(dx, __vneg2(dx)), ay = __vmaxs2(dy, __vneg2(dy));
        const uint32_t t = __vsub2(ay, __vadd2(ax, ax << 1));          // sign: |dy| < 3 |dx| (not decided here)
        const uint32_t c = __vadd2(m, st.kpair);                       // sign: candidate (m > lo)
        const uint32_t d1 = __vsub2(up, m), d2 = __vsub2(m, dn);       // sign: m > up;  m < dn
        const uint32_t h = __vsub2(hpair, m);                          // sign: m > hi
        vs[i] = c & ~t;
        ks[i] = vs[i] & d1 & ~d2;
        ss[i] = ks[i] & h;
    }
    const uint32_t vmask = sign_bits_gather(vs) & cand;                // (halo lanes and lanes outside the image hold cand == 0)
    const uint32_t res = gather_to_columns((sign_bits_gather(ks) & vmask) | ((sign_bits_gather(ss) & vmask) << 16));
    cand &= ~vmask;
    return res;
}

// Non-maximum suppression of strip row y: candidate mask of this lane (gather order) -> kept16 / strong16 (column order).
__device__ __forceinline__ void nms_row(MagRing &R, const CnState &st, NmsStage &S, int y, int lane, uint32_t mycand, int hi,
                                        uint32_t &kept16, uint32_t &strong16)
{
    kept16 = 0; strong16 = 0;
    int cnt = __popc(mycand);
    const int total0 = __reduce_add_sync(FULL, cnt);
    if (total0 == 0) return;
    if (CN_VGATE > 0 && total0 >= CN_VGATE) {
        const uint32_t r = nms_vertical_packed(R, st, y, lane, hi, mycand);
        kept16 = r & 0xFFFFu; strong16 = r >> 16;
        cnt = __popc(mycand);
        if (!__any_sync(FULL, cnt != 0)) return;
    }
    // balanced NMS: list the candidates of the strip row, deal them out evenly to the 32 lanes
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int nb = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += nb; }
    const int total = __shfl_sync(FULL, incl, 31);
    int pos = incl - cnt;
    for (uint32_t c = mycand; c; c &= c - 1) S.list[pos++] = (uint16_t)((lane << 4) | (__ffs((int)c) - 1));
    S.kept[lane] = 0u; S.strong[lane] = 0u;
    __syncwarp();
    for (int q = lane; q < total; q += 32) { const int id = S.list[q]; nms_one(R, st, S, y, id >> 4, id & 15, hi); }
    __syncwarp();
    kept16 |= S.kept[lane]; strong16 |= S.strong[lane];
}

// ALIGNED (source rows 16-byte aligned) is a template parameter: the byte-load fallback stays out of the hot code.
template <bool ALIGNED>
__global__ void __launch_bounds__(32 * CN_WARPS, SYNSEG_CN_MINBLOCKS) canny_classes_kernel(CnParams p)
{
    constexpr bool aligned = ALIGNED;
    __shared__ MagRing rings[CN_WARPS];
    __shared__ NmsStage stages[CN_WARPS];
    __shared__ uint4 Rows[CN_WARPS][CN_DEPTH][32];   // cp.async ring of grey rows: [step][lane]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t task = (int64_t)blockIdx.x * CN_WARPS + warp;
    if (task >= p.tasks) return;                      // warp-uniform
    const int strip = (int)(task % p.strips); task /= p.strips;
    const int band = (int)(task % p.bands);
    const int img = (int)(task / p.bands);
    MagRing &R = rings[warp];
    NmsStage &S = stages[warp];

    int W = p.width, H = p.height;
    const int hi = p.hi;
    if (p.kept.dims) {                                // ragged batch: the grid covers the canvas, this image may be smaller
        const int2 d = p.kept.dims[img];
        W = d.x; H = d.y;
        if (strip * CN_OUT_W >= W || band * p.band_h >= H) return;   // warp-uniform (one task per warp)
    }
    const int x = strip * CN_OUT_W - 16 + 16 * lane;  // first of this lane's 16 columns
    const int y0 = band * p.band_h, y1 = min(y0 + p.band_h, H);
    const uint8_t *base = p.src.p + img * p.src.bs;
    const int64_t rs = p.src.rs;
    const bool out_lane = lane >= 1 && lane <= 30 && x < W;
    const bool writer = out_lane && (lane & 1) && lane <= 29;

    CnState st;
    st.zmask = 0;
    const int lo = p.lo < 0 ? 0 : (p.lo > 4095 ? 4095 : p.lo);
    st.kpair = (uint32_t)(0x7FFF - lo) * 0x00010001u;
    st.cv16 = 0;
#pragma unroll
    for (int c = 0; c < 16; ++c) st.cv16 |= (x + c >= 0 && x + c < W) ? (1u << c) : 0u;

    // Aligned rows: one 16-byte unit per lane and row at a clamped address; the replicated border is produced in
    // registers when the row is consumed (fix16_rep), so loads never wait on data and run CN_DEPTH rows ahead through
    // a cp.async ring in shared memory (no registers in flight; each lane reads back only its own copy).
    const int xl = clamp16_x(x, W);
    const EdgeFix efix = make_edge_fix(x, W);
    uint4 *ring = &Rows[warp][0][lane];
    auto row_ptr = [&](int yy) { return base + (int64_t)min(max(yy, 0), H - 1) * rs; };
    auto load_now = [&](int yy) -> uint4 {
        return aligned ? apply_edge_fix(__ldg((const uint4 *)(row_ptr(yy) + xl)), efix) : load16_rep(row_ptr(yy), x, W, false);
    };
    // running pointers (no 64-bit multiplies in the row loop): the row being prefetched and the two output rows
    int y_pf = y0;                                                  // next row to prefetch (clamped into the image)
    const uint8_t *p_pf = row_ptr(y_pf) + xl;
    auto issue_async = [&](int slot) {
        if (aligned) cp_async16(ring + slot * 32, p_pf);
        cp_async_commit();
        ++y_pf;
        if (y_pf >= 1 && y_pf <= H - 1) p_pf += rs;                 // the pointer only moves while the row index is inside [1, H-1]
    };
    uint32_t *kp = p.kept.p + img * p.kept.bs + (int64_t)y0 * p.kept.wpr + (max(x, 0) >> 5);
    uint32_t *sp = p.strong.p + img * p.strong.bs + (int64_t)y0 * p.strong.wpr + (max(x, 0) >> 5);
    const int k_wpr = p.kept.wpr, s_wpr = p.strong.wpr;
    auto build_hrow = [&](HRow &h, const uint4 v) {
        const uint32_t wl = __shfl_up_sync(FULL, v.w, 1), wr = __shfl_down_sync(FULL, v.x, 1);
        make_hrow(h, v, wl, wr);
    };

    // One loop does everything, including the two priming steps (y = y0-2, y0-1 produce the magnitude rows y0-1, y0
    // without emitting output): a single copy of the row step keeps the loop body inside the instruction cache
    // (three unrolled copies were 3800 SASS instructions and 18 % slower).
    // A = partials of grey row y, B = row y+1, C = free (receives row y+2); rotated by register moves.
    HRow A, B, C;
    build_hrow(A, load_now(y0 - 2));
    build_hrow(B, load_now(y0 - 1));
#pragma unroll
    for (int d = 0; d < CN_DEPTH; ++d) issue_async(d);                           // rows y0 .. y0+CN_DEPTH-1 in flight
    int slot = 0;
    uint32_t cand_cur = 0;
    // Repeated rows (REPEAT_NOTE below): rep = number of consecutive grey rows, ending with the newest, that equal their predecessor in
    // every lane of the strip; last_kw / last_sw = the words written for the previous output row.
    int rep = 0;
    uint4 vprev = load_now(y0 - 1);
    uint32_t last_kw = 0, last_sw = 0;
#pragma unroll 1
    for (int y = y0 - 2; y < y1; ++y) {
        uint4 vcur;
        if (aligned) {
            cp_async_wait<CN_DEPTH - 1>();
            vcur = apply_edge_fix(ring[slot * 32], efix);
            issue_async(slot);
            slot = (slot + 1) & (CN_DEPTH - 1);
        } else vcur = load16_rep(row_ptr(y + 2), x, W, false);
        {
            const bool same = ((vcur.x ^ vprev.x) | (vcur.y ^ vprev.y) | (vcur.z ^ vprev.z) | (vcur.w ^ vprev.w)) == 0u;
            rep = __all_sync(FULL, same) ? rep + 1 : 0;
            vprev = vcur;
            // REPEAT_NOTE.  Output row y is a function of the grey rows y-2 .. y+2.  If the rows y-3 .. y+2 are all equal (rep >= 5) and all of
            // them lie inside the image (no replicated or zero border row among them), output row y equals output row y-1, the
            // magnitude row y+1 equals the three rows in the ring, its candidates equal the current ones and the partials rotate
            // into themselves: nothing has to be computed -- the previous words are written again.  Vertical strokes, bars, rules, the
            // inside of solid boxes: half of the rows of a page of print.
            if (rep >= 5 && y - 3 >= 0 && y + 2 <= H - 1 && y >= y0 + 1) {
                if (writer) { *kp = last_kw; *sp = last_sw; }
                kp += k_wpr; sp += s_wpr;
                continue;
            }
        }
        const uint32_t wl = __shfl_up_sync(FULL, vcur.w, 1), wr = __shfl_down_sync(FULL, vcur.x, 1);
        {
            // Blank paper (45 % of the rows of a text page): the new row is the same constant as the two rows above it in
            // every lane and the row being suppressed has no candidate.  Its partials would equal A's and B's, so the
            // rotation is the identity and the magnitude row is zero: mark the ring slot, store zeros, next row
            // (~45 instructions instead of ~160 through the general path's shortcuts).
            const uint32_t rep = __byte_perm(vcur.x, 0, 0x0000);
            const bool uni = (((vcur.x ^ rep) | (vcur.y ^ rep) | (vcur.z ^ rep)) | ((vcur.w ^ rep) | (wl ^ rep) | (wr ^ rep))) == 0u;
            if (__all_sync(FULL, uni & A.uni & B.uni & (rep == B.rep) & (A.rep == B.rep) & (cand_cur == 0u))) {
                st.zmask |= 1u << ((y + 4) % 3);
                if (y >= y0) {
                    if (writer) { *kp = 0u; *sp = 0u; }
                    kp += k_wpr; sp += s_wpr;
                    last_kw = 0u; last_sw = 0u;
                }
                continue;
            }
        }
        make_hrow(C, vcur, wl, wr);
        const uint32_t cand_next = produce_row(R, st, (y + 4) % 3, (y + 1) & 1, lane, A, B, C, y + 1 >= 0 && y + 1 < H);   // magnitude row y+1
        __syncwarp();
        if (y >= y0) {
            uint32_t kept16, strong16;
            nms_row(R, st, S, y, lane, out_lane ? cand_cur : 0u, hi, kept16, strong16);
            const uint32_t k_up = __shfl_down_sync(FULL, kept16, 1), s_up = __shfl_down_sync(FULL, strong16, 1);
            SS_DEVICE_ASSERT(!writer || (y < H && (max(x, 0) >> 5) < k_wpr));
            last_kw = kept16 | (k_up << 16); last_sw = strong16 | (s_up << 16);
            if (writer) { *kp = last_kw; *sp = last_sw; }
            kp += k_wpr; sp += s_wpr;
        }
        cand_cur = cand_next;
        __syncwarp();          // ring row (y+2) mod 3 == (y-1) mod 3 is overwritten next iteration
        const HRow t = A; A = B; B = C; C = t;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Fused front end: RGB pages -> cv2 grey plane + Canny classes in ONE pass over the RGB bytes, fed by TMA.
//
// The RGB rows of a strip are fetched by the tensor memory accelerator: one elected lane issues
// cp.async.bulk.tensor.2d (SASS UTMALDG) per strip row and the bytes land in a per-warp shared-memory ring while the
// warp works on the rows before; an mbarrier per ring stage signals arrival (complete_tx).  A tightly packed RGB page has
// a row pitch of 3 W bytes (7650 at 300 DPI), not a multiple of 16, so the pages cannot be described as a 2-D image
// tensor; the tensor map instead views the whole batch as rows of 128 bytes ({128, total / 128} u8, box {128, 13}): a box is
// the 128-byte-aligned superset of the 1536 bytes a strip row needs (13 requests of 128 bytes to the memory system), and
// whatever lies beyond the end of the batch is zero-filled by the hardware (no guard code for the last rows).  Each lane then
// takes its 64-byte window (16-byte aligned: the 16-byte part of the row's misalignment is an address offset) with four
// LDS.128, realigns it in registers (the misalignment is uniform per row), converts its 16 pixels with cv2's 15-bit
// fixed-point formula (pixel.cuh) and -- lanes 1..30, rows of the own band -- stores them into the grey plane the
// threshold kernel reads afterwards.  From there on the row goes through exactly the arithmetic of
// canny_classes_kernel above.  Compared with rgb2gray + canny_classes the grey plane is written once and read once
// instead of twice, and the RGB read overlaps the issue-bound stencil instead of being a kernel of its own.
#ifndef SYNSEG_CR_DEPTH
#define SYNSEG_CR_DEPTH 2
#endif
constexpr int CR_DEPTH = SYNSEG_CR_DEPTH;          // RGB rows in flight per warp
constexpr int CR_UNIT = 128;                       // bytes per tensor-map row (the unit a box is aligned to)
constexpr int CR_UNITS = 13;                       // units per box: a strip row starts up to 127 bytes into its first unit; the last lane's 64-byte window ends at
                                                   // 112 + 1488 + 64 = 1664 = 13 x 128
constexpr int CR_BOX = CR_UNITS * CR_UNIT;         // 1664 bytes per TMA box (13 requests of 128 bytes; the first version used 98 units of 16 bytes and
                                                   // kept the warps of blank pages waiting on the TMA unit: 15 % of the stall samples sat in mbar_wait)
constexpr int CR_STAGE = CR_BOX;                   // ring stage (a multiple of the 128-byte alignment a TMA destination needs)
// measured on B200 (profiles/r2_front_end_experiments.txt): 2 stages are enough -- the kernel is bound by instruction issue, not by
// the latency of the loads; deeper rings only cost resident warps (3: +1 %, 4: +3 %, 6: +19 %), an L2 prefetch 8 rows ahead +26 %
#ifdef SYNSEG_CR_MINB
constexpr int CR_MINBLOCKS = SYNSEG_CR_MINB;
#else
// 2 stages: 17 CTAs of 12.9 KB fit the 227 KB of an SM; asking for 17 makes ptxas settle on 96 registers (20 bytes of spills) instead of 128:
// 0.615 -> 0.607 ms per 50 pages
constexpr int CR_MINBLOCKS = CR_DEPTH <= 2 ? 17 : (CR_DEPTH == 3 ? 16 : (CR_DEPTH == 4 ? 14 : (CR_DEPTH == 5 ? 13 : 12)));
#endif
#ifndef SYNSEG_CR_PREFETCH
#define SYNSEG_CR_PREFETCH 0
#endif
constexpr int CR_PREFETCH = SYNSEG_CR_PREFETCH;    // rows ahead of the TMA loads whose boxes are prefetched into L2 (0 = off)

struct CrParams {
    CnParams c;
    Plane gray;                                    // grey plane written (16-byte aligned pitch >= round_up(width, 16))
    int64_t src_bs, src_rs;                        // RGB batch / row stride in bytes
    int64_t total_bytes;                           // bytes of the whole batch (the tensor map covers floor(total / 16) units)
    const uint8_t *src;                            // for the <= 15 tail bytes the tensor map cannot cover
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
#ifndef SYNSEG_CR_WAIT_HINT
#define SYNSEG_CR_WAIT_HINT 0
#endif
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
#if SYNSEG_CR_WAIT_HINT > 0
    // with a suspend-time hint the warp sleeps in the barrier unit instead of re-issuing try_wait (7 % of the warp instructions of the kernel)
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity), "r"((unsigned)SYNSEG_CR_WAIT_HINT) : "memory");
#else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
#endif
}
__device__ __forceinline__ void tma_prefetch_units(const CUtensorMap *map, int unit)
{
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(map), "r"(0), "r"(unit) : "memory");
}
__device__ __forceinline__ void tma_load_units(void *dst, const CUtensorMap *map, int unit, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(0), "r"(unit), "r"(smem_u32(bar))
                 : "memory");
}

// 12 RGB bytes = 4 pixels -> 4 cv2 grey bytes.  Same value as gray.cu:gray4 with every term doubled,
// (19596 R + 38470 G + 7470 B + 32768) >> 16 == (9798 R + 19235 G + 3735 B + 16384) >> 15, so that the grey level is byte 2 of the
// sum (< 2^24) and four of them are gathered with three PRMT instead of four shifts and three merges.
__device__ __forceinline__ uint32_t cv_gray_b2(uint32_t rgbx)
{
    constexpr uint32_t lo = (19596u & 255u) | ((38470u & 255u) << 8) | ((7470u & 255u) << 16);
    constexpr uint32_t hi = (19596u >> 8) | ((38470u >> 8) << 8) | ((7470u >> 8) << 16);
    return __dp4a(rgbx, lo, 32768u) + (__dp4a(rgbx, hi, 0u) << 8);
}
__device__ __forceinline__ uint32_t cv_gray4(uint32_t w0, uint32_t w1, uint32_t w2)
{
    const uint32_t y0 = cv_gray_b2(w0);
    const uint32_t y1 = cv_gray_b2(__funnelshift_r(w0, w1, 24));
    const uint32_t y2 = cv_gray_b2(__funnelshift_r(w1, w2, 16));
    const uint32_t y3 = cv_gray_b2(w2 >> 8);
    return __byte_perm(__byte_perm(y0, y1, 0x0062), __byte_perm(y2, y3, 0x0062), 0x5410);
}

struct CrSmem {
    MagRing ring;
    NmsStage stage;
    alignas(128) uint8_t rows[CR_DEPTH][CR_STAGE];
    alignas(8) uint64_t bar[CR_DEPTH];
};

__global__ void __launch_bounds__(32, CR_MINBLOCKS) canny_rgb_kernel(const __grid_constant__ CUtensorMap tmap, CrParams q)
{
    __shared__ CrSmem sm;
    const CnParams &p = q.c;
    const int lane = threadIdx.x;
    int64_t task = blockIdx.x;
    const int strip = (int)(task % p.strips); task /= p.strips;
    const int band = (int)(task % p.bands);
    const int img = (int)(task / p.bands);
    MagRing &R = sm.ring;
    NmsStage &S = sm.stage;

    const int W = p.width, H = p.height;
    const int hi = p.hi;
    const int x = strip * CN_OUT_W - 16 + 16 * lane;  // first of this lane's 16 columns
    const int y0 = band * p.band_h, y1 = min(y0 + p.band_h, H);
    const bool out_lane = lane >= 1 && lane <= 30 && x < W;
    const bool writer = out_lane && (lane & 1) && lane <= 29;

    CnState st;
    st.zmask = 0;
    const int lo = p.lo < 0 ? 0 : (p.lo > 4095 ? 4095 : p.lo);
    st.kpair = (uint32_t)(0x7FFF - lo) * 0x00010001u;
    st.cv16 = 0;
#pragma unroll
    for (int c = 0; c < 16; ++c) st.cv16 |= (x + c >= 0 && x + c < W) ? (1u << c) : 0u;

    // the strip's bytes of a row start at column cb; this lane's (clamped) 16-pixel group starts 3 (xl - cb) bytes further
    const int cb = max(strip * CN_OUT_W - 16, 0);
    const int xl = clamp16_x(x, W);
    const EdgeFix efix = make_edge_fix(x, W);
    const int lane_off = 3 * (xl - cb);               // multiple of 48
    SS_DEVICE_ASSERT(lane_off >= 0 && lane_off + 64 + (CR_UNIT - 16) <= CR_BOX && (lane_off & 15) == 0);
    if (lane == 0) {
#pragma unroll
        for (int d = 0; d < CR_DEPTH; ++d) mbar_init(&sm.bar[d], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    // byte offset (from the start of the batch) of the strip's first byte in the row being fetched = off0 + rel, rel = (row clamped into the
    // image) x row stride < 2^31 (canny_rgb_supported): the row loop tracks 32-bit quantities only -- tensor-map row unit0 + (lo0 + rel) / 128
    int y_pf = y0 - 2;
    const int64_t off0 = (int64_t)img * q.src_bs + 3 * (int64_t)cb;
    const int unit0 = (int)(off0 >> 7), lo0 = (int)(off0 & (CR_UNIT - 1)), rs32 = (int)q.src_rs;
    int rel_pf = min(max(y_pf, 0), H - 1) * rs32;
    const int64_t tail_start = q.total_bytes & ~(int64_t)(CR_UNIT - 1);      // bytes from here on are not covered by the tensor map
    int y_l2 = y_pf;
    int rel_l2 = rel_pf;
    if (CR_PREFETCH > 0) {
        for (int d = 0; d < CR_DEPTH + CR_PREFETCH; ++d) {
            if (lane == 0 && d >= CR_DEPTH && y_l2 <= y1 + 1) tma_prefetch_units(&tmap, unit0 + ((lo0 + rel_l2) >> 7));
            ++y_l2;
            if ((unsigned)(y_l2 - 1) < (unsigned)(H - 1)) rel_l2 += rs32;
        }
    }
    auto issue = [&](int slot) {
        // rows beyond y1 + 1 are never consumed: nothing may be in flight into this CTA's shared memory when it exits
        if (lane == 0 && y_pf <= y1 + 1) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the stage was read through the generic proxy
            mbar_expect_tx(&sm.bar[slot], CR_BOX);
            tma_load_units(sm.rows[slot], &tmap, unit0 + ((lo0 + rel_pf) >> 7), &sm.bar[slot]);
        }
        ++y_pf;
        if ((unsigned)(y_pf - 1) < (unsigned)(H - 1)) rel_pf += rs32;      // the offset only moves while the row index is inside [1, H-1]
        if (CR_PREFETCH > 0) {                             // the row CR_PREFETCH rows further on starts its way from HBM to L2 now
            if (lane == 0 && y_l2 <= y1 + 1) tma_prefetch_units(&tmap, unit0 + ((lo0 + rel_l2) >> 7));
            ++y_l2;
            if ((unsigned)(y_l2 - 1) < (unsigned)(H - 1)) rel_l2 += rs32;
        }
    };
    int rel_cur = rel_pf;                              // offset of the row consumed next (same recurrence, CR_DEPTH rows behind)
    int y_cur = y_pf;
#pragma unroll
    for (int d = 0; d < CR_DEPTH; ++d) issue(d);
    uint32_t *kp = p.kept.p + img * p.kept.bs + (int64_t)y0 * p.kept.wpr + (max(x, 0) >> 5);
    uint32_t *sp = p.strong.p + img * p.strong.bs + (int64_t)y0 * p.strong.wpr + (max(x, 0) >> 5);
    const int k_wpr = p.kept.wpr, s_wpr = p.strong.wpr;
    uint8_t *gp = q.gray.p + img * q.gray.bs + (int64_t)y0 * q.gray.rs + max(x, 0);     // grey row y0 of this lane's group (advances with the own rows)
    const int64_t g_rs = q.gray.rs;
    // the last 128-byte unit of the batch may be partial and then lies outside the tensor map: only the band holding the last rows checks for it
    const bool tail_band = (q.total_bytes & (CR_UNIT - 1)) != 0 &&
                           (int64_t)img * q.src_bs + (int64_t)min(y1 + 1, H - 1) * q.src_rs + 3 * (int64_t)(cb + CN_OUT_W + 32) > tail_start;

    HRow A, B, C;
    A.uni = false; B.uni = false; A.rep = 0; B.rep = 1;
#pragma unroll
    for (int i = 0; i < 8; ++i) { A.s[i] = 0; A.d[i] = 0; B.s[i] = 0; B.d[i] = 0; }
    int slot = 0;
    uint32_t phases = 0;                               // bit d: parity the next wait on stage d expects
    uint32_t cand_cur = 0;
    int rep = 0;                                       // repeated rows, see REPEAT_NOTE in canny_classes_kernel
    uint4 vprev = make_uint4(0, 0, 0, 0);
    uint32_t last_kw = 0, last_sw = 0;
#pragma unroll 1
    for (int y = y0 - 4; y < y1; ++y) {
        // ---- grey row y + 2 from the RGB stage ------------------------------------------------------------------
        mbar_wait(&sm.bar[slot], (phases >> slot) & 1u);
        phases ^= 1u << slot;
        const int a = (lo0 + rel_cur) & 15;
        const int a_unit = (lo0 + rel_cur) & (CR_UNIT - 16);      // 16-byte pieces between the start of the box and the strip's first byte
        if (tail_band && off0 + rel_cur + 3 * (int64_t)(CN_OUT_W + 32) > tail_start) {
            // last rows of the batch: the final partial 128-byte unit is outside the tensor map (zero-filled); patch it in
            const int64_t first = ((off0 + rel_cur) & ~(int64_t)(CR_UNIT - 1));
            for (int64_t b = tail_start + lane; b < q.total_bytes; b += 32) {
                const int64_t o = b - first;
                if (o >= 0 && o < CR_BOX) sm.rows[slot][o] = q.src[b];
            }
            __syncwarp();
        }
        uint4 vcur;
        {
            const uint4 *wp = (const uint4 *)(sm.rows[slot] + lane_off + a_unit);
            const uint4 q0 = wp[0], q1 = wp[1], q2 = wp[2], q3 = wp[3];
            const uint32_t w[16] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w};
            // 64 equal bytes: whatever the alignment, the 16 pixels are grey level c (R = G = B = c gives c exactly) -- blank paper and
            // the inside of solid boxes skip the realignment and the conversion (the test used to come after the realignment)
            uint32_t rd = 0;
#pragma unroll
            for (int i = 1; i < 16; ++i) rd |= w[i] ^ w[0];
            if ((rd | (w[0] ^ __byte_perm(w[0], 0, 0x0000))) == 0u) {
                vcur = make_uint4(w[0], w[0], w[0], w[0]);
            } else {
                const int ws = a >> 2, bsh = (a & 3) * 8;
                uint32_t v[12];
                switch (ws) {                                  // uniform over the warp: v[] stays in registers
                case 0:
#pragma unroll
                    for (int i = 0; i < 12; ++i) v[i] = __funnelshift_r(w[i], w[i + 1], bsh);
                    break;
                case 1:
#pragma unroll
                    for (int i = 0; i < 12; ++i) v[i] = __funnelshift_r(w[i + 1], w[i + 2], bsh);
                    break;
                case 2:
#pragma unroll
                    for (int i = 0; i < 12; ++i) v[i] = __funnelshift_r(w[i + 2], w[i + 3], bsh);
                    break;
                default:
#pragma unroll
                    for (int i = 0; i < 12; ++i) v[i] = __funnelshift_r(w[i + 3], i + 4 < 16 ? w[i + 4] : 0u, bsh);
                    break;
                }
                vcur.x = cv_gray4(v[0], v[1], v[2]); vcur.y = cv_gray4(v[3], v[4], v[5]);
                vcur.z = cv_gray4(v[6], v[7], v[8]); vcur.w = cv_gray4(v[9], v[10], v[11]);
            }
        }
        __syncwarp();                                      // every lane has read its window: the stage can be refilled
        issue(slot);
        slot = (slot + 1 == CR_DEPTH) ? 0 : slot + 1;
        ++y_cur;
        if ((unsigned)(y_cur - 1) < (unsigned)(H - 1)) rel_cur += rs32;
        if (y + 2 >= y0 && y + 2 < y1) {                   // grey plane, own rows only
            SS_DEVICE_ASSERT(!out_lane || (y + 2 < H && x >= 0 && x + 16 <= g_rs));
            if (out_lane) *(uint4 *)gp = vcur;
            gp += g_rs;
        }
        vcur = apply_edge_fix(vcur, efix);
        {
            const bool same = ((vcur.x ^ vprev.x) | (vcur.y ^ vprev.y) | (vcur.z ^ vprev.z) | (vcur.w ^ vprev.w)) == 0u;
            rep = (y > y0 - 4 && __all_sync(FULL, same)) ? rep + 1 : 0;
            vprev = vcur;
            if (rep >= 5 && y - 3 >= 0 && y + 2 <= H - 1 && y >= y0 + 1) {      // output row y == output row y - 1 (REPEAT_NOTE)
                if (writer) { *kp = last_kw; *sp = last_sw; }
                kp += k_wpr; sp += s_wpr;
                continue;
            }
        }
        const uint32_t wl = __shfl_up_sync(FULL, vcur.w, 1), wr = __shfl_down_sync(FULL, vcur.x, 1);
        if (y < y0 - 2) {                                  // priming: rows y0-2, y0-1 only become partials
            make_hrow(C, vcur, wl, wr);
            const HRow t = A; A = B; B = C; C = t;
            continue;
        }
        {
            const uint32_t rep = __byte_perm(vcur.x, 0, 0x0000);
            const bool uni = (((vcur.x ^ rep) | (vcur.y ^ rep) | (vcur.z ^ rep)) | ((vcur.w ^ rep) | (wl ^ rep) | (wr ^ rep))) == 0u;
            if (__all_sync(FULL, uni & A.uni & B.uni & (rep == B.rep) & (A.rep == B.rep) & (cand_cur == 0u))) {
                st.zmask |= 1u << ((y + 4) % 3);
                if (y >= y0) {
                    if (writer) { *kp = 0u; *sp = 0u; }
                    kp += k_wpr; sp += s_wpr;
                    last_kw = 0u; last_sw = 0u;
                }
                continue;
            }
        }
        make_hrow(C, vcur, wl, wr);
        const uint32_t cand_next = produce_row(R, st, (y + 4) % 3, (y + 1) & 1, lane, A, B, C, y + 1 >= 0 && y + 1 < H);   // magnitude row y+1
        __syncwarp();
        if (y >= y0) {
            uint32_t kept16, strong16;
            nms_row(R, st, S, y, lane, out_lane ? cand_cur : 0u, hi, kept16, strong16);
            const uint32_t k_up = __shfl_down_sync(FULL, kept16, 1), s_up = __shfl_down_sync(FULL, strong16, 1);
            last_kw = kept16 | (k_up << 16); last_sw = strong16 | (s_up << 16);
            if (writer) { *kp = last_kw; *sp = last_sw; }
            kp += k_wpr; sp += s_wpr;
        }
        cand_cur = cand_next;
        __syncwarp();
        const HRow t = A; A = B; B = C; C = t;
    }
}

}  // namespace

// ---- host side of the fused front end ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn tensor_map_encoder()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        else cudaGetLastError();
    }
    return fn;
}

// Can the fused RGB front end take these pages?  (16-byte aligned batch base, flat size below the coordinate range of a tensor map.)
bool canny_rgb_supported(const synseg_img *rgb, const synseg_img *gray)
{
    if (getenv("SYNSEG_NO_TMA")) return false;
    if (((uintptr_t)rgb->data & 15) != 0) return false;
    const int64_t total = (int64_t)(rgb->batch - 1) * rgb->batch_stride + (int64_t)(rgb->height - 1) * rgb->row_stride + 3 * (int64_t)rgb->width;
    if (total < CR_BOX || (total >> 7) >= 0x7fffffffLL) return false;
    if ((int64_t)rgb->height * rgb->row_stride >= 0x7fffff00LL || rgb->row_stride < 0) return false;      // 32-bit row offsets inside an image
    if (!plane_aligned(gray, 16) || gray->row_stride < (int64_t)align_up((size_t)rgb->width, 16)) return false;
    return tensor_map_encoder() != nullptr;
}

// RGB pages -> cv2 grey plane + Canny classes (kept / strong bit planes), one TMA-fed kernel.
int launch_canny_rgb(synseg_ctx *ctx, const synseg_img *rgb, const synseg_img *gray, BitPlane kept, BitPlane strong, int lo, int hi, cudaStream_t st)
{
    CrParams q;
    CnParams &p = q.c;
    p.src = plane_of(gray);
    p.kept = kept; p.strong = strong;
    p.width = rgb->width; p.height = rgb->height; p.lo = lo; p.hi = hi;
    p.strips = cdiv(rgb->width, CN_OUT_W);
    // bands of up to 48 rows: a band converts 4 halo rows besides its own (32-row bands: 0.817 ms, 48: 0.807, 64: 0.842 per 50 pages)
    const int64_t rows_total = (int64_t)rgb->height * rgb->batch * p.strips;
    int band_h = (int)(rows_total / (48 * (int64_t)ctx->sm_count));
    band_h = band_h < 16 ? 16 : (band_h > 48 ? 48 : band_h);
    if (ctx->tune_canny_band > 0) band_h = ctx->tune_canny_band;
    if (band_h > rgb->height) band_h = rgb->height;
    p.band_h = band_h;
    p.bands = cdiv(rgb->height, band_h);
    p.tasks = (int64_t)rgb->batch * p.bands * p.strips;
    q.gray = plane_of(gray);
    q.src_bs = rgb->batch_stride; q.src_rs = rgb->row_stride;
    q.total_bytes = (int64_t)(rgb->batch - 1) * rgb->batch_stride + (int64_t)(rgb->height - 1) * rgb->row_stride + 3 * (int64_t)rgb->width;
    q.src = (const uint8_t *)rgb->data;
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {(cuuint64_t)CR_UNIT, (cuuint64_t)(q.total_bytes / CR_UNIT)};
    const cuuint64_t gstride[1] = {(cuuint64_t)CR_UNIT};
    const cuuint32_t box[2] = {(cuuint32_t)CR_UNIT, (cuuint32_t)CR_UNITS};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = tensor_map_encoder()(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, rgb->data, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { synseg_set_error("canny_rgb: cuTensorMapEncodeTiled failed (%d)", (int)r); return SYNSEG_E_CUDA; }
    if (p.tasks > 0x7fffffffLL) { synseg_set_error("canny_rgb: batch too large"); return SYNSEG_E_INVALID; }
    canny_rgb_kernel<<<(unsigned)p.tasks, 32, 0, st>>>(tmap, q);
    SS_LAUNCH_CHECK(ctx, "canny_rgb", st);
    return SYNSEG_OK;
}

// The fused front end of the page pipeline: RGB -> grey plane + Canny classes (TMA-fed), adaptive threshold of the grey plane
// into `mask`, hysteresis OR-ed into `mask`.  Scratch (two bit planes + the hysteresis arrays) comes from the arena top.
int run_front_rgb(synseg_ctx *ctx, const synseg_img *rgb, const synseg_img *gray, BitPlane mask, int block_size, int C, int lo, int hi, cudaStream_t st)
{
    const int W = rgb->width, H = rgb->height, B = rgb->batch;
    const int wpr = bit_wpr(W);
    const size_t plane_bytes = (size_t)wpr * H * B * 4;
    void *p;
    SS_TRY(arena_alloc(ctx, plane_bytes, &p, st));
    BitPlane kept{(uint32_t *)p, wpr, (int64_t)wpr * H, nullptr};
    SS_TRY(arena_alloc(ctx, plane_bytes, &p, st));
    BitPlane strong{(uint32_t *)p, wpr, (int64_t)wpr * H, nullptr};
    SS_TRY(launch_canny_rgb(ctx, rgb, gray, kept, strong, lo, hi, st));
    SS_TRY(launch_adaptive_mean(ctx, gray, nullptr, mask, block_size, C, 1, st));
    return run_hysteresis(ctx, kept, strong, W, H, B, nullptr, mask, true, st);
}

int launch_canny_classes(synseg_ctx *ctx, const synseg_img *gray, BitPlane kept, BitPlane strong, int lo, int hi, cudaStream_t st)
{
    CnParams p;
    p.src = plane_of(gray);
    p.kept = kept; p.strong = strong;
    p.width = gray->width; p.height = gray->height; p.lo = lo; p.hi = hi;
    p.strips = cdiv(gray->width, CN_OUT_W);
    // Bands: a band recomputes 2 magnitude rows of its neighbours (6 % at 32 rows).  Short bands win: the cost of a
    // band depends on its content (blank rows are ~10x cheaper than text rows), so many small tasks balance the
    // SMs better than few long ones (measured on B200: 128 rows 1.40 ms, 32 rows 0.91 ms per 50 pages).
    const int64_t rows_total = (int64_t)gray->height * gray->batch * p.strips;
    int band_h = (int)(rows_total / (64 * (int64_t)ctx->sm_count));
    band_h = band_h < 16 ? 16 : (band_h > 32 ? 32 : band_h);
    if (ctx->tune_canny_band > 0) band_h = ctx->tune_canny_band;
    if (band_h > gray->height) band_h = gray->height;
    p.band_h = band_h;
    p.bands = cdiv(gray->height, band_h);
    p.tasks = (int64_t)gray->batch * p.bands * p.strips;
    if (plane_aligned(gray, 16)) canny_classes_kernel<true><<<(unsigned)cdiv(p.tasks, CN_WARPS), 32 * CN_WARPS, 0, st>>>(p);
    else canny_classes_kernel<false><<<(unsigned)cdiv(p.tasks, CN_WARPS), 32 * CN_WARPS, 0, st>>>(p);
    SS_LAUNCH_CHECK(ctx, "canny_classes", st);
    return SYNSEG_OK;
}

size_t canny_scratch_bytes(int width, int height, int batch)
{
    return 2 * ((size_t)bit_wpr(width) * height * batch * 4 + 256) + hysteresis_scratch_bytes(width, height, batch) + 256;
}

int run_canny(synseg_ctx *ctx, const synseg_img *gray, const synseg_img *edges_u8, BitPlane edges_bits, bool or_bits,
              int lo, int hi, cudaStream_t st)
{
    const int W = gray->width, H = gray->height, B = gray->batch;
    const int wpr = bit_wpr(W);
    const size_t plane_bytes = (size_t)wpr * H * B * 4;
    void *p;
    SS_TRY(arena_alloc(ctx, plane_bytes, &p, st));
    BitPlane kept{(uint32_t *)p, wpr, (int64_t)wpr * H, edges_bits.dims};      // a ragged batch hands its dims on
    SS_TRY(arena_alloc(ctx, plane_bytes, &p, st));
    BitPlane strong{(uint32_t *)p, wpr, (int64_t)wpr * H, edges_bits.dims};
    SS_TRY(launch_canny_classes(ctx, gray, kept, strong, lo, hi, st));
    return run_hysteresis(ctx, kept, strong, W, H, B, edges_u8, edges_bits, or_bits, st);
}

extern "C" SYNSEG_EXPORT int synseg_canny(synseg_ctx *ctx, const synseg_img *gray, const synseg_img *edges, int lo, int hi, void *stream)
{
    if (!ctx) { synseg_set_error("synseg_canny: ctx is NULL"); return SYNSEG_E_INVALID; }
    SS_ENTER(ctx, stream);
    SS_TRY(validate_img(gray, "gray", 1));
    SS_TRY(validate_img(edges, "edges", 1));
    if (!same_shape(gray, edges)) { synseg_set_error("synseg_canny: shape mismatch"); return SYNSEG_E_INVALID; }
    if (lo < 0 || hi < lo) { synseg_set_error("synseg_canny: need 0 <= lo <= hi"); return SYNSEG_E_INVALID; }
    if (gray->batch > 65535) { synseg_set_error("synseg_canny: batch > 65535"); return SYNSEG_E_INVALID; }
    SS_TRY(arena_ensure(ctx, canny_scratch_bytes(gray->width, gray->height, gray->batch)));
    arena_begin(ctx);
    return run_canny(ctx, gray, edges, BitPlane{nullptr, 0, 0}, false, lo, hi, (cudaStream_t)stream);
}
