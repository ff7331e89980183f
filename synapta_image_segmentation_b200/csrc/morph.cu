// morph.cu -- rectangular erode / dilate / open / close, bit-exact with cv2.erode, cv2.dilate and
// cv2.morphologyEx(MORPH_RECT element, explicit anchor, iterations, default constant border).
//
// Reference call sites: morphologyEx(edges, MORPH_OPEN, rect, iterations=2) at
// pdf_image_segmentation.py:1370,1375 (1 x max(20,H/20) and max(20,W/20) x 1) and :1556,1557
// (25x1, 1x25); dilate / MORPH_CLOSE with a k x k element are north-star primitives.
//
// Semantics (verified against cv2 4.13): dst(x) = op over i in [0,k) of
// src(x + i - anchor), pixels outside the image ignored; `iterations` of a rect element fold into one
// pass with k' = it*(k-1)+1 and anchor' = it*anchor (what cv2 does itself); OPEN = erode then dilate,
// CLOSE = dilate then erode; a rect element is separable into a row pass and a column pass.
//
// Two implementations:
//  * binary fast path (SYNSEG_MORPH_BINARY, and everything inside the fused page pipeline): the mask
//    is bit-packed (1 bit/pixel, 32 pixels per word); a row pass ORs funnel-shifted span-8 words, a
//    column pass is a per-thread in-place doubling over a shared-memory tile of words; erosion runs as
//    the complement of dilation.  HBM traffic of the u8 entry point is the pack (1 B/px read) and the
//    unpack (1 B/px written): 2 algorithmic bytes per pixel, the intermediate bit planes (1/8 B/px)
//    stay in L2.
//  * generic 8-bit path: per-thread in-place doubling (sparse-table) over packed bytes in shared
//    memory with __vmaxu4/__vminu4 -- any grey image, any k.
#include <stdlib.h>

#include "internal.cuh"
#include "pixel.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
// pack / unpack / count
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t nibble_of(uint32_t four_bytes)
{
    // byte != 0 -> bit; (t * 0x204081) >> 21 gathers the four 0/1 bytes into a nibble
    uint32_t t = __vcmpne4(four_bytes, 0u) & 0x01010101u;
    return ((t * 0x00204081u) >> 21) & 0xFu;
}

__device__ __forceinline__ uint32_t bytes_of(uint32_t nib)
{
    return ((nib * 0x00204081u) & 0x01010101u) * 0xFFu;
}

__global__ void __launch_bounds__(256) pack_bits_kernel(Plane src, BitPlane dst, int width, int height, int nw, int64_t total, bool aligned)
{
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= total) return;
    const int w = (int)(i % nw);
    const int64_t row = i / nw;
    const int img = (int)(row / height);
    const int y = (int)(row - (int64_t)img * height);
    const uint8_t *sp = src.p + img * src.bs + y * src.rs + 32 * w;
    uint32_t out = 0;
    if (aligned && 32 * w + 32 <= width) {
        const uint4 a = __ldg((const uint4 *)sp), b = __ldg((const uint4 *)sp + 1);
        out = nibble_of(a.x) | (nibble_of(a.y) << 4) | (nibble_of(a.z) << 8) | (nibble_of(a.w) << 12) |
              (nibble_of(b.x) << 16) | (nibble_of(b.y) << 20) | (nibble_of(b.z) << 24) | (nibble_of(b.w) << 28);
    } else {
        const int n = min(32, width - 32 * w);
        for (int j = 0; j < n; ++j) out |= (uint32_t)(__ldg(sp + j) != 0) << j;
    }
    dst.p[img * dst.bs + (int64_t)y * dst.wpr + w] = out;
}

__global__ void __launch_bounds__(256) unpack_bits_kernel(BitPlane src, Plane dst, int width, int height, int nw, int64_t total, bool aligned)
{
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= total) return;
    const int w = (int)(i % nw);
    const int64_t row = i / nw;
    const int img = (int)(row / height);
    const int y = (int)(row - (int64_t)img * height);
    const uint32_t v = src.p[img * src.bs + (int64_t)y * src.wpr + w];
    uint8_t *dp = dst.p + img * dst.bs + y * dst.rs + 32 * w;
    if (aligned && 32 * w + 32 <= width) {
        uint4 a, b;
        a.x = bytes_of(v & 15u); a.y = bytes_of((v >> 4) & 15u); a.z = bytes_of((v >> 8) & 15u); a.w = bytes_of((v >> 12) & 15u);
        b.x = bytes_of((v >> 16) & 15u); b.y = bytes_of((v >> 20) & 15u); b.z = bytes_of((v >> 24) & 15u); b.w = bytes_of(v >> 28);
        ((uint4 *)dp)[0] = a; ((uint4 *)dp)[1] = b;
    } else {
        const int n = min(32, width - 32 * w);
        for (int j = 0; j < n; ++j) dp[j] = ((v >> j) & 1u) ? 255 : 0;
    }
}

__global__ void __launch_bounds__(256) count_bits_kernel(BitPlane src, int nw, int height, unsigned long long *out, int out_stride)
{
    const int img = blockIdx.y;
    if (src.dims) { const int2 d = src.dims[img]; nw = (d.x + 31) >> 5; height = d.y; }     // ragged batch
    const int64_t total = (int64_t)nw * height;
    unsigned cnt = 0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int w = (int)(i % nw);
        const int64_t y = i / nw;
        cnt += __popc(src.p[img * src.bs + y * src.wpr + w]);
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(out + (int64_t)img * out_stride, (unsigned long long)cnt);
}

// ---------------------------------------------------------------------------------------------
// bit-plane row pass: one warp per row
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ext32(const uint32_t *a, int bitpos)
{
    const int j = bitpos >> 5, sh = bitpos & 31;
    return __funnelshift_r(a[j], a[j + 1], sh);
}

__global__ void __launch_bounds__(256) bitmorph_h_kernel(BitPlane src, BitPlane dst, int width, int height, int batch, int nw,
                                                         int erode, int k, int anchor, int padw)
{
    extern __shared__ uint32_t sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tot = nw + 2 * padw;              // padded words per row buffer
    uint32_t *in = sm + (size_t)warp * 2 * (tot + 1);
    uint32_t *s8 = in + tot + 1;
    const uint32_t last_mask = (width & 31) ? ((1u << (width & 31)) - 1u) : 0xffffffffu;
    const int64_t rows = (int64_t)height * batch;
    for (int64_t row = (int64_t)blockIdx.x * 8 + warp; row < rows; row += (int64_t)gridDim.x * 8) {
        const int img = (int)(row / height);
        const int y = (int)(row - (int64_t)img * height);
        const uint32_t *sp = src.p + img * src.bs + (int64_t)y * src.wpr;
        for (int i = lane; i < tot + 1; i += 32) {
            uint32_t v = 0;
            const int w = i - padw;
            if (w >= 0 && w < nw) {
                v = sp[w];
                if (erode) v = ~v;
                if (w == nw - 1) v &= last_mask;
            }
            in[i] = v;
        }
        __syncwarp();
        for (int i = lane; i < tot; i += 32) {
            unsigned long long v = ((unsigned long long)in[i + 1] << 32) | in[i];
            v |= v >> 1; v |= v >> 2; v |= v >> 4;
            s8[i] = (uint32_t)v;
        }
        if (lane == 0) s8[tot] = 0;
        __syncwarp();
        uint32_t *dp = dst.p + img * dst.bs + (int64_t)y * dst.wpr;
        const int lo = padw * 32 - anchor;         // bit offset of the window start for output bit 0
        for (int w = lane; w < nw; w += 32) {
            uint32_t acc = 0;
            const int b0 = 32 * w + lo;
            if (k >= 8) {
                int d = 0;
                for (; d + 8 <= k; d += 8) acc |= ext32(s8, b0 + d);
                if (d < k) acc |= ext32(s8, b0 + k - 8);
            } else {
                for (int d = 0; d < k; ++d) acc |= ext32(in, b0 + d);
            }
            if (erode) acc = ~acc;
            if (w == nw - 1) acc &= last_mask;
            dp[w] = acc;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// bit-plane column pass: one warp per (32 word-columns x TH rows) tile
// ---------------------------------------------------------------------------------------------
constexpr int BV_TH = 128;

__global__ void __launch_bounds__(32) bitmorph_v_kernel(BitPlane src, BitPlane dst, int width, int height, int nw, int erode, int k,
                                                        int anchor)
{
    extern __shared__ uint32_t sm[];
    const int lane = threadIdx.x;
    const int w = blockIdx.x * 32 + lane;
    const int y0 = blockIdx.y * BV_TH;
    const int img = blockIdx.z;
    const int th = min(BV_TH, height - y0);
    const int n = th + k - 1;                   // input rows y0 - anchor .. y0 - anchor + n - 1
    const bool active = w < nw;
    const uint32_t last_mask = (width & 31) ? ((1u << (width & 31)) - 1u) : 0xffffffffu;
    const uint32_t vmask = (w == nw - 1) ? last_mask : 0xffffffffu;
    const uint32_t *sp = src.p + img * src.bs + w;
    for (int j = 0; j < n; ++j) {
        const int y = y0 - anchor + j;
        uint32_t v = 0;
        if (active && y >= 0 && y < height) {
            v = sp[(int64_t)y * src.wpr];
            if (erode) v = ~v & vmask;
        }
        sm[j * 32 + lane] = v;
    }
    // in-place doubling: after the level with step s, sm[j] = OR of rows j .. j + 2s - 1
    int p = 1;
    while (2 * p <= k) {
        for (int j = 0; j + p < n; ++j) sm[j * 32 + lane] |= sm[(j + p) * 32 + lane];
        p *= 2;
    }
    if (!active) return;
    uint32_t *dp = dst.p + img * dst.bs + w;
    for (int j = 0; j < th; ++j) {
        uint32_t v = sm[j * 32 + lane] | sm[(j + k - p) * 32 + lane];
        if (erode) v = ~v;
        dp[(int64_t)(y0 + j) * dst.wpr] = v & vmask;
    }
}

// ---------------------------------------------------------------------------------------------
// bit-plane row pass, register version: one thread -> 4 consecutive output words (128 pixels)
// ---------------------------------------------------------------------------------------------
// The thread loads the NW words that cover its windows, ORs the array with itself shifted by 1, 2, 4, ...
// (in-place doubling: after the step with shift s every bit holds the OR of the 2s bits starting there), finishes
// with the shift k - span and extracts its four words.  No shared memory, no synchronisation: every output word
// costs ~45 instructions for k = 81 and the kernel runs at full occupancy.  NW = 8 serves k <= 98, NW = 12 k <= 226.
// NOUT = 8 output words per thread (k <= 98, NW = 12) shares the doubling steps among twice as many outputs: 12 x 2 instructions
// per step for 8 words instead of 8 x 2 for 4 (measured on B200: see DESIGN.md section 6).
// Grid: x over the (row, thread-of-row) pairs of one image, y = image -- the index arithmetic stays 32-bit (two 64-bit divisions per
// thread cost as much as the morphology of its eight words).
template <int NW, int NOUT>
__global__ void __launch_bounds__(256) bitmorph_h4_kernel(BitPlane src, BitPlane dst, int width, int height, int nw, int nq,
                                                          unsigned per_image, int erode, int k, int anchor)
{
    const unsigned i = blockIdx.x * 256u + threadIdx.x;
    if (i >= per_image) return;
    const int y = (int)(i / (unsigned)nq);
    const int q4 = (int)(i - (unsigned)y * (unsigned)nq);
    const int img = blockIdx.y;
    const int w0 = NOUT * q4;
    if (src.dims) {                                                   // ragged batch: this image may be smaller than the canvas
        const int2 d = src.dims[img];
        width = d.x; nw = (d.x + 31) >> 5;
        if (y >= d.y || w0 >= nw) return;
    }
    const int aw = (anchor + 31) >> 5, sft = 32 * aw - anchor;        // window of output bit b of word w0+q starts at bit 32q + b + sft of D
    const uint32_t last_mask = (width & 31) ? ((1u << (width & 31)) - 1u) : 0xffffffffu;
    const uint32_t *sp = src.p + img * src.bs + (int64_t)y * src.wpr;
    uint32_t D[NW + 1];
#pragma unroll
    for (int j = 0; j < NW; ++j) {
        const int w = w0 - aw + j;
        uint32_t v = 0;
        if (w >= 0 && w < nw) {
            v = __ldg(sp + w);
            if (erode) v = ~v;
            if (w == nw - 1) v &= last_mask;
        }
        D[j] = v;
    }
    D[NW] = 0;
    int span = 1;
#pragma unroll
    for (int step = 0; step < 5; ++step) {                 // bit shifts 1, 2, 4, 8, 16
        if (2 * span <= k) {
#pragma unroll
            for (int j = 0; j < NW; ++j) D[j] |= __funnelshift_r(D[j], D[j + 1], span);
            span *= 2;
        }
    }
    if (2 * span <= k) {                                   // 32
#pragma unroll
        for (int j = 0; j < NW; ++j) D[j] |= D[j + 1];
        span *= 2;
    }
    if (2 * span <= k) {                                   // 64
#pragma unroll
        for (int j = 0; j < NW; ++j) D[j] |= (j + 2 <= NW) ? D[j + 2 < NW ? j + 2 : NW] : 0u;
        span *= 2;
    }
    if (NW > 8 && 2 * span <= k) {                         // 128
#pragma unroll
        for (int j = 0; j < NW; ++j) D[j] |= D[j + 4 < NW ? j + 4 : NW];
        span *= 2;
    }
    // E = D | (D >> (k - span)), words 0..NOUT.  k - span < span: the word part of the shift is 0 or 1 while span <= 64 (k <= 127)
    const int t = k - span, tw = t >> 5, tb = t & 31;
    constexpr int MAXTW = (NW == 8 || NOUT == 8) ? 2 : 4;       // both are only launched for k <= 98
    uint32_t E[NOUT + 1];
#pragma unroll
    for (int j = 0; j < NOUT + 1; ++j) {
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int m = 0; m < MAXTW; ++m)
            if (tw == m) { lo = D[j + m < NW ? j + m : NW]; hi = D[j + m + 1 < NW ? j + m + 1 : NW]; }
        E[j] = D[j] | __funnelshift_r(lo, hi, tb);
    }
    uint32_t o[NOUT];
#pragma unroll
    for (int q = 0; q < NOUT; ++q) {
        uint32_t v = __funnelshift_r(E[q], E[q + 1], sft);
        if (erode) v = ~v;
        const int w = w0 + q;
        if (w == nw - 1) v &= last_mask;
        if (w >= nw) v = 0;
        o[q] = v;
    }
    uint32_t *dp = dst.p + img * dst.bs + (int64_t)y * dst.wpr + w0;
    *(uint4 *)dp = make_uint4(o[0], o[1], o[2], o[3]);
    if (NOUT == 8 && w0 + 4 < dst.wpr) *(uint4 *)(dp + 4) = make_uint4(o[NOUT - 4], o[NOUT - 3], o[NOUT - 2], o[NOUT - 1]);
}

// ---------------------------------------------------------------------------------------------
// bit-plane column pass, van Herk / Gil-Werman: one thread -> one word column x k output rows
// ---------------------------------------------------------------------------------------------
// Output row ys + i (0 <= i < k) needs input rows [u0 + i, u0 + i + k - 1], u0 = ys - anchor: the suffix OR of block
// [u0, u0 + k) from i on, and the prefix OR of block [u0 + k, u0 + 2k) up to i - 1.  The thread loads the first block
// into its private shared-memory column (conflict-free: [row][thread]), turns it into suffix ORs in place, then
// streams the second block keeping the running prefix OR in a register: ~12 instructions per output word for any k.
// Grid: x over the (segment, word column) pairs of one image, y = image (32-bit index arithmetic).  A segment whose 2k - 1 input rows
// and k output rows all lie inside the image (all but the first and last one or two) takes a path without any bounds test: running
// pointers, cp.async for the first block, VB loads of the second block in flight per batch.
#ifndef SYNSEG_MORPH_VB
#define SYNSEG_MORPH_VB 8
#endif
constexpr int VB = SYNSEG_MORPH_VB;      // loads of the second block in flight per batch (interior segments; 16 measured 3 % slower)
template <bool ERODE>
__global__ void __launch_bounds__(256) bitmorph_v_vh_kernel(BitPlane src, BitPlane dst, int width, int height, int nw, int nseg,
                                                            unsigned per_image, int k, int anchor)
{
    extern __shared__ uint32_t sm[];
    const int T = blockDim.x, tid = threadIdx.x;
    const unsigned id = blockIdx.x * (unsigned)T + tid;
    if (id >= per_image) return;
    const int seg = (int)(id / (unsigned)nw);
    const int w = (int)(id - (unsigned)seg * (unsigned)nw);
    const int img = blockIdx.y;
    if (src.dims) {                                                   // ragged batch: this image may be smaller than the canvas
        const int2 d = src.dims[img];
        width = d.x; height = d.y; nw = (d.x + 31) >> 5;
        if (w >= nw || seg * k >= height) return;                     // no block-level synchronisation below
    }
    const uint32_t last_mask = (width & 31) ? ((1u << (width & 31)) - 1u) : 0xffffffffu;
    const uint32_t vmask = (w == nw - 1) ? last_mask : 0xffffffffu;
    const uint32_t *sp = src.p + img * src.bs + w;
    uint32_t *dp = dst.p + img * dst.bs + w;
    const int ys = seg * k, u0 = ys - anchor;
    const int swpr = src.wpr, dwpr = dst.wpr;
    uint32_t *col = sm + tid;                                         // private column: row j at col[j * T]
    if (u0 >= 0 && u0 + 2 * k - 1 <= height && ys + k <= height) {
        // ---- interior segment ----
        const uint32_t *p = sp + (int64_t)u0 * swpr;
        {
            uint32_t *c = col;
#pragma unroll 4
            for (int j = 0; j < k; ++j) { cp_async4(c, p); c += T; p += swpr; }
        }
        cp_async_commit();
        cp_async_wait<0>();
        {   // suffix ORs in place (erode: on the complemented rows)
            uint32_t acc = 0;
            uint32_t *c = col + (k - 1) * T;
#pragma unroll 4
            for (int j = k - 1; j >= 0; --j) {
                const uint32_t v = *c;
                acc |= ERODE ? (~v & vmask) : v;
                *c = acc;
                c -= T;
            }
        }
        // p = row u0 + k, the first row of the second block; output row ys + i = suffix[i] | OR(second block rows 0 .. i - 1)
        uint32_t *o = dp + (int64_t)ys * dwpr;
        const uint32_t *c = col;
        {
            const uint32_t v0 = *c;
            *o = ERODE ? (~v0 & vmask) : v0;
            o += dwpr; c += T;
        }
        uint32_t g = 0;
        int i = 1;
        for (; i + VB <= k; i += VB) {
            uint32_t t[VB];
#pragma unroll
            for (int q = 0; q < VB; ++q) t[q] = __ldg(p + q * swpr);
            p += VB * swpr;
#pragma unroll
            for (int q = 0; q < VB; ++q) {
                g |= ERODE ? (~t[q] & vmask) : t[q];
                const uint32_t v = c[q * T] | g;
                o[q * dwpr] = ERODE ? (~v & vmask) : v;
            }
            c += VB * T; o += VB * dwpr;
        }
        for (; i < k; ++i) {
            const uint32_t t = __ldg(p);
            g |= ERODE ? (~t & vmask) : t;
            const uint32_t v = *c | g;
            *o = ERODE ? (~v & vmask) : v;
            p += swpr; c += T; o += dwpr;
        }
        return;
    }
    // ---- segment touching the top or bottom border: rows outside the image hold the identity of the raw domain ----
    const uint32_t ident_raw = ERODE ? 0xffffffffu : 0u;
    for (int j = 0; j < k; ++j) {
        const int u = u0 + j;
        if (u >= 0 && u < height) cp_async4(&col[j * T], sp + (int64_t)u * swpr);
        else col[j * T] = ident_raw;
    }
    cp_async_commit();
    cp_async_wait<0>();
    {
        uint32_t acc = 0;
        for (int j = k - 1; j >= 0; --j) {
            uint32_t v = col[j * T];
            if (ERODE) v = ~v & vmask;
            acc |= v;
            col[j * T] = acc;
        }
    }
    auto ld = [&](int u) -> uint32_t {
        if (u < 0 || u >= height) return 0u;
        const uint32_t v = __ldg(sp + (int64_t)u * swpr);
        return ERODE ? (~v & vmask) : v;
    };
    uint32_t g = 0;
    const int n_out = min(k, height - ys);
    for (int i0 = 0; i0 < n_out; i0 += 8) {            // eight loads of the second block in flight per batch (a load per step would make
        uint32_t t[8];                                 // the two border segments of a column the critical path of the whole launch)
#pragma unroll
        for (int q = 0; q < 8; ++q) t[q] = (i0 + q > 0 && i0 + q < n_out) ? ld(u0 + k + i0 + q - 1) : 0u;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int i = i0 + q;
            if (i < n_out) {
                g |= t[q];
                uint32_t v = col[i * T] | g;
                if (ERODE) v = ~v;
                dp[(int64_t)(ys + i) * dwpr] = v & vmask;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// generic 8-bit passes
// ---------------------------------------------------------------------------------------------
template <bool DIL> __device__ __forceinline__ uint32_t vop4(uint32_t a, uint32_t b) { return DIL ? __vmaxu4(a, b) : __vminu4(a, b); }

constexpr int U8H_TW = 256;   // output bytes per tile row
constexpr int U8H_ROWS = 32;  // rows per tile (one lane each)

template <bool DIL>
__global__ void __launch_bounds__(32) morph_u8_h_kernel(Plane src, Plane dst, int width, int height, int k, int anchor, int pitch_w)
{
    extern __shared__ uint32_t sm[];
    const int lane = threadIdx.x;
    const int x0 = blockIdx.x * U8H_TW;
    const int64_t row0 = (int64_t)blockIdx.y * U8H_ROWS;   // global row (over batch * height)
    const int n = U8H_TW + k - 1;                          // input bytes per row: x0 - anchor .. + n - 1
    const uint32_t ident = DIL ? 0u : 0xffffffffu;
    uint8_t *smb = (uint8_t *)sm;
    // cooperative load: each warp iteration loads one tile row
    for (int rr = 0; rr < U8H_ROWS; ++rr) {
        const int64_t grow = row0 + rr;
        uint8_t *srow = smb + (size_t)rr * pitch_w * 4;
        const bool rvalid = grow < (int64_t)height;
        const uint8_t *gp = src.p + grow * src.rs;
        for (int j = lane; j < pitch_w * 4; j += 32) {
            const int x = x0 - anchor + j;
            uint8_t v = (uint8_t)ident;
            if (rvalid && j < n && x >= 0 && x < width) v = __ldg(gp + x);
            srow[j] = v;
        }
    }
    __syncwarp();
    uint32_t *myrow = sm + (size_t)lane * pitch_w;
    const int nwords = (n + 3) / 4;
    int p = 1;
    while (2 * p <= k) {
        const int ws = p >> 2, bsft = (p & 3) * 8;
        for (int j = 0; j < nwords; ++j) {
            const int a = j + ws;
            uint32_t sh;
            if (a + 1 < pitch_w) sh = __funnelshift_r(myrow[a], myrow[a + 1], bsft);
            else if (a < pitch_w) sh = __funnelshift_r(myrow[a], ident, bsft);
            else sh = ident;
            myrow[j] = vop4<DIL>(myrow[j], sh);
        }
        p *= 2;
    }
    {
        const int s = k - p;
        const int ws = s >> 2, bsft = (s & 3) * 8;
        for (int j = 0; j < U8H_TW / 4; ++j) {
            const int a = j + ws;
            uint32_t sh;
            if (a + 1 < pitch_w) sh = __funnelshift_r(myrow[a], myrow[a + 1], bsft);
            else if (a < pitch_w) sh = __funnelshift_r(myrow[a], ident, bsft);
            else sh = ident;
            myrow[j] = vop4<DIL>(myrow[j], sh);
        }
    }
    __syncwarp();
    for (int rr = 0; rr < U8H_ROWS; ++rr) {
        const int64_t grow = row0 + rr;
        if (grow >= (int64_t)height) break;
        const uint8_t *srow = smb + (size_t)rr * pitch_w * 4;
        uint8_t *gp = dst.p + grow * dst.rs;
        for (int j = lane; j < U8H_TW; j += 32)
            if (x0 + j < width) gp[x0 + j] = srow[j];
    }
}

constexpr int U8V_TH = 64;

template <bool DIL>
__global__ void __launch_bounds__(32) morph_u8_v_kernel(Plane src, Plane dst, int width, int height, int k, int anchor)
{
    extern __shared__ uint32_t sm[];
    const int lane = threadIdx.x;
    const int x0 = blockIdx.x * 128 + 4 * lane;
    const int y0 = blockIdx.y * U8V_TH;
    const int img = blockIdx.z;
    const int th = min(U8V_TH, height - y0);
    const int n = th + k - 1;
    const uint32_t ident = DIL ? 0u : 0xffffffffu;
    const uint8_t *base = src.p + img * src.bs;
    for (int j = 0; j < n; ++j) {
        const int y = y0 - anchor + j;
        uint32_t v = ident;
        if (y >= 0 && y < height && x0 < width) {
            const uint8_t *gp = base + y * src.rs + x0;
            v = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                uint32_t byte = (x0 + b < width) ? (uint32_t)__ldg(gp + b) : (ident & 255u);
                v |= byte << (8 * b);
            }
        }
        sm[j * 32 + lane] = v;
    }
    int p = 1;
    while (2 * p <= k) {
        for (int j = 0; j + p < n; ++j) sm[j * 32 + lane] = vop4<DIL>(sm[j * 32 + lane], sm[(j + p) * 32 + lane]);
        p *= 2;
    }
    if (x0 >= width) return;
    uint8_t *dbase = dst.p + img * dst.bs;
    for (int j = 0; j < th; ++j) {
        const uint32_t v = vop4<DIL>(sm[j * 32 + lane], sm[(j + k - p) * 32 + lane]);
        uint8_t *gp = dbase + (y0 + j) * dst.rs + x0;
#pragma unroll
        for (int b = 0; b < 4; ++b)
            if (x0 + b < width) gp[b] = (uint8_t)(v >> (8 * b));
    }
}

}  // namespace

// -------------------------------------------------------------------------------------------------
// launchers
// -------------------------------------------------------------------------------------------------
int launch_pack_bits(synseg_ctx *ctx, const synseg_img *src, BitPlane dst, cudaStream_t st)
{
    const int nw = cdiv(src->width, 32);
    const int64_t total = (int64_t)nw * src->height * src->batch;
    pack_bits_kernel<<<(unsigned)cdiv(total, 256), 256, 0, st>>>(plane_of(src), dst, src->width, src->height, nw, total, plane_aligned(src, 16));
    SS_LAUNCH_CHECK(ctx, "pack_bits", st);
    return SYNSEG_OK;
}

int launch_unpack_bits(synseg_ctx *ctx, BitPlane src, const synseg_img *dst, cudaStream_t st)
{
    const int nw = cdiv(dst->width, 32);
    const int64_t total = (int64_t)nw * dst->height * dst->batch;
    unpack_bits_kernel<<<(unsigned)cdiv(total, 256), 256, 0, st>>>(src, plane_of(dst), dst->width, dst->height, nw, total, plane_aligned(dst, 16));
    SS_LAUNCH_CHECK(ctx, "unpack_bits", st);
    return SYNSEG_OK;
}

int launch_count_bits(synseg_ctx *ctx, BitPlane src, int width, int height, int batch, uint64_t *out, int out_stride, cudaStream_t st)
{
    const int nw = cdiv(width, 32);
    int gx = cdiv((int64_t)nw * height, 256 * 8);
    if (gx < 1) gx = 1;
    count_bits_kernel<<<dim3(gx, batch), 256, 0, st>>>(src, nw, height, (unsigned long long *)out, out_stride);
    SS_LAUNCH_CHECK(ctx, "count_bits", st);
    return SYNSEG_OK;
}

int launch_bitmorph_h(synseg_ctx *ctx, BitPlane src, BitPlane dst, int width, int height, int batch, int op, int k, int anchor,
                      cudaStream_t st)
{
    const int nw = cdiv(width, 32);
    if (k <= 226 && src.p != dst.p) {                      // register kernel (not in place: threads read their neighbours' words)
        const bool wide = k <= 98 && !getenv("SYNSEG_MORPH_H4");          // 8 output words per thread
        const int nq = wide ? cdiv(src.wpr, 8) : src.wpr / 4;
        if (batch > 65535 || (int64_t)nq * height > 0x7fffffffLL) { synseg_set_error("bitmorph_h: batch > 65535 or image too large"); return SYNSEG_E_INVALID; }
        const unsigned per_image = (unsigned)nq * (unsigned)height;
        const dim3 grid((unsigned)cdiv(per_image, 256), (unsigned)batch);
        if (wide) bitmorph_h4_kernel<12, 8><<<grid, 256, 0, st>>>(src, dst, width, height, nw, nq, per_image, op == SYNSEG_MORPH_ERODE, k, anchor);
        else if (k <= 98) bitmorph_h4_kernel<8, 4><<<grid, 256, 0, st>>>(src, dst, width, height, nw, nq, per_image, op == SYNSEG_MORPH_ERODE, k, anchor);
        else bitmorph_h4_kernel<12, 4><<<grid, 256, 0, st>>>(src, dst, width, height, nw, nq, per_image, op == SYNSEG_MORPH_ERODE, k, anchor);
        SS_LAUNCH_CHECK(ctx, "bitmorph_h", st);
        return SYNSEG_OK;
    }
    if (src.dims) { synseg_set_error("bitmorph_h: ragged batches need k <= 226 and distinct planes"); return SYNSEG_E_INVALID; }
    const int padw = cdiv(k, 32) + 1;
    const size_t smem = (size_t)8 * 2 * (nw + 2 * padw + 1) * sizeof(uint32_t);
    if (smem > 200 * 1024) { synseg_set_error("bitmorph_h: row too wide for shared memory"); return SYNSEG_E_INVALID; }
    if (!(ctx->attr_done & ATTR_BITMORPH_H)) {     // the attribute is per device: remembered per context, not per process
        SS_CUDA(cudaFuncSetAttribute(bitmorph_h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); ctx->attr_done |= ATTR_BITMORPH_H;
    }
    const int64_t rows = (int64_t)height * batch;
    int grid = (int)((rows + 7) / 8);
    const int maxg = ctx->sm_count * 8;
    if (grid > maxg) grid = maxg;
    bitmorph_h_kernel<<<grid, 256, smem, st>>>(src, dst, width, height, batch, nw, op == SYNSEG_MORPH_ERODE, k, anchor, padw);
    SS_LAUNCH_CHECK(ctx, "bitmorph_h", st);
    return SYNSEG_OK;
}

int launch_bitmorph_v(synseg_ctx *ctx, BitPlane src, BitPlane dst, int width, int height, int batch, int op, int k, int anchor,
                      cudaStream_t st)
{
    const int nw = cdiv(width, 32);
    if (k <= 384) {                                        // van Herk kernel, k words of shared memory per thread
        // 128 threads per CTA: k words of shared memory per thread limit the CTAs per SM (k = 81: 2 CTAs of 256 threads = 512 threads,
        // 5 CTAs of 128 = 640); measured 0.148 -> 0.127 ms per 50 pages for the two column passes of the page pipeline (64: 0.128)
        int T = 128;
        if (const char *e = getenv("SYNSEG_MORPH_VT")) { const int v = atoi(e); if (v == 64 || v == 128 || v == 256) T = v; }
        const int nseg = cdiv(height, k);
        if (batch > 65535 || (int64_t)nw * nseg > 0x7fffffffLL) { synseg_set_error("bitmorph_v: batch > 65535 or image too large"); return SYNSEG_E_INVALID; }
        const unsigned per_image = (unsigned)nw * (unsigned)nseg;
        if (!(ctx->attr_done & ATTR_BITMORPH_VH)) {
            SS_CUDA(cudaFuncSetAttribute(bitmorph_v_vh_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            SS_CUDA(cudaFuncSetAttribute(bitmorph_v_vh_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            ctx->attr_done |= ATTR_BITMORPH_VH;
        }
        const dim3 grid((unsigned)cdiv(per_image, T), (unsigned)batch);
        const size_t smem = (size_t)k * T * sizeof(uint32_t);
        if (op == SYNSEG_MORPH_ERODE) bitmorph_v_vh_kernel<true><<<grid, T, smem, st>>>(src, dst, width, height, nw, nseg, per_image, k, anchor);
        else bitmorph_v_vh_kernel<false><<<grid, T, smem, st>>>(src, dst, width, height, nw, nseg, per_image, k, anchor);
        SS_LAUNCH_CHECK(ctx, "bitmorph_v", st);
        return SYNSEG_OK;
    }
    if (src.dims) { synseg_set_error("bitmorph_v: ragged batches need k <= 384"); return SYNSEG_E_INVALID; }
    const size_t smem = (size_t)(BV_TH + k - 1) * 32 * sizeof(uint32_t);
    if (smem > 200 * 1024) { synseg_set_error("bitmorph_v: kernel height %d too large", k); return SYNSEG_E_INVALID; }
    if (!(ctx->attr_done & ATTR_BITMORPH_V)) {
        SS_CUDA(cudaFuncSetAttribute(bitmorph_v_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); ctx->attr_done |= ATTR_BITMORPH_V;
    }
    dim3 grid(cdiv(nw, 32), cdiv(height, BV_TH), batch);
    bitmorph_v_kernel<<<grid, 32, smem, st>>>(src, dst, width, height, nw, op == SYNSEG_MORPH_ERODE, k, anchor);
    SS_LAUNCH_CHECK(ctx, "bitmorph_v", st);
    return SYNSEG_OK;
}

// One erode/dilate with a kw x kh rect on bit planes; the result ends up in `cur` (planes swapped as needed).
static int bit_rect_pass(synseg_ctx *ctx, BitPlane &cur, BitPlane &other, int width, int height, int batch, int op, int kw, int kh,
                         int ax, int ay, cudaStream_t st)
{
    if (kw > 1) {
        SS_TRY(launch_bitmorph_h(ctx, cur, other, width, height, batch, op, kw, ax, st));
        BitPlane t = cur; cur = other; other = t;
    }
    if (kh > 1) {
        SS_TRY(launch_bitmorph_v(ctx, cur, other, width, height, batch, op, kh, ay, st));
        BitPlane t = cur; cur = other; other = t;
    }
    return SYNSEG_OK;
}

int run_bitmorph(synseg_ctx *ctx, BitPlane &cur, BitPlane &other, int width, int height, int batch, int op, int kw, int kh, int ax,
                 int ay, int iterations, cudaStream_t st)
{
    const int fkw = iterations * (kw - 1) + 1, fkh = iterations * (kh - 1) + 1;
    const int fax = iterations * ax, fay = iterations * ay;
    const int first = (op == SYNSEG_MORPH_OPEN || op == SYNSEG_MORPH_ERODE) ? SYNSEG_MORPH_ERODE : SYNSEG_MORPH_DILATE;
    SS_TRY(bit_rect_pass(ctx, cur, other, width, height, batch, first, fkw, fkh, fax, fay, st));
    if (op == SYNSEG_MORPH_OPEN || op == SYNSEG_MORPH_CLOSE)
        SS_TRY(bit_rect_pass(ctx, cur, other, width, height, batch, 1 - first, fkw, fkh, fax, fay, st));
    return SYNSEG_OK;
}

int launch_morph_u8_h(synseg_ctx *ctx, const synseg_img *src, const synseg_img *dst, int op, int k, int anchor, cudaStream_t st)
{
    const int n = U8H_TW + k - 1;
    int pitch_w = (n + 3) / 4 + 1;
    pitch_w |= 1;
    const size_t smem = (size_t)U8H_ROWS * pitch_w * 4;
    if (smem > 200 * 1024) { synseg_set_error("morph: kernel width %d too large", k); return SYNSEG_E_INVALID; }
    if (!(ctx->attr_done & ATTR_MORPH_U8_H)) {
        SS_CUDA(cudaFuncSetAttribute(morph_u8_h_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        SS_CUDA(cudaFuncSetAttribute(morph_u8_h_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        ctx->attr_done |= ATTR_MORPH_U8_H;
    }
    for (int b = 0; b < src->batch; ++b) {
        Plane s = plane_of(src), d = plane_of(dst);
        s.p += b * s.bs; d.p += b * d.bs;
        dim3 grid(cdiv(src->width, U8H_TW), cdiv(src->height, U8H_ROWS), 1);
        if (op == SYNSEG_MORPH_DILATE) morph_u8_h_kernel<true><<<grid, 32, smem, st>>>(s, d, src->width, src->height, k, anchor, pitch_w);
        else morph_u8_h_kernel<false><<<grid, 32, smem, st>>>(s, d, src->width, src->height, k, anchor, pitch_w);
        SS_LAUNCH_CHECK(ctx, "morph_u8_h", st);
    }
    return SYNSEG_OK;
}

int launch_morph_u8_v(synseg_ctx *ctx, const synseg_img *src, const synseg_img *dst, int op, int k, int anchor, cudaStream_t st)
{
    const size_t smem = (size_t)(U8V_TH + k - 1) * 32 * 4;
    if (smem > 200 * 1024) { synseg_set_error("morph: kernel height %d too large", k); return SYNSEG_E_INVALID; }
    if (!(ctx->attr_done & ATTR_MORPH_U8_V)) {
        SS_CUDA(cudaFuncSetAttribute(morph_u8_v_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        SS_CUDA(cudaFuncSetAttribute(morph_u8_v_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        ctx->attr_done |= ATTR_MORPH_U8_V;
    }
    dim3 grid(cdiv(src->width, 128), cdiv(src->height, U8V_TH), src->batch);
    if (op == SYNSEG_MORPH_DILATE) morph_u8_v_kernel<true><<<grid, 32, smem, st>>>(plane_of(src), plane_of(dst), src->width, src->height, k, anchor);
    else morph_u8_v_kernel<false><<<grid, 32, smem, st>>>(plane_of(src), plane_of(dst), src->width, src->height, k, anchor);
    SS_LAUNCH_CHECK(ctx, "morph_u8_v", st);
    return SYNSEG_OK;
}

extern "C" SYNSEG_EXPORT int synseg_morph(synseg_ctx *ctx, const synseg_img *src, const synseg_img *dst, int op, int kw, int kh, int ax, int ay,
                            int iterations, int flags, void *stream)
{
    if (!ctx) { synseg_set_error("synseg_morph: ctx is NULL"); return SYNSEG_E_INVALID; }
    SS_ENTER(ctx, stream);
    SS_TRY(validate_img(src, "src", 1));
    SS_TRY(validate_img(dst, "dst", 1));
    if (!same_shape(src, dst)) { synseg_set_error("synseg_morph: shape mismatch"); return SYNSEG_E_INVALID; }
    if (op < 0 || op > 3) { synseg_set_error("synseg_morph: bad op %d", op); return SYNSEG_E_INVALID; }
    if (kw < 1 || kh < 1 || iterations < 1) { synseg_set_error("synseg_morph: bad kernel %dx%d it=%d", kw, kh, iterations); return SYNSEG_E_INVALID; }
    if (ax < 0) ax = kw / 2;
    if (ay < 0) ay = kh / 2;
    if (ax >= kw || ay >= kh) { synseg_set_error("synseg_morph: anchor outside kernel"); return SYNSEG_E_INVALID; }
    if (src->batch > 65535) { synseg_set_error("synseg_morph: batch > 65535"); return SYNSEG_E_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    const int W = src->width, H = src->height, B = src->batch;
    if (flags & SYNSEG_MORPH_BINARY) {
        const int wpr = bit_wpr(W);
        const size_t plane_bytes = (size_t)wpr * H * B * 4;
        SS_TRY(arena_ensure(ctx, 2 * plane_bytes + 1024));
        arena_begin(ctx);
        void *p0, *p1;
        SS_TRY(arena_alloc(ctx, plane_bytes, &p0, st));
        SS_TRY(arena_alloc(ctx, plane_bytes, &p1, st));
        BitPlane cur{(uint32_t *)p0, wpr, (int64_t)wpr * H}, other{(uint32_t *)p1, wpr, (int64_t)wpr * H};
        SS_TRY(launch_pack_bits(ctx, src, cur, st));
        SS_TRY(run_bitmorph(ctx, cur, other, W, H, B, op, kw, kh, ax, ay, iterations, st));
        return launch_unpack_bits(ctx, cur, dst, st);
    }
    // generic grey path: passes ping-pong between dst and a scratch plane so that src is never written
    if (src->data == dst->data) { synseg_set_error("synseg_morph: in-place operation is not supported on the grey path"); return SYNSEG_E_INVALID; }
    const int fkw = iterations * (kw - 1) + 1, fkh = iterations * (kh - 1) + 1;
    const int fax = iterations * ax, fay = iterations * ay;
    synseg_img tmp = *src;
    tmp.row_stride = (int64_t)align_up((size_t)W, 16);
    tmp.batch_stride = tmp.row_stride * H;
    SS_TRY(arena_ensure(ctx, (size_t)tmp.batch_stride * B + 1024));
    arena_begin(ctx);
    void *tp;
    SS_TRY(arena_alloc(ctx, (size_t)tmp.batch_stride * B, &tp, st));
    tmp.data = tp;
    const int first = (op == SYNSEG_MORPH_OPEN || op == SYNSEG_MORPH_ERODE) ? SYNSEG_MORPH_ERODE : SYNSEG_MORPH_DILATE;
    const int nphase = (op >= 2) ? 2 : 1;
    // list of 1-D passes
    struct Pass { int horiz, op, k, a; } passes[4];
    int np = 0;
    for (int ph = 0; ph < nphase; ++ph) {
        const int o = ph == 0 ? first : 1 - first;
        if (fkw > 1) passes[np++] = Pass{1, o, fkw, fax};
        if (fkh > 1) passes[np++] = Pass{0, o, fkh, fay};
    }
    if (np == 0) {   // 1x1 element: copy
        for (int b = 0; b < B; ++b)
            SS_CUDA(cudaMemcpy2DAsync((uint8_t *)dst->data + b * dst->batch_stride, dst->row_stride,
                                      (const uint8_t *)src->data + b * src->batch_stride, src->row_stride, W, H,
                                      cudaMemcpyDeviceToDevice, st));
        return SYNSEG_OK;
    }
    // choose the ping-pong parity so the last pass lands in dst
    const synseg_img *in = src;
    for (int i = 0; i < np; ++i) {
        const bool to_dst = ((np - 1 - i) % 2) == 0;
        const synseg_img *out = to_dst ? dst : &tmp;
        if (passes[i].horiz) SS_TRY(launch_morph_u8_h(ctx, in, out, passes[i].op, passes[i].k, passes[i].a, st));
        else SS_TRY(launch_morph_u8_v(ctx, in, out, passes[i].op, passes[i].k, passes[i].a, st));
        in = out;
    }
    return SYNSEG_OK;
}
