// morph_fused.cu -- dilate(K1 x K1) followed by erode(k2 x k2) on bit planes in ONE kernel, bit-exact with the four 1-D passes of
// morph.cu (and so with cv2.dilate / cv2.morphologyEx(MORPH_CLOSE)).
//
// The page pipeline applies dilate(k) and then close(k) = dilate(k), erode(k): a dilation with the folded (2k-1) element and an
// erosion with k.  As four launches (row / column pass each) the bit planes make four round trips through L2 and every pass
// is latency-bound (0.26 ms per 50 pages at 300 DPI for 106 MB of traffic per pass).  Here a CTA owns a full-width band of
// TH output rows; the band plus its halo of (K1 - 1) + (k2 - 1) rows is loaded once into shared memory and the four passes
// run there -- column dilate (van Herk / Gil-Werman per word column), row dilate, row erode, column erode -- in the order
// that puts the cheap column pass on the tall input and the row passes on the shorter intermediate.  A page is only 80
// words wide, so a band holds whole rows and the row passes need no horizontal halo at all.
// Erosion runs as dilation of the complement (outside the image the complement is 0 = "ignored", cv2's border rule).
//
// Roofline: 1/4 algorithmic byte per pixel (bit plane read + written once); the kernel is bound by shared-memory
// latency / issue, not by HBM (the planes of a 50-page batch stay in L2).
#include <stdlib.h>

#include "internal.cuh"

namespace {

constexpr int MF_THREADS = 512;

struct MfParams {
    BitPlane src, dst;
    int width, height, nw;      // pixels, rows, words per row that hold pixels
    int K1, a1, k2, a2;         // dilate element / anchor, erode element / anchor (per axis, square)
    int TH;                     // output rows per band
    int bands;
};

// Row dilation of 4 consecutive words: out[q] bit b = OR over i in [0, k) of in[32 (w0 + q) + b + i - anchor]; words outside
// [0, nw) read as zero.  In-place doubling on a register array (see morph.cu: bitmorph_h4_kernel).  NW = 8 serves k <= 98, 12 k <= 226.
template <int NW>
__device__ __forceinline__ void hdilate4(const uint32_t *row, int nw, int w0, int k, int anchor, uint32_t (&o)[4])
{
    const int aw = (anchor + 31) >> 5, sft = 32 * aw - anchor;
    uint32_t D[NW + 1];
#pragma unroll
    for (int j = 0; j < NW; ++j) {
        const int w = w0 - aw + j;
        D[j] = (w >= 0 && w < nw) ? row[w] : 0u;
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
        for (int j = 0; j < NW; ++j) D[j] |= D[j + 2 < NW ? j + 2 : NW];
        span *= 2;
    }
    if (NW > 8 && 2 * span <= k) {                         // 128
#pragma unroll
        for (int j = 0; j < NW; ++j) D[j] |= D[j + 4 < NW ? j + 4 : NW];
        span *= 2;
    }
    const int t = k - span, tw = t >> 5, tb = t & 31;      // E = D | (D >> (k - span)), words 0..4
    uint32_t E[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int m = 0; m < 4; ++m)
            if (tw == m) { lo = D[j + m < NW ? j + m : NW]; hi = D[j + m + 1 < NW ? j + m + 1 : NW]; }
        E[j] = D[j] | __funnelshift_r(lo, hi, tb);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) o[q] = __funnelshift_r(E[q], E[q + 1], sft);
}

// Column dilation inside shared memory, in place: A[r] = OR over j in [0, K) of A[r + j] for r < n_out (rows >= n_in read as
// zero; n_in >= n_out + K - 1 so every row an output needs exists).  Doubling: after the pass with shift s every row holds the OR
// of 2s consecutive rows; the last pass adds the row K - span below.  All threads work on 128-bit row pieces; a pass loads
// both operands of every piece a thread owns into registers, a barrier separates the loads from the stores (in place).
// Rows that cannot reach an output any more are left alone.
template <int EPT>
__device__ __forceinline__ void vdilate_inplace(uint32_t *A, int n_in, int n_out, int K, int nq, int pitch)
{
    int off[EPT], row[EPT];
#pragma unroll
    for (int i = 0; i < EPT; ++i) {
        const int e = threadIdx.x + i * MF_THREADS;
        row[i] = e / nq;
        off[i] = row[i] * pitch + 4 * (e - row[i] * nq);
    }
    int span = 1;
    int need = n_out + K - 1;                           // rows whose value can still reach an output
    for (;;) {
        const int sh = (2 * span <= K) ? span : K - span;          // doubling step, or the final combine
        if (sh == 0) break;
        need -= sh;
        uint4 v[EPT];
#pragma unroll
        for (int i = 0; i < EPT; ++i) {
            if (row[i] < need) {
                const uint4 a = *(const uint4 *)(A + off[i]);
                uint4 b = make_uint4(0, 0, 0, 0);
                if (row[i] + sh < n_in) b = *(const uint4 *)(A + off[i] + sh * pitch);
                v[i] = make_uint4(a.x | b.x, a.y | b.y, a.z | b.z, a.w | b.w);
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < EPT; ++i)
            if (row[i] < need) *(uint4 *)(A + off[i]) = v[i];
        __syncthreads();
        if (2 * span <= K) span *= 2; else break;
    }
}

template <int NW1, int NW2, int EPT>
__global__ void __launch_bounds__(MF_THREADS, 2) bit_dilate_erode_kernel(MfParams p)
{
    extern __shared__ uint32_t sm[];
    const int img = blockIdx.y;
    const int y0 = blockIdx.x * p.TH;
    const int th = min(p.TH, p.height - y0);
    const int pitch = p.src.wpr;                        // words per row in memory (multiple of 4, >= nw)
    const int nq = pitch >> 2;
    const int RB = th + p.k2 - 1;                       // rows of the dilated intermediate the erosion reads
    const int RA = RB + p.K1 - 1;                       // source rows the dilation reads
    const int yB0 = y0 - p.a2, yA0 = yB0 - p.a1;
    uint32_t *A = sm, *B = sm + (size_t)(p.TH + p.k2 + p.K1 - 2) * pitch;
    const uint32_t last_mask = (p.width & 31) ? ((1u << (p.width & 31)) - 1u) : 0xffffffffu;
    const uint32_t *sp = p.src.p + img * p.src.bs;

    // ---- load the band and its halo (rows outside the image are empty) --------------------------------------------
    for (int t = threadIdx.x; t < RA * nq; t += MF_THREADS) {
        const int r = t / nq, q = t - r * nq;
        const int y = yA0 + r;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (y >= 0 && y < p.height) v = __ldg((const uint4 *)(sp + (int64_t)y * pitch) + q);
        SS_DEVICE_ASSERT(r < p.TH + p.k2 + p.K1 - 2 && 4 * q + 4 <= pitch);
        *(uint4 *)(A + r * pitch + 4 * q) = v;
    }
    __syncthreads();
    // ---- 1. column dilate K1, in place: A rows [0, RB) ----------------------------------------------------------------
    vdilate_inplace<EPT>(A, RA, RB, p.K1, nq, pitch);
    // ---- 2. row dilate K1: A -> B, stored as the COMPLEMENT inside the image and 0 outside (the erosion's "ignore") ----
    for (int t = threadIdx.x; t < RB * nq; t += MF_THREADS) {
        const int r = t / nq, q = t - r * nq;
        const int y = yB0 + r;
        uint32_t o[4] = {0, 0, 0, 0};
        if (y >= 0 && y < p.height) {
            hdilate4<NW1>(A + r * pitch, p.nw, 4 * q, p.K1, p.a1, o);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int w = 4 * q + i;
                o[i] = ~o[i];
                if (w == p.nw - 1) o[i] &= last_mask;
                if (w >= p.nw) o[i] = 0;
            }
        }
        *(uint4 *)(B + r * pitch + 4 * q) = make_uint4(o[0], o[1], o[2], o[3]);
    }
    __syncthreads();
    // ---- 3. row dilate k2 of the complement: B -> A --------------------------------------------------------------
    for (int t = threadIdx.x; t < RB * nq; t += MF_THREADS) {
        const int r = t / nq, q = t - r * nq;
        uint32_t o[4];
        hdilate4<NW2>(B + r * pitch, p.nw, 4 * q, p.k2, p.a2, o);
        *(uint4 *)(A + r * pitch + 4 * q) = make_uint4(o[0], o[1], o[2], o[3]);
    }
    __syncthreads();
    // ---- 4. column dilate k2 of the complement, in place: A rows [0, th); complement back -> global -------------------
    vdilate_inplace<EPT>(A, RB, th, p.k2, nq, pitch);
    uint32_t *dp = p.dst.p + img * p.dst.bs + (int64_t)y0 * p.dst.wpr;
    for (int t = threadIdx.x; t < th * nq; t += MF_THREADS) {
        const int r = t / nq, q = t - r * nq;
        uint4 v = *(const uint4 *)(A + r * pitch + 4 * q);
        uint32_t o[4] = {~v.x, ~v.y, ~v.z, ~v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int w = 4 * q + i;
            if (w == p.nw - 1) o[i] &= last_mask;
            if (w >= p.nw) o[i] = 0;
        }
        *(uint4 *)(dp + (int64_t)r * p.dst.wpr + 4 * q) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

}  // namespace

// dst = erode(k2 x k2, anchor a2)(dilate(K1 x K1, anchor a1)(src)) on bit planes (src != dst).  Returns SYNSEG_OK and sets *done
// when the fused kernel took the job; *done = false (nothing launched) when the geometry does not fit it.
int launch_bit_dilate_erode(synseg_ctx *ctx, BitPlane src, BitPlane dst, int width, int height, int batch, int K1, int a1, int k2, int a2,
                            bool *done, cudaStream_t st)
{
    *done = false;
    // Opt-in (SYNSEG_FUSED_MORPH=1).  Measured on B200 (profiles/r2_front_end_experiments.txt): 0.262 ms per 50 pages against 0.263 ms for the four
    // separate passes -- both are bound by instruction issue (the doubling steps), not by the four L2 round trips the fusion saves --
    // and a step that runs two page chains on two streams is 2 % SLOWER with it (110 KB of shared memory per CTA leave the other
    // chain's kernels no room beside it).  Kept for geometries where launch count matters, parity-tested like the default path.
    const char *opt = getenv("SYNSEG_FUSED_MORPH");
    if (!opt || atoi(opt) == 0) return SYNSEG_OK;
    if (src.dims || src.p == dst.p || src.wpr != dst.wpr || K1 < 2 || k2 < 2 || K1 > 226 || k2 > 226 || batch > 65535) return SYNSEG_OK;
    const int pitch = src.wpr;
    const int halo = K1 + k2 - 2;
    // two buffers: (TH + halo) and (TH + k2 - 1) rows of `pitch` words; two CTAs per SM when it fits
    const size_t budget = 110 * 1024;
    const long rows2 = (long)(budget / (4 * (size_t)pitch)) - halo - (k2 - 1);
    int TH = (int)(rows2 / 2);
    if (TH > 256) TH = 256;
    if (const char *e = getenv("SYNSEG_MORPH_TH")) { const int v = atoi(e); if (v >= 32 && v < TH) TH = v; }
    TH &= ~7;
    if (TH < 32) return SYNSEG_OK;                      // wide images or huge elements: the separate passes
    if (TH > height) TH = (height + 7) & ~7;
    MfParams p;
    p.src = src; p.dst = dst; p.width = width; p.height = height; p.nw = cdiv(width, 32);
    p.K1 = K1; p.a1 = a1; p.k2 = k2; p.a2 = a2; p.TH = TH; p.bands = cdiv(height, TH);
    constexpr int EPT = 10;                             // 128-bit row pieces a thread holds in a column pass
    while (TH >= 32 && (long)(TH + halo) * (pitch / 4) > (long)EPT * MF_THREADS) TH -= 8;
    if (TH < 32) return SYNSEG_OK;
    p.TH = TH; p.bands = cdiv(height, TH);
    const size_t smem2 = (size_t)4 * pitch * ((size_t)(TH + halo) + (size_t)(TH + k2 - 1));
    auto kern = (K1 <= 98) ? ((k2 <= 98) ? bit_dilate_erode_kernel<8, 8, EPT> : bit_dilate_erode_kernel<8, 12, EPT>)
                           : ((k2 <= 98) ? bit_dilate_erode_kernel<12, 8, EPT> : bit_dilate_erode_kernel<12, 12, EPT>);
    SS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));     // per device and variant; cheap
    kern<<<dim3(p.bands, batch), MF_THREADS, smem2, st>>>(p);
    SS_LAUNCH_CHECK(ctx, "bit_dilate_erode", st);
    *done = true;
    return SYNSEG_OK;
}
