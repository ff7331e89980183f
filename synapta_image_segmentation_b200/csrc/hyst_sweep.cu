// hyst_sweep.cu -- Canny hysteresis by bit-parallel propagation sweeps, in front of the union-find of ccl.cu.
//
// cv2.Canny keeps every pixel that survived non-maximum suppression ("kept") and is 8-connected, through kept pixels, to a
// strong one (pdf_image_segmentation.py:1324 et al.).  The union-find of ccl.cu decides that for any
// image in a fixed number of passes, but it pays its full price (init, unions, compression, flagging, final pass: 0.36 ms per
// 50 pages, 17 % of a step) even on pages where nearly every kept pixel is strong itself or touches a strong one -- print.
// Here the strong plane GROWS in place: a CTA owns a full-width band of TR rows, loads kept and strong with a halo row above
// and below into shared memory and repeats   E <- kept & dilate3x3(E), then flood E along the runs of `kept` inside every word
// (carry trick: (m & ~(m + s)) | s fills upwards from the seeds, the bit-reversed form downwards)   until the band stops
// changing (at most MAX_IT rounds), ORs its rows into the output plane and reports whether anything changed.  Bands only see
// their neighbours through the halo rows, so the sweep is launched up to S times; a sweep in which no band changed anything
// proves the global fixed point (later sweeps return at once).  If the last sweep still changed something (long weak chains
// across many bands: spirals, faint rules) the union-find runs after all, seeded with the grown strong plane -- its kernels
// test the same flag and return at once otherwise.  Bits only ever get set and every set bit is a pixel of the final result,
// so reading a halo row while its owner updates it is harmless.
#include "internal.cuh"
#include "pixel.cuh"

namespace {

#ifndef SYNSEG_HS_TR
#define SYNSEG_HS_TR 32
#endif
#ifndef SYNSEG_HS_MAX_IT
#define SYNSEG_HS_MAX_IT 24
#endif
constexpr int HS_TR = SYNSEG_HS_TR;          // rows per band
constexpr int HS_ROWS_PER_THREAD = 8;
constexpr int HS_MAX_IT = SYNSEG_HS_MAX_IT;  // rounds inside a band per sweep
// Measured on B200, ms per 50-page step, text pages / dense pages (profiles/r2_tune_hysteresis_sweeps.txt); sweeps x rounds:
// 4 x 3: 1.984 / 2.85, 4 x 6: 1.987 / 2.72, 4 x 12: 1.922 / 2.74, 4 x 16: 1.918 / 2.60, 4 x 24: 1.915 / 2.63, 3 x 32: 1.908 / 2.67,
// 6 x 6: 1.940 / 2.70; 16-row bands 2.002.  Without the sweeps: 2.13 / 2.67.

__device__ __forceinline__ uint32_t flood_word(uint32_t m, uint32_t s)
{
    // s subset of m: extend the seeds along the runs of m in both directions (inside this word)
    const uint32_t up = (m & ~(m + s)) | s;
    const uint32_t mr = __brev(m), sr = __brev(s);
    const uint32_t dn = __brev((mr & ~(mr + sr)) | sr);
    return up | dn;
}

// sw[img][s] = number of bands of image img that changed in sweep s (zeroed by the launcher); sweep s > 0 returns at once when sw[img][s - 1] == 0
__global__ void __launch_bounds__(512) hyst_sweep_kernel(BitPlane kept, BitPlane strong, BitPlane out, int or_bits, int width, int height,
                                                         int32_t *sw_all, int n_sweeps, int sweep)
{
    int32_t *sw = sw_all + blockIdx.y * (n_sweeps + 1);       // one set of flags per image: a page of print is not held up by a photograph
    if (sweep > 0) {
        if (sw[sweep - 1] == 0) return;                        // converged
        // Hopeless: in the first sweep more than an eighth of the bands were still changing after HS_MAX_IT rounds (photographs,
        // noise: long weak chains everywhere).  Further sweeps would only delay the union-find: hand the image over now.
        if (sw[n_sweeps] * 8 > (int)gridDim.x) {
            if (blockIdx.x == 0 && threadIdx.x == 0) sw[n_sweeps - 1] = 1;
            return;
        }
    }
    extern __shared__ __align__(16) uint32_t sm[];
    const int pitch = kept.wpr;
    const int nw = (width + 31) >> 5;
    const int img = blockIdx.y, y0 = blockIdx.x * HS_TR;
    const int rows = min(HS_TR, height - y0);
    uint32_t *K = sm, *E = sm + (HS_TR + 2) * pitch;           // row r of the band lives at index r + 1; rows 0 and rows + 1 are the halo
    uint32_t *O = E + (HS_TR + 2) * pitch;                     // or_bits: the rows of the output plane the result is OR-ed into (row r at index r)
    const uint32_t *kp = kept.p + img * kept.bs, *sp = strong.p + img * strong.bs;
    const uint32_t *outp = out.p + img * out.bs;
    const int nq = pitch >> 2;
    // All loads of the band go out at once as 16-byte cp.async copies (no registers, L1 bypassed -- the strong plane grows while the
    // sweep runs), the rows of the output plane included: the kernel waits for memory once, not once per loop and again before the
    // final OR (60 % of the stall samples of the first version).  (r, q) advance without a division.
    const int dr = blockDim.x / nq, dq = blockDim.x - dr * nq;
    const int r_start = threadIdx.x / nq, q_start = threadIdx.x - r_start * nq;
    for (int r = r_start, q = q_start; r < HS_TR + 2; ) {
        const int y = y0 - 1 + r;
        if (y >= 0 && y < height && r <= rows + 1 && 4 * q < nw) {
            cp_async16(K + r * pitch + 4 * q, kp + (int64_t)y * pitch + 4 * q);
            cp_async16(E + r * pitch + 4 * q, sp + (int64_t)y * pitch + 4 * q);
            if (or_bits && r >= 1 && r <= rows) cp_async16(O + (r - 1) * pitch + 4 * q, outp + (int64_t)y * out.wpr + 4 * q);
        } else {
            *(uint4 *)(K + r * pitch + 4 * q) = make_uint4(0, 0, 0, 0);
            *(uint4 *)(E + r * pitch + 4 * q) = make_uint4(0, 0, 0, 0);
        }
        r += dr; q += dq;
        if (q >= nq) { q -= nq; ++r; }
    }
    cp_async_commit();
    cp_async_wait<0>();
    bool weak = false;
    for (int r = r_start, q = q_start; r < HS_TR + 2; ) {      // the same (r, q) pairs: every thread reads back its own copies
        const int y = y0 - 1 + r;
        if (y >= 0 && y < height && r <= rows + 1 && 4 * q < nw) {
            uint4 k = *(const uint4 *)(K + r * pitch + 4 * q), e = *(const uint4 *)(E + r * pitch + 4 * q);
            if (4 * q + 3 >= nw) {
                // words at or beyond nw are padding that nobody initialises: they must not leak into the neighbourhood of word nw - 1
                if (4 * q + 1 >= nw) { k.y = 0; e.y = 0; }
                if (4 * q + 2 >= nw) { k.z = 0; e.z = 0; }
                k.w = 0; e.w = 0;
                *(uint4 *)(K + r * pitch + 4 * q) = k;
                *(uint4 *)(E + r * pitch + 4 * q) = e;
            }
            if (r >= 1 && r <= rows) weak |= ((k.x & ~e.x) | (k.y & ~e.y) | (k.z & ~e.z) | (k.w & ~e.w)) != 0u;
        }
        r += dr; q += dq;
        if (q >= nq) { q -= nq; ++r; }
    }
    const bool any_weak = __syncthreads_or(weak);
    bool band_changed = false;
    if (any_weak) {
        // thread -> word column c, HS_ROWS_PER_THREAD consecutive rows, top-down (a change moves down inside the same round)
        const int c = threadIdx.x % pitch, grp = threadIdx.x / pitch;
        const int r_first = 1 + grp * HS_ROWS_PER_THREAD;
        const bool active = c < nw && grp * HS_ROWS_PER_THREAD < rows;
        for (int it = 0; it < HS_MAX_IT; ++it) {
            bool ch = false;
            if (active) {
                auto hrow = [&](int r) -> uint32_t {               // E of row r dilated by one pixel to the left and right, word c
                    const uint32_t *row = E + r * pitch;
                    const uint32_t m = row[c];
                    const uint32_t l = c > 0 ? row[c - 1] : 0u, rr = c + 1 < pitch ? row[c + 1] : 0u;
                    return m | (m << 1) | (m >> 1) | (l >> 31) | (rr << 31);
                };
                uint32_t h_up = hrow(r_first - 1), h_mid = hrow(r_first);
                const int r_end = min(r_first + HS_ROWS_PER_THREAD, rows + 1);
                for (int r = r_first; r < r_end; ++r) {
                    const uint32_t h_dn = hrow(r + 1);
                    const uint32_t k = K[r * pitch + c], e = E[r * pitch + c];
                    uint32_t ne = k & (h_up | h_mid | h_dn);
                    if (ne & ~e) {
                        ne = flood_word(k, ne | e);
                        E[r * pitch + c] = ne;
                        ch = true;
                        h_up = hrow(r);                             // with the new value
                    } else h_up = h_mid;
                    h_mid = h_dn;
                }
            }
            if (!__syncthreads_or(ch)) break;
            band_changed = true;
            if (it == HS_MAX_IT - 1 && sweep == 0 && threadIdx.x == 0) atomicAdd(&sw[n_sweeps], 1);      // did not settle inside the band
        }
    }
    // rows of this band -> strong plane (grown) and output plane
    for (int r = r_start, q = q_start; r < rows; ) {
        const uint4 e = *(const uint4 *)(E + (r + 1) * pitch + 4 * q);
        if (band_changed) *(uint4 *)(strong.p + img * strong.bs + (int64_t)(y0 + r) * pitch + 4 * q) = e;
        if (sweep == 0 || band_changed) {
            uint4 *op = (uint4 *)(out.p + img * out.bs + (int64_t)(y0 + r) * out.wpr) + q;
            uint4 o = e;
            if (or_bits && 4 * q < nw) {                       // copied by some thread of the CTA before the barrier above
                const uint4 old = *(const uint4 *)(O + r * pitch + 4 * q);
                o.x |= old.x; o.y |= old.y; o.z |= old.z; o.w |= old.w;
            }
            *op = o;
        }
        r += dr; q += dq;
        if (q >= nq) { q -= nq; ++r; }
    }
    if (band_changed && threadIdx.x == 0) atomicAdd(&sw[sweep], 1);
}

}  // namespace

// Up to n_sweeps propagation sweeps over (kept, strong); strong grows in place, `out` receives (or is OR-ed with) the result.
// sw: device int32[batch][n_sweeps + 1], zeroed here (per image: bands changed per sweep, then the bands that did not settle in the
// first sweep); afterwards sw[img][n_sweeps - 1] != 0 means "image img has not converged: run the union-find".
// Returns false in *used when the geometry does not fit (nothing launched).
int launch_hyst_sweeps(synseg_ctx *ctx, BitPlane kept, BitPlane strong, BitPlane out, bool or_bits, int width, int height, int batch,
                       int32_t *sw, int n_sweeps, bool *used, cudaStream_t st)
{
    *used = false;
    const int pitch = kept.wpr;
    if (kept.dims || strong.wpr != pitch || out.wpr != pitch || pitch > 128 || batch > 65535) return SYNSEG_OK;
    int threads = pitch * (HS_TR / HS_ROWS_PER_THREAD);
    threads = (threads + 31) & ~31;
    if (threads > 512) return SYNSEG_OK;
    if (threads < 64) threads = 64;
    SS_CUDA(cudaMemsetAsync(sw, 0, sizeof(int32_t) * (n_sweeps + 1) * (size_t)batch, st));
    const size_t smem = (size_t)(2 * (HS_TR + 2) + HS_TR) * pitch * sizeof(uint32_t);
    if (smem > 48 * 1024 && !(ctx->attr_done & ATTR_HYST_SWEEP)) {
        SS_CUDA(cudaFuncSetAttribute(hyst_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)); ctx->attr_done |= ATTR_HYST_SWEEP;
    }
    const dim3 grid(cdiv(height, HS_TR), batch);
    for (int s = 0; s < n_sweeps; ++s) {
        hyst_sweep_kernel<<<grid, threads, smem, st>>>(kept, strong, out, or_bits ? 1 : 0, width, height, sw, n_sweeps, s);
        SS_LAUNCH_CHECK(ctx, "hyst_sweep", st);
    }
    *used = true;
    return SYNSEG_OK;
}
