// pixel.cuh -- per-pixel fixed-point conversions shared by several kernels (bit-exact with cv2 / PIL).
#pragma once
#include <stdint.h>

#include "../../include/synseg.h"

template <int MODE> struct GrayCoef;
template <> struct GrayCoef<SYNSEG_GRAY_CV> {  // cv2 COLOR_RGB2GRAY: (9798 R + 19235 G + 3735 B + 16384) >> 15
    static constexpr uint32_t lo = (9798u & 255u) | ((19235u & 255u) << 8) | ((3735u & 255u) << 16);
    static constexpr uint32_t hi = (9798u >> 8) | ((19235u >> 8) << 8) | ((3735u >> 8) << 16);
    static constexpr uint32_t rnd = 16384u;
    static constexpr int sh = 15;
};
template <> struct GrayCoef<SYNSEG_GRAY_PIL> {  // PIL convert('L'): (19595 R + 38470 G + 7471 B + 32768) >> 16
    static constexpr uint32_t lo = (19595u & 255u) | ((38470u & 255u) << 8) | ((7471u & 255u) << 16);
    static constexpr uint32_t hi = (19595u >> 8) | ((38470u >> 8) << 8) | ((7471u >> 8) << 16);
    static constexpr uint32_t rnd = 32768u;
    static constexpr int sh = 16;
};

// rgbx: R | G<<8 | B<<16 | (ignored)<<24
template <int MODE>
__device__ __forceinline__ uint32_t gray1(uint32_t rgbx)
{
    using K = GrayCoef<MODE>;
    uint32_t l = __dp4a(rgbx, K::lo, K::rnd);
    uint32_t h = __dp4a(rgbx, K::hi, 0u);
    return (l + (h << 8)) >> K::sh;
}

__device__ __forceinline__ uint32_t gray_dyn(uint32_t rgbx, int mode)
{
    return mode == SYNSEG_GRAY_CV ? gray1<SYNSEG_GRAY_CV>(rgbx) : gray1<SYNSEG_GRAY_PIL>(rgbx);
}

// cv2 COLOR_RGB2HSV (8-bit) S and V: V = max, S = (diff * sdiv[V] + 2048) >> 12,
// sdiv[v] = rint((255 << 12) / v) (never a tie for v <= 255, so rint = floor((2*1044480 + v) / 2v)).
__device__ __forceinline__ uint32_t hsv_sdiv(uint32_t v) { return v ? (2u * 1044480u + v) / (2u * v) : 0u; }

// The reference's dominant-colour mask: S > 30 && V > 40 && V < 240 (pdf_image_segmentation.py:1574)
__device__ __forceinline__ bool hsv_mask_px(uint32_t r, uint32_t g, uint32_t b, const uint32_t *sdiv_tab)
{
    const uint32_t v = max(r, max(g, b)), mn = min(r, min(g, b)), diff = v - mn;
    // greys and near-greys need no table look-up: 9 diff < v  =>  diff * sdiv[v] + 2048 < 1044480 / 9 + 128 + 2048  =>  S <= 28
    if (9u * diff < v || v <= 40u || v >= 240u) return false;
    const uint32_t s = (diff * sdiv_tab[v] + 2048u) >> 12;
    return s > 30u;
}

// ---- 16 consecutive grey pixels of a row starting at column x (any x), BORDER_REPLICATE ------------------------
// Aligned rows (base and stride 16-byte aligned, x a multiple of 16) are read with exactly ONE 128-bit load per
// call for every lane: the address is clamped to the row's first / last 16-byte unit (clamp16_x) and the edge pixel
// is replicated afterwards with register arithmetic (fix16_rep), so the load itself never depends on loaded data
// and can be issued far ahead (registers or cp.async).  The vector path may read (never write) up to 15 bytes past
// column W-1 inside the row stride.
__device__ __forceinline__ int clamp16_x(int x, int W) { return min(max(x, 0), (W - 1) & ~15); }

// Border replication of one lane, prepared once per kernel (it depends on x and W only).
struct EdgeFix {
    bool interior;          // all 16 columns inside the image: nothing to do
    int sel_w, sel_sh;      // word / bit shift of the byte that is replicated
    uint32_t m0, m1, m2, m3; // bytes kept per word (0 everywhere for a lane wholly outside the image)
};

__device__ __forceinline__ EdgeFix make_edge_fix(int x, int W)
{
    EdgeFix f;
    const int xl = clamp16_x(x, W);
    const int nv = W - xl;                             // valid bytes from xl on (>= 1)
    f.interior = (x == xl && nv >= 16);
    const int bi = (x < 0) ? 0 : (min(nv, 16) - 1);    // byte that is replicated
    f.sel_w = bi >> 2; f.sel_sh = 8 * (bi & 3);
    if (x != xl) { f.m0 = f.m1 = f.m2 = f.m3 = 0u; }   // wholly outside the image
    else {
        const int k1 = nv - 4, k2 = nv - 8, k3 = nv - 12;
        f.m0 = nv >= 4 ? 0xFFFFFFFFu : ((1u << (8 * nv)) - 1u);
        f.m1 = k1 >= 4 ? 0xFFFFFFFFu : (k1 <= 0 ? 0u : ((1u << (8 * k1)) - 1u));
        f.m2 = k2 >= 4 ? 0xFFFFFFFFu : (k2 <= 0 ? 0u : ((1u << (8 * k2)) - 1u));
        f.m3 = k3 >= 4 ? 0xFFFFFFFFu : (k3 <= 0 ? 0u : ((1u << (8 * k3)) - 1u));
    }
    return f;
}

// v = the 16 bytes at column clamp16_x(x, W); returns the replicated-border pixels of columns x .. x+15
__device__ __forceinline__ uint4 apply_edge_fix(uint4 v, const EdgeFix &f)
{
    if (f.interior) return v;
    const uint32_t lw = f.sel_w == 0 ? v.x : (f.sel_w == 1 ? v.y : (f.sel_w == 2 ? v.z : v.w));
    const uint32_t rep = ((lw >> f.sel_sh) & 0xFFu) * 0x01010101u;
    v.x = (v.x & f.m0) | (rep & ~f.m0); v.y = (v.y & f.m1) | (rep & ~f.m1);
    v.z = (v.z & f.m2) | (rep & ~f.m2); v.w = (v.w & f.m3) | (rep & ~f.m3);
    return v;
}

__device__ __forceinline__ uint4 fix16_rep(uint4 v, int x, int W) { return apply_edge_fix(v, make_edge_fix(x, W)); }

// generic form: any alignment (byte loads on unaligned rows)
__device__ __forceinline__ uint4 load16_rep(const uint8_t *row, int x, int W, bool aligned)
{
    if (aligned) return fix16_rep(__ldg((const uint4 *)(row + clamp16_x(x, W))), x, W);
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        w[q] = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = min(max(x + 4 * q + j, 0), W - 1);
            w[q] |= (uint32_t)__ldg(row + c) << (8 * j);
        }
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// ---- cp.async (LDGSTS) helpers: 16-byte global -> shared copies that occupy no registers while in flight --------
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
