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
    const uint32_t v = max(r, max(g, b)), mn = min(r, min(g, b));
    const uint32_t s = ((v - mn) * sdiv_tab[v] + 2048u) >> 12;
    return s > 30u && v > 40u && v < 240u;
}

// 16 consecutive grey pixels of a row starting at column x (any x), replicated outside [0, W)
// (BORDER_REPLICATE).  `aligned`: row base and stride are 16-byte aligned (x must then be a multiple of 16).
__device__ __forceinline__ uint4 load16_rep(const uint8_t *row, int x, int W, bool aligned)
{
    if (aligned && x >= 0 && x + 15 < W) return __ldg((const uint4 *)(row + x));
    if (x + 15 < 0 || x >= W) {
        const uint32_t b = (uint32_t)__ldg(row + (x < 0 ? 0 : W - 1)) * 0x01010101u;
        return make_uint4(b, b, b, b);
    }
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
