// gauss.cuh — the unit-peak Gaussian window `rectify` pastes at a decoded arg-max
// (utils.py:77-109 of the reference), shared by the decode kernel (which materialises the
// rectified map) and the fused loss step (which evaluates it on the fly instead of reading
// a materialised map).  Both paths call the same functions, so their values are identical.
#pragma once

#include <cmath>

#include "common.cuh"

namespace udape {

// window geometry, derived on the host from sigma exactly as utils.py:81,93-98 derives it
struct GaussWindow {
    float tmp;    // 3*sigma                       (utils.py:81)
    int n;        // len(arange(0, 2*tmp+1, 1))    (utils.py:93-94)
    float x0;     // (2*tmp+1) // 2                (utils.py:96)
    float denom;  // 2*sigma**2                    (utils.py:98)
};

inline GaussWindow make_window(double sigma) {
    GaussWindow g;
    const double tmp = 3.0 * sigma;
    const double size = 2.0 * tmp + 1.0;
    g.tmp = static_cast<float>(tmp);
    g.n = static_cast<int>(std::ceil(size));
    g.x0 = static_cast<float>(std::floor(size / 2.0));
    g.denom = static_cast<float>(2.0 * sigma * sigma);
    return g;
}

// Every plane of a launch pastes the SAME (6*sigma+1)^2 window (only its position differs), so a
// CTA tabulates it once in shared memory and the per-element work of the writers / the fused loss
// step becomes one LDS instead of a division and an expf.  The table entries are produced by
// window_value(), the function the direct route calls, so both routes give identical bits.
// Windows wider than kWinTabN (sigma > 3) are evaluated directly.
constexpr int kWinTabN = 19;

__device__ __forceinline__ float window_value(int gx, int gy, const GaussWindow& g) {
    const float dx = static_cast<float>(gx) - g.x0, dy = static_cast<float>(gy) - g.x0;
    const float d2 = dx * dx + dy * dy;
    return expf(-(d2 / g.denom));
}

// all threads of the CTA; the caller synchronises before the first lookup.  Returns the table
// pointer to pass to rectified_value (nullptr when the window is too wide to tabulate).
__device__ __forceinline__ const float* build_window_table(float* tab, const GaussWindow& g) {
    if (g.n > kWinTabN) return nullptr;
    for (int i = threadIdx.x; i < g.n * g.n; i += blockDim.x) {
        const int gy = i / g.n;
        tab[i] = window_value(i - gy * g.n, gy, g);
    }
    return tab;
}

// placement of the window inside one h x w plane, utils.py:84-107 (including its use of h
// for the x bound and w for the y bound)
struct RectGeom {
    int ul_x, ul_y;              // upper-left corner of the (unclipped) window
    int x0i, x1i, y0i, y1i;      // clipped image range that receives window values
};

__device__ __forceinline__ RectGeom rect_geometry(float mu_x, float mu_y, int h, int w,
                                                  const GaussWindow& g) {
    RectGeom r;
    r.ul_x = static_cast<int>(mu_x - g.tmp);
    r.ul_y = static_cast<int>(mu_y - g.tmp);
    const int br_x = static_cast<int>(mu_x + g.tmp + 1.0f), br_y = static_cast<int>(mu_y + g.tmp + 1.0f);
    const bool skip = (mu_x >= static_cast<float>(h)) || (mu_y >= static_cast<float>(w));  // utils.py:89
    r.x0i = max(0, r.ul_x);
    // the window holds g.n samples per axis (utils.py:101-102 slices g with these bounds)
    r.x1i = min(min(min(br_x, h), w), r.ul_x + g.n);
    r.y0i = max(0, r.ul_y);
    r.y1i = min(min(min(br_y, w), h), r.ul_y + g.n);
    if (skip) { r.x1i = r.x0i = 0; r.y1i = r.y0i = 0; }
    return r;
}

// zeros + clipped Gaussian window
__device__ __forceinline__ float rectified_value(int x, int y, const RectGeom& r, const GaussWindow& g,
                                                 const float* __restrict__ tab = nullptr) {
    if (x < r.x0i || x >= r.x1i || y < r.y0i || y >= r.y1i) return 0.0f;
    const int gx = x - r.ul_x, gy = y - r.ul_y;
    return tab ? tab[gy * g.n + gx] : window_value(gx, gy, g);
}

}  // namespace udape
