// Integer colour arithmetic shared by the grid and per-cell k-means kernels (internal header).
#pragma once
#include "ofc_common.cuh"

namespace ofc {

// cv2 BGR2HSV (8-bit, H range 180) hue of one pixel: integer table formula
__device__ __forceinline__ int hue_of_bgr(int b, int g, int r) {
    int v = max(max(b, g), r), mn = min(min(b, g), r);
    int d = v - mn;
    if (d == 0) return 0;
    int h = (v == r) ? (g - b) : ((v == g) ? (b - r + 2 * d) : (r - g + 4 * d));
    int hdiv = (int)rint((double)(180 << 12) / (6.0 * (double)d));
    h = (h * hdiv + (1 << 11)) >> 12;
    if (h < 0) h += 180;
    return h;
}

// np.rint(sum / n) in exact integer arithmetic (round half to even)
__device__ __forceinline__ unsigned rint_div(unsigned long long s, unsigned n) {
    unsigned long long q = s / n, r = s - q * n;
    if (2 * r > n || (2 * r == n && (q & 1))) ++q;
    return (unsigned)q;
}

}  // namespace ofc
