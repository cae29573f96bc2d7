// Grid-cell segmented reductions over the 8-bit flow visualisation.
//
// One CTA per (cell, frame).  Replaces the reference's Python cell loop
//   overlayGridAndComputeAvgColor  KmeanGrids.py:52-113, drawGridsAndOutputCSV.py:47-135
// and, for n_clusters == 1, the per-cell preprocess_image + KMeans(1) fit
//   KmeanGrids.py:269-339 / color_kmeans.py:35-135
// whose fitted centre is the column mean (integer-exact).
// Integer arithmetic throughout: results are bit-exact by construction.
#include "ofc_common.cuh"
#include "grid_kernels.cuh"
#include "color_math.cuh"

namespace ofc {

__global__ void __launch_bounds__(256) grid_cells_kernel(GridParams p) {
    const int cell = blockIdx.x, frame = blockIdx.y;
    const int cy = cell / p.cols, cx = cell - cy * p.cols;
    const int x1 = cx * p.x_step, y1 = cy * p.y_step;
    const int cw = min(x1 + p.x_step, p.W) - x1, chh = min(y1 + p.y_step, p.H) - y1;
    const unsigned char* img = p.bgr + (int64_t)frame * p.frame_stride;
    // state of the grid lines when the reference takes the cell mean (Q3): row 0
    // is white iff cy > 0, column 0 iff cx > 0; by k-means time both always are.
    const bool mean_row = p.draw_lines && cy > 0, mean_col = p.draw_lines && cx > 0;

    unsigned sa[3] = {0, 0, 0};          // sums for the mean stage
    unsigned sb[4] = {0, 0, 0, 0};       // sums for the k=1 stage (thresholded, + alpha count)
    const int n = cw * chh;
    auto add_pixel = [&](unsigned c0, unsigned c1, unsigned c2, int ly, int lx) {
        const bool wa = (ly == 0 && mean_row) || (lx == 0 && mean_col);
        sa[0] += wa ? 255u : c0; sa[1] += wa ? 255u : c1; sa[2] += wa ? 255u : c2;
        const bool wb = p.draw_lines && (ly == 0 || lx == 0);
        unsigned k0 = wb ? 255u : c0, k1 = wb ? 255u : c1, k2 = wb ? 255u : c2;
        if (p.threshold) {
            k0 = k0 < (unsigned)p.threshold ? 0u : k0;
            k1 = k1 < (unsigned)p.threshold ? 0u : k1;
            k2 = k2 < (unsigned)p.threshold ? 0u : k2;
        }
        // alpha = 255 * (BGR2GRAY(thresholded) > 0)
        const unsigned gray = (3735u * k0 + 19235u * k1 + 9798u * k2 + 16384u) >> 15;
        sb[0] += k0; sb[1] += k1; sb[2] += k2; sb[3] += gray > 0 ? 1u : 0u;
    };
    const bool quads = (cw & 3) == 0 && ((x1 * 3) & 3) == 0 && ((p.W * 3) & 3) == 0 && (p.frame_stride & 3) == 0 &&
                       ((uintptr_t)p.bgr & 3) == 0;
    if (quads) {
        // 4 pixels = 12 bytes = three aligned words per thread step
        const int qpr = cw >> 2, nq = qpr * chh;
        for (int i = threadIdx.x; i < nq; i += blockDim.x) {
            const int ly = i / qpr, lq = i - ly * qpr;
            const unsigned* wp = reinterpret_cast<const unsigned*>(img + ((int64_t)(y1 + ly) * p.W + x1 + lq * 4) * 3);
            const unsigned w0 = wp[0], w1 = wp[1], w2 = wp[2];
            add_pixel(w0 & 255u, (w0 >> 8) & 255u, (w0 >> 16) & 255u, ly, lq * 4);
            add_pixel(w0 >> 24, w1 & 255u, (w1 >> 8) & 255u, ly, lq * 4 + 1);
            add_pixel((w1 >> 16) & 255u, w1 >> 24, w2 & 255u, ly, lq * 4 + 2);
            add_pixel((w2 >> 8) & 255u, (w2 >> 16) & 255u, w2 >> 24, ly, lq * 4 + 3);
        }
    } else {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int ly = i / cw, lx = i - ly * cw;
            const unsigned char* px = img + ((int64_t)(y1 + ly) * p.W + (x1 + lx)) * 3;
            add_pixel(px[0], px[1], px[2], ly, lx);
        }
    }
    __shared__ unsigned s_red[8][7];
    unsigned vals[7] = {sa[0], sa[1], sa[2], sb[0], sb[1], sb[2], sb[3]};
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        unsigned v = vals[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        vals[k] = v;
    }
    if ((threadIdx.x & 31) == 0)
        for (int k = 0; k < 7; ++k) s_red[threadIdx.x >> 5][k] = vals[k];
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t[7] = {0, 0, 0, 0, 0, 0, 0};
        for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv)
            for (int k = 0; k < 7; ++k) t[k] += s_red[wv][k];
        const int64_t o = (int64_t)frame * gridDim.x + cell;
        if (n > 0) {
            unsigned a0 = t[0] / n, a1 = t[1] / n, a2 = t[2] / n;           // floor: .astype(uint8)
            if (p.avg_bgr) { p.avg_bgr[o * 3] = (unsigned char)a0; p.avg_bgr[o * 3 + 1] = (unsigned char)a1; p.avg_bgr[o * 3 + 2] = (unsigned char)a2; }
            if (p.avg_hue) p.avg_hue[o] = (unsigned char)hue_of_bgr((int)a0, (int)a1, (int)a2);
            unsigned k0 = rint_div(t[3], n), k1 = rint_div(t[4], n), k2 = rint_div(t[5], n);
            unsigned k3 = rint_div((unsigned long long)t[6] * 255ull, n);
            if (p.km_centre) { p.km_centre[o * 4] = (unsigned char)k0; p.km_centre[o * 4 + 1] = (unsigned char)k1; p.km_centre[o * 4 + 2] = (unsigned char)k2; p.km_centre[o * 4 + 3] = (unsigned char)k3; }
            if (p.km_hue) p.km_hue[o] = (unsigned char)hue_of_bgr((int)k0, (int)k1, (int)k2);
            if (p.km_sums) { p.km_sums[o * 4] = t[3]; p.km_sums[o * 4 + 1] = t[4]; p.km_sums[o * 4 + 2] = t[5]; p.km_sums[o * 4 + 3] = t[6] * 255u; }
        }
    }
}

// cv2.rectangle(frame,(x1,y1),(x2,y2),(255,255,255),1) for every cell: white
// rows at y = cy*y_step (cy = 0..rows) and columns at x = cx*x_step (cx = 0..cols),
// clipped to the grid extent and the image (KmeanGrids.py:108).
__global__ void __launch_bounds__(256) draw_grid_kernel(unsigned char* bgr, int64_t frame_stride, int W, int H,
                                                         int rows, int cols, int x_step, int y_step) {
    unsigned char* img = bgr + (int64_t)blockIdx.z * frame_stride;
    const int x_end = min(cols * x_step, W - 1), y_end = min(rows * y_step, H - 1);
    const int line = blockIdx.y;                       // 0..rows : horizontal, rows+1.. : vertical
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (line <= rows) {
        int y = line * y_step;
        if (y < H && t <= x_end) {
            unsigned char* px = img + ((int64_t)y * W + t) * 3;
            px[0] = px[1] = px[2] = 255;
        }
    } else {
        int x = (line - rows - 1) * x_step;
        if (x < W && t <= y_end) {
            unsigned char* px = img + ((int64_t)t * W + x) * 3;
            px[0] = px[1] = px[2] = 255;
        }
    }
}

// cv2.cvtColor(..., COLOR_BGR2HSV) on 8-bit pixels (H in 0..179): the integer table formula
// of SURVEY.md A.3 (hdiv/sdiv tables, +2048 >> 12).  Used on cell means and cluster centres
// (KmeanGrids.py:92,336, color_kmeans.py:121).
__global__ void __launch_bounds__(256) bgr2hsv_kernel(const unsigned char* bgr, unsigned char* hsv, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int b = bgr[i * 3], g = bgr[i * 3 + 1], r = bgr[i * 3 + 2];
    const int v = max(max(b, g), r), mn = min(min(b, g), r);
    const int d = v - mn;
    int s = 0;
    if (v > 0) {
        const int sdiv = (int)rint((double)(255 << 12) / (double)v);
        s = (d * sdiv + (1 << 11)) >> 12;
    }
    hsv[i * 3] = (unsigned char)hue_of_bgr(b, g, r);
    hsv[i * 3 + 1] = (unsigned char)s;
    hsv[i * 3 + 2] = (unsigned char)v;
}

int launch_bgr2hsv(const unsigned char* bgr, unsigned char* hsv, int64_t n, void* stream) {
    if (n <= 0) return OFC_OK;
    ProfScope prof(PK_GRID, stream);
    OFC_LAUNCH(bgr2hsv_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, stream, bgr, hsv, n);
    OFC_CHECK_LAUNCH("bgr2hsv");
    return OFC_OK;
}

int launch_grid_cells(const GridParams& p, int n_frames, void* stream) {
    if (n_frames <= 0) return OFC_OK;
    ProfScope prof(PK_GRID, stream);
    OFC_LAUNCH(grid_cells_kernel, dim3(p.rows * p.cols, n_frames), dim3(256), 0, stream, p);
    OFC_CHECK_LAUNCH("grid_cells");
    return OFC_OK;
}

int launch_draw_grid(unsigned char* bgr, int64_t frame_stride, int W, int H, int rows, int cols,
                     int x_step, int y_step, int n_frames, void* stream) {
    if (n_frames <= 0) return OFC_OK;
    int longest = max(W, H);
    dim3 grid(cdiv(longest, 256), rows + cols + 2, n_frames);
    ProfScope prof(PK_DRAW, stream);
    OFC_LAUNCH(draw_grid_kernel, grid, dim3(256), 0, stream, bgr, frame_stride, W, H, rows, cols, x_step, y_step);
    OFC_CHECK_LAUNCH("draw_grid");
    return OFC_OK;
}

}  // namespace ofc
