// Launchers of grid_kernels.cu (internal header).
#pragma once
#include "ofc_common.cuh"

namespace ofc {

struct GridParams {
    const unsigned char* bgr;    // [n_frames][H][W][3]
    int64_t frame_stride;        // bytes between frames
    int W, H, rows, cols, x_step, y_step;
    int draw_lines;              // reproduce the reference's white-rectangle state
    int threshold;               // preprocess_image: channel < threshold -> 0 (reference: 30); 0 = off
    unsigned char* avg_bgr;      // [n_frames][cells][3] floor(mean)            or null
    unsigned char* avg_hue;      // [n_frames][cells]    BGR2HSV hue of avg_bgr or null
    unsigned char* km_centre;    // [n_frames][cells][4] rint(mean of c0,c1,c2,alpha) or null
    unsigned char* km_hue;       // [n_frames][cells]    hue of km_centre[0..2] or null
    unsigned* km_sums;           // [n_frames][cells][4] channel sums of the k-means input or null
};

int launch_grid_cells(const GridParams& p, int n_frames, void* stream);
int launch_bgr2hsv(const unsigned char* bgr, unsigned char* hsv, int64_t n, void* stream);
int launch_draw_grid(unsigned char* bgr, int64_t frame_stride, int W, int H, int rows, int cols,
                     int x_step, int y_step, int n_frames, void* stream);

}  // namespace ofc
