// Launchers of viz_kernels.cu (internal header).
#pragma once
#include "ofc_common.cuh"

namespace ofc {

struct VizParams {
    const float2* flow;        // [n_frames][n_px]
    int64_t n_px;              // H*W
    int width;                 // W (cv2's HSV2BGR treats the last W % 32 pixels of a row differently)
    const unsigned* minmax;    // [n_frames][2] float bits of min/max |flow|
    unsigned char* bgr;        // [n_frames][n_px][3]
    double* mag_sum;           // [n_frames] (zeroed by the caller) or null
    unsigned char* hsv;        // [n_frames][n_px][3] the reference's `mask` (H, 255, V) or null
};

int launch_bgr2gray(const unsigned char* bgr, unsigned char* gray, int64_t n_px, void* stream);
int launch_flow_encode(const VizParams& p, int n_frames, void* stream);
struct GridParams;
// visualisation + grid pass in one kernel (one CTA per cell and frame); gp.bgr is ignored (p.bgr is written)
int launch_flow_encode_grid(const VizParams& p, const GridParams& gp, int n_frames, void* stream);
int launch_flow_minmax(const float2* flow, int64_t n_px, int n_frames, unsigned* minmax, void* stream);

}  // namespace ofc
