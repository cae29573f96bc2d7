// 8-bit colour kernels of the flow path: BGR->gray and the flow visualisation
// (cartToPolar -> hue byte -> min-max value byte -> HSV->BGR).
//
// Replaces the cv2 calls of the reference's ComputeOpticalFLow.compute
// (k-means-color-clustering/computeOpticalFlowModule.py:19,25-33; the same chain
// inline at computeOpticalFlow.py:96-120).  Every formula is the verified
// restatement of SURVEY.md Appendix A.3; rounding-sensitive steps use explicit
// round-to-nearest intrinsics so nvcc cannot contract or reassociate them.
#include "ofc_common.cuh"
#include "viz_kernels.cuh"

namespace ofc {

// (3735*B + 19235*G + 9798*R + 16384) >> 15, four pixels per thread
__global__ void __launch_bounds__(256) bgr2gray_kernel(const unsigned char* __restrict__ bgr,
                                                       unsigned char* __restrict__ gray, int64_t n_px) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // group of 4 pixels
    int64_t px = q * 4;
    if (px >= n_px) return;
    if (px + 3 < n_px) {
        const unsigned* s = reinterpret_cast<const unsigned*>(bgr + px * 3);
        unsigned w0 = s[0], w1 = s[1], w2 = s[2];
        unsigned b0 = w0 & 255, g0 = (w0 >> 8) & 255, r0 = (w0 >> 16) & 255;
        unsigned b1 = w0 >> 24, g1 = w1 & 255, r1 = (w1 >> 8) & 255;
        unsigned b2 = (w1 >> 16) & 255, g2 = w1 >> 24, r2 = w2 & 255;
        unsigned b3 = (w2 >> 8) & 255, g3 = (w2 >> 16) & 255, r3 = w2 >> 24;
        unsigned y0 = (3735u * b0 + 19235u * g0 + 9798u * r0 + 16384u) >> 15;
        unsigned y1 = (3735u * b1 + 19235u * g1 + 9798u * r1 + 16384u) >> 15;
        unsigned y2 = (3735u * b2 + 19235u * g2 + 9798u * r2 + 16384u) >> 15;
        unsigned y3 = (3735u * b3 + 19235u * g3 + 9798u * r3 + 16384u) >> 15;
        *reinterpret_cast<unsigned*>(gray + px) = y0 | (y1 << 8) | (y2 << 16) | (y3 << 24);
    } else {
        for (; px < n_px; ++px) {
            unsigned b = bgr[px * 3], g = bgr[px * 3 + 1], r = bgr[px * 3 + 2];
            gray[px] = (unsigned char)((3735u * b + 19235u * g + 9798u * r + 16384u) >> 15);
        }
    }
}

__device__ __forceinline__ float magnitude(float x, float y) {
    return sqrtf(__fmaf_rn(x, x, __fmul_rn(y, y)));
}

// cv::cartToPolar angle (degrees), cv2 4.13 polynomial with fused multiply-adds
__device__ __forceinline__ float angle_deg(float x, float y) {
    const float p1 = 57.283627f, p3 = -18.667446f, p5 = 8.9140005f, p7 = -2.5397246f;
    float ax = fabsf(x), ay = fabsf(y);
    float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    float c = __fdiv_rn(mn, __fadd_rn(mx, 2.220446e-16f));
    float c2 = __fmul_rn(c, c);
    float a = __fmaf_rn(__fmaf_rn(__fmaf_rn(c2, p7, p5), c2, p3), c2, p1);
    a = __fmul_rn(a, c);
    if (ax < ay) a = __fsub_rn(90.f, a);
    if (x < 0.f) a = __fsub_rn(180.f, a);
    if (y < 0.f) a = __fsub_rn(360.f, a);
    return a;
}

// round_tail: cv2 finishes the last W % 32 pixels of a row with scalar code whose
// float->u8 step rounds (half to even) where the SIMD body truncates.
__device__ __forceinline__ void encode_pixel(float fxv, float fyv, float fscale, float fshift, bool round_tail,
                                             unsigned char& B, unsigned char& G, unsigned char& Rr, float& mag) {
    mag = magnitude(fxv, fyv);
    float rad = __fmul_rn(angle_deg(fxv, fyv), 0.017453292f);          // f32(pi/180)
    // mask[...,0] = angle*180/np.pi/2 evaluated in float32, truncated to uint8
    float hv = __fmul_rn(__fdiv_rn(__fmul_rn(rad, 180.f), 3.1415927f), 0.5f);      // /2 is exact either way
    int H = (int)hv;
    // cv.normalize(NORM_MINMAX): fmaf(m, f32(scale), f32(shift)), truncated to uint8
    float vv = __fmaf_rn(mag, fscale, fshift);
    int V = (int)vv;
    V = V < 0 ? 0 : (V > 255 ? 255 : V);
    // cv.cvtColor(HSV2BGR), S = 255, float sector formula, truncation
    float h = __fmul_rn((float)H, 0.033333335f);                      // f32(6/180)
    float v = __fmul_rn((float)V, 0.003921569f);                     // f32(1/255)
    const float s = __fmul_rn(255.f, 0.003921569f);
    int sec = (int)floorf(h);
    float fr = __fsub_rn(h, (float)sec);
    sec %= 6;
    if (sec < 0) sec += 6;
    float t0 = v;
    float t1 = __fmul_rn(v, __fsub_rn(1.f, s));
    float t2 = __fmul_rn(v, __fsub_rn(1.f, __fmul_rn(s, fr)));
    float t3 = __fmul_rn(v, __fsub_rn(1.f, __fmul_rn(s, __fsub_rn(1.f, fr))));
    float b, g, r;
    switch (sec) {
        case 0: b = t1; g = t3; r = t0; break;
        case 1: b = t1; g = t0; r = t2; break;
        case 2: b = t3; g = t0; r = t1; break;
        case 3: b = t0; g = t2; r = t1; break;
        case 4: b = t0; g = t1; r = t3; break;
        default: b = t2; g = t1; r = t0; break;
    }
    b = __fmul_rn(b, 255.f); g = __fmul_rn(g, 255.f); r = __fmul_rn(r, 255.f);
    int bi = round_tail ? __float2int_rn(b) : (int)b;
    int gi = round_tail ? __float2int_rn(g) : (int)g;
    int ri = round_tail ? __float2int_rn(r) : (int)r;
    B = (unsigned char)(bi < 0 ? 0 : (bi > 255 ? 255 : bi));
    G = (unsigned char)(gi < 0 ? 0 : (gi > 255 ? 255 : gi));
    Rr = (unsigned char)(ri < 0 ? 0 : (ri > 255 ? 255 : ri));
}

// x may run past the row end inside a quad: wrap to the next row's column
__device__ __forceinline__ bool is_tail(int x, int width, int tail_from) {
    if (x >= width) x -= width;
    return x >= tail_from;
}

// flow -> BGR u8 (+ optional per-frame sum of magnitudes for the reference's
// "Average Magnitude" CSV, computeOpticalFlow.py:114-117).  blockIdx.y = frame.
__global__ void __launch_bounds__(256) flow_encode_kernel(VizParams p) {
    const int frame = blockIdx.y;
    const float2* flow = p.flow + (int64_t)frame * p.n_px;
    unsigned char* out = p.bgr + (int64_t)frame * p.n_px * 3;
    float mn = __uint_as_float(p.minmax[2 * frame]), mxv = __uint_as_float(p.minmax[2 * frame + 1]);
    double range = (double)mxv - (double)mn;
    double scale = 255.0 * (range > 2.220446049250313e-16 ? 1.0 / range : 0.0);
    // cv2 4.13 derives the float shift from the float-rounded scale (verified
    // bit-exact on 3000 random arrays; see oracle/viz_np.py)
    const float fscale = (float)scale;
    const float fshift = -__fmul_rn(mn, fscale);

    const int tail_from = p.width - (p.width % 32);
    double local = 0.0;
    const int64_t n_quads = (p.n_px + 3) / 4;
    const bool aligned = (p.n_px & 3) == 0;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n_quads; q += (int64_t)gridDim.x * blockDim.x) {
        int64_t px = q * 4;
        if (px + 3 < p.n_px && aligned) {
            const float4* f4 = reinterpret_cast<const float4*>(flow + px);
            float4 a = f4[0], b = f4[1];
            unsigned char c[12];
            float m0, m1, m2, m3;
            const int x = (int)((unsigned)px % (unsigned)p.width);          // a frame has < 2^32 pixels
            encode_pixel(a.x, a.y, fscale, fshift, is_tail(x, p.width, tail_from), c[0], c[1], c[2], m0);
            encode_pixel(a.z, a.w, fscale, fshift, is_tail(x + 1, p.width, tail_from), c[3], c[4], c[5], m1);
            encode_pixel(b.x, b.y, fscale, fshift, is_tail(x + 2, p.width, tail_from), c[6], c[7], c[8], m2);
            encode_pixel(b.z, b.w, fscale, fshift, is_tail(x + 3, p.width, tail_from), c[9], c[10], c[11], m3);
            unsigned* o = reinterpret_cast<unsigned*>(out + px * 3);
            o[0] = c[0] | (c[1] << 8) | (c[2] << 16) | ((unsigned)c[3] << 24);
            o[1] = c[4] | (c[5] << 8) | (c[6] << 16) | ((unsigned)c[7] << 24);
            o[2] = c[8] | (c[9] << 8) | (c[10] << 16) | ((unsigned)c[11] << 24);
            local += (double)m0 + (double)m1 + (double)m2 + (double)m3;
        } else {
            for (int64_t i = px; i < px + 4 && i < p.n_px; ++i) {
                float2 f = flow[i];
                float m;
                encode_pixel(f.x, f.y, fscale, fshift, (int)((unsigned)i % (unsigned)p.width) >= tail_from, out[i * 3], out[i * 3 + 1],
                             out[i * 3 + 2], m);
                local += (double)m;
            }
        }
    }
    if (p.mag_sum) {
        __shared__ double s_part[8];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
        if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = local;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += s_part[i];
            // order-independent accumulation: the CTA's (deterministic) partial is added as 36.28 fixed
            // point, so the frame sum does not depend on which CTA gets there first; the word is turned
            // into a double by mag_sum_finalize_kernel.  A frame's sum of |flow| stays far below 2^35.
            atomicAdd(reinterpret_cast<unsigned long long*>(p.mag_sum) + frame,
                      (unsigned long long)(t * 268435456.0 + 0.5));
        }
    }
}

__global__ void mag_sum_finalize_kernel(double* mag_sum, int n_frames) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_frames) {
        const unsigned long long fx = reinterpret_cast<unsigned long long*>(mag_sum)[i];
        mag_sum[i] = (double)fx * (1.0 / 268435456.0);
    }
}

// stand-alone min/max of |flow| per frame (used when the flow did not come from
// the fused last Farneback iteration, e.g. ofc_flow_to_bgr on a user flow field)
__global__ void __launch_bounds__(256) flow_minmax_kernel(const float2* flow, int64_t n_px, unsigned* minmax) {
    const int frame = blockIdx.y;
    const float2* f = flow + (int64_t)frame * n_px;
    float lmin = 3.402823466e38f, lmax = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_px; i += (int64_t)gridDim.x * blockDim.x) {
        float2 v = f[i];
        float m = magnitude(v.x, v.y);
        lmin = fminf(lmin, m);
        lmax = fmaxf(lmax, m);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lmin = fminf(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
        lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(minmax + 2 * frame, __float_as_uint(lmin));
        atomicMax(minmax + 2 * frame + 1, __float_as_uint(lmax));
    }
}

int launch_bgr2gray(const unsigned char* bgr, unsigned char* gray, int64_t n_px, void* stream) {
    if (n_px <= 0) return OFC_OK;
    int64_t quads = (n_px + 3) / 4;
    dim3 grid((unsigned)((quads + 255) / 256));
    ProfScope prof(PK_GRAY, stream);
    OFC_LAUNCH(bgr2gray_kernel, grid, dim3(256), 0, stream, bgr, gray, n_px);
    OFC_CHECK_LAUNCH("bgr2gray");
    return OFC_OK;
}

int launch_flow_encode(const VizParams& p, int n_frames, void* stream) {
    if (n_frames <= 0 || p.n_px <= 0) return OFC_OK;
    int64_t quads = (p.n_px + 3) / 4;
    int bx = (int)((quads + 255) / 256);
    if (bx > 148 * 8) bx = 148 * 8;
    ProfScope prof(PK_ENCODE, stream);
    OFC_LAUNCH(flow_encode_kernel, dim3(bx, n_frames), dim3(256), 0, stream, p);
    OFC_CHECK_LAUNCH("flow_encode");
    if (p.mag_sum) {
        OFC_LAUNCH(mag_sum_finalize_kernel, dim3(cdiv(n_frames, 128)), dim3(128), 0, stream, p.mag_sum, n_frames);
        OFC_CHECK_LAUNCH("mag_sum_finalize");
    }
    return OFC_OK;
}

int launch_flow_minmax(const float2* flow, int64_t n_px, int n_frames, unsigned* minmax, void* stream) {
    if (n_frames <= 0 || n_px <= 0) return OFC_OK;
    int bx = (int)((n_px + 255) / 256);
    if (bx > 148 * 4) bx = 148 * 4;
    ProfScope prof(PK_FLOW_MINMAX, stream);
    OFC_LAUNCH(flow_minmax_kernel, dim3(bx, n_frames), dim3(256), 0, stream, flow, n_px, minmax);
    OFC_CHECK_LAUNCH("flow_minmax");
    return OFC_OK;
}

}  // namespace ofc
