// 8-bit colour kernels of the flow path: BGR->gray and the flow visualisation
// (cartToPolar -> hue byte -> min-max value byte -> HSV->BGR).
//
// Replaces the cv2 calls of the reference's ComputeOpticalFLow.compute
// (k-means-color-clustering/computeOpticalFlowModule.py:19,25-33; the same chain
// inline at computeOpticalFlow.py:96-120).  Every formula is the verified
// restatement of SURVEY.md Appendix A.3; rounding-sensitive steps use explicit
// round-to-nearest intrinsics so nvcc cannot contract or reassociate them.
#include "ofc_common.cuh"
#include "color_math.cuh"
#include "grid_kernels.cuh"
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

// ---------------------------------------------------------------------------
// The same bytes with far fewer instructions (the production path).  encode_pixel above stays as the
// plain statement of the arithmetic and as the slow path.
//   * hue: the byte only needs trunc(hv).  hv is evaluated with a fast reciprocal instead of the two IEEE
//     divisions and without the rad -> deg round trip: |hv' - hv| < 1e-4 (1-2 ulp in the quotient, at most
//     four roundings of the conversion chain, on values <= 180).  When hv' is further than 2e-3 from an
//     integer its truncation is hv's; the few pixels nearer than that run the exact chain.
//   * value: trunc(v) for 0 <= v < 2^23 is the low mantissa of (v + 2^23) rounded toward zero -- FADD.RZ is
//     a full-rate instruction, float<->int conversions are quarter rate (rint likewise with FADD.RN).
//   * HSV -> BGR with S = 255: every channel is  trunc(((V / 255) * m) * 255)  with m one of
//     {1, 1 - s, 1 - s fr, 1 - s (1 - fr)} chosen by the hue sector.  The three multipliers per hue (0..180)
//     are a 181-entry table built per CTA with the exact operations of encode_pixel.
// ---------------------------------------------------------------------------
constexpr int HUE_TAB = 181;

// table entry H -> (m_b, m_g, m_r): multipliers of v in encode_pixel's sector switch
__device__ __forceinline__ float4 hue_multipliers(int H) {
    const float h = __fmul_rn((float)H, 0.033333335f);
    int sec = (int)floorf(h);
    const float fr = __fsub_rn(h, (float)sec);
    sec %= 6;
    if (sec < 0) sec += 6;
    const float s = __fmul_rn(255.f, 0.003921569f);
    const float m0 = 1.f;
    const float m1 = __fsub_rn(1.f, s);
    const float m2 = __fsub_rn(1.f, __fmul_rn(s, fr));
    const float m3 = __fsub_rn(1.f, __fmul_rn(s, __fsub_rn(1.f, fr)));
    switch (sec) {
        case 0: return make_float4(m1, m3, m0, 0.f);
        case 1: return make_float4(m1, m0, m2, 0.f);
        case 2: return make_float4(m3, m0, m1, 0.f);
        case 3: return make_float4(m0, m2, m1, 0.f);
        case 4: return make_float4(m0, m1, m3, 0.f);
        default: return make_float4(m2, m1, m0, 0.f);
    }
}

__device__ __forceinline__ void build_hue_table(float4* tab) {
    for (int i = threadIdx.x; i < HUE_TAB; i += blockDim.x) tab[i] = hue_multipliers(i);
}

// exact hue byte (the chain of encode_pixel)
__device__ __forceinline__ float exact_hue(float fxv, float fyv) {
    const float rad = __fmul_rn(angle_deg(fxv, fyv), 0.017453292f);
    const float hv = __fmul_rn(__fdiv_rn(__fmul_rn(rad, 180.f), 3.1415927f), 0.5f);
    return (float)(int)hv;
}

// low 8 bits: trunc(y) (or rint(y) for the scalar tail of a row), y clamped at 0, y < 256
__device__ __forceinline__ unsigned to_byte_bits(float y, bool round_tail) {
    y = fmaxf(y, 0.f);
    return __float_as_uint(round_tail ? __fadd_rn(y, 8388608.f) : __fadd_rz(y, 8388608.f));
}

// one pixel: returns B | G << 8 | R << 16 (and H, V bytes through hv_out when asked: H | 255 << 8 | V << 16)
__device__ __forceinline__ unsigned encode_pixel_fast(float fxv, float fyv, float fscale, float fshift, bool round_tail,
                                                      const float4* __restrict__ tab, float& mag, unsigned* hv_out) {
    mag = magnitude(fxv, fyv);
    // value byte
    const float vv = __fmaf_rn(mag, fscale, fshift);
    float Vf = __fadd_rz(fmaxf(vv, 0.f), 8388608.f) - 8388608.f;
    Vf = fminf(Vf, 255.f);
    // hue byte
    const float ax = fabsf(fxv), ay = fabsf(fyv);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float c = __fdividef(mn, __fadd_rn(mx, 2.220446e-16f));
    const float c2 = c * c;
    const float p1 = 57.283627f, p3 = -18.667446f, p5 = 8.9140005f, p7 = -2.5397246f;
    float a = fmaf(fmaf(fmaf(c2, p7, p5), c2, p3), c2, p1) * c;
    if (ax < ay) a = 90.f - a;
    if (fxv < 0.f) a = 180.f - a;
    if (fyv < 0.f) a = 360.f - a;
    const float hv = a * 0.5f;
    float Hf = __fadd_rz(hv, 8388608.f) - 8388608.f;
    const float frac = hv - Hf;
    if (!(frac > 2e-3f && frac < 1.f - 2e-3f)) Hf = exact_hue(fxv, fyv);
    const unsigned Hi = __float_as_uint(Hf + 8388608.f) & 255u;
    // HSV -> BGR
    const float4 m = tab[Hi < HUE_TAB ? Hi : HUE_TAB - 1];
    const float v = __fmul_rn(Vf, 0.003921569f);
    const unsigned b = to_byte_bits(__fmul_rn(__fmul_rn(v, m.x), 255.f), round_tail);
    const unsigned g = to_byte_bits(__fmul_rn(__fmul_rn(v, m.y), 255.f), round_tail);
    const unsigned r = to_byte_bits(__fmul_rn(__fmul_rn(v, m.z), 255.f), round_tail);
    if (hv_out) *hv_out = Hi | (255u << 8) | ((__float_as_uint(Vf + 8388608.f) & 255u) << 16);
    return (b & 255u) | ((g & 255u) << 8) | ((r & 255u) << 16);
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

    __shared__ float4 s_tab[HUE_TAB];
    build_hue_table(s_tab);
    __syncthreads();
    unsigned char* hsv = p.hsv ? p.hsv + (int64_t)frame * p.n_px * 3 : nullptr;
    const int tail_from = p.width - (p.width % 32);
    double local = 0.0;
    const int64_t n_quads = (p.n_px + 3) / 4;
    const bool aligned = (p.n_px & 3) == 0;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n_quads; q += (int64_t)gridDim.x * blockDim.x) {
        int64_t px = q * 4;
        if (px + 3 < p.n_px && aligned) {
            const float4* f4 = reinterpret_cast<const float4*>(flow + px);
            float4 a = f4[0], b = f4[1];
            float m0, m1, m2, m3;
            unsigned h0, h1, h2, h3;
            const int x = (int)((unsigned)px % (unsigned)p.width);          // a frame has < 2^32 pixels
            const unsigned c0 = encode_pixel_fast(a.x, a.y, fscale, fshift, is_tail(x, p.width, tail_from), s_tab, m0, hsv ? &h0 : nullptr);
            const unsigned c1 = encode_pixel_fast(a.z, a.w, fscale, fshift, is_tail(x + 1, p.width, tail_from), s_tab, m1, hsv ? &h1 : nullptr);
            const unsigned c2 = encode_pixel_fast(b.x, b.y, fscale, fshift, is_tail(x + 2, p.width, tail_from), s_tab, m2, hsv ? &h2 : nullptr);
            const unsigned c3 = encode_pixel_fast(b.z, b.w, fscale, fshift, is_tail(x + 3, p.width, tail_from), s_tab, m3, hsv ? &h3 : nullptr);
            unsigned* o = reinterpret_cast<unsigned*>(out + px * 3);
            o[0] = c0 | (c1 << 24);
            o[1] = (c1 >> 8) | (c2 << 16);
            o[2] = (c2 >> 16) | (c3 << 8);
            if (hsv) {
                unsigned* ho = reinterpret_cast<unsigned*>(hsv + px * 3);
                ho[0] = h0 | (h1 << 24);
                ho[1] = (h1 >> 8) | (h2 << 16);
                ho[2] = (h2 >> 16) | (h3 << 8);
            }
            local += (double)m0 + (double)m1 + (double)m2 + (double)m3;
        } else {
            for (int64_t i = px; i < px + 4 && i < p.n_px; ++i) {
                float2 f = flow[i];
                float m;
                unsigned hh;
                const unsigned c = encode_pixel_fast(f.x, f.y, fscale, fshift, (int)((unsigned)i % (unsigned)p.width) >= tail_from, s_tab, m,
                                                     hsv ? &hh : nullptr);
                out[i * 3] = (unsigned char)c; out[i * 3 + 1] = (unsigned char)(c >> 8); out[i * 3 + 2] = (unsigned char)(c >> 16);
                if (hsv) { hsv[i * 3] = (unsigned char)hh; hsv[i * 3 + 1] = 255; hsv[i * 3 + 2] = (unsigned char)(hh >> 16); }
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

// ---------------------------------------------------------------------------
// Visualisation + grid pass fused: one CTA per (grid cell, frame) encodes the cell's pixels (and the frame's
// right / bottom remainder next to the last column / row of cells), writes the BGR bytes once and keeps the
// per-cell sums of overlayGridAndComputeAvgColor and of the k = 1 colour cluster in registers -- the
// visualisation is never read back (grid_cells_kernel's arithmetic, KmeanGrids.py:52-113, 269-339).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) flow_encode_grid_kernel(VizParams p, GridParams gp) {
    const int cell = blockIdx.x, frame = blockIdx.y;
    const int cy = cell / gp.cols, cx = cell - cy * gp.cols;
    const int W = gp.W, H = gp.H;
    const int x1 = cx * gp.x_step, y1 = cy * gp.y_step;
    const int cw = min(x1 + gp.x_step, W) - x1, chh = min(y1 + gp.y_step, H) - y1;          // the cell proper
    const int rw = (cx == gp.cols - 1 ? W : x1 + gp.x_step) - x1;                           // region this CTA encodes
    const int rh = (cy == gp.rows - 1 ? H : y1 + gp.y_step) - y1;
    const float2* flow = p.flow + (int64_t)frame * p.n_px;
    unsigned char* out = p.bgr + (int64_t)frame * p.n_px * 3;
    const float mn = __uint_as_float(p.minmax[2 * frame]), mxv = __uint_as_float(p.minmax[2 * frame + 1]);
    const double range = (double)mxv - (double)mn;
    const double scale = 255.0 * (range > 2.220446049250313e-16 ? 1.0 / range : 0.0);
    const float fscale = (float)scale;
    const float fshift = -__fmul_rn(mn, fscale);
    __shared__ float4 s_tab[HUE_TAB];
    build_hue_table(s_tab);
    __syncthreads();
    const int tail_from = W - (W % 32);
    const bool mean_row = gp.draw_lines && cy > 0, mean_col = gp.draw_lines && cx > 0;

    unsigned sa[3] = {0, 0, 0}, sb[4] = {0, 0, 0, 0};
    double local = 0.0;
    auto add_pixel = [&](unsigned c, int ly, int lx) {
        if (ly >= chh || lx >= cw) return;                       // remainder pixels belong to no cell
        const unsigned c0 = c & 255u, c1 = (c >> 8) & 255u, c2 = (c >> 16) & 255u;
        const bool wa = (ly == 0 && mean_row) || (lx == 0 && mean_col);
        sa[0] += wa ? 255u : c0; sa[1] += wa ? 255u : c1; sa[2] += wa ? 255u : c2;
        const bool wb = gp.draw_lines && (ly == 0 || lx == 0);
        unsigned k0 = wb ? 255u : c0, k1 = wb ? 255u : c1, k2 = wb ? 255u : c2;
        if (gp.threshold) {
            k0 = k0 < (unsigned)gp.threshold ? 0u : k0;
            k1 = k1 < (unsigned)gp.threshold ? 0u : k1;
            k2 = k2 < (unsigned)gp.threshold ? 0u : k2;
        }
        const unsigned gray = (3735u * k0 + 19235u * k1 + 9798u * k2 + 16384u) >> 15;
        sb[0] += k0; sb[1] += k1; sb[2] += k2; sb[3] += gray > 0 ? 1u : 0u;
    };
    const bool quads = (rw & 3) == 0 && (x1 & 3) == 0 && (W & 3) == 0 && ((uintptr_t)p.flow & 15) == 0 && ((uintptr_t)p.bgr & 3) == 0 &&
                       (p.n_px & 3) == 0;
    if (quads) {
        const int qpr = rw >> 2;
        const int rows_per_pass = 256 / qpr > 0 ? 256 / qpr : 1;
        const int lq = threadIdx.x % qpr, ly0 = threadIdx.x / qpr;
        if (ly0 < rows_per_pass) {
            for (int ly = ly0; ly < rh; ly += rows_per_pass) {
                for (int q = lq; q < qpr; q += 256) {            // one step unless a region row is wider than 1024 px
                    const int x = x1 + q * 4;
                    const int64_t px = (int64_t)(y1 + ly) * W + x;
                    const float4* f4 = reinterpret_cast<const float4*>(flow + px);
                    const float4 a = f4[0], b = f4[1];
                    float m0, m1, m2, m3;
                    const unsigned c0 = encode_pixel_fast(a.x, a.y, fscale, fshift, x >= tail_from, s_tab, m0, nullptr);
                    const unsigned c1 = encode_pixel_fast(a.z, a.w, fscale, fshift, x + 1 >= tail_from, s_tab, m1, nullptr);
                    const unsigned c2 = encode_pixel_fast(b.x, b.y, fscale, fshift, x + 2 >= tail_from, s_tab, m2, nullptr);
                    const unsigned c3 = encode_pixel_fast(b.z, b.w, fscale, fshift, x + 3 >= tail_from, s_tab, m3, nullptr);
                    unsigned* o = reinterpret_cast<unsigned*>(out + px * 3);
                    o[0] = c0 | (c1 << 24);
                    o[1] = (c1 >> 8) | (c2 << 16);
                    o[2] = (c2 >> 16) | (c3 << 8);
                    local += (double)m0 + (double)m1 + (double)m2 + (double)m3;
                    add_pixel(c0, ly, q * 4); add_pixel(c1, ly, q * 4 + 1); add_pixel(c2, ly, q * 4 + 2); add_pixel(c3, ly, q * 4 + 3);
                }
            }
        }
    } else {
        const int n_reg = rw * rh;
        for (int i = threadIdx.x; i < n_reg; i += 256) {
            const int ly = i / rw, lx = i - ly * rw;
            const int64_t px = (int64_t)(y1 + ly) * W + x1 + lx;
            const float2 f = flow[px];
            float m;
            const unsigned c = encode_pixel_fast(f.x, f.y, fscale, fshift, x1 + lx >= tail_from, s_tab, m, nullptr);
            out[px * 3] = (unsigned char)c; out[px * 3 + 1] = (unsigned char)(c >> 8); out[px * 3 + 2] = (unsigned char)(c >> 16);
            local += (double)m;
            add_pixel(c, ly, lx);
        }
    }
    // cell results (grid_cells_kernel's tail) and the frame's magnitude sum
    __shared__ unsigned s_red[8][7];
    __shared__ double s_part[8];
    unsigned vals[7] = {sa[0], sa[1], sa[2], sb[0], sb[1], sb[2], sb[3]};
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        unsigned v = vals[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        vals[k] = v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) {
        for (int k = 0; k < 7; ++k) s_red[threadIdx.x >> 5][k] = vals[k];
        s_part[threadIdx.x >> 5] = local;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t[7] = {0, 0, 0, 0, 0, 0, 0};
        double tm = 0.0;
        for (int wv = 0; wv < 8; ++wv) {
            for (int k = 0; k < 7; ++k) t[k] += s_red[wv][k];
            tm += s_part[wv];
        }
        if (p.mag_sum)
            atomicAdd(reinterpret_cast<unsigned long long*>(p.mag_sum) + frame, (unsigned long long)(tm * 268435456.0 + 0.5));
        const int n = cw * chh;
        const int64_t o = (int64_t)frame * gridDim.x + cell;
        if (n > 0) {
            const unsigned a0 = t[0] / n, a1 = t[1] / n, a2 = t[2] / n;           // floor: .astype(uint8)
            if (gp.avg_bgr) { gp.avg_bgr[o * 3] = (unsigned char)a0; gp.avg_bgr[o * 3 + 1] = (unsigned char)a1; gp.avg_bgr[o * 3 + 2] = (unsigned char)a2; }
            if (gp.avg_hue) gp.avg_hue[o] = (unsigned char)hue_of_bgr((int)a0, (int)a1, (int)a2);
            const unsigned k0 = rint_div(t[3], n), k1 = rint_div(t[4], n), k2 = rint_div(t[5], n);
            const unsigned k3 = rint_div((unsigned long long)t[6] * 255ull, n);
            if (gp.km_centre) { gp.km_centre[o * 4] = (unsigned char)k0; gp.km_centre[o * 4 + 1] = (unsigned char)k1; gp.km_centre[o * 4 + 2] = (unsigned char)k2; gp.km_centre[o * 4 + 3] = (unsigned char)k3; }
            if (gp.km_hue) gp.km_hue[o] = (unsigned char)hue_of_bgr((int)k0, (int)k1, (int)k2);
            if (gp.km_sums) { gp.km_sums[o * 4] = t[3]; gp.km_sums[o * 4 + 1] = t[4]; gp.km_sums[o * 4 + 2] = t[5]; gp.km_sums[o * 4 + 3] = t[6] * 255u; }
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

int launch_flow_encode_grid(const VizParams& p, const GridParams& gp, int n_frames, void* stream) {
    if (n_frames <= 0 || p.n_px <= 0) return OFC_OK;
    ProfScope prof(PK_ENCODE, stream);
    OFC_LAUNCH(flow_encode_grid_kernel, dim3(gp.rows * gp.cols, n_frames), dim3(256), 0, stream, p, gp);
    OFC_CHECK_LAUNCH("flow_encode_grid");
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
