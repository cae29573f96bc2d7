// Farneback dense optical flow for sm_100a: Gaussian pre-filter + resample,
// separable polynomial expansion, and the fused
// update-matrices -> box blur -> 2x2 solve iteration.
//
// Replaces the arithmetic behind the reference's call
//   cv.calcOpticalFlowFarneback(prev, next, None, 0.5, 3, 15, 3, 5, 1.2, 0)
// (reference k-means-color-clustering/computeOpticalFlowModule.py:20-22,
// computeOpticalFlow.py:99-101).  Algorithm: SURVEY.md Appendix A.1/A.2.
//
// Layout in HBM (per pyramid level, per frame):
//   I   float  [h][w]      pre-filtered, resampled image
//   RA  float4 [h][w]      polynomial coefficients R0..R3
//   RB  float  [h][w]      polynomial coefficient  R4
// and per frame pair: flow float2 [h][w] (two ping-pong buffers per level).
// R is split float4 + float so the bilinear gather of the warped frame is one
// 128-bit and one 32-bit load per tap instead of five scalar loads.
#include <stdlib.h>

#include <type_traits>

#include "ofc_common.cuh"
#include "flow_kernels.cuh"

namespace ofc {

// ---------------------------------------------------------------------------
// INTER_LINEAR source coordinate (cv::resize): index + fraction, clamped
// ---------------------------------------------------------------------------
__device__ __forceinline__ void src_coord(int d, double scale, int n_src, int& i, float& fr) {
    float f = (float)(((double)d + 0.5) * scale - 0.5);
    int ii = (int)floorf(f);
    fr = f - (float)ii;
    if (ii < 0) { ii = 0; fr = 0.f; }
    if (ii >= n_src - 1) { ii = n_src - 1; fr = 0.f; }
    i = ii;
}

// asynchronous copies into shared memory (LDGSTS) and tensor-memory loads / stores, with their host-emulation forms
#ifndef OFC_EMULATE
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16_cg(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
// dst / src byte offsets as immediates: one address register serves several copies
template <int DOFF, int SOFF> __device__ __forceinline__ void cp_async16_o(unsigned dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0+%2], [%1+%3], 16;" ::"r"(dst), "l"(src), "n"(DOFF), "n"(SOFF) : "memory");
}
template <int DOFF, int SOFF> __device__ __forceinline__ void cp_async4_o(unsigned dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0+%2], [%1+%3], 4;" ::"r"(dst), "l"(src), "n"(DOFF), "n"(SOFF) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tmem_ld8(unsigned taddr, float (&v)[8]) {
    unsigned r0, r1, r2, r3, r4, r5, r6, r7;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
    v[4] = __uint_as_float(r4); v[5] = __uint_as_float(r5); v[6] = __uint_as_float(r6); v[7] = __uint_as_float(r7);
}
__device__ __forceinline__ void tmem_st8(unsigned taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
                 "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                 "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
#else
// host-side debug emulation: copies are immediate, the "tensor memory" is a per-thread array
static float ofc_emu_tmem[1024][128];
static inline void cp_async16(void* dst, const void* src) { memcpy(dst, src, 16); }
static inline void cp_async4(void* dst, const void* src) { memcpy(dst, src, 4); }
static inline void cp_async16_cg(void* dst, const void* src) { memcpy(dst, src, 16); }
static inline unsigned long long smem_u32(const void* p) { return (unsigned long long)p; }
template <int DOFF, int SOFF> static inline void cp_async16_o(unsigned long long dst, const void* src) {
    memcpy((char*)dst + DOFF, (const char*)src + SOFF, 16);
}
template <int DOFF, int SOFF> static inline void cp_async4_o(unsigned long long dst, const void* src) {
    memcpy((char*)dst + DOFF, (const char*)src + SOFF, 4);
}
static inline void cp_async_commit() {}
template <int N> static inline void cp_async_wait() {}
static inline void tmem_ld8(unsigned taddr, float (&v)[8]) { memcpy(v, &ofc_emu_tmem[threadIdx.x][taddr & 127u], 32); }
static inline void tmem_st8(unsigned taddr, const float (&v)[8]) { memcpy(&ofc_emu_tmem[threadIdx.x][taddr & 127u], v, 32); }
static inline void tmem_wait_st() {}
#endif

// ---------------------------------------------------------------------------
// K2: Gaussian pre-filter at full resolution fused with the bilinear
// down-sample to this level.  Only the blurred samples the resize reads are
// computed: a block stages the u8 source window in shared memory, runs the
// horizontal taps for the two source columns of every output column, then the
// vertical taps for the two source rows of every output row, then the lerp.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) prefilter_kernel(PrefilterParams p) {
    OFC_DYN_SMEM(unsigned char, smem);
    float* taps = reinterpret_cast<float*>(smem);
    float* hbuf = taps + p.taps_pad;                                   // [in_rows][tx][2]
    unsigned char* tile = reinterpret_cast<unsigned char*>(hbuf + (size_t)p.in_rows * p.tx * 2);  // [in_rows][in_pitch]
    __shared__ int s_ci[32], s_ri[32];        // per output column / row: source index
    __shared__ float s_fx[32], s_fy[32];      //                           and lerp fraction

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int r = p.ksz >> 1;
    const int x0 = blockIdx.x * p.tx, y0 = blockIdx.y * p.ty;
    const int nx = min(p.tx, p.w - x0), ny = min(p.ty, p.h - y0);
    const unsigned char* src = p.gray + (int64_t)blockIdx.z * p.gray_stride;

    if (tid < nx) src_coord(x0 + tid, p.sx, p.W, s_ci[tid], s_fx[tid]);
    if (tid >= 32 && tid - 32 < ny) src_coord(y0 + tid - 32, p.sy, p.H, s_ri[tid - 32], s_fy[tid - 32]);
    for (int i = tid; i < p.ksz; i += blockDim.x) taps[i] = p.taps[i];
    __syncthreads();
    const int c_lo = s_ci[0] - r, r_lo = s_ri[0] - r;
    const int ncols = min(s_ci[nx - 1] + 1, p.W - 1) + r - c_lo + 1;
    const int nrows = min(s_ri[ny - 1] + 1, p.H - 1) + r - r_lo + 1;

    for (int rr = warp; rr < nrows; rr += nwarps) {
        const unsigned char* row = src + (int64_t)reflect101(r_lo + rr, p.H) * p.W;
        for (int cc = lane; cc < ncols; cc += 32) tile[rr * p.in_pitch + cc] = row[reflect101(c_lo + cc, p.W)];
    }
    __syncthreads();

    // horizontal taps at the two sample columns of each output column
    for (int rr = warp; rr < nrows; rr += nwarps) {
        for (int j2 = lane; j2 < nx * 2; j2 += 32) {
            int xx = j2 >> 1, c = j2 & 1;
            int ci = min(s_ci[xx] + c, p.W - 1);
            const unsigned char* t = tile + rr * p.in_pitch + (ci - r - c_lo);
            float s = 0.f;
            for (int j = 0; j < p.ksz; ++j) s = fmaf(taps[j], (float)t[j], s);
            hbuf[(rr * p.tx + xx) * 2 + c] = s;
        }
    }
    __syncthreads();

    float* out = p.out + (int64_t)blockIdx.z * p.out_stride;
    for (int yy = warp; yy < ny; yy += nwarps) {
        const int ri = s_ri[yy];
        const float fy = s_fy[yy];
        const int ra = ri - r - r_lo, rb = min(ri + 1, p.H - 1) - r - r_lo;
        for (int xx = lane; xx < nx; xx += 32) {
            const float fx = s_fx[xx];
            float b00 = 0.f, b01 = 0.f, b10 = 0.f, b11 = 0.f;
            for (int j = 0; j < p.ksz; ++j) {
                float t = taps[j];
                const float2 ha = *reinterpret_cast<const float2*>(hbuf + ((ra + j) * p.tx + xx) * 2);
                const float2 hb = *reinterpret_cast<const float2*>(hbuf + ((rb + j) * p.tx + xx) * 2);
                b00 = fmaf(t, ha.x, b00); b01 = fmaf(t, ha.y, b01);
                b10 = fmaf(t, hb.x, b10); b11 = fmaf(t, hb.y, b11);
            }
            float top = b00 * (1.f - fx) + b01 * fx;
            float bot = b10 * (1.f - fx) + b11 * fx;
            out[(int64_t)(y0 + yy) * p.w + x0 + xx] = top * (1.f - fy) + bot * fy;
        }
    }
}

// Level 0 of the pyramid is never resampled and always uses the fixed 3-tap
// kernel [1/4, 1/2, 1/4] (sigma = 0).  Its products and sums are exact in
// float32 for 8-bit input, so any evaluation order gives the reference's bits.
// One thread = 4 consecutive output pixels of one row; 3 rows x 6 bytes in.
__global__ void __launch_bounds__(256) prefilter_identity3_kernel(PrefilterParams p) {
    const int gx = (blockIdx.x * 64 + (threadIdx.x & 63)) * 4;
    const int gy = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (gx >= p.W || gy >= p.H) return;
    const unsigned char* src = p.gray + (int64_t)blockIdx.z * p.gray_stride;
    float h[3][4];
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
        const unsigned char* row = src + (int64_t)reflect101(gy + dy, p.H) * p.W;
        float v[6];
        if (gx >= 4 && gx + 8 <= p.W && (p.W & 3) == 0) {
            unsigned a = *reinterpret_cast<const unsigned*>(row + gx - 4);
            unsigned b = *reinterpret_cast<const unsigned*>(row + gx);
            unsigned c = *reinterpret_cast<const unsigned*>(row + gx + 4);
            v[0] = (float)(a >> 24);
            v[1] = (float)(b & 255); v[2] = (float)((b >> 8) & 255); v[3] = (float)((b >> 16) & 255); v[4] = (float)(b >> 24);
            v[5] = (float)(c & 255);
        } else {
#pragma unroll
            for (int i = 0; i < 6; ++i) v[i] = (float)row[reflect101(gx - 1 + i, p.W)];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) h[dy + 1][i] = 0.25f * v[i] + 0.5f * v[i + 1] + 0.25f * v[i + 2];
    }
    float o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = 0.25f * h[0][i] + 0.5f * h[1][i] + 0.25f * h[2][i];
    float* out = p.out + (int64_t)blockIdx.z * p.out_stride + (int64_t)gy * p.W + gx;
    if (gx + 4 <= p.W && (p.W & 3) == 0) {
        *reinterpret_cast<float4*>(out) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
        for (int i = 0; i < 4 && gx + i < p.W; ++i) out[i] = o[i];
    }
}

// ---------------------------------------------------------------------------
// K3: separable polynomial expansion (FarnebackPolyExp), I -> (RA, RB).
// 64x32 output tile per 256-thread block; the vertical pass keeps 4 rows per
// thread in registers, the horizontal pass 4 columns per thread fed by 128-bit
// shared-memory loads.
// ---------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(256) polyexp_kernel(PolyParams p) {
    constexpr int TW = 64, TH = 32;
    constexpr int IW = TW + 2 * N, IH = TH + 2 * N;
    constexpr int PITCH = (IW + 3) / 4 * 4;
    __shared__ __align__(16) float s_in[IH * PITCH];
    __shared__ __align__(16) float s_v[3][TH * PITCH];

    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const float* I = p.I + (int64_t)blockIdx.z * p.in_stride;

    for (int i = tid; i < IH * IW; i += 256) {
        int rr = i / IW, cc = i - rr * IW;
        int sy = clampi(y0 - N + rr, 0, p.h - 1), sx = clampi(x0 - N + cc, 0, p.w - 1);
        s_in[rr * PITCH + cc] = I[(int64_t)sy * p.w + sx];
    }
    __syncthreads();

    // vertical pass: r0 = sum g*I, r1 = sum xg*(I+ - I-), r2 = sum xxg*(I+ + I-)
    for (int t = tid; t < IW * (TH / 4); t += 256) {
        int cc = t % IW, seg = t / IW;
        float v[4 + 2 * N];
#pragma unroll
        for (int j = 0; j < 4 + 2 * N; ++j) v[j] = s_in[(seg * 4 + j) * PITCH + cc];
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            float c = v[o + N];
            float r0 = c * p.g[0], r1 = 0.f, r2 = 0.f;
#pragma unroll
            for (int k = 1; k <= N; ++k) {
                float up = v[o + N - k], dn = v[o + N + k];
                float s = dn + up;
                r0 = fmaf(p.g[k], s, r0);
                r1 = fmaf(p.xg[k], dn - up, r1);
                r2 = fmaf(p.xxg[k], s, r2);
            }
            int row = seg * 4 + o;
            s_v[0][row * PITCH + cc] = r0;
            s_v[1][row * PITCH + cc] = r1;
            s_v[2][row * PITCH + cc] = r2;
        }
    }
    __syncthreads();

    float4* RA = p.RA + (int64_t)blockIdx.z * p.out_stride;
    float* RB = p.RB + (int64_t)blockIdx.z * p.out_stride;
    for (int t = tid; t < TH * (TW / 4); t += 256) {
        int seg = t % (TW / 4), row = t / (TW / 4);
        int gy = y0 + row, gx = x0 + seg * 4;
        if (gy >= p.h || gx >= p.w) continue;
        constexpr int NV = (4 + 2 * N + 3) / 4 * 4;
        float a0[NV], a1[NV], a2[NV];
#pragma unroll
        for (int j = 0; j < NV / 4; ++j) {
            float4 q0 = *reinterpret_cast<const float4*>(&s_v[0][row * PITCH + seg * 4 + j * 4]);
            float4 q1 = *reinterpret_cast<const float4*>(&s_v[1][row * PITCH + seg * 4 + j * 4]);
            float4 q2 = *reinterpret_cast<const float4*>(&s_v[2][row * PITCH + seg * 4 + j * 4]);
            a0[j * 4 + 0] = q0.x; a0[j * 4 + 1] = q0.y; a0[j * 4 + 2] = q0.z; a0[j * 4 + 3] = q0.w;
            a1[j * 4 + 0] = q1.x; a1[j * 4 + 1] = q1.y; a1[j * 4 + 2] = q1.z; a1[j * 4 + 3] = q1.w;
            a2[j * 4 + 0] = q2.x; a2[j * 4 + 1] = q2.y; a2[j * 4 + 2] = q2.z; a2[j * 4 + 3] = q2.w;
        }
        float4 ra[4];
        float rb[4];
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            const int c = o + N;
            float b1 = a0[c] * p.g[0], b3 = a1[c] * p.g[0], b5 = a2[c] * p.g[0];
            float b2 = 0.f, b4 = 0.f, b6 = 0.f;
#pragma unroll
            for (int k = 1; k <= N; ++k) {
                float tg = a0[c + k] + a0[c - k];
                b1 = fmaf(tg, p.g[k], b1);
                b4 = fmaf(tg, p.xxg[k], b4);
                b2 = fmaf(a0[c + k] - a0[c - k], p.xg[k], b2);
                b3 = fmaf(a1[c + k] + a1[c - k], p.g[k], b3);
                b6 = fmaf(a1[c + k] - a1[c - k], p.xg[k], b6);
                b5 = fmaf(a2[c + k] + a2[c - k], p.g[k], b5);
            }
            ra[o].x = b3 * p.ig11;
            ra[o].y = b2 * p.ig11;
            ra[o].z = fmaf(b1, p.ig03, b5 * p.ig33);
            ra[o].w = fmaf(b1, p.ig03, b4 * p.ig33);
            rb[o] = b6 * p.ig55;
        }
        int64_t base = (int64_t)gy * p.w + gx;
        if (gx + 3 < p.w) {
#pragma unroll
            for (int o = 0; o < 4; ++o) RA[base + o] = ra[o];
            if ((p.w & 3) == 0) {
                *reinterpret_cast<float4*>(RB + base) = make_float4(rb[0], rb[1], rb[2], rb[3]);
            } else {
#pragma unroll
                for (int o = 0; o < 4; ++o) RB[base + o] = rb[o];
            }
        } else {
            for (int o = 0; o < 4 && gx + o < p.w; ++o) { RA[base + o] = ra[o]; RB[base + o] = rb[o]; }
        }
    }
}

// ---------------------------------------------------------------------------
// K2, lean tiled form for the down-sampled levels (tap count known at compile time).
// A 256-thread CTA produces a 32x8 output tile:
//   1  source window (8-bit) of the tile -> shared memory, aligned 32-bit loads in the
//      interior, reflected byte loads at the frame border;
//   2  horizontal taps for every window row at the two sample columns of each output
//      column (they are neighbours, so one pass over KSZ+1 bytes feeds both), float32,
//      tap order 0..KSZ-1 like cv::sepFilter2D;
//   3  vertical taps at the two sample rows of each output row + the bilinear lerp.
// Taps live in registers and every tap loop is unrolled.
// ---------------------------------------------------------------------------
template <int KSZ>
__global__ void __launch_bounds__(256) prefilter_tile_kernel(PrefilterParams p, int in_rows, int in_pitch) {
    constexpr int RR = KSZ / 2;
    constexpr int TX = 32, TY = 8;
    OFC_DYN_SMEM(unsigned char, smem);
    float2* hbuf = reinterpret_cast<float2*>(smem);                        // [in_rows][TX] (ha, hb)
    unsigned char* tile = smem + (size_t)in_rows * TX * sizeof(float2);     // [in_rows][in_pitch]
    __shared__ int s_ci[TX], s_ri[TY];
    __shared__ float s_fx[TX], s_fy[TY];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const int nx = min(TX, p.w - x0), ny = min(TY, p.h - y0);
    const unsigned char* src = p.gray + (int64_t)blockIdx.z * p.gray_stride;
    if (tid < TX) src_coord(min(x0 + tid, p.w - 1), p.sx, p.W, s_ci[tid], s_fx[tid]);
    else if (tid < TX + TY) src_coord(min(y0 + tid - TX, p.h - 1), p.sy, p.H, s_ri[tid - TX], s_fy[tid - TX]);
    float tp[KSZ];
#pragma unroll
    for (int j = 0; j < KSZ; ++j) tp[j] = p.taps[j];
    __syncthreads();
    // window: columns [c_lo, c_lo + ncols), rows [r_lo, r_lo + nrows); c_lo rounded down to a word
    const int c_first = s_ci[0] - RR, r_lo = s_ri[0] - RR;
    const int c_lo = c_first & ~3;
    const int ncols = min(s_ci[nx - 1] + 1, p.W - 1) + RR - c_lo + 1;
    const int nrows = min(s_ri[ny - 1] + 1, p.H - 1) + RR - r_lo + 1;
    const bool interior = c_lo >= 0 && c_lo + ((ncols + 3) & ~3) <= p.W && r_lo >= 0 && r_lo + nrows <= p.H &&
                          (p.W & 3) == 0 && (p.gray_stride & 3) == 0 && ((uintptr_t)p.gray & 3) == 0;
    if (interior) {
        const int nwords = (ncols + 3) >> 2;
        for (int i = tid; i < nrows * nwords; i += 256) {
            const int rr = i / nwords, wc = i - rr * nwords;
            *reinterpret_cast<unsigned*>(tile + rr * in_pitch + wc * 4) =
                *reinterpret_cast<const unsigned*>(src + (int64_t)(r_lo + rr) * p.W + c_lo + wc * 4);
        }
    } else {
        for (int rr = warp; rr < nrows; rr += 8) {
            const unsigned char* row = src + (int64_t)reflect101(r_lo + rr, p.H) * p.W;
            for (int cc = lane; cc < ncols; cc += 32) tile[rr * in_pitch + cc] = row[reflect101(c_lo + cc, p.W)];
        }
    }
    __syncthreads();
    // horizontal taps: lane = output column, warps stride over the window rows
    {
        const int xx = lane < nx ? lane : nx - 1;
        const int ci = s_ci[xx];
        const int d = min(ci + 1, p.W - 1) - ci;         // 1, or 0 at the right frame border
        const unsigned char* base = tile + (ci - RR - c_lo);
        for (int rr = warp; rr < nrows; rr += 8) {
            const unsigned char* t = base + rr * in_pitch;
            float ha = 0.f, hb = 0.f;
            float prev = (float)t[0];
#pragma unroll
            for (int j = 0; j < KSZ; ++j) {
                const float nxt = (float)t[j + 1];       // one byte past the window is inside the padded pitch
                ha = fmaf(tp[j], prev, ha);
                hb = fmaf(tp[j], d ? nxt : prev, hb);
                prev = nxt;
            }
            hbuf[rr * TX + lane] = make_float2(ha, hb);
        }
    }
    __syncthreads();
    // vertical taps + lerp: one output per thread
    if (lane < nx && warp < ny) {
        const int ri = s_ri[warp];
        const int ra = ri - RR - r_lo, rb = min(ri + 1, p.H - 1) - RR - r_lo;
        float b00 = 0.f, b01 = 0.f, b10 = 0.f, b11 = 0.f;
#pragma unroll
        for (int j = 0; j < KSZ; ++j) {
            const float2 ha = hbuf[(ra + j) * TX + lane], hb = hbuf[(rb + j) * TX + lane];
            b00 = fmaf(tp[j], ha.x, b00); b01 = fmaf(tp[j], ha.y, b01);
            b10 = fmaf(tp[j], hb.x, b10); b11 = fmaf(tp[j], hb.y, b11);
        }
        const float fx = s_fx[lane], fy = s_fy[warp];
        const float top = b00 * (1.f - fx) + b01 * fx;
        const float bot = b10 * (1.f - fx) + b11 * fx;
        p.out[(int64_t)blockIdx.z * p.out_stride + (int64_t)(y0 + warp) * p.w + x0 + lane] = top * (1.f - fy) + bot * fy;
    }
}

// ---------------------------------------------------------------------------
// K2, pyramid form (production when a level is an exact 1/S fraction of the frame, S = 2, 4, 8, 16 -- every
// level of pyr_scale = 0.5 on frames whose sides divide): ALL such levels in one launch.
//
// With an exact ratio the resize samples sit at column S x + S/2 - 1/2, i.e. between source columns
// ci = S x + S/2 - 1 and ci + 1 with weight 1/2 each (rows alike), so an output needs the blurred image at
// a 2 x 2 block only.  A thread owns one output column and walks down the source rows of its row segment:
//   * the CTA stages 8 (or S) source rows of its column range in shared memory as bytes (aligned words);
//   * per source row the thread reads its KSZ + 1 window bytes as words and forms BOTH horizontal sums
//     (columns ci and ci + 1) from the same registers, taps in ascending order like cv::sepFilter2D;
//   * the row then feeds the vertical chains of the (at most three) output rows whose windows contain it --
//     tap indices are compile-time because the walk is unrolled over one period of S rows -- again in
//     ascending tap order, so every output is bit-identical to prefilter_tile_kernel's;
//   * an output row is finished S rows after the next one started: lerp (weights 1/2), one coalesced store.
// No per-tile prologue, no float64 coordinate math, every source byte converted once per output column pair
// instead of once per tap, and three launches become one.
// ---------------------------------------------------------------------------
struct PyrLevel {
    float* out;
    int64_t out_stride;
    const float* taps;
    int w, h, S;
    int blocks_x, segs, rows_per_seg, first_block;
};
struct PyrParams {
    const unsigned char* gray;
    int64_t gray_stride;
    int W, H, n_levels;
    PyrLevel lv[4];
};
constexpr int PYR_NT = 128;

__device__ __forceinline__ float byte_to_float_sel(unsigned w, int t) {
#ifndef OFC_EMULATE
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540 + t)) - 8388608.f;
#else
    return (float)((w >> (8 * t)) & 255u);
#endif
}

template <int S, int KSZ>
__device__ __forceinline__ void pyr_level(const PyrParams& p, const PyrLevel& L, int local_block, int frame, unsigned* srow) {
    constexpr int NT = PYR_NT;
    constexpr int RR = KSZ / 2;
    constexpr int C0 = RR + 1 - S / 2;                          // tap index (first sample row) of group m's row 0 for output row m
    constexpr int PH_END = KSZ - S - C0;                         // phase at which output row m - 1 receives its last tap
    constexpr int A_OFF = ((S / 2 - 1 - RR) % 4 + 4) % 4;        // byte offset of thread 0's window in the staged row
    constexpr int RB = S < 8 ? 8 : S;                            // source rows staged per barrier
    constexpr int GB = RB / S;                                   // groups of S rows per staged batch
    constexpr int PW = (S * (NT - 1) + A_OFF + KSZ + 1 + 3) / 4; // words per staged row
    constexpr int NWD = S == 2 ? 2 : (A_OFF + KSZ + 1 + 3) / 4;  // words a thread reads per row
    static_assert(PH_END >= 0 && PH_END < S && C0 < S && 2 * S + C0 > KSZ, "three output rows in flight");

    const int t = threadIdx.x;
    const int seg = local_block / L.blocks_x, bx = local_block - seg * L.blocks_x;
    const int ox0 = bx * NT, x = ox0 + t;
    const int oy0 = seg * L.rows_per_seg, oy1 = min(L.h, oy0 + L.rows_per_seg);
    const int W = p.W, H = p.H;
    const unsigned char* __restrict__ gray = p.gray + (int64_t)frame * p.gray_stride;
    float* __restrict__ out = L.out + (int64_t)frame * L.out_stride;
    const int cb = S * ox0 + S / 2 - 1 - RR - A_OFF;             // source column of staged byte 0 (a multiple of 4)
    float tp[KSZ];
#pragma unroll
    for (int j = 0; j < KSZ; ++j) tp[j] = L.taps[j];
    const int o = S * t + A_OFF;                                 // byte offset of this thread's window
    const int wofs = o >> 2;
    const int shift = (o & 3) * 8;                               // S >= 4: A_OFF * 8 for every thread

    float acc[3][4];                                             // [slot: output row m+1, m, m-1][b00, b01, b10, b11]
#pragma unroll
    for (int q = 0; q < 3; ++q) acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.f;

    const int m_lo = oy0 - 1, n_groups = oy1 - oy0 + 2;
    const int n_batches = (n_groups + GB - 1) / GB;
    for (int bi = 0; bi < n_batches; ++bi) {
        __syncthreads();                                         // the previous batch has been consumed
        const int r0 = S * m_lo + bi * RB;
        // warp w stages rows w, w + 4, ...; lanes stride over the row's words (no index division; the row reflection
        // needs no modulo: a batch overshoots the frame by less than a frame height)
        for (int lr = t >> 5; lr < RB; lr += NT / 32) {
            const unsigned char* row = gray + (int64_t)reflect101_near(r0 + lr, H) * W;
            unsigned* dst = srow + lr * PW;
            for (int wc = t & 31; wc < PW; wc += 32) {
                const int gc = cb + 4 * wc;
                unsigned v;
                if (gc >= 0 && gc + 3 < W) {
                    v = *reinterpret_cast<const unsigned*>(row + gc);
                } else {
                    v = (unsigned)row[reflect101(gc, W)] | ((unsigned)row[reflect101(gc + 1, W)] << 8) |
                        ((unsigned)row[reflect101(gc + 2, W)] << 16) | ((unsigned)row[reflect101(gc + 3, W)] << 24);
                }
                dst[wc] = v;
            }
        }
        __syncthreads();
#pragma unroll
        for (int g = 0; g < GB; ++g) {
            const int m = m_lo + bi * GB + g;
#pragma unroll
            for (int ph = 0; ph < S; ++ph) {
                const unsigned* rw = srow + (g * S + ph) * PW + wofs;
                unsigned wd[NWD];
#pragma unroll
                for (int q = 0; q < NWD; ++q) wd[q] = rw[q];
                float ha = 0.f, hb = 0.f;
                if (S == 2) {
                    const unsigned v = __funnelshift_r(wd[0], wd[1], shift);
                    float prev = byte_to_float_sel(v, 0);
#pragma unroll
                    for (int j = 0; j < KSZ; ++j) {
                        const float nxt = byte_to_float_sel(v, j + 1);
                        ha = fmaf(tp[j], prev, ha);
                        hb = fmaf(tp[j], nxt, hb);
                        prev = nxt;
                    }
                } else {
                    float prev = byte_to_float_sel(wd[A_OFF >> 2], A_OFF & 3);
#pragma unroll
                    for (int j = 0; j < KSZ; ++j) {
                        const float nxt = byte_to_float_sel(wd[(A_OFF + j + 1) >> 2], (A_OFF + j + 1) & 3);
                        ha = fmaf(tp[j], prev, ha);
                        hb = fmaf(tp[j], nxt, hb);
                        prev = nxt;
                    }
                }
#pragma unroll
                for (int dl = -1; dl <= 1; ++dl) {
                    const int j0 = S * dl + ph + C0;             // tap of this row for the first sample row of output m - dl
                    const int slot = dl + 1;
                    if (j0 >= 0 && j0 < KSZ) {
                        acc[slot][0] = fmaf(tp[j0 < KSZ && j0 >= 0 ? j0 : 0], ha, acc[slot][0]);
                        acc[slot][1] = fmaf(tp[j0 < KSZ && j0 >= 0 ? j0 : 0], hb, acc[slot][1]);
                    }
                    if (j0 - 1 >= 0 && j0 - 1 < KSZ) {
                        acc[slot][2] = fmaf(tp[j0 - 1 >= 0 && j0 - 1 < KSZ ? j0 - 1 : 0], ha, acc[slot][2]);
                        acc[slot][3] = fmaf(tp[j0 - 1 >= 0 && j0 - 1 < KSZ ? j0 - 1 : 0], hb, acc[slot][3]);
                    }
                }
                if (ph == PH_END) {
                    const int y = m - 1;
                    if (y >= oy0 && y < oy1 && x < L.w) {
                        const float fx = 0.5f, fy = 0.5f;
                        const float top = acc[2][0] * (1.f - fx) + acc[2][1] * fx;
                        const float bot = acc[2][2] * (1.f - fx) + acc[2][3] * fx;
                        out[(int64_t)y * L.w + x] = top * (1.f - fy) + bot * fy;
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) { acc[2][c] = acc[1][c]; acc[1][c] = acc[0][c]; acc[0][c] = 0.f; }
        }
    }
}

__global__ void __launch_bounds__(PYR_NT) prefilter_pyr_kernel(PyrParams p) {
    OFC_DYN_SMEM(unsigned, srow);
    const int frame = blockIdx.y;
    int b = blockIdx.x;
    for (int l = 0; l < p.n_levels; ++l) {
        const PyrLevel& L = p.lv[l];
        const int nb = L.blocks_x * L.segs;
        if (b < nb) {
            switch (L.S) {
                case 2: pyr_level<2, 3>(p, L, b, frame, srow); break;
                case 4: pyr_level<4, 9>(p, L, b, frame, srow); break;
                case 8: pyr_level<8, 19>(p, L, b, frame, srow); break;
                default: pyr_level<16, 39>(p, L, b, frame, srow); break;
            }
            return;
        }
        b -= nb;
    }
}

// ---------------------------------------------------------------------------
// K3, strip-walk form (production): thread t owns column x0-N+t and walks down a
// strip of rows with the 2N+G-row vertical window of I in registers; every G rows
// the vertical results (r0,r1,r2) of G rows go through shared memory and the CTA
// does the horizontal pass, 4 pixels per thread, writing RA/RB with 128-bit
// stores.  Only the horizontal halo is recomputed (1.08x at TW=128) and the
// vertical halo once per strip, against 1.52x for the 64x32 tile kernel above.
// FUSE3: the level is full resolution, so I is the fixed 3x3 blur [1 2 1]^2/16 of the
// 8-bit frame (exact in integers): it is produced on the fly from `gray` and
// written to p.I only as a by-product.
// ---------------------------------------------------------------------------
// STAGE (FUSE3 only, frame width and base address multiples of 16): the 8-bit rows the next group's I needs are copied
// into a small shared-memory ring with cp.async one group ahead, instead of 18 byte loads per thread whose first use
// stalled the walk (ncu r02h: 27 % of the stall samples on the [1 2 1] sums right behind those loads).
template <int N, int TW, int NT, bool FUSE3, bool STAGE = false>
// (minimum CTAs per SM: the staged form needs ~100 registers for its 18 shared-memory bytes in flight; at 5 CTAs of 160
// threads it spills and runs slower than unstaged, at 3 it loses occupancy: 1.02 / 0.69 / 0.73 ms per 33 frames at 5 / 4 / 3
// against 0.755 ms unstaged, r02t-r02u)
__global__ void __launch_bounds__(NT, STAGE ? (NT >= 160 ? 4 : 6) : 0) polyexp_strip_kernel(
    PolyParams p, const unsigned char* __restrict__ gray_all,
                                                           int64_t gray_stride, float* __restrict__ I_out, int n_cols,
                                                           int64_t total_rows) {
    constexpr int G = 4;
    constexpr int CW = TW + 2 * N;
    constexpr int CP = (CW + 3) / 4 * 4 + 4;
    constexpr int SEG = TW / 4;
    constexpr int NV = (4 + 2 * N + 3) / 4 * 4;
    static_assert(NT >= CW && NT >= G * SEG, "thread count");
    static_assert(!STAGE || FUSE3, "staging is for the 8-bit frame");
    // two hand-over buffers, alternating per group: one barrier per group (between the vertical and the horizontal pass)
    __shared__ __align__(16) float vbuf2[2][G][3][CP];
    // staged 8-bit rows: ring of 32 rows (virtual row number & 31), GPB bytes of columns [a0, a0 + GPB) each
    constexpr int GCH = (CW + 2 + 15 + 15) / 16;         // 16-byte chunks that cover CW + 2 columns from any alignment
    constexpr int GPB = GCH * 16;
    static_assert(!STAGE || NT >= (G + 2) * GCH, "one chunk per thread");
    __shared__ __align__(16) unsigned char gring[STAGE ? 32 : 1][STAGE ? GPB : 16];

    const int t = threadIdx.x;
    const int w = p.w, h = p.h;
    // persistent CTAs over the flattened rows of all (frame, column strip) units (see flow_iter_strip_kernel)
    const int64_t range_lo = total_rows * blockIdx.x / gridDim.x;
    const int64_t range_hi = total_rows * (blockIdx.x + 1) / gridDim.x;
    for (int64_t cur = range_lo; cur < range_hi;) {
    const int unit = (int)(cur / h);
    const int ys = (int)(cur - (int64_t)unit * h);
    const int64_t left = range_hi - cur;
    const int ye = left < (int64_t)(h - ys) ? ys + (int)left : h;
    cur += ye - ys;
    const int frame = unit / n_cols;
    const int x0 = (unit - frame * n_cols) * TW;
    const bool col_thread = t < CW;
    const int gx = clampi(x0 - N + t, 0, w - 1);
    const float* __restrict__ I = FUSE3 ? nullptr : p.I + (int64_t)frame * p.in_stride;
    const unsigned char* __restrict__ gray = FUSE3 ? gray_all + (int64_t)frame * gray_stride : nullptr;
    float4* __restrict__ RA = p.RA + (int64_t)frame * p.out_stride;
    float* __restrict__ RB = p.RB + (int64_t)frame * p.out_stride;
    float* __restrict__ Iw = (FUSE3 && I_out) ? I_out + (int64_t)frame * p.out_stride : nullptr;
    const int xl = FUSE3 ? reflect101_near(gx - 1, w) : 0, xr = FUSE3 ? reflect101_near(gx + 1, w) : 0;

    // I at (clamped row, this thread's column)
    auto hsum = [&](int row) -> int {                   // [1 2 1] across columns of one (reflected) source row
        const unsigned char* r = gray + (int64_t)reflect101_near(row, h) * w;
        return (int)r[xl] + 2 * (int)r[gx] + (int)r[xr];
    };
    auto load_I = [&](int row) -> float {
        const int rc = clampi(row, 0, h - 1);
        if (!FUSE3) return I[(int64_t)rc * w + gx];
        const int acc = hsum(rc - 1) + 2 * hsum(rc) + hsum(rc + 1);
        return (float)acc * 0.0625f;                   // exact: acc <= 4080
    };
    // staged rows: columns from a0 (a multiple of 16) on; rows first .. first + n - 1 (virtual numbers, reflected here)
    const int a0 = STAGE ? (max(x0 - N - 1, 0) & ~15) : 0;
    auto stage_rows = [&](int first, int n) {
        if constexpr (STAGE) {
            if (t < n * GCH) {
                const int rr = t / GCH, ch = t - rr * GCH;
                const int col = a0 + ch * 16;
                if (col < w)                             // w is a multiple of 16: the chunk is inside the row
                    cp_async16(&gring[(first + rr + 32) & 31][ch * 16], gray + (int64_t)reflect101_near(first + rr, h) * w + col);
            }
            cp_async_commit();
        }
    };
    auto hsum_staged = [&](int vrow) -> int {
        const unsigned char* r = gring[(vrow + 32) & 31];
        return (int)r[xl - a0] + 2 * (int)r[gx - a0] + (int)r[xr - a0];
    };
    // G consecutive un-clamped rows starting at `first`: each row needs only one new horizontal sum
    auto load_I_run = [&](int first, float* dst, auto staged) {
        if constexpr (FUSE3) {
            if (first >= 0 && first + G <= h) {
                if constexpr (STAGE && decltype(staged)::value) {
                    int hs[G + 2];
#pragma unroll
                    for (int i = 0; i < G + 2; ++i) hs[i] = hsum_staged(first - 1 + i);
#pragma unroll
                    for (int i = 0; i < G; ++i) dst[i] = (float)(hs[i] + 2 * hs[i + 1] + hs[i + 2]) * 0.0625f;
                    return;
                } else {
                    int hs[G + 2];
#pragma unroll
                    for (int i = 0; i < G + 2; ++i) hs[i] = hsum(first - 1 + i);
#pragma unroll
                    for (int i = 0; i < G; ++i) dst[i] = (float)(hs[i] + 2 * hs[i + 1] + hs[i + 2]) * 0.0625f;
                    return;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < G; ++i) dst[i] = load_I(first + i);
    };

    float v[2 * N + G];                                // I(rows r-N .. r+G-1+N) of the current group
    float nxt[G];
    if (col_thread) {
#pragma unroll
        for (int j = 0; j < 2 * N; ++j) v[G + j] = load_I(ys - N + j);      // becomes v[0..2N) after the first shift
        load_I_run(ys + N, nxt, std::false_type{});
    }
    if (STAGE) {
        // the 8-bit rows group 0 turns into the I rows of group 1 (the ring may still be read by a slower warp of the
        // previous segment: barrier first)
        __syncthreads();
        stage_rows(ys + G + N - 1, G + 2);
        stage_rows(ys + 2 * G + N + 1, G);               // and the new rows of the group after
        cp_async_wait<0>();
    }
    __syncthreads();                                     // staged rows visible; hand-over buffers of the previous segment free
    for (int g0 = 0; g0 < ye - ys; g0 += G) {
        float (*vbuf)[3][CP] = vbuf2[(g0 / G) & 1];
        if (col_thread) {
#pragma unroll
            for (int j = 0; j < 2 * N; ++j) v[j] = v[j + G];
#pragma unroll
            for (int i = 0; i < G; ++i) v[2 * N + i] = nxt[i];
            // rows of the next group: staged one group ago (STAGE), else in flight while this group is processed
            load_I_run(ys + g0 + G + N, nxt, std::true_type{});
        }
        // 8-bit rows the group after the next converts: G new rows, two groups ahead (the copies of the previous group are
        // waited for before this group's barrier, which publishes them)
        if (STAGE) stage_rows(ys + g0 + 3 * G + N + 1, G);
        if (col_thread) {
            if (FUSE3 && Iw && t >= N && t < N + TW && x0 + t - N < w) {
#pragma unroll
                for (int i = 0; i < G; ++i)
                    if (ys + g0 + i < ye) Iw[(int64_t)(ys + g0 + i) * w + gx] = v[N + i];
            }
#pragma unroll
            for (int o = 0; o < G; ++o) {
                const float c = v[o + N];
                float r0 = c * p.g[0], r1 = 0.f, r2 = 0.f;
#pragma unroll
                for (int k = 1; k <= N; ++k) {
                    const float up = v[o + N - k], dn = v[o + N + k];
                    const float sm = dn + up;
                    r0 = fmaf(p.g[k], sm, r0);
                    r1 = fmaf(p.xg[k], dn - up, r1);
                    r2 = fmaf(p.xxg[k], sm, r2);
                }
                vbuf[o][0][t] = r0;
                vbuf[o][1][t] = r1;
                vbuf[o][2][t] = r2;
            }
        }
        if (STAGE) cp_async_wait<1>();                   // all but this group's copies have landed
        __syncthreads();
        if (t < G * SEG) {
            const int o = t / SEG, seg = t - o * SEG;
            const int gy = ys + g0 + o, gx0 = x0 + seg * 4;
            if (gy < ye && gx0 < w) {
                float a[NV];
                float b1[4], b2[4], b3[4], b4[4], b5[4], b6[4];
                // channel r0 -> b1, b4, b2
#pragma unroll
                for (int j = 0; j < NV / 4; ++j) {
                    const float4 q = *reinterpret_cast<const float4*>(&vbuf[o][0][seg * 4 + j * 4]);
                    a[j * 4 + 0] = q.x; a[j * 4 + 1] = q.y; a[j * 4 + 2] = q.z; a[j * 4 + 3] = q.w;
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int c = e + N;
                    float x1 = a[c] * p.g[0], x4 = 0.f, x2 = 0.f;
#pragma unroll
                    for (int k = 1; k <= N; ++k) {
                        const float tg = a[c + k] + a[c - k];
                        x1 = fmaf(tg, p.g[k], x1);
                        x4 = fmaf(tg, p.xxg[k], x4);
                        x2 = fmaf(a[c + k] - a[c - k], p.xg[k], x2);
                    }
                    b1[e] = x1; b4[e] = x4; b2[e] = x2;
                }
                // channel r1 -> b3, b6
#pragma unroll
                for (int j = 0; j < NV / 4; ++j) {
                    const float4 q = *reinterpret_cast<const float4*>(&vbuf[o][1][seg * 4 + j * 4]);
                    a[j * 4 + 0] = q.x; a[j * 4 + 1] = q.y; a[j * 4 + 2] = q.z; a[j * 4 + 3] = q.w;
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int c = e + N;
                    float x3 = a[c] * p.g[0], x6 = 0.f;
#pragma unroll
                    for (int k = 1; k <= N; ++k) {
                        x3 = fmaf(a[c + k] + a[c - k], p.g[k], x3);
                        x6 = fmaf(a[c + k] - a[c - k], p.xg[k], x6);
                    }
                    b3[e] = x3; b6[e] = x6;
                }
                // channel r2 -> b5
#pragma unroll
                for (int j = 0; j < NV / 4; ++j) {
                    const float4 q = *reinterpret_cast<const float4*>(&vbuf[o][2][seg * 4 + j * 4]);
                    a[j * 4 + 0] = q.x; a[j * 4 + 1] = q.y; a[j * 4 + 2] = q.z; a[j * 4 + 3] = q.w;
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int c = e + N;
                    float x5 = a[c] * p.g[0];
#pragma unroll
                    for (int k = 1; k <= N; ++k) x5 = fmaf(a[c + k] + a[c - k], p.g[k], x5);
                    b5[e] = x5;
                }
                float4 ra[4];
                float rb[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    ra[e].x = b3[e] * p.ig11;
                    ra[e].y = b2[e] * p.ig11;
                    ra[e].z = fmaf(b1[e], p.ig03, b5[e] * p.ig33);
                    ra[e].w = fmaf(b1[e], p.ig03, b4[e] * p.ig33);
                    rb[e] = b6[e] * p.ig55;
                }
                const int64_t base = (int64_t)gy * w + gx0;
                if (gx0 + 3 < w) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) RA[base + e] = ra[e];
                    if ((w & 3) == 0) {
                        *reinterpret_cast<float4*>(RB + base) = make_float4(rb[0], rb[1], rb[2], rb[3]);
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e) RB[base + e] = rb[e];
                    }
                } else {
                    for (int e = 0; e < 4 && gx0 + e < w; ++e) { RA[base + e] = ra[e]; RB[base + e] = rb[e]; }
                }
            }
        }
    }
    if (STAGE) cp_async_wait<0>();
    }   // segments of this CTA's row range
}

// The 2x2 solve of one pixel from the five window sums S (box window): OpenCV evaluates, in float64,
//     g = S * scale,  idet = 1 / (g11 g22 - g12^2 + 1e-3),  flow = ((g11 h2 - g12 h1) idet, (g22 h1 - g12 h2) idet)
// and rounds to float32.  The scale cancels: flow = N / (D + 1e-3 / scale^2) with N, D the same 2x2 determinants of the
// raw sums.  Both determinants are evaluated in float32 with error-free products (a b = p + e exactly, e from one fma),
// so the cancellation in D and N costs nothing: the results are within about one float32 ulp of the float64 evaluation
// (tests: same bars against cv2 as before; measured r02l: emulated mean end-point error against cv2 1.69e-7 -> 1.74e-7 px),
// at a fifth of the float64-pipe and conversion work: the three level-0 launches 3.62 -> 3.43 ms per 32 pairs (r02n).  `c_hi + c_lo` is
// 1e-3 / scale^2 split into two floats on the host.
__device__ __forceinline__ float det_eft(float a, float b, float c, float d, float& lo) {
    // a b - c d = (p - q) + (e - f) with p + e = a b and q + f = c d exactly
    const float p = __fmul_rn(a, b), q = __fmul_rn(c, d);
    const float e = __fmaf_rn(a, b, -p), f = __fmaf_rn(c, d, -q);
    lo = __fsub_rn(e, f);
    return __fsub_rn(p, q);
}

__device__ __forceinline__ float rcp_seed(float x) {
#ifndef OFC_EMULATE
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
}

__device__ __forceinline__ float2 solve2x2(const IterParams& p, float s11, float s12, float s22, float sh1, float sh2) {
    float dl, n1l, n2l;
    const float dh = det_eft(s11, s22, s12, s12, dl);
    const float n1h = det_eft(s11, sh2, s12, sh1, n1l);
    const float n2h = det_eft(s22, sh1, s12, sh2, n2l);
    // D + c > 0: D >= 0 up to rounding (a sum of positive semi-definite matrices), c = 50.625 for winsize 15
    const float den = __fadd_rn(__fadd_rn(dh, p.solve_c_hi), __fadd_rn(dl, p.solve_c_lo));
    const float n1 = __fadd_rn(n1h, n1l), n2 = __fadd_rn(n2h, n2l);
    const float r0 = rcp_seed(den);
    const float r = __fmaf_rn(__fmaf_rn(-den, r0, 1.f), r0, r0);               // one Newton step: ~1 ulp
    const float q1 = __fmul_rn(n1, r), q2 = __fmul_rn(n2, r);
    // residual correction: the quotient is correctly rounded except in rare half-way cases
    return make_float2(__fmaf_rn(__fmaf_rn(-q1, den, n1), r, q1), __fmaf_rn(__fmaf_rn(-q2, den, n2), r, q2));
}

// ---------------------------------------------------------------------------
// K4+K5+K6 fused: one Farneback iteration.
//   phase 1  M(y,x) for the tile + box-filter halo, straight from R0, warped R1
//            and the current flow (optionally the x2 up-sampled coarse flow):
//            M never goes to HBM;
//   phase 2  vertical (2R+1)-row sliding sums, in place, one column per thread;
//   phase 3  horizontal sliding sums for 8 pixels per thread (128-bit shared
//            loads), the 2x2 solve, the flow store, and -- on the last
//            iteration of level 0 -- the magnitude min/max the visualisation
//            needs (warp-shuffle reduce + one atomic pair per warp).
// ---------------------------------------------------------------------------
// cv::resize(INTER_LINEAR) of the coarser level's flow at one pixel, times 1/pyr_scale: the ONE expression every
// consumer uses (tile kernel, stand-alone up-sample kernel, fused strip walk), so they agree bit for bit
__device__ __forceinline__ float2 bilerp_raw(float2 q00, float2 q01, float2 q10, float2 q11, float fx, float fy) {
    const float ax = 1.f - fx, ay = 1.f - fy;
    const float tx0 = q00.x * ax + q01.x * fx, tx1 = q10.x * ax + q11.x * fx;
    const float ty0 = q00.y * ax + q01.y * fx, ty1 = q10.y * ax + q11.y * fx;
    return make_float2(tx0 * ay + tx1 * fy, ty0 * ay + ty1 * fy);
}
__device__ __forceinline__ float2 bilerp_flow(float2 q00, float2 q01, float2 q10, float2 q11, float fx, float fy, double mul) {
    const float2 r = bilerp_raw(q00, q01, q10, q11, fx, fy);
    return make_float2((float)((double)r.x * mul), (float)((double)r.y * mul));
}

__device__ __forceinline__ float2 load_flow(const IterParams& p, const float2* fin, int gx, int gy) {
    if (fin == nullptr) return make_float2(0.f, 0.f);
    if (!p.upsample) return fin[(int64_t)gy * p.w + gx];
    int xi, yi; float fx, fy;
    src_coord(gx, p.usx, p.wc, xi, fx);
    src_coord(gy, p.usy, p.hc, yi, fy);
    int xj = min(xi + 1, p.wc - 1), yj = min(yi + 1, p.hc - 1);
    float2 q00 = fin[(int64_t)yi * p.wc + xi], q01 = fin[(int64_t)yi * p.wc + xj];
    float2 q10 = fin[(int64_t)yj * p.wc + xi], q11 = fin[(int64_t)yj * p.wc + xj];
    return bilerp_flow(q00, q01, q10, q11, fx, fy, p.flow_mul);
}

template <int R, int TH, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) flow_iter_kernel(IterParams p) {
    constexpr int TW = 64;
    static_assert(NT >= 8 * TH, "one phase-3 task per thread");
    constexpr int MW = TW + 2 * R, MH = TH + 2 * R;
    constexpr int P0 = (MW + 2 + 3) / 4 * 4;                 // room for the 128-bit over-read
    constexpr int PITCH = ((P0 / 4) % 2 == 0) ? P0 + 4 : P0; // pitch/4 odd: conflict-free LDS.128
    constexpr int CH = MH * PITCH;
    OFC_DYN_SMEM(float, sm);                                 // [5][MH][PITCH]

    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const int pair = blockIdx.z;
    const int w = p.w, h = p.h;
    const float4* RA0 = p.RA + (int64_t)pair * p.r_stride;
    const float* RB0 = p.RB + (int64_t)pair * p.r_stride;
    const float4* RA1 = RA0 + p.r_next;
    const float* RB1 = RB0 + p.r_next;
    const float2* fin = p.flow_in ? p.flow_in + (int64_t)pair * p.flow_in_stride : nullptr;

    // ---- phase 1: update matrices on the haloed tile -----------------------
    for (int i = tid; i < MH * MW; i += NT) {
        int my = i / MW, mx = i - my * MW;
        int gy = clampi(y0 - R + my, 0, h - 1), gx = clampi(x0 - R + mx, 0, w - 1);
        float2 fl = load_flow(p, fin, gx, gy);
        float dx = fl.x, dy = fl.y;
        float fx = (float)gx + dx, fy = (float)gy + dy;
        int x1 = (int)floorf(fx), y1 = (int)floorf(fy);
        fx -= (float)x1; fy -= (float)y1;
        int64_t o0 = (int64_t)gy * w + gx;
        float4 a = RA0[o0];
        float b = RB0[o0];
        float r2, r3, r4, r5, r6;
        if ((unsigned)x1 < (unsigned)(w - 1) && (unsigned)y1 < (unsigned)(h - 1)) {
            float a01 = fx * (1.f - fy), a10 = (1.f - fx) * fy, a11 = fx * fy, a00 = (1.f - fx) * (1.f - fy);
            int64_t o1 = (int64_t)y1 * w + x1;
            float4 q00 = RA1[o1], q01 = RA1[o1 + 1], q10 = RA1[o1 + w], q11 = RA1[o1 + w + 1];
            float s00 = RB1[o1], s01 = RB1[o1 + 1], s10 = RB1[o1 + w], s11 = RB1[o1 + w + 1];
            r2 = a00 * q00.x + a01 * q01.x + a10 * q10.x + a11 * q11.x;
            r3 = a00 * q00.y + a01 * q01.y + a10 * q10.y + a11 * q11.y;
            r4 = a00 * q00.z + a01 * q01.z + a10 * q10.z + a11 * q11.z;
            r5 = a00 * q00.w + a01 * q01.w + a10 * q10.w + a11 * q11.w;
            r6 = a00 * s00 + a01 * s01 + a10 * s10 + a11 * s11;
            r4 = (a.z + r4) * 0.5f;
            r5 = (a.w + r5) * 0.5f;
            r6 = (b + r6) * 0.25f;
        } else {
            r2 = r3 = 0.f;
            r4 = a.z; r5 = a.w; r6 = b * 0.5f;
        }
        r2 = (a.x - r2) * 0.5f;
        r3 = (a.y - r3) * 0.5f;
        r2 += r4 * dy + r6 * dx;
        r3 += r6 * dy + r5 * dx;
        if ((unsigned)(gx - 5) >= (unsigned)(w - 10) || (unsigned)(gy - 5) >= (unsigned)(h - 10)) {
            float sc = (gx < 5 ? p.border[gx] : 1.f) * (gx >= w - 5 ? p.border[w - gx - 1] : 1.f) *
                       (gy < 5 ? p.border[gy] : 1.f) * (gy >= h - 5 ? p.border[h - gy - 1] : 1.f);
            r2 *= sc; r3 *= sc; r4 *= sc; r5 *= sc; r6 *= sc;
        }
        float* d = sm + my * PITCH + mx;
        d[0 * CH] = r4 * r4 + r6 * r6;
        d[1 * CH] = (r4 + r5) * r6;
        d[2 * CH] = r5 * r5 + r6 * r6;
        d[3 * CH] = r4 * r2 + r6 * r3;
        d[4 * CH] = r6 * r2 + r5 * r3;
    }
    __syncthreads();

    // ---- phase 2: vertical sliding sums, in place --------------------------
    for (int t = tid; t < 5 * MW; t += NT) {
        int ch = t / MW, mx = t - ch * MW;
        float* col = sm + ch * CH + mx;
        float ring[2 * R + 1];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < MH; ++i) {
            float v = col[i * PITCH];
            if (i >= 2 * R + 1) s -= ring[i % (2 * R + 1)];
            ring[i % (2 * R + 1)] = v;
            s += v;
            if (i >= 2 * R) col[(i - 2 * R) * PITCH] = s;
        }
    }
    __syncthreads();

    // ---- phase 3: horizontal sums (8 px / thread) + solve ------------------
    const int row = tid % TH, seg = tid / TH;
    const int gy = y0 + row, gx0 = x0 + seg * 8;
    float lmin = 3.402823466e38f, lmax = 0.f;
    if (seg < 8 && gy < h && gx0 < w) {
        constexpr int NV = (8 + 2 * R + 3) / 4 * 4;
        float S[5][8];
#pragma unroll
        for (int ch = 0; ch < 5; ++ch) {
            const float* src = sm + ch * CH + row * PITCH + seg * 8;
            float v[NV];
#pragma unroll
            for (int j = 0; j < NV / 4; ++j) {
                float4 q = *reinterpret_cast<const float4*>(src + j * 4);
                v[j * 4 + 0] = q.x; v[j * 4 + 1] = q.y; v[j * 4 + 2] = q.z; v[j * 4 + 3] = q.w;
            }
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 2 * R + 1; ++j) s += v[j];
            S[ch][0] = s;
#pragma unroll
            for (int j = 1; j < 8; ++j) {
                s += v[j + 2 * R] - v[j - 1];
                S[ch][j] = s;
            }
        }
        float2 res[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            res[j] = solve2x2(p, S[0][j], S[1][j], S[2][j], S[3][j], S[4][j]);
        }
        float2* out = p.flow_out + (int64_t)pair * p.flow_out_stride + (int64_t)gy * w + gx0;
        if (gx0 + 7 < w && (w & 1) == 0) {
            float4* o4 = reinterpret_cast<float4*>(out);
#pragma unroll
            for (int j = 0; j < 4; ++j) o4[j] = make_float4(res[2 * j].x, res[2 * j].y, res[2 * j + 1].x, res[2 * j + 1].y);
        } else {
            for (int j = 0; j < 8 && gx0 + j < w; ++j) out[j] = res[j];
        }
        if (p.minmax) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (gx0 + j < w) {
                    float m = sqrtf(__fmaf_rn(res[j].x, res[j].x, __fmul_rn(res[j].y, res[j].y)));
                    lmin = fminf(lmin, m);
                    lmax = fmaxf(lmax, m);
                }
            }
        }
    }
    if (p.minmax) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lmin = fminf(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
            lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
        }
        if ((tid & 31) == 0) {
            // magnitudes are >= 0, so the IEEE bit patterns order like unsigned ints
            atomicMin(p.minmax + 2 * pair, __float_as_uint(lmin));
            atomicMax(p.minmax + 2 * pair + 1, __float_as_uint(lmax));
        }
    }
}

// ---------------------------------------------------------------------------
// K4: x2 bilinear up-sample of the coarser level's flow (cv::resize INTER_LINEAR
// of the float2 field, then * 1/pyr_scale) into this level's flow buffer.  Used
// by the strip-walk iteration kernel, which reads a plain flow field.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) flow_upsample_kernel(IterParams p, float2* dst, int64_t dst_stride) {
    // 64 x 16 output tile per CTA; the resize taps of its 64 columns and 16 rows are worked out once
    // (float64 coordinate like cv::resize) and shared
    __shared__ int s_xi[64], s_yi[16];
    __shared__ float s_fx[64], s_fy[16];
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const int x0 = blockIdx.x * 64, y0 = blockIdx.y * 16;
    if (threadIdx.x < 64) src_coord(min(x0 + tx, p.w - 1), p.usx, p.wc, s_xi[tx], s_fx[tx]);
    else if (threadIdx.x < 80) src_coord(min(y0 + (int)threadIdx.x - 64, p.h - 1), p.usy, p.hc, s_yi[threadIdx.x - 64], s_fy[threadIdx.x - 64]);
    __syncthreads();
    const int gx = x0 + tx;
    if (gx >= p.w) return;
    const float2* fin = p.flow_in + (int64_t)blockIdx.z * p.flow_in_stride;
    float2* out = dst + (int64_t)blockIdx.z * dst_stride;
    const int xi = s_xi[tx], xj = min(xi + 1, p.wc - 1);
    const float fx = s_fx[tx];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int ly = ty + r * 4, gy = y0 + ly;
        if (gy >= p.h) break;
        const int yi = s_yi[ly], yj = min(yi + 1, p.hc - 1);
        const float fy = s_fy[ly];
        const float2 q00 = fin[yi * p.wc + xi], q01 = fin[yi * p.wc + xj];
        const float2 q10 = fin[yj * p.wc + xi], q11 = fin[yj * p.wc + xj];
        out[(int64_t)gy * p.w + gx] = bilerp_flow(q00, q01, q10, q11, fx, fy, p.flow_mul);
    }
}

// The same resize when the level is exactly twice the coarser one (w = 2 wc, h = 2 hc, w % 4 == 0: every level of the
// reference's 1080p / 720p pyramids): cv::resize's taps are then fixed -- output column c reads source columns
// ((c - 1) >> 1, +1) with weights (0.75, 0.25) for even c and (0.25, 0.75) for odd c, rows alike, the frame border
// clamped with weight 0 -- so a thread produces a 4 x 2 output block from a 4 x 3 source block (nine loads for eight
// pixels instead of 32) without any coordinate arithmetic.  Same expressions (bilerp_flow) and therefore the same bits
// as flow_upsample_kernel.
__global__ void __launch_bounds__(256) flow_upsample_x2_kernel(IterParams p, float2* dst, int64_t dst_stride) {
    const int k = blockIdx.x * 64 + (threadIdx.x & 63);           // output columns 4k .. 4k+3
    const int j = blockIdx.y * 4 + (threadIdx.x >> 6);            // output rows 2j, 2j+1
    if (4 * k >= p.w || 2 * j >= p.h) return;
    const float2* fin = p.flow_in + (int64_t)blockIdx.z * p.flow_in_stride;
    float2* out = dst + (int64_t)blockIdx.z * dst_stride;
    const int wc = p.wc, hc = p.hc;
    // source block: columns 2k-1 .. 2k+2 and rows j-1 .. j+1, clamped to the frame
    const int cl = max(2 * k - 1, 0), cr = min(2 * k + 2, wc - 1);
    float2 q[3][4];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const float2* row = fin + (int64_t)clampi(j - 1 + r, 0, hc - 1) * wc;
        const float4 mid = *reinterpret_cast<const float4*>(row + 2 * k);     // columns 2k, 2k+1 (wc is even: 16-byte aligned)
        q[r][0] = row[cl];
        q[r][1] = make_float2(mid.x, mid.y);
        q[r][2] = make_float2(mid.z, mid.w);
        q[r][3] = row[cr];
    }
    // src_coord at scale 0.5: (c + 0.5) / 2 - 0.5 = c / 2 - 0.25 -> tap (c - 1) >> 1 with fraction 0.75 (even c) or 0.25 (odd c);
    // c = 0 reads taps (0, 1) with fraction 0 and the last column / row its clamped tap twice with fraction 0
    const bool first_col = k == 0, last_col = 2 * k + 1 >= wc - 1;
    const bool first_row = j == 0, last_row = j >= hc - 1;
    const float fx0 = first_col ? 0.f : 0.75f, fx3 = last_col ? 0.f : 0.25f;
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
        float2 a[4], b[4];                                         // the two source rows of this output row
        float fy;
        if (rr == 0) {
            fy = first_row ? 0.f : 0.75f;
#pragma unroll
            for (int c = 0; c < 4; ++c) { a[c] = first_row ? q[1][c] : q[0][c]; b[c] = first_row ? q[2][c] : q[1][c]; }
        } else {
            fy = last_row ? 0.f : 0.25f;
#pragma unroll
            for (int c = 0; c < 4; ++c) { a[c] = q[1][c]; b[c] = q[2][c]; }
        }
        float2 res[4];
        res[0] = bilerp_flow(first_col ? a[1] : a[0], first_col ? a[2] : a[1], first_col ? b[1] : b[0], first_col ? b[2] : b[1], fx0, fy, p.flow_mul);
        res[1] = bilerp_flow(a[1], a[2], b[1], b[2], 0.25f, fy, p.flow_mul);
        res[2] = bilerp_flow(a[1], a[2], b[1], b[2], 0.75f, fy, p.flow_mul);
        res[3] = bilerp_flow(a[2], a[3], b[2], b[3], fx3, fy, p.flow_mul);
        float4* o4 = reinterpret_cast<float4*>(out + (int64_t)(2 * j + rr) * p.w + 4 * k);
        o4[0] = make_float4(res[0].x, res[0].y, res[1].x, res[1].y);
        o4[1] = make_float4(res[2].x, res[2].y, res[3].x, res[3].y);
    }
}

// ---------------------------------------------------------------------------
// K5+K6 fused, strip-walk form (the production iteration kernel, winsize 15).
//
// A CTA owns TW output columns and a strip of output rows and walks down the
// strip.  Thread t < TW+2R owns halo column x0-R+t for the whole walk:
//   A  per row it evaluates M (5 channels) for its column from R0, the warped R1
//      and the current flow.  Everything that depends only on the column (clamped
//      x, border factor) is hoisted out of the walk, and the walk is software-
//      pipelined: flow / R0 of a row are loaded G rows ahead, the 8 bilinear taps
//      of R1 one row ahead (double-buffered in registers), so a thread always has
//      several rows of loads in flight;
//   B  it keeps the vertical (2R+1)-row sums of its column as float64 running sums
//      (OpenCV's vsum is double too), retiring the row that leaves the window from
//      a ring buffer in shared memory that only this thread touches -- no barrier;
//   C  every G rows the column sums of G output rows are handed over through shared
//      memory and the CTA does the horizontal sliding sums (4 pixels per thread,
//      128-bit shared loads), the 2x2 solve in float64 and the flow store.
// Compared with the square-tile kernel above: the vertical halo is paid once per
// strip instead of once per 30 rows (about 1.25x instead of 1.79x redundant M work
// at 1080p) and shared memory does not grow with the strip (ring + hand-over).
// ---------------------------------------------------------------------------
// ---------------------------------------------------------------------------
// Phase C of the strip-walk kernels: the column sums of G output rows sit in `hand`
// ([G][5][CP] floats, halo columns included); every task is 4 neighbouring pixels of one
// row: horizontal sliding sums (128-bit shared loads), 2x2 solve in float64, flow store.
// ---------------------------------------------------------------------------
template <int R, int TW, int NT, int G, bool MINMAX>
__device__ __forceinline__ void solve_rows(const IterParams& p, const float* hand, float2* __restrict__ fout, int t, int g0,
                                           int nrows, int x0, int ys, int w, float& lmin, float& lmax) {
    constexpr int K = 2 * R + 1;
    constexpr int CW = TW + 2 * R;
    constexpr int CP = (CW + 3) / 4 * 4 + 4;
    // 4 pixels per task: 8 (fewer shared loads, half the threads busy) and a balanced-tree form of the sums both measured
    // slower (r02o: 3.62 / 3.54 against 3.50 ms)
    constexpr int PX = 4;
    constexpr int SEG = TW / PX;
    static_assert(TW % PX == 0, "whole tasks per row");
    for (int task = t; task < G * SEG; task += NT) {
        const int i = task / SEG, seg = task - i * SEG;
        const int ri = g0 + i;
        const int gx0 = x0 + seg * PX;
        if (ri < 2 * R || ri >= nrows || gx0 >= w) continue;
        const int gy = ys + ri - 2 * R;
        constexpr int NV = (PX + 2 * R + 3) / 4 * 4;
        float S[5][PX];
#pragma unroll
        for (int ch = 0; ch < 5; ++ch) {
            const float* src = hand + i * (5 * CP) + ch * CP + seg * PX;
            float v[NV];
#pragma unroll
            for (int j = 0; j < NV / 4; ++j) {
                const float4 q = *reinterpret_cast<const float4*>(src + j * 4);
                v[j * 4 + 0] = q.x; v[j * 4 + 1] = q.y; v[j * 4 + 2] = q.z; v[j * 4 + 3] = q.w;
            }
            {
                float sum = 0.f;
#pragma unroll
                for (int j = 0; j < K; ++j) sum += v[j];
                S[ch][0] = sum;
#pragma unroll
                for (int j = 1; j < PX; ++j) {
                    sum += v[j + 2 * R] - v[j - 1];
                    S[ch][j] = sum;
                }
            }
        }
        float2 res[PX];
#pragma unroll
        for (int j = 0; j < PX; ++j) {
            res[j] = solve2x2(p, S[0][j], S[1][j], S[2][j], S[3][j], S[4][j]);
        }
        float2* out = fout + (int64_t)gy * w + gx0;
        if (gx0 + PX - 1 < w && (w & 1) == 0) {
            float4* o4 = reinterpret_cast<float4*>(out);
#pragma unroll
            for (int j = 0; j < PX / 2; ++j) o4[j] = make_float4(res[2 * j].x, res[2 * j].y, res[2 * j + 1].x, res[2 * j + 1].y);
        } else {
            for (int j = 0; j < PX && gx0 + j < w; ++j) out[j] = res[j];
        }
        if (MINMAX) {
#pragma unroll
            for (int j = 0; j < PX; ++j) {
                if (gx0 + j < w) {
                    const float m = sqrtf(__fmaf_rn(res[j].x, res[j].x, __fmul_rn(res[j].y, res[j].y)));
                    lmin = fminf(lmin, m);
                    lmax = fmaxf(lmax, m);
                }
            }
        }
    }
}

struct Taps {
    float4 q00, q01, q10, q11;
    float s00, s01, s10, s11;
};

// Where the warped sample of (gx, gy) falls: integer corner, fractions, in-bounds flag.  Cheap enough
// to evaluate twice (when the taps are issued and when they are consumed) instead of carrying it.
__device__ __forceinline__ void warp_point(float fgx, int gy, float dx, float dy, int w, int h,
                                           int& x1, int& y1, float& fx, float& fy, bool& inb) {
    fx = fgx + dx;
    fy = (float)gy + dy;
    x1 = (int)floorf(fx);
    y1 = (int)floorf(fy);
    fx -= (float)x1;
    fy -= (float)y1;
    inb = (unsigned)x1 < (unsigned)(w - 1) && (unsigned)y1 < (unsigned)(h - 1);
}

// `RA` / `RB` point at the pair's "prev" frame; the "next" frame is r_next pixels further
__device__ __forceinline__ void issue_taps(Taps& g, const float4* __restrict__ RA, const float* __restrict__ RB, int r_next,
                                           int w, int h, float fgx, int gy, float dx, float dy) {
    int x1, y1;
    float fx, fy;
    bool inb;
    warp_point(fgx, gy, dx, dy, w, h, x1, y1, fx, fy, inb);
    // loads are unconditional (clamped address) so that they can be issued back to back;
    // a level has < 2^30 pixels, so pixel offsets are 32-bit integers
    const int o1 = clampi(y1, 0, h - 2) * w + clampi(x1, 0, w - 2) + r_next;
    const float4* pa = RA + o1;
    const float* pb = RB + o1;
    g.q00 = pa[0]; g.q01 = pa[1]; g.q10 = pa[w]; g.q11 = pa[w + 1];
    g.s00 = pb[0]; g.s01 = pb[1]; g.s10 = pb[w]; g.s11 = pb[w + 1];
}

template <int R, int TW, int NT, int G, int MINB, bool MINMAX>
__global__ void __launch_bounds__(NT, MINB) flow_iter_strip_kernel(IterParams p, int n_cols, int64_t total_rows) {
    constexpr int K = 2 * R + 1;
    constexpr int CW = TW + 2 * R;                       // halo'd columns
    constexpr int CP = (CW + 3) / 4 * 4 + 4;             // hand-over pitch (room for the 128-bit over-read)
    constexpr int RING = (K * 5 * CW + 3) / 4 * 4;       // keeps the hand-over buffer 16-byte aligned
    static_assert(NT >= CW, "one thread per halo column");
    static_assert(G % 2 == 0, "the tap double buffer alternates by row parity");
    OFC_DYN_SMEM(float, sm);
    float* ring = sm;                                    // [K][5][CW]
    float* hand = sm + RING;                             // [G][5][CP]

    const int t = threadIdx.x;
    const int w = p.w, h = p.h;
    // Persistent CTAs: the output rows of all (pair, column strip) units are laid end to end
    // (total_rows = n_pairs * n_cols * h) and cut into gridDim.x equal ranges; a CTA walks its
    // range, restarting the pipeline where it crosses into the next unit.
    const int64_t range_lo = total_rows * blockIdx.x / gridDim.x;
    const int64_t range_hi = total_rows * (blockIdx.x + 1) / gridDim.x;
    for (int64_t cur = range_lo; cur < range_hi;) {
    const int unit = (int)(cur / h);
    const int ys = (int)(cur - (int64_t)unit * h);
    const int64_t left = range_hi - cur;
    const int ye = left < (int64_t)(h - ys) ? ys + (int)left : h;
    cur += ye - ys;
    const int pair = unit / n_cols;
    const int x0 = (unit - pair * n_cols) * TW;
    const float4* __restrict__ RA0 = p.RA + (int64_t)pair * p.r_stride;
    const float* __restrict__ RB0 = p.RB + (int64_t)pair * p.r_stride;
    const int r_next = (int)p.r_next;
    const float2* __restrict__ fin = p.flow_in ? p.flow_in + (int64_t)pair * p.flow_in_stride : nullptr;
    float2* __restrict__ fout = p.flow_out + (int64_t)pair * p.flow_out_stride;

    // ---- per-column constants ---------------------------------------------
    const bool col_thread = t < CW;
    const int gx = clampi(x0 - R + t, 0, w - 1);
    const float bxs = (gx < 5 ? p.border[gx] : 1.f) * (gx >= w - 5 ? p.border[w - gx - 1] : 1.f);
    const bool x_edge = (unsigned)(gx - 5) >= (unsigned)(w - 10);
    const float fgx = (float)gx;
    const int nrows = (ye - ys) + 2 * R;
    const int row0 = ys - R;

    double cs0 = 0.0, cs1 = 0.0, cs2 = 0.0, cs3 = 0.0, cs4 = 0.0;
    float lmin = 3.402823466e38f, lmax = 0.f;
    int slot = 0;
    // an all-zero ring lets every row retire "the row K steps back" unconditionally
    for (int i = t; i < K * 5 * CW; i += NT) ring[i] = 0.f;
    __syncthreads();

    // ---- pipeline prologue: flow / R0 of the first G rows, taps of row 0 ----
    constexpr int PD = 2;                                // rows of flow / R0 in flight ahead of the arithmetic
    static_assert(G % PD == 0, "prefetch slots are indexed by the unrolled row number");
    float2 fl[PD];
    float4 na;                                           // R0 of the next row (distance 1, like the taps)
    float nb;
    Taps tp[2];
    if (col_thread) {
#pragma unroll
        for (int i = 0; i < PD; ++i) {
            const int o = clampi(row0 + i, 0, h - 1) * w + gx;
            fl[i] = fin ? fin[o] : make_float2(0.f, 0.f);
            if (i == 0) { na = RA0[o]; nb = RB0[o]; }
        }
        issue_taps(tp[0], RA0, RB0, r_next, w, h, fgx, clampi(row0, 0, h - 1), fl[0].x, fl[0].y);
    }
    int gy_next = clampi(row0, 0, h - 1);

    for (int g0 = 0; g0 < nrows; g0 += G) {
        if (col_thread) {
#pragma unroll
            for (int i = 0; i < G; ++i) {
                const int ri = g0 + i;
                const int gy = gy_next;
                // taps of the next row go out before this row's arithmetic
                {
                    const int in = (i + 1) % PD;
                    gy_next = clampi(row0 + ri + 1, 0, h - 1);
                    issue_taps(tp[(i + 1) & 1], RA0, RB0, r_next, w, h, fgx, gy_next, fl[in].x, fl[in].y);
                }
                const Taps& g = tp[i & 1];
                const float dx = fl[i % PD].x, dy = fl[i % PD].y;
                const float4 a = na;
                const float b = nb;
                {
                    const int o = gy_next * w + gx;      // R0 of the next row
                    na = RA0[o];
                    nb = RB0[o];
                }
                // flow of the row PD steps ahead reuses this row's registers
                if (fin) fl[i % PD] = fin[clampi(row0 + ri + PD, 0, h - 1) * w + gx];
                float r2, r3, r4, r5, r6;
                int wx1, wy1;
                float fx, fy;
                bool inb;
                warp_point(fgx, gy, dx, dy, w, h, wx1, wy1, fx, fy, inb);
                {
                    const float a01 = fx * (1.f - fy), a10 = (1.f - fx) * fy, a11 = fx * fy, a00 = (1.f - fx) * (1.f - fy);
                    r2 = a00 * g.q00.x + a01 * g.q01.x + a10 * g.q10.x + a11 * g.q11.x;
                    r3 = a00 * g.q00.y + a01 * g.q01.y + a10 * g.q10.y + a11 * g.q11.y;
                    r4 = a00 * g.q00.z + a01 * g.q01.z + a10 * g.q10.z + a11 * g.q11.z;
                    r5 = a00 * g.q00.w + a01 * g.q01.w + a10 * g.q10.w + a11 * g.q11.w;
                    r6 = a00 * g.s00 + a01 * g.s01 + a10 * g.s10 + a11 * g.s11;
                    r4 = (a.z + r4) * 0.5f;
                    r5 = (a.w + r5) * 0.5f;
                    r6 = (b + r6) * 0.25f;
                }
                if (!inb) {
                    r2 = r3 = 0.f;
                    r4 = a.z; r5 = a.w; r6 = b * 0.5f;
                }
                r2 = (a.x - r2) * 0.5f;
                r3 = (a.y - r3) * 0.5f;
                r2 += r4 * dy + r6 * dx;
                r3 += r6 * dy + r5 * dx;
                if (x_edge || (unsigned)(gy - 5) >= (unsigned)(h - 10)) {
                    const float sc = bxs * (gy < 5 ? p.border[gy] : 1.f) * (gy >= h - 5 ? p.border[h - gy - 1] : 1.f);
                    r2 *= sc; r3 *= sc; r4 *= sc; r5 *= sc; r6 *= sc;
                }
                const float m0 = r4 * r4 + r6 * r6, m1 = (r4 + r5) * r6, m2 = r5 * r5 + r6 * r6;
                const float m3 = r4 * r2 + r6 * r3, m4 = r6 * r2 + r5 * r3;
                if (ri < nrows) {
                    // vertical running sums: retire the row that leaves the window
                    // float64 sums of <= 2R+1 float32 terms are exact (no rounding until the spread of the
                    // terms exceeds 2^29), so the result does not depend on where the walk started:
                    // any partition of the rows into strips gives the same bits
                    float* rs = ring + slot * (5 * CW) + t;
                    cs0 -= (double)rs[0 * CW]; cs1 -= (double)rs[1 * CW]; cs2 -= (double)rs[2 * CW];
                    cs3 -= (double)rs[3 * CW]; cs4 -= (double)rs[4 * CW];
                    rs[0 * CW] = m0; rs[1 * CW] = m1; rs[2 * CW] = m2; rs[3 * CW] = m3; rs[4 * CW] = m4;
                    cs0 += (double)m0; cs1 += (double)m1; cs2 += (double)m2; cs3 += (double)m3; cs4 += (double)m4;
                    slot = slot + 1 == K ? 0 : slot + 1;
                    if (ri >= 2 * R) {
                        float* hd = hand + i * (5 * CP) + t;
                        hd[0 * CP] = (float)cs0; hd[1 * CP] = (float)cs1; hd[2 * CP] = (float)cs2;
                        hd[3 * CP] = (float)cs3; hd[4 * CP] = (float)cs4;
                    }
                }
            }
        }
        __syncthreads();
        solve_rows<R, TW, NT, G, MINMAX>(p, hand, fout, t, g0, nrows, x0, ys, w, lmin, lmax);
        __syncthreads();
    }
    if (MINMAX) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lmin = fminf(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
            lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
        }
        if ((t & 31) == 0) {
            atomicMin(p.minmax + 2 * pair, __float_as_uint(lmin));
            atomicMax(p.minmax + 2 * pair + 1, __float_as_uint(lmax));
        }
    }
    }   // segments of this CTA's row range
}

// ---------------------------------------------------------------------------
// K5+K6 fused, strip walk with the ring in TENSOR MEMORY and asynchronous tap landing.
//
// Same algorithm and results as flow_iter_strip_kernel; what changes is where state lives:
//   * the (2R+1)-row ring of M that a column thread keeps is thread-private and only ever
//     read back by its writer -- exactly the access pattern of tcgen05.ld/st.32x32b (lane =
//     thread, columns = words).  It moves from shared memory (42 KB per 128 columns) to TMEM
//     (256 KB per SM, otherwise idle on this path): 8 columns per ring slot, 120 per thread;
//   * the shared memory that frees up becomes a landing zone for cp.async (LDGSTS): the 8
//     bilinear taps of R1 and R0 of a row are requested two rows ahead without occupying
//     registers, so every thread keeps two rows of gathers in flight where the register
//     version keeps one.  Flow rows are pre-loaded four rows ahead in registers (8 B).
// Geometry: 256 threads = 254 halo columns of a 240-column strip (1920 = 8 x 240, halo
// 1.06x); 2 CTAs per SM x 256 TMEM columns = all 512 columns.
// ---------------------------------------------------------------------------

// UPS != 0: flow_in is the COARSER level's flow (1 general ratio, 2 exact x2); every flow row is up-sampled on the fly (bilerp_flow, the expression of
// flow_upsample_kernel), so the first iteration of a level needs no up-sample launch and no full-size scratch field.
//
// SHARE: the two lower taps of a row's warped sample are the two upper taps of the next row's whenever the flow is smooth
// (the next row's sample sits exactly one row further down: same clamped address + w).  The lower taps therefore land in
// a ring of their own, four rows deep, and a row requests its upper taps only when that test fails (rows at motion
// boundaries, replicated border rows, the first row of a segment); otherwise it reads the previous row's lower taps.
// Six cp.async per row instead of ten, 40 % fewer tap requests to L1 -- same bytes, same arithmetic, same results
// (bit-identical in the emulated and GPU tests); measured r02z: the three level-0 launches of 32 pairs 3.30 -> 3.15 ms,
// level 1 0.894 -> 0.859 ms.  OFC_TMEM_SHARE=0 selects the form without sharing.
// (Sharing in the x direction as well -- the right taps read from the right neighbour's slots after a shuffle of the tap
// addresses, two __syncwarp per row -- measured slower: 3.21 against 3.05 ms, r03a.)
template <int R, int TW, int NT, int G, bool MINMAX, int UPS, int SHARE = 0>
__global__ void __launch_bounds__(NT, 2) flow_iter_tmem_kernel(IterParams p, int n_cols, int64_t total_rows) {
    constexpr int K = 2 * R + 1;
    constexpr int CW = TW + 2 * R;
    constexpr int CP = (CW + 3) / 4 * 4 + 4;
    constexpr int S = 3;                                 // landing slots: rows r, r+1, r+2
    constexpr int TCOLS = 256;                           // TMEM columns per CTA (two warps share a lane quarter)
    static_assert(NT == 256 && CW <= NT, "8 warps, one thread per halo column");
    static_assert(K * 8 <= 128, "ring slots of 8 columns must fit the thread's 128 columns");
    static_assert(G == 4, "the flow pre-load ring is indexed by the unrolled row number");
    OFC_DYN_SMEM(float, sm);
    float* hand = sm;                                                    // [G][5][CP]
    // !SHARE: [S][4][NT] taps of R1 per row.  SHARE: land_q = [SB][2][NT] lower taps (row & 3) then [S][2][NT] upper taps
    // of the rows that could not share; land_s alike
    constexpr int SB = 4;
    constexpr int QROWS = SHARE ? SB * 2 + S * 2 : S * 4;
    float4* land_q = reinterpret_cast<float4*>(sm + G * 5 * CP);         // taps of R1 (RA part)
    float4* land_a = land_q + QROWS * NT;                                // [S][NT]    R0 (RA part)
    float* land_s = reinterpret_cast<float*>(land_a + S * NT);           // taps of R1 (RB part)
    float* land_b = land_s + QROWS * NT;                                 // [S][NT]    R0 (RB part)
    float4* land_qt = land_q + SB * 2 * NT;                              // SHARE: upper taps, [S][2][NT]
    float* land_st = land_s + SB * 2 * NT;
    __shared__ unsigned s_tmem_base;

    const int t = threadIdx.x;
    const int w = p.w, h = p.h;
#ifndef OFC_EMULATE
    if (t < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "n"(TCOLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    // this thread's ring: lane quarter of its warp, upper or lower half of the CTA's columns
    const unsigned ring_base = s_tmem_base + ((unsigned)(((t >> 5) & 3) * 32) << 16) + (unsigned)((t >> 7) * 128);
#else
    const unsigned ring_base = 0;
    if (t == 0) s_tmem_base = 0;
#endif

    const int64_t range_lo = total_rows * blockIdx.x / gridDim.x;
    const int64_t range_hi = total_rows * (blockIdx.x + 1) / gridDim.x;
    for (int64_t cur = range_lo; cur < range_hi;) {
    const int unit = (int)(cur / h);
    const int ys = (int)(cur - (int64_t)unit * h);
    const int64_t left = range_hi - cur;
    const int ye = left < (int64_t)(h - ys) ? ys + (int)left : h;
    cur += ye - ys;
    const int pair = unit / n_cols;
    const int x0 = (unit - pair * n_cols) * TW;
    const float4* __restrict__ RA0 = p.RA + (int64_t)pair * p.r_stride;
    const float* __restrict__ RB0 = p.RB + (int64_t)pair * p.r_stride;
    const int r_next = (int)p.r_next;
    const float2* __restrict__ fin = p.flow_in ? p.flow_in + (int64_t)pair * p.flow_in_stride : nullptr;
    float2* __restrict__ fout = p.flow_out + (int64_t)pair * p.flow_out_stride;

    const int gx = clampi(x0 - R + t, 0, w - 1);
    // columns past the frame's replicated border (only in a strip that sticks out of the frame) carry no
    // data: they skip all loads -- hundreds of lanes copying one address serialise in the LDGSTS path
    const bool live = x0 - R + t <= w - 1 + R;
    const float bxs = (gx < 5 ? p.border[gx] : 1.f) * (gx >= w - 5 ? p.border[w - gx - 1] : 1.f);
    const bool x_edge = (unsigned)(gx - 5) >= (unsigned)(w - 10);
    const float fgx = (float)gx;
    const int nrows = (ye - ys) + 2 * R;
    const int row0 = ys - R;

    double cs0 = 0.0, cs1 = 0.0, cs2 = 0.0, cs3 = 0.0, cs4 = 0.0;
    float lmin = 3.402823466e38f, lmax = 0.f;
    {   // all-zero ring: every row retires "the row K steps back" unconditionally
        const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int q = 0; q < K; ++q) tmem_st8(ring_base + q * 8, z);
        tmem_wait_st();
    }
    int slot = 0;                                        // ring slot of the current row
    int ls = 0;                                          // landing slot of the current row (row index mod S)

    // request the taps of R1 and R0 for `row` (already clamped) into landing slot `dst_slot`; SHARE: `bslot` = the row's
    // position in the ring of lower taps (its un-clamped number & 3)
    int o1_last = -0x40000000;                           // SHARE: tap address of the previous request (none yet)
    unsigned shq = 0;                                    // SHARE: "upper taps shared" flags of the last requests, newest in bit 0
    auto request_row = [&](int row, float dx, float dy, int dst_slot, int bslot) {
        int x1, y1;
        float fx, fy;
        bool inb;
        warp_point(fgx, row, dx, dy, w, h, x1, y1, fx, fy, inb);
        const int o1 = clampi(y1, 0, h - 2) * w + clampi(x1, 0, w - 2) + r_next;
        if (!live) { cp_async_commit(); return; }
        const float4* pa = RA0 + o1;
        const float* pb = RB0 + o1;
        // R0 first (it never depends on the flow: 3.06 -> 3.03 ms, r03c), past L1 (.cg: it is read once per column and
        // L1 is needed for the four-fold reuse of the taps of R1; 3.50 -> 3.42 ms, r02o)
        const int o0 = row * w + gx;
        cp_async16_cg(land_a + dst_slot * NT + t, RA0 + o0);
        cp_async4(land_b + dst_slot * NT + t, RB0 + o0);
        if (SHARE) {
            const bool sh = o1 == o1_last + w;
            o1_last = o1;
            shq = (shq << 1) | (sh ? 1u : 0u);
            const auto lqb = smem_u32(land_q + (bslot * 2) * NT + t);
            const auto lsb = smem_u32(land_s + (bslot * 2) * NT + t);
            cp_async16_o<0, 0>(lqb, pa + w);
            cp_async4_o<0, 0>(lsb, pb + w);
            cp_async16_o<NT * 16, 16>(lqb, pa + w);
            cp_async4_o<NT * 4, 4>(lsb, pb + w);
            if (!sh) {
                const auto lqt = smem_u32(land_qt + (dst_slot * 2) * NT + t);
                const auto lst = smem_u32(land_st + (dst_slot * 2) * NT + t);
                cp_async16_o<0, 0>(lqt, pa); cp_async16_o<NT * 16, 16>(lqt, pa);
                cp_async4_o<0, 0>(lst, pb); cp_async4_o<NT * 4, 4>(lst, pb);
            }
        } else {
            const auto lq = smem_u32(land_q + (dst_slot * 4) * NT + t);
            const auto lsb = smem_u32(land_s + (dst_slot * 4) * NT + t);
            cp_async16_o<0, 0>(lq, pa); cp_async16_o<NT * 16, 16>(lq, pa);
            cp_async16_o<2 * NT * 16, 0>(lq, pa + w); cp_async16_o<3 * NT * 16, 16>(lq, pa + w);
            cp_async4_o<0, 0>(lsb, pb); cp_async4_o<NT * 4, 4>(lsb, pb);
            cp_async4_o<2 * NT * 4, 0>(lsb, pb + w); cp_async4_o<3 * NT * 4, 4>(lsb, pb + w);
        }
        cp_async_commit();
    };

    // flow of one (clamped) row at this thread's column
    int ups_xi = 0, ups_xj = 0;
    float ups_fx = 0.f;
    if (UPS) {
        src_coord(gx, p.usx, p.wc, ups_xi, ups_fx);
        ups_xj = min(ups_xi + 1, p.wc - 1);
    }
    // the level is exactly twice the coarser one and 1/pyr_scale is a power of two (pyr_scale = 0.5 on frames whose
    // sides divide): the row's resize coordinate is integer arithmetic and the scaling a float multiply -- the same
    // bits as src_coord / the float64 multiply of bilerp_flow (exact operations), without their conversions
    constexpr bool ups_fast = UPS == 2;
    const float ups_mul = (float)p.flow_mul;
    auto flow_at = [&](int row) -> float2 {
        if (!UPS) return fin[row * w + gx];
        int yi;
        float fy;
        if (ups_fast) {
            yi = (row - 1) >> 1;
            fy = (row & 1) ? 0.25f : 0.75f;
            if (row == 0) { yi = 0; fy = 0.f; }
            if (yi >= p.hc - 1) { yi = p.hc - 1; fy = 0.f; }
        } else {
            src_coord(row, p.usy, p.hc, yi, fy);
        }
        const int yj = min(yi + 1, p.hc - 1);
        const float2* r0p = fin + yi * p.wc;
        const float2* r1p = fin + yj * p.wc;
        if (ups_fast) {
            const float2 r = bilerp_raw(r0p[ups_xi], r0p[ups_xj], r1p[ups_xi], r1p[ups_xj], ups_fx, fy);
            return make_float2(r.x * ups_mul, r.y * ups_mul);
        }
        return bilerp_flow(r0p[ups_xi], r0p[ups_xj], r1p[ups_xi], r1p[ups_xj], ups_fx, fy, p.flow_mul);
    };

    // ---- prologue: flow of rows 0..3 in registers, requests for rows 0 and 1 ------------
    float2 fl[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
        fl[i] = (fin && live) ? flow_at(clampi(row0 + i, 0, h - 1)) : make_float2(0.f, 0.f);
    request_row(clampi(row0, 0, h - 1), fl[0].x, fl[0].y, 0, 0);
    request_row(clampi(row0 + 1, 0, h - 1), fl[1].x, fl[1].y, 1, 1);

    for (int g0 = 0; g0 < nrows; g0 += G) {
#pragma unroll
        for (int i = 0; i < G; ++i) {
            const int ri = g0 + i;
            const int gy = clampi(row0 + ri, 0, h - 1);
            // row ri+2 goes out first; its flow was loaded four steps ago
            {
                const int l2 = ls + 2 >= S ? ls + 2 - S : ls + 2;
                request_row(clampi(row0 + ri + 2, 0, h - 1), fl[(i + 2) & 3].x, fl[(i + 2) & 3].y, l2, (i + 2) & 3);
            }
            const float dx = fl[i].x, dy = fl[i].y;
            if (fin && live) fl[i] = flow_at(clampi(row0 + ri + 4, 0, h - 1));        // (ahead of the requests: slower, r03c)
            // old ring row (leaves the window) -- TMEM read overlaps the wait for the landing zone
            float old[8];
            // (issuing this load before the row's requests and waiting for it here measured slower, r02s)
            tmem_wait_st();                              // last row's ring store
            tmem_ld8(ring_base + slot * 8, old);
            cp_async_wait<2>();                          // everything but the two newest requests has landed
            float4 q00, q01, q10, q11;
            float s00, s01, s10, s11;
            if (SHARE) {
                // this row's flag went in two requests ago (rows ri + 1 and ri + 2 followed)
                const bool sh = (shq >> 2) & 1u;
                const int top = sh ? (((i + 3) & 3) * 2) * NT + t : (SB * 2 + ls * 2) * NT + t;       // previous row's lower taps, or its own upper ones
                const int low = ((i & 3) * 2) * NT + t;
                q00 = land_q[top]; q01 = land_q[top + NT];
                s00 = land_s[top]; s01 = land_s[top + NT];
                q10 = land_q[low]; q11 = land_q[low + NT];
                s10 = land_s[low]; s11 = land_s[low + NT];
            } else {
                q00 = land_q[(ls * 4 + 0) * NT + t]; q01 = land_q[(ls * 4 + 1) * NT + t];
                q10 = land_q[(ls * 4 + 2) * NT + t]; q11 = land_q[(ls * 4 + 3) * NT + t];
                s00 = land_s[(ls * 4 + 0) * NT + t]; s01 = land_s[(ls * 4 + 1) * NT + t];
                s10 = land_s[(ls * 4 + 2) * NT + t]; s11 = land_s[(ls * 4 + 3) * NT + t];
            }
            const float4 a = land_a[ls * NT + t];
            const float b = land_b[ls * NT + t];
            float r2, r3, r4, r5, r6;
            int wx1, wy1;
            float fx, fy;
            bool inb;
            warp_point(fgx, gy, dx, dy, w, h, wx1, wy1, fx, fy, inb);
            {
                const float a01 = fx * (1.f - fy), a10 = (1.f - fx) * fy, a11 = fx * fy, a00 = (1.f - fx) * (1.f - fy);
                r2 = a00 * q00.x + a01 * q01.x + a10 * q10.x + a11 * q11.x;
                r3 = a00 * q00.y + a01 * q01.y + a10 * q10.y + a11 * q11.y;
                r4 = a00 * q00.z + a01 * q01.z + a10 * q10.z + a11 * q11.z;
                r5 = a00 * q00.w + a01 * q01.w + a10 * q10.w + a11 * q11.w;
                r6 = a00 * s00 + a01 * s01 + a10 * s10 + a11 * s11;
                r4 = (a.z + r4) * 0.5f;
                r5 = (a.w + r5) * 0.5f;
                r6 = (b + r6) * 0.25f;
            }
            if (!inb) {
                r2 = r3 = 0.f;
                r4 = a.z; r5 = a.w; r6 = b * 0.5f;
            }
            r2 = (a.x - r2) * 0.5f;
            r3 = (a.y - r3) * 0.5f;
            r2 += r4 * dy + r6 * dx;
            r3 += r6 * dy + r5 * dx;
            if (x_edge || (unsigned)(gy - 5) >= (unsigned)(h - 10)) {
                const float sc = bxs * (gy < 5 ? p.border[gy] : 1.f) * (gy >= h - 5 ? p.border[h - gy - 1] : 1.f);
                r2 *= sc; r3 *= sc; r4 *= sc; r5 *= sc; r6 *= sc;
            }
            float m[8];
            m[0] = r4 * r4 + r6 * r6; m[1] = (r4 + r5) * r6; m[2] = r5 * r5 + r6 * r6;
            m[3] = r4 * r2 + r6 * r3; m[4] = r6 * r2 + r5 * r3;
            m[5] = m[6] = m[7] = 0.f;
            {
                // float64 sums of <= 2R+1 float32 terms are exact: the result does not depend on
                // where the walk started (see flow_iter_strip_kernel).  No row tests here: rows before the window is
                // full (ri < 2R) and rows past the end of the segment hand over values that solve_rows skips, and the
                // sums are re-initialised with the next segment -- straight-line code schedules better than two branches
                cs0 -= (double)old[0]; cs1 -= (double)old[1]; cs2 -= (double)old[2]; cs3 -= (double)old[3]; cs4 -= (double)old[4];
                cs0 += (double)m[0]; cs1 += (double)m[1]; cs2 += (double)m[2]; cs3 += (double)m[3]; cs4 += (double)m[4];
                float* hd = hand + i * (5 * CP) + t;
                hd[0 * CP] = (float)cs0; hd[1 * CP] = (float)cs1; hd[2 * CP] = (float)cs2;
                hd[3 * CP] = (float)cs3; hd[4 * CP] = (float)cs4;
            }
            // the store is warp-collective (rows past the end rewrite the slot with values nobody reads)
            tmem_st8(ring_base + slot * 8, m);
            slot = slot + 1 == K ? 0 : slot + 1;
            ls = ls + 1 == S ? 0 : ls + 1;
        }
        __syncthreads();
        solve_rows<R, TW, NT, G, MINMAX>(p, hand, fout, t, g0, nrows, x0, ys, w, lmin, lmax);
        __syncthreads();
    }
    cp_async_wait<0>();                                  // drain requests that ran past the segment
    tmem_wait_st();
    if (MINMAX) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lmin = fminf(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
            lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
        }
        if ((t & 31) == 0) {
            atomicMin(p.minmax + 2 * pair, __float_as_uint(lmin));
            atomicMax(p.minmax + 2 * pair + 1, __float_as_uint(lmax));
        }
    }
    }   // segments of this CTA's row range
#ifndef OFC_EMULATE
    __syncthreads();
    if (t < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem_base), "n"(TCOLS));
#endif
}

// ---------------------------------------------------------------------------
// OPTFLOW_FARNEBACK_GAUSSIAN (cv2 flag 256; FarnebackUpdateFlow_GaussianBlur): the same fused
// iteration with a separable Gaussian window of 2R+1 taps instead of the box.  Square tiles
// (32 x 16 outputs + R halo), R at run time.  OpenCV's order of operations is kept: vertical pass
// first, symmetric pairs added before the multiply, float32 throughout without contraction, rows and
// columns replicated at the border (M evaluated at the clamped coordinate), solve in float64.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void eval_update_matrix(const IterParams& p, const float4* __restrict__ RA0, const float* __restrict__ RB0,
                                                   const float4* __restrict__ RA1, const float* __restrict__ RB1,
                                                   const float2* __restrict__ fin, int gx, int gy, float (&m)[5]) {
    const int w = p.w, h = p.h;
    const float2 fl = load_flow(p, fin, gx, gy);
    const float dx = fl.x, dy = fl.y;
    float fx = (float)gx + dx, fy = (float)gy + dy;
    const int x1 = (int)floorf(fx), y1 = (int)floorf(fy);
    fx -= (float)x1; fy -= (float)y1;
    const int64_t o0 = (int64_t)gy * w + gx;
    const float4 a = RA0[o0];
    const float b = RB0[o0];
    float r2, r3, r4, r5, r6;
    if ((unsigned)x1 < (unsigned)(w - 1) && (unsigned)y1 < (unsigned)(h - 1)) {
        const float a01 = fx * (1.f - fy), a10 = (1.f - fx) * fy, a11 = fx * fy, a00 = (1.f - fx) * (1.f - fy);
        const int64_t o1 = (int64_t)y1 * w + x1;
        const float4 q00 = RA1[o1], q01 = RA1[o1 + 1], q10 = RA1[o1 + w], q11 = RA1[o1 + w + 1];
        const float s00 = RB1[o1], s01 = RB1[o1 + 1], s10 = RB1[o1 + w], s11 = RB1[o1 + w + 1];
        r2 = a00 * q00.x + a01 * q01.x + a10 * q10.x + a11 * q11.x;
        r3 = a00 * q00.y + a01 * q01.y + a10 * q10.y + a11 * q11.y;
        r4 = a00 * q00.z + a01 * q01.z + a10 * q10.z + a11 * q11.z;
        r5 = a00 * q00.w + a01 * q01.w + a10 * q10.w + a11 * q11.w;
        r6 = a00 * s00 + a01 * s01 + a10 * s10 + a11 * s11;
        r4 = (a.z + r4) * 0.5f;
        r5 = (a.w + r5) * 0.5f;
        r6 = (b + r6) * 0.25f;
    } else {
        r2 = r3 = 0.f;
        r4 = a.z; r5 = a.w; r6 = b * 0.5f;
    }
    r2 = (a.x - r2) * 0.5f;
    r3 = (a.y - r3) * 0.5f;
    r2 += r4 * dy + r6 * dx;
    r3 += r6 * dy + r5 * dx;
    if ((unsigned)(gx - 5) >= (unsigned)(w - 10) || (unsigned)(gy - 5) >= (unsigned)(h - 10)) {
        const float sc = (gx < 5 ? p.border[gx] : 1.f) * (gx >= w - 5 ? p.border[w - gx - 1] : 1.f) *
                         (gy < 5 ? p.border[gy] : 1.f) * (gy >= h - 5 ? p.border[h - gy - 1] : 1.f);
        r2 *= sc; r3 *= sc; r4 *= sc; r5 *= sc; r6 *= sc;
    }
    m[0] = r4 * r4 + r6 * r6;
    m[1] = (r4 + r5) * r6;
    m[2] = r5 * r5 + r6 * r6;
    m[3] = r4 * r2 + r6 * r3;
    m[4] = r6 * r2 + r5 * r3;
}

__global__ void __launch_bounds__(256) flow_iter_gauss_kernel(IterParams p, GaussWindow gw) {
    constexpr int TW = 32, TH = 16, NT = 256;
    const int R = gw.r;
    const int MW = TW + 2 * R, MH = TH + 2 * R;
    OFC_DYN_SMEM(float, sm);                       // M [5][MH][MW], then V [5][TH][MW]
    float* sv = sm + 5 * MH * MW;
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH, pair = blockIdx.z;
    const int w = p.w, h = p.h;
    const float4* RA0 = p.RA + (int64_t)pair * p.r_stride;
    const float* RB0 = p.RB + (int64_t)pair * p.r_stride;
    const float4* RA1 = RA0 + p.r_next;
    const float* RB1 = RB0 + p.r_next;
    const float2* fin = p.flow_in ? p.flow_in + (int64_t)pair * p.flow_in_stride : nullptr;
    for (int i = tid; i < MH * MW; i += NT) {
        const int my = i / MW, mx = i - my * MW;
        float m[5];
        eval_update_matrix(p, RA0, RB0, RA1, RB1, fin, clampi(x0 - R + mx, 0, w - 1), clampi(y0 - R + my, 0, h - 1), m);
#pragma unroll
        for (int c = 0; c < 5; ++c) sm[(c * MH + my) * MW + mx] = m[c];
    }
    __syncthreads();
    for (int i = tid; i < 5 * TH * MW; i += NT) {
        const int c = i / (TH * MW), r = i - c * TH * MW, ty = r / MW, mx = r - ty * MW;
        const float* col = sm + (c * MH + ty + R) * MW + mx;
        float v = __fmul_rn(col[0], gw.k[0]);
        for (int t = 1; t <= R; ++t) v = __fadd_rn(v, __fmul_rn(__fadd_rn(col[t * MW], col[-t * MW]), gw.k[t]));
        sv[i] = v;
    }
    __syncthreads();
    float lmin = 3.402823466e38f, lmax = 0.f;
    for (int i = tid; i < TH * TW; i += NT) {
        const int ty = i / TW, tx = i - ty * TW;
        const int gx = x0 + tx, gy = y0 + ty;
        if (gx >= w || gy >= h) continue;
        float S[5];
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            const float* row = sv + (c * TH + ty) * MW + tx + R;
            float v = __fmul_rn(row[0], gw.k[0]);
            for (int t = 1; t <= R; ++t) v = __fadd_rn(v, __fmul_rn(__fadd_rn(row[t], row[-t]), gw.k[t]));
            S[c] = v;
        }
        const double g11 = S[0] * gw.scale, g12 = S[1] * gw.scale, g22 = S[2] * gw.scale, h1 = S[3] * gw.scale, h2 = S[4] * gw.scale;
        const double idet = 1.0 / (g11 * g22 - g12 * g12 + 1e-3);
        const float2 res = make_float2((float)((g11 * h2 - g12 * h1) * idet), (float)((g22 * h1 - g12 * h2) * idet));
        p.flow_out[(int64_t)pair * p.flow_out_stride + (int64_t)gy * w + gx] = res;
        if (p.minmax) {
            const float m = sqrtf(__fmaf_rn(res.x, res.x, __fmul_rn(res.y, res.y)));
            lmin = fminf(lmin, m);
            lmax = fmaxf(lmax, m);
        }
    }
    if (p.minmax) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lmin = fminf(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
            lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
        }
        if ((tid & 31) == 0) {
            atomicMin(p.minmax + 2 * pair, __float_as_uint(lmin));
            atomicMax(p.minmax + 2 * pair + 1, __float_as_uint(lmax));
        }
    }
}

// ---------------------------------------------------------------------------
// OPTFLOW_USE_INITIAL_FLOW (cv2 flag 4): the caller's full-resolution flow seeds the coarsest level
// through cv::resize(INTER_AREA) times the level's scale.  One thread per output pixel walks the
// area tables in OpenCV's order (whole-number ratios: float32 box sum in row-major order times
// 1/area; fractional ratios: per source row a float32 weighted row sum, then times the row weight).
// ---------------------------------------------------------------------------
struct AreaSpan { int first, mid0, mid1, last; float wf, wm, wl; };     // first/last = -1 when absent
__device__ __forceinline__ AreaSpan area_span(int dx, double scale, int ssize) {
    AreaSpan a;
    const double f1 = dx * scale, f2 = f1 + scale;
    const double cell = fmin(scale, (double)ssize - f1);
    int s1 = (int)ceil(f1), s2 = (int)floor(f2);
    s2 = min(s2, ssize - 1);
    s1 = min(s1, s2);
    a.first = (s1 - f1 > 1e-3) ? s1 - 1 : -1;
    a.wf = (float)((s1 - f1) / cell);
    a.mid0 = s1; a.mid1 = s2;
    a.wm = (float)(1.0 / cell);
    a.last = (f2 - s2 > 1e-3) ? s2 : -1;
    a.wl = (float)(fmin(fmin(f2 - s2, 1.0), cell) / cell);
    return a;
}

__global__ void __launch_bounds__(256) flow_area_seed_kernel(const float2* __restrict__ src, int64_t src_stride, int W, int H,
                                                             float2* __restrict__ dst, int64_t dst_stride, int w, int h, float mul) {
    const int dx = blockIdx.x * 32 + (threadIdx.x & 31), dy = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (dx >= w || dy >= h) return;
    const float2* S = src + (int64_t)blockIdx.z * src_stride;
    const double sx = (double)W / w, sy = (double)H / h;
    const int ix = (int)nearbyint(sx), iy = (int)nearbyint(sy);
    float ax = 0.f, ay = 0.f;
    if (fabs(sx - ix) < 2.2e-16 && fabs(sy - iy) < 2.2e-16) {
        for (int a = 0; a < iy; ++a)
            for (int b = 0; b < ix; ++b) {
                const float2 v = S[(int64_t)(dy * iy + a) * W + dx * ix + b];
                ax = __fadd_rn(ax, v.x); ay = __fadd_rn(ay, v.y);
            }
        const float inv = (float)(1.0 / (ix * iy));
        ax = __fmul_rn(ax, inv); ay = __fmul_rn(ay, inv);
    } else {
        const AreaSpan cx = area_span(dx, sx, W), cy = area_span(dy, sy, H);
        auto row_sum = [&](int yy, float beta) {
            const float2* r = S + (int64_t)yy * W;
            float bx = 0.f, by = 0.f;
            if (cx.first >= 0) { bx = __fadd_rn(bx, __fmul_rn(r[cx.first].x, cx.wf)); by = __fadd_rn(by, __fmul_rn(r[cx.first].y, cx.wf)); }
            for (int xx = cx.mid0; xx < cx.mid1; ++xx) { bx = __fadd_rn(bx, __fmul_rn(r[xx].x, cx.wm)); by = __fadd_rn(by, __fmul_rn(r[xx].y, cx.wm)); }
            if (cx.last >= 0) { bx = __fadd_rn(bx, __fmul_rn(r[cx.last].x, cx.wl)); by = __fadd_rn(by, __fmul_rn(r[cx.last].y, cx.wl)); }
            ax = __fadd_rn(ax, __fmul_rn(bx, beta)); ay = __fadd_rn(ay, __fmul_rn(by, beta));
        };
        if (cy.first >= 0) row_sum(cy.first, cy.wf);
        for (int yy = cy.mid0; yy < cy.mid1; ++yy) row_sum(yy, cy.wm);
        if (cy.last >= 0) row_sum(cy.last, cy.wl);
    }
    dst[(int64_t)blockIdx.z * dst_stride + (int64_t)dy * w + dx] = make_float2(__fmul_rn(ax, mul), __fmul_rn(ay, mul));
}

__global__ void minmax_init_kernel(unsigned* mm, int n_pairs) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_pairs) {
        mm[2 * i] = 0x7f7fffffu;   // FLT_MAX
        mm[2 * i + 1] = 0u;
    }
}

// ---------------------------------------------------------------------------
// host-side launchers
// ---------------------------------------------------------------------------
static int num_sms() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaDeviceProp prop;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 148;
        n = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : 148;
    }
    return n;
}

static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

int launch_prefilter(const PrefilterParams& p, int n_frames, size_t smem, void* stream) {
    static const int legacy = env_int("OFC_PREFILTER_LEGACY", 0);
    if (p.w == p.W && p.h == p.H && p.ksz == 3 && p.identity3) {
        dim3 g(cdiv(p.W, 256), cdiv(p.H, 4), n_frames);
        ProfScope prof(PK_PREFILTER, stream);
        OFC_LAUNCH(prefilter_identity3_kernel, g, dim3(256), 0, stream, p);
        OFC_CHECK_LAUNCH("prefilter_identity3");
        return OFC_OK;
    }
    // tap counts of the reference's pyramid (pyr_scale 0.5 -> 3, 9, 19 taps) have unrolled lean kernels
    if (!legacy && (p.ksz == 3 || p.ksz == 9 || p.ksz == 19 || p.ksz == 39)) {
        // window of a 32x8 tile (+4 columns: word alignment and the one-byte over-read of the tap loop)
        const int in_rows = (int)ceil(7 * p.sy) + p.ksz + 3;
        const int in_pitch = (((int)ceil(31 * p.sx) + p.ksz + 3 + 8) + 3) / 4 * 4;
        const size_t sm = (size_t)in_rows * 32 * 8 + (size_t)in_rows * in_pitch + 16;
        if (sm <= 200 * 1024) {
            dim3 g(cdiv(p.w, 32), cdiv(p.h, 8), n_frames);
            ProfScope prof(PK_PREFILTER, stream);
#define OFC_PF_TILE(KS)                                                                                                \
    {                                                                                                                  \
        OFC_SMEM_OPTIN(prefilter_tile_kernel<KS>, sm);                                                                 \
        OFC_LAUNCH(prefilter_tile_kernel<KS>, g, dim3(256), sm, stream, p, in_rows, in_pitch);                          \
    }
            if (p.ksz == 3) OFC_PF_TILE(3)
            else if (p.ksz == 9) OFC_PF_TILE(9)
            else if (p.ksz == 19) OFC_PF_TILE(19)
            else OFC_PF_TILE(39)
#undef OFC_PF_TILE
            OFC_CHECK_LAUNCH("prefilter_tile");
            return OFC_OK;
        }
    }
    dim3 grid(cdiv(p.w, p.tx), cdiv(p.h, p.ty), n_frames);
    OFC_SMEM_OPTIN(prefilter_kernel, smem);
    ProfScope prof(PK_PREFILTER, stream);
    OFC_LAUNCH(prefilter_kernel, grid, dim3(256), smem, stream, p);
    OFC_CHECK_LAUNCH("prefilter");
    return OFC_OK;
}

static size_t pyr_smem(int S, int ksz) {
    const int RR = ksz / 2, A = ((S / 2 - 1 - RR) % 4 + 4) % 4, RB = S < 8 ? 8 : S;
    const int PW = (S * (PYR_NT - 1) + A + ksz + 1 + 3) / 4;
    return (size_t)RB * PW * 4;
}

int launch_prefilter_pyramid(const PrefilterParams* levels, const size_t* smem_fallback, int n_levels, int n_frames,
                             bool* done, void* stream) {
    (void)smem_fallback;
    const int off = env_int("OFC_PREFILTER_PYR", 1) == 0;          // read per call: tests compare both forms in one process
    PyrParams pp;
    pp.n_levels = 0;
    size_t smem = 0;
    int blocks = 0;
    for (int i = 0; i < n_levels; ++i) done[i] = false;
    if (off || n_levels <= 0) return OFC_OK;
    // heaviest levels (largest S) first: their CTAs start first
    for (int pass = 16; pass >= 2; pass >>= 1) {
        for (int i = 0; i < n_levels && pp.n_levels < 4; ++i) {
            const PrefilterParams& p = levels[i];
            if (p.identity3 || p.w <= 0) continue;
            const int S = p.w > 0 ? p.W / p.w : 0;
            if (S != pass || p.W != p.w * S || p.H != p.h * S) continue;
            const int want_ksz = S == 2 ? 3 : (S == 4 ? 9 : (S == 8 ? 19 : 39));
            if (p.ksz != want_ksz || (p.W & 3) || (p.gray_stride & 3) || ((uintptr_t)p.gray & 3)) continue;
            if (p.H < 4 * S + p.ksz + 16) continue;        // a staged batch may overshoot the frame by < H rows (single reflection)
            if (pp.n_levels == 0) { pp.gray = p.gray; pp.gray_stride = p.gray_stride; pp.W = p.W; pp.H = p.H; }
            PyrLevel& L = pp.lv[pp.n_levels++];
            L.out = p.out; L.out_stride = p.out_stride; L.taps = p.taps; L.w = p.w; L.h = p.h; L.S = S;
            L.blocks_x = cdiv(p.w, PYR_NT);
            L.rows_per_seg = 128 / S < 8 ? 8 : 128 / S;
            L.segs = cdiv(p.h, L.rows_per_seg);
            L.first_block = blocks;
            blocks += L.blocks_x * L.segs;
            const size_t sm = pyr_smem(S, p.ksz);
            if (sm > smem) smem = sm;
            done[i] = true;
        }
    }
    if (pp.n_levels == 0) return OFC_OK;
    ProfScope prof(PK_PREFILTER, stream);
    OFC_SMEM_OPTIN(prefilter_pyr_kernel, smem);
    OFC_LAUNCH(prefilter_pyr_kernel, dim3(blocks, n_frames), dim3(PYR_NT), smem, stream, pp);
    OFC_CHECK_LAUNCH("prefilter_pyr");
    return OFC_OK;
}

template <int N, int TW, int NT, bool FUSE3, bool STAGE = false>
static int launch_polyexp_strip(const PolyParams& p, const unsigned char* gray, int64_t gray_stride, float* I_out,
                                int n_frames, void* stream) {
    static int resident = 0;
    if (!resident) {
        OFC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, polyexp_strip_kernel<N, TW, NT, FUSE3, STAGE>, NT, 0));
        if (resident < 1) resident = 1;
    }
    const int cols = cdiv(p.w, TW);
    const int64_t total_rows = (int64_t)n_frames * cols * p.h;
    // leave part of every SM to the coarse-level iterations that run beside this kernel on the caller's
    // stream (ofc_api.cu forks the per-frame work onto a side stream)
    static const int cap = env_int("OFC_POLY_RESIDENT", 0);
    const int use = (cap > 0 && cap < resident) ? cap : resident;
    int64_t ctas = (int64_t)num_sms() * use;
    const int64_t max_ctas = (total_rows + 23) / 24;     // keep ranges >= 24 rows (vertical halo 2N per segment)
    if (ctas > max_ctas) ctas = max_ctas;
    ProfScope prof(PK_POLYEXP, stream);
    OFC_LAUNCH((polyexp_strip_kernel<N, TW, NT, FUSE3, STAGE>), dim3((unsigned)ctas), dim3(NT), 0, stream, p, gray, gray_stride, I_out,
               cols, total_rows);
    OFC_CHECK_LAUNCH("polyexp_strip");
    return OFC_OK;
}

// gray != null: the level is full resolution with the fixed 3-tap pre-filter -> fused
int launch_polyexp(const PolyParams& p, int poly_n, int n_frames, const unsigned char* gray, int64_t gray_stride,
                   float* I_out, void* stream) {
    static const int legacy = env_int("OFC_POLYEXP_LEGACY", 0);
    if (!legacy && poly_n == 5) {
        if (gray) {
            const int stage = env_int("OFC_POLY_STAGE", 1);
            const bool aligned = p.w % 16 == 0 && gray_stride % 16 == 0 && ((uintptr_t)gray & 15) == 0 && p.h > 16;
            if (p.w > 512) {
                if (stage && aligned) return launch_polyexp_strip<5, 128, 160, true, true>(p, gray, gray_stride, I_out, n_frames, stream);
                return launch_polyexp_strip<5, 128, 160, true>(p, gray, gray_stride, I_out, n_frames, stream);
            }
            if (stage && aligned) return launch_polyexp_strip<5, 64, 96, true, true>(p, gray, gray_stride, I_out, n_frames, stream);
            return launch_polyexp_strip<5, 64, 96, true>(p, gray, gray_stride, I_out, n_frames, stream);
        }
        if (p.w > 512) return launch_polyexp_strip<5, 128, 160, false>(p, nullptr, 0, nullptr, n_frames, stream);
        return launch_polyexp_strip<5, 64, 96, false>(p, nullptr, 0, nullptr, n_frames, stream);
    }
    if (gray) {     // un-fused fallback: materialise I first
        PrefilterParams pf;
        pf.gray = gray; pf.gray_stride = gray_stride; pf.out = I_out; pf.out_stride = p.out_stride;
        pf.W = p.w; pf.H = p.h; pf.w = p.w; pf.h = p.h; pf.ksz = 3; pf.sx = pf.sy = 1.0; pf.taps = nullptr;
        pf.tx = pf.ty = pf.in_rows = pf.in_pitch = pf.taps_pad = 0; pf.identity3 = 1;
        int rc = launch_prefilter(pf, n_frames, 0, stream);
        if (rc != OFC_OK) return rc;
    }
    dim3 grid(cdiv(p.w, 64), cdiv(p.h, 32), n_frames);
    ProfScope prof(PK_POLYEXP, stream);
    if (poly_n == 5) {
        OFC_LAUNCH(polyexp_kernel<5>, grid, dim3(256), 0, stream, p);
    } else if (poly_n == 7) {
        OFC_LAUNCH(polyexp_kernel<7>, grid, dim3(256), 0, stream, p);
    } else {
        set_error("poly_n=%d unsupported (5 or 7)", poly_n);
        return OFC_ERR_UNSUPPORTED;
    }
    OFC_CHECK_LAUNCH("polyexp");
    return OFC_OK;
}

template <int R, int TH, int NT, int MINB>
static int launch_iter_r(const IterParams& p, int n_pairs, void* stream) {
    constexpr int MW = 64 + 2 * R, MH = TH + 2 * R;
    constexpr int P0 = (MW + 2 + 3) / 4 * 4;
    constexpr int PITCH = ((P0 / 4) % 2 == 0) ? P0 + 4 : P0;
    constexpr size_t smem = (size_t)5 * MH * PITCH * sizeof(float);
    OFC_SMEM_OPTIN((flow_iter_kernel<R, TH, NT, MINB>), smem);
    dim3 grid(cdiv(p.w, 64), cdiv(p.h, TH), n_pairs);
    ProfScope prof(PK_ITER_L0 + (g_prof_level < 8 ? g_prof_level : 7), stream);
    OFC_LAUNCH((flow_iter_kernel<R, TH, NT, MINB>), grid, dim3(NT), smem, stream, p);
    OFC_CHECK_LAUNCH("flow_iter");
    return OFC_OK;
}

template <int R, int TW, int NT, int G, int MINB, bool MINMAX>
static int launch_strip_rm(const IterParams& p, int n_pairs, void* stream) {
    constexpr int K = 2 * R + 1, CW = TW + 2 * R, CP = (CW + 3) / 4 * 4 + 4;
    constexpr size_t smem = (size_t)((K * 5 * CW + 3) / 4 * 4 + G * 5 * CP) * sizeof(float);
    OFC_SMEM_OPTIN((flow_iter_strip_kernel<R, TW, NT, G, MINB, MINMAX>), smem);
    // persistent grid: every SM holds as many CTAs as fit, each gets an equal share of the rows
    static int resident = 0;
    if (!resident) {
        OFC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, flow_iter_strip_kernel<R, TW, NT, G, MINB, MINMAX>, NT, smem));
        if (resident < 1) resident = 1;
    }
    const int cols = cdiv(p.w, TW);
    const int64_t total_rows = (int64_t)n_pairs * cols * p.h;
    int64_t ctas = (int64_t)num_sms() * resident;
    const int64_t max_ctas = (total_rows + 15) / 16;     // keep ranges >= 16 rows
    if (ctas > max_ctas) ctas = max_ctas;
    ProfScope prof(PK_ITER_L0 + (g_prof_level < 8 ? g_prof_level : 7), stream);
    OFC_LAUNCH((flow_iter_strip_kernel<R, TW, NT, G, MINB, MINMAX>), dim3((unsigned)ctas), dim3(NT), smem, stream, p, cols, total_rows);
    OFC_CHECK_LAUNCH("flow_iter_strip");
    return OFC_OK;
}

template <int R, int TW, int NT, int G, int MINB>
static int launch_strip_r(const IterParams& p, int n_pairs, void* stream) {
    return p.minmax ? launch_strip_rm<R, TW, NT, G, MINB, true>(p, n_pairs, stream)
                    : launch_strip_rm<R, TW, NT, G, MINB, false>(p, n_pairs, stream);
}

template <bool MINMAX, int UPS, int SHARE = 0>
static int launch_tmem(const IterParams& p, int n_pairs, void* stream) {
    constexpr int R = 7, TW = 240, NT = 256, G = 4;
    constexpr int CW = TW + 2 * R, CP = (CW + 3) / 4 * 4 + 4, S = 3;
    constexpr int QROWS = SHARE ? 4 * 2 + S * 2 : S * 4;
    constexpr size_t smem = (size_t)(G * 5 * CP) * 4 + (size_t)(QROWS + S) * NT * 16 + (size_t)(QROWS + S) * NT * 4;
    OFC_SMEM_OPTIN((flow_iter_tmem_kernel<R, TW, NT, G, MINMAX, UPS, SHARE>), smem);
    const int cols = cdiv(p.w, TW);
    const int64_t total_rows = (int64_t)n_pairs * cols * p.h;
    int64_t ctas = (int64_t)num_sms() * 2;              // 2 CTAs per SM: 2 x 256 TMEM columns
    // Frame f's expansion is read twice per launch: as R0 by pair f and through the warped taps by pair f-1.  With a
    // grid that is a multiple of the pair count every pair is cut into ranges at the same rows, so the CTAs of
    // neighbouring pairs walk the same rows at the same time and the second read hits L2 instead of DRAM (measured r02q,
    // 32 pairs at 1080p: 288 CTAs 3.30 ms against 296 CTAs 3.34 ms for the three level-0 launches, although eight CTA
    // slots stay empty).  Taken when it costs at most 1/16 of the slots.
    if (n_pairs > 1 && n_pairs <= ctas) {
        const int64_t aligned = ctas / n_pairs * n_pairs;
        if (aligned * 16 >= ctas * 15) ctas = aligned;
    }
    {
        static const int force = env_int("OFC_TMEM_CTAS", 0);          // tuning override: grid size
        if (force > 0) ctas = force;
    }
    const int64_t max_ctas = (total_rows + 15) / 16;
    if (ctas > max_ctas) ctas = max_ctas;
    ProfScope prof(PK_ITER_L0 + (g_prof_level < 8 ? g_prof_level : 7), stream);
    OFC_LAUNCH((flow_iter_tmem_kernel<R, TW, NT, G, MINMAX, UPS, SHARE>), dim3((unsigned)ctas), dim3(NT), smem, stream, p, cols, total_rows);
    OFC_CHECK_LAUNCH("flow_iter_tmem");
    return OFC_OK;
}

// The strip-walk kernel reads a plain flow field: when the input is the coarser level, up-sample
// it first into `scratch` (this level's other ping-pong buffer, not otherwise live on iteration 0).
static int launch_strip(const IterParams& p_in, int n_pairs, float2* scratch, void* stream) {
    IterParams p = p_in;
    const int use_tmem0 = env_int("OFC_ITER_TMEM", 513), any_w0 = env_int("OFC_TMEM_ANYW", 1);
    const bool tmem_path = use_tmem0 && (p.w >= use_tmem0 || p.w % 240 == 0) && (any_w0 || p.w % 240 == 0);
    // OFC_FUSE_UPSAMPLE=1: the tensor-memory walk up-samples the coarser flow itself (bit-identical).  Measured SLOWER
    // than the separate up-sample launch (r02c: +0.29 ms on the level-0 launch against 0.26 ms of up-sample kernels per
    // 32 pairs: the four coarse loads sit on the walk's critical path, two rows before the tap addresses they feed),
    // so it is off by default
    if (p.upsample && tmem_path && env_int("OFC_FUSE_UPSAMPLE", 0) && (int64_t)p.wc * p.hc < ((int64_t)1 << 30))
    {
        if (p.ups_fast) return p.minmax ? launch_tmem<true, 2, 1>(p, n_pairs, stream) : launch_tmem<false, 2, 1>(p, n_pairs, stream);
        return p.minmax ? launch_tmem<true, 1, 1>(p, n_pairs, stream) : launch_tmem<false, 1, 1>(p, n_pairs, stream);
    }
    if (p.upsample) {
        const int x2 = env_int("OFC_UPSAMPLE_X2", 1);
        const bool exact2 = x2 && p.w == 2 * p.wc && p.h == 2 * p.hc && p.w % 4 == 0 && p.usx == 0.5 && p.usy == 0.5 &&
                            p.flow_in_stride % 2 == 0 && ((uintptr_t)p.flow_in & 15) == 0;
        {
            ProfScope prof(PK_UPSAMPLE, stream);
            if (exact2) {
                OFC_LAUNCH(flow_upsample_x2_kernel, dim3(cdiv(p.w / 4, 64), cdiv(p.h / 2, 4), n_pairs), dim3(256), 0, stream, p, scratch,
                           (int64_t)p.w * p.h);
            } else {
                OFC_LAUNCH(flow_upsample_kernel, dim3(cdiv(p.w, 64), cdiv(p.h, 16), n_pairs), dim3(256), 0, stream, p, scratch,
                           (int64_t)p.w * p.h);
            }
            OFC_CHECK_LAUNCH("flow_upsample");
        }
        p.flow_in = scratch;
        p.flow_in_stride = (int64_t)p.w * p.h;
        p.upsample = 0;
    }
    static const int use_tmem = env_int("OFC_ITER_TMEM", 513);      // minimum level width; 0 = off
    // (OFC_TMEM_ANYW=0 restricts it to widths that are a multiple of its 240-column strips)
    static const int any_w = env_int("OFC_TMEM_ANYW", 1);
    if (use_tmem && (p.w >= use_tmem || p.w % 240 == 0) && (any_w || p.w % 240 == 0)) {
        const int share = env_int("OFC_TMEM_SHARE", 1);
        if (share) return p.minmax ? launch_tmem<true, 0, 1>(p, n_pairs, stream) : launch_tmem<false, 0, 1>(p, n_pairs, stream);
        return p.minmax ? launch_tmem<true, 0>(p, n_pairs, stream) : launch_tmem<false, 0>(p, n_pairs, stream);
    }
    static const int minb4 = env_int("OFC_STRIP_MINB4", 1);
    if (p.w > 512 && minb4) return launch_strip_r<7, 128, 160, 4, 4>(p, n_pairs, stream);
    if (p.w > 512) return launch_strip_r<7, 128, 160, 4, 3>(p, n_pairs, stream);
    return launch_strip_r<7, 64, 96, 4, 5>(p, n_pairs, stream);
}

int launch_flow_area_seed(const float2* src, int64_t src_stride, int W, int H, float2* dst, int64_t dst_stride, int w, int h,
                          float mul, int n_pairs, void* stream) {
    ProfScope prof(PK_UPSAMPLE, stream);
    OFC_LAUNCH(flow_area_seed_kernel, dim3(cdiv(w, 32), cdiv(h, 8), n_pairs), dim3(256), 0, stream, src, src_stride, W, H, dst,
               dst_stride, w, h, mul);
    OFC_CHECK_LAUNCH("flow_area_seed");
    return OFC_OK;
}

static int launch_iter_gauss(const IterParams& p, const GaussWindow& gw, int n_pairs, void* stream) {
    const int MW = 32 + 2 * gw.r, MH = 16 + 2 * gw.r;
    const size_t smem = (size_t)5 * (MH + 16) * MW * sizeof(float);
    if (smem > 200 * 1024) { set_error("Gaussian window of %d taps needs too much shared memory", 2 * gw.r + 1); return OFC_ERR_UNSUPPORTED; }
    OFC_SMEM_OPTIN(flow_iter_gauss_kernel, smem);
    ProfScope prof(PK_ITER_L0 + (g_prof_level < 8 ? g_prof_level : 7), stream);
    OFC_LAUNCH(flow_iter_gauss_kernel, dim3(cdiv(p.w, 32), cdiv(p.h, 16), n_pairs), dim3(256), smem, stream, p, gw);
    OFC_CHECK_LAUNCH("flow_iter_gauss");
    return OFC_OK;
}

int launch_flow_iter(const IterParams& p, int winsize, int n_pairs, float2* scratch, void* stream, const GaussWindow* gw) {
    if (gw) return launch_iter_gauss(p, *gw, n_pairs, stream);
    static int variant = -1;
    if (variant < 0) { const char* e = getenv("OFC_ITER_VARIANT"); variant = e ? atoi(e) : 0; }
    // winsize 15 (the reference's literal) on the wide levels runs the strip-walk kernel; the small
    // pyramid levels (too few rows per SM for a column walk to hide latency) and other window sizes
    // run the square-tile kernel
    static const int strip_min_w = env_int("OFC_STRIP_MIN_W", 513);
    // a narrower level also walks strips when the batch gives every persistent CTA a long enough range (>= 48 rows per
    // CTA of 2 x 148) and its width is a whole number of 240-column strips: the 480 x 270 level of a 32-pair 1080p chunk,
    // 0.38 -> 0.30 ms for its three launches (r02r; neutral before the float32 solve, r02c).  OFC_STRIP_SMALL_ROWS=0: off
    static const int small_rows = env_int("OFC_STRIP_SMALL_ROWS", 48);
    const bool wide = p.w >= strip_min_w;
    const bool batched = small_rows > 0 && p.w % 240 == 0 &&
                         (int64_t)n_pairs * (p.w / 240) * p.h >= (int64_t)2 * num_sms() * small_rows;
    if (variant == 0 && winsize == 15 && scratch != nullptr && (wide || batched))
        return launch_strip(p, n_pairs, scratch, stream);
    switch (winsize / 2) {
        case 2: return launch_iter_r<2, 32, 256, 3>(p, n_pairs, stream);
        case 3: return launch_iter_r<3, 32, 256, 3>(p, n_pairs, stream);
        case 4: return launch_iter_r<4, 32, 256, 3>(p, n_pairs, stream);
        case 5: return launch_iter_r<5, 32, 256, 3>(p, n_pairs, stream);
        case 6: return launch_iter_r<6, 32, 256, 3>(p, n_pairs, stream);
        case 7: {
            // winsize 15 (the reference's value): 64x30 tile = 73.9 KB -> 3 CTAs/SM
            if (variant == 2) return launch_iter_r<7, 48, 384, 2>(p, n_pairs, stream);
            return launch_iter_r<7, 30, 256, 3>(p, n_pairs, stream);
        }
        case 10: return launch_iter_r<10, 32, 256, 2>(p, n_pairs, stream);
        case 12: return launch_iter_r<12, 32, 256, 2>(p, n_pairs, stream);
        default: {
            // any other window: the run-time-radius kernel with unit taps (box sums in float32, scaled like OpenCV
            // by 1 / winsize^2 -- the window itself is 2 (winsize / 2) + 1 wide)
            if (winsize / 2 < 1 || winsize / 2 > 32) { set_error("winsize=%d unsupported (4..65)", winsize); return OFC_ERR_UNSUPPORTED; }
            GaussWindow box;
            box.r = winsize / 2;
            for (int i = 0; i <= box.r; ++i) box.k[i] = 1.f;
            box.scale = p.blur_scale;
            return launch_iter_gauss(p, box, n_pairs, stream);
        }
    }
}

int launch_minmax_init(unsigned* mm, int n_pairs, void* stream) {
    ProfScope prof(PK_MINMAX_INIT, stream);
    OFC_LAUNCH(minmax_init_kernel, dim3(cdiv(n_pairs, 128)), dim3(128), 0, stream, mm, n_pairs);
    OFC_CHECK_LAUNCH("minmax_init");
    return OFC_OK;
}

}  // namespace ofc
