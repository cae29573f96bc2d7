// Cosine similarity kernels and the grid-cell gather for per-cell k-means.
//
//   sliding cosine   findCosineDifferentVectors.py:5-66  (window of the long vector vs the
//                    short one, zero norm -> 0, maximum and LAST arg-max)
//   row cosine       the 1M x D "grid vectors vs query" sweep of BASELINE.json configs[4]
//                    (same formula as calculate_cosine_similarity, one row per window)
//   vector distance  computeVectorDistance.py:22-43 (true cosine, the outer-product "row 0"
//                    quirk, sum |a-b|)
//   extract cells    KmeanGrids.py:85,113 (image_dict ROI per cell) + preprocess_image :269-286
//
// The reference's vectors are integer hues; products and sums of integers below 2^53 are
// exact in float64, so dot / (sqrt(na) * sqrt(nb)) has the reference's bits (IEEE sqrt and
// division).  Reductions use a fixed order: no floating-point atomics.
#include "ofc_common.cuh"
#include "kmeans_kernels.cuh"

namespace ofc {

// order-preserving map double -> uint64 (for atomicMax on similarities)
__device__ __forceinline__ unsigned long long ordered_bits(double v) {
    unsigned long long u = (unsigned long long)__double_as_longlong(v);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double from_ordered_bits(unsigned long long u) {
    u = (u >> 63) ? (u & 0x7fffffffffffffffull) : ~u;
    return __longlong_as_double((long long)u);
}

// one thread per window offset; the short vector and its norm come from shared memory
__global__ void __launch_bounds__(256) sliding_cosine_kernel(const double* a, int n, const double* b, int64_t m,
                                                             double* sims, unsigned long long* best_bits) {
    OFC_DYN_SMEM(double, sa);                          // [n]
    __shared__ double s_na;
    for (int i = threadIdx.x; i < n; i += blockDim.x) sa[i] = a[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < n; ++i) s = fma(sa[i], sa[i], s);
        s_na = sqrt(s);
    }
    __syncthreads();
    const int64_t nwin = m - n + 1;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long mine = 0ull;
    if (i < nwin) {
        double dot = 0.0, nb = 0.0;
        for (int t = 0; t < n; ++t) {
            const double v = b[i + t];
            dot = fma(sa[t], v, dot);
            nb = fma(v, v, nb);
        }
        const double nbs = sqrt(nb);
        const double sim = (s_na == 0.0 || nbs == 0.0) ? 0.0 : dot / (s_na * nbs);
        sims[i] = sim;
        mine = ordered_bits(sim);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, mine, o);
        mine = other > mine ? other : mine;
    }
    if ((threadIdx.x & 31) == 0 && mine) atomicMax(best_bits, mine);
}

// last index whose similarity equals the maximum (findCosineDifferentVectors.py:60-61)
__global__ void __launch_bounds__(256) sliding_argmax_kernel(const double* sims, int64_t nwin,
                                                             const unsigned long long* best_bits, double* best,
                                                             long long* best_idx) {
    const double mx = from_ordered_bits(*best_bits);
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    long long mine = -1;
    if (i < nwin && sims[i] == mx) mine = i;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const long long other = __shfl_xor_sync(0xffffffffu, mine, o);
        mine = other > mine ? other : mine;
    }
    if ((threadIdx.x & 31) == 0 && mine >= 0) atomicMax(best_idx, mine);
    (void)best;
}

// the scratch word held the ordered bits of the maximum: turn it into the double in place
__global__ void sliding_finish_kernel(double* best) {
    *best = from_ordered_bits(*reinterpret_cast<unsigned long long*>(best));
}

__global__ void sliding_init_kernel(unsigned long long* best_bits, long long* best_idx) {
    *best_bits = 0ull;
    *best_idx = -1;
}

// cos(X[i,:], q) for every row; `lpr` lanes (power of two) cooperate on a row
template <typename T>
__global__ void __launch_bounds__(256) row_cosine_kernel(const T* X, int64_t n, int d, const double* q, int lpr,
                                                         double* out) {
    OFC_DYN_SMEM(double, sq);                          // [d]
    __shared__ double s_nq;
    for (int i = threadIdx.x; i < d; i += blockDim.x) sq[i] = q[i];
    __syncthreads();
    if (threadIdx.x < 32) {
        double s = 0.0;
        for (int i = threadIdx.x; i < d; i += 32) s = fma(sq[i], sq[i], s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) s_nq = sqrt(s);
    }
    __syncthreads();
    const int rows_per_block = blockDim.x / lpr;
    const int sub = threadIdx.x % lpr, rib = threadIdx.x / lpr;
    for (int64_t r0 = (int64_t)blockIdx.x * rows_per_block; r0 < n; r0 += (int64_t)gridDim.x * rows_per_block) {
        const int64_t r = r0 + rib;
        double dot = 0.0, nx = 0.0;
        if (r < n) {
            const T* row = X + r * d;
            // eight loads in flight per lane; the fma chains keep their order (t = sub, sub + lpr, ...), so the
            // result is bit for bit what the one-load-at-a-time loop gives
            int t = sub;
            for (; t + 7 * lpr < d; t += 8 * lpr) {
                T a[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) a[u] = row[t + u * lpr];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const double v = (double)a[u];
                    dot = fma(v, sq[t + u * lpr], dot);
                    nx = fma(v, v, nx);
                }
            }
            for (; t < d; t += lpr) {
                const double v = (double)row[t];
                dot = fma(v, sq[t], dot);
                nx = fma(v, v, nx);
            }
        }
        for (int o = lpr >> 1; o > 0; o >>= 1) {
            dot += __shfl_xor_sync(0xffffffffu, dot, o);
            nx += __shfl_xor_sync(0xffffffffu, nx, o);
        }
        if (r < n && sub == 0) {
            const double nxs = sqrt(nx);
            out[r] = (s_nq == 0.0 || nxs == 0.0) ? 0.0 : dot / (s_nq * nxs);
        }
    }
}

// computeVectorDistance.py on two (N,1) columns: one CTA, fixed-order reductions.
// The true cosine follows sklearn.metrics.pairwise.cosine_similarity: both vectors are
// normalised first (x / ||x||), then multiplied (computeVectorDistance.py:26).
__device__ __forceinline__ double block_sum_fixed(double v, double* s_w) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s_w[w];
    __syncthreads();
    return t;                                          // same value in every thread
}

__global__ void __launch_bounds__(256) vector_distance_kernel(const double* a, const double* b, int64_t n,
                                                              double* cos_out, double* quirk_row, double* l1) {
    __shared__ double s_w[8];
    double na = 0.0, nb = 0.0, dist = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 256) {
        const double x = a[i], y = b[i];
        na = fma(x, x, na); nb = fma(y, y, nb);
        dist += fabs(x - y);
        // row 0 of  hsv1 @ hsv2.T / (|hsv1_j| * |hsv2_j|)  for (N,1) inputs (computeVectorDistance.py:25)
        if (quirk_row) quirk_row[i] = (a[0] * y) / (sqrt(x * x) * sqrt(y * y));
    }
    const double na_s = sqrt(block_sum_fixed(na, s_w)), nb_s = sqrt(block_sum_fixed(nb, s_w));
    const double dist_t = block_sum_fixed(dist, s_w);
    const double da = na_s == 0.0 ? 1.0 : na_s, db = nb_s == 0.0 ? 1.0 : nb_s;   // sklearn: zero norms -> 1
    double dot = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 256) dot = fma(a[i] / da, b[i] / db, dot);
    const double dot_t = block_sum_fixed(dot, s_w);
    if (threadIdx.x == 0) {
        if (cos_out) *cos_out = dot_t;
        if (l1) *l1 = dist_t;
    }
}

// frame -> [cells][y_step*x_step][4] (c0, c1, c2, alpha) with the grid-line state of the
// k-means stage (row 0 / column 0 of every cell white) and preprocess_image
// applied; swap_rb reproduces read_image's BGR->RGB of color_kmeans.py:32-33.
__global__ void __launch_bounds__(256) extract_cells_kernel(const unsigned char* bgr, int64_t frame_stride, int W, int H,
                                                            int cols, int x_step, int y_step, int draw_lines,
                                                            int threshold, int swap_rb, unsigned char* out) {
    const int cell = blockIdx.x, frame = blockIdx.y;
    const int cy = cell / cols, cx = cell - cy * cols;
    const int x1 = cx * x_step, y1 = cy * y_step;
    const int n = x_step * y_step;
    const unsigned char* img = bgr + (int64_t)frame * frame_stride;
    uchar4* dst = reinterpret_cast<uchar4*>(out) + ((int64_t)frame * gridDim.x + cell) * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int ly = i / x_step, lx = i - ly * x_step;
        const unsigned char* px = img + ((int64_t)(y1 + ly) * W + (x1 + lx)) * 3;
        unsigned c0 = px[0], c1 = px[1], c2 = px[2];
        if (draw_lines && (ly == 0 || lx == 0)) c0 = c1 = c2 = 255u;
        if (swap_rb) { unsigned t = c0; c0 = c2; c2 = t; }
        if (threshold) {
            c0 = c0 < (unsigned)threshold ? 0u : c0;
            c1 = c1 < (unsigned)threshold ? 0u : c1;
            c2 = c2 < (unsigned)threshold ? 0u : c2;
        }
        // cv2.cvtColor(image, COLOR_BGR2GRAY) is applied to whatever channel order is there (Q5)
        const unsigned gray = (3735u * c0 + 19235u * c1 + 9798u * c2 + 16384u) >> 15;
        dst[i] = make_uchar4((unsigned char)c0, (unsigned char)c1, (unsigned char)c2, gray > 0 ? 255 : 0);
    }
    (void)H;
}

int launch_sliding_cosine(const double* a, int n, const double* b, int64_t m, double* sims, double* best,
                          long long* best_idx, void* stream) {
    const int64_t nwin = m - n + 1;
    ProfScope prof(PK_COSINE, stream);
    // *best holds the ordered bits of the running maximum until sliding_finish_kernel converts it
    unsigned long long* bits = reinterpret_cast<unsigned long long*>(best);
    OFC_LAUNCH(sliding_init_kernel, dim3(1), dim3(1), 0, stream, bits, best_idx);
    OFC_CHECK_LAUNCH("sliding_init");
    const unsigned blocks = (unsigned)((nwin + 255) / 256);
    OFC_LAUNCH(sliding_cosine_kernel, dim3(blocks), dim3(256), (size_t)n * sizeof(double), stream, a, n, b, m, sims, bits);
    OFC_CHECK_LAUNCH("sliding_cosine");
    OFC_LAUNCH(sliding_argmax_kernel, dim3(blocks), dim3(256), 0, stream, sims, nwin, bits, best, best_idx);
    OFC_CHECK_LAUNCH("sliding_argmax");
    OFC_LAUNCH(sliding_finish_kernel, dim3(1), dim3(1), 0, stream, best);
    OFC_CHECK_LAUNCH("sliding_finish");
    return OFC_OK;
}

int launch_row_cosine(const void* X, int dtype, int64_t n, int d, const double* q, double* out, void* stream) {
    int lpr = 1;
    while (lpr < 32 && lpr * 4 < d) lpr <<= 1;
    const int rows_per_block = 256 / lpr;
    int64_t blocks = (n + rows_per_block - 1) / rows_per_block;
    if (blocks > 148 * 8) blocks = 148 * 8;
    const size_t smem = (size_t)d * sizeof(double);
    ProfScope prof(PK_COSINE, stream);
#define OFC_ROWCOS(TT)                                                                                         \
    {                                                                                                          \
        OFC_SMEM_OPTIN(row_cosine_kernel<TT>, smem);                                                           \
        OFC_LAUNCH(row_cosine_kernel<TT>, dim3((unsigned)blocks), dim3(256), smem, stream, (const TT*)X, n, d, q, lpr, out); \
    }
    if (dtype == DT_U8) OFC_ROWCOS(unsigned char)
    else if (dtype == DT_F32) OFC_ROWCOS(float)
    else OFC_ROWCOS(double)
#undef OFC_ROWCOS
    OFC_CHECK_LAUNCH("row_cosine");
    return OFC_OK;
}

int launch_vector_distance(const double* a, const double* b, int64_t n, double* cos_out, double* quirk_row,
                           double* l1, void* stream) {
    ProfScope prof(PK_COSINE, stream);
    OFC_LAUNCH(vector_distance_kernel, dim3(1), dim3(256), 0, stream, a, b, n, cos_out, quirk_row, l1);
    OFC_CHECK_LAUNCH("vector_distance");
    return OFC_OK;
}

int launch_extract_cells(const unsigned char* bgr, int n_frames, int H, int W, int rows, int cols, int draw_lines,
                         int threshold, int swap_rb, unsigned char* out, void* stream) {
    if (n_frames <= 0) return OFC_OK;
    ProfScope prof(PK_GRID, stream);
    OFC_LAUNCH(extract_cells_kernel, dim3(rows * cols, n_frames), dim3(256), 0, stream, bgr, (int64_t)H * W * 3, W, H, cols,
               W / cols, H / rows, draw_lines, threshold, swap_rb, out);
    OFC_CHECK_LAUNCH("extract_cells");
    return OFC_OK;
}

}  // namespace ofc
