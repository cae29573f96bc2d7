// Parameter blocks and launchers of kmeans_kernels.cu / cosine_kernels.cu (internal header).
#pragma once
#include "ofc_common.cuh"

namespace ofc {

enum { DT_U8 = 0, DT_F32 = 1, DT_F64 = 2 };

static inline size_t dtype_size(int dtype) { return dtype == DT_U8 ? 1 : (dtype == DT_F32 ? 4 : 8); }

struct KmAssignParams {
    const void* X;               // [batch][n][d]
    int dtype;
    int64_t n;
    int d, k;
    const double* mean;          // [batch][d] subtracted from every row, or null
    const double* centres;       // [batch][k][d]
    const double* c2;            // [batch][k] squared norms (generic path only)
    int32_t* labels;             // [batch][n] out
    const int32_t* prev_labels;  // [batch][n] or null
    unsigned long long* n_changed;   // [batch] (+= labels that differ from prev_labels) or null
    double* inertia_partial;     // [batch][gridDim.x] or null
    double* min_dist;            // [batch][n] squared distance to the chosen centre, or null
    const unsigned char* active; // [batch] problems to process (null = all)
};

struct KmSumsParams {
    const void* X;
    int dtype;
    int64_t n;
    int d, k;
    const double* mean;          // [batch][d] or null
    const int32_t* labels;       // [batch][n] or null (everything in cluster 0)
    int square;                  // accumulate (x-mean)^2 instead of (x-mean)
    double* partial;             // [batch][splits][k][d]
    long long* cnt_partial;      // [batch][splits][k]
    int splits;
    int dt;                      // dims per lane group (power of two <= 32)
    int kt;                      // clusters per shared-memory pass
    const unsigned char* active; // [batch] or null
};

int launch_kmeans_assign(KmAssignParams p, int batch, double* c2_ws, void* stream);
int launch_kmeans_sums(const KmSumsParams& p, int batch, double* sums, long long* counts, void* stream);
// fused E-step + M-step sums for uint8 rows whose [k][d] accumulators fit shared memory (one pass over X)
int kmeans_step_grid(int64_t n, int batch);
size_t kmeans_step_smem(int d, int k);
int launch_kmeans_step_u8(KmAssignParams p, int batch, double* partial, long long* cnt_partial, double* sums, long long* counts,
                          void* stream);
int launch_kmeans_centres(int batch, int d, int k, const double* sums, const long long* counts, const double* mean_sub,
                          int use_reciprocal, double* centres, double* shift_tot, double* shift_ws,
                          const unsigned char* active, void* stream);
int launch_kmeans_relocate(const void* X, int dtype, int batch, int64_t n, int d, int k, const double* mean,
                           const int32_t* labels, const double* centres_old, double* sums, long long* counts,
                           int raw_sums, const unsigned char* active, void* stream);
int launch_kmeans_far_points(const void* X, int dtype, int batch, int64_t n, int d, int k, const double* mean, const int32_t* labels,
                             const double* centres_old, int n_far, double* out_val, long long* out_idx, void* stream);
int launch_inertia_reduce(const double* partial, int parts, int batch, double* inertia, const unsigned char* active,
                          void* stream);
int launch_kmeans_cells(const unsigned char* X, int batch, int64_t n, int d, int k, const double* init,
                        unsigned long long seed, int max_iter, double tol, int32_t* labels, double* centres,
                        double* inertia, int32_t* n_iter, long long* counts, double* scratch, void* stream);
int kmeans_assign_grid(int64_t n);
int kmeans_sums_splits(int64_t n, int batch);

int launch_sliding_cosine(const double* a, int n, const double* b, int64_t m, double* sims, double* best,
                          long long* best_idx, void* stream);
int launch_row_cosine(const void* X, int dtype, int64_t n, int d, const double* q, double* out, void* stream);
int launch_vector_distance(const double* a, const double* b, int64_t n, double* cos_out, double* quirk_row,
                           double* l1, void* stream);
int launch_extract_cells(const unsigned char* bgr, int n_frames, int H, int W, int rows, int cols, int draw_lines,
                         int threshold, int swap_rb, unsigned char* out, void* stream);

}  // namespace ofc
