// Common definitions for the sm_100a kernels of libofc.
//
// Two build modes:
//   * nvcc -gencode arch=compute_100a,code=sm_100a  -> the product (libofc.so)
//   * g++ -DOFC_EMULATE -include tests/emu/cuda_emu.h -> a host-side *debug
//     emulation* of the same kernel sources (fibers standing in for CUDA
//     threads).  It exists only so kernel indexing can be debugged in a
//     container without a GPU; it is built and loaded by tests/ only and is
//     never reachable from the Python package (see tests/emu/README.md).
#pragma once

#include <stdint.h>
#include <stddef.h>

#ifndef OFC_EMULATE
#include <cuda_runtime.h>
#define OFC_LAUNCH(kernel, grid, block, smem, stream, ...) \
    kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__)
#define OFC_DYN_SMEM(type, name) extern __shared__ __align__(16) unsigned char name##_raw_[]; \
    type* name = reinterpret_cast<type*>(name##_raw_)
#else
#define OFC_LAUNCH(kernel, grid, block, smem, stream, ...) \
    ofc_emu::launch((grid), (block), (smem), [=]() { kernel(__VA_ARGS__); })
#define OFC_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(ofc_emu::dyn_smem())
#endif

#define OFC_OK 0
#define OFC_ERR_INVALID (-1)
#define OFC_ERR_UNSUPPORTED (-2)
#define OFC_ERR_CUDA (-3)
#define OFC_ERR_WORKSPACE (-4)

namespace ofc {

void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);

#define OFC_CUDA(call)                                              \
    do {                                                            \
        int ofc_rc_ = ::ofc::check_cuda((call), #call);             \
        if (ofc_rc_ != OFC_OK) return ofc_rc_;                      \
    } while (0)

#define OFC_CHECK_LAUNCH(name) OFC_CUDA(cudaGetLastError())

// Opt a kernel in to more than 48 KB of dynamic shared memory.  The attribute is per device, so the
// "already done" cache is keyed by (current device, kernel) and guarded by a mutex: one process may
// drive several GPUs and several host threads may launch (ofc_api.cu).
int smem_optin(const void* kernel, size_t bytes);
#define OFC_SMEM_OPTIN(kernel, bytes)                                                   \
    do {                                                                                \
        if ((size_t)(bytes) > 48 * 1024) {                                              \
            int ofc_rc_ = ::ofc::smem_optin(reinterpret_cast<const void*>(kernel), (size_t)(bytes)); \
            if (ofc_rc_ != OFC_OK) return ofc_rc_;                                      \
        }                                                                               \
    } while (0)

#define OFC_REQUIRE(cond, ...)                                      \
    do {                                                            \
        if (!(cond)) {                                              \
            ::ofc::set_error(__VA_ARGS__);                          \
            return OFC_ERR_INVALID;                                 \
        }                                                           \
    } while (0)

// Optional per-kernel timing (ofc_profile_begin / ofc_profile_end in ofc.h): when
// enabled every launcher brackets its kernel with two cudaEvents on the launch
// stream.  Off by default; costs nothing then.
enum ProfKind {
    PK_GRAY = 0, PK_PREFILTER, PK_POLYEXP, PK_MINMAX_INIT, PK_ENCODE, PK_GRID, PK_FLOW_MINMAX, PK_DRAW,
    PK_KMEANS, PK_COSINE, PK_UPSAMPLE, PK_RESERVED1,
    PK_ITER_L0 = 12,             // flow_iter at full resolution; +1 per coarser level (up to 8)
    PK_COUNT = 20
};
struct ProfScope {
    int kind; void* stream; bool on; int index;
    ProfScope(int kind, void* stream);
    ~ProfScope();
};
extern int g_prof_level;         // pyramid level (0 = finest) of the flow_iter launch being issued

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

__host__ __device__ __forceinline__ int clampi(int v, int lo, int hi) {
    return v < lo ? lo : (v > hi ? hi : v);
}

// BORDER_REFLECT_101 for |overshoot| < n (single reflection is enough on this
// path: the Gaussian radius is always smaller than the image).
__host__ __device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    const int p = 2 * (n - 1);
    i %= p;
    if (i < 0) i += p;
    return i >= n ? p - i : i;
}

// same map for -n < i < 2n-1 without the modulo (n >= 2)
__host__ __device__ __forceinline__ int reflect101_near(int i, int n) {
    return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i);
}

// byte T of a word as a float without an integer->float conversion: 0x4B0000bb is 2^23 + bb
template <int T> __device__ __forceinline__ float byte_to_float(unsigned w) {
#ifndef OFC_EMULATE
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540 + T)) - 8388608.f;
#else
    return (float)((w >> (8 * T)) & 255u);
#endif
}

}  // namespace ofc
