// Host-side debug emulation of the small CUDA subset the libofc kernels use.
//
// TEST TOOLING ONLY.  `g++ -DOFC_EMULATE -include tests/emu/cuda_emu.h` turns
// the .cu sources into a host library in which every CUDA thread of a block is
// a ucontext fiber (so __syncthreads / warp shuffles behave), blocks run one
// after another, and "device" pointers are host pointers.  It lets kernel
// indexing be debugged in a container without a GPU.  The Python package never
// loads it; only tests/test_emulated_kernels.py does, explicitly by path.
#pragma once
#ifndef OFC_EMULATE
#error "cuda_emu.h is only for -DOFC_EMULATE builds"
#endif

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <type_traits>
#include <ucontext.h>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __restrict__ __restrict
#define __shared__ static
#define __constant__ static
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct double2 { double x, y; };
struct uchar2 { unsigned char x, y; };
struct uchar3 { unsigned char x, y, z; };
struct __attribute__((aligned(4))) uchar4 { unsigned char x, y, z, w; };
struct int2 { int x, y; };
struct __attribute__((aligned(16))) int4 { int x, y, z, w; };
struct __attribute__((aligned(8))) uint2 { unsigned x, y; };
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static inline uchar4 make_uchar4(unsigned char x, unsigned char y, unsigned char z, unsigned char w) { return uchar4{x, y, z, w}; }
static inline uchar3 make_uchar3(unsigned char x, unsigned char y, unsigned char z) { return uchar3{x, y, z}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }

typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = aligned_alloc(256, (n + 255) / 256 * 256); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* p, int v, size_t n) { memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
typedef int cudaEvent_t;
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = 0; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = 0; return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
struct cudaDeviceProp { int multiProcessorCount; };
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { p->multiProcessorCount = 3; return cudaSuccess; }
template <class F> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, F, int, size_t) { *n = 2; return cudaSuccess; }

namespace ofc_emu {

struct Fiber {
    ucontext_t ctx;
    bool done = false;
    uint3 tid;
};

struct State {
    std::vector<Fiber> fibers;
    std::vector<char*> stacks;
    ucontext_t sched;
    int cur = -1;
    int nthreads = 0;
    int alive = 0;
    int bar_count = 0;
    unsigned bar_gen = 0;
    // per-warp sync state
    int warp_count[32];
    unsigned warp_gen[32];
    int warp_alive[32];
    unsigned long long warp_buf[32][32];
    std::function<void()> body;
    std::vector<unsigned char> dyn;
};

inline State& st() { static State s; return s; }

}  // namespace ofc_emu

// CUDA built-in index variables (one OS thread, updated on every fiber switch)
inline uint3 threadIdx, blockIdx;
inline dim3 blockDim, gridDim;
static const int warpSize = 32;

namespace ofc_emu {

static const size_t kStack = 256 * 1024;

inline void* dyn_smem() { return st().dyn.data(); }

inline void yield_() {
    State& s = st();
    int me = s.cur;
    swapcontext(&s.fibers[me].ctx, &s.sched);
    threadIdx = s.fibers[me].tid;
}

inline void fiber_entry() {
    State& s = st();
    int me = s.cur;
    threadIdx = s.fibers[me].tid;
    s.body();
    s.fibers[me].done = true;
    s.alive--;
    int w = me / 32;
    s.warp_alive[w]--;
    // a thread that exits counts as arrived at pending barriers
    if (s.alive > 0 && s.bar_count >= s.alive) { s.bar_count = 0; s.bar_gen++; }
    if (s.warp_alive[w] > 0 && s.warp_count[w] >= s.warp_alive[w]) { s.warp_count[w] = 0; s.warp_gen[w]++; }
    swapcontext(&s.fibers[me].ctx, &s.sched);
}

inline void run_block(unsigned nthreads) {
    State& s = st();
    if (s.fibers.size() < nthreads) s.fibers.resize(nthreads);
    while (s.stacks.size() < nthreads) s.stacks.push_back((char*)malloc(kStack));
    s.nthreads = s.alive = (int)nthreads;
    s.bar_count = 0;
    for (int w = 0; w < 32; ++w) { s.warp_count[w] = 0; s.warp_alive[w] = 0; }
    for (unsigned i = 0; i < nthreads; ++i) {
        Fiber& f = s.fibers[i];
        f.done = false;
        f.tid.x = i % blockDim.x;
        f.tid.y = (i / blockDim.x) % blockDim.y;
        f.tid.z = i / (blockDim.x * blockDim.y);
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = s.stacks[i];
        f.ctx.uc_stack.ss_size = kStack;
        f.ctx.uc_link = nullptr;
        makecontext(&f.ctx, (void (*)())fiber_entry, 0);
        s.warp_alive[i / 32]++;
    }
    while (s.alive > 0) {
        for (unsigned i = 0; i < nthreads; ++i) {
            if (s.fibers[i].done) continue;
            s.cur = (int)i;
            swapcontext(&s.sched, &s.fibers[i].ctx);
        }
    }
    s.cur = -1;
}

template <class F>
inline void launch(dim3 grid, dim3 block, size_t smem, F f) {
    State& s = st();
    s.body = f;
    if (s.dyn.size() < smem + 16) s.dyn.resize(smem + 16);
    gridDim = grid;
    blockDim = block;
    unsigned nthreads = block.x * block.y * block.z;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                blockIdx = uint3{bx, by, bz};
                run_block(nthreads);
            }
}

inline void sync_block() {
    State& s = st();
    unsigned gen = s.bar_gen;
    if (++s.bar_count >= s.alive) { s.bar_count = 0; s.bar_gen++; return; }
    while (s.bar_gen == gen) yield_();
}

inline void sync_warp() {
    State& s = st();
    int w = s.cur / 32;
    unsigned gen = s.warp_gen[w];
    if (++s.warp_count[w] >= s.warp_alive[w]) { s.warp_count[w] = 0; s.warp_gen[w]++; return; }
    while (s.warp_gen[w] == gen) yield_();
}

template <class T>
inline T warp_exchange(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    State& s = st();
    int w = s.cur / 32, lane = s.cur % 32;
    unsigned long long bits = 0;
    memcpy(&bits, &v, sizeof(T));
    s.warp_buf[w][lane] = bits;
    sync_warp();
    int nl = std::min(32, s.nthreads - w * 32);
    unsigned long long got = s.warp_buf[w][(src_lane >= 0 && src_lane < nl) ? src_lane : lane];
    sync_warp();
    T r;
    memcpy(&r, &got, sizeof(T));
    return r;
}

}  // namespace ofc_emu

static inline void __syncthreads() { ofc_emu::sync_block(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { ofc_emu::sync_warp(); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}

template <class T> static inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
    int lane = ofc_emu::st().cur % 32;
    int base = lane / width * width;
    return ofc_emu::warp_exchange(v, base + (src % width));
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32) {
    int lane = ofc_emu::st().cur % 32;
    (void)width;
    return ofc_emu::warp_exchange(v, lane ^ m);
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int width = 32) {
    int lane = ofc_emu::st().cur % 32;
    int src = lane + (int)d;
    if (src / width != lane / width) src = lane;
    return ofc_emu::warp_exchange(v, src);
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int width = 32) {
    int lane = ofc_emu::st().cur % 32;
    int src = lane - (int)d;
    if (src < 0 || src / width != lane / width) src = lane;
    return ofc_emu::warp_exchange(v, src);
}
static inline unsigned __ballot_sync(unsigned, int pred) {
    unsigned r = 0;
    for (int l = 0; l < 32; ++l) {
        int p = ofc_emu::warp_exchange(pred ? 1 : 0, l);
        // warp_exchange returns own value for out-of-range lanes: mask them
        int nl = std::min(32, ofc_emu::st().nthreads - (ofc_emu::st().cur / 32) * 32);
        if (l < nl && p) r |= 1u << l;
    }
    return r;
}

// atomics: fibers never preempt, so plain read-modify-write is atomic
template <class T> static inline T atomicAdd(T* a, T v) { T o = *a; *a = o + v; return o; }
template <class T> static inline T atomicMin(T* a, T v) { T o = *a; if (v < o) *a = v; return o; }
template <class T> static inline T atomicMax(T* a, T v) { T o = *a; if (v > o) *a = v; return o; }
template <class T> static inline T atomicExch(T* a, T v) { T o = *a; *a = v; return o; }
template <class T> static inline T atomicCAS(T* a, T c, T v) { T o = *a; if (o == c) *a = v; return o; }
template <class T> static inline T atomicOr(T* a, T v) { T o = *a; *a = o | v; return o; }

// math / conversion intrinsics
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
// round-toward-zero float add: the double sum of two floats is exact here (exponents < 2^29 apart)
static inline float __fadd_rz(float a, float b) {
    const double d = (double)a + (double)b;
    float r = (float)d;
    if (std::fabs((double)r) > std::fabs(d)) r = std::nextafterf(r, 0.0f);
    return r;
}
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return std::sqrt(a); }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline int __float2int_rz(float a) { return (int)a; }
static inline int __float2int_rn(float a) { return (int)nearbyintf(a); }
static inline int __float2int_rd(float a) { return (int)std::floor(a); }
static inline int __double2int_rd(double a) { return (int)std::floor(a); }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
static inline long long __double_as_longlong(double d) { long long u; memcpy(&u, &d, 8); return u; }
static inline double __longlong_as_double(long long u) { double d; memcpy(&d, &u, 8); return d; }
template <class T> static inline T __ldg(const T* p) { return *p; }
// funnel shift right: low 32 bits of ((hi:lo) >> (shift & 31))
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned shift) {
    shift &= 31u;
    return (unsigned)(((((unsigned long long)hi) << 32) | lo) >> shift);
}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
#include <math.h>
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline float min(float a, float b) { return a < b ? a : b; }
static inline float max(float a, float b) { return a > b ? a : b; }
static inline double min(double a, double b) { return a < b ? a : b; }
static inline double max(double a, double b) { return a > b ? a : b; }
template <class T> static inline T __ldcs(const T* p) { return *p; }
