// One-shot exchange between the GPUs of one node over NVLink peer memory: the collective of the sharded k-means
// (the all-reduce of the per-cluster sums / counts after the M-step, and the all-gather of the far-point lists of the
// empty-cluster relocation) as ONE kernel per rank instead of a library call per buffer.
//
// What it replaces: the exchange between workers that `KMeans.fit` would need if the rows of KmeanGrids.py:299-304 /
// color_kmeans.py:65-78 were split over GPUs; sizes are 264 bytes (k = 8, 4 features) to 1 MB (k = 1024, 128 features)
// per iteration, i.e. pure latency -- a pull over NVLink with two flags is shorter than a ring.
//
// Protocol (every rank runs the same kernel on its own stream; `bufs[r]` is rank r's buffer mapped into this process):
//   buffer = PeerHeader | slot 0 | slot 1 | gather region.   Exchange number e = header.seq + 1 (kept on the device, so
//   the kernel can sit in a CUDA graph).
//   1. the data of exchange e was written into the region by earlier kernels of this stream;
//   2. CTA 0: fence.sys, then flag[rank] = e in EVERY peer's header (st.release.sys);
//   3. every CTA waits until its own header shows flag[src] >= e for every src (ld.acquire.sys);
//   4. pull: out[i] = sum over src IN RANK ORDER of bufs[src].region[i]  -- the same order on every rank, so all ranks
//      hold bit-identical results (and for uint8 rows the sums are exact integers, so they equal the 1-GPU sums);
//      mode "gather": out[src][i] = bufs[src].region[i];
//   5. the last CTA to finish stores header.seq = e.
// A region is rewritten only two exchanges later (slots alternate per Lloyd iteration; a gather is always separated
// from the next one by an all-reduce): a rank that has completed exchange e + 1 has seen every peer's flag e + 1, which
// a peer sets only after it finished reading exchange e.  The wait has a time limit (the error word is set and the fit
// raises) so that a lost rank cannot hang the GPU.
#include <string.h>

#include "ofc_common.cuh"

namespace ofc {

constexpr int kPeerMaxWorld = 16;
constexpr size_t kPeerHeaderBytes = 256;

struct PeerHeader {
    unsigned int flag[kPeerMaxWorld];   // flag[src]: last exchange rank src has published (written by src)
    unsigned int seq;                   // last exchange this rank completed (local)
    unsigned int done;                  // CTAs of the running exchange that have finished (local)
    int err;                            // 1: a wait timed out
    int pad;
};
static_assert(sizeof(PeerHeader) <= kPeerHeaderBytes, "header");

#ifndef OFC_EMULATE
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_peer_f64(const double* p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ long long ld_peer_i64(const long long* p) {
    long long v;
    asm volatile("ld.relaxed.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#else
static inline void st_release_sys(unsigned int* p, unsigned int v) { *p = v; }
static inline unsigned int ld_acquire_sys(const unsigned int* p) { return *(const volatile unsigned int*)p; }
static inline double ld_peer_f64(const double* p) { return *p; }
static inline long long ld_peer_i64(const long long* p) { return *p; }
static inline unsigned long long timer_ns() { static unsigned long long t = 0; return t += 1000000ull; }
static inline void __threadfence_system() {}
#endif

// mode 0: sum n_f64 doubles then n_i64 int64 of the region at `region_off` (bytes from the buffer start) over all ranks
// mode 1: gather n_f64 doubles of every rank: out_f64[src * n_f64 + i]
// gate (optional): int64[gate_n]; the exchange only happens when one of them is zero (all ranks hold the same values)
__global__ void __launch_bounds__(256) peer_exchange_kernel(void* const* __restrict__ bufs, int world, int rank, size_t region_off,
                                                            int mode, long long n_f64, long long n_i64, double* __restrict__ out_f64,
                                                            long long* __restrict__ out_i64, const long long* __restrict__ gate,
                                                            int gate_n, unsigned long long timeout_ns) {
    __shared__ unsigned int s_seq;
    __shared__ int s_gate;
    const int tid = threadIdx.x;
    PeerHeader* me = reinterpret_cast<PeerHeader*>(bufs[rank]);
    if (gate) {
        if (tid == 0) s_gate = 0;
        __syncthreads();
        int any = 0;
        for (int i = tid; i < gate_n; i += blockDim.x) any |= gate[i] == 0;
        if (any) s_gate = 1;
        __syncthreads();
        if (!s_gate) return;                       // same decision on every rank and in every CTA: nothing is published
    }
    if (tid == 0) s_seq = *reinterpret_cast<volatile unsigned int*>(&me->seq) + 1u;
    __syncthreads();
    const unsigned int seq = s_seq;
    if (blockIdx.x == 0 && tid < world) {
        __threadfence_system();                    // the region was written by earlier kernels of this stream
        st_release_sys(&reinterpret_cast<PeerHeader*>(bufs[tid])->flag[rank], seq);
    }
    if (tid < world) {
        const unsigned long long t0 = timer_ns();
        while ((int)(ld_acquire_sys(&me->flag[tid]) - seq) < 0) {
            if (timer_ns() - t0 > timeout_ns) { me->err = 1; break; }
        }
    }
    __syncthreads();

    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long t = (long long)blockIdx.x * blockDim.x + tid;
    if (mode == 0) {
        for (long long i = t; i < n_f64; i += stride) {
            double acc = 0.0;
            for (int src = 0; src < world; ++src)
                acc += ld_peer_f64(reinterpret_cast<const double*>(static_cast<const char*>(bufs[src]) + region_off) + i);
            out_f64[i] = acc;
        }
        for (long long i = t; i < n_i64; i += stride) {
            long long acc = 0;
            for (int src = 0; src < world; ++src)
                acc += ld_peer_i64(reinterpret_cast<const long long*>(static_cast<const char*>(bufs[src]) + region_off) + n_f64 + i);
            out_i64[i] = acc;
        }
    } else {
        for (int src = 0; src < world; ++src) {
            const double* from = reinterpret_cast<const double*>(static_cast<const char*>(bufs[src]) + region_off);
            for (long long i = t; i < n_f64; i += stride) out_f64[(long long)src * n_f64 + i] = ld_peer_f64(from + i);
        }
    }

    __syncthreads();
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(&me->done, 1u) == gridDim.x - 1) {
            me->done = 0;
            *reinterpret_cast<volatile unsigned int*>(&me->seq) = seq;
            __threadfence();
        }
    }
}

}  // namespace ofc

using namespace ofc;

extern "C" {

size_t ofc_peer_header_bytes(void) { return kPeerHeaderBytes; }

int ofc_peer_alloc(size_t bytes, void** ptr, unsigned char* handle64) {
    OFC_REQUIRE(ptr && bytes >= kPeerHeaderBytes, "ofc_peer_alloc: need at least the %zu-byte header", kPeerHeaderBytes);
    void* p = nullptr;
    OFC_CUDA(cudaMalloc(&p, bytes));
    OFC_CUDA(cudaMemset(p, 0, bytes));
    OFC_CUDA(cudaDeviceSynchronize());
#ifndef OFC_EMULATE
    if (handle64) {
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        cudaIpcMemHandle_t h;
        cudaError_t e = cudaIpcGetMemHandle(&h, p);
        if (e != cudaSuccess) {
            cudaFree(p);
            return check_cuda(e, "cudaIpcGetMemHandle");
        }
        memcpy(handle64, &h, 64);
    }
#else
    if (handle64) memset(handle64, 0, 64);
#endif
    *ptr = p;
    return OFC_OK;
}

int ofc_peer_open(const unsigned char* handle64, void** ptr) {
    OFC_REQUIRE(handle64 && ptr, "ofc_peer_open: null argument");
#ifndef OFC_EMULATE
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    OFC_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return OFC_OK;
#else
    set_error("ofc_peer_open: no inter-process mapping in the host emulation");
    return OFC_ERR_UNSUPPORTED;
#endif
}

int ofc_peer_close(void* ptr) {
#ifndef OFC_EMULATE
    if (ptr) OFC_CUDA(cudaIpcCloseMemHandle(ptr));
#endif
    return OFC_OK;
}

int ofc_peer_free(void* ptr) {
    if (ptr) OFC_CUDA(cudaFree(ptr));
    return OFC_OK;
}

int ofc_peer_exchange(const void* bufs, int world, int rank, size_t region_offset, int mode, int64_t n_f64, int64_t n_i64,
                      double* out_f64, int64_t* out_i64, const int64_t* gate, int gate_n, double timeout_s, void* stream) {
    OFC_REQUIRE(bufs && world >= 1 && world <= kPeerMaxWorld && rank >= 0 && rank < world, "ofc_peer_exchange: world %d rank %d", world, rank);
    OFC_REQUIRE(mode == 0 || mode == 1, "ofc_peer_exchange: mode %d", mode);
    OFC_REQUIRE(n_f64 >= 0 && n_i64 >= 0 && (n_f64 == 0 || out_f64) && (n_i64 == 0 || (out_i64 && mode == 0)), "ofc_peer_exchange: outputs");
    OFC_REQUIRE(region_offset >= kPeerHeaderBytes && region_offset % 16 == 0, "ofc_peer_exchange: region offset %zu", region_offset);
    const long long work = (long long)(n_f64 + n_i64) * (mode == 1 ? world : 1);
    int grid = (int)((work + 256 * 4 - 1) / (256 * 4));
    grid = grid < 1 ? 1 : (grid > 64 ? 64 : grid);          // every CTA must be resident while it waits: far below 148 SMs
    const unsigned long long tmo = (unsigned long long)((timeout_s > 0 ? timeout_s : 10.0) * 1e9);
    OFC_LAUNCH(peer_exchange_kernel, grid, 256, 0, stream, (void* const*)bufs, world, rank, region_offset, mode, (long long)n_f64,
               (long long)n_i64, out_f64, (long long*)out_i64, (const long long*)gate, gate_n, tmo);
    OFC_CHECK_LAUNCH("peer_exchange");
    return OFC_OK;
}

int ofc_peer_error(const void* own_buffer, int* err) {
    OFC_REQUIRE(own_buffer && err, "ofc_peer_error: null argument");
    PeerHeader h;
    OFC_CUDA(cudaMemcpy(&h, own_buffer, sizeof(h), cudaMemcpyDeviceToHost));
    *err = h.err;
    return OFC_OK;
}

}  // extern "C"
