// tests/hostsim/cuda_runtime.h -- TEST INFRASTRUCTURE, never part of the product.
//
// A stand-in for <cuda_runtime.h> that lets g++ compile die_b200/csrc/*.cu{,h} UNCHANGED (after
// tests/hostsim/build.py rewrote the three pieces of syntax g++ cannot parse: `k<<<g, b, s, st>>>(...)`,
// `extern __shared__ T x[];` and the inline PTX prefetch) into a host library that executes every kernel
// thread by thread on the CPU: one cooperative fiber per CUDA thread, __syncthreads / warp shuffles /
// ballots as fiber barriers (tests/hostsim/hostsim.cpp).  "Device memory" is host memory, streams and
// events are no-ops.
//
// Purpose: `-m "not gpu"` tests run the REAL kernel source through the C ABI against the oracle, so a logic
// error in a kernel shows up without a GPU.  It says nothing about speed or about races between CTAs (blocks
// run one after the other).  die_b200 itself never loads this library: the product path is CUDA only.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>

// ---- qualifiers --------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __shared__ static      /* blocks run one at a time, so one static copy per kernel is "per block" */
#define __constant__

// ---- vector types -----------------------------------------------------------------------------
struct uint2 { unsigned x, y; };
struct uint3 { unsigned x, y, z; };
struct uint4 { unsigned x, y, z, w; };
struct alignas(16) double2 { double x, y; };
struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) int2 { int x, y; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
static inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
static inline int2 make_int2(int x, int y) { int2 r; r.x = x; r.y = y; return r; }

// ---- the running thread ------------------------------------------------------------------------
namespace hostsim {
struct ThreadCtx {
    uint3 tid, bid;
    dim3 bdim, gdim;
};
extern ThreadCtx* cur;
void* dyn_smem();
void syncthreads();
int syncthreads_or(int pred);
void warp_exchange(uint64_t mine, uint64_t* all32);      // all32[l] = value deposited by lane l (stale if it exited)
unsigned lane_id();
void run_grid(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
void run_grid(dim3 grid, dim3 block, unsigned cluster, size_t smem, const std::function<void()>& body);   // thread-block clusters
unsigned cluster_rank();
void cluster_sync();
void* cluster_map(void* p, unsigned rank);
void* dev_alloc(size_t bytes);
void dev_free(void* p);
extern int last_error;
}  // namespace hostsim

#define threadIdx (hostsim::cur->tid)
#define blockIdx (hostsim::cur->bid)
#define blockDim (hostsim::cur->bdim)
#define gridDim (hostsim::cur->gdim)

// ---- arithmetic intrinsics (compile with -ffp-contract=off: no contraction, as nvcc -fmad=false) ---
using std::max;
using std::min;
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
static inline int __double2loint(double d) { uint64_t u; memcpy(&u, &d, 8); return (int)(uint32_t)u; }
static inline int __double2hiint(double d) { uint64_t u; memcpy(&u, &d, 8); return (int)(uint32_t)(u >> 32); }
static inline double __hiloint2double(int hi, int lo) {
    const uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double d; memcpy(&d, &u, 8); return d;
}
static inline long long __double_as_longlong(double d) { long long u; memcpy(&u, &d, 8); return u; }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline T __ldcg(const T* p) { return *p; }
static inline int atomicMax(int* p, int v) { const int old = *p; if (v > old) *p = v; return old; }
static inline unsigned long long atomicCAS(unsigned long long* p, unsigned long long cmp, unsigned long long val) {
    const unsigned long long old = *p;          // (fibers switch only at barriers: a plain read-modify-write is atomic here)
    if (old == cmp) *p = val;
    return old;
}

// ---- warp collectives -------------------------------------------------------------------------
template <typename T>
static inline T hostsim_shfl_(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle of a type wider than 8 bytes");
    uint64_t mine = 0, all[32];
    memcpy(&mine, &v, sizeof(T));
    hostsim::warp_exchange(mine, all);
    T r;
    memcpy(&r, &all[src_lane & 31], sizeof(T));
    return r;
}
template <typename T> static inline T __shfl_sync(unsigned, T v, int src) { return hostsim_shfl_(v, src); }
template <typename T> static inline T __shfl_down_sync(unsigned, T v, unsigned delta) {
    const unsigned l = hostsim::lane_id();
    const T r = hostsim_shfl_(v, (int)(l + delta));                  // everybody takes part in the exchange
    return (l + delta < 32) ? r : v;
}
static inline unsigned __ballot_sync(unsigned, int pred) {
    uint64_t all[32];
    hostsim::warp_exchange(pred ? 1 : 0, all);
    unsigned w = 0;
    for (int l = 0; l < 32; ++l) w |= (unsigned)(all[l] & 1u) << l;
    return w;
}
static inline void __syncthreads() { hostsim::syncthreads(); }
static inline int __syncthreads_or(int pred) { return hostsim::syncthreads_or(pred); }
static inline void __syncwarp(unsigned = 0xffffffffu) { uint64_t all[32]; hostsim::warp_exchange(0, all); }
static inline void __trap() { abort(); }

// ---- die_async.cuh under the emulator: mbarrier + 1-D bulk copies (see hostsim.cpp) ----------------
namespace hostsim {
void mbar_init(void* bar, int count);
void mbar_arrive_expect_tx(void* bar, uint32_t bytes);
void bulk_g2s(void* dst, const void* src, uint32_t bytes, void* bar);
void mbar_wait(void* bar, uint32_t parity);
}  // namespace hostsim
#if defined(DIE_HOSTSIM)
namespace die {
typedef unsigned long long mbar_t;
static inline void mbar_init(mbar_t* bar, int count) { hostsim::mbar_init(bar, count); }
static inline void mbar_fence_init() {}
static inline void mbar_arrive_expect_tx(mbar_t* bar, uint32_t bytes) { hostsim::mbar_arrive_expect_tx(bar, bytes); }
static inline void bulk_g2s(void* dst, const void* src, uint32_t bytes, mbar_t* bar) { hostsim::bulk_g2s(dst, src, bytes, bar); }
static inline void mbar_wait(mbar_t* bar, uint32_t parity) { hostsim::mbar_wait(bar, parity); }
static inline void proxy_fence_async() {}
// die_cluster.cuh under the emulator: the CTAs of a cluster run concurrently, each with its own shared memory
static inline unsigned cluster_rank() { return hostsim::cluster_rank(); }
static inline void cluster_sync() { hostsim::cluster_sync(); }
template <typename T> static inline T* cluster_map(T* p, unsigned rank) { return (T*)hostsim::cluster_map((void*)p, rank); }
}  // namespace die
#endif

// ---- runtime API -------------------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorLaunchFailure = 719 };
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum { cudaDevAttrMultiProcessorCount = 16 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };

static inline const char* cudaGetErrorString(cudaError_t e) {
    switch (e) {
        case cudaSuccess: return "no error";
        case cudaErrorInvalidValue: return "invalid argument";
        case cudaErrorMemoryAllocation: return "out of memory";
        default: return "hostsim: kernel deadlock (a barrier some thread never reached)";
    }
}
template <typename T> static inline cudaError_t cudaMalloc(T** p, size_t n) {
    *p = (T*)hostsim::dev_alloc(n);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
static inline cudaError_t cudaFree(void* p) { hostsim::dev_free(p); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* p, int v, size_t n) { memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) {
    memmove(d, s, n);
    return cudaSuccess;
}
static inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dpitch, const void* s, size_t spitch, size_t width, size_t height,
                                            cudaMemcpyKind, cudaStream_t) {
    for (size_t r = 0; r < height; ++r) memmove((char*)d + r * dpitch, (const char*)s + r * spitch, width);
    return cudaSuccess;
}
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) { *v = 2; return cudaSuccess; }   // "2 SMs"
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
enum cudaLimit { cudaLimitMaxL2FetchGranularity = 5 };
static inline cudaError_t cudaDeviceGetLimit(size_t* v, cudaLimit) { *v = 64; return cudaSuccess; }
static inline cudaError_t cudaDeviceSetLimit(cudaLimit, size_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (void*)1; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (void*)1; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (void*)1; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
static inline cudaError_t cudaGetLastError() {
    const int e = hostsim::last_error;
    hostsim::last_error = cudaSuccess;
    return e;
}

// ---- kernel launch (build.py rewrites `k<<<g, b, s, st>>>(args)` into `hostsim::launch(g, b, s, st, k, args)`) ---
namespace hostsim {
template <typename K, typename... A>
static inline void launch(dim3 grid, dim3 block, size_t smem, cudaStream_t, K kern, A&&... args) {
    run_grid(grid, block, smem, [&]() { kern(args...); });
}
}  // namespace hostsim
#if defined(DIE_HOSTSIM)
namespace die {
template <typename A>
static inline cudaError_t launch_cluster(void (*kern)(const A), unsigned grid, unsigned block, unsigned cluster_x,
                                         size_t smem, cudaStream_t, const A& arg) {
    hostsim::run_grid(dim3(grid), dim3(block), cluster_x, smem, [&]() { kern(arg); });
    return cudaGetLastError();
}
}  // namespace die
#endif
