// =============================================================================
// tests/emu/cuda_emu.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A small host emulation of the CUDA execution model, just wide enough to compile the product's
// kernel sources (graphvite_b200/csrc/*.cu) with g++ and run them on a CPU: one fiber per CUDA
// thread, CTAs one after another, warp collectives (shfl / ballot / syncwarp) and CTA barriers
// (__syncthreads, bar.sync id) as cooperative rendezvous points, a malloc-backed CUDA runtime.
// It exists so that kernels written while no GPU is available can be EXECUTED against the oracle
// (tests/test_emulated_kernels.py) instead of only desk-checked.  It says nothing about speed,
// alignment faults that only the hardware raises, or inter-CTA races (CTAs run sequentially).
//
// Force-included (-include) by tests/emu/Makefile before every kernel source; GV_EMULATE selects
// the matching branch of graphvite_b200/csrc/gv_device.cuh.  The product never loads this build.
// =============================================================================
#pragma once
#define GV_EMULATE 1

#include <cuda_runtime.h>  // host-side types (float4, dim3, cudaStream_t ...) and API declarations

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>

// ---- qualifiers ------------------------------------------------------------------------------
#undef __shared__
#define __shared__ static  // CTAs run one at a time: one static instance per kernel is one CTA's copy
#undef __launch_bounds__
#define __launch_bounds__(...)
#undef __forceinline__
#define __forceinline__ inline
#undef __global__
#define __global__
#undef __device__
#define __device__
#undef __host__
#define __host__

namespace gv_emu {

// ---- scheduler interface (cuda_emu.cpp) ------------------------------------------------------
void run(dim3 grid, dim3 block, size_t shared_bytes, const std::function<void()> &thread_body);
// Two-pass sources (code that switches on __CUDA_ARCH__, i.e. the reference): the kernel a launch names is the
// HOST-pass instantiation; its device-pass twin lives in the library GV_EMU_DEVICE_LIBRARY, compiled from the same
// source with __CUDA_ARCH__ defined and its namespace renamed (GV_EMU_DEVICE_NAMESPACE="from=to").  Returns the
// twin's address (looked up by mangled name), or `kernel` itself when no device library is configured.
void *resolve_kernel(void *kernel);
void *dynamic_shared();
// all lanes in `mask` deposit a 64-bit value; returns the 32 deposited values once everyone arrived
const uint64_t *warp_gather(unsigned mask, uint64_t value);
void block_barrier(int id, int threads);  // threads <= 0: every live thread of the CTA
int lane_id();

template<class T>
inline uint64_t pack(T value) {
    static_assert(sizeof(T) <= 8, "warp collectives move at most 8 bytes");
    uint64_t bits = 0;
    memcpy(&bits, &value, sizeof(T));
    return bits;
}
template<class T>
inline T unpack(uint64_t bits) {
    T value;
    memcpy(&value, &bits, sizeof(T));
    return value;
}

template<class... P>
struct Launcher {
    dim3 grid, block;
    size_t shared;
    void (*kernel)(P...);
    template<class... A>
    void operator()(A &&...args) const {
        void (*k)(P...) = reinterpret_cast<void (*)(P...)>(resolve_kernel(reinterpret_cast<void *>(kernel)));
        // parameters are converted once, by value, like a kernel launch does
        auto bound = [k](P... converted) {
            return std::function<void()>([=]() { k(converted...); });
        }(std::forward<A>(args)...);
        run(grid, block, shared, bound);
    }
};

template<class... P>
inline Launcher<P...> launcher(dim3 grid, dim3 block, size_t shared, cudaStream_t, void (*kernel)(P...)) {
    return Launcher<P...>{grid, block, shared, kernel};
}

// kernel<<<grid, block[, shared[, stream]]>>>(args) of sources that cannot use GV_LAUNCH (the reference, see
// oracle/emulate_reference.py) is rewritten to gv_emu::LaunchConfig(grid, block[, shared[, stream]])(kernel)(args)
struct LaunchConfig {
    dim3 grid, block;
    size_t shared;
    LaunchConfig(dim3 _grid, dim3 _block, size_t _shared = 0, cudaStream_t = nullptr)
        : grid(_grid), block(_block), shared(_shared) {}
    template<class... P>
    Launcher<P...> operator()(void (*kernel)(P...)) const {
        return Launcher<P...>{grid, block, shared, kernel};
    }
};

}  // namespace gv_emu

// ---- built-in variables ------------------------------------------------------------------------
// (const like <curand_mtgp32_kernel.h> declares them for host code; defined writable in cuda_emu_vars.cpp)
extern const uint3 threadIdx, blockIdx;
extern const dim3 blockDim, gridDim;
static const int warpSize = 32;

// the kernel-pointer overloads <cuda_runtime.h> only defines for nvcc
template<class T>
inline cudaError_t cudaFuncSetAttribute(T *entry, enum cudaFuncAttribute attribute, int value) {
    return ::cudaFuncSetAttribute(reinterpret_cast<const void *>(entry), attribute, value);
}
template<class T>
inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *blocks, T *entry, int threads, size_t shared) {
    return ::cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks, reinterpret_cast<const void *>(entry), threads, shared);
}

// ---- what gv_device.cuh provides in the nvcc build ------------------------------------------------
#define GV_LAUNCH(grid, block, shared, stream, ...) gv_emu::launcher(grid, block, shared, stream, __VA_ARGS__)
#define GV_DYNAMIC_SHARED(type, name) type *name = static_cast<type *>(gv_emu::dynamic_shared())

inline void gv_named_barrier(int id, int threads) {
    gv_emu::block_barrier(id, threads);
}
unsigned long long gv_global_timer_ns();
inline float gv_fast_exp(float x) {
    return expf(x);
}
inline float gv_load_again(const float *address) { return *address; }
inline void gv_prefetch_l2(const void *) {}
inline void gv_prefetch_row_line(const void *address) {  // checked: must lie inside a live allocation
    (void)*static_cast<const volatile unsigned char *>(address);
}
inline void gv_wait_for(float &) {}  // a scheduling fence on the GPU; nothing to wait for here
inline float gv_fast_divide(float a, float b) {
    return a / b;
}
inline float gv_fast_rcp(float x) {
    return 1.f / x;
}
inline float gv_fast_sqrt(float x) {
    return sqrtf(x);
}
inline float gv_fast_rsqrt(float x) {
    return 1.f / sqrtf(x);
}

// ---- synchronisation ------------------------------------------------------------------------------
inline void __syncthreads() {
    gv_emu::block_barrier(0, 0);
}
inline void __syncwarp(unsigned mask = 0xFFFFFFFFu) {
    gv_emu::warp_gather(mask, 0);
}
inline void __threadfence() {}
inline void __threadfence_block() {}
inline void __threadfence_system() {
    __sync_synchronize();  // peers are other processes sharing memory
}
void __nanosleep(unsigned ns);

// ---- warp collectives -------------------------------------------------------------------------------
template<class T>
inline T __shfl_sync(unsigned mask, T value, int source, int width = 32) {
    const uint64_t *all = gv_emu::warp_gather(mask, gv_emu::pack(value));
    const int lane = gv_emu::lane_id();
    return gv_emu::unpack<T>(all[(lane & ~(width - 1)) | (source & (width - 1))]);
}
template<class T>
inline T __shfl_xor_sync(unsigned mask, T value, int lane_mask, int width = 32) {
    const uint64_t *all = gv_emu::warp_gather(mask, gv_emu::pack(value));
    const int lane = gv_emu::lane_id(), source = lane ^ lane_mask;
    return gv_emu::unpack<T>(all[(source & ~(width - 1)) == (lane & ~(width - 1)) ? source : lane]);
}
template<class T>
inline T __shfl_up_sync(unsigned mask, T value, unsigned delta, int width = 32) {
    const uint64_t *all = gv_emu::warp_gather(mask, gv_emu::pack(value));
    const int lane = gv_emu::lane_id(), source = lane - int(delta);
    return gv_emu::unpack<T>(all[source >= (lane & ~(width - 1)) ? source : lane]);
}
template<class T>
inline T __shfl_down_sync(unsigned mask, T value, unsigned delta, int width = 32) {
    const uint64_t *all = gv_emu::warp_gather(mask, gv_emu::pack(value));
    const int lane = gv_emu::lane_id(), source = lane + int(delta);
    return gv_emu::unpack<T>(all[source <= (lane | (width - 1)) ? source : lane]);
}
inline unsigned __ballot_sync(unsigned mask, int predicate) {
    const uint64_t *all = gv_emu::warp_gather(mask, predicate ? 1 : 0);
    unsigned result = 0;
    for (int l = 0; l < 32; l++)
        if ((mask >> l & 1) && all[l])
            result |= 1u << l;
    return result;
}
inline int __any_sync(unsigned mask, int predicate) {
    return __ballot_sync(mask, predicate) != 0;
}
inline int __all_sync(unsigned mask, int predicate) {
    return __ballot_sync(mask, !predicate) == 0;
}
unsigned __activemask();

// ---- integer intrinsics -------------------------------------------------------------------------------
inline int __ffs(int x) {
    return __builtin_ffs(x);
}
inline int __popc(unsigned x) {
    return __builtin_popcount(x);
}
inline int __popcll(unsigned long long x) {
    return __builtin_popcountll(x);
}
inline int __clz(int x) {
    return x ? __builtin_clz(unsigned(x)) : 32;
}
inline float __uint_as_float(unsigned x) {
    return gv_emu::unpack<float>(x);
}
inline unsigned __float_as_uint(float x) {
    return unsigned(gv_emu::pack(x));
}
inline float __int_as_float(int x) {
    return gv_emu::unpack<float>(uint64_t(unsigned(x)));
}
inline int __float_as_int(float x) {
    return int(gv_emu::pack(x));
}

// CUDA's global min / max overloads
#define GV_EMU_MINMAX(T)            \
    inline T min(T a, T b) {        \
        return a < b ? a : b;       \
    }                               \
    inline T max(T a, T b) {        \
        return a > b ? a : b;       \
    }
GV_EMU_MINMAX(int)
GV_EMU_MINMAX(unsigned)
GV_EMU_MINMAX(long long)
GV_EMU_MINMAX(unsigned long long)
GV_EMU_MINMAX(long)
GV_EMU_MINMAX(unsigned long)
GV_EMU_MINMAX(float)
GV_EMU_MINMAX(double)
#undef GV_EMU_MINMAX

// ---- memory: cache-hinted accesses are plain accesses; vector accesses must be naturally aligned ------
namespace gv_emu {
void misaligned(const void *pointer, size_t alignment);
template<class T>
inline void check_aligned(const T *pointer) {
    if (reinterpret_cast<uintptr_t>(pointer) % alignof(T) != 0 || reinterpret_cast<uintptr_t>(pointer) % sizeof(T) != 0)
        misaligned(pointer, sizeof(T));
}
}  // namespace gv_emu
#define GV_EMU_LOAD(name)                 \
    template<class T>                     \
    inline T name(const T *pointer) {     \
        gv_emu::check_aligned(pointer);   \
        return *pointer;                  \
    }
GV_EMU_LOAD(__ldg)
GV_EMU_LOAD(__ldcg)
GV_EMU_LOAD(__ldca)
GV_EMU_LOAD(__ldcs)
GV_EMU_LOAD(__ldlu)
GV_EMU_LOAD(__ldcv)
#undef GV_EMU_LOAD
#define GV_EMU_STORE(name)                       \
    template<class T>                            \
    inline void name(T *pointer, T value) {      \
        gv_emu::check_aligned(pointer);          \
        *pointer = value;                        \
    }
GV_EMU_STORE(__stcg)
GV_EMU_STORE(__stcs)
GV_EMU_STORE(__stwb)
GV_EMU_STORE(__stwt)
#undef GV_EMU_STORE

// ---- atomics: fibers are cooperative, a read-modify-write is atomic by construction ----------------------
template<class T>
inline T atomicAdd(T *address, T value) {
    const T old = *address;
    *address = old + value;
    return old;
}
template<class T>
inline T atomicMax(T *address, T value) {
    const T old = *address;
    *address = old > value ? old : value;
    return old;
}
template<class T>
inline T atomicMin(T *address, T value) {
    const T old = *address;
    *address = old < value ? old : value;
    return old;
}
template<class T>
inline T atomicExch(T *address, T value) {
    const T old = *address;
    *address = value;
    return old;
}
template<class T>
inline T atomicCAS(T *address, T compare, T value) {
    const T old = *address;
    if (old == compare)
        *address = value;
    return old;
}
template<class T>
inline T atomicOr(T *address, T value) {
    const T old = *address;
    *address = old | value;
    return old;
}
