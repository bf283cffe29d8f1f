// Device-side helpers shared by the CUDA translation units of libgv_b200: the kernel launch macro,
// dynamic shared memory, and the few inline-PTX / fast-math primitives the kernels use.
//
// The product is compiled by nvcc for sm_100a and that is the only configuration that ships.
// GV_EMULATE is defined only by tests/emu/Makefile, which compiles these very kernel sources for
// the host (g++, one fiber per CUDA thread, warp collectives and barriers emulated) so that the
// kernels' logic can be executed against the oracle on machines without a GPU.  That build lives
// under tests/ and is test infrastructure like oracle/: the package never loads it.
#pragma once

#if defined(GV_EMULATE)

// tests/emu/cuda_emu.h (force-included by the emulation build) provides GV_LAUNCH,
// GV_DYNAMIC_SHARED, gv_named_barrier, gv_global_timer_ns, gv_wait_for, gv_load_again, gv_prefetch_l2, gv_fast_exp and gv_fast_divide.

#else

// kernel<<<grid, block, shared, stream>>>(args...)  ==  GV_LAUNCH(grid, block, shared, stream, kernel)(args...)
#define GV_LAUNCH(grid, block, shared, stream, ...) __VA_ARGS__<<<grid, block, shared, stream>>>

// the kernel's dynamic shared memory as `type name[]`
#define GV_DYNAMIC_SHARED(type, name) extern __shared__ __align__(16) type name[]

// bar.sync id, threads: barrier over `threads` threads (a multiple of 32) of the CTA; id in [1, 15]
__device__ __forceinline__ void gv_named_barrier(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

__device__ __forceinline__ unsigned long long gv_global_timer_ns() {
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    return now;
}

// a scheduling fence on a loaded value: what follows is issued only once `value` has arrived
__device__ __forceinline__ void gv_wait_for(float &value) {
    asm volatile("" : "+f"(value)::"memory");
}

// a plain ld.global that the compiler may neither drop nor merge with an earlier load of the same address
__device__ __forceinline__ float gv_load_again(const float *address) {
    float value;
    asm volatile("ld.global.f32 %0, [%1];" : "=f"(value) : "l"(address) : "memory");
    return value;
}

// bring the line holding `address` into L2 without waiting for it
__device__ __forceinline__ void gv_prefetch_l2(const void *address) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(address));
}

// ex2.approx / rcp.approx based (<= 2 ulp each)
__device__ __forceinline__ float gv_fast_exp(float x) {
    return __expf(x);
}
__device__ __forceinline__ float gv_fast_divide(float a, float b) {
    return __fdividef(a, b);
}

#endif
