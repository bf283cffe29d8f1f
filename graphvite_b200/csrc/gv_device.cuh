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
// GV_DYNAMIC_SHARED, gv_named_barrier, gv_global_timer_ns, gv_wait_for, gv_load_again, gv_prefetch_l2, gv_fast_exp,
// gv_fast_divide, gv_fast_rcp, gv_fast_sqrt, gv_fast_rsqrt and gv_prefetch_row_line (which READS a byte there, so that
// a prefetch address outside its allocation faults at the guard page instead of passing unnoticed).
#define GV_DEVICE_INLINE inline

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

// the same for a line of an embedding row (the emulation build dereferences the address to check it)
__device__ __forceinline__ void gv_prefetch_row_line(const void *address) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(address));
}

// ex2.approx / rcp.approx based (<= 2 ulp each)
__device__ __forceinline__ float gv_fast_exp(float x) {
    return __expf(x);
}
__device__ __forceinline__ float gv_fast_divide(float a, float b) {
    return __fdividef(a, b);
}
// MUFU.RCP / MUFU.SQRT / MUFU.RSQ, one instruction each (<= 2 ulp; denormals flush to zero)
__device__ __forceinline__ float gv_fast_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float gv_fast_sqrt(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float gv_fast_rsqrt(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
#define GV_DEVICE_INLINE __device__ __forceinline__

#endif

// sin and cos of one argument in ~22 straight-line instructions, <= 1.6 ulp for |x| <= 48000 (checked against double
// precision over 1.2e8 arguments; tests/test_host_runtime.py repeats the check through the emulation build):
// Cody-Waite reduction by pi/2 in three FMAs (the quotient from the round-to-nearest of a 1.5 * 2^23 bias), the
// classic degree-7 / degree-8 minimax polynomials on [-pi/4, pi/4], quadrant fix-up.  sincosf() costs about twice as
// much per call, carries a Payne-Hanek slow path (local memory, a divergent region per call site) and the
// knowledge-graph kernels call it per element pair and target.  Larger arguments, infinities and NaN take sincosf().
GV_DEVICE_INLINE void gv_sincos(float x, float *sine, float *cosine) {
    if (!(fabsf(x) <= 48000.f)) {
        sincosf(x, sine, cosine);
        return;
    }
    float j = fmaf(x, 0.636619772f, 12582912.f);
    const int quadrant = __float_as_int(j);
    j -= 12582912.f;
    float r = fmaf(j, -1.57079601e+00f, x);
    r = fmaf(j, -3.13916473e-07f, r);
    r = fmaf(j, -5.39030253e-15f, r);
    const float z = r * r;
    float s = -1.95152959e-4f;
    s = fmaf(s, z, 8.33216087e-3f);
    s = fmaf(s, z, -1.66666546e-1f);
    s = fmaf(s * z, r, r);
    float c = 2.44331571e-5f;
    c = fmaf(c, z, -1.38873163e-3f);
    c = fmaf(c, z, 4.16666457e-2f);
    c = fmaf(c, z, -0.5f);
    c = fmaf(c, z, 1.0f);
    const float a = (quadrant & 1) ? c : s, b = (quadrant & 1) ? s : c;
    *sine = (quadrant & 2) ? -a : a;
    *cosine = ((quadrant + 1) & 2) ? -b : b;
}
