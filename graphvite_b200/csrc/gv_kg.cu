// =============================================================================
// gv_kg.cu -- knowledge-graph embedding kernels, hand-written for sm_100a.
//
// Replaces gpu::knowledge_graph::train / train_1_moment / train_2_moment / predict (reference
// include/instance/gpu/knowledge_graph.cuh:38-366) for the models TransE, DistMult, ComplEx, SimplE,
// RotatE and QuatE (include/instance/model/knowledge_graph.h:34-860) with the update rules of
// include/core/optimizer.h:161-210 and the uniform negative draw of gpu::Sample
// (include/base/alias_table.cuh:148-152,175-183 over the all-ones table of
// instance/knowledge_graph.cuh:316-319).
//
// STATUS: sm_100a, 239 registers for 8 floats per thread with RotatE / Adam, no spills.  Parity: tests/test_gpu_zz_kg_*.py
// and tests/test_gpu_zzzz_kg_full_size.py, green on B200s (GPUTEST_r01; profiles/r02_kg_kernel_tests_after_diet.txt for
// the math of this revision) against the oracle and against golden vectors recorded from the reference's own kernels;
// the same files also run under the CUDA emulation of tests/emu.  Measured (profiles/r02b_summary.md): RotatE d = 2048,
// k = 64, Adam: 6.5e5 positives/s per B200 = 0.38 of the HBM roofline; bound by instruction issue at 8 warps per SM
// (1 340 instructions per thread and target), not by memory (DRAM 23 % busy) -- see the math policy below.
//
// Design.  One positive sample = 1 + k targets that share the relation row and, each, either the
// positive head or the positive tail.  The reference walks the targets with one warp and
// read-modify-writes head, tail and relation rows (plus moments) in global memory for every target:
// ~7.8 MB per positive at d = 2048, k = 64 with Adam.  Here a *group* of dim / E threads (E = 2, 4 or 8
// floats per thread as 64- / 128-bit units interleaved over the threads, so that a warp's access is one
// contiguous 512-B run; 256 threads at d = 2048) owns the sample and keeps the relation row,
// the positive head row and the positive tail row -- with their moments -- in registers across all
// targets; only the negative rows travel: read once for the self-adversarial normaliser, then one
// read-modify-write with their moments.  That is ~3.6 MB per positive, with identical sequential
// semantics inside the group (rows that alias a cached row are served from the registers; a target
// whose head and tail are the same row of the same matrix takes an in-order slow path).  Groups race
// against each other Hogwild-style exactly like the reference's warps.
// The logit is a group-wide sum: butterfly shuffles inside a warp, one shared-memory hop (double
// buffered, a single named barrier) across the warps of a group.
// Latency hiding: the logits of the normaliser pass are independent, so kPass1Batch targets are loaded together and
// reduced with one barrier; in the update pass the targets depend on each other through the cached rows, but the
// next target's negative row (+ moments) is requested while the current one is reduced and updated (never a row
// the current target writes).
// =============================================================================
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <string>

#include "gv_common.h"
#include "gv_device.cuh"

namespace gv {
namespace device {

namespace {

constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr float kEps = 1e-15f;  // util/common.h:28
constexpr int kCtaThreads = 256;  // groups are packed into CTAs of this size ...
constexpr int kMaxGroupThreads = 512;  // ... except a group of 4-float threads at d = 2048 (kg_flags bit 2), a CTA of its own
template<int E>
constexpr int max_cta_threads() {
    return E == 8 ? kCtaThreads : kMaxGroupThreads;
}
constexpr int kPass1Batch = 4;  // targets per barrier in the normaliser pass

// tunable `kg_flags` (gv_cuda_set_tunable; environment GV_KG_FLAGS gives the initial value).
// bit 0: IEEE square roots, divisions and sincosf() in the train kernel instead of the MUFU / gv_sincos versions
// bit 1: no L2 prefetch of the negative rows ahead of their targets
// bit 2: 4 floats per thread also for rows of 1024 .. 2048 floats (groups of up to 512 threads)
int &kg_flags() {
    static int flags = getenv("GV_KG_FLAGS") ? atoi(getenv("GV_KG_FLAGS")) : 0;
    return flags;
}

struct KgParams {
    int dim;
    uint32_t num_head;
    int shared;  // head and tail blocks are the same memory
    float *head, *tail, *relation;
    float *head_m1, *tail_m1, *relation_m1;
    float *head_m2, *tail_m2, *relation_m2;
    const uint32_t *batch;  // [n][3] {relation, tail, head}
    unsigned long long num_sample;
    int num_negative;
    const uint32_t *negatives;
    const double *random;
    uint32_t negative_count;
    uint32_t *negatives_out;
    gv_device_optimizer_t optimizer;
    const float *lr_per_batch;
    uint32_t batch_size;
    float relation_lr_multiplier, margin_or_l3, temperature;
    float *loss_per_sample, *loss_per_batch;
    int flags;  // kg_flags
};

// ---- a thread's N floats of a row: N / U vector units of U floats (128-, 64- or 32-bit accesses) ------------
// Unit i of thread c lives at unit index c + i * (threads of the group): consecutive lanes touch consecutive
// units, so every warp-wide access is one contiguous run (32 x 16 B = 512 B for float4) -- with 8 floats per
// thread a thread therefore owns two float4 units half a row apart, not 8 contiguous floats.  `stride` is the
// distance between a thread's units in floats.
template<int N, int U>
__device__ __forceinline__ void load_vec(float (&dst)[N], const float *src, size_t stride) {
    static_assert(N % U == 0 && (U == 4 || U == 2 || U == 1), "unit width");
#pragma unroll
    for (int i = 0; i < N / U; i++) {
        const float *unit = src + i * stride;
        if constexpr (U == 4) {
            const float4 v = __ldcg(reinterpret_cast<const float4 *>(unit));
            dst[i * 4] = v.x, dst[i * 4 + 1] = v.y, dst[i * 4 + 2] = v.z, dst[i * 4 + 3] = v.w;
        } else if constexpr (U == 2) {
            const float2 v = __ldcg(reinterpret_cast<const float2 *>(unit));
            dst[i * 2] = v.x, dst[i * 2 + 1] = v.y;
        } else
            dst[i] = __ldcg(unit);
    }
}

template<int N, int U>
__device__ __forceinline__ void store_vec(float *dst, const float (&src)[N], size_t stride) {
    static_assert(N % U == 0 && (U == 4 || U == 2 || U == 1), "unit width");
#pragma unroll
    for (int i = 0; i < N / U; i++) {
        float *unit = dst + i * stride;
        if constexpr (U == 4)
            __stcg(reinterpret_cast<float4 *>(unit),
                   make_float4(src[i * 4], src[i * 4 + 1], src[i * 4 + 2], src[i * 4 + 3]));
        else if constexpr (U == 2)
            __stcg(reinterpret_cast<float2 *>(unit), make_float2(src[i * 2], src[i * 2 + 1]));
        else
            __stcg(unit, src[i]);
    }
}

// a thread's slice of one (entity) row and of its NM moment rows
template<int N, int NM>
struct Slice {
    static constexpr int U = N >= 4 ? 4 : N;  // floats per vector unit
    float v[N];
    float m1[NM >= 1 ? N : 1];
    float m2[NM >= 2 ? N : 1];
};

template<int N, int NM>
__device__ __forceinline__ void load_slice(Slice<N, NM> &s, const float *v, const float *m1, const float *m2,
                                           size_t offset, size_t stride, bool active) {
    if (active) {
        load_vec<N, Slice<N, NM>::U>(s.v, v + offset, stride);
        if constexpr (NM >= 1)
            load_vec<N, Slice<N, NM>::U>(s.m1, m1 + offset, stride);
        if constexpr (NM >= 2)
            load_vec<N, Slice<N, NM>::U>(s.m2, m2 + offset, stride);
    } else {
#pragma unroll
        for (int i = 0; i < N; i++)
            s.v[i] = 0.f;
    }
}

template<int N, int NM>
__device__ __forceinline__ void store_slice(const Slice<N, NM> &s, float *v, float *m1, float *m2, size_t offset,
                                            size_t stride, bool active) {
    if (!active)
        return;
    store_vec<N, Slice<N, NM>::U>(v + offset, s.v, stride);
    if constexpr (NM >= 1)
        store_vec<N, Slice<N, NM>::U>(m1 + offset, s.m1, stride);
    if constexpr (NM >= 2)
        store_vec<N, Slice<N, NM>::U>(m2 + offset, s.m2, stride);
}

// ---- optimizers, core/optimizer.h:161-210: the step to subtract from `parameter` --------------
struct Opt {
    int type;
    float lr, wd, a, b, eps;
};

// ---- math policy ------------------------------------------------------------------------------------
// FAST (the default, `kg_flags` bit 0 clear): square roots and divisions of the update rules are single MUFU
// instructions (<= 2 ulp each) and RotatE's rotation is gv_sincos (<= 1.6 ulp, no slow path), evaluated ONCE per
// target and shared by the logit and the gradient.  The IEEE versions nvcc emits for `/`, sqrtf() and sincosf() are
// 8-40 instructions each with a divergent slow path per call site; with them the RotatE / Adam kernel executes ~2 000
// instructions per thread and target and is bound by instruction issue at one 256-thread group per SM (0.22 of the
// HBM roofline, even with every row in L2 -- profiles/r02_bench/bench_rotate_n4.json).  !FAST keeps that code.
template<bool FAST>
__device__ __forceinline__ float kg_sqrt(float x) {
    if constexpr (FAST)
        return gv_fast_sqrt(x);
    else
        return sqrtf(x);
}
template<bool FAST>
__device__ __forceinline__ float kg_divide(float a, float b) {
    if constexpr (FAST)
        return a * gv_fast_rcp(b);
    else
        return a / b;
}
template<bool FAST>
__device__ __forceinline__ void kg_sincos(float x, float *sine, float *cosine) {
    if constexpr (FAST)
        gv_sincos(x, sine, cosine);
    else
        sincosf(x, sine, cosine);
}

template<int NM, bool FAST>
__device__ __forceinline__ float step(const Opt &o, float parameter, float gradient, float &m1, float &m2,
                                      float weight) {
    if constexpr (NM == 0)
        return o.lr * weight * (gradient + o.wd * parameter);
    const float regularized = weight * (gradient + o.wd * parameter);
    if constexpr (NM == 2) {
        m1 = o.a * m1 + (1 - o.a) * regularized;
        m2 = o.b * m2 + (1 - o.b) * regularized * regularized;
        return kg_divide<FAST>(o.lr * m1, kg_sqrt<FAST>(m2) + o.eps);
    }
    if (o.type == GV_OPT_MOMENTUM) {
        m1 = o.a * m1 + (1 - o.a) * regularized;
        return o.lr * m1;
    }
    if (o.type == GV_OPT_ADAGRAD) {
        m1 += regularized * regularized;
        return kg_divide<FAST>(o.lr * regularized, kg_sqrt<FAST>(m1) + o.eps);
    }
    m1 = o.a * m1 + (1 - o.a) * regularized * regularized;  // RMSprop
    if constexpr (FAST)
        return o.lr * regularized * gv_fast_rsqrt(m1 + o.eps);
    else
        return o.lr * regularized / sqrtf(m1 + o.eps);
}

// util/math.h:30-44 (precise exponentials: the loss and the adversarial weights are compared with the reference)
template<bool FAST>
__device__ __forceinline__ float sigmoid(float x) {
    if constexpr (FAST) {  // one exponential, one MUFU.RCP: the same two branches of util/math.h
        const float e = expf(-fabsf(x));
        const float inverse = gv_fast_rcp(1 + e);
        return x > 0 ? inverse : e * inverse;
    } else
        return x > 0 ? 1 / (1 + expf(-x)) : expf(x) / (expf(x) + 1);
}
__device__ __forceinline__ float safe_exp(float x) {
    return expf(fminf(fmaxf(x, -80.f), 80.f));
}

// ---- per-model slice geometry --------------------------------------------------------------------
// entity rows: E floats per thread at offset c * E.  Relation VALUES: RotatE keeps dim/2 phases, the
// thread owning pairs [c*E/2, (c+1)*E/2) reads them at offset c * E/2; every other model uses the
// entity geometry.  Relation MOMENTS: RotatE and ComplEx index them by pair (ComplEx updates the same
// moment for the real and the imaginary part, model/knowledge_graph.h:302-306,336-340).
template<int E, int MODEL>
struct Geometry {
    static constexpr int RV = MODEL == GV_KG_ROTATE ? E / 2 : E;
    static constexpr int RM = (MODEL == GV_KG_ROTATE || MODEL == GV_KG_COMPLEX) ? E / 2 : E;
    // floats per vector unit: entity rows, relation values, relation moments.  An entity unit of UE floats holds
    // UE / 2 complex pairs, whose phases (RotatE) / shared moments (RotatE, ComplEx) form a unit of UE / 2 floats at
    // the same unit index, so the local order of pairs is the same in all three.
    static constexpr int UE = E >= 4 ? 4 : E;
    static constexpr int URV = MODEL == GV_KG_ROTATE ? UE / 2 : UE;
    static constexpr int URM = (MODEL == GV_KG_ROTATE || MODEL == GV_KG_COMPLEX) ? UE / 2 : UE;
};

template<int E, int MODEL, int NM>
struct Relation {
    float v[Geometry<E, MODEL>::RV];
    float m1[NM >= 1 ? Geometry<E, MODEL>::RM : 1];
    float m2[NM >= 2 ? Geometry<E, MODEL>::RM : 1];
};

// Model::forward restricted to a thread's slice (model/knowledge_graph.h:44-49,117-123,208-223,359-366,
// 453-468); the caller sums over the group and applies `margin - sum` for TransE / RotatE.
template<int E, int MODEL, bool FAST>
__device__ __forceinline__ float partial_logit(const float (&h)[E], const float (&t)[E],
                                               const float (&r)[Geometry<E, MODEL>::RV]) {
    float output = 0.f;
    if constexpr (MODEL == GV_KG_TRANSE) {
#pragma unroll
        for (int i = 0; i < E; i++)
            output += fabsf(h[i] + r[i] - t[i]);
    } else if constexpr (MODEL == GV_KG_DISTMULT) {
#pragma unroll
        for (int i = 0; i < E; i++)
            output += h[i] * r[i] * t[i];
    } else if constexpr (MODEL == GV_KG_SIMPLE) {
#pragma unroll
        for (int i = 0; i < E; i++)
            output += h[i] * r[i] * t[i ^ 1];
    } else if constexpr (MODEL == GV_KG_COMPLEX) {
#pragma unroll
        for (int i = 0; i < E / 2; i++) {
            const float product_re = h[i * 2] * r[i * 2] - h[i * 2 + 1] * r[i * 2 + 1];
            const float product_im = h[i * 2] * r[i * 2 + 1] + h[i * 2 + 1] * r[i * 2];
            output += product_re * t[i * 2] + product_im * t[i * 2 + 1];
        }
    } else if constexpr (MODEL == GV_KG_QUATE) {  // model/knowledge_graph.h:594-618
#pragma unroll
        for (int i = 0; i < E / 4; i++) {
            const float h_r = h[i * 4], h_i = h[i * 4 + 1], h_j = h[i * 4 + 2], h_k = h[i * 4 + 3];
            const float r_r = r[i * 4], r_i = r[i * 4 + 1], r_j = r[i * 4 + 2], r_k = r[i * 4 + 3];
            const float t_r = t[i * 4], t_i = t[i * 4 + 1], t_j = t[i * 4 + 2], t_k = t[i * 4 + 3];
            const float r_norm = kg_sqrt<FAST>(r_r * r_r + r_i * r_i + r_j * r_j + r_k * r_k);
            const float product_r = h_r * r_r - h_i * r_i - h_j * r_j - h_k * r_k;
            const float product_i = h_r * r_i + h_i * r_r + h_j * r_k - h_k * r_j;
            const float product_j = h_r * r_j - h_i * r_k + h_j * r_r + h_k * r_i;
            const float product_k = h_r * r_k + h_i * r_j - h_j * r_i + h_k * r_r;
            output += kg_divide<FAST>(product_r * t_r + product_i * t_i + product_j * t_j + product_k * t_k,
                                      r_norm + kEps);
        }
    } else {  // RotatE
#pragma unroll
        for (int i = 0; i < E / 2; i++) {
            float r_re, r_im;
            kg_sincos<FAST>(r[i], &r_im, &r_re);
            const float distance_re = h[i * 2] * r_re - h[i * 2 + 1] * r_im - t[i * 2];
            const float distance_im = h[i * 2] * r_im + h[i * 2 + 1] * r_re - t[i * 2 + 1];
            output += kg_sqrt<FAST>(distance_re * distance_re + distance_im * distance_im);
        }
    }
    return output;
}

// RotatE with the relation's rotation (cos, sin of its phases) already evaluated: the relation row does not change
// during the normaliser pass, so its sincosf are hoisted out of the k targets (same values, same order of operations)
template<int E, bool FAST>
__device__ __forceinline__ float partial_logit_rotated(const float (&h)[E], const float (&t)[E],
                                                       const float (&r_re)[E / 2], const float (&r_im)[E / 2]) {
    float output = 0.f;
#pragma unroll
    for (int i = 0; i < E / 2; i++) {
        const float distance_re = h[i * 2] * r_re[i] - h[i * 2 + 1] * r_im[i] - t[i * 2];
        const float distance_im = h[i * 2] * r_im[i] + h[i * 2 + 1] * r_re[i] - t[i * 2 + 1];
        output += kg_sqrt<FAST>(distance_re * distance_re + distance_im * distance_im);
    }
    return output;
}

// Model::backward on a thread's slices, statement order of the reference kept
// (model/knowledge_graph.h:51-108,125-190,225-340,368-433,470-575).  ALIAS: head and tail are the SAME
// row of the same matrix; every tail access then goes to the head slice (and its moments), which
// reproduces the reference's two successive read-modify-writes of one memory location.
// `rotation_re / rotation_im` (RotatE with FAST only): cos / sin of R.v as the logit of this target used them.
template<int E, int MODEL, int NM, bool ALIAS, bool FAST>
__device__ __forceinline__ void backward(Slice<E, NM> &H, Slice<E, NM> &Tin, Relation<E, MODEL, NM> &R, const Opt &o,
                                         float margin_or_l3, float gradient, float relation_lr_multiplier,
                                         float weight, const float (&rotation_re)[E / 2],
                                         const float (&rotation_im)[E / 2]) {
    Slice<E, NM> &T = ALIAS ? H : Tin;
    if constexpr (MODEL == GV_KG_TRANSE) {
#pragma unroll
        for (int i = 0; i < E; i++) {
            const float h = H.v[i], t = T.v[i], r = R.v[i];
            const float s = h + r - t > 0 ? 1.f : -1.f;
            H.v[i] -= step<NM, FAST>(o, h, -gradient * s, H.m1[NM >= 1 ? i : 0], H.m2[NM >= 2 ? i : 0], weight);
            T.v[i] -= step<NM, FAST>(o, t, gradient * s, T.m1[NM >= 1 ? i : 0], T.m2[NM >= 2 ? i : 0], weight);
            R.v[i] -= relation_lr_multiplier *
                      step<NM, FAST>(o, r, -gradient * s, R.m1[NM >= 1 ? i : 0], R.m2[NM >= 2 ? i : 0], weight);
        }
    } else if constexpr (MODEL == GV_KG_DISTMULT || MODEL == GV_KG_SIMPLE) {
        const float l3 = margin_or_l3 * 3;
#pragma unroll
        for (int i = 0; i < E; i++) {
            const int j = MODEL == GV_KG_SIMPLE ? (i ^ 1) : i;
            const float h = H.v[i], t = T.v[j], r = R.v[i];
            H.v[i] -= step<NM, FAST>(o, h, gradient * r * t + l3 * fabsf(h) * h, H.m1[NM >= 1 ? i : 0],
                               H.m2[NM >= 2 ? i : 0], weight);
            T.v[j] -= step<NM, FAST>(o, t, gradient * h * r + l3 * fabsf(t) * t, T.m1[NM >= 1 ? j : 0],
                               T.m2[NM >= 2 ? j : 0], weight);
            R.v[i] -= relation_lr_multiplier * step<NM, FAST>(o, r, gradient * h * t + l3 * fabsf(r) * r,
                                                        R.m1[NM >= 1 ? i : 0], R.m2[NM >= 2 ? i : 0], weight);
        }
    } else if constexpr (MODEL == GV_KG_COMPLEX) {
        const float l3 = margin_or_l3 * 3;
#pragma unroll
        for (int i = 0; i < E / 2; i++) {
            const int re = i * 2, im = i * 2 + 1;
            const float h_re = H.v[re], h_im = H.v[im], t_re = T.v[re], t_im = T.v[im];
            const float r_re = R.v[re], r_im = R.v[im];
            const float h_re_grad = gradient * (r_re * t_re + r_im * t_im);
            const float h_im_grad = gradient * (-r_im * t_re + r_re * t_im);
            H.v[re] -= step<NM, FAST>(o, h_re, h_re_grad + l3 * fabsf(h_re) * h_re, H.m1[NM >= 1 ? re : 0],
                                H.m2[NM >= 2 ? re : 0], weight);
            H.v[im] -= step<NM, FAST>(o, h_im, h_im_grad + l3 * fabsf(h_im) * h_im, H.m1[NM >= 1 ? im : 0],
                                H.m2[NM >= 2 ? im : 0], weight);
            const float t_re_grad = gradient * (h_re * r_re - h_im * r_im);
            const float t_im_grad = gradient * (h_re * r_im + h_im * r_re);
            T.v[re] -= step<NM, FAST>(o, t_re, t_re_grad + l3 * fabsf(t_re) * t_re, T.m1[NM >= 1 ? re : 0],
                                T.m2[NM >= 2 ? re : 0], weight);
            T.v[im] -= step<NM, FAST>(o, t_im, t_im_grad + l3 * fabsf(t_im) * t_im, T.m1[NM >= 1 ? im : 0],
                                T.m2[NM >= 2 ? im : 0], weight);
            const float r_re_grad = gradient * (h_re * t_re + h_im * t_im);
            const float r_im_grad = gradient * (-h_im * t_re + h_re * t_im);
            R.v[re] -= relation_lr_multiplier * step<NM, FAST>(o, r_re, r_re_grad + l3 * fabsf(r_re) * r_re,
                                                         R.m1[NM >= 1 ? i : 0], R.m2[NM >= 2 ? i : 0], weight);
            R.v[im] -= relation_lr_multiplier * step<NM, FAST>(o, r_im, r_im_grad + l3 * fabsf(r_im) * r_im,
                                                         R.m1[NM >= 1 ? i : 0], R.m2[NM >= 2 ? i : 0], weight);
        }
    } else if constexpr (MODEL == GV_KG_QUATE) {  // model/knowledge_graph.h:620-674 (and the moment overloads)
        const float l3 = margin_or_l3 * 3;
#pragma unroll
        for (int i = 0; i < E / 4; i++) {
            const int q = i * 4;
            const float h_r = H.v[q], h_i = H.v[q + 1], h_j = H.v[q + 2], h_k = H.v[q + 3];
            const float r_r = R.v[q], r_i = R.v[q + 1], r_j = R.v[q + 2], r_k = R.v[q + 3];
            const float t_r = T.v[q], t_i = T.v[q + 1], t_j = T.v[q + 2], t_k = T.v[q + 3];
            const float r_norm = kg_sqrt<FAST>(r_r * r_r + r_i * r_i + r_j * r_j + r_k * r_k);
            const float grad = kg_divide<FAST>(gradient, r_norm + kEps);
            const float head_grad[4] = {grad * (r_r * t_r + r_i * t_i + r_j * t_j + r_k * t_k),
                                        grad * (-r_i * t_r + r_r * t_i - r_k * t_j + r_j * t_k),
                                        grad * (-r_j * t_r + r_k * t_i + r_r * t_j - r_i * t_k),
                                        grad * (-r_k * t_r - r_j * t_i + r_i * t_j + r_r * t_k)};
            const float head_old[4] = {h_r, h_i, h_j, h_k};
#pragma unroll
            for (int c = 0; c < 4; c++)
                H.v[q + c] -= step<NM, FAST>(o, head_old[c], head_grad[c] + l3 * fabsf(head_old[c]) * head_old[c],
                                       H.m1[NM >= 1 ? q + c : 0], H.m2[NM >= 2 ? q + c : 0], weight);
            const float tail_grad[4] = {grad * (h_r * r_r - h_i * r_i - h_j * r_j - h_k * r_k),
                                        grad * (h_r * r_i + h_i * r_r + h_j * r_k - h_k * r_j),
                                        grad * (h_r * r_j - h_i * r_k + h_j * r_r + h_k * r_i),
                                        grad * (h_r * r_k + h_i * r_j - h_j * r_i + h_k * r_r)};
            const float tail_old[4] = {t_r, t_i, t_j, t_k};
#pragma unroll
            for (int c = 0; c < 4; c++)
                T.v[q + c] -= step<NM, FAST>(o, tail_old[c], tail_grad[c] + l3 * fabsf(tail_old[c]) * tail_old[c],
                                       T.m1[NM >= 1 ? q + c : 0], T.m2[NM >= 2 ? q + c : 0], weight);
            const float relation_grad[4] = {grad * (h_r * t_r + h_i * t_i + h_j * t_j + h_k * t_k),
                                            grad * (-h_i * t_r + h_r * t_i + h_k * t_j - h_j * t_k),
                                            grad * (-h_j * t_r - h_k * t_i + h_r * t_j + h_i * t_k),
                                            grad * (-h_k * t_r + h_j * t_i - h_i * t_j + h_r * t_k)};
            const float relation_old[4] = {r_r, r_i, r_j, r_k};
#pragma unroll
            for (int c = 0; c < 4; c++)
                R.v[q + c] -= relation_lr_multiplier *
                              step<NM, FAST>(o, relation_old[c],
                                       relation_grad[c] + l3 * fabsf(relation_old[c]) * relation_old[c],
                                       R.m1[NM >= 1 ? q + c : 0], R.m2[NM >= 2 ? q + c : 0], weight);
        }
    } else {  // RotatE
#pragma unroll
        for (int i = 0; i < E / 2; i++) {
            const int re = i * 2, im = i * 2 + 1;
            const float phase = R.v[i];
            float r_re, r_im;
            if constexpr (FAST)
                r_re = rotation_re[i], r_im = rotation_im[i];
            else
                sincosf(phase, &r_im, &r_re);
            const float h_re = H.v[re], h_im = H.v[im], t_re = T.v[re], t_im = T.v[im];
            const float distance_re = h_re * r_re - h_im * r_im - t_re;
            const float distance_im = h_re * r_im + h_im * r_re - t_im;
            const float grad =
                kg_divide<FAST>(gradient, kg_sqrt<FAST>(distance_re * distance_re + distance_im * distance_im) + kEps);
            const float head_re_grad = -grad * (distance_re * r_re + distance_im * r_im);
            const float head_im_grad = -grad * (-distance_re * r_im + distance_im * r_re);
            H.v[re] -= step<NM, FAST>(o, h_re, head_re_grad, H.m1[NM >= 1 ? re : 0], H.m2[NM >= 2 ? re : 0], weight);
            H.v[im] -= step<NM, FAST>(o, h_im, head_im_grad, H.m1[NM >= 1 ? im : 0], H.m2[NM >= 2 ? im : 0], weight);
            T.v[re] -= step<NM, FAST>(o, t_re, grad * distance_re, T.m1[NM >= 1 ? re : 0], T.m2[NM >= 2 ? re : 0], weight);
            T.v[im] -= step<NM, FAST>(o, t_im, grad * distance_im, T.m1[NM >= 1 ? im : 0], T.m2[NM >= 2 ? im : 0], weight);
            const float relation_grad = -grad * (distance_re * (h_re * -r_im + h_im * -r_re) +
                                                 distance_im * (h_re * r_re + h_im * -r_im));
            R.v[i] -= relation_lr_multiplier *
                      step<NM, FAST>(o, phase, relation_grad, R.m1[NM >= 1 ? i : 0], R.m2[NM >= 2 ? i : 0], weight);
        }
    }
}

// ---- a group of threads that owns one sample --------------------------------------------------------
struct Group {
    int threads;       // multiple of 32
    int id_in_cta;     // named-barrier id - 1
    int warps;         // threads / 32
    int lane, warp;    // of this thread inside the group
    float *scratch;    // [2][warps]
    int parity;

    __device__ __forceinline__ void sync() const {
        if (warps == 1)
            __syncwarp();
        else
            gv_named_barrier(id_in_cta + 1, threads);
    }
    // sum over the group, the same value in every thread
    __device__ __forceinline__ float sum(float value) {
#pragma unroll
        for (int delta = 16; delta > 0; delta >>= 1)
            value += __shfl_xor_sync(kFull, value, delta);
        if (warps == 1)
            return value;
        // (both kinds of sum alternate between the SAME two regions: a fast warp may already write the next sum's
        // partials while a slow one still reads this one's, so consecutive sums must never share memory)
        float *slot = scratch + parity * kPass1Batch * warps;
        if (lane == 0)
            slot[warp] = value;
        sync();
        float total = 0.f;
        for (int w = 0; w < warps; w++)
            total += slot[w];
        parity ^= 1;  // the next sum uses the other buffer: one barrier per sum is enough
        return total;
    }
    // N independent sums with ONE barrier (scratch holds [2][N][warps]); every value is reduced exactly like sum()
    template<int N>
    __device__ __forceinline__ void sum_many(float (&value)[N]) {
#pragma unroll
        for (int n = 0; n < N; n++)
#pragma unroll
            for (int delta = 16; delta > 0; delta >>= 1)
                value[n] += __shfl_xor_sync(kFull, value[n], delta);
        if (warps == 1)
            return;
        float *slot = scratch + parity * N * warps;
        if (lane == 0)
#pragma unroll
            for (int n = 0; n < N; n++)
                slot[n * warps + warp] = value[n];
        sync();
#pragma unroll
        for (int n = 0; n < N; n++) {
            float total = 0.f;
            for (int w = 0; w < warps; w++)
                total += slot[n * warps + w];
            value[n] = total;
        }
        parity ^= 1;
    }
};



template<int MODEL>
__device__ __forceinline__ float finish_logit(float sum, float margin_or_l3) {
    return (MODEL == GV_KG_TRANSE || MODEL == GV_KG_ROTATE) ? margin_or_l3 - sum : sum;
}

// gpu::Sample over the all-ones alias table: prob == 1 and alias[i] == i, so the draw is the index
__device__ __forceinline__ uint32_t uniform_negative(uint32_t count, double random1) {
    const float rand1 = float(random1);
    const uint32_t index = uint32_t(double(rand1) * double(count));
    return min(index, count - 1);
}

// -----------------------------------------------------------------------------
// The train kernel.  E floats per thread, NM moments per row.
// Dynamic shared memory per group: 2 * kPass1Batch * warps floats (sums) + num_negative ids.
// -----------------------------------------------------------------------------
template<int E, int MODEL, int NM, bool FAST>
__global__ void __launch_bounds__(max_cta_threads<E>()) kg_train_kernel(const KgParams p) {
    GV_DYNAMIC_SHARED(unsigned char, shared_bytes);
    using G = Geometry<E, MODEL>;
    const int chunks = p.dim / E;                       // active threads of a group
    const int group_threads = (chunks + 31) / 32 * 32;  // blockDim.x is a multiple of this
    const int groups_per_cta = blockDim.x / group_threads;
    Group g;
    g.threads = group_threads;
    g.warps = group_threads / 32;
    g.id_in_cta = threadIdx.x / group_threads;
    const int c = threadIdx.x % group_threads;
    g.lane = c & 31;
    g.warp = c >> 5;
    g.parity = 0;
    const size_t scratch_bytes = size_t(2) * kPass1Batch * g.warps * sizeof(float);
    const size_t per_group = scratch_bytes + size_t(p.num_negative) * sizeof(uint32_t);
    unsigned char *mine = shared_bytes + per_group * g.id_in_cta;
    g.scratch = reinterpret_cast<float *>(mine);
    uint32_t *negative_ids = reinterpret_cast<uint32_t *>(mine + scratch_bytes);
    const bool active = c < chunks;
    // this thread's first vector unit inside an entity row / relation row / relation moment row, and the distance
    // between its units (see load_vec)
    const size_t slice = size_t(c) * G::UE, slice_stride = size_t(chunks) * G::UE;
    const size_t relation_value = size_t(c) * G::URV, value_stride = size_t(chunks) * G::URV;
    const size_t relation_moment = size_t(c) * G::URM, moment_stride = size_t(chunks) * G::URM;
    const size_t dim = p.dim;
    const int k = p.num_negative;
    const bool adversarial = p.temperature > kEps;

    Opt o;
    o.type = p.optimizer.type;
    o.wd = p.optimizer.weight_decay;
    o.a = p.optimizer.a;
    o.b = p.optimizer.b;
    o.eps = p.optimizer.epsilon;

    // L2 prefetch geometry (see prefetch_row): this thread's array (0 = row, 1 / 2 = moments) and 128-byte line
    const bool l2_prefetch = !(p.flags & 2);
    const int row_lines = (p.dim * 4 + 127) / 128;
    const int prefetch_array = c / row_lines, prefetch_line = c % row_lines;

    const unsigned long long first = (unsigned long long)blockIdx.x * groups_per_cta + g.id_in_cta;
    const unsigned long long stride = (unsigned long long)gridDim.x * groups_per_cta;
    for (unsigned long long sample = first; sample < p.num_sample; sample += stride) {
        const uint32_t relation_id = __ldg(p.batch + sample * 3);
        const uint32_t positive_tail = __ldg(p.batch + sample * 3 + 1);
        const uint32_t positive_head = __ldg(p.batch + sample * 3 + 2);
        o.lr = __ldg(p.lr_per_batch + sample / p.batch_size);

        // the sample's negatives: given, or drawn from the random stream like train_batch does
        for (int s = c; s < k; s += group_threads) {
            const unsigned long long t = sample * k + s;
            uint32_t negative;
            if (p.negatives)
                negative = __ldcs(p.negatives + t);
            else
                negative = uniform_negative(p.negative_count, __ldcs(p.random + t * 2));
            negative_ids[s] = negative;
            if (p.negatives_out)
                p.negatives_out[t] = negative;
        }
        g.sync();

        // cached rows: relation, positive head, positive tail (+ moments)
        Relation<E, MODEL, NM> R;
        {
            const size_t base = size_t(relation_id) * dim;
            if (active) {
                load_vec<G::RV, G::URV>(R.v, p.relation + base + relation_value, value_stride);
                if constexpr (NM >= 1)
                    load_vec<G::RM, G::URM>(R.m1, p.relation_m1 + base + relation_moment, moment_stride);
                if constexpr (NM >= 2)
                    load_vec<G::RM, G::URM>(R.m2, p.relation_m2 + base + relation_moment, moment_stride);
            } else {
#pragma unroll
                for (int i = 0; i < G::RV; i++)
                    R.v[i] = 0.f;
            }
        }
        // a self loop inside one block: the two cached rows would be the same memory -- do not cache
        const bool cached = !(p.shared && positive_head == positive_tail);
        Slice<E, NM> PH, PT;
        if (cached) {
            load_slice<E, NM>(PH, p.head, p.head_m1, p.head_m2, size_t(positive_head) * dim + slice, slice_stride, active);
            load_slice<E, NM>(PT, p.tail, p.tail_m1, p.tail_m2, size_t(positive_tail) * dim + slice, slice_stride, active);
        }

        // ids of target s (s == k: the positive triple)
        auto target = [&](int s, uint32_t &head_id, uint32_t &tail_id) {
            head_id = positive_head;
            tail_id = positive_tail;
            if (s < k) {
                const uint32_t negative = negative_ids[s];
                if (negative < p.num_head)
                    head_id = negative;
                else
                    tail_id = negative - p.num_head;
            }
        };

        // The negative rows of a sample are scattered 8-KB reads whose ids are known from the start, and one group per
        // SM has only the next target's row in flight (NX below): rows further ahead are pulled into L2 -- the row
        // and its moments are (1 + NM) * row_lines 128-byte lines, one per thread (always fewer than the group has
        // threads).  A prefetch is not a read: the values are loaded when the target is reached.
        auto prefetch_row = [&](int s, int arrays) {
            if (!l2_prefetch || s >= k || prefetch_array >= arrays)
                return;
            const uint32_t negative = negative_ids[s];
            const bool is_head = negative < p.num_head;
            const size_t row = size_t(is_head ? negative : negative - p.num_head) * dim;
            const float *base = prefetch_array == 0 ? (is_head ? p.head : p.tail)
                                : prefetch_array == 1 ? (is_head ? p.head_m1 : p.tail_m1)
                                                      : (is_head ? p.head_m2 : p.tail_m2);
            gv_prefetch_row_line(reinterpret_cast<const char *>(base + row) + prefetch_line * 128);
        };
        auto prefetch_row_of = [&](int s) {  // values only, line prefetch_line of target s
            if (!l2_prefetch || s >= k)
                return;
            const uint32_t negative = negative_ids[s];
            const bool is_head = negative < p.num_head;
            const float *row = (is_head ? p.head : p.tail) + size_t(is_head ? negative : negative - p.num_head) * dim;
            gv_prefetch_row_line(reinterpret_cast<const char *>(row) + prefetch_line * 128);
        };
        prefetch_row(0, 1 + NM);  // pass 2 starts with these; pass 1 (values only) runs in between
        prefetch_row(1, 1 + NM);

        // pass 1: normaliser of the self-adversarial weights (gpu/knowledge_graph.cuh:59-77).  The k logits are
        // independent of each other: kPass1Batch targets are loaded together (that many rows in flight per
        // thread) and reduced with one barrier; the normaliser is still accumulated in target order.
        float bias = 0.f, normalizer = 0.f;
        float rotation_re[E / 2], rotation_im[E / 2];
        if constexpr (MODEL == GV_KG_ROTATE)
            if (adversarial) {
#pragma unroll
                for (int i = 0; i < E / 2; i++)
                    kg_sincos<FAST>(R.v[i], &rotation_im[i], &rotation_re[i]);
            }
        if (adversarial)
            for (int s0 = 0; s0 < k; s0 += kPass1Batch) {
                // the rows of the NEXT batch of targets -> L2: thread c takes line (c % row_lines) of target
                // s0 + kPass1Batch + c / row_lines
                if (prefetch_array < kPass1Batch)
                    prefetch_row_of(s0 + kPass1Batch + prefetch_array);
                float partial[kPass1Batch];
#pragma unroll
                for (int b = 0; b < kPass1Batch; b++) {
                    partial[b] = 0.f;
                    if (s0 + b < k && active) {
                        uint32_t head_id, tail_id;
                        target(s0 + b, head_id, tail_id);
                        float h[E], t[E];
                        if (cached && head_id == positive_head) {
#pragma unroll
                            for (int i = 0; i < E; i++)
                                h[i] = PH.v[i];
                        } else
                            load_vec<E, G::UE>(h, p.head + size_t(head_id) * dim + slice, slice_stride);
                        if (cached && tail_id == positive_tail) {
#pragma unroll
                            for (int i = 0; i < E; i++)
                                t[i] = PT.v[i];
                        } else
                            load_vec<E, G::UE>(t, p.tail + size_t(tail_id) * dim + slice, slice_stride);
                        if constexpr (MODEL == GV_KG_ROTATE)
                            partial[b] = partial_logit_rotated<E, FAST>(h, t, rotation_re, rotation_im);
                        else
                            partial[b] = partial_logit<E, MODEL, FAST>(h, t, R.v);
                    }
                }
                g.sum_many<kPass1Batch>(partial);
#pragma unroll
                for (int b = 0; b < kPass1Batch; b++)
                    if (s0 + b < k) {
                        const float logit = finish_logit<MODEL>(partial[b], p.margin_or_l3);
                        if (s0 + b == 0)
                            bias = logit;
                        normalizer += safe_exp(kg_divide<FAST>(logit - bias, p.temperature));
                    }
            }

        // pass 2: negatives first, the positive triple last (gpu/knowledge_graph.cuh:79-118)
        // The targets of a sample are sequentially dependent through the cached rows, but the NEXT target's
        // negative row (with its moments) can already be in flight while this one is reduced and updated: it is
        // requested right after this target's own rows, unless it is a row this target is about to write.
        float sample_loss = 0.f;
        Slice<E, NM> NX;
        bool next_valid = false, next_is_head = false;
        for (int s = 0; s <= k; s++) {
            uint32_t head_id, tail_id;
            target(s, head_id, tail_id);
            const bool head_cached = cached && head_id == positive_head;
            const bool tail_cached = cached && tail_id == positive_tail;
            const bool alias = p.shared && head_id == tail_id;  // one row of one matrix in both roles
            const size_t head_offset = size_t(head_id) * dim + slice, tail_offset = size_t(tail_id) * dim + slice;
            // Work on copies WH / WT so that every register array keeps static indexing; a cached row is
            // copied in and back (register moves).  A target that aliases never has both rows cached (a
            // cached pair is never a self loop); if one of them is cached, that slice IS the row.
            Slice<E, NM> WH, WT;
            const int head_source = head_cached ? 1 : ((alias && tail_cached) ? 2 : 0);  // 0 global, 1 PH, 2 PT
            if (head_source == 1)
                WH = PH;
            else if (head_source == 2)
                WH = PT;
            else if (next_valid && next_is_head)
                WH = NX;
            else
                load_slice<E, NM>(WH, p.head, p.head_m1, p.head_m2, head_offset, slice_stride, active);
            if (!alias) {
                if (tail_cached)
                    WT = PT;
                else if (next_valid && !next_is_head)
                    WT = NX;
                else
                    load_slice<E, NM>(WT, p.tail, p.tail_m1, p.tail_m2, tail_offset, slice_stride, active);
            }
            next_valid = false;
            prefetch_row(s + 2, 1 + NM);  // two targets ahead -> L2; the next one -> registers:
            if (s + 1 < k && cached) {  // target s + 1 is a negative: exactly one of its rows may be uncached
                uint32_t next_head, next_tail;
                target(s + 1, next_head, next_tail);
                const bool next_head_cached = next_head == positive_head, next_tail_cached = next_tail == positive_tail;
                const bool next_alias = p.shared && next_head == next_tail;
                if (!next_alias && next_head_cached != next_tail_cached) {
                    next_is_head = !next_head_cached;
                    const uint32_t next_id = next_is_head ? next_head : next_tail;
                    bool written_now = false;  // is it a row this target stores below?
                    if (head_source == 0)
                        written_now |= (p.shared || next_is_head) && next_id == head_id;
                    if (!alias && !tail_cached)
                        written_now |= (p.shared || !next_is_head) && next_id == tail_id;
                    if (!written_now) {
                        const size_t offset = size_t(next_id) * dim + slice;
                        if (next_is_head)
                            load_slice<E, NM>(NX, p.head, p.head_m1, p.head_m2, offset, slice_stride, active);
                        else
                            load_slice<E, NM>(NX, p.tail, p.tail_m1, p.tail_m2, offset, slice_stride, active);
                        next_valid = true;
                    }
                }
            }

            float partial = 0.f;
            if constexpr (MODEL == GV_KG_ROTATE && FAST) {
                // the rotation of this target's relation row, once: the logit and the gradient use the same values
                if (active) {
#pragma unroll
                    for (int i = 0; i < E / 2; i++)
                        gv_sincos(R.v[i], &rotation_im[i], &rotation_re[i]);
                    partial = alias ? partial_logit_rotated<E, FAST>(WH.v, WH.v, rotation_re, rotation_im)
                                    : partial_logit_rotated<E, FAST>(WH.v, WT.v, rotation_re, rotation_im);
                }
            } else if (active)
                partial = alias ? partial_logit<E, MODEL, FAST>(WH.v, WH.v, R.v)
                                : partial_logit<E, MODEL, FAST>(WH.v, WT.v, R.v);
            const float logit = finish_logit<MODEL>(g.sum(partial), p.margin_or_l3);
            const float prob = sigmoid<FAST>(logit);
            float gradient, weight;
            if (s == k) {
                gradient = prob - 1;
                weight = 1;
                if (c == 0)  // the loss is reported by the group's first thread only
                    sample_loss += weight * -logf(prob + kEps);
            } else {
                gradient = prob;
                if (adversarial)
                    weight = fminf(kg_divide<FAST>(safe_exp(kg_divide<FAST>(logit - bias, p.temperature)), normalizer),
                                   1.f);
                else
                    weight = float(1.0 / k);
                if (c == 0)
                    sample_loss += weight * -logf(1 - prob + kEps);
            }
            if (active) {
                if (alias)
                    backward<E, MODEL, NM, true, FAST>(WH, WH, R, o, p.margin_or_l3, gradient,
                                                       p.relation_lr_multiplier, weight, rotation_re, rotation_im);
                else
                    backward<E, MODEL, NM, false, FAST>(WH, WT, R, o, p.margin_or_l3, gradient,
                                                        p.relation_lr_multiplier, weight, rotation_re, rotation_im);
            }
            if (head_source == 1)
                PH = WH;
            else if (head_source == 2)
                PT = WH;
            else
                store_slice<E, NM>(WH, p.head, p.head_m1, p.head_m2, head_offset, slice_stride, active);
            if (!alias) {
                if (tail_cached)
                    PT = WT;
                else
                    store_slice<E, NM>(WT, p.tail, p.tail_m1, p.tail_m2, tail_offset, slice_stride, active);
            }
        }

        // write the cached rows back
        if (cached) {
            store_slice<E, NM>(PH, p.head, p.head_m1, p.head_m2, size_t(positive_head) * dim + slice, slice_stride, active);
            store_slice<E, NM>(PT, p.tail, p.tail_m1, p.tail_m2, size_t(positive_tail) * dim + slice, slice_stride, active);
        }
        if (active) {
            const size_t base = size_t(relation_id) * dim;
            store_vec<G::RV, G::URV>(p.relation + base + relation_value, R.v, value_stride);
            if constexpr (NM >= 1)
                store_vec<G::RM, G::URM>(p.relation_m1 + base + relation_moment, R.m1, moment_stride);
            if constexpr (NM >= 2)
                store_vec<G::RM, G::URM>(p.relation_m2 + base + relation_moment, R.m2, moment_stride);
        }
        if (c == 0) {
            const float loss = sample_loss / 2;
            if (p.loss_per_sample)
                p.loss_per_sample[sample] = loss;
            if (p.loss_per_batch)
                atomicAdd(p.loss_per_batch + sample / p.batch_size, loss);
        }
        g.sync();  // negative_ids are rewritten by the next sample
    }
}

// gpu::knowledge_graph::predict, instance/gpu/knowledge_graph.cuh:341-366: batch rows {relation, tail, head}
template<int E, int MODEL>
__global__ void __launch_bounds__(kCtaThreads) kg_predict_kernel(const float *head, const float *tail,
                                                                 const float *relation, int dim_, const uint32_t *batch,
                                                                 unsigned long long num_sample, float margin,
                                                                 float *logits) {
    GV_DYNAMIC_SHARED(unsigned char, shared_bytes);
    using G = Geometry<E, MODEL>;
    const int chunks = dim_ / E;
    const int group_threads = (chunks + 31) / 32 * 32;
    const int groups_per_cta = blockDim.x / group_threads;
    Group g;
    g.threads = group_threads;
    g.warps = group_threads / 32;
    g.id_in_cta = threadIdx.x / group_threads;
    const int c = threadIdx.x % group_threads;
    g.lane = c & 31;
    g.warp = c >> 5;
    g.parity = 0;
    g.scratch = reinterpret_cast<float *>(shared_bytes) + size_t(2) * kPass1Batch * g.warps * g.id_in_cta;
    const bool active = c < chunks;
    const size_t dim = dim_;
    for (unsigned long long sample = (unsigned long long)blockIdx.x * groups_per_cta + g.id_in_cta; sample < num_sample;
         sample += (unsigned long long)gridDim.x * groups_per_cta) {
        const uint32_t relation_id = __ldg(batch + sample * 3);
        const uint32_t tail_id = __ldg(batch + sample * 3 + 1);
        const uint32_t head_id = __ldg(batch + sample * 3 + 2);
        float partial = 0.f;
        if (active) {
            float h[E], t[E], r[G::RV];
            load_vec<E, G::UE>(h, head + size_t(head_id) * dim + size_t(c) * G::UE, size_t(chunks) * G::UE);
            load_vec<E, G::UE>(t, tail + size_t(tail_id) * dim + size_t(c) * G::UE, size_t(chunks) * G::UE);
            load_vec<G::RV, G::URV>(r, relation + size_t(relation_id) * dim + size_t(c) * G::URV,
                                    size_t(chunks) * G::URV);
            partial = partial_logit<E, MODEL, false>(h, t, r);
        }
        const float logit = finish_logit<MODEL>(g.sum(partial), margin);
        if (c == 0)
            logits[sample] = logit;
    }
}

int floats_per_thread(int dim, int model, bool train) {
    // kg_flags bit 2 (train kernel only): 4 floats per thread for the largest rows as well -- twice the warps per
    // sample (16 per SM at d = 2048), half the registers and instructions per thread
    const bool narrow = train && (kg_flags() & 4) && dim / 4 <= kMaxGroupThreads;
    if (dim % 8 == 0 && dim >= 256 && !narrow)
        return 8;
    if (dim % 4 == 0 && (dim >= 64 || model == GV_KG_QUATE))  // a quaternion never straddles two threads
        return 4;
    return 2;
}

template<int E, int MODEL, bool FAST>
cudaError_t launch_train_math(const KgParams &p, int num_moment, dim3 grid, dim3 block, size_t shared,
                              cudaStream_t s) {
    if (num_moment == 0)
        GV_LAUNCH(grid, block, shared, s, kg_train_kernel<E, MODEL, 0, FAST>)(p);
    else if (num_moment == 1)
        GV_LAUNCH(grid, block, shared, s, kg_train_kernel<E, MODEL, 1, FAST>)(p);
    else
        GV_LAUNCH(grid, block, shared, s, kg_train_kernel<E, MODEL, 2, FAST>)(p);
    return cudaGetLastError();
}

template<int E, int MODEL>
cudaError_t launch_train_nm(const KgParams &p, int num_moment, dim3 grid, dim3 block, size_t shared, cudaStream_t s) {
    if (kg_flags() & 1)  // IEEE sqrt / division / sincosf
        return launch_train_math<E, MODEL, false>(p, num_moment, grid, block, shared, s);
    return launch_train_math<E, MODEL, true>(p, num_moment, grid, block, shared, s);
}

template<int E>
cudaError_t launch_train(const KgParams &p, int model, int num_moment, dim3 grid, dim3 block, size_t shared,
                         cudaStream_t s) {
    switch (model) {
        case GV_KG_TRANSE: return launch_train_nm<E, GV_KG_TRANSE>(p, num_moment, grid, block, shared, s);
        case GV_KG_DISTMULT: return launch_train_nm<E, GV_KG_DISTMULT>(p, num_moment, grid, block, shared, s);
        case GV_KG_COMPLEX: return launch_train_nm<E, GV_KG_COMPLEX>(p, num_moment, grid, block, shared, s);
        case GV_KG_SIMPLE: return launch_train_nm<E, GV_KG_SIMPLE>(p, num_moment, grid, block, shared, s);
        case GV_KG_QUATE:
            if constexpr (E % 4 == 0)
                return launch_train_nm<E, GV_KG_QUATE>(p, num_moment, grid, block, shared, s);
            return cudaErrorInvalidValue;
        default: return launch_train_nm<E, GV_KG_ROTATE>(p, num_moment, grid, block, shared, s);
    }
}

template<int E>
cudaError_t launch_predict(int model, const float *head, const float *tail, const float *relation, int dim,
                           const uint32_t *batch, unsigned long long n, float margin, float *logits, dim3 grid,
                           dim3 block, size_t shared, cudaStream_t s) {
#define GV_PREDICT(M) GV_LAUNCH(grid, block, shared, s, kg_predict_kernel<E, M>)(head, tail, relation, dim, batch, n, margin, logits)
    switch (model) {
        case GV_KG_TRANSE: GV_PREDICT(GV_KG_TRANSE); break;
        case GV_KG_DISTMULT: GV_PREDICT(GV_KG_DISTMULT); break;
        case GV_KG_COMPLEX: GV_PREDICT(GV_KG_COMPLEX); break;
        case GV_KG_SIMPLE: GV_PREDICT(GV_KG_SIMPLE); break;
        case GV_KG_QUATE:
            if constexpr (E % 4 == 0)
                GV_PREDICT(GV_KG_QUATE);
            else
                return cudaErrorInvalidValue;
            break;
        default: GV_PREDICT(GV_KG_ROTATE);
    }
#undef GV_PREDICT
    return cudaGetLastError();
}

int check_geometry(const char *who, int dim, int model, int &E, int &group_threads, bool train = false) {
    if (model < GV_KG_TRANSE || model > GV_KG_QUATE)
        return fail(std::string(who) + ": unknown model");
    if (dim < 2 || dim % 2 != 0 || dim > 2048)
        return fail(std::string(who) + ": dim must be even and at most 2048");
    if (model == GV_KG_QUATE && dim % 4 != 0)
        return fail(std::string(who) + ": QuatE needs a dimension divisible by 4");
    E = floats_per_thread(dim, model, train);
    if (dim % E != 0)
        return fail(std::string(who) + ": dim must be a multiple of " + std::to_string(E));
    group_threads = (dim / E + 31) / 32 * 32;
    if (group_threads > (train && E != 8 ? kMaxGroupThreads : kCtaThreads))
        return fail(std::string(who) + ": dim too large for one CTA");
    return 0;
}

}  // namespace

}  // namespace device

int kg_kernel_flags() {
    return device::kg_flags();
}
void set_kg_kernel_flags(int value) {
    device::kg_flags() = value;
}

}  // namespace gv

using namespace gv;
using namespace gv::device;

extern "C" {

int gv_cuda_kg_train_block(const gv_kg_matrices_t *m, int model, const uint32_t *batch, uint64_t num_sample,
                           int num_negative, const uint32_t *negatives, const double *random, uint32_t negative_count,
                           uint32_t *negatives_out, const gv_device_optimizer_t *optimizer, const float *lr_per_batch,
                           uint32_t batch_size, float relation_lr_multiplier, float margin_or_l3,
                           float adversarial_temperature, float *loss_per_sample, float *loss_per_batch, int num_group,
                           void *stream) {
    if (num_sample == 0)
        return 0;
    if (!m || !batch || !optimizer || !lr_per_batch || !m->head || !m->tail || !m->relation || batch_size == 0)
        return fail("gv_cuda_kg_train_block: null argument");
    if (num_negative < 0 || (num_negative > 0 && !negatives && (!random || negative_count == 0)))
        return fail("gv_cuda_kg_train_block: negatives need either ids or a random stream and a count");
    int E, group_threads;
    if (check_geometry("gv_cuda_kg_train_block", m->dim, model, E, group_threads, true))
        return -1;
    const int type = optimizer->type;
    const int num_moment = type == GV_OPT_SGD ? 0 : (type == GV_OPT_ADAM ? 2 : 1);
    if (type < GV_OPT_SGD || type > GV_OPT_ADAM)
        return fail("gv_cuda_kg_train_block: unknown optimizer");
    if ((num_moment >= 1 && (!m->head_m1 || !m->tail_m1 || !m->relation_m1)) ||
        (num_moment >= 2 && (!m->head_m2 || !m->tail_m2 || !m->relation_m2)))
        return fail("gv_cuda_kg_train_block: moment matrices missing for this optimizer");
    KgParams p;
    p.dim = m->dim;
    p.num_head = m->num_head;
    p.shared = m->head == m->tail;
    p.head = m->head, p.tail = m->tail, p.relation = m->relation;
    p.head_m1 = m->head_m1, p.tail_m1 = m->tail_m1, p.relation_m1 = m->relation_m1;
    p.head_m2 = m->head_m2, p.tail_m2 = m->tail_m2, p.relation_m2 = m->relation_m2;
    p.batch = batch;
    p.num_sample = num_sample;
    p.num_negative = num_negative;
    p.negatives = negatives;
    p.random = random;
    p.negative_count = negative_count;
    p.negatives_out = negatives_out;
    p.optimizer = *optimizer;
    p.lr_per_batch = lr_per_batch;
    p.batch_size = batch_size;
    p.relation_lr_multiplier = relation_lr_multiplier;
    p.margin_or_l3 = margin_or_l3;
    p.temperature = adversarial_temperature;
    p.loss_per_sample = loss_per_sample;
    p.loss_per_batch = loss_per_batch;
    p.flags = kg_flags();

    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int device = 0, num_sm = 0;
    GV_CUDA_OK(cudaGetDevice(&device));
    GV_CUDA_OK(cudaDeviceGetAttribute(&num_sm, cudaDevAttrMultiProcessorCount, device));
    // num_group == 1: one group, samples in order (the parity tests); 0: fill the device
    const int groups_per_cta = num_group == 1 ? 1 : std::max(1, kCtaThreads / group_threads);
    const dim3 block(groups_per_cta * group_threads);
    const size_t per_group =
        size_t(2) * kPass1Batch * (group_threads / 32) * sizeof(float) + size_t(num_negative) * sizeof(uint32_t);
    const size_t shared = per_group * groups_per_cta;
    if (shared > 48 * 1024)
        return fail("gv_cuda_kg_train_block: too many negatives per sample for the shared-memory id buffer");
    unsigned long long ctas = (num_sample + groups_per_cta - 1) / groups_per_cta;
    if (num_group == 1)
        ctas = 1;
    else if (num_group > 1)
        ctas = std::min<unsigned long long>(ctas, (unsigned long long)(num_group + groups_per_cta - 1) / groups_per_cta);
    else
        ctas = std::min<unsigned long long>(ctas, (unsigned long long)num_sm * 2);
    const dim3 grid((unsigned)ctas);
    cudaError_t status;
    if (E == 8)
        status = launch_train<8>(p, model, num_moment, grid, block, shared, s);
    else if (E == 4)
        status = launch_train<4>(p, model, num_moment, grid, block, shared, s);
    else
        status = launch_train<2>(p, model, num_moment, grid, block, shared, s);
    GV_CUDA_OK(status);
    return 0;
}

int gv_cuda_kg_predict(const gv_kg_matrices_t *m, int model, const uint32_t *batch, uint64_t num_sample, float margin,
                       float *logits, void *stream) {
    if (num_sample == 0)
        return 0;
    if (!m || !batch || !logits || !m->head || !m->tail || !m->relation)
        return fail("gv_cuda_kg_predict: null argument");
    int E, group_threads;
    if (check_geometry("gv_cuda_kg_predict", m->dim, model, E, group_threads))
        return -1;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int device = 0, num_sm = 0;
    GV_CUDA_OK(cudaGetDevice(&device));
    GV_CUDA_OK(cudaDeviceGetAttribute(&num_sm, cudaDevAttrMultiProcessorCount, device));
    const int groups_per_cta = kCtaThreads / group_threads;
    const dim3 block(groups_per_cta * group_threads);
    const size_t shared = size_t(2) * kPass1Batch * (group_threads / 32) * sizeof(float) * groups_per_cta;
    const unsigned long long ctas =
        std::min<unsigned long long>((num_sample + groups_per_cta - 1) / groups_per_cta, (unsigned long long)num_sm * 8);
    const dim3 grid((unsigned)ctas);
    cudaError_t status;
    if (E == 8)
        status = launch_predict<8>(model, m->head, m->tail, m->relation, m->dim, batch, num_sample, margin, logits, grid,
                                   block, shared, s);
    else if (E == 4)
        status = launch_predict<4>(model, m->head, m->tail, m->relation, m->dim, batch, num_sample, margin, logits, grid,
                                   block, shared, s);
    else
        status = launch_predict<2>(model, m->head, m->tail, m->relation, m->dim, batch, num_sample, margin, logits, grid,
                                   block, shared, s);
    GV_CUDA_OK(status);
    return 0;
}

}  // extern "C"
