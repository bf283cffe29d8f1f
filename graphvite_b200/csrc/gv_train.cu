// =============================================================================
// gv_train.cu -- the negative-sampling SGD hot loop, hand-written for sm_100a.
//
// Replaces gpu::graph::train / train_1_moment / train_2_moment / predict
// (reference include/instance/gpu/graph.cuh:36-279), LINE::forward/backward
// (include/instance/model/graph.h:40-85), the optimizer update rules
// (include/core/optimizer.h:161-210) and gpu::Sample
// (include/base/alias_table.cuh:175-183).
//
// Design:
//  * one warp owns one positive sample and its k negatives at a time; the d-dim
//    row is spread over the lanes as float4 (128-bit LDG/STG; a 128-d fp32 row is
//    exactly one 512-B warp-wide transaction), the vertex row lives in registers
//    for the whole sample, dot products are butterfly warp-shuffle reductions and
//    the Hogwild update is written back in place;
//  * cache policy (TrainParams::flags / hot_rows): rows are read either L2-only
//    (ld.global.cg) or through L1 (ld.global.ca, what the reference's plain loads compile to);
//    the L1 path is incoherent across SMs exactly like the reference's, and how much of it is
//    used decides how the Hogwild races fall -- the shipped setting is the one that passed the
//    statistical parity test against the unmodified reference (tests/test_gpu_zzzzz_parity.py);
//  * a warp stages 32 samples at a time: lane i loads sample i's {tail, head} pair, draws its k
//    negatives from the alias table (fused gpu::Sample) and parks the ids in shared memory, so
//    index latency is paid once per 32 samples and every row address is known up front.  The 32
//    samples are either consecutive pool entries or (flags & 16) entries one grid-width apart,
//    which is the reference's concurrency structure: neighbouring pool entries are trained by
//    different warps at the same time, a warp's own samples lie far apart in the pool;
//  * persistent grid: one launch consumes a whole pool block (any number of reference
//    batches); the per-batch learning rate comes from a small array.
// =============================================================================
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <map>
#include <mutex>
#include <string>
#include <utility>

#include "gv_common.h"
#include "gv_device.cuh"

namespace gv {
namespace device {

constexpr unsigned kFullMask = 0xFFFFFFFFu;
constexpr int kBlockThreads = 256;
constexpr float kEpsilon = 1e-15f;  // util/common.h:28

struct TrainParams {
    float *vertex, *context, *vertex_m1, *context_m1, *vertex_m2, *context_m2;
    const uint2 *pool;
    unsigned long long num_sample;
    int num_negative;
    const uint32_t *negatives;
    const double *random;
    const gv_alias_entry_t *negative_table;
    uint32_t negative_count;
    uint32_t *negatives_out;
    gv_device_optimizer_t optimizer;
    const float *lr_per_batch;
    uint32_t batch_size;
    float negative_weight;
    float *loss_per_sample, *loss_per_batch;
    int flags;  // 1 = L1-cached (.ca) loads for ALL rows, 4 = never use train_sgd_kernel,
                // 8 = warps claim their 32-sample chunks from a counter instead of a fixed stride,
                // 16 = interleaved mapping: warp w trains samples w, w + G, w + 2G ... (G = warps of the grid), so
                //      that neighbouring pool entries run concurrently on different warps like the reference's
                //      one-warp-per-sample grid (instance/gpu/graph.cuh:54-60); ignored with 8,
                // 64 = SGD with the persistent kernels below (32 consecutive pool entries per warp, software
                //      pipelining; faster, but a walk's samples no longer race the way the reference's do --
                //      see train_sample_per_warp_kernel); flags 1 / 8 / 16 / 32 and hot_rows only act on those,
                // 32 = rows are stored with the default write-back policy (st.global, what the reference's stores
                //      compile to: the writing SM's L1 copy stays current) instead of st.global.cg
    unsigned int *work_counter;  // flags & 8: {next ticket, warps done}; both zero between launches
    uint32_t hot_rows;  // rows with a local id below this are loaded through L1 (.ca), the rest L2-only (.cg)
    uint32_t prefetch_blocks;  // one-warp-per-sample kernel: index lines of the block this far ahead are pulled into L2
};

// Which 32-sample chunk a warp takes after `chunk`.  Static: the grid-stride successor.  Dynamic (work_counter):
// the first num_warp chunks go by warp index, every further one by ticket -- a CTA that starts late (its SM was
// busy with a sampler kernel when the launch began) or sits on hot rows takes fewer chunks instead of becoming the
// straggler of the launch.  One warp alone still visits the chunks in order (the parity tests' mode).
__device__ __forceinline__ unsigned long long next_chunk(const TrainParams &p, unsigned long long chunk,
                                                         unsigned long long num_warp, int lane) {
    if (!p.work_counter)
        return chunk + num_warp;
    unsigned int ticket = 0;
    if (lane == 0)
        ticket = atomicAdd(p.work_counter, 1u);
    ticket = __shfl_sync(0xFFFFFFFFu, ticket, 0);
    return num_warp + ticket;
}

// the last warp to leave a launch re-arms the counter for the next one (launches of a stream are serial)
__device__ __forceinline__ void release_work_counter(const TrainParams &p, unsigned long long num_warp, int lane) {
    if (p.work_counter && lane == 0) {
        const unsigned int done = atomicAdd(p.work_counter + 1, 1u);
        if (done + 1 == num_warp) {
            p.work_counter[0] = 0;
            p.work_counter[1] = 0;
        }
    }
}

// -----------------------------------------------------------------------------
// A d-dim fp32 row spread over a warp: pass p, lane l holds elements
// [(p*32+l)*4, +4).  DIM must be a multiple of 4; lanes past the end hold zeros.
// -----------------------------------------------------------------------------
template<int DIM>
struct Row {
    static constexpr int kPass = (DIM + 127) / 128;
    float4 x[kPass];
};

template<int DIM>
__device__ __forceinline__ bool lane_active(int pass, int lane) {
    return (DIM % 128 == 0) || ((pass * 32 + lane) * 4 < DIM);
}

template<int DIM>
__device__ __forceinline__ void load_row(Row<DIM> &row, const float *base, int lane, bool l1 = false) {
#pragma unroll
    for (int p = 0; p < Row<DIM>::kPass; p++) {
        if (lane_active<DIM>(p, lane))
            row.x[p] = l1 ? __ldca(reinterpret_cast<const float4 *>(base) + p * 32 + lane)
                          : __ldcg(reinterpret_cast<const float4 *>(base) + p * 32 + lane);
        else
            row.x[p] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// a plain ld.global / st.global per lane: exactly the reference's memory policy (L1-cached loads, write-back stores;
// SASS LDG.E.128 / STG.E.128 where the reference has LDG.E / STG.E)
template<int DIM>
__device__ __forceinline__ void load_row_plain(Row<DIM> &row, const float *base, int lane) {
#pragma unroll
    for (int p = 0; p < Row<DIM>::kPass; p++) {
        if (lane_active<DIM>(p, lane))
            row.x[p] = reinterpret_cast<const float4 *>(base)[p * 32 + lane];
        else
            row.x[p] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

template<int DIM>
__device__ __forceinline__ void store_row(const Row<DIM> &row, float *base, int lane, bool write_back = false) {
#pragma unroll
    for (int p = 0; p < Row<DIM>::kPass; p++)
        if (lane_active<DIM>(p, lane)) {
            if (write_back)
                reinterpret_cast<float4 *>(base)[p * 32 + lane] = row.x[p];
            else
                __stcg(reinterpret_cast<float4 *>(base) + p * 32 + lane, row.x[p]);
        }
}

// Which samples a warp stages in its m-th visit.  Consecutive: chunk c covers pool entries [32c, 32c + 32).
// Interleaved (flags & 16): the grid's G warps sweep the pool together -- in visit m warp w takes the entries
// (32m + l) * G + w for l = 0..31, so at any moment the samples in flight form a window of about G consecutive
// pool entries, one per warp.  One warp alone (G = 1, the parity tests' mode) visits the pool in order either way.
struct ChunkMap {
    unsigned long long first, step;
    __device__ __forceinline__ unsigned long long sample(int lane) const { return first + step * lane; }
};
__device__ __forceinline__ ChunkMap map_chunk(bool interleaved, unsigned long long chunk, unsigned long long num_warp) {
    if (!interleaved)
        return {chunk * 32, 1};
    return {(chunk / num_warp * 32) * num_warp + chunk % num_warp, num_warp};
}

__device__ __forceinline__ float warp_sum(float value) {
#pragma unroll
    for (int delta = 16; delta > 0; delta >>= 1)
        value += __shfl_xor_sync(kFullMask, value, delta);
    return value;
}

template<int DIM>
__device__ __forceinline__ float dot(const Row<DIM> &a, const Row<DIM> &b) {
    float acc = 0.f;
#pragma unroll
    for (int p = 0; p < Row<DIM>::kPass; p++) {
        acc = fmaf(a.x[p].x, b.x[p].x, acc);
        acc = fmaf(a.x[p].y, b.x[p].y, acc);
        acc = fmaf(a.x[p].z, b.x[p].z, acc);
        acc = fmaf(a.x[p].w, b.x[p].w, acc);
    }
    return warp_sum(acc);
}

// util/math.h:30-33: sigmoid(x) = x > 0 ? 1 / (1 + exp(-x)) : exp(x) / (exp(x) + 1).
// Both branches share e = exp(-|x|) and r = 1 / (1 + e).  The training path uses the hardware
// exponential and reciprocal (ex2.approx / rcp.approx, <= 2 ulp each): the kernel is issue-bound on
// the precise libdevice sequences otherwise, and the difference (~1e-7 relative on the gradient) is
// three orders of magnitude below the Hogwild noise of the algorithm itself.
__device__ __forceinline__ float sigmoid(float x) {
    const float e = gv_fast_exp(-fabsf(x));
    const float r = gv_fast_divide(1.f, 1.f + e);
    return x > 0 ? r : e * r;
}

// -----------------------------------------------------------------------------
// core/optimizer.h:161-210 -- returns the step to subtract from `parameter`
// -----------------------------------------------------------------------------
template<int OPT>
__device__ __forceinline__ float update(const gv_device_optimizer_t &o, float lr, float parameter, float gradient,
                                        float &moment1, float &moment2, float weight) {
    float regularized = weight * (gradient + o.weight_decay * parameter);
    if (OPT == GV_OPT_MOMENTUM) {
        moment1 = o.a * moment1 + (1 - o.a) * regularized;
        return lr * moment1;
    }
    if (OPT == GV_OPT_ADAGRAD) {
        moment1 += regularized * regularized;
        return lr * regularized / (sqrtf(moment1) + o.epsilon);
    }
    if (OPT == GV_OPT_RMSPROP) {
        moment1 = o.a * moment1 + (1 - o.a) * regularized * regularized;
        return lr * regularized / sqrtf(moment1 + o.epsilon);
    }
    if (OPT == GV_OPT_ADAM) {
        moment1 = o.a * moment1 + (1 - o.a) * regularized;
        moment2 = o.b * moment2 + (1 - o.b) * regularized * regularized;
        return lr * moment1 / (sqrtf(moment2) + o.epsilon);
    }
    return lr * regularized;
}

// LINE::backward, instance/model/graph.h:47-85: both updates read the pre-update v and c.
// SGD (optimizer.h:161-164): x -= lr * w * (g * y + wd * x)  ==  x * (1 - lr*w*wd) - (lr*w*g) * y,
// two instructions per element instead of five.
template<int DIM, int OPT>
__device__ __forceinline__ void backward(const gv_device_optimizer_t &o, float lr, float gradient, float weight,
                                         Row<DIM> &v, Row<DIM> &c, Row<DIM> &vm1, Row<DIM> &cm1, Row<DIM> &vm2,
                                         Row<DIM> &cm2) {
    if constexpr (OPT == GV_OPT_SGD) {
        const float scale = lr * weight;
        const float alpha = 1.f - scale * o.weight_decay, beta = scale * gradient;
#pragma unroll
        for (int p = 0; p < Row<DIM>::kPass; p++) {
#define GV_ELEMENT(f)                                      \
    {                                                      \
        const float vv = v.x[p].f, cc = c.x[p].f;          \
        v.x[p].f = fmaf(-beta, cc, alpha * vv);            \
        c.x[p].f = fmaf(-beta, vv, alpha * cc);            \
    }
            GV_ELEMENT(x)
            GV_ELEMENT(y)
            GV_ELEMENT(z)
            GV_ELEMENT(w)
#undef GV_ELEMENT
        }
    } else {
#pragma unroll
    for (int p = 0; p < Row<DIM>::kPass; p++) {
#define GV_ELEMENT(f)                                                                        \
    {                                                                                        \
        float vv = v.x[p].f, cc = c.x[p].f;                                                  \
        v.x[p].f = vv - update<OPT>(o, lr, vv, gradient * cc, vm1.x[p].f, vm2.x[p].f, weight); \
        c.x[p].f = cc - update<OPT>(o, lr, cc, gradient * vv, cm1.x[p].f, cm2.x[p].f, weight); \
    }
        GV_ELEMENT(x)
        GV_ELEMENT(y)
        GV_ELEMENT(z)
        GV_ELEMENT(w)
#undef GV_ELEMENT
    }
    }
}

// AliasTable::sample as called by gpu::Sample (base/alias_table.cuh:148-152,175-183):
// both randoms are narrowed to float first; index = Index(double(rand1) * count).
// cuRAND doubles lie in (0,1]: rand1 == 1 would index one past the table in the
// reference; we clamp to count-1 and change no other outcome.
__device__ __forceinline__ uint32_t alias_sample_narrowed(const gv_alias_entry_t *table, uint32_t count, double random1,
                                                          double random2) {
    float rand1 = float(random1), rand2 = float(random2);
    uint32_t index = uint32_t(double(rand1) * double(count));
    index = min(index, count - 1);
    const uint2 entry = __ldg(reinterpret_cast<const uint2 *>(table) + index);
    return rand2 < __uint_as_float(entry.x) ? index : entry.y;
}

// -----------------------------------------------------------------------------
// The train kernel.  NM moment rows accompany every embedding row.  LOSS = false skips the two
// logf per target (the reference logs the loss of one batch in log_frequency, core/solver.h:1541).
//
// Memory-level parallelism: all rows of a sample are requested before the first one is used (the
// next target's context row is prefetched while the current one is processed), and for SGD the
// first two rows of the NEXT sample are requested before the current sample is computed.  Rows a
// sample is about to overwrite are never consumed stale by the same warp: equal ids reuse the
// updated registers, and a prefetched row that the current sample turned out to write is reloaded.
// A single warp therefore executes exactly the sequential semantics of the reference's per-sample
// loop, which is what the parity tests check; across warps the updates race (Hogwild), as in the
// reference.
// -----------------------------------------------------------------------------
template<int DIM, int OPT, bool LOSS>
__global__ void __launch_bounds__(kBlockThreads) train_kernel(const TrainParams p) {
    constexpr int NM = OPT == GV_OPT_SGD ? 0 : (OPT == GV_OPT_ADAM ? 2 : 1);
    constexpr bool kCross = NM == 0 && Row<DIM>::kPass <= 2;  // cross-sample prefetch (register budget)
    GV_DYNAMIC_SHARED(uint32_t, shared_ids);

    const int lane = threadIdx.x & 31;
    const int warp_in_block = threadIdx.x >> 5;
    const int k = p.num_negative;
    const int stride = k + 2;  // head, k negatives, positive tail
    uint32_t *ids = shared_ids + warp_in_block * 32 * stride;

    const unsigned long long num_chunk = (p.num_sample + 31) / 32;
    const unsigned long long num_warp = (unsigned long long)gridDim.x * (blockDim.x >> 5);
    const gv_device_optimizer_t o = p.optimizer;
    const bool l1 = p.flags & 1, wb = p.flags & 32;
    const bool interleaved = (p.flags & 16) && !p.work_counter;

    // interleaved: every warp owns pool entries however few there are; its loop ends at the `break` below
    const unsigned long long chunk_limit = interleaved ? ~0ull : num_chunk;
    for (unsigned long long chunk = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + warp_in_block;
         chunk < chunk_limit; chunk = next_chunk(p, chunk, num_warp, lane)) {
        const ChunkMap map = map_chunk(interleaved, chunk, num_warp);
        if (map.first >= p.num_sample)
            break;  // interleaved: this warp's share of the pool is exhausted (later visits start even further)
        const unsigned long long i = map.sample(lane);
        const bool valid = i < p.num_sample;
        float lr_lane = 0.f;
        uint32_t batch_lane = 0;
        if (valid) {
            // {tail, head}; consecutive: streamed once, interleaved: the CTA's warps share the 32-byte sectors
            const uint2 pair = interleaved ? __ldg(p.pool + i) : __ldcs(p.pool + i);
            ids[lane * stride] = pair.y;
            ids[lane * stride + 1 + k] = pair.x;
            for (int s = 0; s < k; s++) {
                const unsigned long long t = i * k + s;
                uint32_t negative;
                if (p.negatives)
                    negative = __ldcs(p.negatives + t);
                else {
                    const double2 r = __ldcs(reinterpret_cast<const double2 *>(p.random) + t);
                    negative = alias_sample_narrowed(p.negative_table, p.negative_count, r.x, r.y);
                }
                ids[lane * stride + 1 + s] = negative;
                if (p.negatives_out)
                    p.negatives_out[t] = negative;
            }
            batch_lane = uint32_t(i / p.batch_size);
            lr_lane = __ldg(p.lr_per_batch + batch_lane);
        }
        __syncwarp();

        const int count = __popc(__ballot_sync(kFullMask, valid));  // valid lanes are a prefix
        float loss_lane = 0.f;
        Row<DIM> v, vm1, vm2, c, cm1, cm2, c_next, cm1_next, cm2_next, v_ahead, c_ahead;
        uint32_t head = ids[0], tail = ids[1];
        if (kCross) {
            load_row<DIM>(v_ahead, p.vertex + size_t(head) * DIM, lane, l1 || head < p.hot_rows);
            load_row<DIM>(c_ahead, p.context + size_t(tail) * DIM, lane, l1 || tail < p.hot_rows);
        }
        for (int t = 0; t < count; t++) {
            const uint32_t *sample = ids + t * stride;
            const float lr = __shfl_sync(kFullMask, lr_lane, t);
            const size_t head_offset = size_t(head) * DIM;
            // ---- this sample's first rows: already in flight (SGD) or requested now ----
            uint32_t head_ahead = head, tail_ahead = tail;
            const bool more = t + 1 < count;
            if (kCross) {
                v = v_ahead;
                c = c_ahead;
                if (more) {  // request the next sample's first two rows before computing this one
                    head_ahead = sample[stride];
                    tail_ahead = sample[stride + 1];
                    if (head_ahead != head)
                        load_row<DIM>(v_ahead, p.vertex + size_t(head_ahead) * DIM, lane, l1 || head_ahead < p.hot_rows);
                    load_row<DIM>(c_ahead, p.context + size_t(tail_ahead) * DIM, lane, l1 || tail_ahead < p.hot_rows);
                }
            } else {
                load_row<DIM>(v, p.vertex + head_offset, lane, l1 || head < p.hot_rows);
                load_row<DIM>(c, p.context + size_t(tail) * DIM, lane, l1 || tail < p.hot_rows);
                if (NM >= 1) {
                    load_row<DIM>(vm1, p.vertex_m1 + head_offset, lane, head < p.hot_rows);
                    load_row<DIM>(cm1, p.context_m1 + size_t(tail) * DIM, lane, tail < p.hot_rows);
                }
                if (NM >= 2) {
                    load_row<DIM>(vm2, p.vertex_m2 + head_offset, lane, head < p.hot_rows);
                    load_row<DIM>(cm2, p.context_m2 + size_t(tail) * DIM, lane, tail < p.hot_rows);
                }
                if (more) {
                    head_ahead = sample[stride];
                    tail_ahead = sample[stride + 1];
                }
            }
            bool stale_ahead = false;  // did this sample write the context row prefetched for the next one?
            float sample_loss = 0.f;
            for (int s = 0; s <= k; s++) {
                // prefetch the next target's rows while this one is being processed
                uint32_t tail_next = tail;
                if (s < k) {
                    tail_next = sample[2 + s];
                    if (tail_next != tail) {
                        load_row<DIM>(c_next, p.context + size_t(tail_next) * DIM, lane, l1 || tail_next < p.hot_rows);
                        if (NM >= 1)
                            load_row<DIM>(cm1_next, p.context_m1 + size_t(tail_next) * DIM, lane, tail_next < p.hot_rows);
                        if (NM >= 2)
                            load_row<DIM>(cm2_next, p.context_m2 + size_t(tail_next) * DIM, lane, tail_next < p.hot_rows);
                    }
                }
                // forward (LINE::forward) and the gradient of the logistic loss, gpu/graph.cuh:73-88
                const float logit = dot<DIM>(v, c);
                const float prob = sigmoid(logit);
                float gradient, weight;
                if (s == k) {
                    gradient = prob - 1;
                    weight = 1;
                    if (LOSS)
                        sample_loss += weight * -logf(prob + kEpsilon);
                } else {
                    gradient = prob;
                    weight = p.negative_weight;
                    if (LOSS)
                        sample_loss += weight * -logf(1 - prob + kEpsilon);
                }
                backward<DIM, OPT>(o, lr, gradient, weight, v, c, vm1, cm1, vm2, cm2);
                store_row<DIM>(c, p.context + size_t(tail) * DIM, lane, wb);
                if (NM >= 1)
                    store_row<DIM>(cm1, p.context_m1 + size_t(tail) * DIM, lane, wb);
                if (NM >= 2)
                    store_row<DIM>(cm2, p.context_m2 + size_t(tail) * DIM, lane, wb);
                stale_ahead |= tail == tail_ahead;
                if (s < k && tail_next != tail) {
                    c = c_next;
                    if (NM >= 1)
                        cm1 = cm1_next;
                    if (NM >= 2)
                        cm2 = cm2_next;
                }  // else: the same row again -- keep the just-updated registers
                tail = tail_next;
            }
            store_row<DIM>(v, p.vertex + head_offset, lane, wb);
            if (NM >= 1)
                store_row<DIM>(vm1, p.vertex_m1 + head_offset, lane, wb);
            if (NM >= 2)
                store_row<DIM>(vm2, p.vertex_m2 + head_offset, lane, wb);
            if (LOSS) {
                sample_loss = sample_loss / (1 + k * p.negative_weight);  // gpu/graph.cuh:91-92
                if (lane == t)
                    loss_lane = sample_loss;
            }
            if (kCross && more) {
                if (head_ahead == head)
                    v_ahead = v;  // same vertex again: continue from the updated registers
                if (stale_ahead)  // program order: this load follows the store above in the same thread
                    load_row<DIM>(c_ahead, p.context + size_t(tail_ahead) * DIM, lane, l1 || tail_ahead < p.hot_rows);
            }
            head = head_ahead;
            tail = tail_ahead;
        }
        if (LOSS) {
            if (p.loss_per_sample && valid)
                p.loss_per_sample[i] = loss_lane;
            if (p.loss_per_batch) {
                // a 32-sample chunk may straddle batches: one reduction + atomic per distinct batch
                unsigned remaining = __ballot_sync(kFullMask, valid);
                while (remaining) {
                    const int leader = __ffs(remaining) - 1;
                    const uint32_t batch = __shfl_sync(kFullMask, batch_lane, leader);
                    const bool mine = valid && batch_lane == batch;
                    const float sum = warp_sum(mine ? loss_lane : 0.f);
                    if (lane == leader)
                        atomicAdd(p.loss_per_batch + batch, sum);
                    remaining &= ~__ballot_sync(kFullMask, mine);
                }
            }
        }
        __syncwarp();  // ids[] is rewritten by the next chunk
    }
    release_work_counter(p, num_warp, lane);
}

// -----------------------------------------------------------------------------
// SGD with a compile-time number of negatives (the shipped configs all use k = 1): the hot
// kernel.  Same staging and the same per-sample semantics as train_kernel above, but ALL rows
// of sample t+1 (vertex + K+1 context rows) are requested before sample t is computed, so a warp
// keeps 2 * (K + 2) rows in flight and the load latency of a sample hides behind the whole
// previous sample.  Rows that the current sample writes and the next one reads (same vertex, or
// a context row in both) are forwarded register-to-register instead of being reloaded, which
// keeps a single warp exactly sequential.
// -----------------------------------------------------------------------------
template<int DIM, int K>
struct SampleRows {
    Row<DIM> v;
    Row<DIM> c[K + 1];
    uint32_t head;
    uint32_t tail[K + 1];
};

template<int DIM, int K>
__device__ __forceinline__ void request_sample(SampleRows<DIM, K> &rows, const uint32_t *sample, const TrainParams &p,
                                               int lane, bool l1) {
    rows.head = sample[0];
#pragma unroll
    for (int s = 0; s <= K; s++)
        rows.tail[s] = sample[1 + s];
    load_row<DIM>(rows.v, p.vertex + size_t(rows.head) * DIM, lane, l1 || rows.head < p.hot_rows);
#pragma unroll
    for (int s = 0; s <= K; s++)
        load_row<DIM>(rows.c[s], p.context + size_t(rows.tail[s]) * DIM, lane, l1 || rows.tail[s] < p.hot_rows);
}

template<int DIM, int K, bool LOSS>
__device__ __forceinline__ float process_sample(SampleRows<DIM, K> &cur, SampleRows<DIM, K> &next, bool has_next,
                                                float lr, const TrainParams &p, int lane) {
    float sample_loss = 0.f;
    Row<DIM> unused;
#pragma unroll
    for (int s = 0; s <= K; s++) {
        // a row that an earlier target of this sample already updated: continue from those registers
#pragma unroll
        for (int e = 0; e < s; e++)
            if (cur.tail[e] == cur.tail[s])
                cur.c[s] = cur.c[e];
        const float prob = sigmoid(dot<DIM>(cur.v, cur.c[s]));
        float gradient, weight;
        if (s == K) {
            gradient = prob - 1;
            weight = 1;
            if (LOSS)
                sample_loss += weight * -logf(prob + kEpsilon);
        } else {
            gradient = prob;
            weight = p.negative_weight;
            if (LOSS)
                sample_loss += weight * -logf(1 - prob + kEpsilon);
        }
        backward<DIM, GV_OPT_SGD>(p.optimizer, lr, gradient, weight, cur.v, cur.c[s], unused, unused, unused, unused);
        store_row<DIM>(cur.c[s], p.context + size_t(cur.tail[s]) * DIM, lane, p.flags & 32);
    }
    // consecutive samples of one walk often share the head (DeepWalk / node2vec: k = 1..augmentation_step
    // with shuffle_base 1): the row is carried in registers and written once at the end of the run
    const bool same_head = has_next && next.head == cur.head;
    if (!same_head)
        store_row<DIM>(cur.v, p.vertex + size_t(cur.head) * DIM, lane, p.flags & 32);
    if (has_next) {  // forward what the next sample requested before these stores were issued
        if (same_head)
            next.v = cur.v;
#pragma unroll
        for (int s = 0; s <= K; s++)
#pragma unroll
            for (int e = 0; e <= K; e++)
                if (next.tail[s] == cur.tail[e])
                    next.c[s] = cur.c[e];
    }
    return sample_loss / (1 + K * p.negative_weight);  // gpu/graph.cuh:91-92
}

template<int DIM, int K, bool LOSS>
__global__ void __launch_bounds__(kBlockThreads) train_sgd_kernel(const TrainParams p) {
    GV_DYNAMIC_SHARED(uint32_t, shared_ids);
    const int lane = threadIdx.x & 31;
    const int warp_in_block = threadIdx.x >> 5;
    constexpr int stride = K + 2;  // head, K negatives, positive tail
    uint32_t *ids = shared_ids + warp_in_block * 32 * stride;
    const unsigned long long num_chunk = (p.num_sample + 31) / 32;
    const unsigned long long num_warp = (unsigned long long)gridDim.x * (blockDim.x >> 5);
    const bool l1 = p.flags & 1;
    const bool interleaved = (p.flags & 16) && !p.work_counter;

    // interleaved: every warp owns pool entries however few there are; its loop ends at the `break` below
    const unsigned long long chunk_limit = interleaved ? ~0ull : num_chunk;
    for (unsigned long long chunk = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + warp_in_block;
         chunk < chunk_limit; chunk = next_chunk(p, chunk, num_warp, lane)) {
        const ChunkMap map = map_chunk(interleaved, chunk, num_warp);
        if (map.first >= p.num_sample)
            break;
        const unsigned long long i = map.sample(lane);
        const bool valid = i < p.num_sample;
        float lr_lane = 0.f;
        uint32_t batch_lane = 0;
        if (valid) {
            const uint2 pair = interleaved ? __ldg(p.pool + i) : __ldcs(p.pool + i);  // {tail, head}
            ids[lane * stride] = pair.y;
            ids[lane * stride + 1 + K] = pair.x;
#pragma unroll
            for (int s = 0; s < K; s++) {
                const unsigned long long t = i * K + s;
                uint32_t negative;
                if (p.negatives)
                    negative = __ldcs(p.negatives + t);
                else {
                    const double2 r = __ldcs(reinterpret_cast<const double2 *>(p.random) + t);
                    negative = alias_sample_narrowed(p.negative_table, p.negative_count, r.x, r.y);
                }
                ids[lane * stride + 1 + s] = negative;
                if (p.negatives_out)
                    p.negatives_out[t] = negative;
            }
            batch_lane = uint32_t(i / p.batch_size);
            lr_lane = __ldg(p.lr_per_batch + batch_lane);
        }
        __syncwarp();

        const int count = __popc(__ballot_sync(kFullMask, valid));
        float loss_lane = 0.f;
        SampleRows<DIM, K> a, b;
        request_sample<DIM, K>(a, ids, p, lane, l1);
        for (int t = 0; t < count; t += 2) {
            bool more = t + 1 < count;
            if (more)
                request_sample<DIM, K>(b, ids + (t + 1) * stride, p, lane, l1);
            float loss = process_sample<DIM, K, LOSS>(a, b, more, __shfl_sync(kFullMask, lr_lane, t), p, lane);
            if (LOSS && lane == t)
                loss_lane = loss;
            if (!more)
                break;
            more = t + 2 < count;
            if (more)
                request_sample<DIM, K>(a, ids + (t + 2) * stride, p, lane, l1);
            loss = process_sample<DIM, K, LOSS>(b, a, more, __shfl_sync(kFullMask, lr_lane, t + 1), p, lane);
            if (LOSS && lane == t + 1)
                loss_lane = loss;
        }
        if (LOSS) {
            if (p.loss_per_sample && valid)
                p.loss_per_sample[i] = loss_lane;
            if (p.loss_per_batch) {
                unsigned remaining = __ballot_sync(kFullMask, valid);
                while (remaining) {
                    const int leader = __ffs(remaining) - 1;
                    const uint32_t batch = __shfl_sync(kFullMask, batch_lane, leader);
                    const bool mine = valid && batch_lane == batch;
                    const float sum = warp_sum(mine ? loss_lane : 0.f);
                    if (lane == leader)
                        atomicAdd(p.loss_per_batch + batch, sum);
                    remaining &= ~__ballot_sync(kFullMask, mine);
                }
            }
        }
        __syncwarp();  // ids[] is rewritten by the next chunk
    }
    release_work_counter(p, num_warp, lane);
}


// -----------------------------------------------------------------------------
// The shipped SGD kernel: the reference's own launch geometry, because the result of Hogwild
// training is a property of WHICH samples race, not only of the arithmetic.
//
// gpu::graph::train (instance/gpu/graph.cuh:36-95, launched <<<8192, 512>>> once per batch,
// instance/graph.cuh:487) gives every sample of a batch its own warp: 16 consecutive pool entries per
// 512-thread block, blocks dispatched in order, 4 blocks = 64 warps resident per SM.  At any moment
// the samples in flight are a window of ~64 * #SM CONSECUTIVE pool entries, each read-modify-writing
// its rows without atomics.  Consecutive pool entries come from the same random walk (pseudo shuffle,
// instance/graph.cuh:427-447), a walk revisits a vertex two steps later with probability 1 / <degree>,
// and of two concurrent updates of one row one is lost: on the Youtube-shaped graph the reference
// loses ~10 % of the updates this way, and a kernel that trains a walk's samples one after another
// (the persistent kernels above: 32 consecutive entries per warp) keeps them all -- measured on a
// B200, 100 epochs: |vertex| +9.1 %, |context| -9.7 % against the unmodified reference however the
// cache policy was set (profiles/r02_parity_sweep.md).  Within a row the arithmetic is identical.
//
// So this kernel keeps the reference's concurrency structure and memory policy and only changes how a
// warp moves its bytes:
//  * one warp = one sample, 16 warps per block, blocks in pool order, 64 warps per SM
//    (__launch_bounds__(512, 4): 32 registers -- the reference's kernel has 30), one launch per batch
//    (the launch boundary invalidates L1 like the reference's);
//  * rows are read through L1 (ld.global.ca = the reference's plain loads) when they are needed -- the
//    vertex row first, every target's context row right before its dot product -- and written back with
//    plain stores right after the update, so the read-to-write window of every row is the reference's;
//  * but a row is ONE 128-bit load / store per lane (a 128-d row = one 512-B transaction) held in
//    registers, the dot product a butterfly reduction, the sigmoid ex2/rcp, the update 2 FMAs per
//    element; the negatives are drawn by a pre-pass like the reference's gpu::Sample launch.
// Latency is hidden by occupancy (64 warps x 1.5 KB in flight per SM), not by software pipelining.
// -----------------------------------------------------------------------------
constexpr int kSampleBlockThreads = 512;

// an index load the compiler keeps where it is written (the reference loads head, negative and tail ids one after
// another, each right before it is needed); `word` selects .x / .y of a {tail, head} pool entry
__device__ __forceinline__ uint32_t gv_load_again_u32(const uint2 *entry, int word) {
    return __float_as_uint(gv_load_again(reinterpret_cast<const float *>(entry) + word));
}
__device__ __forceinline__ uint32_t gv_load_again_u32(const uint32_t *entry, int) {
    return __float_as_uint(gv_load_again(reinterpret_cast<const float *>(entry)));
}  // instance/gpu/graph.cuh: kThreadPerBlock = 512 -> 16 samples per block

template<int DIM>
constexpr int sample_blocks_per_sm() {
    return Row<DIM>::kPass == 1 ? 4 : (Row<DIM>::kPass == 2 ? 2 : 1);
}

template<int DIM, bool LOSS>
__global__ void __launch_bounds__(1024, sample_blocks_per_sm<DIM>() / 2)
    train_sample_per_warp_kernel(const TrainParams p) {
    const int lane = threadIdx.x & 31;
    const uint32_t num_warp = gridDim.x * (blockDim.x >> 5);  // the host keeps launches below 2^32 samples
    const uint32_t num_sample = uint32_t(p.num_sample);
    const int k = p.num_negative;
    if (p.prefetch_blocks && threadIdx.x < 32) {
        // The pool entries and negative ids are streamed once, so every warp would begin with a DRAM round trip before
        // it can even ask for its vertex row.  The first warp of a block pulls the index lines of the block that runs
        // `prefetch_blocks` later into L2: by then they are an L2 hit, the warp's rows are requested sooner, and a
        // larger share of its lifetime is the read-modify-write window -- as in the reference, whose warps live 10 us
        // and spend almost all of it between reading and writing their vertex row.
        const unsigned long long first = (unsigned long long)(blockIdx.x + p.prefetch_blocks) * (blockDim.x >> 5);
        const unsigned long long pool_bytes = (unsigned long long)(blockDim.x >> 5) * sizeof(uint2);
        const unsigned long long negative_bytes = (unsigned long long)(blockDim.x >> 5) * k * sizeof(uint32_t);
        if (first < num_sample) {
            if (lane * 128ull < pool_bytes)
                gv_prefetch_l2(reinterpret_cast<const char *>(p.pool + first) + lane * 128);
            if (lane * 128ull < negative_bytes)
                gv_prefetch_l2(reinterpret_cast<const char *>(p.negatives + first * k) + lane * 128);
        }
    }
    // the reference's grid-stride loop (gpu/graph.cuh:54); one iteration unless the launch was capped
    for (uint32_t i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < num_sample; i += num_warp) {
        const uint2 pair = __ldg(p.pool + i);  // {tail, head}: one broadcast load per warp
        const uint32_t batch = i / p.batch_size;
        const float lr = __ldg(p.lr_per_batch + batch);
        Row<DIM> v, c, unused;
        const bool l2_only = p.flags & 256;  // experiment: no L1 at all (ld.cg / st.cg)
        if (l2_only)
            load_row<DIM>(v, p.vertex + size_t(pair.y) * DIM, lane, false);
        else
            load_row_plain<DIM>(v, p.vertex + size_t(pair.y) * DIM, lane);
        if (p.flags & 128) {
            // experiment: the vertex row has arrived before the first context row is requested, as in the reference,
            // whose copy loop into shared memory completes first (gpu/graph.cuh:58-59)
            gv_wait_for(v.x[0].x);
        }
        float sample_loss = 0.f;
        for (int s = 0; s <= k; s++) {  // negatives first, then the positive (gpu/graph.cuh:62-71)
            // negatives were drawn by gv_cuda_sample_negatives before this launch (gpu::Sample, solver.h:1536-1539)
            const uint32_t tail = s < k ? __ldg(p.negatives + size_t(i) * k + s) : pair.x;
            float *context = p.context + size_t(tail) * DIM;
            if (l2_only)
                load_row<DIM>(c, context, lane, false);
            else
                load_row_plain<DIM>(c, context, lane);
            const float prob = sigmoid(dot<DIM>(v, c));
            float gradient, weight;
            if (s == k) {
                gradient = prob - 1;
                weight = 1;
                if (LOSS)
                    sample_loss += weight * -logf(prob + kEpsilon);
            } else {
                gradient = prob;
                weight = p.negative_weight;
                if (LOSS)
                    sample_loss += weight * -logf(1 - prob + kEpsilon);
            }
            backward<DIM, GV_OPT_SGD>(p.optimizer, lr, gradient, weight, v, c, unused, unused, unused, unused);
            store_row<DIM>(c, context, lane, !l2_only);
        }
        store_row<DIM>(v, p.vertex + size_t(pair.y) * DIM, lane, !l2_only);
        if (LOSS && lane == 0) {
            sample_loss = sample_loss / (1 + k * p.negative_weight);  // gpu/graph.cuh:91-92
            if (p.loss_per_sample)
                p.loss_per_sample[i] = sample_loss;
            if (p.loss_per_batch)
                atomicAdd(p.loss_per_batch + batch, sample_loss);
        }
    }
}


// -----------------------------------------------------------------------------
// kernel_flags & 1024: the same one-warp-per-sample training with RESIDENT blocks.  A block of 16 warps takes groups
// of 16 consecutive pool entries by ticket (one atomicAdd per group: groups start in pool order, like the reference's
// blocks) and, while it trains group t, already knows its next group t':
//   * every warp loads the k + 2 indices of its next sample (one lane each) during the current sample, so the next
//     vertex row is requested the moment the current one is stored -- no index round trip in front of it.  The share
//     of a warp's time that is a read-modify-write window rises to what it is in the reference;
//   * the rows of the next sample are pulled into L2 (prefetch.global.L2, 4 lanes per row): a prefetch is not a read
//     -- the values are read from L2 when they are needed, after every write that reached L2 before -- so the race
//     semantics are untouched, but a read no longer waits for DRAM.
// How many updates of a row are lost depends on how many samples are between reading and writing it at the same
// time, i.e. on the number of resident blocks: `sample_resident_blocks` (default 4 per SM) is the one calibration knob,
// set so that the Youtube-shaped model's norms match the unmodified reference's (tools/parity_sweep.py,
// profiles/r02_parity6_*.jsonl).
// -----------------------------------------------------------------------------
template<int DIM, bool LOSS>
__global__ void __launch_bounds__(kSampleBlockThreads, sample_blocks_per_sm<DIM>())
    train_resident_groups_kernel(const TrainParams p) {
    __shared__ unsigned int next_group[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t num_sample = uint32_t(p.num_sample);
    const uint32_t group_size = blockDim.x >> 5;
    const int k = p.num_negative;  // k + 2 <= 32 (checked by the host)
    const bool l2_only = p.flags & 256;

    // index l of sample i: head (l = 0), the k negatives, the positive tail (l = k + 1) -- one lane each
    auto load_index = [&](uint32_t i) -> uint32_t {
        if (i >= num_sample || lane > k + 1)
            return 0u;
        if (lane == 0)
            return __ldg(reinterpret_cast<const uint32_t *>(p.pool + i) + 1);
        if (lane == k + 1)
            return __ldg(reinterpret_cast<const uint32_t *>(p.pool + i));
        return __ldg(p.negatives + size_t(i) * k + (lane - 1));
    };

    if (threadIdx.x == 0)
        next_group[0] = atomicAdd(p.work_counter, 1u);
    __syncthreads();
    uint32_t i = next_group[0] * group_size + warp;
    uint32_t index = load_index(i);
    int parity = 1;
    while (true) {
        // the ticket of the group after this one (a uniform decision: every warp sees the same value)
        if (threadIdx.x == 0)
            next_group[parity] = atomicAdd(p.work_counter, 1u);
        __syncthreads();
        if (i - warp >= num_sample)
            break;  // this group starts past the end: so does every later ticket
        const uint32_t i_next = next_group[parity] * group_size + warp;
        parity ^= 1;
        const uint32_t index_next = load_index(i_next);  // in flight while this sample is trained
        if (i < num_sample) {
            const uint32_t head = __shfl_sync(kFullMask, index, 0);
            const uint32_t batch = i / p.batch_size;
            const float lr = __ldg(p.lr_per_batch + batch);
            Row<DIM> v, c, unused;
            float *vertex = p.vertex + size_t(head) * DIM;
            if (l2_only)
                load_row<DIM>(v, vertex, lane, false);
            else
                load_row_plain<DIM>(v, vertex, lane);
            float sample_loss = 0.f;
            for (int s = 0; s <= k; s++) {  // negatives first, then the positive (gpu/graph.cuh:62-71)
                const uint32_t tail = __shfl_sync(kFullMask, index, s + 1);
                float *context = p.context + size_t(tail) * DIM;
                if (l2_only)
                    load_row<DIM>(c, context, lane, false);
                else
                    load_row_plain<DIM>(c, context, lane);
                if (s == 0 && p.prefetch_blocks && i_next < num_sample) {
                    // rows of the next sample -> L2: lanes 4r .. 4r + 3 take the four 128-byte lines of row r
                    constexpr int kLines = (DIM * 4 + 127) / 128, kRows = 32 / kLines;
                    for (int first = 0; first <= k + 1; first += kRows) {
                        const int r = first + lane / kLines;
                        const uint32_t row = __shfl_sync(kFullMask, index_next, r & 31);
                        if (r <= k + 1 && lane < kRows * kLines) {
                            const float *base = (r == 0 ? p.vertex : p.context) + size_t(row) * DIM;
                            gv_prefetch_l2(reinterpret_cast<const char *>(base) + (lane % kLines) * 128);
                        }
                    }
                }
                const float prob = sigmoid(dot<DIM>(v, c));
                float gradient, weight;
                if (s == k) {
                    gradient = prob - 1;
                    weight = 1;
                    if (LOSS)
                        sample_loss += weight * -logf(prob + kEpsilon);
                } else {
                    gradient = prob;
                    weight = p.negative_weight;
                    if (LOSS)
                        sample_loss += weight * -logf(1 - prob + kEpsilon);
                }
                backward<DIM, GV_OPT_SGD>(p.optimizer, lr, gradient, weight, v, c, unused, unused, unused, unused);
                store_row<DIM>(c, context, lane, !l2_only);
            }
            store_row<DIM>(v, vertex, lane, !l2_only);
            if (LOSS && lane == 0) {
                sample_loss = sample_loss / (1 + k * p.negative_weight);  // gpu/graph.cuh:91-92
                if (p.loss_per_sample)
                    p.loss_per_sample[i] = sample_loss;
                if (p.loss_per_batch)
                    atomicAdd(p.loss_per_batch + batch, sample_loss);
            }
        }
        i = i_next;
        index = index_next;
    }
    // the last block to leave re-arms the ticket counter for the next launch of this stream
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(p.work_counter + 1, 1u);
        if (done + 1 == gridDim.x) {
            p.work_counter[0] = 0;
            p.work_counter[1] = 0;
        }
    }
}

// -----------------------------------------------------------------------------
// kernel_flags & 512: the reference's memory TIMELINE as well as its geometry -- a measuring instrument, not a fast
// path.  The reference's warp touches a row 128 bytes at a time (lane l owns elements l, l + 32, ..., base/vector.h:
// 62-71, util/gpu.cuh:24-27) and its loops are not unrolled (SASS of gpu::graph::train: LDG; STS; BRA for the vertex
// copy, LDG; LDS; FFMA; BRA for the dot product, LDG; LDS; ...; STS; STG; BRA for the update): every 128-byte segment
// is a separate dependent round trip, the context row is read a second time in the backward loop, and a warp lives
// ~10 us of which the vertex row's read-modify-write window is >80 %.  How long those windows are, relative to a
// warp's lifetime, decides how many concurrent updates of a row are lost.  This variant reproduces that timeline with
// the arithmetic of the kernels above; it runs at the reference kernel's own speed (~1e9 edges/s) and exists to show
// that the residual between train_sample_per_warp_kernel and the reference is the timeline's.
// -----------------------------------------------------------------------------
template<int DIM, bool LOSS>
__global__ void __launch_bounds__(kSampleBlockThreads, 4) train_reference_timeline_kernel(const TrainParams p) {
    constexpr int N = DIM / 32;
    const int lane = threadIdx.x & 31;
    const uint32_t num_warp = gridDim.x * (blockDim.x >> 5);
    const uint32_t num_sample = uint32_t(p.num_sample);
    const int k = p.num_negative;
    for (uint32_t i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < num_sample; i += num_warp) {
        const uint32_t head = gv_load_again_u32(p.pool + i, 1);  // batch[sample_id * 2 + 1]
        const uint32_t batch = i / p.batch_size;
        const float lr = __ldg(p.lr_per_batch + batch);
        float *vertex = p.vertex + size_t(head) * DIM;
        float v[N];
#pragma unroll 1
        for (int j = 0; j < N; j++) {  // vertex_buffer = vertex: one segment per round trip
            v[j] = gv_load_again(vertex + lane + 32 * j);
            gv_wait_for(v[j]);
        }
        float sample_loss = 0.f;
        for (int s = 0; s <= k; s++) {
            const uint32_t tail = s < k ? gv_load_again_u32(p.negatives + size_t(i) * k + s, 0)
                                        : gv_load_again_u32(p.pool + i, 0);
            float *context = p.context + size_t(tail) * DIM;
            float acc = 0.f;
#pragma unroll 1
            for (int j = 0; j < N; j++) {  // LINE::forward: the dot product, segment by segment
                float c = gv_load_again(context + lane + 32 * j);
                gv_wait_for(c);
                acc = fmaf(v[j], c, acc);
            }
            const float prob = sigmoid(warp_sum(acc));
            float gradient, weight;
            if (s == k) {
                gradient = prob - 1;
                weight = 1;
                if (LOSS)
                    sample_loss += weight * -logf(prob + kEpsilon);
            } else {
                gradient = prob;
                weight = p.negative_weight;
                if (LOSS)
                    sample_loss += weight * -logf(1 - prob + kEpsilon);
            }
            const float scale = lr * weight;
            const float alpha = 1.f - scale * p.optimizer.weight_decay, beta = scale * gradient;
#pragma unroll 1
            for (int j = 0; j < N; j++) {  // LINE::backward: the context row is read again (an L1 hit), then written
                const float c = gv_load_again(context + lane + 32 * j);
                const float vv = v[j];
                v[j] = fmaf(-beta, c, alpha * vv);
                context[lane + 32 * j] = fmaf(-beta, vv, alpha * c);
            }
        }
#pragma unroll 1
        for (int j = 0; j < N; j++)  // vertex = vertex_buffer
            vertex[lane + 32 * j] = v[j];
        if (LOSS && lane == 0) {
            sample_loss = sample_loss / (1 + k * p.negative_weight);
            if (p.loss_per_sample)
                p.loss_per_sample[i] = sample_loss;
            if (p.loss_per_batch)
                atomicAdd(p.loss_per_batch + batch, sample_loss);
        }
    }
}

// gpu::Sample, base/alias_table.cuh:175-183
__global__ void __launch_bounds__(256) sample_negatives_kernel(const gv_alias_entry_t *table, uint32_t count,
                                                               const double *random, unsigned long long num,
                                                               uint32_t *out) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < num; t += stride) {
        const double2 r = __ldcs(reinterpret_cast<const double2 *>(random) + t);
        out[t] = alias_sample_narrowed(table, count, r.x, r.y);
    }
}

// gpu::graph::predict, instance/gpu/graph.cuh:250-279
template<int DIM>
__global__ void __launch_bounds__(kBlockThreads) predict_kernel(const float *vertex, const float *context,
                                                                const uint2 *batch, unsigned long long num,
                                                                float *logits) {
    const int lane = threadIdx.x & 31;
    const unsigned long long num_warp = (unsigned long long)gridDim.x * (blockDim.x >> 5);
    for (unsigned long long i = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < num;
         i += num_warp) {
        const uint2 pair = __ldg(batch + i);  // {tail, head}
        Row<DIM> v, c;
        load_row<DIM>(v, vertex + size_t(pair.y) * DIM, lane);
        load_row<DIM>(c, context + size_t(pair.x) * DIM, lane);
        const float logit = dot<DIM>(v, c);
        if (lane == 0)
            logits[i] = logit;
    }
}

// -----------------------------------------------------------------------------
// tunables (gv_cuda_set_tunable; environment GV_HOT_ROWS / GV_KERNEL_FLAGS give the initial values)
// -----------------------------------------------------------------------------
// Rows are stored in degree order (partition() sorts by weighted degree, core/solver.h:877-878), so
// "local id < hot_rows" selects the hub rows.  On a power-law graph the few hottest 128-B lines
// serialise in their L2 slices (measured: +28 % at P=1, +70 % at P=2 once the top rows are read through
// L1).  The price is the reference's own staleness -- its kernels load every row through L1 -- bounded
// here by the launch granularity (L1 is invalidated between launches).
static uint32_t g_hot_rows = getenv("GV_HOT_ROWS") ? uint32_t(atol(getenv("GV_HOT_ROWS"))) : 128;
static int g_kernel_flags = getenv("GV_KERNEL_FLAGS") ? atoi(getenv("GV_KERNEL_FLAGS")) : 0;
// resident train CTAs per SM (0 = as many as fit).  One less than the maximum leaves registers and
// thread slots for the samplers' kernels, which otherwise only run between train launches.
static int g_blocks_per_sm = getenv("GV_TRAIN_BLOCKS_PER_SM") ? atoi(getenv("GV_TRAIN_BLOCKS_PER_SM")) : 0;
// experiment: size the persistent grid for this many SMs fewer than the device has, so that some SMs keep room for
// the samplers' kernels while a train launch is resident (0 = use every SM)
static int g_reserve_sms = getenv("GV_TRAIN_RESERVE_SMS") ? atoi(getenv("GV_TRAIN_RESERVE_SMS")) : 0;
// experiment: cap the grid of the walk kernels (grid-stride loops), so that they take a slice of the device next to a
// resident train launch instead of all of it between two launches (0 = one thread per walk)
// comparison switch: 1 = fill the single-block pool with one thread per walk (scattered 8-byte stores) instead of the
// tiled, coalesced fill
static int g_fill_per_walk = getenv("GV_FILL_PER_WALK") ? atoi(getenv("GV_FILL_PER_WALK")) : 0;
static int g_sampler_max_ctas = getenv("GV_SAMPLER_MAX_CTAS") ? atoi(getenv("GV_SAMPLER_MAX_CTAS")) : 0;
// threads per block of the one-warp-per-sample kernel (the reference: 512 = 16 samples per block, 64 warps per SM)
static int g_sample_block_threads = getenv("GV_SAMPLE_BLOCK_THREADS") ? atoi(getenv("GV_SAMPLE_BLOCK_THREADS")) : 512;
// one-warp-per-sample kernel: how many blocks ahead the index lines are prefetched into L2 (0 = off; 592 = one
// resident wave of 148 SMs x 4 blocks, the setting closest to the reference in profiles/r02_parity4_*.jsonl)
// resident blocks of train_resident_groups_kernel (0 = 4 per SM): the calibration knob of its race statistics
static int g_sample_resident_blocks = getenv("GV_SAMPLE_RESIDENT_BLOCKS") ? atoi(getenv("GV_SAMPLE_RESIDENT_BLOCKS")) : 0;
static int g_sample_prefetch_blocks = getenv("GV_SAMPLE_PREFETCH_BLOCKS") ? atoi(getenv("GV_SAMPLE_PREFETCH_BLOCKS")) : 592;

// -----------------------------------------------------------------------------
// launch helpers
// -----------------------------------------------------------------------------
// kernel_flags & 8: the ticket counter of the launches of one stream (two zeroed words; the kernel re-arms them)
static unsigned int *work_counter_for(cudaStream_t stream) {
    static std::mutex mutex;
    static std::map<std::pair<int, cudaStream_t>, unsigned int *> counters;
    int device = 0;
    if (cudaGetDevice(&device) != cudaSuccess)
        return nullptr;
    std::lock_guard<std::mutex> lock(mutex);
    auto found = counters.find({device, stream});
    if (found != counters.end())
        return found->second;
    unsigned int *counter = nullptr;
    if (cudaMalloc(&counter, 2 * sizeof(unsigned int)) != cudaSuccess ||
        cudaMemset(counter, 0, 2 * sizeof(unsigned int)) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;  // fall back to the fixed stride
    }
    counters[{device, stream}] = counter;
    return counter;
}

// grow-only device buffer for the negatives of one launch, per (device, stream): launches of a stream are serial
static uint32_t *negative_scratch_for(cudaStream_t stream, unsigned long long count) {
    static std::mutex mutex;
    static std::map<std::pair<int, cudaStream_t>, std::pair<uint32_t *, unsigned long long>> buffers;
    int device = 0;
    if (cudaGetDevice(&device) != cudaSuccess)
        return nullptr;
    std::lock_guard<std::mutex> lock(mutex);
    auto &entry = buffers[{device, stream}];
    if (entry.second < count) {
        if (entry.first) {
            cudaStreamSynchronize(stream);  // an earlier launch may still read the old buffer
            cudaFree(entry.first);
        }
        entry = {nullptr, 0};
        const unsigned long long capacity = count + count / 4;
        if (cudaMalloc(&entry.first, capacity * sizeof(uint32_t)) != cudaSuccess) {
            cudaGetLastError();
            entry.first = nullptr;
            return nullptr;
        }
        entry.second = capacity;
    }
    return entry.first;
}

static int device_sm_count() {
    int device = 0, sms = 0;
    if (cudaGetDevice(&device) != cudaSuccess)
        return 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || sms <= 0)
        return 148;
    return sms;
}

static int launch_train(void (*kernel)(const TrainParams), const TrainParams &p, int num_warps,
                        cudaStream_t stream) {
    int threads = kBlockThreads;
    int blocks;
    if (num_warps > 0) {
        threads = num_warps >= kBlockThreads / 32 ? kBlockThreads : num_warps * 32;
        blocks = (num_warps * 32 + threads - 1) / threads;
    }
    const size_t shared_bytes = size_t(threads / 32) * 32 * (p.num_negative + 2) * sizeof(uint32_t);
    if (shared_bytes > 200 * 1024)
        return fail("num_negative too large for the train kernel's shared-memory staging");
    if (shared_bytes > 48 * 1024)
        GV_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(shared_bytes)));
    if (num_warps <= 0) {
        int per_sm = 0;
        GV_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, shared_bytes));
        if (g_blocks_per_sm > 0 && per_sm > g_blocks_per_sm)
            per_sm = g_blocks_per_sm;
        if (per_sm < 1)
            per_sm = 1;
        const int sms = device_sm_count();
        blocks = (g_reserve_sms > 0 && g_reserve_sms < sms ? sms - g_reserve_sms : sms) * per_sm;  // one resident wave
        const unsigned long long needed = ((p.num_sample + 31) / 32 + threads / 32 - 1) / (threads / 32);
        if ((unsigned long long)blocks > needed)
            blocks = int(needed);
    }
    if (blocks < 1)
        blocks = 1;
    GV_LAUNCH(blocks, threads, shared_bytes, stream, kernel)(p);
    GV_CUDA_OK(cudaGetLastError());
    return 0;
}

template<int DIM, int K>
static int launch_sgd(const TrainParams &p, int num_warps, cudaStream_t stream) {
    if (p.loss_per_sample || p.loss_per_batch)
        return launch_train(train_sgd_kernel<DIM, K, true>, p, num_warps, stream);
    return launch_train(train_sgd_kernel<DIM, K, false>, p, num_warps, stream);
}


// one warp per sample, 16 samples per block, blocks in pool order (the reference's <<<8192, 512>>> covers a batch of up
// to 131 072 samples in one pass; longer launches simply get more blocks so that the order is kept)
template<int DIM>
static int launch_sample_per_warp(const TrainParams &p, int num_warps, cudaStream_t stream) {
    // experiment knob: warps per block -> resident warps per SM (32 registers: floor(2048 / threads) blocks)
    int threads = g_sample_block_threads >= 32 && g_sample_block_threads <= 1024 ? g_sample_block_threads / 32 * 32
                                                                                 : kSampleBlockThreads;
    if (Row<DIM>::kPass > 1)
        threads = kSampleBlockThreads;
    unsigned long long blocks = (p.num_sample + threads / 32 - 1) / (threads / 32);
    if (num_warps > 0) {  // tests: a fixed number of warps (1 = sequential) walking the pool with the grid stride
        threads = num_warps >= kSampleBlockThreads / 32 ? kSampleBlockThreads : num_warps * 32;
        blocks = (num_warps * 32 + threads - 1) / threads;
    }
    if (blocks > 0x7FFFFFFFull || p.num_sample > 0xFFFFFFFFull)
        return fail("gv_cuda_train_block: too many samples for one launch");
    TrainParams q = p;
    if (!q.negatives && q.num_negative > 0) {
        // gpu::Sample as its own launch (core/solver.h:1536-1539): into the caller's capture buffer, else a scratch
        const unsigned long long count = q.num_sample * q.num_negative;
        uint32_t *drawn = q.negatives_out ? q.negatives_out : negative_scratch_for(stream, count);
        if (!drawn)
            return fail("gv_cuda_train_block: cannot allocate the negative sample buffer");
        unsigned long long draw_blocks = (count + 255) / 256;
        draw_blocks = std::min<unsigned long long>(draw_blocks, (unsigned long long)device_sm_count() * 8);
        GV_LAUNCH(int(draw_blocks), 256, 0, stream, sample_negatives_kernel)(q.negative_table, q.negative_count, q.random,
                                                                            count, drawn);
        GV_CUDA_OK(cudaGetLastError());
        q.negatives = drawn;
    } else if (q.negatives && q.negatives_out && q.num_negative > 0)
        GV_CUDA_OK(cudaMemcpyAsync(q.negatives_out, q.negatives, q.num_sample * q.num_negative * sizeof(uint32_t),
                                   cudaMemcpyDeviceToDevice, stream));
    if ((p.flags & 1024) && num_warps <= 0 && q.num_negative + 2 <= 32 && DIM % 32 == 0) {
        // resident blocks taking groups of 16 samples by ticket (train_resident_groups_kernel)
        q.work_counter = work_counter_for(stream);
        if (!q.work_counter)
            return fail("gv_cuda_train_block: cannot allocate the ticket counter");
        void (*resident)(const TrainParams) = (p.loss_per_sample || p.loss_per_batch)
                                                  ? train_resident_groups_kernel<DIM, true>
                                                  : train_resident_groups_kernel<DIM, false>;
        static bool resident_configured[2] = {false, false};
        const int slot = (p.loss_per_sample || p.loss_per_batch) ? 1 : 0;
        if (!resident_configured[slot]) {
            cudaFuncSetAttribute(resident, cudaFuncAttributePreferredSharedMemoryCarveout, 25);
            cudaGetLastError();
            resident_configured[slot] = true;
        }
        const unsigned long long groups = (p.num_sample + kSampleBlockThreads / 32 - 1) / (kSampleBlockThreads / 32);
        unsigned long long resident_blocks = g_sample_resident_blocks > 0
                                                 ? (unsigned long long)g_sample_resident_blocks
                                                 : (unsigned long long)device_sm_count() * sample_blocks_per_sm<DIM>();
        resident_blocks = std::min(resident_blocks, groups);
        GV_LAUNCH(int(resident_blocks), kSampleBlockThreads, 0, stream, resident)(q);
        GV_CUDA_OK(cudaGetLastError());
        return 0;
    }
    const bool timeline = (p.flags & 512) && DIM % 32 == 0;
    void (*kernel)(const TrainParams) =
        (p.loss_per_sample || p.loss_per_batch)
            ? (timeline ? train_reference_timeline_kernel<DIM, true> : train_sample_per_warp_kernel<DIM, true>)
            : (timeline ? train_reference_timeline_kernel<DIM, false> : train_sample_per_warp_kernel<DIM, false>);
    static bool configured[4] = {false, false, false, false};
    const int which = ((p.loss_per_sample || p.loss_per_batch) ? 1 : 0) + (timeline ? 2 : 0);
    if (timeline && num_warps <= 0 && threads != kSampleBlockThreads) {
        threads = kSampleBlockThreads;
        blocks = (p.num_sample + threads / 32 - 1) / (threads / 32);
    }
    if (!configured[which]) {
        // the reference's kernel keeps 4 x 9 KB of shared memory per SM: ask for the same L1 / shared split, so that an
        // L1 line lives about as long here as there (best effort -- a hint, not an error)
        cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 25);
        cudaGetLastError();
        configured[which] = true;
    }
    GV_LAUNCH(int(blocks), threads, 0, stream, kernel)(q);
    GV_CUDA_OK(cudaGetLastError());
    return 0;
}

template<int DIM, int OPT>
static int dispatch_loss(const TrainParams &p, int num_warps, cudaStream_t stream) {
    // SGD: the reference's geometry (one warp per sample) unless the persistent kernels are asked for
    if (OPT == GV_OPT_SGD && !(p.flags & 64))
        return launch_sample_per_warp<DIM>(p, num_warps, stream);
    // the pipelined SGD kernel for the usual small k, when two samples' rows fit in registers
    if (OPT == GV_OPT_SGD && Row<DIM>::kPass <= 2 && !(p.flags & 4)) {
        switch (p.num_negative) {
            case 1: return launch_sgd<DIM, 1>(p, num_warps, stream);
            case 2: return launch_sgd<DIM, 2>(p, num_warps, stream);
            case 3: return launch_sgd<DIM, 3>(p, num_warps, stream);
        }
    }
    if (p.loss_per_sample || p.loss_per_batch)
        return launch_train(train_kernel<DIM, OPT, true>, p, num_warps, stream);
    return launch_train(train_kernel<DIM, OPT, false>, p, num_warps, stream);
}

template<int DIM>
static int dispatch_optimizer(const TrainParams &p, int num_warps, cudaStream_t stream) {
    switch (p.optimizer.type) {
        case GV_OPT_SGD: return dispatch_loss<DIM, GV_OPT_SGD>(p, num_warps, stream);
        case GV_OPT_MOMENTUM: return dispatch_loss<DIM, GV_OPT_MOMENTUM>(p, num_warps, stream);
        case GV_OPT_ADAGRAD: return dispatch_loss<DIM, GV_OPT_ADAGRAD>(p, num_warps, stream);
        case GV_OPT_RMSPROP: return dispatch_loss<DIM, GV_OPT_RMSPROP>(p, num_warps, stream);
        case GV_OPT_ADAM: return dispatch_loss<DIM, GV_OPT_ADAM>(p, num_warps, stream);
    }
    return fail("unknown optimizer type " + std::to_string(p.optimizer.type));
}

}  // namespace device

int sampler_max_ctas() {
    return device::g_sampler_max_ctas;
}

bool direct_fill_per_walk() {
    return device::g_fill_per_walk != 0;
}

}  // namespace gv

using namespace gv;
using namespace gv::device;

extern "C" {

int gv_cuda_train_block(const gv_matrices_t *m, const uint32_t *pool, uint64_t num_sample, int num_negative,
                        const uint32_t *negatives, const double *random, const gv_alias_entry_t *negative_table,
                        uint32_t negative_count, uint32_t *negatives_out, const gv_device_optimizer_t *optimizer,
                        const float *lr_per_batch, uint32_t batch_size, float negative_weight,
                        float *loss_per_sample, float *loss_per_batch, int num_warps, void *stream) {
    if (num_sample == 0)
        return 0;
    if (!m || !pool || !optimizer || !lr_per_batch)
        return fail("gv_cuda_train_block: null argument");
    if (num_negative < 0 || batch_size == 0)
        return fail("gv_cuda_train_block: invalid num_negative / batch_size");
    if (!negatives && num_negative > 0 && (!random || !negative_table || negative_count == 0))
        return fail("gv_cuda_train_block: need either `negatives` or (`random`, `negative_table`)");
    int num_moment = optimizer->type == GV_OPT_SGD ? 0 : (optimizer->type == GV_OPT_ADAM ? 2 : 1);
    if (!m->vertex || !m->context || (num_moment >= 1 && (!m->vertex_m1 || !m->context_m1)) ||
        (num_moment >= 2 && (!m->vertex_m2 || !m->context_m2)))
        return fail("gv_cuda_train_block: missing embedding / moment matrix");
    TrainParams p;
    p.vertex = m->vertex;
    p.context = m->context;
    p.vertex_m1 = m->vertex_m1;
    p.context_m1 = m->context_m1;
    p.vertex_m2 = m->vertex_m2;
    p.context_m2 = m->context_m2;
    p.pool = reinterpret_cast<const uint2 *>(pool);
    p.num_sample = num_sample;
    p.num_negative = num_negative;
    p.negatives = negatives;
    p.random = random;
    p.negative_table = negative_table;
    p.negative_count = negative_count;
    p.negatives_out = negatives_out;
    p.optimizer = *optimizer;
    p.lr_per_batch = lr_per_batch;
    p.batch_size = batch_size;
    p.negative_weight = negative_weight;
    p.loss_per_sample = loss_per_sample;
    p.loss_per_batch = loss_per_batch;
    p.flags = g_kernel_flags;
    p.hot_rows = g_hot_rows;
    p.prefetch_blocks = uint32_t(g_sample_prefetch_blocks);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    p.work_counter = (g_kernel_flags & 8) ? work_counter_for(s) : nullptr;
    switch (m->dim) {  // src/graphvite.cu:52-59 instantiates exactly these dimensions
        case 32: return dispatch_optimizer<32>(p, num_warps, s);
        case 64: return dispatch_optimizer<64>(p, num_warps, s);
        case 96: return dispatch_optimizer<96>(p, num_warps, s);
        case 128: return dispatch_optimizer<128>(p, num_warps, s);
        case 256: return dispatch_optimizer<256>(p, num_warps, s);
        case 512: return dispatch_optimizer<512>(p, num_warps, s);
    }
    return fail("unsupported embedding dimension " + std::to_string(m->dim) + " (32, 64, 96, 128, 256, 512)");
}

int gv_cuda_set_tunable(const char *name, long value) {
    const std::string key = name ? name : "";
    if (key == "hot_rows")
        g_hot_rows = value < 0 ? 0u : uint32_t(value);
    else if (key == "kernel_flags")
        g_kernel_flags = int(value);
    else if (key == "train_blocks_per_sm")
        g_blocks_per_sm = int(value);
    else if (key == "fill_per_walk")
        g_fill_per_walk = int(value);
    else if (key == "sampler_max_ctas")
        g_sampler_max_ctas = int(value);
    else if (key == "train_reserve_sms")
        g_reserve_sms = int(value);
    else if (key == "sample_block_threads")
        g_sample_block_threads = int(value);
    else if (key == "sample_prefetch_blocks")
        g_sample_prefetch_blocks = int(value < 0 ? 0 : value);
    else if (key == "sample_resident_blocks")
        g_sample_resident_blocks = int(value < 0 ? 0 : value);
    else if (key == "kg_flags")
        set_kg_kernel_flags(int(value));
    else
        return fail("unknown tunable `" + key + "`");
    return 0;
}

long gv_cuda_get_tunable(const char *name) {
    const std::string key = name ? name : "";
    if (key == "hot_rows")
        return long(g_hot_rows);
    if (key == "kernel_flags")
        return g_kernel_flags;
    if (key == "train_blocks_per_sm")
        return g_blocks_per_sm;
    if (key == "fill_per_walk")
        return g_fill_per_walk;
    if (key == "sampler_max_ctas")
        return g_sampler_max_ctas;
    if (key == "train_reserve_sms")
        return g_reserve_sms;
    if (key == "sample_block_threads")
        return g_sample_block_threads;
    if (key == "sample_prefetch_blocks")
        return g_sample_prefetch_blocks;
    if (key == "sample_resident_blocks")
        return g_sample_resident_blocks;
    if (key == "kg_flags")
        return kg_kernel_flags();
    fail("unknown tunable `" + key + "`");
    return -1;
}

int gv_cuda_sample_negatives(const gv_alias_entry_t *table, uint32_t count, const double *random, uint64_t num,
                             uint32_t *out, void *stream) {
    if (num == 0)
        return 0;
    if (!table || !random || !out || count == 0)
        return fail("gv_cuda_sample_negatives: null argument");
    unsigned long long blocks = (num + 255) / 256;
    const unsigned long long cap = (unsigned long long)device_sm_count() * 8;
    if (blocks > cap)
        blocks = cap;
    GV_LAUNCH(int(blocks), 256, 0, static_cast<cudaStream_t>(stream), sample_negatives_kernel)(table, count, random, num, out);
    GV_CUDA_OK(cudaGetLastError());
    return 0;
}

int gv_cuda_predict(int dim, const float *vertex, const float *context, const uint32_t *batch, uint64_t num,
                    float *logits, void *stream) {
    if (num == 0)
        return 0;
    if (!vertex || !context || !batch || !logits)
        return fail("gv_cuda_predict: null argument");
    unsigned long long blocks = (num + kBlockThreads / 32 - 1) / (kBlockThreads / 32);
    const unsigned long long cap = (unsigned long long)device_sm_count() * 8;
    if (blocks > cap)
        blocks = cap;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const uint2 *pairs = reinterpret_cast<const uint2 *>(batch);
    switch (dim) {
        case 32: GV_LAUNCH(int(blocks), kBlockThreads, 0, s, predict_kernel<32>)(vertex, context, pairs, num, logits); break;
        case 64: GV_LAUNCH(int(blocks), kBlockThreads, 0, s, predict_kernel<64>)(vertex, context, pairs, num, logits); break;
        case 96: GV_LAUNCH(int(blocks), kBlockThreads, 0, s, predict_kernel<96>)(vertex, context, pairs, num, logits); break;
        case 128: GV_LAUNCH(int(blocks), kBlockThreads, 0, s, predict_kernel<128>)(vertex, context, pairs, num, logits); break;
        case 256: GV_LAUNCH(int(blocks), kBlockThreads, 0, s, predict_kernel<256>)(vertex, context, pairs, num, logits); break;
        case 512: GV_LAUNCH(int(blocks), kBlockThreads, 0, s, predict_kernel<512>)(vertex, context, pairs, num, logits); break;
        default: return fail("unsupported embedding dimension " + std::to_string(dim));
    }
    GV_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // extern "C"
