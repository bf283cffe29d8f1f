// =============================================================================
// gv_kg_sampler.cu -- the positive-sample draw of the knowledge-graph solver and the write-back of the
// global relation matrix, on the device.
//
// Replaces, for KnowledgeGraphSolver:
//   SamplerMixin::sample                 include/core/solver.h:1011-1055 (draw part; the append into the
//                                        pool blocks is gv_sampler.cu's stable partition with attributes)
//   KnowledgeGraphSampler::get_attributes include/instance/knowledge_graph.cuh:300-302
//   WorkerMixin::write_embedding (kGlobal) include/core/solver.h:1413-1420: global -= loaded - trained
//
// Draw d of a sample() call consumes doubles [2d, 2d+2) of the sampler's stream (5e6 is even, so refill
// buffers are used up exactly and the draws of a call are contiguous in the stream): one thread per draw.
// =============================================================================
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "gv_common.h"
#include "gv_device.cuh"

namespace gv {
namespace device {

// edge_table.sample(random[r++], random[r++]) with gcc's right-to-left evaluation:
// rand1 (index, kept double) = random[r + 1], rand2 (accept, narrowed to float) = random[r]
__global__ void __launch_bounds__(256) kg_draw_kernel(const gv_device_kgraph_t g, const double *random,
                                                      uint32_t num_draw, gv_location_t *chains, uint32_t *relations) {
    const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= num_draw)
        return;
    const double2 draw = __ldcs(reinterpret_cast<const double2 *>(random) + d);
    unsigned long long index = (unsigned long long)(draw.y * double(g.num_edge));
    if (index >= g.num_edge)  // cuRAND doubles lie in (0, 1]: clamp rand1 == 1 (the reference reads out of bounds)
        index = g.num_edge - 1;
    const unsigned long long edge = float(draw.x) < __ldg(g.edge_prob + index) ? index : __ldg(g.edge_alias + index);
    const uint2 *locations = reinterpret_cast<const uint2 *>(g.locations);
    uint2 *out = reinterpret_cast<uint2 *>(chains);
    out[d] = __ldg(locations + __ldg(g.edge_h + edge));                    // chain position 0: head
    out[size_t(num_draw) + d] = __ldg(locations + __ldg(g.edge_t + edge));  // chain position 1: tail
    relations[d] = __ldg(g.edge_r + edge);
}

__global__ void __launch_bounds__(256) kg_relation_delta_kernel(const float *global, const float *work, float *delta,
                                                                unsigned long long n) {
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x)
        delta[i] = global[i] - work[i];
}

__global__ void __launch_bounds__(256) kg_relation_apply_kernel(float *global, float *work, const float *delta,
                                                                unsigned long long n) {
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const float value = global[i] - delta[i];
        global[i] = value;
        work[i] = value;
    }
}

}  // namespace device
}  // namespace gv

using namespace gv;
using namespace gv::device;

extern "C" {

int gv_cuda_kg_draw(const gv_device_kgraph_t *graph, const double *random, uint32_t num_draw, gv_location_t *chains,
                    uint32_t *relations, void *stream) {
    if (num_draw == 0)
        return 0;
    if (!graph || !random || !chains || !relations || graph->num_edge == 0)
        return fail("gv_cuda_kg_draw: invalid argument");
    GV_LAUNCH((num_draw + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream), kg_draw_kernel)(*graph, random, num_draw,
                                                                                               chains, relations);
    GV_CUDA_OK(cudaGetLastError());
    return 0;
}

static unsigned elementwise_blocks(uint64_t n) {
    const uint64_t blocks = (n + 255) / 256;
    return unsigned(blocks < 1 ? 1 : (blocks > 148 * 8 ? 148 * 8 : blocks));
}

int gv_cuda_kg_relation_delta(const float *global, const float *work, float *delta, uint64_t n, void *stream) {
    if (n == 0)
        return 0;
    if (!global || !work || !delta)
        return fail("gv_cuda_kg_relation_delta: null argument");
    GV_LAUNCH(elementwise_blocks(n), 256, 0, static_cast<cudaStream_t>(stream), kg_relation_delta_kernel)(global, work, delta,
                                                                                                        n);
    GV_CUDA_OK(cudaGetLastError());
    return 0;
}

int gv_cuda_kg_relation_apply(float *global, float *work, const float *delta, uint64_t n, void *stream) {
    if (n == 0)
        return 0;
    if (!global || !work || !delta)
        return fail("gv_cuda_kg_relation_apply: null argument");
    GV_LAUNCH(elementwise_blocks(n), 256, 0, static_cast<cudaStream_t>(stream), kg_relation_apply_kernel)(global, work, delta,
                                                                                                        n);
    GV_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // extern "C"
