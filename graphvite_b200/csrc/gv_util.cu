// Row gather / scatter between a [num_vertex][dim] matrix in global-id order and a partition
// block in local-id order.  Device-side replacement for Memory::gather / Memory::scatter
// (reference include/base/memory.h:194-217), which the reference runs on one CPU thread.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <string>

#include "gv_common.h"
#include "gv_device.cuh"

namespace gv {
namespace device {

// one warp per row, float4 lanes; dim % 4 == 0
__global__ void __launch_bounds__(256) move_rows_kernel(float *dst, const float *src, const uint32_t *ids,
                                                        unsigned long long num_row, int dim, bool gather) {
    const int lane = threadIdx.x & 31;
    const unsigned long long num_warp = (unsigned long long)gridDim.x * (blockDim.x >> 5);
    const int vec = dim / 4;
    for (unsigned long long row = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
         row < num_row; row += num_warp) {
        const size_t global = __ldg(ids + row);
        const float4 *from = reinterpret_cast<const float4 *>(src) + (gather ? global : row) * vec;
        float4 *to = reinterpret_cast<float4 *>(dst) + (gather ? row : global) * vec;
        for (int i = lane; i < vec; i += 32)
            to[i] = from[i];
    }
}

// what the host would otherwise compute and upload for a graph whose edges all weigh the same: a constant, the
// identity alias column, and the source vertex of every CSR slot
__global__ void __launch_bounds__(256) fill_float_kernel(float *dst, unsigned long long n, float value) {
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x)
        dst[i] = value;
}

__global__ void __launch_bounds__(256) fill_identity_kernel(unsigned long long *dst, unsigned long long n) {
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x)
        dst[i] = i;
}

// edge_u[e] = v for offsets[v] <= e < offsets[v + 1]: a warp per vertex, lanes stride over its range
__global__ void __launch_bounds__(256) expand_sources_kernel(const unsigned long long *offsets, uint32_t num_vertex,
                                                             uint32_t *edge_u) {
    const int lane = threadIdx.x & 31;
    const unsigned long long num_warp = (unsigned long long)gridDim.x * (blockDim.x >> 5);
    for (unsigned long long v = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); v < num_vertex;
         v += num_warp) {
        const unsigned long long begin = __ldg(offsets + v), end = __ldg(offsets + v + 1);
        for (unsigned long long e = begin + lane; e < end; e += 32)
            edge_u[e] = uint32_t(v);
    }
}

}  // namespace device
}  // namespace gv

extern "C" {

int gv_cuda_fill_float(float *dst, uint64_t n, float value, void *stream) {
    if (n == 0)
        return 0;
    if (!dst)
        return gv::fail("gv_cuda_fill_float: null argument");
    const unsigned blocks = unsigned(std::min<uint64_t>((n + 255) / 256, 148 * 16));
    GV_LAUNCH(blocks, 256, 0, static_cast<cudaStream_t>(stream), gv::device::fill_float_kernel)(dst, n, value);
    GV_CUDA_OK(cudaGetLastError());
    return 0;
}

int gv_cuda_fill_identity(uint64_t *dst, uint64_t n, void *stream) {
    if (n == 0)
        return 0;
    if (!dst)
        return gv::fail("gv_cuda_fill_identity: null argument");
    const unsigned blocks = unsigned(std::min<uint64_t>((n + 255) / 256, 148 * 16));
    GV_LAUNCH(blocks, 256, 0, static_cast<cudaStream_t>(stream), gv::device::fill_identity_kernel)(
        reinterpret_cast<unsigned long long *>(dst), n);
    GV_CUDA_OK(cudaGetLastError());
    return 0;
}

int gv_cuda_expand_sources(const uint64_t *offsets, uint32_t num_vertex, uint32_t *edge_u, void *stream) {
    if (num_vertex == 0)
        return 0;
    if (!offsets || !edge_u)
        return gv::fail("gv_cuda_expand_sources: null argument");
    const unsigned blocks = unsigned(std::min<uint64_t>((uint64_t(num_vertex) + 7) / 8, 148 * 16));
    GV_LAUNCH(blocks, 256, 0, static_cast<cudaStream_t>(stream), gv::device::expand_sources_kernel)(
        reinterpret_cast<const unsigned long long *>(offsets), num_vertex, edge_u);
    GV_CUDA_OK(cudaGetLastError());
    return 0;
}

// dst[i] = src[ids[i]] (gather != 0) or dst[ids[i]] = src[i] (gather == 0), rows of `dim` floats
int gv_cuda_move_rows(float *dst, const float *src, const uint32_t *ids, uint64_t num_row, int dim, int gather,
                      void *stream) {
    if (num_row == 0)
        return 0;
    if (!dst || !src || !ids || dim <= 0 || dim % 4 != 0)
        return gv::fail("gv_cuda_move_rows: invalid argument");
    unsigned long long blocks = (num_row + 7) / 8;
    if (blocks > 148 * 16)
        blocks = 148 * 16;
    GV_LAUNCH(int(blocks), 256, 0, static_cast<cudaStream_t>(stream), gv::device::move_rows_kernel)(dst, src, ids, num_row,
                                                                                             dim, gather != 0);
    GV_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // extern "C"
