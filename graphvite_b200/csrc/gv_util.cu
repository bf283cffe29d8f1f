// Row gather / scatter between a [num_vertex][dim] matrix in global-id order and a partition
// block in local-id order.  Device-side replacement for Memory::gather / Memory::scatter
// (reference include/base/memory.h:194-217), which the reference runs on one CPU thread.
#include <cuda_runtime.h>

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

}  // namespace device
}  // namespace gv

extern "C" {

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
