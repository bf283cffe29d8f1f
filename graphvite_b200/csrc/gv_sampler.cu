// =============================================================================
// gv_sampler.cu -- the positive-sample path on the device.
//
// Replaces the CPU sampler threads of the reference:
//   GraphSampler::sample_random_walk   include/instance/graph.cuh:376-450
//   SamplerMixin::sample               include/core/solver.h:1011-1055
//   AliasTable::sample (CPU call shape) include/base/alias_table.cuh:148-152
//
// The reference walks sequentially, but with no dead ends every walk consumes exactly
// 2*L doubles, so walk w owns random[2*L*w, 2*L*(w+1)) and walks are independent given
// the stream: one thread per walk.  The sequential part -- appending each pair to its
// (head part, tail part) block until that block's slice is full -- is a stable
// partition, done as count / scan / scatter so that the pools are bit-identical to the
// reference's (stream order preserved inside every block, pseudo-shuffle included).
// =============================================================================
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "gv_common.h"

namespace gv {
namespace device {

// AliasTable::sample with the CPU call shape table.sample(random[r++], random[r++]):
// gcc evaluates the arguments right to left, so rand1 (index draw) = random[r+1] and
// rand2 (accept draw) = random[r].  rand1 is NOT narrowed here.
template<class Count>
__device__ __forceinline__ Count alias_index(double rand1, Count count) {
    Count index = Count(rand1 * double(count));
    return index < count ? index : count - 1;  // cuRAND doubles lie in (0,1]; clamp rand1 == 1
}

__global__ void __launch_bounds__(256) random_walk_kernel(const gv_device_graph_t g, const double *random,
                                                          uint32_t num_walk, int walk_length,
                                                          gv_location_t *chains) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= num_walk)
        return;
    const double2 *r = reinterpret_cast<const double2 *>(random) + size_t(w) * walk_length;
    uint2 *out = reinterpret_cast<uint2 *>(chains) + w;
    const uint2 *locations = reinterpret_cast<const uint2 *>(g.locations);

    // first edge from the global edge table (AliasTable<float, size_t>)
    double2 draw = __ldcs(r);  // .x = random[r] (accept), .y = random[r+1] (index)
    unsigned long long index = alias_index<unsigned long long>(draw.y, g.num_edge);
    unsigned long long edge = float(draw.x) < __ldg(g.edge_prob + index) ? index : __ldg(g.edge_alias + index);
    uint32_t current = __ldg(g.edge_u + edge);
    out[0] = __ldg(locations + current);
    current = __ldg(g.edge_v + edge);
    out[num_walk] = __ldg(locations + current);
    // remaining steps from the per-vertex tables (AliasTable<float, Index>)
    for (int j = 2; j <= walk_length; j++) {
        const unsigned long long begin = __ldg(g.offsets + current);
        const uint32_t degree = uint32_t(__ldg(g.offsets + current + 1) - begin);
        if (degree == 0) {  // dead end: the host refuses such graphs; never read out of bounds
            for (; j <= walk_length; j++)
                out[size_t(j) * num_walk] = __ldg(locations + current);
            break;
        }
        draw = __ldcs(r + j - 1);
        const uint32_t slot = alias_index<uint32_t>(draw.y, degree);
        const uint2 entry = __ldg(reinterpret_cast<const uint2 *>(g.vertex_tables) + begin + slot);
        const uint32_t neighbor = float(draw.x) < __uint_as_float(entry.x) ? slot : entry.y;
        current = __ldg(g.edge_v + begin + neighbor);
        out[size_t(j) * num_walk] = __ldg(locations + current);
    }
}

// ---- pool fill -----------------------------------------------------------------
struct FillParams {
    int num_partition, walk_length, augmentation_step, shuffle_base;
    unsigned long long pool_size, start, slice;  // slice = end - start
};

// pass 1: counts[w][b] = pairs of walk w that belong to block b
__global__ void __launch_bounds__(256) fill_count_kernel(const FillParams p, const uint2 *chains, uint32_t num_walk,
                                                         uint32_t *counts) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= num_walk)
        return;
    const int num_block = p.num_partition * p.num_partition;
    uint32_t *row = counts + size_t(w) * num_block;
    for (int b = 0; b < num_block; b++)
        row[b] = 0;
    for (int j = 0; j < p.walk_length; j++) {
        const uint32_t head_part = chains[size_t(j) * num_walk + w].x;
        for (int k = 1; k <= p.augmentation_step && j + k <= p.walk_length; k++) {
            const uint32_t tail_part = chains[size_t(j + k) * num_walk + w].x;
            row[head_part * p.num_partition + tail_part]++;
        }
    }
}

// pass 2: one CTA per block b: exclusive scan of counts[:, b] over the walks, seeded with fill[b];
// bases saturate at the slice length (anything beyond is dropped anyway).
__global__ void __launch_bounds__(1024) fill_scan_kernel(const FillParams p, uint32_t num_walk, uint32_t *counts,
                                                         unsigned long long *fill) {
    __shared__ unsigned long long partial[1024];
    const int num_block = p.num_partition * p.num_partition;
    const int b = blockIdx.x;
    const uint32_t per_thread = (num_walk + blockDim.x - 1) / blockDim.x;
    const uint32_t begin = min(num_walk, threadIdx.x * per_thread), end = min(num_walk, begin + per_thread);
    unsigned long long sum = 0;
    for (uint32_t w = begin; w < end; w++)
        sum += counts[size_t(w) * num_block + b];
    partial[threadIdx.x] = sum;
    __syncthreads();
    // Hillis-Steele inclusive scan over the 1024 partial sums
    for (int offset = 1; offset < blockDim.x; offset <<= 1) {
        unsigned long long add = threadIdx.x >= offset ? partial[threadIdx.x - offset] : 0;
        __syncthreads();
        partial[threadIdx.x] += add;
        __syncthreads();
    }
    const unsigned long long seed = fill[b];
    unsigned long long running = seed + partial[threadIdx.x] - sum;
    for (uint32_t w = begin; w < end; w++) {
        const uint32_t count = counts[size_t(w) * num_block + b];
        counts[size_t(w) * num_block + b] = uint32_t(min(running, p.slice));
        running += count;
    }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1)
        fill[b] = seed + partial[threadIdx.x];
}

// pass 3: every walk re-emits its pairs in order and writes the ones that still fit
__global__ void __launch_bounds__(256) fill_scatter_kernel(const FillParams p, const uint2 *chains, uint32_t num_walk,
                                                           unsigned long long first_walk, uint32_t *counts,
                                                           uint32_t *const *pool_blocks,
                                                           unsigned long long *last_walk) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= num_walk)
        return;
    const int num_block = p.num_partition * p.num_partition;
    uint32_t *row = counts + size_t(w) * num_block;
    const unsigned long long shuffle_stride = p.pool_size / p.shuffle_base;
    bool completed = false;
    for (int j = 0; j < p.walk_length; j++) {
        const uint2 head = chains[size_t(j) * num_walk + w];
        for (int k = 1; k <= p.augmentation_step && j + k <= p.walk_length; k++) {
            const uint2 tail = chains[size_t(j + k) * num_walk + w];
            const int b = head.x * p.num_partition + tail.x;
            const unsigned long long in_slice = row[b];
            if (in_slice < p.slice) {
                row[b] = uint32_t(in_slice + 1);
                const unsigned long long offset = p.start + in_slice;
                // pseudo shuffle, instance/graph.cuh:440-441
                const unsigned long long shuffled = offset % p.shuffle_base * shuffle_stride + offset / p.shuffle_base;
                uint2 *block = reinterpret_cast<uint2 *>(pool_blocks[b]);
                if (block)
                    block[shuffled] = make_uint2(tail.y, head.y);  // {tail_local, head_local}
                completed |= in_slice + 1 == p.slice;
            }
        }
    }
    if (completed)
        atomicMax(last_walk, first_walk + w);
}

// single-block fast path (num_partition == 1): the slice offset of a pair is its stream index
__global__ void __launch_bounds__(256) fill_direct_kernel(const FillParams p, const uint2 *chains, uint32_t num_walk,
                                                          unsigned long long first_walk, uint32_t pairs_per_walk,
                                                          const unsigned long long *fill,
                                                          uint32_t *const *pool_blocks,
                                                          unsigned long long *last_walk) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= num_walk)
        return;
    uint2 *block = reinterpret_cast<uint2 *>(pool_blocks[0]);
    const unsigned long long shuffle_stride = p.pool_size / p.shuffle_base;
    unsigned long long in_slice = fill[0] + (unsigned long long)w * pairs_per_walk;
    bool completed = false;
    for (int j = 0; j < p.walk_length && in_slice < p.slice; j++) {
        const uint2 head = chains[size_t(j) * num_walk + w];
        for (int k = 1; k <= p.augmentation_step && j + k <= p.walk_length; k++, in_slice++) {
            if (in_slice >= p.slice)
                break;
            const uint2 tail = chains[size_t(j + k) * num_walk + w];
            const unsigned long long offset = p.start + in_slice;
            const unsigned long long shuffled = offset % p.shuffle_base * shuffle_stride + offset / p.shuffle_base;
            if (block)
                block[shuffled] = make_uint2(tail.y, head.y);
            completed |= in_slice + 1 == p.slice;
        }
    }
    if (completed)
        atomicMax(last_walk, first_walk + w);
}

__global__ void fill_advance_kernel(unsigned long long *fill, unsigned long long amount) {
    fill[0] += amount;
}

static uint32_t pairs_per_walk(int walk_length, int augmentation_step) {
    uint32_t n = 0;
    for (int j = 0; j < walk_length; j++)
        for (int k = 1; k <= augmentation_step && j + k <= walk_length; k++)
            n++;
    return n;
}

}  // namespace device
}  // namespace gv

using namespace gv;
using namespace gv::device;

extern "C" {

int gv_cuda_random_walk(const gv_device_graph_t *graph, const double *random, uint32_t num_walk, int walk_length,
                        gv_location_t *chains, void *stream) {
    if (num_walk == 0)
        return 0;
    if (!graph || !random || !chains || walk_length < 1)
        return fail("gv_cuda_random_walk: invalid argument");
    if (walk_length > 1 && !graph->vertex_tables)
        return fail("gv_cuda_random_walk: per-vertex alias tables are required for walk_length > 1");
    const int threads = 256;
    const uint32_t blocks = (num_walk + threads - 1) / threads;
    random_walk_kernel<<<blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(*graph, random, num_walk,
                                                                                  walk_length, chains);
    GV_CUDA_OK(cudaGetLastError());
    return 0;
}

size_t gv_cuda_fill_scratch_bytes(uint32_t num_walk, int num_partition) {
    if (num_partition <= 1)
        return 16;
    return size_t(num_walk) * num_partition * num_partition * sizeof(uint32_t);
}

int gv_cuda_fill_pool(const gv_fill_params_t *params, const gv_location_t *chains, uint32_t num_walk,
                      uint64_t first_walk, uint32_t *const *pool_blocks, unsigned long long *fill,
                      unsigned long long *last_walk, void *scratch, void *stream) {
    if (num_walk == 0)
        return 0;
    if (!params || !chains || !pool_blocks || !fill || !last_walk)
        return fail("gv_cuda_fill_pool: null argument");
    if (params->num_partition < 1 || params->num_partition > 16)
        return fail("gv_cuda_fill_pool: num_partition must be in [1, 16]");
    if (params->shuffle_base < 1 || params->pool_size % params->shuffle_base != 0)
        return fail("Can't perform pseudo shuffle: episode size must be a multiple of the shuffle base");
    if (params->end < params->start || params->end > params->pool_size)
        return fail("gv_cuda_fill_pool: invalid slice");
    FillParams p;
    p.num_partition = params->num_partition;
    p.walk_length = params->walk_length;
    p.augmentation_step = params->augmentation_step;
    p.shuffle_base = params->shuffle_base;
    p.pool_size = params->pool_size;
    p.start = params->start;
    p.slice = params->end - params->start;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const uint2 *c = reinterpret_cast<const uint2 *>(chains);
    const int threads = 256;
    const uint32_t blocks = (num_walk + threads - 1) / threads;
    if (p.num_partition == 1) {
        // every walk is full length, so pair (w, j, k) sits at stream index w * pairs_per_walk + f(j, k);
        // fill[0] is read on the device and advanced by a 1-thread kernel behind the fill (same stream)
        const uint32_t per_walk = pairs_per_walk(p.walk_length, p.augmentation_step);
        fill_direct_kernel<<<blocks, threads, 0, s>>>(p, c, num_walk, first_walk, per_walk, fill, pool_blocks,
                                                      last_walk);
        GV_CUDA_OK(cudaGetLastError());
        fill_advance_kernel<<<1, 1, 0, s>>>(fill, (unsigned long long)num_walk * per_walk);
        GV_CUDA_OK(cudaGetLastError());
        return 0;
    }
    if (!scratch)
        return fail("gv_cuda_fill_pool: scratch required for num_partition > 1");
    uint32_t *counts = static_cast<uint32_t *>(scratch);
    fill_count_kernel<<<blocks, threads, 0, s>>>(p, c, num_walk, counts);
    GV_CUDA_OK(cudaGetLastError());
    fill_scan_kernel<<<p.num_partition * p.num_partition, 1024, 0, s>>>(p, num_walk, counts, fill);
    GV_CUDA_OK(cudaGetLastError());
    fill_scatter_kernel<<<blocks, threads, 0, s>>>(p, c, num_walk, first_walk, counts, pool_blocks, last_walk);
    GV_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // extern "C"
